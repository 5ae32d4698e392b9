"""Load tests/golden/*.npz fixtures (made by oracle/make_golden.py from the executed reference)."""
import json
import os
import re

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    kwargs = json.loads(bytes(z["kwargs_json"]).decode())
    sd, arrays = {}, {}
    for k in z.files:
        if k == "kwargs_json":
            continue
        if k.startswith("sd::"):
            sd[k[4:]] = torch.from_numpy(z[k])
        else:
            arrays[k] = torch.from_numpy(z[k])
    # re-create the aliases of a shared ParameterList (dropped by make_golden.sd_np)
    if "fourier_weight.0" in sd:
        layers = {int(m.group(1)) for k in sd for m in [re.match(r"spectral_layers\.(\d+)\.", k)] if m}
        n_axes = sum(1 for k in sd if re.match(r"fourier_weight\.\d+$", k))
        for l in layers:
            for a in range(n_axes):
                sd[f"spectral_layers.{l}.fourier_weight.{a}"] = sd[f"fourier_weight.{a}"]
    return kwargs, sd, arrays


def rel_err(y, ref):
    """max|y-ref| / max|ref| — the parity metric of SURVEY.md §8(c)."""
    y, ref = y.detach().double().cpu(), ref.detach().double().cpu()
    return ((y - ref).abs().max() / ref.abs().max().clamp(min=1e-30)).item()


def load_geo(name):
    """geo_* fixtures (FNOFactorizedPointCloud2D): complex weights stored as view_as_real, shared-weight aliases
    ``convs.{i}.fourier_weight.{a}`` re-created."""
    kw, sd, arrays = load(name)
    for k in list(sd):
        if re.match(r"convs\.\d+\.weights[12]$", k):
            sd[k] = torch.view_as_complex(sd[k].contiguous())
    if "fourier_weight.0" in sd:
        for i in range(1, kw["n_layers"]):
            for a in range(2):
                sd[f"convs.{i}.fourier_weight.{a}"] = sd[f"fourier_weight.{a}"]
    return kw, sd, arrays
