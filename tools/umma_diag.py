"""Hardware diagnostics of the tcgen05 layer (run on the B200 box): each case runs in its own subprocess so a
trap/sticky CUDA error in one case cannot poison the others.  Writes gpurun_out/umma_diag.json.

    python tools/umma_diag.py            # all cases
    python tools/umma_diag.py CASE ...   # internal: run one case in this process, print JSON
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def selftest(N, K, a_mn, b_mn, variant):
    import torch
    from fourierflow_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(N * 1000 + K + a_mn * 7 + b_mn * 13)
    A = torch.randn(128, K, generator=g).bfloat16()
    B = torch.randn(N, K, generator=g).bfloat16()
    ref = A.float() @ B.float().t()
    Ad, Bd = A.cuda(), B.cuda()
    D = torch.zeros(128, N, device="cuda")
    st = lib.ffno_umma_selftest(Ad.data_ptr(), Bd.data_ptr(), D.data_ptr(), N, K, a_mn, b_mn, variant,
                                C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    out = {"status": st}
    if st == 0:
        d = D.cpu()
        err = (d - ref).abs()
        out["max_err"] = err.max().item()
        out["ref_max"] = ref.abs().max().item()
        bad = (err > 1e-2 * ref.abs().max()).nonzero()
        out["n_bad"] = int(bad.shape[0])
        out["bad_rows"] = sorted(set(bad[:, 0].tolist()))[:16]
        out["bad_cols"] = sorted(set(bad[:, 1].tolist()))[:16]
        out["n_bad_rows"] = len(set(bad[:, 0].tolist()))
        out["n_bad_cols"] = len(set(bad[:, 1].tolist()))
        out["zero_frac"] = (d == 0).float().mean().item()
    return out


def ff_case(P, what="both"):
    import torch
    import fourierflow_b200.modules as M
    from golden_like import rel
    torch.manual_seed(0)
    m = M.FNOFactorized2DBlock(modes=16, width=64, n_layers=1, input_dim=3, share_weight=True, factor=4,
                               ff_weight_norm=True, gain=0.1).cuda().eval()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn_like(p))
    layer = m.spectral_layers[0]
    side = 64
    B = max(1, P // (side * side))
    s = torch.randn(B, side, side, 64, device="cuda")
    x = torch.randn(B, side, side, 64, device="cuda")
    res = {}
    with torch.no_grad():
        os.environ["FFNO_B200_PATH"] = "generic"
        pg = layer._plan(s)
        yg = pg.ff_forward(0, 0, s, x)
        bg = pg.ff_forward(0, 0, s, None)
        sg = pg.spectral_forward(0, x)
        os.environ["FFNO_B200_PATH"] = "umma"
        pu = layer._plan(s)
        assert pu.uses_umma
        if what in ("both", "ff"):
            yu = pu.ff_forward(0, 0, s, x)
            bu = pu.ff_forward(0, 0, s, None)
            torch.cuda.synchronize()
            res["ff_residual_rel"] = rel(yu, yg)
            res["ff_rel"] = rel(bu, bg)
        if what in ("both", "spectral"):
            su = pu.spectral_forward(0, x)
            torch.cuda.synchronize()
            res["spectral_rel"] = rel(su, sg)
    return res


def block_case():
    import torch
    import fourierflow_b200.modules as M
    from golden_like import rel
    torch.manual_seed(0)
    m = M.FNOFactorized2DBlock(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4,
                               ff_weight_norm=True, gain=0.1).cuda().eval()
    x = torch.randn(32, 64, 64, 3, device="cuda")
    res = {}
    with torch.no_grad():
        os.environ["FFNO_B200_PATH"] = "generic"
        yg = m(x)["forecast"]
        os.environ["FFNO_B200_PATH"] = "umma"
        yu = m(x)["forecast"]
        torch.cuda.synchronize()
        res["block24_rel"] = rel(yu, yg)
        import time
        for name in ("generic", "umma"):
            os.environ["FFNO_B200_PATH"] = name
            for _ in range(3):
                m(x)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(10):
                m(x)
            torch.cuda.synchronize()
            res[f"ms_{name}"] = (time.perf_counter() - t0) * 100
    return res


CASES = {}
for N, K in ((64, 64), (128, 128), (256, 256), (64, 256), (256, 64), (32, 64), (16, 16)):
    CASES[f"st_k_N{N}_K{K}"] = (selftest, (N, K, 0, 0, 0))
for v in (0, 1):
    CASES[f"st_amn_v{v}"] = (selftest, (64, 64, 1, 0, v))
    CASES[f"st_bmn_v{v}"] = (selftest, (64, 64, 0, 1, v))
    CASES[f"st_amn_bmn_N128_K128_v{v}"] = (selftest, (128, 128, 1, 1, v))
CASES["st_atmem_N64"] = (selftest, (64, 64, 0, 0, 2))
CASES["st_atmem_N128"] = (selftest, (128, 64, 0, 0, 2))
CASES["ff_4096"] = (ff_case, (4096,))
CASES["ff_131072"] = (ff_case, (131072,))
CASES["ffonly_131072"] = (ff_case, (131072, "ff"))
CASES["speconly_131072"] = (ff_case, (131072, "spectral"))
CASES["block24"] = (block_case, ())


def main():
    if len(sys.argv) == 2 and sys.argv[1] in CASES and os.environ.get("FFNO_DIAG_CHILD") == "1":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        fn, args = CASES[sys.argv[1]]
        print("RESULT " + json.dumps(fn(*args)))
        return
    names = sys.argv[1:] or list(CASES)
    out = {}
    for name in names:
        try:
            env = dict(os.environ, FFNO_DIAG_CHILD="1")
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], capture_output=True, text=True,
                               timeout=180, env=env)
            line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
            out[name] = json.loads(line[-1][7:]) if line else {"rc": r.returncode, "stderr": r.stderr[-600:]}
        except subprocess.TimeoutExpired:
            out[name] = {"timeout": True}
        print(name, out[name], flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "umma_diag.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
