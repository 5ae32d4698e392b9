"""``python -m fourierflow_b200.predict <config.yaml> [key=value ...]`` — the reference's own inference-time measurement
(fourierflow/commands/predict.py:87-105, also the tail of commands/train.py:134-148) on the B200 backend:

    routine = instantiate(config.routine); routine.load_lightning_model_state(ckpt)        # optional here
    batch = builder.inference_data()            # first 512 samples, [512, X, Y, T]
    routine.cuda(); routine.warmup(); time routine.infer(batch)
    inference_time = elapsed / n_samples / (step_size * n_steps)     # seconds per sample and simulated time unit

Differences, all stated in the printed JSON: the dataset is not available offline, so the batch is synthetic N(0, 1)
frames of the same shape (``--samples``, ``--grid``, ``--frames``) unless ``--mat`` names a .mat file with the reference's
``'u'`` array (builders/ns_markov.py:57-59); the normaliser statistics come from that batch when the checkpoint has none;
the timed region is synchronised and repeated (the reference times one unsynchronised shot).
"""
from __future__ import annotations

import argparse
import json
import sys
import time

import torch


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m fourierflow_b200.predict", description=__doc__.split("\n\n")[0])
    ap.add_argument("config", help="experiment YAML of the reference (e.g. experiments/torus_li/markov/24_layers/config.yaml)")
    ap.add_argument("overrides", nargs="*", help="hydra-style key=value overrides (routine.n_steps=10 ...)")
    ap.add_argument("--checkpoint", help="Lightning checkpoint written by the reference")
    ap.add_argument("--mat", help=".mat file holding the reference's 'u' array [N, X, Y, T]")
    ap.add_argument("--samples", type=int, default=512)
    ap.add_argument("--grid", type=int, default=64)
    ap.add_argument("--frames", type=int, default=20, help="T of the synthetic batch (rollout uses the last n_steps + 1)")
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args(argv)

    from .config import load_routine
    routine, _ = load_routine(args.config, args.overrides)
    if not hasattr(routine, "conv"):
        raise SystemExit("predict: this command times Grid2DMarkovExperiment-style routines (routine.conv)")
    if args.checkpoint:
        routine.load_lightning_model_state(args.checkpoint)
    if not torch.cuda.is_available():
        raise SystemExit("predict: no CUDA device — the F-FNO backend has no CPU path")

    if args.mat:
        import scipy.io
        data = torch.from_numpy(scipy.io.loadmat(args.mat)["u"].astype("float32")[:args.samples])
        source = args.mat
    else:
        data = torch.randn(args.samples, args.grid, args.grid, args.frames, generator=torch.Generator().manual_seed(1))
        source = "synthetic N(0,1) frames"
    routine = routine.cuda().eval()
    batch = {"data": data.cuda()}
    if float(routine.normalizer.count) == 0:             # no statistics in the (absent) checkpoint: std would be 1e-8
        routine.accumulate_statistics(batch["data"])
    T = batch["data"].shape[-1]
    n_steps = routine.n_steps or (T - 1)
    if T < n_steps + 1:
        raise SystemExit(f"predict: the batch has {T} frames, a {n_steps}-step rollout needs {n_steps + 1}")
    routine.warmup()
    routine.infer(batch)                                  # plan creation, parameter packing, graph capture
    routine.infer(batch)
    torch.cuda.synchronize()
    times = []
    for _ in range(max(1, args.repeats)):
        t0 = time.time()
        routine.infer(batch)
        torch.cuda.synchronize()
        times.append(time.time() - t0)
    elapsed = min(times)
    n = len(batch["data"])
    loss = float(routine.infer(batch)[0])                 # sum over the steps of the batch-mean relative L2 (:313-315)
    print(json.dumps({
        "rollout_loss": loss,
        "inference_time": elapsed / n / (routine.step_size * n_steps),
        "unit": "s per sample and simulated time unit (commands/predict.py:103-104)",
        "elapsed_s": elapsed, "samples": n, "n_steps": n_steps, "step_size": routine.step_size,
        "grid": list(batch["data"].shape[1:3]), "data": source, "checkpoint": args.checkpoint,
        "sample_steps_per_s": n * n_steps / elapsed,
    }))
    return 0


if __name__ == "__main__":
    sys.exit(main())
