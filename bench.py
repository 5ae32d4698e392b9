#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native F-FNO forward (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Metric (BASELINE.json): Navier–Stokes 64x64 samples/s of the 24-layer F-FNO forward.
Workload at N=1: configs[1] = torus_li/markov/24_layers, 64x64 grid, batch 32, fp32, random-init weights of that
architecture, synthetic N(0,1) input features.  One "step" = one forward of the 24-layer stack over one batch of 32
samples per GPU (weak scaling: 32 samples per rank; N=8 is configs[2], batch 256 sharded over 8 GPUs).

  value : samples/s with the inputs already resident in HBM (CUDA events, L2 flushed between timed steps).
  e2e   : the same metric through the C-ABI host-buffer entry point (ffno_block_fwd_host): pinned host input
          copied H2D, forward, forecast copied D2H, all inside the timed region.
  roofline / cpu_baseline : see DESIGN.md §Measurement.

  parity : the timed kernels' forecast of the first 8 samples of rank 0 against the reference's own modules run on
          the host cores on the same input and weights (max|y-ref| / max|ref|); the line is refused above 1e-4.
  eager_gpu / vs_eager : the reference's own FNOFactorized2DBlock (oracle/_ref, unmodified) moved to the same GPU,
          TF32 off, eager PyTorch, timed in this run — the north star's ">= 10x reference eager" denominator.

  train_step : (informational, rank 0, after the timed region) forward + backward of the reference's one-step training
          loss through ffno_block_bwd, next to the reference modules' autograd on the same GPU, and the relative L2
          distance between the two whole gradients.

`--impl reference` times the reference's own CPU implementation (oracle/_ref: the unmodified fourierflow.modules
files, copied there by oracle/make_ref.sh; the torch-CPU oracle port only if that copy is missing) on the host cores,
same metric and config, all 32 samples per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE JSON line.  Libraries write banners there (NCCL prints "NCCL version ..." to stdout at
# NCCL_DEBUG=VERSION whatever NCCL_DEBUG_FILE says), so file descriptor 1 is pointed at stderr for the whole run and
# the result line goes to a private duplicate of the original stdout.
_RESULT_OUT = os.fdopen(os.dup(1), "w")
sys.stdout.flush()
os.dup2(2, 1)


def emit(line: dict) -> None:
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()

C2 = dict(modes=16, width=64, n_layers=24, input_dim=3, share_weight=True, factor=4, ff_weight_norm=True, gain=0.1)
GRID = 64
BATCH_PER_GPU = 32
WORKLOAD = "torus_li/markov/24_layers FNOFactorized2DBlock fwd, 64x64, batch 32/GPU, fp32 (BASELINE configs[1]; N=8 -> configs[2])"
METRIC = "navier_stokes_64x64_samples_per_sec_24layer_ffno_fwd"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# --------------------------------------------------------------------------------------------------
# clocks sampling during the timed region (B200_PROFILING.md "clocks line")
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle port) on the host cores
# --------------------------------------------------------------------------------------------------
def reference_model(state_dict=None):
    """The reference's own FNOFactorized2DBlock (oracle/_ref, see oracle/make_ref.sh) with the C2 arguments and the
    seed-0 initialisation; None when the copy is not there.  Only bench legs that TIME or CHECK use it."""
    ref_root = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.exists(os.path.join(ref_root, "fourierflow", "modules", "factorized_fno", "grid_2d.py")):
        return None
    if ref_root not in sys.path:
        sys.path.insert(0, ref_root)
    import warnings
    warnings.filterwarnings("ignore")
    from fourierflow.modules.factorized_fno.grid_2d import FNOFactorized2DBlock as RefBlock
    torch.manual_seed(0)
    m = RefBlock(**C2).eval()
    if state_dict is not None:
        m.load_state_dict(state_dict, strict=True)
    return m


def oracle_setup(sample_batch: int, seed: int = 1):
    """(step, kind): one forward of `sample_batch` samples of the seed-`seed` input on the host cores through the
    reference modules ("reference") or, without oracle/_ref, the oracle port ("port")."""
    from fourierflow_b200.modules import FNOFactorized2DBlock
    torch.manual_seed(0)
    m = FNOFactorized2DBlock(**C2).eval()
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    x = torch.randn(BATCH_PER_GPU, GRID, GRID, 3, generator=torch.Generator().manual_seed(seed))[:sample_batch]
    ref = reference_model(sd)
    if ref is not None:
        def step():
            with torch.no_grad():
                return ref(x)["forecast"]
        return step, "reference"
    from oracle import ffno_oracle as O          # the one other place bench.py touches oracle/: the CPU baseline

    def step():
        with torch.no_grad():
            return O.block_grid2d_forward(sd, x, modes=C2["modes"], n_layers=C2["n_layers"])["forecast"]
    return step, "port"


def pick_cpu_threads(step):
    """All the host threads the box can USE: the GPU hosts expose 128 logical CPUs but a container quota makes
    torch with 128 threads far slower than with fewer, so time one step at a few thread counts and keep the best."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, avail) if c <= avail} | {min(avail, 8)})
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        step()
        t0 = time.perf_counter()
        step()
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = c, dt
        if dt > 4 * best_t:
            break
    torch.set_num_threads(best)
    return best


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = BATCH_PER_GPU
    step, kind = oracle_setup(sample)
    cores = pick_cpu_threads(step)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = sample / dt
    what = ("the reference's own FNOFactorized2DBlock (oracle/_ref, unmodified fourierflow.modules files)" if kind == "reference"
            else "torch CPU oracle port of fourierflow.modules (oracle/_ref missing)")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": sample, "grid": [GRID, GRID], "n_layers": 24},
        "cpu_baseline": {"value": v, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"all {sample} samples of one step per forward of the 24-layer stack, {what}"},
        "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def ev_pair():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, steps, warmup, flush):
    """Per-step CUDA-event timing on the current stream; the L2 flush sits outside the events."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(steps):
        flush()
        a, b = ev_pair()
        a.record()
        fn()
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
    return tot / steps          # ms per step


def run_ours(args, rank, world, local):
    import torch.distributed as dist
    from fourierflow_b200 import build, distributed as D
    build.build()
    from fourierflow_b200.modules import FNOFactorized2DBlock, LpLoss

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    torch.manual_seed(0)
    model = FNOFactorized2DBlock(**C2).to(dev).eval()
    B = BATCH_PER_GPU
    gen = torch.Generator().manual_seed(1 + rank)
    x_host = torch.randn(B, GRID, GRID, 3, generator=gen).pin_memory()
    out_host = torch.empty(B, GRID, GRID, 1).pin_memory()
    x = x_host.to(dev)
    target = torch.randn(B, GRID, GRID, 1, device=dev)
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    plan = model.plan_for(dev, (GRID, GRID))
    loss_fn = LpLoss()
    # the one collective of a sharded run: per-sample losses for the LpLoss mean, gathered on a side stream into
    # preallocated buffers so that the (latency-bound) all-gather overlaps the next step's kernels
    gatherer = D.LossGatherer(1, B, B * world, dev) if world > 1 else None

    def flush():
        flush_buf.zero_()

    def step():
        with torch.no_grad():
            y = model(x)["forecast"]
            if gatherer is not None:
                gatherer.submit(loss_fn.rel_per_sample(y, target).unsqueeze(0))
        return y

    def step_e2e():
        plan.block_forward_host(x_host, out_host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure():
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        ms = timed(step, args.steps, args.warmup, flush)
        if gatherer is not None:
            gatherer.result()                 # every gather of the timed steps has completed
        launches = plan.last_launch_count
        ms_e2e = timed(step_e2e, args.steps, max(3, args.warmup), flush)
        barrier()
        clocks = sampler.stop()
        return ms, ms_e2e, launches, clocks

    ms, ms_e2e, launches, clocks = measure()
    if set(clocks.get("reasons", [])) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}:
        ms, ms_e2e, launches, clocks = measure()          # re-measure once

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()

    # ---- roofline of the two per-layer operators, timed alone with CUDA events on the launching stream --------
    # spectral operator = exactly the launches the layer loop issues (forward transforms of both axes, mode mixes,
    # inverse transforms: 3 kernels on the tcgen05 path); FF = the single ff_ts_kernel launch.  Cold L2 (flushed).
    peaks, peak_src = measured_peaks()
    layer = model.spectral_layers[0]
    xs = torch.randn(B, GRID, GRID, 64, device=dev)
    with torch.no_grad():
        lplan = layer._plan(xs)
        bufs = [torch.empty_like(xs), torch.empty_like(xs)]
        t_spec = timed(lambda: lplan.spectral_split_forward(0, xs, bufs), 20, 5, flush)
        n_spec = lplan.last_launch_count
        s = lplan.spectral_forward(0, xs)
        t_ff = timed(lambda: lplan.ff_forward(0, 0, s, xs), 20, 5, flush)
    P = B * GRID * GRID
    Cw, K = 64, C2["modes"]
    U = P * Cw * 4
    spec_bytes = 2 * U + 2 * Cw * Cw * K * 8                      # 2U + sum_a C*C*K_a*8
    ff_flops = 4 * C2["factor"] * Cw * Cw * P                      # 4 f C^2 P
    spec_gbs = spec_bytes / (t_spec * 1e-3) / 1e9
    ff_tflops = ff_flops / (t_ff * 1e-3) / 1e12
    traffic = {}
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp))
    roof_spec = {"kernel": f"spectral operator forward_fourier ({n_spec} launches: axis_pipe_kernel fwd, mix_pipe_kernel, "
                           "axis_pipe_kernel inv)", "bound": "hbm", "achieved": spec_gbs,
                 "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": spec_gbs / peaks["hbm_gbs"],
                 "traffic": traffic.get("spectral"), "ms": t_spec, "algorithmic_bytes": spec_bytes, "peak_source": peak_src}
    roof_ff = {"kernel": "ff_ts_kernel (feed-forward C->4C->C + residual, 3xBF16 tcgen05)", "bound": "tensor",
               "achieved": ff_tflops, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": ff_tflops / peaks["bf16_tflops"],
               "traffic": traffic.get("ff"), "ms": t_ff, "algorithmic_flops": ff_flops, "peak_source": peak_src,
               "note": "fp32-equivalent flops; every product costs 3 BF16 tensor passes, so frac <= 1/3 by construction"}
    # the spectral operator is the kernel family the north star names and holds the larger share of the step
    dominant = roof_spec if 1.0 * t_spec >= t_ff else roof_ff

    # ---- the reference's own module, eager PyTorch on this GPU (rank 0; TF32 off: the reference's default) ------
    eager = None
    y_gpu_first8 = None
    if rank == 0:
        with torch.no_grad():
            y_gpu_first8 = model(x)["forecast"][:8].float().cpu()
        ref = None if args.no_eager_gpu else reference_model({k: v.detach().cpu() for k, v in model.state_dict().items()})
        if ref is not None:
            torch.backends.cuda.matmul.allow_tf32 = False
            torch.backends.cudnn.allow_tf32 = False
            ref = ref.to(dev)
            eager_steps = max(10, min(args.steps, 100))
            sampler = ClockSampler(local)
            sampler.start()
            with torch.no_grad():
                ms_eager = timed(lambda: ref(x)["forecast"], eager_steps, 10, flush)
                y_eager = ref(x)["forecast"][:8].float().cpu()
            eager_clocks = sampler.stop()
            scale = y_eager.abs().max().clamp(min=1e-30)
            eager = {"value": B / (ms_eager * 1e-3), "unit": "samples/s", "ms_per_step": ms_eager, "steps": eager_steps,
                     "warmup": 10, "kind": "reference", "allow_tf32": False, "clocks": eager_clocks,
                     "what": "oracle/_ref FNOFactorized2DBlock(**C2).cuda(), eager PyTorch, same input and weights, L2 "
                             "flushed between timed forwards, CUDA events",
                     "max_rel_diff_vs_ours": float((y_gpu_first8 - y_eager).abs().max() / scale)}

    # ---- training step (backward row): forward + backward of the reference's one-step loss
    #      (routines/grid_2d_markov.py:172-193) through ffno_block_bwd, next to the reference's autograd on this GPU ----
    train = None
    if rank == 0 and not args.no_train_step:
        from fourierflow_b200.modules import LpLoss
        l2 = LpLoss(size_average=True)
        yt = torch.randn(B, GRID, GRID, 1, device=dev, generator=torch.Generator(dev).manual_seed(7))

        def ours_train():
            model.zero_grad(set_to_none=True)
            loss = l2(model(x)["forecast"].reshape(B, -1), yt.reshape(B, -1))
            loss.backward()
            return loss

        ms_train = timed(ours_train, 10, 3, flush)
        plan.set_backward_mode("fast")
        ms_train_fast = timed(ours_train, 10, 3, flush)
        plan.set_backward_mode("default")
        train = {"ms_per_step": ms_train, "samples_per_s": B / (ms_train * 1e-3), "steps": 10, "warmup": 3,
                 "fast_mode_ms_per_step": ms_train_fast,
                 "what": "forward (tcgen05 kernels) + backward (ffno_block_bwd) of LpLoss(forecast, target), batch "
                         f"{B}, no optimizer step; default mode = FP32 forward recompute + tcgen05 spectral adjoint, fast "
                         "mode = recompute on the tcgen05 kernels too (include/ffno_b200.h: ffno_plan_set_backward_mode); "
                         "gradients checked against the reference's in tests/test_gpu_backward.py"}
        if ref is not None:
            def ref_train():
                ref.zero_grad(set_to_none=True)
                f = ref(x)["forecast"]
                loss = (torch.linalg.vector_norm((f - yt).reshape(B, -1), dim=1)
                        / torch.linalg.vector_norm(yt.reshape(B, -1), dim=1)).mean()
                loss.backward()
                return loss

            ms_ref_train = timed(ref_train, 10, 3, flush)
            train["reference_eager_ms_per_step"] = ms_ref_train
            train["vs_eager"] = ms_ref_train / ms_train
            # whole-gradient agreement with the reference's autograd on identical weights and data
            ours_train()
            ref_train()
            rp = dict(ref.named_parameters())
            num = den = 0.0
            for k, prm in model.named_parameters():
                d = prm.grad.double() - rp[k].grad.double()
                num += float((d * d).sum())
                den += float((rp[k].grad.double() ** 2).sum())
            train["gradient_rel_l2_vs_reference"] = (num / max(den, 1e-300)) ** 0.5
        model.zero_grad(set_to_none=True)
    ref = None

    if world > 1:
        dist.barrier()                    # the other ranks leave here; rank 0's CPU leg below holds no GPU busy
        dist.destroy_process_group()
    if rank != 0:
        return

    # ---- CPU baseline + parity gate: the reference modules on the host cores, bounded sample -------------------
    cpu_sample, cpu_reps = 8, 3
    parity = None
    if args.no_cpu_baseline:
        cores, cpu_v, cpu_kind = 0, None, None
    else:
        cstep, cpu_kind = oracle_setup(cpu_sample, seed=1)        # rank 0's input: seed 1 + rank
        cores = pick_cpu_threads(cstep)
        t0 = time.perf_counter()
        for _ in range(cpu_reps):
            y_ref = cstep()
        cpu_v = cpu_sample * cpu_reps / (time.perf_counter() - t0)
        err = float((y_gpu_first8.double() - y_ref.double()).abs().max() / y_ref.double().abs().max())
        parity = {"max_rel_err": err, "n": cpu_sample, "tolerance": 1e-4, "against": cpu_kind,
                  "what": "forecast of the first 8 samples of the timed batch: CUDA path vs the reference modules on the "
                          "host cores, max|y-ref|/max|ref|"}
        if not err < 1e-4:
            raise SystemExit(f"bench.py: parity gate failed — max|y-ref|/max|ref| = {err:.3e} over the first {cpu_sample} "
                             "samples exceeds 1e-4; no value is reported for kernels that compute the wrong answer")

    N = world
    value = N * B / (ms * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": N, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": N * B, "grid": [GRID, GRID], "n_layers": 24,
                   "parallelism": f"dp{N} (batch shards, no data-path collective; loss all-gather only)",
                   "l2": "flushed between timed steps (256 MiB memset outside the timed events)",
                   "kernel_path": "tcgen05" if plan.uses_umma else "generic-fp32",
                   "stage_pipelined": plan.pipeline_unit(B) > 0, "cuda_graph": plan.graph_active},
        "clocks": clocks,
        "e2e": {"value": N * B / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": x_host.numel() * 4, "d2h_bytes_per_step": out_host.numel() * 4,
                "api": "ffno_block_fwd_host (pinned host buffers)"},
        # this repo's kernels launched inside the timed region, all ranks (the sharded step adds one rel_l2 launch)
        "gpu_launches": (int(launches) + (1 if N > 1 else 0)) * args.steps * N,
        "parity": parity,
        "eager_gpu": eager, "vs_eager": (value / N / eager["value"]) if eager else None,
        "train_step": train,
        "roofline": dominant, "roofline_spectral": roof_spec, "roofline_ff": roof_ff,
        "cpu_baseline": {"value": cpu_v, "unit": "samples/s", "cores": cores, "kind": cpu_kind,
                         "sample": f"{cpu_reps} forwards of {cpu_sample} samples through the 24-layer stack "
                                   "(the reference modules of oracle/_ref on the host cores; 'port' = oracle restatement)"},
    }
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 100; reference arm: 5)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg and the parity gate (profiling runs under ncu)")
    ap.add_argument("--no-eager-gpu", action="store_true", help="skip the reference-eager-on-GPU leg (profiling runs under ncu)")
    ap.add_argument("--no-train-step", action="store_true", help="skip the forward + backward leg (rank 0, after the timed region)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.steps is None:
        args.steps = 100 if args.impl == "ours" else 5

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the F-FNO backend has no CPU path (use --impl reference)")
    if world > 1:
        from fourierflow_b200 import distributed as D
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries exactly one JSON line
        D.init_from_env("nccl")
    run_ours(args, rank, world, local)          # (destroys the process group itself, before rank 0's CPU leg)


if __name__ == "__main__":
    main()
