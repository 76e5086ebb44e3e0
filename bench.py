"""Benchmark of the LSTM hot path: SimpleLSTM fp32 training step, B=64 per GPU x 300 frames
(BASELINE.json configs[1]); metric = training frames/sec (whole job, all ranks).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (sm_100a kernels)
    python bench.py --impl reference ...                            # the reference's CPU path, same config
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One JSON line on stdout (rank 0).  `value` is timed with inputs resident in HBM; `e2e` runs the same step
through the public module API from pinned HOST buffers (H2D copy of the batch and D2H read of the loss
inside the timed region).  `roofline` describes the dominant kernel (the persistent recurrent kernels),
timed live with CUDA events on their own stream through the library's measurement hooks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, T_FRAMES, ACOUSTIC, POSE, HIDDEN, LAYERS = 64, 300, 80, 6, 256, 2
WORKLOAD = ("simple_lstm fp32 train step (fwd+bwd+AdamW), B=64/GPU x T=300 frames, stereo 2x40-d log-mel "
            "+ 6-d head pose -> 6-d next-frame motion, hidden 256 x 2 layers per stack [BASELINE configs[1]]")
# algorithmic HBM bytes per frame per LSTM layer at I=H=256, fp32
BYTES_FWD_PER_FRAME, BYTES_BWD_PER_FRAME = 7168, 9216


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def synthetic_batch(seed: int, batch: int, pin: bool):
    g = torch.Generator().manual_seed(seed)
    acoustic = torch.randn(batch, T_FRAMES, ACOUSTIC, generator=g)
    motion = torch.randn(batch, T_FRAMES, POSE, generator=g)
    target = torch.randn(batch, 1, POSE, generator=g)
    if pin:
        acoustic, motion, target = acoustic.pin_memory(), motion.pin_memory(), target.pin_memory()
    return acoustic, motion, target


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
            out, _ = self.proc.communicate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def timed_region(fn, steps, world):
    """barrier + sync, K steps between CUDA events, sync + barrier; returns max-over-ranks milliseconds."""
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(steps):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    return float(ms.item())


def cpu_reference_steps(batch: int, steps: int, warmup: int, threads: int):
    """The reference's CPU path (oracle port over torch.nn.LSTM / oneDNN): fwd + bwd + AdamW per step."""
    from multimodalreactiongeneration_b200.mr_gen.configs import simple_lstm_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
    from oracle import ref_port
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    model = SimpleLSTM(*simple_lstm_cfg(HIDDEN, LAYERS, False, ACOUSTIC, POSE))
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items()}
    opt = torch.optim.AdamW(list(sd.values()), lr=5e-6, weight_decay=1e-2)
    batch_t = synthetic_batch(1234, batch, pin=False)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = ref_port.simple_lstm_training_step(sd, batch_t)
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, float(loss)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    warm = min(args.warmup, 2)
    # bounded sample: every step processes `rows` of the B_PER_GPU sequences of the config, chosen from a calibration step
    # so that the whole K + W run takes about two minutes of CPU time (the full batch costs ~8 s per step on 16 cores)
    calib, _ = cpu_reference_steps(8, 1, 1, threads)
    t_row = calib[0] / 8.0
    rows = int(120.0 / ((args.steps + warm) * t_row))
    rows = max(8, min(B_PER_GPU, rows))
    times, _ = cpu_reference_steps(rows, args.steps, warm, threads)
    total = sum(times)
    value = rows * T_FRAMES * len(times) / total
    line = {
        "impl": "reference", "metric": "train_frames_per_sec", "value": value, "unit": "frames/s",
        "n_gpus": args.gpus, "steps": len(times), "warmup": warm,
        "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "device": "host CPU (reference's torch.nn.LSTM / oneDNN path)"},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": threads, "kind": "port",
                         "sample": f"{len(times)} steps of {rows} of the {B_PER_GPU} sequences x T={T_FRAMES} each "
                                   f"(bounded to ~2 min of CPU work; {total:.1f} s timed)"},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    from multimodalreactiongeneration_b200 import _cabi
    from multimodalreactiongeneration_b200.mr_gen.configs import simple_lstm_cfg
    from multimodalreactiongeneration_b200.mr_gen.model.simple_lstm.simple_lstm import SimpleLSTM
    from multimodalreactiongeneration_b200.mr_gen.tainer.trainer import Trainer, init_distributed

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the LSTM path has no CPU fallback (use --impl reference)")
    world = init_distributed()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _cabi.lib()  # fail loudly before timing anything if the extension is missing

    torch.manual_seed(0)
    model = SimpleLSTM(*simple_lstm_cfg(HIDDEN, LAYERS, False, ACOUSTIC, POSE)).to(dev)
    trainer = Trainer(model)
    n_pool = 4
    host = [synthetic_batch(1234 + rank * 100 + i, B_PER_GPU, pin=True) for i in range(n_pool)]
    resident = [tuple(t.to(dev) for t in b) for b in host]
    staging = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
    h2d = sum(t.numel() * 4 for t in host[0])

    # kernels of this library per step, counted on an eager step (a graph replay launches the same
    # kernels without going through the library's host code)
    trainer.train_step(resident[0])
    l0 = _cabi.launch_count()
    trainer.train_step(resident[0])
    launches_per_step = _cabi.launch_count() - l0
    graphed = not args.no_graph
    graph_note = "whole step replayed from one captured CUDA graph"
    if graphed:
        try:
            trainer.enable_cuda_graph(resident[0])
        except Exception as exc:  # e.g. a collective that cannot be captured on this stack
            graphed = False
            graph_note = f"eager (graph capture failed: {type(exc).__name__})"
    else:
        graph_note = "eager (--no-graph)"
    run_step = trainer.train_step_graphed if graphed else trainer.train_step

    def step_resident(i):
        run_step(resident[i % n_pool])

    losses = []

    def step_e2e(i):
        # host (pinned) -> device copy of this step's batch, the step, device -> host read of its loss
        if graphed:
            loss = run_step(host[i % n_pool])
        else:
            dst = staging[i % 2]
            for d, s in zip(dst, host[i % n_pool]):
                d.copy_(s, non_blocking=True)
            loss = run_step(dst)
        losses.append(float(loss.item()))

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms = timed_region(step_resident, args.steps, world)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if sampler else None

    for i in range(2):
        step_e2e(i)
    ms_e2e = timed_region(step_e2e, args.steps, world)

    # live per-kernel timing of the recurrent kernels (CUDA events on their own stream)
    _cabi.profile_enable(True)
    prof_steps = 3
    for i in range(prof_steps):
        trainer.train_step(resident[i % n_pool])   # eager: the library records events around its launches
    torch.cuda.synchronize()
    prof = _cabi.profile_read()
    _cabi.profile_enable(False)

    # the same recurrent kernels timed ALONE on the GPU (one nn.LSTM layer, B x T x 256, all clusters): inside the step
    # the two encoder stacks share the SMs on two streams, which stretches every individual launch
    from multimodalreactiongeneration_b200 import lstm_layer
    kk = 1.0 / HIDDEN ** 0.5
    iso_w = [torch.empty(4 * HIDDEN, HIDDEN, device=dev).uniform_(-kk, kk).requires_grad_(True) for _ in range(2)] + \
            [torch.empty(4 * HIDDEN, device=dev).uniform_(-kk, kk).requires_grad_(True) for _ in range(2)]
    iso_x = torch.randn(T_FRAMES, B_PER_GPU, HIDDEN, device=dev, requires_grad=True)
    for it in range(7):
        if it == 2:
            torch.cuda.synchronize()
            _cabi.profile_enable(True)
        lstm_layer(iso_x, iso_w, HIDDEN, 1)[0].sum().backward()
    torch.cuda.synchronize()
    iso = _cabi.profile_read()
    _cabi.profile_enable(False)

    if graphed:
        trainer.release_cuda_graph()   # before the process group goes away (see Trainer.release_cuda_graph)
    if rank != 0:
        return
    frames = world * B_PER_GPU * T_FRAMES
    value = frames * args.steps / (ms * 1e-3)
    peak, peak_src = peaks()
    (fwd_ms, fwd_n), (bwd_ms, bwd_n), (gemm_ms, gemm_n) = prof["rec_fwd"], prof["rec_bwd"], prof["gemm"]
    frames_per_launch = B_PER_GPU * T_FRAMES
    dom = "rec_bwd" if bwd_ms >= fwd_ms else "rec_fwd"
    dom_ms, dom_n = (bwd_ms, bwd_n) if dom == "rec_bwd" else (fwd_ms, fwd_n)
    per_frame = BYTES_BWD_PER_FRAME if dom == "rec_bwd" else BYTES_FWD_PER_FRAME
    in_step_avg_ms = dom_ms / max(1, dom_n)
    # roofline of the dominant kernel from its launches timed ALONE (burst peak applies); inside the step four of the
    # six launches per direction share the GPU with the other encoder's kernel on a second stream
    avg_ms = iso[dom][0] / max(1, iso[dom][1])
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r1d_traffic.json")
    if os.path.exists(tpath):  # dram__bytes_read.sum + dram__bytes_write.sum of that kernel, one ncu --set full capture
        with open(tpath) as fh:
            t = json.load(fh).get(dom)
        if t:
            traffic = t["dram_bytes_read"] + t["dram_bytes_write"]
    achieved = frames_per_launch * per_frame / (avg_ms * 1e-3) / 1e9
    line = {
        "metric": "train_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": world * B_PER_GPU, "frames_per_step": frames,
                   "parallelism": f"dp{world} (batch sharded by sequence; flat gradient bucket, decoder / attention slices "
                                  f"all-reduced during the encoders' BPTT, the rest in one trailing all-reduce)",
                   "cuda_graph": graph_note,
                   "l2": "no explicit flush: each step rewrites ~0.6 GB of reserve/activations, >> 126 MB L2"},
        "clocks": clocks,
        "e2e": {"value": frames * args.steps / (ms_e2e * 1e-3), "unit": "frames/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "kernel": f"{dom}2_kernel<256, 4>", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "avg_launch_ms": avg_ms, "in_step_avg_launch_ms": in_step_avg_ms,
                     "algorithmic_bytes_per_launch": frames_per_launch * per_frame,
                     "note": "latency-bound recurrence: T dependent steps per launch; see latency_us_per_timestep"},
        "latency_us_per_timestep": {"rec_fwd": 1e3 * iso["rec_fwd"][0] / max(1, iso["rec_fwd"][1]) / T_FRAMES,
                                    "rec_bwd": 1e3 * iso["rec_bwd"][0] / max(1, iso["rec_bwd"][1]) / T_FRAMES,
                                    "how": "one LSTM layer (B=64, T=300, I=H=256) alone on the GPU, 5 launches each"},
        "latency_us_per_timestep_in_step": {"rec_fwd": 1e3 * fwd_ms / max(1, fwd_n) / T_FRAMES,
                                            "rec_bwd": 1e3 * bwd_ms / max(1, bwd_n) / T_FRAMES,
                                            "how": "average over the 12 launches of a step; the two encoder stacks "
                                                   "run side by side on two streams with half the clusters each"},
        "kernel_ms_per_step": {"rec_fwd": fwd_ms / prof_steps, "rec_bwd": bwd_ms / prof_steps,
                               "gemm": gemm_ms / prof_steps, "rec_launches": (fwd_n + bwd_n) // prof_steps,
                               "gemm_launches": gemm_n // prof_steps},
        "loss_last": losses[-1] if losses else None,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        t0 = time.perf_counter()
        times, _ = cpu_reference_steps(8, 8, 2, threads)  # BASELINE configs[0]: B=8 x 300 on CPU
        line["cpu_baseline"] = {
            "value": 8 * T_FRAMES * len(times) / sum(times), "unit": "frames/s", "cores": threads,
            "kind": "port",
            "sample": f"{len(times)} steps of the same model at B=8 x T={T_FRAMES} (BASELINE configs[0]), "
                      f"{time.perf_counter() - t0:.1f} s of CPU work"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="enqueue every kernel from Python instead of "
                    "replaying the captured CUDA graph of the step")
    args = ap.parse_args()
    args.warmup = max(3, args.warmup) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    _leave()


def _leave():
    """Multi-rank exit that cannot hang: every rank drains its device, meets the others at a barrier (so no peer is
    still inside a collective) and then leaves WITHOUT the NCCL communicator teardown — at 8 ranks
    ``destroy_process_group`` after CUDA-graph-captured collectives blocked for minutes after the result line had
    been printed.  Single-process runs return normally."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return
    sys.stdout.flush()
    sys.stderr.flush()
    try:
        if torch.cuda.is_available():
            torch.cuda.synchronize()
        dist.barrier()
        if torch.cuda.is_available():
            torch.cuda.synchronize()
    finally:
        os._exit(0)


if __name__ == "__main__":
    main()
