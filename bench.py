#!/usr/bin/env python
"""bench.py -- points/s of the PIXEL Helmholtz training step (fwd -> triple backward).

    python bench.py --gpus N --steps K --warmup W            # our CUDA path, rank 0 prints one JSON line
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port)

Workload (BASELINE.json config 5, the one the metric is quoted on): 2D cosine multicell,
cells [4,16,256,256] fp32 replicated on every rank, 2^25 collocation points in total,
partitioned over the N ranks, processed in chunks of 2^20 points (each chunk is exactly
BASELINE.json config 3), MLP head Linear(16,16)-Tanh-Linear(16,1) on the GPU, residual
f = u_xx + u_yy + k^2 u, loss = mean f^2, gradients w.r.t. cells and head; one NCCL
all-reduce of the gradients per step when N > 1.  `--workload cfg3|cfg4` select the other
single-GPU configurations.

A "step" is one pass over all points.  `value` = points of the whole job / max-over-ranks
device time.  Each [N,C,P] stream of a chunk is 256 MiB (> the 126 MB L2), so no L2 flush is
needed between iterations (stated in `config`).
"""
import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_OUT = sys.stdout

WORKLOADS = {
    # name: (dim, cells shape, total points, chunk, kernel name, residual)
    "cfg5": (2, (4, 16, 256, 256), 2 ** 25, 2 ** 20, "cosine", "helmholtz"),
    "cfg3": (2, (4, 16, 256, 256), 2 ** 20, 2 ** 20, "cosine", "helmholtz"),
    "cfg4": (3, (4, 16, 64, 64, 64), 2 ** 22, 2 ** 20, "smooth-step", "laplace"),
}
# minimal algorithmic bytes per (cell, point) pair of one step, BASELINE.md section 3
CHAIN_BYTES_PER_PAIR = {"cfg5": 1544, "cfg3": 1544, "cfg4": 2432}
CHAIN_FIELD_TERMS = {"cfg5": 19, "cfg3": 19, "cfg4": 27}     # grid-shaped tensors per step


def ncu_traffic_bytes(label):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a stage kernel, from the committed
    `ncu --set full` capture of the same shapes (profiles/r1/ncu_stages_cfg{3,4}_summary.json; the
    fused kernels: ncu_fused_cfg{3,4}_summary.json, captured at 2^20 / 2^22 points per launch)."""
    for name in ("ncu_stages_cfg3_summary.json", "ncu_stages_cfg4_summary.json",
                 "ncu_fused_cfg3_summary.json", "ncu_fused_cfg4_summary.json"):
        try:
            with open(os.path.join(ROOT, "profiles", "r1", name)) as f:
                kernels = json.load(f)["kernels"]
            if label in kernels:
                return int(kernels[label]["traffic_bytes"])
        except Exception:
            pass
    return None


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


class StageProfiler:
    """Collects CUDA-event timings of every stage call issued inside the timed region."""

    def __init__(self):
        self.records = []

    def record(self, label, nbytes, start, end):
        self.records.append((label, nbytes, start, end))

    def summary(self):
        agg = {}
        for label, nbytes, s, e in self.records:
            a = agg.setdefault(label, [0, 0.0, 0])
            a[0] += 1
            a[1] += s.elapsed_time(e)
            a[2] += nbytes
        return {k: {"launches": v[0], "ms_total": v[1], "ms_avg": v[1] / v[0],
                    "bytes_per_launch": v[2] // v[0]} for k, v in agg.items()}


# ---------------------------------------------------------------------------------------
def make_inputs(workload, rank, world, device, dtype):
    import torch
    dim, shape, total, chunk, kernel, residual = WORKLOADS[workload]
    from cosinesampler_b200 import chain, dp
    g = torch.Generator().manual_seed(0)
    cells = torch.rand(shape, generator=g, dtype=dtype)                 # U(0,1), replicated
    s, e = dp.shard_range(total, rank, world)
    gr = torch.Generator().manual_seed(1000 + rank)
    coords_host = torch.rand(e - s, dim, generator=gr, dtype=dtype) * 2 - 1     # U(-1,1), unsorted
    head = chain.make_head(shape[1], seed=0, device=device, dtype=dtype)
    return cells, coords_host, head


def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from cosinesampler_b200 import _lib, chain, dp, ops
    from cosine_sampler_2d import CosineSampler2d
    from cosine_sampler_3d import CosineSampler3d

    dim, shape, total, chunk, kernel, residual = WORKLOADS[args.workload]
    if args.points:
        total = args.points
    N, C = shape[:2]
    S = CosineSampler2d if dim == 2 else CosineSampler3d
    sampler = lambda c, g: S.apply(c, g, "zeros", True, kernel, True)
    cells_h, coords_host, head = make_inputs(args.workload, rank, world, device, torch.float32)
    if args.points:
        s, e = dp.shard_range(total, rank, world)
        coords_host = coords_host[: e - s]
    cells = torch.nn.Parameter(cells_h.to(device))
    coords_pinned = coords_host.pin_memory()
    coords_dev = coords_host.to(device)
    stepper = dp.PointShardedStep(sampler, cells, head, residual=residual, chunk=chunk)
    from cosinesampler_b200 import jet
    fused_kw = dict(kernel=kernel, multicell=True)
    # the fused step keeps [1+2*dim, C, chunk] jets instead of [N, C, chunk] streams: 4x the chunk of
    # the drop-in arm is the same footprint per stream
    fchunk = 4 * chunk
    reduce_kind = "none (single GPU)"
    fstepper = None
    if world > 1 and not args.nccl_reduce:
        try:        # gradient reduce fused with the layout change over NVLink peer memory (peer.py)
            fstepper = dp.PointShardedStep(None, cells, head, residual=residual, chunk=fchunk, fused=fused_kw,
                                           peer_reduce=True)
            reduce_kind = "peer-memory kernel cs_peer_allreduce_from_channel_last (symmetric memory over NVLink)"
        except Exception as exc:          # no symmetric memory on this box: NCCL
            print("peer-memory reduce unavailable (%s); using NCCL" % exc, file=sys.stderr)
            fstepper = None
        # all ranks must take the same path: one vote, NCCL unless every rank has its reducer
        vote = torch.tensor([1.0 if fstepper is not None else 0.0], device=device)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN)
        if float(vote.item()) < 0.5:
            fstepper = None
    if fstepper is None:
        fstepper = dp.PointShardedStep(None, cells, head, residual=residual, chunk=fchunk, fused=fused_kw)
        if world > 1:
            reduce_kind = "NCCL all-reduce of one flat bucket"

    def fused_resident():
        fstepper.zero_grad()
        return fstepper.step(coords_dev, total)

    def step_resident():
        stepper.zero_grad()
        cols = [coords_dev[:, a:a + 1] for a in range(dim)]
        return stepper.step(cols, total)

    loss_host = torch.zeros(1, pin_memory=True)

    copy_stream = torch.cuda.Stream(device=device)

    def step_e2e():
        """Same step from HOST buffers: every chunk of coordinates is copied from pinned host
        memory inside the timed region (on a copy stream, one chunk ahead of the compute stream) and
        the step's loss is read back to the host."""
        stepper.zero_grad()
        nloc = coords_pinned.shape[0]
        spans = [(s0, min(nloc, s0 + chunk)) for s0 in range(0, nloc, chunk)]
        cur = torch.cuda.current_stream()

        def fetch(span):
            with torch.cuda.stream(copy_stream):
                t = coords_pinned[span[0]:span[1]].to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return t, ev

        nxt = fetch(spans[0])
        acc = None
        for i, span in enumerate(spans):
            t, ev = nxt
            if i + 1 < len(spans):
                nxt = fetch(spans[i + 1])
            cur.wait_event(ev)
            t.record_stream(cur)
            loss = chain.training_step(sampler, cells, [t[:, a:a + 1] for a in range(dim)], head,
                                       residual=residual, loss_scale=(span[1] - span[0]) / float(total))
            acc = loss if acc is None else acc + loss
        dp.allreduce_grads(stepper.params())
        loss_host.copy_(acc.reshape(1), non_blocking=True)               # D2H of the step's result
        cur.synchronize()
        return loss_host

    def fused_e2e():
        """The fused step from HOST buffers, same copy pipeline as step_e2e."""
        fstepper.zero_grad()
        nloc = coords_pinned.shape[0]
        spans = [(s0, min(nloc, s0 + fchunk)) for s0 in range(0, nloc, fchunk)]
        cur = torch.cuda.current_stream()

        def fetch(span):
            with torch.cuda.stream(copy_stream):
                t = coords_pinned[span[0]:span[1]].to(device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            return t, ev

        fs = jet.FusedPdeStep(cells, head, residual, kernel=kernel, multicell=True)
        fs.begin(fstepper.reducer)
        nxt = fetch(spans[0])
        for i, span in enumerate(spans):
            t, ev = nxt
            if i + 1 < len(spans):
                nxt = fetch(spans[i + 1])
            cur.wait_event(ev)
            t.record_stream(cur)
            fs.add(t, 1.0 / float(total))
        loss = fs.finish()
        if fstepper.reducer is None:
            dp.allreduce_grads(fstepper.params())
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        cur.synchronize()
        return loss_host

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            fn()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident()
    barrier()

    clocks = ClockSampler(local)
    prof = StageProfiler()
    if rank == 0:
        clocks.start()
    ops.profiler = prof
    n0 = _lib.launch_count()
    ms = timed(step_resident, args.steps)
    launches = _lib.launch_count() - n0
    ops.profiler = None
    clock_info = clocks.stop() if rank == 0 else None
    loss_t = step_resident().detach().clone()
    if world > 1:
        dist.all_reduce(loss_t)                      # each rank holds its share of the mean
    loss_val = float(loss_t.item())

    for _ in range(max(1, min(args.warmup, 2))):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- the opt-in fused jet path (jet.FusedPdeStep), same inputs, same outputs
    for _ in range(args.warmup):
        fused_resident()
    barrier()
    fprof = StageProfiler()
    ops.profiler = fprof
    n0 = _lib.launch_count()
    ms_f = timed(fused_resident, args.steps)
    flaunches = _lib.launch_count() - n0
    ops.profiler = None
    floss_t = fused_resident().detach().clone()
    if world > 1 and not fstepper.loss_is_global:
        dist.all_reduce(floss_t)
    floss_val = float(floss_t.item())
    for _ in range(max(1, min(args.warmup, 2))):
        fused_e2e()
    ms_f_e2e = timed(fused_e2e, args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    pts_per_s = total * args.steps / (ms * 1e-3)
    e2e_pts_per_s = total * args.steps / (ms_e2e * 1e-3)
    stage = prof.summary()
    top = max(stage, key=lambda k: stage[k]["ms_total"]) if stage else None
    roofline = None
    if top:
        a = stage[top]["bytes_per_launch"] / (stage[top]["ms_avg"] * 1e-3) / 1e9
        roofline = {"kernel": top, "bound": "hbm", "achieved": round(a, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(a / peak, 4), "traffic": ncu_traffic_bytes(top), "peak_source": peak_src,
                    "launches": stage[top]["launches"], "ms_avg": round(stage[top]["ms_avg"], 4),
                    "bytes_per_launch": stage[top]["bytes_per_launch"],
                    "share_of_step": round(stage[top]["ms_total"] / ms, 4)}
    G = 4 * N * C * (shape[2] * shape[3] * (shape[4] if dim == 3 else 1))
    chain_bytes = total * N * CHAIN_BYTES_PER_PAIR[args.workload] + \
        CHAIN_FIELD_TERMS[args.workload] * G * math.ceil(total / world / chunk) * world
    chain_gbs = chain_bytes * args.steps / (ms * 1e-3) / 1e9 / world
    out = {
        "metric": "points/s fwd->triple-bwd (PIXEL Helmholtz step, 2D cosine multicell)"
                  if dim == 2 else "points/s fwd->triple-bwd (3D smoothstep multicell Laplacian step)",
        "value": pts_per_s, "unit": "points/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: cells %s, %d points total, chunk %d, kernel %s, multicell, residual %s"
                               % (args.workload, list(shape), total, chunk, kernel, residual),
                   "points_per_rank": coords_host.shape[0], "parallelism": "dp%d over points" % world,
                   "l2": "inputs larger than L2 (each [N,C,P] stream of a chunk is %d MiB)"
                         % (4 * N * C * min(chunk, total) // 2 ** 20)},
        "e2e": {"value": e2e_pts_per_s, "unit": "points/s",
                "h2d_bytes_per_step": coords_pinned.numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": roofline,
        "chain_roofline": {"bound": "hbm", "bytes_per_step_minimal": chain_bytes,
                           "achieved": round(chain_gbs, 1), "peak": peak, "unit": "GB/s per GPU",
                           "frac": round(chain_gbs / peak, 4)},
        "stages": {k: {"launches": v["launches"], "ms_avg": round(v["ms_avg"], 4),
                       "GBps": round(v["bytes_per_launch"] / (v["ms_avg"] * 1e-3) / 1e9, 1)}
                   for k, v in sorted(stage.items())},
        "loss": loss_val,
        "index_mode": ops.get_index_mode(),
    }
    fstage = fprof.summary()
    ftop = max(fstage, key=lambda k: fstage[k]["ms_total"]) if fstage else None
    fchain_gbs = chain_bytes * args.steps / (ms_f * 1e-3) / 1e9 / world
    out["fused"] = {
        "api": "cosinesampler_b200.jet.FusedPdeStep (opt-in; SURVEY 8f ranks 1+2): jets in one gather pass, "
               "head + residual + gradients in one kernel, one scatter pass; same loss and gradients",
        "chunk": fchunk, "gradient_reduce": reduce_kind,
        "value": total * args.steps / (ms_f * 1e-3), "unit": "points/s", "ms_per_step": ms_f / args.steps,
        "e2e": {"value": total * args.steps / (ms_f_e2e * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": coords_pinned.numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_f_e2e / args.steps},
        "gpu_launches": int(flaunches), "loss": floss_val,
        "speedup_vs_dropin": ms / ms_f,
        "stages": {k: {"launches": v["launches"], "ms_avg": round(v["ms_avg"], 4),
                       "GBps": round(v["bytes_per_launch"] / (v["ms_avg"] * 1e-3) / 1e9, 1),
                       "frac": round(v["bytes_per_launch"] / (v["ms_avg"] * 1e-3) / 1e9 / peak, 4)}
                   for k, v in sorted(fstage.items())},
        "roofline": None if not ftop else {
            "kernel": ftop, "bound": "hbm", "unit": "GB/s", "peak": peak,
            "achieved": round(fstage[ftop]["bytes_per_launch"] / (fstage[ftop]["ms_avg"] * 1e-3) / 1e9, 1),
            "frac": round(fstage[ftop]["bytes_per_launch"] / (fstage[ftop]["ms_avg"] * 1e-3) / 1e9 / peak, 4),
            "bytes_per_launch": fstage[ftop]["bytes_per_launch"], "ms_avg": round(fstage[ftop]["ms_avg"], 4),
            "traffic": None if ncu_traffic_bytes(ftop) is None else int(
                ncu_traffic_bytes(ftop) * (min(fchunk, coords_host.shape[0]) / float(2 ** 20 if dim == 2 else 2 ** 22))),
            "traffic_note": "ncu dram bytes of the committed capture, scaled from its points per launch to this run's",
            "share_of_step": round(fstage[ftop]["ms_total"] / ms_f, 4)},
        "chain_roofline": {"note": "the reference formulation's minimal bytes per step (BASELINE.md section 3) "
                                   "over the fused step's time", "achieved": round(fchain_gbs, 1),
                           "peak": peak, "unit": "GB/s per GPU", "frac": round(fchain_gbs / peak, 4)},
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.workload, args.cpu_sample, repeats=2)
    print(json.dumps(out), file=_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------
# CPU legs: the reference's own CPU path is its pure-PyTorch sampler under torch autograd
# (test/grid_sampler.py, restated device-agnostically in oracle/); timed on the host cores.
# ---------------------------------------------------------------------------------------
def _cpu_step_fn(workload, sample_points):
    import torch
    from oracle.grid_sampler_oracle import grid_sample_2d, grid_sample_3d
    from cosinesampler_b200 import chain
    dim, shape, total, chunk, kernel, residual = WORKLOADS[workload]
    step_name = {"smooth-step": "smoothstep"}.get(kernel, kernel)
    fn = grid_sample_2d if dim == 2 else grid_sample_3d
    sampler = lambda c, g: fn(c, g, step=step_name, offset=True)
    g = torch.Generator().manual_seed(0)
    cells = torch.nn.Parameter(torch.rand(shape, generator=g))
    coords = torch.rand(sample_points, dim, generator=torch.Generator().manual_seed(1000)) * 2 - 1
    head = chain.make_head(shape[1], seed=0)

    def step():
        cells.grad = None
        for p in head.parameters():
            p.grad = None
        return chain.training_step(sampler, cells, [coords[:, a:a + 1] for a in range(dim)], head,
                                   residual=residual)
    return step


def cpu_baseline(workload, sample_points, repeats=2):
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _cpu_step_fn(workload, sample_points)
    step()                                           # warm-up
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        step()
        best = min(best, time.perf_counter() - t0)
    return {"value": sample_points / best, "unit": "points/s", "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": "%d points of the same workload (oracle/grid_sampler_oracle.py under torch CPU "
                      "autograd, fp32, best of %d after 1 warm-up)" % (sample_points, repeats)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dim, shape, total, chunk, kernel, residual = WORKLOADS[args.workload]
    sample = args.cpu_sample
    step = _cpu_step_fn(args.workload, sample)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    desc = {"value": v, "unit": "points/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "each step = %d points of the workload on the host cores" % sample}
    out = {
        "impl": "reference",
        "metric": "points/s fwd->triple-bwd (PIXEL Helmholtz step, 2D cosine multicell)"
                  if dim == 2 else "points/s fwd->triple-bwd (3D smoothstep multicell Laplacian step)",
        "value": v, "unit": "points/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "%s: cells %s, %d points total, chunk %d, kernel %s, multicell, residual %s"
                               % (args.workload, list(shape), total, chunk, kernel, residual),
                   "note": "reference CPU path = its pure-PyTorch sampler (test/grid_sampler.py, oracle "
                           "port) under torch CPU autograd; bounded sample per step"},
        "cpu_baseline": desc,
        "e2e": {"value": v, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), file=_OUT, flush=True)


def _claim_stdout():
    """Keep stdout for the one JSON line: libraries (NCCL prints its version banner there) get
    stderr instead.  Returns a file object bound to the real stdout."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    return real


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=sorted(WORKLOADS))
    ap.add_argument("--points", type=int, default=0, help="override the total number of points")
    ap.add_argument("--cpu-sample", type=int, default=2 ** 19,
                    help="points per CPU-baseline step (bounded sample of the workload)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--nccl-reduce", action="store_true",
                    help="fused arm: reduce the gradients with NCCL instead of the peer-memory kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
