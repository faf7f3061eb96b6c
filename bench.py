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
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the committed
    `ncu --set full` captures of the same shapes (profiles/r2_session2/*_summary.json, else profiles/r2/, r1/; the
    captures drive the same call pattern as this bench: an expanded gOut for the stage kernels, 2^25 (2D) /
    2^22 (3D) binned points per launch for the one-pass kernel).  -> (bytes, points per launch of the capture)"""
    for rnd, name in (("r2_session2", "ncu_onepass_cfg3_summary.json"), ("r2_session2", "ncu_onepass_cfg4_summary.json"),
                      ("r2", "ncu_stages_cfg3_summary.json"), ("r2", "ncu_stages_cfg4_summary.json"),
                      ("r2", "ncu_onepass_cfg3_summary.json"), ("r2", "ncu_onepass_cfg4_summary.json"),
                      ("r1", "ncu_stages_cfg3_summary.json"), ("r1", "ncu_stages_cfg4_summary.json"),
                      ("r1", "ncu_fused_cfg3_summary.json"), ("r1", "ncu_fused_cfg4_summary.json")):
        try:
            with open(os.path.join(ROOT, "profiles", rnd, name)) as f:
                doc = json.load(f)
            if label in doc["kernels"]:
                return (int(doc["kernels"][label]["traffic_bytes"]), doc.get("points_per_launch"),
                        "profiles/%s/%s" % (rnd, name), doc["kernels"][label].get("warp_instructions"))
        except Exception:
            pass
    return None, None, None, None


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
    """Collects CUDA-event timings of every kernel bracket (ops._timed) issued inside the timed region."""

    def __init__(self):
        self.records = []

    def record(self, label, nbytes, start, end, moved=None):
        self.records.append((label, nbytes, start, end, nbytes if moved is None else moved))

    def summary(self):
        agg = {}
        for label, nbytes, s, e, moved in self.records:
            a = agg.setdefault(label, [0, 0.0, 0, 0])
            a[0] += 1
            a[1] += s.elapsed_time(e)
            a[2] += nbytes
            a[3] += moved
        return {k: {"launches": v[0], "ms_total": v[1], "ms_avg": v[1] / v[0],
                    "bytes_per_launch": v[2] // v[0], "moved_per_launch": v[3] // v[0]} for k, v in agg.items()}


def stage_table(stage, peak):
    return {k: {"launches": v["launches"], "ms_avg": round(v["ms_avg"], 4),
                "GBps": round(v["bytes_per_launch"] / (v["ms_avg"] * 1e-3) / 1e9, 1),
                "frac": round(v["bytes_per_launch"] / (v["ms_avg"] * 1e-3) / 1e9 / peak, 4),
                "frac_as_moved": round(v["moved_per_launch"] / (v["ms_avg"] * 1e-3) / 1e9 / peak, 4)}
            for k, v in sorted(stage.items())}


def kernel_roofline(stage, top, peak, peak_src, ms_step_total, points_per_launch=None, sms=148, sm_clock_hz=None):
    v = stage[top]
    a = v["bytes_per_launch"] / (v["ms_avg"] * 1e-3) / 1e9
    m = v["moved_per_launch"] / (v["ms_avg"] * 1e-3) / 1e9
    traffic, cap_points, src, winstr = ncu_traffic_bytes(top)
    if traffic is not None and cap_points and points_per_launch:
        traffic = int(traffic * points_per_launch / float(cap_points))
    issue = None
    if winstr and cap_points and points_per_launch and sm_clock_hz:
        # instruction-issue roofline of a kernel that is not memory-bound: warp instructions per point from the
        # committed ncu capture x the points of this launch / its measured time, against one warp instruction per
        # cycle and SM sub-partition (4 per SM)
        per_point = winstr / float(cap_points)
        peak_issue = sms * 4 * sm_clock_hz
        ach = per_point * points_per_launch / (v["ms_avg"] * 1e-3)
        issue = {"bound": "instruction issue", "warp_instructions_per_point": round(per_point, 1),
                 "achieved": round(ach / 1e9, 1), "peak": round(peak_issue / 1e9, 1), "unit": "G warp instructions/s",
                 "frac": round(ach / peak_issue, 4),
                 "note": "instructions per point from the ncu capture in traffic_source; peak = SMs x 4 schedulers x SM clock"}
    return {"kernel": top, "issue_roofline": issue, "bound": "hbm", "achieved": round(a, 1), "peak": peak, "unit": "GB/s",
            "frac": round(a / peak, 4), "achieved_as_moved": round(m, 1), "frac_as_moved": round(m / peak, 4),
            "traffic": traffic, "traffic_source": src, "peak_source": peak_src,
            "launches": v["launches"], "ms_avg": round(v["ms_avg"], 4), "bytes_per_launch": v["bytes_per_launch"],
            "moved_per_launch": v["moved_per_launch"], "share_of_step": round(v["ms_total"] / ms_step_total, 4)}


# ---------------------------------------------------------------------------------------
def make_inputs(workload, rank, world, device, dtype, total=None):
    import torch
    dim, shape, wtotal, chunk, kernel, residual = WORKLOADS[workload]
    total = total or wtotal
    from cosinesampler_b200 import chain, dp
    g = torch.Generator().manual_seed(0)
    cells = torch.rand(shape, generator=g, dtype=dtype)                 # U(0,1), replicated
    s, e = dp.shard_range(total, rank, world)
    gr = torch.Generator().manual_seed(1000 + rank)
    coords_host = torch.rand(e - s, dim, generator=gr, dtype=dtype) * 2 - 1     # U(-1,1), unsorted
    head = chain.make_head(shape[1], seed=0, device=device, dtype=dtype)
    return cells, coords_host, head


def measure_workload(args, workload, steps, warmup, ctx, with_jets):
    """All ranks run this; rank 0 gets the result dictionary (others None)."""
    import torch
    import torch.distributed as dist
    from cosinesampler_b200 import _lib, chain, dp, fused, jet, ops
    from cosine_sampler_2d import CosineSampler2d
    from cosine_sampler_3d import CosineSampler3d
    rank, world, device = ctx["rank"], ctx["world"], ctx["device"]

    dim, shape, total, chunk, kernel, residual = WORKLOADS[workload]
    if args.points and workload == args.workload:
        total = args.points
    N, C = shape[:2]
    S = CosineSampler2d if dim == 2 else CosineSampler3d
    sampler = lambda c, g: S.apply(c, g, "zeros", True, kernel, True)
    cells_h, coords_host, head = make_inputs(workload, rank, world, device, torch.float32, total)
    cells = torch.nn.Parameter(cells_h.to(device))
    coords_pinned = coords_host.pin_memory()
    coords_dev = coords_host.to(device)
    nloc = coords_host.shape[0]
    stepper = dp.PointShardedStep(sampler, cells, head, residual=residual, chunk=chunk)
    fused_kw = dict(kernel=kernel, multicell=True)
    # the one-pass step keeps nothing per point but the coordinates: a whole shard is one chunk (the more points
    # a chunk holds, the more of them share a texel and the fewer reds leave the SMs); the end-to-end arm cuts
    # the shard into pieces so that the host->device copy of one piece overlaps the pass over the previous one
    fchunk = max(1, nloc)
    fchunk_e2e = max(2 ** 21, (nloc + 3) // 4)
    reduce_kind = "none (single GPU)"
    fstepper = None
    want_peer = world > 1 and not args.nccl_reduce
    if want_peer:
        # all ranks must take the same path: vote BEFORE the collective rendezvous of the reducer
        ok = 1.0
        try:
            import torch.distributed._symmetric_memory  # noqa: F401
        except Exception as exc:
            print("symmetric memory unavailable (%s); using NCCL" % exc, file=sys.stderr)
            ok = 0.0
        vote = torch.tensor([ok], device=device)
        dist.all_reduce(vote, op=dist.ReduceOp.MIN)
        if float(vote.item()) > 0.5:
            fstepper = dp.PointShardedStep(None, cells, head, residual=residual, chunk=fchunk, fused=fused_kw,
                                           peer_reduce=True)
            reduce_kind = "peer-memory kernel: " + fstepper.reducer.kind
    if fstepper is None:
        fstepper = dp.PointShardedStep(None, cells, head, residual=residual, chunk=fchunk, fused=fused_kw)
        if world > 1:
            reduce_kind = "NCCL all-reduce of one flat bucket"
    jstepper = None
    if with_jets and world == 1:
        jstepper = dp.PointShardedStep(None, cells, head, residual=residual, chunk=4 * chunk,
                                       fused=dict(fused_kw, mode="jets"))

    def fused_resident():
        fstepper.zero_grad()
        return fstepper.step(coords_dev, total)

    def jets_resident():
        jstepper.zero_grad()
        return jstepper.step(coords_dev, total)

    def step_resident():
        stepper.zero_grad()
        cols = [coords_dev[:, a:a + 1] for a in range(dim)]
        return stepper.step(cols, total)

    loss_host = torch.zeros(1, pin_memory=True)
    copy_stream = torch.cuda.Stream(device=device)

    def fetch(span):
        with torch.cuda.stream(copy_stream):
            t = coords_pinned[span[0]:span[1]].to(device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return t, ev

    def h2d_probe():
        """Bandwidth of the copy the end-to-end arms depend on: the rank's pinned coordinates to the device,
        alone on the copy stream (best of 3).  A box whose host memory system is slow shows up here."""
        best = None
        for _ in range(3):
            e0 = torch.cuda.Event(enable_timing=True)
            e1 = torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            with torch.cuda.stream(copy_stream):
                e0.record(copy_stream)
                t = coords_pinned.to(device, non_blocking=True)
                e1.record(copy_stream)
            torch.cuda.synchronize()
            del t
            ms_ = e0.elapsed_time(e1)
            best = ms_ if best is None else min(best, ms_)
        return coords_pinned.numel() * 4 / (best * 1e-3) / 1e9 if best and best > 0 else None

    # input batches are double-buffered across steps as well: the copy of a step's FIRST chunk is issued while the
    # previous step finishes (its reduce, post-mix and loss read-back), like a training loop that prefetches its next
    # batch of collocation points.  Every step still copies all of its coordinates inside the timed region.
    pending = {}

    def first_chunk(arm, span):
        nxt = pending.pop(arm, None)
        return nxt if nxt is not None else fetch(span)

    def step_e2e():
        """Same step from HOST buffers: every chunk of coordinates is copied from pinned host
        memory inside the timed region (on a copy stream, one chunk ahead of the compute stream) and
        the step's loss is read back to the host."""
        stepper.zero_grad()
        spans = [(s0, min(nloc, s0 + chunk)) for s0 in range(0, nloc, chunk)]
        cur = torch.cuda.current_stream()
        nxt = first_chunk("dropin", spans[0])
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
        pending["dropin"] = fetch(spans[0])                               # the next step's first chunk
        dp.allreduce_grads(stepper.params())
        loss_host.copy_(acc.reshape(1), non_blocking=True)               # D2H of the step's result
        cur.synchronize()
        return loss_host

    def fused_e2e():
        """The fused step from HOST buffers, same copy pipeline as step_e2e."""
        fstepper.zero_grad()
        spans = [(s0, min(nloc, s0 + fchunk_e2e)) for s0 in range(0, nloc, fchunk_e2e)]
        cur = torch.cuda.current_stream()
        if fstepper.mode == "onepass":
            fs = fused.OnePassPdeStep(cells, head, residual, kernel=kernel, multicell=True)
        else:
            fs = jet.FusedPdeStep(cells, head, residual, kernel=kernel, multicell=True)
        fs.begin(fstepper.reducer, scale=1.0 / float(total))
        nxt = first_chunk("fused", spans[0])
        for i, span in enumerate(spans):
            t, ev = nxt
            if i + 1 < len(spans):
                nxt = fetch(spans[i + 1])
            cur.wait_event(ev)
            t.record_stream(cur)
            fs.add(t, 1.0 / float(total))
        pending["fused"] = fetch(spans[0])                                # the next step's first chunk
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

    def timed(fn, nsteps):
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(nsteps):
            fn()
        t1.record()
        barrier()
        ms = torch.tensor([t0.elapsed_time(t1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- the drop-in operator (the reference's API; what PIXEL runs unchanged)
    for _ in range(warmup):
        step_resident()
    barrier()
    clocks = ClockSampler(ctx["local"])
    prof = StageProfiler()
    if rank == 0:
        clocks.start()
    ops.profiler = prof
    n0 = _lib.launch_count()
    ms = timed(step_resident, steps)
    launches = _lib.launch_count() - n0
    ops.profiler = None
    clock_info = clocks.stop() if rank == 0 else None
    loss_t = step_resident().detach().clone()
    if world > 1:
        dist.all_reduce(loss_t)                      # each rank holds its share of the mean
    loss_val = float(loss_t.item())
    h2d_gbs = h2d_probe()
    for _ in range(max(1, min(warmup, 2))):
        step_e2e()
    ms_e2e = timed(step_e2e, steps)
    pending.clear()

    # ---- the opt-in fused step, same inputs, same outputs
    for _ in range(warmup):
        fused_resident()
    barrier()
    fprof = StageProfiler()
    ops.profiler = fprof
    n0 = _lib.launch_count()
    ms_f = timed(fused_resident, steps)
    flaunches = _lib.launch_count() - n0
    ops.profiler = None
    floss_t = fused_resident().detach().clone()
    if world > 1 and not fstepper.loss_is_global:
        dist.all_reduce(floss_t)
    floss_val = float(floss_t.item())
    for _ in range(max(1, min(warmup, 2))):
        fused_e2e()
    ms_f_e2e = timed(fused_e2e, steps)
    pending.clear()
    ms_j = None
    if jstepper is not None:
        for _ in range(2):
            jets_resident()
        ms_j = timed(jets_resident, 2) / 2

    if rank != 0:
        return None

    peak, peak_src = measured_peak_gbs()
    stage = prof.summary()
    sampler_ms = sum(v["ms_total"] for v in stage.values()) / steps
    kernel_stages = {k: v for k, v in stage.items() if not k.startswith("AUX")}
    top = max(kernel_stages, key=lambda k: kernel_stages[k]["ms_total"]) if kernel_stages else None
    T = 1
    for s_ in shape[2:]:
        T *= s_
    G = 4 * N * C * T
    chain_bytes = total * N * CHAIN_BYTES_PER_PAIR[workload] + \
        CHAIN_FIELD_TERMS[workload] * G * math.ceil(total / world / chunk) * world
    chain_gbs = chain_bytes * steps / (ms * 1e-3) / 1e9 / world
    sampler_gbs = chain_bytes / world / (sampler_ms * 1e-3) / 1e9
    wl = "%s: cells %s, %d points total, chunk %d, kernel %s, multicell, residual %s" \
        % (workload, list(shape), total, chunk, kernel, residual)
    out = {
        "metric": "points/s fwd->triple-bwd (PIXEL Helmholtz step, 2D cosine multicell)"
                  if dim == 2 else "points/s fwd->triple-bwd (3D smoothstep multicell Laplacian step)",
        "value": total * steps / (ms * 1e-3), "unit": "points/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms / steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "points_per_rank": nloc, "parallelism": "dp%d over points" % world,
                   "l2": "inputs larger than L2 (each [N,C,P] stream of a chunk is %d MiB)"
                         % (4 * N * C * min(chunk, total) // 2 ** 20),
                   "e2e_pipeline": "coordinates copied from pinned host memory in chunks on a copy stream, one chunk "
                                   "ahead of the compute stream; the first chunk of the next step is prefetched while "
                                   "the current step finishes"},
        "e2e": {"value": total * steps / (ms_e2e * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": coords_pinned.numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / steps, "h2d_GBps_alone": round(h2d_gbs, 1) if h2d_gbs else None},
        "gpu_launches": int(launches),
        "clocks": clock_info,
        "roofline": kernel_roofline(stage, top, peak, peak_src, ms) if top else None,
        "chain_roofline": {"bound": "hbm", "bytes_per_step_minimal": chain_bytes,
                           "achieved": round(chain_gbs, 1), "peak": peak, "unit": "GB/s per GPU",
                           "frac": round(chain_gbs / peak, 4),
                           "note": "BASELINE.md's minimal bytes of the step over the whole step time, the "
                                   "caller's torch head included"},
        "sampler_only": {"ms_per_step": round(sampler_ms, 4),
                         "ms_per_2^20_points": round(sampler_ms * 2 ** 20 * world / total, 4),
                         "achieved": round(sampler_gbs, 1), "unit": "GB/s per GPU",
                         "frac": round(sampler_gbs / peak, 4),
                         "share_of_step": round(sampler_ms / (ms / steps), 4),
                         "note": "sum of the CUDA-event brackets of every kernel this library launches in a step "
                                 "(stage kernels + accumulator memsets + layout transposes), BASELINE.md's "
                                 "minimal bytes over that time: the operator's own roofline fraction"},
        "stages": stage_table(stage, peak),
        "loss": loss_val,
        "index_mode": ops.get_index_mode(),
    }
    fstage = fprof.summary()
    fkern = {k: v for k, v in fstage.items() if k.startswith("ONEPASS") or k.startswith("JET") or k.startswith("HEAD")}
    ftop = max(fkern, key=lambda k: fkern[k]["ms_total"]) if fkern else None
    fchain_gbs = chain_bytes * steps / (ms_f * 1e-3) / 1e9 / world
    out["fused"] = {
        "api": "cosinesampler_b200.fused.OnePassPdeStep (opt-in; SURVEY 8f ranks 1+2): cells mixed with the head's first "
               "layer once per step, points binned by texel, gather -> tanh / residual / gradients -> scatter in ONE "
               "kernel, register-aggregated reds; same loss and gradients" if fstepper.mode == "onepass" else
               "cosinesampler_b200.jet.FusedPdeStep (jets -> tensor-core head -> scatter)",
        "mode": fstepper.mode, "chunk": fchunk, "chunk_e2e": fchunk_e2e, "gradient_reduce": reduce_kind,
        "value": total * steps / (ms_f * 1e-3), "unit": "points/s", "ms_per_step": ms_f / steps,
        "ms_per_2^20_points": round(ms_f / steps * 2 ** 20 * world / total, 4),
        "e2e": {"value": total * steps / (ms_f_e2e * 1e-3), "unit": "points/s",
                "h2d_bytes_per_step": coords_pinned.numel() * 4, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_f_e2e / steps, "h2d_GBps_alone": round(h2d_gbs, 1) if h2d_gbs else None,
                "h2d_ms_alone": round(coords_pinned.numel() * 4 / (h2d_gbs * 1e9) * 1e3, 3) if h2d_gbs else None},
        "gpu_launches": int(flaunches), "loss": floss_val,
        "speedup_vs_dropin": ms / ms_f,
        "stages": stage_table(fstage, peak),
        "roofline": kernel_roofline(fstage, ftop, peak, peak_src, ms_f, points_per_launch=min(fchunk, nloc),
                                    sms=torch.cuda.get_device_properties(device).multi_processor_count,
                                    sm_clock_hz=(clock_info or {}).get("sm_mhz", 0) and clock_info["sm_mhz"] * 1e6) if ftop else None,
        "roofline_note": "the one-pass kernel moves 4*dim bytes per point plus the two grid-shaped fields: it is bound by "
                         "instruction issue and the L1 / red paths, not by HBM (profiles/README.md); its HBM fraction "
                         "is reported for completeness and roofline.issue_roofline gives the fraction of the issue slots",
        "chain_roofline": {"note": "the reference formulation's minimal bytes per step (BASELINE.md section 3) "
                                   "over the fused step's time", "achieved": round(fchain_gbs, 1),
                           "peak": peak, "unit": "GB/s per GPU", "frac": round(fchain_gbs / peak, 4)},
    }
    # the opt-in arm's two numbers also at the top level of the line (same units as value / e2e)
    out["fused_value"] = out["fused"]["value"]
    out["fused_e2e_value"] = out["fused"]["e2e"]["value"]
    if ms_j is not None:
        out["fused"]["round1_jets_path"] = {"value": total / (ms_j * 1e-3), "unit": "points/s", "ms_per_step": ms_j,
                                            "note": "jet.FusedPdeStep (jets -> tensor-core head -> scatter, 3 launches "
                                                    "per chunk of %d points), 2 timed steps" % (4 * chunk)}
    return out


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
    ctx = {"rank": rank, "world": world, "local": local, "device": device}
    out = measure_workload(args, args.workload, args.steps, args.warmup, ctx, with_jets=not args.no_extras)
    # the 3D configuration rides along in the default single-GPU run (a driver-run record of config 4)
    extra = None
    if args.workload == "cfg5" and world == 1 and not args.no_extras and not args.points:
        torch.cuda.empty_cache()
        extra = measure_workload(args, "cfg4", max(2, min(args.steps, 5)), 3, ctx, with_jets=True)
    if rank == 0:
        if extra is not None:
            for k in ("warmup", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "clocks", "index_mode", "n_gpus"):
                extra.pop(k, None)
            out["cfg4"] = extra
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
                   "sample_points_per_step": sample,
                   "note": "reference CPU path = its pure-PyTorch sampler (test/grid_sampler.py, oracle "
                           "port) under torch CPU autograd; each step is a bounded sample of %d points of the "
                           "workload (points/s is size-normalised)" % sample},
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
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the config-4 (3D) measurement and the round-1 jets path that ride along at N = 1")
    ap.add_argument("--nccl-reduce", action="store_true",
                    help="fused arm: reduce the gradients with NCCL instead of the peer-memory kernel")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
