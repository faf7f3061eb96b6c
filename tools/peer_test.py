"""torchrun script: the fused step's gradient reduce over peer memory (peer.PeerReducer) against the
NCCL path (dp.allreduce_grads) -- same gradients on every rank, and the time of a whole step each way.
    python -m torch.distributed.run --nproc-per-node N tools/peer_test.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from cosinesampler_b200 import dp
    from cosinesampler_b200.chain import make_head
    total = 1 << 22
    g = torch.Generator().manual_seed(0)
    cells0 = torch.rand(4, 16, 256, 256, generator=g)
    s, e = dp.shard_range(total, rank, world)
    xy = (torch.rand(e - s, 2, generator=torch.Generator().manual_seed(100 + rank)) * 2 - 1).to(dev)
    res = {}
    steppers = {}
    kinds = {}
    # nccl: one-pass step + NCCL all-reduce;  peer: one-pass step + cs_peer_allreduce (multimem when the box has
    # multicast objects);  peer_ld: the same with 16-byte peer loads;  jets_peer: the round-1 path with the
    # transposing cs_peer_allreduce_from_channel_last
    for mode in ("nccl", "peer", "peer_ld", "jets_peer"):
        cells = torch.nn.Parameter(cells0.clone().to(dev))
        head = make_head(16, seed=0, device=dev)
        fk = dict(kernel="cosine", multicell=True)
        if mode == "jets_peer":
            fk["mode"] = "jets"
        st = dp.PointShardedStep(None, cells, head, residual="helmholtz", chunk=1 << 22, fused=fk,
                                 peer_reduce=(mode != "nccl"), peer_multicast=(False if mode == "peer_ld" else None))
        kinds[mode] = st.reducer.kind if st.reducer is not None else "NCCL all-reduce"
        steppers[mode] = st
        st.zero_grad()
        loss = st.step(xy, total).detach().clone()
        if not st.loss_is_global:
            dist.all_reduce(loss)
        res[mode] = (loss, cells.grad.clone(), [p.grad.clone() for p in head.parameters()])
    torch.cuda.synchronize()
    ok = True
    msgs = []
    for mode in ("peer", "peer_ld", "jets_peer"):
        for name, a, b in [("loss", res[mode][0], res["nccl"][0]), ("cells.grad", res[mode][1], res["nccl"][1])] + \
                [("head%d" % i, a, b) for i, (a, b) in enumerate(zip(res[mode][2], res["nccl"][2]))]:
            err = float((a - b).abs().max() / (b.abs().max() + 1e-30))
            msgs.append("%s %s %.2e" % (mode, name, err))
            ok = ok and err < (2e-4 if name.startswith("head") else 2e-5)
    # every rank must hold the same bits after the peer reduce
    gsum = res["peer"][1].double().sum().reshape(1)
    lst = [torch.zeros_like(gsum) for _ in range(world)]
    dist.all_gather(lst, gsum)
    same = all(float(x) == float(lst[0]) for x in lst)
    times = {}
    for mode, st in steppers.items():
        for _ in range(3):
            st.zero_grad(); st.step(xy, total)
        dist.barrier(); torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(10):
            st.zero_grad(); st.step(xy, total)
        t1.record()
        dist.barrier(); torch.cuda.synchronize()
        ms = torch.tensor([t0.elapsed_time(t1) / 10], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times[mode] = float(ms)
    if rank == 0:
        print(json.dumps({"world": world, "points": total, "match": ok, "same_bits_on_all_ranks": same,
                          "rel_err": msgs, "ms_step": times, "reducers": kinds}), flush=True)
    dist.destroy_process_group()
    if not (ok and same):
        sys.exit(1)


if __name__ == "__main__":
    main()
