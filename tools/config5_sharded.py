"""BASELINE config 5: N x 1 GB G_dna shards, one per GPU; type
classification + LMS flags + (byte,type) histogram + LMS positions of the whole
N GB text with NCCL collectives only for the tiny summaries.
Launch:  python -m torch.distributed.run --nproc-per-node N tools/config5_sharded.py [shard_bytes]
Rank 0 prints one JSON line (aggregate GB/s of the sharded phases, max over ranks)."""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suffix_b200 import _lib, gen, sharded  # noqa: E402


def ref_types_window(w):
    """S bits of a byte window; positions whose type depends on bytes beyond the window end are marked 2."""
    n = len(w)
    S = np.full(n, 2, dtype=np.uint8)
    nxt_c, nxt_t = int(w[-1]), 2
    for i in range(n - 2, -1, -1):
        c = int(w[i])
        t = 1 if c < nxt_c else 0 if c > nxt_c else nxt_t
        S[i] = t
        nxt_c, nxt_t = c, t
    return S


def main():
    shard_bytes = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    text = gen.dna(shard_bytes, seed=gen.SEED_DNA + rank)
    shard = torch.from_numpy(text).to(dev)
    ctx = _lib.Context(local)
    eng = sharded.CudaShardEngine(ctx, torch)
    d = dist if world > 1 else None
    res = None
    times = []
    for it in range(4):                      # 1 warm-up + 3 timed
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = sharded.classify_sharded(eng, shard, dist=d, device=dev)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
    dt = torch.tensor([min(times[1:])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    # ---- checks: global invariants + boundary windows against a numpy restatement
    ok = int(res.hist_global.sum()) == shard_bytes * world
    ok = ok and int(res.hist_global[512:].sum()) == res.m_total
    W = 4096
    head = torch.from_numpy(text[:W].copy()).to(dev)
    heads = [torch.empty_like(head) for _ in range(world)]
    if world > 1:
        dist.all_gather(heads, head)
    else:
        heads = [head]
    if rank + 1 < world:
        window = np.concatenate([text[-W:], heads[rank + 1].cpu().numpy()])
        S = ref_types_window(window)[:W]
        mine = np.unpackbits(res.stype_words.cpu().numpy().view(np.uint8), bitorder="little")[:shard_bytes][-W:]
        known = S != 2
        ok = ok and bool(np.array_equal(mine[known], S[known])) and bool(known.sum() > W - 64)
    okt = torch.tensor([1 if ok else 0], device=dev)
    if world > 1:
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        total = shard_bytes * world
        print(json.dumps({"config": "BASELINE config 5: sharded classify + LMS flags + histogram + LMS positions",
                          "n_gpus": world, "bytes_total": total, "seconds": round(float(dt.item()), 5),
                          "GBps_aggregate": round(total / 1e9 / float(dt.item()), 1),
                          "m_total": res.m_total, "checks_ok": bool(okt.item()),
                          "collectives": "all_gather 3+1+1 ints/rank, all_reduce 768 x i64"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
