"""Text-position sharding of the classification phases over several GPUs
(BASELINE config 5): one process per GPU, each holding one
contiguous shard of the text; only tiny summaries cross NVLink.

Collectives (torch.distributed; NCCL on GPUs, gloo in the CPU tests):
  * all_gather of 4 ints per rank  (first byte, last byte, shard state, length)
  * all_reduce(sum) of the 768-bin (byte, type) histogram
  * all_gather of the per-shard LMS counts (exclusive prefix = global LMS rank offset)

The induce recursion itself is single-device (north_star); this module covers
SuffixTypes::compute (src/table.rs:592-615), Bins::find_sizes (:686-704) and the
LMS position list (P15, :512-520) for a sharded text.
"""
from dataclasses import dataclass

import numpy as np

ST_L, ST_S, ST_P = 0, 1, 2


class CudaShardEngine:
    """The product engine: the CUDA kernels behind b200sa_shard_* (no CPU path)."""

    def __init__(self, ctx, torch_mod):
        self.ctx = ctx
        self.torch = torch_mod

    def edge_bytes(self, shard):
        return int(shard[0].item()), int(shard[-1].item())

    def summary(self, shard, next_char):
        return self.ctx.shard_summary(shard.data_ptr(), shard.numel(), next_char,
                                      self.torch.cuda.current_stream().cuda_stream)

    def classify(self, shard, prev_char, next_char, tail_carry):
        t = self.torch
        n = shard.numel()
        nw = (n + 31) // 32
        stype = t.empty(nw, dtype=t.int32, device=shard.device)
        lms = t.empty(nw, dtype=t.int32, device=shard.device)
        lmspos = t.empty(n // 2 + 1, dtype=t.int32, device=shard.device)
        hist, m = self.ctx.shard_classify(shard.data_ptr(), n, prev_char, next_char, tail_carry,
                                          stype.data_ptr(), lms.data_ptr(), lmspos.data_ptr(), lmspos.numel(),
                                          t.cuda.current_stream().cuda_stream)
        t.cuda.synchronize()
        return stype, lms, lmspos[:m], hist, m


@dataclass
class ShardResult:
    lo: int                 # global offset of this shard
    stype_words: object     # S-type bitmap of the shard (engine's array type)
    lms_words: object       # LMS bitmap of the shard
    lmspos_local: object    # shard-local LMS positions, ascending
    m_local: int
    m_offset: int           # number of LMS positions in earlier shards
    m_total: int
    hist_global: np.ndarray  # 768 x u64: L / S-non-LMS / LMS counts per byte over the WHOLE text
    state: int
    tail_carry: int


def resolve_tail_carries(states):
    """tail_carry[r] = first state != P among shards r+1.. (ST_L if none: unused,
    the last shard ends the text and position n-1 is Descending, src/table.rs:602)."""
    out = []
    for r in range(len(states)):
        tc = ST_L
        for s in states[r + 1:]:
            if s != ST_P:
                tc = s
                break
        out.append(tc)
    return out


def classify_sharded(engine, shard, dist=None, device=None):
    """Collective call: every rank passes its own contiguous shard (rank order =
    text order).  `dist` is torch.distributed (initialised) or None for a single
    shard.  Returns a ShardResult."""
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    n_local = int(shard.shape[0]) if hasattr(shard, "shape") else len(shard)
    first, last = engine.edge_bytes(shard)

    def all_gather_ints(vals):
        t = torch.tensor(vals, dtype=torch.int64, device=device)
        if world == 1:
            return [list(vals)]
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [o.cpu().tolist() for o in out]

    edges = all_gather_ints([first, last, n_local])
    next_char = edges[rank + 1][0] if rank + 1 < world else -1
    prev_char = edges[rank - 1][1] if rank > 0 else -1
    lo = sum(e[2] for e in edges[:rank])

    state = engine.summary(shard, next_char)
    states = [s[0] for s in all_gather_ints([state])]
    tail = resolve_tail_carries(states)[rank]

    stype, lms, lmspos, hist, m = engine.classify(shard, prev_char, next_char, tail)

    h = torch.from_numpy(hist.astype(np.int64)).to(device) if device is not None else torch.from_numpy(hist.astype(np.int64))
    if world > 1:
        dist.all_reduce(h)                       # sum over shards: global bucket sizes
    ms = [x[0] for x in all_gather_ints([m])]
    return ShardResult(lo=lo, stype_words=stype, lms_words=lms, lmspos_local=lmspos, m_local=m,
                       m_offset=sum(ms[:rank]), m_total=sum(ms), hist_global=h.cpu().numpy().astype(np.uint64),
                       state=state, tail_carry=tail)


# ---------------------------------------------------------------------------------------
# Sharded LMS-suffix sort (BASELINE config 5): the collectives run
# INSIDE the library (NCCL behind the C-ABI, b200sa_shard_lms_sort); Python only hands
# the 128-byte NCCL unique id from rank 0 to the other ranks.
_comm_ready = {}


def ensure_comm(ctx, dist=None):
    """Creates the library-side communicator of `ctx` once (rank / world of torch.distributed)."""
    key = id(ctx)
    if key in _comm_ready:
        return
    if dist is None or dist.get_world_size() == 1:
        _comm_ready[key] = True
        return
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", ctx.device)
    uid = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        uid = torch.frombuffer(bytearray(ctx.comm_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(uid, src=0)
    ctx.comm_init(world, rank, bytes(uid.cpu().numpy().tobytes()))
    _comm_ready[key] = True


def lms_sort_sharded(ctx, shard, dist=None, cap=None):
    """Collective: every rank passes its contiguous shard (a CUDA uint8 tensor, rank order = text
    order).  Returns (gpos, names, stats): this rank's slice of the LMS suffixes of the whole text
    ordered by their first `stats['kc']` characters (int64 global positions), the dense global
    window names, and the statistics of b200sa_shard_stats.  stats['ties_total'] == 0 means the
    concatenated slices are the LMS suffixes in exact suffix order."""
    import torch
    ensure_comm(ctx, dist)
    n = shard.numel()
    if cap is None:                      # slices are balanced by the sampled splitters: ~ m_total / world entries
        tot = torch.tensor([n], dtype=torch.int64, device=shard.device)
        if dist is not None and dist.get_world_size() > 1:
            dist.all_reduce(tot)
        world = dist.get_world_size() if dist is not None else 1
        cap = int(tot.item()) // 2 // world * 2 + 65536 if world > 1 else n // 2 + 4096
    cap = int(cap)
    gpos = torch.empty(cap, dtype=torch.int64, device=shard.device)
    names = torch.empty(cap, dtype=torch.int32, device=shard.device)
    st = ctx.shard_lms_sort(shard.data_ptr(), n, gpos.data_ptr(), names.data_ptr(), cap,
                            torch.cuda.current_stream().cuda_stream)
    k = int(st["recv_count"])
    return gpos[:k], names[:k], st
