"""numpy stand-in for the shard engine (TEST INFRASTRUCTURE): same interface as
suffix_b200.sharded.CudaShardEngine, computed on the CPU so that the N>1
coordination logic can run under gloo without a GPU."""
import numpy as np

ST_L, ST_S, ST_P = 0, 1, 2


def _types_with_edge(t, next_char, tail_carry):
    """S bit per position of shard t given the char after it and the type of that position."""
    n = len(t)
    S = np.zeros(n, dtype=np.uint8)
    nxt_c = next_char
    nxt_t = tail_carry
    for i in range(n - 1, -1, -1):
        c = int(t[i])
        if nxt_c < 0:
            ty = ST_L                      # position n-1 of the text
        elif c < nxt_c:
            ty = ST_S
        elif c > nxt_c:
            ty = ST_L
        else:
            ty = nxt_t
        S[i] = ty
        nxt_c, nxt_t = c, ty
    return S


class NumpyShardEngine:
    def edge_bytes(self, shard):
        return int(shard[0]), int(shard[-1])

    def summary(self, shard, next_char):
        t = np.asarray(shard)
        S = _types_with_edge(t, next_char, ST_P)
        return int(S[0])

    def classify(self, shard, prev_char, next_char, tail_carry):
        t = np.asarray(shard)
        n = len(t)
        S = _types_with_edge(t, next_char, tail_carry)
        assert not (S == ST_P).any()
        prev_t = 1 if prev_char < 0 else (1 if prev_char < t[0] else 0 if prev_char > t[0] else int(S[0]))
        before = np.concatenate([[prev_t], S[:-1]])
        lms = (S == 1) & (before == 0)
        hist = np.zeros(768, dtype=np.uint64)
        cls = S.astype(np.int64) + lms.astype(np.int64)
        np.add.at(hist, t.astype(np.int64) + 256 * cls, 1)
        pack = lambda b: np.packbits(np.concatenate([b, np.zeros((-n) % 32, dtype=np.uint8)]), bitorder="little").view(np.uint32)
        pos = np.flatnonzero(lms).astype(np.uint32)
        return pack(S.astype(np.uint8)), pack(lms.astype(np.uint8)), pos, hist, len(pos)


# ---------------------------------------------------------------------------------------
# numpy mirror of b200sa_shard_lms_sort (suffix_b200/csrc/b200sa.cu): same steps, same
# conventions (64-byte heads as halo source, window keys base sigma, splitters from an
# all-gathered sample, stable partition, chunks received in DESCENDING source rank, stable
# sort, truncated windows are groups of their own), with torch.distributed collectives
# (gloo).  TEST INFRASTRUCTURE: validates the distributed formulation on the CPU.
HEAD = 64
PER = 64


def lms_sort_sharded_model(shard, dist):
    import torch
    t = np.asarray(shard, dtype=np.uint8)
    world, rank = dist.get_world_size(), dist.get_rank()
    eng = NumpyShardEngine()

    def all_gather_obj(x):
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    recs = all_gather_obj((len(t), int(t[0]), int(t[-1]), bytes(t[:HEAD])))
    lens = [r[0] for r in recs]
    lo = sum(lens[:rank]); n_total = sum(lens)
    next_char = recs[rank + 1][1] if rank + 1 < world else -1
    prev_char = recs[rank - 1][2] if rank > 0 else -1
    halo = b""
    for r in range(rank + 1, world):
        if len(halo) >= HEAD:
            break
        halo += recs[r][3][:min(lens[r], HEAD)]
    halo = np.frombuffer(halo[:HEAD], dtype=np.uint8)
    state = eng.summary(t, next_char)
    states = all_gather_obj(state)
    tail = ST_L
    for s in states[rank + 1:]:
        if s != ST_P:
            tail = s
            break
    _, _, lmspos, hist, m = eng.classify(t, prev_char, next_char, tail if next_char >= 0 else ST_L)
    hsum = sum(np.asarray(h, dtype=np.uint64) for h in all_gather_obj(hist))
    present = (hsum[:256] + hsum[256:512] + hsum[512:]) > 0
    code_of = np.cumsum(present) - present
    sigma = max(2, int(present.sum()))
    kc, r_ = 0, 1
    while kc < 32 and r_ * sigma <= (1 << 64):
        r_ *= sigma
        kc += 1
    ext = np.concatenate([t, halo])

    def window(p):
        key = 0
        for j in range(kc):
            c = int(code_of[ext[p + j]]) if p + j < len(ext) else 0
            key = key * sigma + c
        return key

    pos_desc = lmspos[::-1].astype(np.int64)
    keys = [window(int(p)) for p in pos_desc]
    samp = [keys[i * m // PER] if m else (1 << 64) - 1 for i in range(PER)]
    allsamp = sorted(x for s in all_gather_obj(samp) for x in s)
    split = [allsamp[(i + 1) * len(allsamp) // world] for i in range(world - 1)]
    dest = [sum(1 for s in split if s <= k) for k in keys]
    send = [[(keys[i], int(pos_desc[i])) for i in range(m) if dest[i] == d] for d in range(world)]   # stable
    recv = [None] * world
    gathered = all_gather_obj(send)                       # world x world lists (model: exchange through all-gather)
    for src in range(world):
        recv[src] = gathered[src][rank]
    items = []
    for src in range(world - 1, -1, -1):                  # descending source rank
        items += [(k, sum(lens[:src]) + p) for k, p in recv[src]]
    items.sort(key=lambda kp: kp[0])                      # Python's sort is stable
    gpos = [p for _, p in items]
    heads = []
    for i in range(len(items)):
        tr = lambda j: gpos[j] + kc > n_total
        heads.append(i == 0 or items[i][0] != items[i - 1][0] or tr(i - 1) or tr(i))
    distinct = sum(heads)
    ties = 0
    for i in range(len(items)):
        tl = (i + 1 == len(items)) or heads[i + 1]
        if not (heads[i] and tl):
            ties += 1
    dl = all_gather_obj((distinct, ties))
    off = sum(d for d, _ in dl[:rank])
    names, run = [], -1
    for h in heads:
        run += 1 if h else 0
        names.append(off + run)
    return np.array(gpos, dtype=np.int64), np.array(names, dtype=np.int64), {
        "ties_total": sum(x for _, x in dl), "kc": kc, "m_total": sum(len(x) for row in gathered for x in row) // 1,
        "lo": lo, "n_total": n_total}
