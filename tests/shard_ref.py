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
