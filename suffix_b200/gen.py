"""Deterministic synthetic inputs, numpy-vectorised.

PRNG: splitmix64 with state s0 = seed; output k uses s = seed + (k+1)*GOLDEN.
"""
import os

import numpy as np

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
SEED_DNA = 0x5AFE5EED0000D7A4
SEED_BYTES = 0x5AFE5EED000000FF
SEED_ENGLISH = 0x5AFE5EED0000E416

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, count: int, start: int = 0) -> np.ndarray:
    """Outputs start..start+count-1 of the splitmix64 stream seeded with `seed`."""
    with np.errstate(over="ignore"):
        k = np.arange(start + 1, start + count + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + k * GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        return z ^ (z >> np.uint64(31))


def dna(n: int, seed: int = SEED_DNA, newline_tail: bool = False) -> np.ndarray:
    """G_dna: 32 symbols per 64-bit output, low bits first, "ACGT"[(z>>2k)&3]."""
    out = np.empty(n, dtype=np.uint8)
    lut = np.frombuffer(b"ACGT", dtype=np.uint8)
    chunk = 1 << 20  # outputs per chunk (32 MiB of text)
    pos = 0
    k0 = 0
    shifts = (np.arange(32, dtype=np.uint64) * np.uint64(2))[None, :]
    while pos < n:
        cnt = min(chunk, (n - pos + 31) // 32)
        z = splitmix64(seed, cnt, k0)
        sym = ((z[:, None] >> shifts) & np.uint64(3)).astype(np.uint8).reshape(-1)
        take = min(len(sym), n - pos)
        out[pos:pos + take] = lut[sym[:take]]
        pos += take
        k0 += cnt
    if newline_tail and n > 0:
        out[n - 1] = 0x0A
    return out


def rand_bytes(n: int, seed: int = SEED_BYTES) -> np.ndarray:
    """G_bytes: each output yields 8 bytes little-endian."""
    out = np.empty(n, dtype=np.uint8)
    chunk = 1 << 22
    pos = 0
    k0 = 0
    while pos < n:
        cnt = min(chunk, (n - pos + 7) // 8)
        z = splitmix64(seed, cnt, k0).astype("<u8")
        b = z.view(np.uint8)
        take = min(len(b), n - pos)
        out[pos:pos + take] = b[:take]
        pos += take
        k0 += cnt
    return out


def tiled(data: np.ndarray, n: int) -> np.ndarray:
    """G_tiled: `data` repeated and truncated to n bytes."""
    reps = (n + len(data) - 1) // len(data)
    return np.tile(data, reps)[:n].copy()


_TWO = ["é", "è", "ü", "ö", "ä", "ñ", "ç", "ß"]
_THREE = ["—", "’", "“", "”", "€", "…", "☃", "日", "本", "語"]
_FOUR = "😀"


def _vocab(seed: int, V: int = 65536):
    words = []
    for k in range(V):
        r = splitmix64(seed ^ ((k * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF), 12)
        ln = 1 + int(r[0] % np.uint64(10))
        w = []
        for j in range(ln):
            x = int(r[1 + j])
            if x % 1024 == 0:
                w.append(_FOUR)
            elif x % 128 == 1:
                w.append(_THREE[(x >> 10) % len(_THREE)])
            elif x % 16 == 2:
                w.append(_TWO[(x >> 10) % len(_TWO)])
            else:
                w.append(chr(ord("a") + (x >> 10) % 26))
        words.append("".join(w).encode("utf-8"))
    return words


def english(n: int, seed: int = SEED_ENGLISH) -> np.ndarray:
    """G_english: Zipf-like (log-uniform rank) words over a 65,536-word
    vocabulary with 2/3/4-byte code points; always valid UTF-8; padded with
    ' ' to exactly n bytes."""
    V = 65536
    words = _vocab(seed, V)
    # word + separator variants flattened so assembly is one gather per chunk
    seps = [b" ", b". ", b"\n"]
    flat = bytearray()
    off = np.zeros((V, 3), dtype=np.int64)
    ln = np.zeros((V, 3), dtype=np.int64)
    for k, w in enumerate(words):
        for s, sep in enumerate(seps):
            off[k, s] = len(flat)
            ln[k, s] = len(w) + len(sep)
            flat += w + sep
    flat = np.frombuffer(bytes(flat), dtype=np.uint8)
    out = np.full(n, 0x20, dtype=np.uint8)
    pos = 0
    widx = 0
    chunk = 1 << 20
    logV = np.log(V)
    while pos < n:
        z = splitmix64(seed, chunk, widx)
        u = (z >> np.uint64(11)).astype(np.float64) / float(1 << 53)
        rank = np.minimum(np.maximum(np.floor(np.exp(u * logV)).astype(np.int64) - 1, 0), V - 1)
        i = np.arange(widx + 1, widx + chunk + 1)
        sep = np.where(i % 80 == 0, 2, np.where(i % 12 == 0, 1, 0))
        l = ln[rank, sep]
        o = off[rank, sep]
        ends = np.cumsum(l)
        fit = int(np.searchsorted(ends, n - pos, side="right"))
        if fit == 0:
            break
        l, o, ends = l[:fit], o[:fit], ends[:fit]
        total = int(ends[-1])
        starts = ends - l
        src = np.repeat(o - starts, l) + np.arange(total, dtype=np.int64)
        out[pos:pos + total] = flat[src]
        pos += total
        widx += fit
        if fit < chunk:
            break
    return out


def fixture(name: str) -> np.ndarray:
    """The reference's own fixtures (tests/AP009048_*.fasta), committed
    verbatim under tests/golden/ because /root/reference is absent on the GPU box."""
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = os.path.join(here, "tests", "golden", name)
    return np.fromfile(p, dtype=np.uint8)
