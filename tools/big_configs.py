"""BASELINE configs 3 and 4 (and larger DNA) on one GPU: timing + size-independent
parity properties (permutation, sampled adjacent order, LCP spot checks)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suffix_b200 import _lib, gen  # noqa: E402


def check(t, sa, lcp, samples=20000, seed=1):
    n = len(t)
    seen = np.zeros(n, dtype=np.uint8)
    seen[sa] = 1
    ok_perm = int(seen.sum()) == n
    del seen
    rng = np.random.default_rng(seed)
    bad = 0
    for i in rng.integers(1, n, samples).tolist():
        a, b = int(sa[i - 1]), int(sa[i])
        k = 64
        while True:
            x, y = t[a:a + k].tobytes(), t[b:b + k].tobytes()
            if x != y or a + k >= n or b + k >= n:
                break
            k *= 4
        if not (x < y):
            bad += 1
        if lcp is not None:
            h = int(lcp[i])
            if t[a:a + h].tobytes() != t[b:b + h].tobytes() or not (a + h == n or b + h == n or t[a + h] != t[b + h]):
                bad += 1
    return ok_perm, bad


def run(ctx, name, t, want_lcp=True):
    dev = torch.device("cuda:0")
    n = len(t)
    d_t = torch.from_numpy(t).to(dev)
    d_sa = torch.empty(n, dtype=torch.int32, device=dev)
    d_lcp = torch.empty(n, dtype=torch.int32, device=dev) if want_lcp else None
    s = torch.cuda.current_stream().cuda_stream
    ctx.set_timing(True)
    out = {"config": name, "n": n}
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        ctx.build_dev(d_t.data_ptr(), n, d_sa.data_ptr(), s)
        torch.cuda.synchronize(); out["sa_ms_%d" % rep] = round((time.perf_counter() - t0) * 1e3, 2)
    out["sa_phases_ms"] = {k: round(v, 3) for k, v in ctx.phase_times()}
    out["stats"] = ctx.stats()
    if want_lcp:
        for rep in range(2):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            ctx.lcp_dev(d_t.data_ptr(), n, d_sa.data_ptr(), d_lcp.data_ptr(), s)
            torch.cuda.synchronize(); out["lcp_ms_%d" % rep] = round((time.perf_counter() - t0) * 1e3, 2)
        out["lcp_phases_ms"] = {k: round(v, 3) for k, v in ctx.phase_times()}
    out["sa_MBps"] = round(n / 1e6 / (out["sa_ms_1"] / 1e3), 1)
    sa = d_sa.cpu().numpy().view(np.uint32)
    lcp = d_lcp.cpu().numpy().view(np.uint32) if want_lcp else None
    ok_perm, bad = check(t, sa, lcp, samples=(0 if name.startswith("tiled") else 20000))   # tiled: LCP ~ n/2 makes the sampled compare quadratic
    out["permutation"] = ok_perm
    out["bad_samples"] = bad
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["bytes100m", "english1g", "dna1g"]
    ctx = _lib.Context(0)
    for w in which:
        t0 = time.time()
        if w == "bytes100m":
            t = gen.rand_bytes(100_000_000)
        elif w == "english1g":
            t = gen.english(1_000_000_000)
        elif w == "english100m":
            t = gen.english(100_000_000)
        elif w == "dna1g":
            t = gen.dna(1_000_000_000)
        elif w == "dna3g":
            t = gen.dna(3_000_000_000)
        elif w == "tiled100m":
            t = gen.tiled(gen.fixture("AP009048_100000.fasta"), 100_000_000)
        else:
            raise SystemExit("unknown " + w)
        print("# generated %s in %.1fs" % (w, time.time() - t0), flush=True)
        run(ctx, w, t, want_lcp=(w not in ("dna3g", "tiled100m")))
