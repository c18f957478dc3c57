"""Per-phase device times of the build (CUDA events inside the library) for a
few synthetic inputs; diagnostic, prints one JSON line per input."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from suffix_b200 import _lib, gen  # noqa: E402


def run(ctx, name, t, reps=2, lcp=True):
    dev = torch.device("cuda:0")
    d_t = torch.from_numpy(t).to(dev)
    n = len(t)
    d_sa = torch.empty(n, dtype=torch.int32, device=dev)
    d_lcp = torch.empty(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ctx.set_timing(True)
    out = {"input": name, "n": n}
    for r in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.build_dev(d_t.data_ptr(), n, d_sa.data_ptr(), stream)
        torch.cuda.synchronize()
        out["wall_ms_%d" % r] = round((time.perf_counter() - t0) * 1e3, 3)
    out["phases_ms"] = {k: round(v, 3) for k, v in ctx.phase_times()}
    out["stats"] = ctx.stats()
    if lcp:
        ctx.lcp_dev(d_t.data_ptr(), n, d_sa.data_ptr(), d_lcp.data_ptr(), stream)
        torch.cuda.synchronize()
        out["lcp_phases_ms"] = {k: round(v, 3) for k, v in ctx.phase_times()}
    total = sum(out["phases_ms"].values())
    out["sa_MBps"] = round(n / 1e6 / (total / 1e3), 2) if total > 0 else None
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    kinds = ["dna", "dna_nl", "bytes", "english"]
    args = []
    for a in sys.argv[1:]:
        if a.startswith("--kinds="):
            kinds = a.split("=", 1)[1].split(",")
        else:
            args.append(a)
    sizes = [int(x) for x in args] or [1_000_000, 10_000_000, 100_000_000]
    ctx = _lib.Context(0)
    makers = {"dna": lambda n: gen.dna(n), "dna_nl": lambda n: gen.dna(n, newline_tail=True),
              "bytes": lambda n: gen.rand_bytes(n), "english": lambda n: gen.english(n),
              "tiled": lambda n: gen.tiled(gen.fixture("AP009048_100000.fasta"), n)}
    for n in sizes:
        for k in kinds:
            run(ctx, k, makers[k](n))
