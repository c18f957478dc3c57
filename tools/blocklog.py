"""Per-block phase times of ONE big step of the persistent induce kernel
(B200SA_STEPLOG=1 B200SA_BLOCKLOG=<pass>:<big step index>): count end, first grid sync passed,
scatter end, second grid sync passed, relative to the earliest block start; with the SM id."""
import os, sys
os.environ["B200SA_STEPLOG"] = "1"
os.environ.setdefault("B200SA_BLOCKLOG", sys.argv[1] if len(sys.argv) > 1 else "1:0")
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from suffix_b200 import _lib, gen
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000_000
t = gen.dna(n)
ctx = _lib.Context(0)
d_t = torch.from_numpy(t).cuda(); d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.build_dev(d_t.data_ptr(), n, d_sa.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
raw = ctx.debug_fetch(7, 16384).view(np.uint64)
rec = raw[4096:4096 + 6 * 444].reshape(-1, 6).astype(np.int64)
rec = rec[rec[:, 0] > 0]
t0 = rec[:, 0].min()
print("step", os.environ["B200SA_BLOCKLOG"], "blocks", len(rec))
d = (rec[:, :5] - t0) / 1e3
cnt = d[:, 1] - d[:, 0]; sc = d[:, 3] - d[:, 2]
print("count   us: min %.1f med %.1f max %.1f" % (cnt.min(), np.median(cnt), cnt.max()))
print("scatter us: min %.1f med %.1f max %.1f" % (sc.min(), np.median(sc), sc.max()))
print("sync1 passed at %.1f..%.1f, sync2 passed at %.1f..%.1f" % (d[:, 2].min(), d[:, 2].max(), d[:, 4].min(), d[:, 4].max()))
print(" bid  sm  start  cnt_end  sync1  scat_end  (count, scatter)")
for b in range(len(rec)):
    if b % 12 == 0 or sc[b] > np.percentile(sc, 97):
        print("%4d %3d %6.1f %7.1f %7.1f %8.1f   %6.1f %6.1f" % (b, rec[b, 5], d[b, 0], d[b, 1], d[b, 2], d[b, 3], cnt[b], sc[b]))
# by SM: mean scatter time
sms = rec[:, 5]
by = {}
for b in range(len(rec)):
    by.setdefault(int(sms[b]), []).append(sc[b])
m = sorted((np.mean(v), k, len(v)) for k, v in by.items())
print("slowest SMs (mean scatter us, sm, blocks):", [(round(a, 1), k, c) for a, k, c in m[-8:]])
print("fastest SMs:", [(round(a, 1), k, c) for a, k, c in m[:8]])
