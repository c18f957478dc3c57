"""Step timeline of the persistent induce kernels (B200SA_STEPLOG=1): per step the list
length and the time until the next step starts (block 0, globaltimer)."""
import os, sys
os.environ["B200SA_STEPLOG"] = "1"
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from suffix_b200 import _lib, gen
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
t = gen.dna(n)
ctx = _lib.Context(0)
d_t = torch.from_numpy(t).cuda(); d_sa = torch.empty(n, dtype=torch.int32, device="cuda")
for _ in range(2):
    ctx.build_dev(d_t.data_ptr(), n, d_sa.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
raw = ctx.debug_fetch(7, 8192).view(np.uint64)
k = int(raw[0])
rec = raw[1:1 + 2 * k].reshape(-1, 2)
t0 = int(rec[0, 0])
print("records", k)
names = {1: "count", 2: "sync1", 3: "matrix", 4: "scatter"}
i = 0
rows = []
while i < k - 1:
    if int(rec[i, 1]) >> 28 == 0xF:
        i += 1
        continue
    j = i + 1
    marks = []
    while j < k and int(rec[j, 1]) >> 28 == 0xF:
        marks.append((int(rec[j, 1]) & 0xff, int(rec[j, 0])))
        j += 1
    if j >= k:
        break
    t_prev = int(rec[i, 0])
    parts = []
    for code, tm in marks:
        parts.append("%s=%.1f" % (names.get(code, str(code)), (tm - t_prev) / 1e3))
        t_prev = tm
    parts.append("sync2+peek=%.1f" % ((int(rec[j, 0]) - t_prev) / 1e3))
    ln = int(rec[i, 1])
    print("%4d t=%9.1f us len=%10d tiles=%6d dur=%8.1f us  %s" % (i, (int(rec[i, 0]) - t0) / 1e3, ln, (ln + 2047) // 2048, (int(rec[j, 0]) - int(rec[i, 0])) / 1e3, " ".join(parts) if marks else "(small)"))
    i = j
sys.exit(0)
for i in range(k - 1):
    dt = (int(rec[i + 1, 0]) - int(rec[i, 0])) / 1e3
    print("%4d  t=%9.1f us  len=%10d  tiles=%6d  dur=%8.1f us%s" % (i, (int(rec[i, 0]) - t0) / 1e3, int(rec[i, 1]), (int(rec[i, 1]) + 2047) // 2048, dt, "  (small)" if int(rec[i, 1]) <= 2048 else ""))
