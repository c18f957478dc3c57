"""f-1 throughput: a batch of positions() queries on the device against the host mirror.
Prints one JSON line (queries/s of k_positions, of the Python host mirror, of the oracle's C
restatement of src/table.rs:223-259, and the agreement count)."""
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from oracle import oracle
from suffix_b200 import SuffixTable, _lib, gen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
rng = np.random.default_rng(5)
t = gen.dna(n)
dev = torch.device("cuda:0")
ctx = _lib.default_context(0)
d_t = torch.from_numpy(t).to(dev)
d_sa = torch.empty(n, dtype=torch.int32, device=dev)
ts = torch.cuda.Stream()                      # a real stream: handle 0 would mean "the library's own stream"
torch.cuda.set_stream(ts)
stream = ts.cuda_stream
ctx.build_dev(d_t.data_ptr(), n, d_sa.data_ptr(), stream)
torch.cuda.synchronize()
lens = rng.integers(8, 33, nq)
starts = rng.integers(0, n - 40, nq)
miss = rng.random(nq) < 0.2
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
flat = np.empty(int(off[-1]), dtype=np.uint8)
for k in range(nq):
    q = t[starts[k]:starts[k] + lens[k]]
    flat[off[k]:off[k + 1]] = q if not miss[k] else np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, lens[k])]
d_q = torch.from_numpy(flat).to(dev); d_off = torch.from_numpy(off).to(dev)
d_s = torch.zeros(nq, dtype=torch.int32, device=dev); d_e = torch.zeros(nq, dtype=torch.int32, device=dev)
best = None
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ts)
    ctx.positions_dev(d_t.data_ptr(), n, d_sa.data_ptr(), d_q.data_ptr(), d_off.data_ptr(), nq, d_s.data_ptr(), d_e.data_ptr(), stream)
    e1.record(ts); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = ms if best is None or (it > 0 and ms < best) else best
s, e = d_s.cpu().numpy(), d_e.cpu().numpy()
sa = d_sa.cpu().numpy().view(np.uint32)
tab = SuffixTable.from_parts(t.tobytes(), sa)
ks = rng.integers(0, nq, 20000)
t0 = time.perf_counter(); agree = 0
for k in ks.tolist():
    agree += (int(s[k]), int(e[k])) == oracle.positions(t, sa, flat[off[k]:off[k + 1]])
t_or = time.perf_counter() - t0
t0 = time.perf_counter()
for k in ks[:2000].tolist():
    tab.positions(flat[off[k]:off[k + 1]].tobytes())
t_py = time.perf_counter() - t0
print(json.dumps({"n": n, "queries": nq, "query_len": "8..32", "miss_fraction": 0.2,
                  "device_ms": round(best, 3), "device_queries_per_s": round(nq / (best / 1e3)),
                  "oracle_c_queries_per_s_1core": round(len(ks) / t_or), "python_mirror_queries_per_s": round(2000 / t_py),
                  "checked_vs_oracle": int(len(ks)), "agree": int(agree), "hits": int((e > s).sum())}))
