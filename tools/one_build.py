"""Two device-resident SA+LCP builds (1 warm-up + 1) of a G_dna text; the
short command ncu wraps for --set full captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from suffix_b200 import _lib, gen

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "dna"
t = gen.dna(n) if kind == "dna" else gen.rand_bytes(n)
dev = torch.device("cuda:0")
d_t = torch.from_numpy(t).to(dev)
d_sa = torch.empty(n, dtype=torch.int32, device=dev)
d_lcp = torch.empty(n, dtype=torch.int32, device=dev)
ctx = _lib.Context(0)
s = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    ctx.build_dev(d_t.data_ptr(), n, d_sa.data_ptr(), s)
    ctx.lcp_dev(d_t.data_ptr(), n, d_sa.data_ptr(), d_lcp.data_ptr(), s)
torch.cuda.synchronize()
print("launches per build+lcp:", ctx.stats())
