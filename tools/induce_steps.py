import sys, os
sys.path.insert(0, os.getcwd())
import torch
from suffix_b200 import _lib, gen
ctx = _lib.Context(0)
for name, t in (("dna100m", gen.dna(100_000_000)), ("bytes100m", gen.rand_bytes(100_000_000)), ("english100m", gen.english(100_000_000))):
    d_t = torch.from_numpy(t).cuda(); d_sa = torch.empty(len(t), dtype=torch.int32, device="cuda")
    ctx.build_dev(d_t.data_ptr(), len(t), d_sa.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    e = ctx.debug_fetch(5, 16)
    print(name, "L pass: big steps %d, small steps %d, big tiles %d | S pass: big %d small %d tiles %d" % tuple(int(x) for x in e[4:10]))
