#!/bin/bash
# compute-sanitizer on small builds (memcheck, racecheck, initcheck-lite)
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from suffix_b200 import _lib, gen
from oracle import oracle
ctx = _lib.Context(0)
for name, t in [("dna", gen.dna(300_000)), ("bytes", gen.rand_bytes(200_000)), ("english", gen.english(150_000)),
                ("runs", np.concatenate([gen.dna(5000), np.full(20000, 78, np.uint8), gen.dna(5000), np.full(300, 65, np.uint8)])),
                ("periodic", np.tile(np.frombuffer(b"abcab", np.uint8), 20000))]:
    sa, lcp = ctx.build_lcp(t)
    want = oracle.sais(t)
    assert np.array_equal(sa, want), name
    assert np.array_equal(lcp, oracle.lcp_kasai(t, want)), name
    os.environ["B200SA_LCP_LINEAR"] = "1"
    assert np.array_equal(ctx.lcp(t, sa), lcp), name
    del os.environ["B200SA_LCP_LINEAR"]
print("sanitize workload ok")
PY
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/san_$tool.log 2>&1
  echo "$tool exit $?"; tail -4 gpurun_out/san_$tool.log
done
