"""Writes tests/golden/full_size.json: SHA-256 of the suffix array and the LCP
array (packed little-endian u32) of BASELINE configs 2 and 3 at full size
(100 MB G_dna, 100 MB G_bytes; SURVEY.md Appendix C generators), computed with
the CPU oracle (oracle_sais = restated src/table.rs:388-574, oracle_lcp_quadratic
= src/table.rs:348-361).  The oracle itself is pinned to the reference's KATs by
tests/test_oracle.py.  Runs in about two minutes on one core:

    python tests/golden/make_full_size.py
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle          # noqa: E402
from suffix_b200 import gen        # noqa: E402

N = 100_000_000
out = {}
for name, text in (("G_dna_100MB", gen.dna(N)), ("G_bytes_100MB", gen.rand_bytes(N))):
    t0 = time.time()
    sa = oracle.sais(text)
    lcp = oracle.lcp_quadratic(text, sa)
    out[name] = {
        "n": N,
        "text_sha256": hashlib.sha256(text.tobytes()).hexdigest(),
        "sa_sha256": hashlib.sha256(sa.astype("<u4").tobytes()).hexdigest(),
        "lcp_sha256": hashlib.sha256(lcp.astype("<u4").tobytes()).hexdigest(),
        "sa_head": [int(x) for x in sa[:8]], "lcp_head": [int(x) for x in lcp[:8]],
        "lcp_max": int(lcp.max()), "lcp_sum": int(lcp.astype(np.uint64).sum()),
    }
    print(name, out[name], "%.1f s" % (time.time() - t0), flush=True)
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "full_size.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
    f.write("\n")
