"""Per-phase times on FASTA-shaped inputs (4-bit packed path): DNA + trailing newline
(the reference fixtures' shape, sigma = 5) and wrapped FASTA with poly-N runs."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from suffix_b200 import _lib, gen
from tools.phase_times import run

def wrapped_fasta(n):
    d = gen.dna(n)
    out = d.copy()
    out[70::71] = 10                      # newline every 71st byte
    rng = np.random.default_rng(5)
    for _ in range(200):                  # poly-N runs, a few long
        a = int(rng.integers(0, n - 400000)); L = int(rng.choice([50, 500, 5000, 300000]))
        out[a:a + L] = ord("N")
    return out

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
    ctx = _lib.Context(0)
    run(ctx, "dna_newline_tail (sigma=5)", gen.dna(n, newline_tail=True))
    run(ctx, "wrapped_fasta_polyN (sigma=6)", wrapped_fasta(n))
    run(ctx, "fixture100k", gen.fixture("AP009048_100000.fasta"))
