"""Input families shared by the CPU and GPU parity tests: the reference's own
KAT strings plus the adversarial shapes of SURVEY.md section 4 (iv)."""
import json
import os

import numpy as np

from suffix_b200 import gen

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def kat():
    with open(os.path.join(GOLDEN, "kat.json")) as f:
        return json.load(f)


def full_size():
    """SHA-256 goldens of BASELINE configs 2 and 3 at full size (make_full_size.py)."""
    with open(os.path.join(GOLDEN, "full_size.json")) as f:
        return json.load(f)


def fib_word(n):
    a, b = b"a", b"ab"
    while len(b) < n:
        a, b = b, b + a
    return b[:n]


def adversarial(scale=1):
    """(name, bytes) pairs; `scale` multiplies the base lengths."""
    s = scale
    rng = np.random.default_rng(12345)
    fx = gen.fixture("AP009048_10000.fasta").tobytes()
    out = [
        ("a^n", b"a" * (5000 * s)),
        ("(ab)^n", b"ab" * (5000 * s)),
        ("(ba)^n", b"ba" * (2500 * s)),
        ("fib", fib_word(10946 * s)),
        ("fixture_x3", fx * 3),
        ("ramp_desc_x10", bytes(range(255, -1, -1)) * 10),
        ("ramp_asc_x10", bytes(range(256)) * 10),
        ("binary", rng.integers(0, 2, 20000 * s, dtype=np.uint8).tobytes()),
        ("a^k b repeated", (b"a" * 37 + b"b") * (300 * s)),
        ("b a^k repeated", (b"b" + b"a" * 37) * (300 * s)),
        ("zeros", bytes(3000 * s)),
        ("ff", b"\xff" * (3000 * s)),
        ("zeros_then_ff", bytes(1500 * s) + b"\xff" * (1500 * s)),
        ("ff_then_zeros", b"\xff" * (1500 * s) + bytes(1500 * s)),
        ("dna_small", gen.dna(20000 * s).tobytes()),
        ("dna_nl", gen.dna(20001 * s, newline_tail=True).tobytes()),
        ("bytes_small", gen.rand_bytes(20000 * s).tobytes()),
        ("runs", bytes(np.repeat(rng.integers(0, 4, 2000 * s, dtype=np.uint8) + 65,
                                 rng.integers(1, 40, 2000 * s)).tolist())),
        ("long_runs", bytes(np.repeat(rng.integers(0, 3, 40, dtype=np.uint8) + 97,
                                      rng.integers(1, 3000, 40)).tolist())),
        ("english_small", gen.english(30000 * s).tobytes()),
        ("bacbacbc", b"bacbacbc"), ("bababaa", b"bababaa"),
        ("two_distinct_desc", b"ba"), ("three", b"cab"), ("aab", b"aab"), ("baa", b"baa"),
    ]
    return out
