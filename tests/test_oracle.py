"""Pins the CPU oracle (oracle/sais_oracle.c) against the reference's own
golden vectors: tests/tests.rs KATs, doc tests, fixture SHA-256s, and the
reference's own test oracle (naive suffix sort)."""
import hashlib

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import oracle
from suffix_b200 import gen
from tests import families

KAT = families.kat()


@pytest.mark.parametrize("case", KAT["kat"], ids=lambda c: repr(c["text"])[:24])
def test_kat_sa_lcp(case):
    t = case["text"].encode("utf-8")
    sa = oracle.sais(t)
    assert sa.tolist() == case["sa"]
    assert oracle.naive_sa(t).tolist() == case["sa"]       # tests/tests.rs:22-70
    assert oracle.lcp_quadratic(t, sa).tolist() == case["lcp"]
    assert oracle.lcp_kasai(t, sa).tolist() == case["lcp"]


@pytest.mark.parametrize("text,want", list(KAT["types"].items()))
def test_kat_types(text, want):
    ty = oracle.types(text.encode("utf-8"))
    assert "".join("SLV"[x] for x in ty) == want


@pytest.mark.parametrize("case", KAT["positions"], ids=lambda c: repr((c["text"], c["query"]))[:40])
def test_kat_positions(case):
    t = case["text"].encode("utf-8")
    q = case["query"].encode("utf-8")
    sa = oracle.sais(t)
    s, e = oracle.positions(t, sa, q)
    assert sa[s:e].tolist() == case["positions"]
    hit = oracle.any_position(t, sa, q)
    if case["positions"]:
        assert hit in case["positions"]
    else:
        assert hit is None


@pytest.mark.parametrize("name", sorted(KAT["fixtures"]))
def test_fixture_sha256(name):
    info = KAT["fixtures"][name]
    t = gen.fixture(name)
    assert len(t) == info["n"]
    sa = oracle.sais(t)
    assert sa[:8].tolist() == info["sa_head"]
    assert hashlib.sha256(sa.astype("<u4").tobytes()).hexdigest() == info["sa_sha256"]
    lcp = oracle.lcp_quadratic(t, sa)
    assert lcp[:8].tolist() == info["lcp_head"]
    assert hashlib.sha256(lcp.astype("<u4").tobytes()).hexdigest() == info["lcp_sha256"]
    assert np.array_equal(oracle.lcp_kasai(t, sa), lcp)
    if info["n"] <= 20000:
        assert np.array_equal(oracle.naive_sa(t), sa)


@pytest.mark.parametrize("name,data", families.adversarial(), ids=lambda x: x if isinstance(x, str) else "")
def test_adversarial_sais_equals_naive(name, data):
    sa = oracle.sais(data)
    assert np.array_equal(sa, oracle.naive_sa(data)), name
    assert np.array_equal(oracle.lcp_kasai(data, sa), oracle.lcp_quadratic(data, sa))


# tests/tests.rs:73-96 prop_naive_equals_sais / prop_matches_naive
@settings(max_examples=500, deadline=None)
@given(st.text(max_size=200))
def test_prop_text(s):
    t = s.encode("utf-8")
    sa = oracle.sais(t)
    assert len(sa) == len(t)                                # prop_length :215-221
    assert np.array_equal(sa, oracle.naive_sa(t))


@settings(max_examples=500, deadline=None)
@given(st.binary(max_size=300))
def test_prop_binary(t):
    assert np.array_equal(oracle.sais(t), oracle.naive_sa(t))


@settings(max_examples=300, deadline=None)
@given(st.text(alphabet="ab", max_size=64), st.text(alphabet="ab", min_size=1, max_size=3))
def test_prop_positions(s, q):
    # tests/tests.rs:223-243 prop_contains / prop_positions
    t, qb = s.encode(), q.encode()
    sa = oracle.sais(t)
    a, b = oracle.positions(t, sa, qb)
    want = [i for i in range(len(t)) if t.startswith(qb, i)]
    assert sorted(sa[a:b].tolist()) == want
    assert (oracle.any_position(t, sa, qb) is not None) == bool(want)


def test_generators_shapes():
    d = gen.dna(1000)
    assert set(d.tolist()) <= set(b"ACGT") and len(d) == 1000
    assert np.array_equal(gen.dna(1000)[:777], gen.dna(777))
    b = gen.rand_bytes(4099)
    assert len(b) == 4099 and np.array_equal(b[:4096], gen.rand_bytes(4096))
    e = gen.english(5000)
    e.tobytes().decode("utf-8")
    assert len(e) == 5000


@settings(max_examples=300, deadline=None)
@given(st.binary(min_size=2, max_size=200), st.data())
def test_verify_sa(t, data):
    """oracle.verify_sa accepts exactly the suffix array (used for BASELINE config 4,
    where sais() itself is too slow to be a per-test oracle)."""
    sa = oracle.naive_sa(t)
    assert oracle.verify_sa(t, sa) == 0
    i = data.draw(st.integers(0, len(t) - 2))
    bad = sa.copy()
    bad[[i, i + 1]] = bad[[i + 1, i]]          # any transposition breaks the (unique) order
    assert oracle.verify_sa(t, bad) != 0
    dup = sa.copy()
    dup[i] = dup[i + 1]
    assert oracle.verify_sa(t, dup) != 0


@settings(max_examples=300, deadline=None)
@given(st.binary(max_size=120), st.binary(min_size=1, max_size=4))
def test_host_mirror_positions(t, q):
    """The Python mirror's positions()/any_position() (host binary search over text +
    table, src/table.rs:223-293) against the oracle's restatement, on tables built by
    the oracle (from_parts: no GPU involved)."""
    from suffix_b200.table import SuffixTable
    sa = oracle.naive_sa(t)
    tab = SuffixTable.from_parts(t, sa)
    s, e = oracle.positions(t, sa, q)
    assert tab.positions(q).tolist() == sa[s:e].tolist()
    want = [i for i in range(len(t)) if t.startswith(q, i)]
    assert sorted(tab.positions(q).tolist()) == want
    hit = tab.any_position(q)
    assert (hit in want) if want else hit is None


def test_full_size_goldens_present():
    fs = families.full_size()
    assert set(fs) == {"G_dna_100MB", "G_bytes_100MB"} and all(len(v["sa_sha256"]) == 64 for v in fs.values())
