"""Rows (f) of SURVEY.md section 8 ("next"): batched positions() on the device, generalized
suffix array, LCP-interval tree -- each against the oracle / a brute-force statement."""
import numpy as np
import pytest

from oracle import oracle
from suffix_b200 import GeneralizedSuffixTable, SuffixTable, _lib, gen

pytestmark = pytest.mark.gpu


def _queries(t, rng, nq):
    tb = t.tobytes()
    qs = []
    for _ in range(nq):
        k = rng.integers(0, 10)
        ln = int(rng.integers(1, 24))
        if k < 7:                                   # a substring of the text (hits)
            s = int(rng.integers(0, len(tb) - ln))
            qs.append(tb[s:s + ln])
        elif k < 9:                                 # random bytes over the text's alphabet (mostly misses when long)
            qs.append(bytes(rng.choice(np.frombuffer(tb[:4096], dtype=np.uint8), ln).tolist()))
        else:                                       # bytes that do not occur
            qs.append(bytes([255] * ln))
    return qs


@pytest.mark.parametrize("maker", ["dna", "english"])
def test_positions_dev_large_batch(maker):
    """f-1: 100,000 queries in one launch, every answer against oracle.positions
    (src/table.rs:223-259 restated in C)."""
    import torch
    rng = np.random.default_rng(17)
    t = gen.dna(1_000_000) if maker == "dna" else gen.english(1_000_000)
    tab = SuffixTable(t.tobytes())
    sa = tab.table()
    qs = _queries(t, rng, 100_000)
    flat = np.frombuffer(b"".join(qs), dtype=np.uint8)
    off = np.cumsum([0] + [len(q) for q in qs]).astype(np.int64)
    dev = torch.device("cuda:0")
    d_t = torch.from_numpy(t.copy()).to(dev)
    d_sa = torch.from_numpy(sa.astype(np.int64)).to(dev).to(torch.int32)
    d_q = torch.from_numpy(flat.copy()).to(dev)
    d_off = torch.from_numpy(off).to(dev)
    d_s = torch.zeros(len(qs), dtype=torch.int32, device=dev)
    d_e = torch.zeros(len(qs), dtype=torch.int32, device=dev)
    ctx = _lib.default_context(0)
    ctx.positions_dev(d_t.data_ptr(), len(t), d_sa.data_ptr(), d_q.data_ptr(), d_off.data_ptr(), len(qs),
                      d_s.data_ptr(), d_e.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    s, e = d_s.cpu().numpy(), d_e.cpu().numpy()
    hits = 0
    for k, q in enumerate(qs):
        assert (int(s[k]), int(e[k])) == oracle.positions(t, sa, q), (k, q)
        hits += int(e[k] > s[k])
    assert 0.5 * len(qs) < hits < len(qs)


def test_generalized_suffix_table():
    """f-3: documents joined by a separator + document lookup (reference README.md:60-74)."""
    rng = np.random.default_rng(3)
    docs = [gen.dna(int(rng.integers(1, 3000)), seed=100 + k).tobytes() for k in range(40)] + [b"", b"ACGT", b"ACGT"]
    g = GeneralizedSuffixTable(docs)
    for q in (b"ACGT", b"GATTACA", b"T", b"ACGTAC", b"N", docs[5][10:40]):
        got = sorted(map(tuple, g.positions(q).tolist()))
        want = sorted((d, i) for d, doc in enumerate(docs) for i in range(len(doc)) if doc.startswith(q, i))
        assert got == want, q
        assert g.contains(q) == bool(want)
    with pytest.raises(ValueError):
        GeneralizedSuffixTable([b"a\x00b"])
    with pytest.raises(ValueError):
        g.positions(b"A\x00C")
    # locate(): every position of the concatenation, separators included (they close their document)
    n = len(g.table())
    loc = g.locate(np.arange(n, dtype=np.uint32))
    starts = g.doc_starts()
    want_doc = np.searchsorted(starts, np.arange(n), side="right") - 1
    assert np.array_equal(loc[:, 0], want_doc) and np.array_equal(loc[:, 1], np.arange(n) - starts[want_doc])


def _ansv_cpu(lcp):
    n = len(lcp)
    psv = np.full(n, 0xFFFFFFFF, dtype=np.uint32)
    nsv = np.full(n, n, dtype=np.uint32)
    st = []
    for i in range(n):
        while st and lcp[st[-1]] >= lcp[i]:
            st.pop()
        if st:
            psv[i] = st[-1]
        st.append(i)
    st = []
    for i in range(n - 1, -1, -1):
        while st and lcp[st[-1]] >= lcp[i]:
            st.pop()
        if st:
            nsv[i] = st[-1]
        st.append(i)
    return psv, nsv


@pytest.mark.parametrize("text", [b"banana", b"mississippi", b"abracadabra" * 50, b"a" * 3000, gen.dna(200_000).tobytes(),
                                  gen.english(150_000).tobytes(), gen.fixture("AP009048_100000.fasta").tobytes()],
                         ids=["banana", "mississippi", "abra", "a^n", "dna", "english", "fixture"])
def test_lcp_intervals(text):
    """f-4: all-nearest-smaller-values of the LCP array = the LCP-interval tree.  The number of
    distinct intervals equals the number of internal nodes (root excluded) of the suffix tree the
    reference builds from the same SA + LCP (suffix_tree/src/lib.rs:392-505)."""
    st_ = SuffixTable(text)
    lcp, psv, nsv = st_.lcp_intervals()
    want_p, want_n = _ansv_cpu(lcp)
    assert np.array_equal(psv, want_p) and np.array_equal(nsv, want_n)
    n = len(lcp)
    nodes = {(int(psv[i]), int(nsv[i]), int(lcp[i])) for i in range(1, n) if lcp[i] > 0}
    # serial construction of the same node set: a stack of open intervals over the LCP array
    stack, serial = [], set()
    for i in range(1, n + 1):
        cur = int(lcp[i]) if i < n else 0
        left = i - 1
        while stack and stack[-1][0] > cur:
            d, l = stack.pop()
            serial.add((l, i, d))
            left = l
        if cur > 0 and (not stack or stack[-1][0] < cur):
            stack.append((cur, left))
    # an interval [l, r) of depth d is owned by every boundary i in it with lcp[i] == d: psv < l.. map to (l-ish)
    assert len(nodes) == len(serial)
