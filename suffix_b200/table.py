"""Host-side mirror of the reference's `SuffixTable` (src/table.rs:54-294).

Construction (`SuffixTable(text)` == `SuffixTable::new`, src/table.rs:78-85)
and `lcp_lens` (src/table.rs:130-138) call the CUDA library through the
C-ABI; the bookkeeping and the O(m log n) queries stay on the host exactly
like the reference (src/table.rs:197-293).
"""
import threading

import numpy as np

from . import _lib

_lock = threading.Lock()


def _as_bytes(text) -> bytes:
    if isinstance(text, str):
        return text.encode("utf-8")          # byte-level, like `Utf8` (src/table.rs:778-800)
    if isinstance(text, (bytes, bytearray, memoryview)):
        return bytes(text)
    return np.ascontiguousarray(text, dtype=np.uint8).tobytes()


class SuffixTable:
    """A sequence of lexicographically sorted suffixes (u32 byte offsets)."""

    def __init__(self, text, *, device: int = 0, _table=None):
        """SuffixTable::new (src/table.rs:78-85): O(n) construction on the GPU.
        Raises if the text exceeds B200SA_MAX_N = 2^32-4096 bytes (the reference panics above 2^32-1, :380)."""
        self._text = _as_bytes(text)
        self._device = device
        if _table is not None:
            self._table = _table
            return
        if len(self._text) > 0xFFFFF000:   # B200SA_MAX_N
            raise OverflowError("text longer than 2^32-4096 bytes (B200SA_MAX_N)")
        t = np.frombuffer(self._text, dtype=np.uint8)
        with _lock:                           # default context is not thread-safe
            ctx = _lib.default_context(device)
            self._table = ctx.build(t)
            self._stats = ctx.stats()

    # src/table.rs:111-119
    @classmethod
    def from_parts(cls, text, table) -> "SuffixTable":
        tb = _as_bytes(text)
        table = np.ascontiguousarray(table, dtype=np.uint32)
        assert len(tb) == len(table), "text and table lengths differ"   # assert_eq!, :117
        return cls(tb, _table=table)

    # src/table.rs:125-127
    def into_parts(self):
        return self._text, self._table

    # src/table.rs:130-138 (semantics of lcp_lens_quadratic, :348-361)
    def lcp_lens(self) -> np.ndarray:
        t = np.frombuffer(self._text, dtype=np.uint8)
        with _lock:
            return _lib.default_context(self._device).lcp(t, self._table)

    def table(self) -> np.ndarray:            # :142-144
        return self._table

    def text(self) -> bytes:                  # :148-150 (bytes; the reference returns &str)
        return self._text

    def __len__(self) -> int:                 # :156-158
        return len(self._table)

    def len(self) -> int:
        return len(self._table)

    def is_empty(self) -> bool:               # :162-164
        return len(self._table) == 0

    def suffix(self, i: int) -> str:          # :168-170
        return self._text[int(self._table[i]):].decode("utf-8")

    def suffix_bytes(self, i: int) -> bytes:  # :174-176
        return self._text[int(self._table[i]):]

    def __eq__(self, other):                  # derived PartialEq, :54
        return isinstance(other, SuffixTable) and self._text == other._text and \
            np.array_equal(self._table, other._table)

    # ---- queries (host side, like the reference)
    def contains(self, query) -> bool:        # :197-199
        return self.any_position(query) is not None

    def positions(self, query) -> np.ndarray:
        """src/table.rs:223-259: sub-slice of the table (SA order, unsorted)."""
        text, q = self._text, _as_bytes(query)
        n, tab = len(text), self._table
        empty = tab[0:0]
        if n == 0 or len(q) == 0:
            return empty
        m = len(q)
        # bounded heads instead of whole suffixes: `q <= text[s..]` <=> `q <= text[s..s+m]`
        # (O(m) per probe, no copy of the suffix -- like the reference's slice compares)
        h0 = text[int(tab[0]):int(tab[0]) + m]
        hl = text[int(tab[n - 1]):int(tab[n - 1]) + m]
        if q < h0 or q > hl:                  # :228-235 (q < s0 && !s0.starts_with(q)) || q > last
            return empty
        lo, hi = 0, n                         # binary_search, :900-914
        while lo < hi:
            mid = (lo + hi) // 2
            s = int(tab[mid])
            if q <= text[s:s + m]:
                hi = mid
            else:
                lo = mid + 1
        start = lo
        lo, hi = 0, n - start
        while lo < hi:
            mid = (lo + hi) // 2
            if not text.startswith(q, int(tab[start + mid])):
                hi = mid
            else:
                lo = mid + 1
        end = start + lo
        return empty if start > end else tab[start:end]

    def any_position(self, query):
        """src/table.rs:279-293: some position of `query`, or None."""
        text, q = self._text, _as_bytes(query)
        if len(q) == 0:
            return None
        tab = self._table
        lo, hi = 0, len(tab)
        m = len(q)
        while lo < hi:
            mid = (lo + hi) // 2
            s = int(tab[mid])
            head = text[s:s + m]
            if head == q:
                return s
            if head < q:
                lo = mid + 1
            else:
                hi = mid
        return None

    # ---- LCP-interval tree (the node set of the reference's suffix tree,
    # suffix_tree/src/lib.rs:392-505): (psv, nsv) per rank from b200sa_lcp_intervals_dev
    def lcp_intervals(self, lcp=None):
        """Returns (lcp, psv, nsv): the internal node owning the boundary before rank i is the
        interval [psv[i], nsv[i]) of string depth lcp[i] (psv == 0xFFFFFFFF: none to the left)."""
        import torch
        lcp = self.lcp_lens() if lcp is None else np.ascontiguousarray(lcp, dtype=np.uint32)
        n = len(lcp)
        dev = torch.device("cuda", self._device)
        d_l = torch.from_numpy(lcp.astype(np.int64)).to(dev).to(torch.int32)
        d_p = torch.empty(n, dtype=torch.int32, device=dev)
        d_n = torch.empty(n, dtype=torch.int32, device=dev)
        with _lock:
            _lib.default_context(self._device).lcp_intervals_dev(d_l.data_ptr(), n, d_p.data_ptr(), d_n.data_ptr(),
                                                                 torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return lcp, d_p.cpu().numpy().view(np.uint32), d_n.cpu().numpy().view(np.uint32)

    # ---- persistence: the only "wire format" the reference API implies is
    # from_parts/into_parts (src/table.rs:111-127): text bytes + little-endian u32 table.
    def save(self, path: str) -> None:
        """Writes `<path>.text` (raw bytes) and `<path>.sa` (raw little-endian u32)."""
        with open(path + ".text", "wb") as f:
            f.write(self._text)
        self._table.astype("<u4").tofile(path + ".sa")

    @classmethod
    def load(cls, path: str, mmap: bool = True) -> "SuffixTable":
        """from_parts over files written by save(); the table is memory-mapped by default."""
        with open(path + ".text", "rb") as f:
            text = f.read()
        table = np.memmap(path + ".sa", dtype="<u4", mode="r") if mmap else np.fromfile(path + ".sa", dtype="<u4")
        assert len(text) == len(table), "text and table lengths differ"
        return cls(text, _table=table)

    def last_stats(self) -> dict:
        return getattr(self, "_stats", {})

    def __repr__(self):
        return "SuffixTable(n=%d)" % len(self._table)
