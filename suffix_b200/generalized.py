"""Generalized suffix array over several documents.

The reference does not implement one; its README (README.md:60-74) describes the recipe this
class follows: append the documents with a separator byte that occurs in none of them,
remember where each document starts, build ONE `SuffixTable` over the giant string and map
match positions back to documents with a binary search (on the device for batches:
b200sa_doc_ids_dev).  A query that does not contain the separator can never match across a
document boundary, so `positions()` needs no filtering.
"""
import numpy as np

from . import _lib
from .table import SuffixTable, _as_bytes


class GeneralizedSuffixTable:
    def __init__(self, docs, sep: bytes = b"\x00", device: int = 0):
        self._docs = [_as_bytes(d) for d in docs]
        assert len(sep) == 1, "the separator is one byte"
        for d in self._docs:
            if sep in d:
                raise ValueError("separator byte occurs in a document")
        self._sep = sep
        self._device = device
        starts, pos = [], 0
        for d in self._docs:
            starts.append(pos)
            pos += len(d) + 1
        self._starts = np.asarray(starts, dtype=np.uint32)
        self._table = SuffixTable(sep.join(self._docs) + (sep if self._docs else b""), device=device)

    def table(self) -> SuffixTable:
        return self._table

    def doc_starts(self) -> np.ndarray:
        return self._starts

    def locate(self, positions) -> np.ndarray:
        """(document, offset) for each text position of the concatenation (device batch)."""
        import torch
        p = np.ascontiguousarray(positions, dtype=np.uint32)
        if len(p) == 0 or len(self._docs) == 0:
            return np.zeros((0, 2), dtype=np.uint32)
        dev = torch.device("cuda", self._device)
        d_p = torch.from_numpy(p.astype(np.int64)).to(dev).to(torch.int32)
        d_s = torch.from_numpy(self._starts.astype(np.int64)).to(dev).to(torch.int32)
        d_doc = torch.empty(len(p), dtype=torch.int32, device=dev)
        d_off = torch.empty(len(p), dtype=torch.int32, device=dev)
        ctx = _lib.default_context(self._device)
        ctx.doc_ids_dev(d_p.data_ptr(), len(p), d_s.data_ptr(), len(self._starts), d_doc.data_ptr(), d_off.data_ptr(),
                        torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        return np.stack([d_doc.cpu().numpy().view(np.uint32), d_off.cpu().numpy().view(np.uint32)], axis=1)

    def positions(self, query) -> np.ndarray:
        """All (document, offset) pairs where `query` occurs (SA order of the concatenation)."""
        q = _as_bytes(query)
        if self._sep in q:
            raise ValueError("query contains the separator byte")
        return self.locate(self._table.positions(q))

    def contains(self, query) -> bool:
        return self._table.contains(query)
