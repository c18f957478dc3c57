"""Sharded classification: host coordination under gloo with a numpy
engine (CPU, world_size 2 and 3) and the CUDA kernels' halo logic on one GPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle
from suffix_b200 import gen, sharded
from tests import families
from tests.shard_ref import NumpyShardEngine


def _expected(t):
    ty = oracle.types(t)
    S = (ty != 1).astype(np.uint8)
    lms = (ty == 2).astype(np.uint8)
    hist = np.zeros(768, dtype=np.uint64)
    np.add.at(hist, t.astype(np.int64) + 256 * np.array([1, 0, 2])[ty], 1)
    return S, lms, hist


def _unpack(words, n):
    return np.unpackbits(np.asarray(words).view(np.uint8), bitorder="little")[:n]


def _cuts(n, parts, rng):
    c = sorted(rng.choice(np.arange(1, n), size=parts - 1, replace=False).tolist()) if n > parts else list(range(1, parts))
    return [0] + c + [n]


def _worker(rank, world, port, texts, cutsets, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = NumpyShardEngine()
    out = []
    for t, cuts in zip(texts, cutsets):
        shard = t[cuts[rank]:cuts[rank + 1]]
        r = sharded.classify_sharded(eng, shard, dist=dist)
        out.append((r.lo, np.asarray(r.stype_words), np.asarray(r.lms_words), np.asarray(r.lmspos_local),
                    r.m_offset, r.m_total, r.hist_global, len(shard)))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_gloo(world):
    rng = np.random.default_rng(5 + world)
    texts = [np.frombuffer(d, dtype=np.uint8) for _, d in families.adversarial() if len(d) >= 8][:18]
    texts += [gen.dna(50_000), np.full(5000, 65, dtype=np.uint8),
              np.concatenate([np.full(3000, 66, dtype=np.uint8), np.full(3000, 65, dtype=np.uint8)])]
    cutsets = [_cuts(len(t), world, rng) for t in texts]
    cutsets[-2] = [0] + [len(texts[-2]) * k // world for k in range(1, world)] + [len(texts[-2])]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, texts, cutsets, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for k, t in enumerate(texts):
        S, lms, hist = _expected(t)
        pos_all = []
        for r in range(world):
            lo, sw, lw, pos, moff, mtot, hg, n_loc = res[r][k]
            assert lo == cutsets[k][r]
            assert np.array_equal(_unpack(sw, n_loc), S[lo:lo + n_loc]), (k, r)
            assert np.array_equal(_unpack(lw, n_loc), lms[lo:lo + n_loc]), (k, r)
            assert np.array_equal(hg, hist)
            assert moff == len(pos_all) and mtot == int(lms.sum())
            pos_all += (pos.astype(np.int64) + lo).tolist()
        assert pos_all == np.flatnonzero(lms).tolist()


@pytest.mark.gpu
def test_sharded_cuda_kernels_single_gpu():
    """The CUDA shard kernels with halos, shards processed one after another on
    one GPU (same code path a rank runs; the collectives are covered above)."""
    from suffix_b200 import _lib
    ctx = _lib.Context(0)
    eng = sharded.CudaShardEngine(ctx, torch)
    rng = np.random.default_rng(11)
    texts = [np.frombuffer(d, dtype=np.uint8) for _, d in families.adversarial() if len(d) >= 64]
    texts += [gen.dna(3_000_017), gen.rand_bytes(1_000_003), np.full(100_000, 65, dtype=np.uint8)]
    for t in texts:
        S, lms, hist = _expected(t)
        n = len(t)
        parts = 3
        cuts = [0] + sorted((rng.integers(1, n // 16, parts - 1) * 16 % n).tolist()) + [n]   # 16-byte aligned cuts
        cuts = sorted(set(cuts))
        d_t = torch.from_numpy(t.copy()).cuda()
        shards = [d_t[cuts[i]:cuts[i + 1]] for i in range(len(cuts) - 1)]
        edges = [eng.edge_bytes(s) for s in shards]
        states = [eng.summary(s, edges[i + 1][0] if i + 1 < len(shards) else -1) for i, s in enumerate(shards)]
        tails = sharded.resolve_tail_carries(states)
        hsum = np.zeros(768, dtype=np.uint64)
        pos_all = []
        for i, s in enumerate(shards):
            nxt = edges[i + 1][0] if i + 1 < len(shards) else -1
            prv = edges[i - 1][1] if i > 0 else -1
            sw, lw, pos, h, m = eng.classify(s, prv, nxt, tails[i])
            lo, n_loc = cuts[i], s.numel()
            assert np.array_equal(_unpack(sw.cpu().numpy(), n_loc), S[lo:lo + n_loc])
            assert np.array_equal(_unpack(lw.cpu().numpy(), n_loc), lms[lo:lo + n_loc])
            hsum += h
            pos_all += (pos.cpu().numpy().view(np.uint32).astype(np.int64) + lo).tolist()
        assert np.array_equal(hsum, hist)
        assert pos_all == np.flatnonzero(lms).tolist()


def _nccl_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from suffix_b200 import _lib
    ctx = _lib.Context(rank)
    eng = sharded.CudaShardEngine(ctx, torch)
    n = 8_000_000
    t = gen.dna(n * world)                         # every rank generates the same text, keeps its shard
    shard = torch.from_numpy(t[rank * n:(rank + 1) * n].copy()).cuda()
    r = sharded.classify_sharded(eng, shard, dist=dist, device=torch.device("cuda", rank))
    S, lms, hist = _expected(t)
    ok = np.array_equal(_unpack(r.stype_words.cpu().numpy(), n), S[r.lo:r.lo + n]) and \
        np.array_equal(_unpack(r.lms_words.cpu().numpy(), n), lms[r.lo:r.lo + n]) and \
        np.array_equal(r.hist_global, hist) and r.m_total == int(lms.sum())
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}


# ---------------------------------------------------------------------------------------
# sharded LMS-suffix sort: numpy mirror under gloo; the CUDA + NCCL
# product path on one and two GPUs
def _sorted_lms_oracle(t):
    ty = oracle.types(t)
    sa = oracle.sais(t)
    return sa[ty[sa] == 2].astype(np.int64)


def _check_global_order(t, gpos_all, names_all, ties, kc):
    want = _sorted_lms_oracle(t)
    assert len(gpos_all) == len(want)
    assert sorted(gpos_all.tolist()) == sorted(want.tolist())
    tb = t.tobytes()
    win = [tb[p:p + kc] for p in gpos_all.tolist()]
    assert all(win[i] <= win[i + 1] for i in range(len(win) - 1))          # ordered by the first kc bytes
    nm = names_all.tolist()
    assert all(0 <= nm[i + 1] - nm[i] <= 1 for i in range(len(nm) - 1)) and (not nm or nm[0] == 0)
    if ties == 0:
        assert np.array_equal(gpos_all, want)                                # exact suffix order
        assert nm == list(range(len(nm)))


def _lms_worker(rank, world, port, texts, cutsets, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.shard_ref import lms_sort_sharded_model
    out = []
    for t, cuts in zip(texts, cutsets):
        g, nm, st = lms_sort_sharded_model(t[cuts[rank]:cuts[rank + 1]], dist)
        out.append((g, nm, st["ties_total"], st["kc"]))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lms_sort_model_gloo(world):
    rng = np.random.default_rng(50 + world)
    texts = [np.frombuffer(d, dtype=np.uint8) for nm_, d in families.adversarial() if 8 <= len(d) <= 12000][:14]
    texts += [gen.dna(30_000), gen.rand_bytes(8_000), gen.english(9_000), gen.dna(2_001, newline_tail=True)]
    cutsets = [_cuts(len(t), world, rng) for t in texts]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lms_worker, args=(r, world, port, texts, cutsets, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for k, t in enumerate(texts):
        g = np.concatenate([res[r][k][0] for r in range(world)])
        nm = np.concatenate([res[r][k][1] for r in range(world)])
        _check_global_order(t, g, nm, res[0][k][2], res[0][k][3])


@pytest.mark.gpu
def test_sharded_lms_sort_cuda_world1():
    """b200sa_shard_lms_sort on one GPU (a world of 1: same kernels, self-exchange)."""
    from suffix_b200 import _lib
    ctx = _lib.Context(0)
    cases = [gen.dna(3_000_000), gen.dna(500_001, newline_tail=True), gen.rand_bytes(1_000_000), gen.english(400_000),
             gen.fixture("AP009048_100000.fasta")]
    cases += [np.frombuffer(d, dtype=np.uint8) for _, d in families.adversarial() if len(d) >= 64]
    for t in cases:
        d_t = torch.from_numpy(np.ascontiguousarray(t).copy()).cuda()
        g, nm, st = sharded.lms_sort_sharded(ctx, d_t)
        torch.cuda.synchronize()
        assert st["n_total"] == len(t) and st["m_total"] == st["recv_count"] == g.numel()
        _check_global_order(np.ascontiguousarray(t), g.cpu().numpy(), nm.cpu().numpy().astype(np.int64), st["ties_total"], st["kc"])
    ctx.close()


def _nccl_lms_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from suffix_b200 import _lib
    ctx = _lib.Context(rank)
    out = []
    for t, cuts in ((gen.dna(4_000_000), [0, 1_999_984, 4_000_000]), (gen.rand_bytes(600_000), [0, 16, 600_000]),
                    (gen.dna(300_001, newline_tail=True), [0, 299_984, 300_001])):
        shard = torch.from_numpy(t[cuts[rank]:cuts[rank + 1]].copy()).cuda()
        g, nm, st = sharded.lms_sort_sharded(ctx, shard, dist=dist, cap=len(t) // 2 + 4096)
        torch.cuda.synchronize()
        out.append((g.cpu().numpy(), nm.cpu().numpy().astype(np.int64), st["ties_total"], st["kc"], st["bytes_sent"]))
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_lms_sort_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_lms_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    texts = [gen.dna(4_000_000), gen.rand_bytes(600_000), gen.dna(300_001, newline_tail=True)]
    for k, t in enumerate(texts):
        g = np.concatenate([res[r][k][0] for r in range(2)])
        nm = np.concatenate([res[r][k][1] for r in range(2)])
        _check_global_order(t, g, nm, res[0][k][2], res[0][k][3])
        assert res[0][k][4] > 0 or res[1][k][4] > 0          # something crossed NVLink


# ---------------------------------------------------------------------------------------
# sharded LCP: text-range-sharded Phi / PLCP
def _lcp_range_worker(rank, world, port, texts, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = []
    for t in texts:
        n = len(t)
        objs = [None]
        if rank == 0:
            objs = [oracle.sais(t)]                        # rank 0 holds the table; broadcast like the product does
        dist.broadcast_object_list(objs, src=0)
        sa = objs[0]
        per = ((n + world - 1) // world + 1023) // 1024 * 1024
        lo, hi = min(n, rank * per), min(n, rank * per + per)
        phi = {}
        for r in range(n):                                  # Phi restricted to the rank's text range
            i = int(sa[r])
            if lo <= i < hi:
                phi[i] = int(sa[r - 1]) if r else -1
        plcp = np.zeros(max(hi - lo, 0), dtype=np.int64)
        h = 0
        for i in range(lo, hi):                             # PLCP with the h-1 carry, restart at the range start
            j = phi[i]
            if j < 0:
                h = 0
            else:
                while i + h < n and j + h < n and t[i + h] == t[j + h]:
                    h += 1
            plcp[i - lo] = h
            h = max(h - 1, 0)
        parts = [None] * world
        dist.all_gather_object(parts, plcp)
        full = np.concatenate(parts)[:n]
        lcp_slice = full[sa[lo:hi]]                         # this rank's RANK range
        dist.all_gather_object(parts, lcp_slice)
        out.append(np.concatenate(parts)[:n])
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lcp_model_gloo(world):
    texts = [gen.dna(5000), gen.rand_bytes(3000), np.frombuffer(b"abracadabra" * 200, dtype=np.uint8), gen.english(4000)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_lcp_range_worker, args=(r, world, port, texts, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    for k, t in enumerate(texts):
        want = oracle.lcp_kasai(t, oracle.sais(t)).astype(np.int64)
        for r in range(world):
            assert np.array_equal(res[r][k], want), (k, r)


@pytest.mark.gpu
def test_sharded_lcp_cuda_world1():
    from suffix_b200 import _lib
    ctx = _lib.Context(0)
    for t in (gen.dna(2_000_000), gen.rand_bytes(700_001), gen.fixture("AP009048_100000.fasta"), gen.english(300_000)):
        sa = oracle.sais(t)
        d_t = torch.from_numpy(t.copy()).cuda()
        d_sa = torch.from_numpy(sa.astype(np.int64)).cuda().to(torch.int32)
        d_lcp = torch.empty(len(t), dtype=torch.int32, device="cuda")
        ctx.lcp_sharded(d_t.data_ptr(), len(t), d_sa.data_ptr(), d_lcp.data_ptr(), False,
                        torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(d_lcp.cpu().numpy().view(np.uint32), oracle.lcp_kasai(t, sa))
    ctx.close()


def _nccl_lcp_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from suffix_b200 import _lib
    ctx = _lib.Context(rank)
    sharded.ensure_comm(ctx, dist)
    ok = True
    for t in (gen.dna(3_000_000), gen.rand_bytes(1_000_001)):
        n = len(t)
        sa = oracle.sais(t)
        dev = torch.device("cuda", rank)
        if rank == 0:                                       # only rank 0 holds text and table
            d_t = torch.from_numpy(t.copy()).to(dev)
            d_sa = torch.from_numpy(sa.astype(np.int64)).to(dev).to(torch.int32)
        else:
            d_t = torch.zeros(n, dtype=torch.uint8, device=dev)
            d_sa = torch.zeros(n, dtype=torch.int32, device=dev)
        d_lcp = torch.empty(n, dtype=torch.int32, device=dev)
        ctx.lcp_sharded(d_t.data_ptr(), n, d_sa.data_ptr(), d_lcp.data_ptr(), False, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ok = ok and bool(np.array_equal(d_lcp.cpu().numpy().view(np.uint32), oracle.lcp_kasai(t, sa)))
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_lcp_nccl_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_lcp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: True, 1: True}
