"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: object partitioning and the
model-sharded merge (max / logsumexp / argmax all-reduce, PDF partial sum)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden
from oracle import fz_oracle as fo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from frankenz_b200.distributed import merge_pass1, merge_pdfs, shard_bounds
        g = golden("bruteforce_c1small.npz")
        m, me, mm = g["models"], g["models_err"], g["models_mask"]
        x, xe, xm = g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy()
        lab, labe = g["labels"], g["label_errs"]
        zgrid = np.arange(0, 7 + 1e-5, 0.01)
        kd = fo.KernelDict(zgrid, np.linspace(0.005, 2, 500))
        lo, hi = shard_bounds(len(m), world, rank)
        fit = fo.bruteforce_fit(m[lo:hi], me[lo:hi], mm[lo:hi], x, xe, xm)
        lp = fit["lnprob"]
        pmax = lp.max(axis=1)
        psum = np.exp(lp - pmax[:, None]).sum(axis=1)
        pbest = lp.argmax(axis=1)
        if rank == 1:       # exercise the NaN / all -inf branches on two objects
            pmax[0], psum[0] = np.nan, np.nan
        lmap, levid, best = merge_pass1(torch.from_numpy(pmax), torch.from_numpy(psum), torch.from_numpy(pbest), lo)
        yi, si = kd.fit(lab[lo:hi], labe[lo:hi])
        part = np.zeros((len(x), kd.Ngrid))
        for i in range(len(x)):
            if i == 0:
                continue
            wt = np.exp(lp[i] - levid[i].item())
            wt = np.where(wt > 1e-3 * np.exp(lmap[i].item() - levid[i].item()), wt, 0.0)
            part[i] = fo.kde_dict(kd, yi, si, y_wt=wt, wt_thresh=None, cdf_thresh=None)
        pdfs = merge_pdfs(torch.from_numpy(part))
        if rank == 0:
            q.put((lmap.numpy(), levid.numpy(), best.numpy(), pdfs.numpy()))
    finally:
        dist.destroy_process_group()


def test_model_sharded_merge_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    lmap, levid, best, pdfs = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    g = golden("bruteforce_c1small.npz")
    assert np.isnan(lmap[0]) and np.isnan(levid[0])
    assert np.allclose(lmap[1:], g["lmap"][1:], rtol=1e-13) and np.allclose(levid[1:], g["levid"][1:], rtol=1e-12)
    assert np.array_equal(best[1:], g["fit_lnprob"][1:].argmax(axis=1))
    assert np.max(np.sum(np.abs(pdfs[1:] - g["pdf_dict"][1:]), axis=1)) < 1e-12


def _worker_chunked(rank, world, port, q):
    """The collective sequence of ModelShardedBruteForce on CPU tensors over gloo: ONE all-gather of the packed pass-1
    partials, merge_gathered, ONE reduce-scatter of fp32 PDF partials per chunk, normalisation by the owner."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from frankenz_b200.distributed import merge_gathered, owned_rows, shard_bounds
        g = golden("bruteforce_c1small.npz")
        m, me, mm = g["models"], g["models_err"], g["models_mask"]
        x, xe, xm = g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy()
        lab, labe = g["labels"], g["label_errs"]
        zgrid = np.arange(0, 7 + 1e-5, 0.01)
        kd = fo.KernelDict(zgrid, np.linspace(0.005, 2, 500))
        lo, hi = shard_bounds(len(m), world, rank)
        yi, si = kd.fit(lab[lo:hi], labe[lo:hi])
        chunk = 10                                   # 32 objects -> chunks of 10, 10, 10, 2 (ragged tail)
        rows, idx, lm_all, best_all = [], [], [], []
        for c0 in range(0, len(x), chunk):
            sl = slice(c0, min(len(x), c0 + chunk))
            nc = sl.stop - sl.start
            lp = fo.bruteforce_fit(m[lo:hi], me[lo:hi], mm[lo:hi], x[sl], xe[sl], xm[sl])["lnprob"]
            pmax = lp.max(axis=1)
            packed = np.stack([pmax, np.exp(lp - pmax[:, None]).sum(axis=1),
                               (lp.argmax(axis=1) + lo).astype(np.int64).view(np.float64)])
            if rank == 1 and c0 == 0:
                packed[0, 0] = packed[1, 0] = np.nan          # a poisoned object
            gathered = torch.empty((world * 3, nc), dtype=torch.float64)
            dist.all_gather_into_tensor(gathered, torch.from_numpy(packed))
            lmap, levid, best = merge_gathered(gathered.view(world, 3, nc))
            olo, ohi, qrows = owned_rows(nc, world, rank)
            part = np.zeros((qrows * world, kd.Ngrid), dtype=np.float32)
            for i in range(nc):
                if np.isnan(lmap[i].item()):
                    continue
                wt = np.exp(lp[i] - lmap[i].item())            # weights relative to the GLOBAL maximum
                wt = np.where(wt > 1e-3, wt, 0.0)
                part[i] = fo.kde_dict(kd, yi, si, y_wt=wt, wt_thresh=None, cdf_thresh=None)
            own = torch.empty((qrows, kd.Ngrid), dtype=torch.float32)
            dist.reduce_scatter_tensor(own, torch.from_numpy(part), op=dist.ReduceOp.SUM)
            own = own[:ohi - olo].double()
            rows.append((own / own.sum(dim=1, keepdim=True)).numpy())
            idx.append(np.arange(c0 + olo, c0 + ohi))
            lm_all.append(lmap.numpy())
            best_all.append(best.numpy())
        q.put((rank, np.concatenate(idx), np.concatenate(rows), np.concatenate(lm_all), np.concatenate(best_all)))
    finally:
        dist.destroy_process_group()


def test_chunked_allgather_reduce_scatter_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_chunked, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(), q.get()]
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    g = golden("bruteforce_c1small.npz")
    seen = np.zeros(32, dtype=int)
    for rank, idx, rows, lmap, best in got:
        seen[idx] += 1
        ok = idx != 0
        assert np.max(np.sum(np.abs(rows[ok] - g["pdf_dict"][idx[ok]]), axis=1)) < 5e-6       # fp32 partials
        assert np.isnan(lmap[0]) and np.allclose(lmap[1:], g["lmap"][1:], rtol=1e-13)
        assert np.array_equal(best[1:], g["fit_lnprob"][1:].argmax(axis=1))
    assert np.all(seen == 1)                      # every object is owned by exactly one rank


def test_owned_rows_partition():
    from frankenz_b200.distributed import owned_rows
    for n in (1, 2, 7, 8, 65536, 65537):
        for w in (1, 2, 3, 8):
            parts = [owned_rows(n, w, r) for r in range(w)]
            assert parts[0][0] == 0 and max(p[1] for p in parts) == n
            assert sum(p[1] - p[0] for p in parts) == n
            assert all(p[2] * w >= n for p in parts)


def test_shard_bounds_cover_everything():
    from frankenz_b200.distributed import shard_bounds
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_merge_single_process_identity():
    from frankenz_b200.distributed import merge_pass1, merge_pdfs
    pmax = torch.tensor([-3.0, -float("inf"), 2.0], dtype=torch.float64)
    psum = torch.tensor([2.0, 0.0, 1.0], dtype=torch.float64)
    lmap, levid, best = merge_pass1(pmax, psum, torch.tensor([4, 0, 1]), 10)
    assert torch.allclose(lmap[[0, 2]], pmax[[0, 2]]) and lmap[1] == -float("inf")
    assert torch.allclose(levid[[0, 2]], torch.tensor([-3.0 + np.log(2.0), 2.0], dtype=torch.float64))
    assert levid[1] == -float("inf") and best.tolist() == [14, 10, 11]
    p = merge_pdfs(torch.tensor([[1.0, 3.0], [2.0, 2.0]], dtype=torch.float64))
    assert torch.allclose(p, torch.tensor([[0.25, 0.75], [0.5, 0.5]], dtype=torch.float64))


def _knn_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from frankenz_b200.distributed import merge_topk, shard_bounds
        rs = np.random.RandomState(3)
        feats = rs.normal(size=(2, 400, 5)).astype(np.float32)
        feats[1, 300] = feats[1, 20]                    # an exact tie across the two shards
        qf = rs.normal(size=(6, 5))
        qf[0] = feats[1, 20].astype(np.float64)
        k = 7
        lo, hi = shard_bounds(feats.shape[1], world, rank)
        idx = np.empty((len(qf), 2, k), dtype=np.int64)
        dd = np.empty((len(qf), 2, k))
        for i in range(len(qf)):                        # the per-shard exact search (the GPU kernel's job), by the oracle
            oi, od = fo.knn_query_exact(feats[:, lo:hi], qf[i], k, 2)
            idx[i], dd[i] = oi, od
        dm, im = merge_topk(torch.from_numpy(dd), torch.from_numpy(idx + lo), k)
        if rank == 0:
            q.put((feats, qf, im.numpy(), dm.numpy()))
    finally:
        dist.destroy_process_group()


def test_knn_row_sharded_merge_gloo():
    """Per-shard top-k lists merged over two gloo ranks = the exact search of the whole set (knn.py:362-365), ties to the
    lowest row index."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_knn_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    feats, qf, im, dm = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for i in range(len(qf)):
        oi, od = fo.knn_query_exact(feats, qf[i], 7, 2)
        assert np.array_equal(im[i], oi) and np.array_equal(dm[i], od)
