"""Model-sharded mode over real NCCL ranks: two processes, one GPU each, when the box has two
GPUs (skipped otherwise); plus the same kernels driven shard by shard on one GPU."""
import os
import socket

import numpy as np
import pytest

import bench_data

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(nm=6000, no=700):
    m, me, mm, z, x, xe, xm, _ = bench_data.c1_dataset(nm, no)
    x = x.copy()
    x[5, 1] = np.nan                      # cleaned band (pdf.py:310-311)
    xm = xm.copy()
    xm[9] = 0.0                           # all-masked object: NaN row, poisons lmap / levid on every rank
    zgrid, sig = bench_data.c3_kde()
    return m, me, mm, z, np.full(nm, 0.05), x, xe, xm, zgrid, sig


def _nccl_worker(rank, world, port, q):
    """Never leaves the parent waiting: whatever happens, one item per rank reaches the queue."""
    try:
        _nccl_worker_body(rank, world, port, q)
    except BaseException:
        import traceback
        q.put((rank, "ERROR: " + traceback.format_exc()))


def _nccl_worker_body(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), FZB_DEVICE=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import frankenz_b200 as fz
        from frankenz_b200.distributed import ModelShardedBruteForce
        m, me, mm, z, labe, x, xe, xm, zgrid, sig = _problem()
        rdict = fz.pdf.PDFDict(zgrid, sig)
        out = {}
        for name, kw in (("default", dict()), ("free", dict(free_scale=True, ignore_model_err=True))):
            sb = ModelShardedBruteForce(m, me, mm, device=rank, chunk=256)      # 700 objects: chunks 256, 256, 188
            p, (lm, le), best = sb.fit_predict(x.copy(), xe.copy(), xm.copy(), z, labe, label_dict=rdict,
                                               lprob_kwargs=kw, return_best=True, gather=True)
            p_own, _ = sb.fit_predict(x.copy(), xe.copy(), xm.copy(), z, labe, label_dict=rdict, lprob_kwargs=kw,
                                      gather=False)
            own = sb.owned_indices(len(x))
            assert sb.last["collectives"] == 2 * 3 and sb.last["nccl_bytes"] > 0
            assert np.allclose(p_own, p[own], rtol=0, atol=2e-6, equal_nan=True)      # fp32 atomics: run-to-run order
            out[name] = (p, lm, le, best, own)
            sb.close()
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_model_sharded_over_two_nccl_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import frankenz_b200 as fz
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        items = [q.get(timeout=300), q.get(timeout=300)]
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.kill()
    for rank, payload in items:
        assert not isinstance(payload, str), "rank %d failed:\n%s" % (rank, payload)
    got = dict(items)
    m, me, mm, z, labe, x, xe, xm, zgrid, sig = _problem()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    seen = np.zeros(len(x), dtype=int)
    for r in (0, 1):
        seen[got[r]["default"][4]] += 1
    assert np.all(seen == 1)
    for name, kw in (("default", dict()), ("free", dict(free_scale=True, ignore_model_err=True))):
        bf = fz.BruteForce(m, me, mm)
        p1, (lm1, le1) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), z, labe, label_dict=rdict, return_gof=True,
                                        verbose=False, save_fits=False, lprob_kwargs=dict(kw, precision="fp64"))
        good = np.isfinite(lm1)
        assert not good[9]
        for r in (0, 1):
            p, lm, le, best, _ = got[r][name]
            assert np.array_equal(np.isfinite(lm), good) and np.array_equal(np.isfinite(le), good)
            assert np.max(np.sum(np.abs(p[good] - p1[good]), axis=1)) <= 1e-5
            assert np.all(np.abs(lm[good] - lm1[good]) <= 1e-5 * np.maximum(1, np.abs(lm1[good])))
            assert np.all(np.abs(le[good] - le1[good]) <= 1e-5 * np.maximum(1, np.abs(le1[good])))
            assert np.mean(best[good] == bf.best_idx[good]) > 0.99
        assert np.array_equal(got[0][name][0], got[1][name][0], equal_nan=True)       # identical on every rank (one gather)


@pytest.mark.parametrize("kw", [dict(), dict(free_scale=True, ignore_model_err=True)])
def test_shard_kernels_three_shards_one_gpu(kw):
    """pass1_packed -> k_shard_merge -> pass2_f32 -> sum -> k_shard_normalise with three shards on one GPU (the
    collectives replaced by their definitions): against the unsharded float64 run and against the torch restatement
    of the merge."""
    import ctypes as C
    import torch
    import frankenz_b200 as fz
    from frankenz_b200 import _lib
    from frankenz_b200._engine import Engine, make_config
    from frankenz_b200.distributed import merge_gathered, shard_bounds
    m, me, mm, z, labe, x, xe, xm, zgrid, sig = _problem()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    n, nm, W = len(x), len(m), 3
    tx = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x, xe, xm)]
    cfg = make_config(kw, None)
    engs = []
    gathered = torch.empty((W, 3, n), dtype=torch.float64).cuda()
    for r in range(W):
        lo, hi = shard_bounds(nm, W, r)
        e = Engine(m[lo:hi], me[lo:hi], mm[lo:hi])
        e.set_kde(z[lo:hi], labe[lo:hi], label_dict=rdict)
        _lib.check(e.lib.fzb_shard_pass1_packed_dev(e.h, tx[0].data_ptr(), tx[1].data_ptr(), tx[2].data_ptr(), n,
                                                    C.byref(cfg), lo, gathered[r].data_ptr()))
        engs.append(e)
    torch.cuda.synchronize()
    lmap = torch.empty(n, dtype=torch.float64).cuda()
    levid = torch.empty(n, dtype=torch.float64).cuda()
    best = torch.empty(n, dtype=torch.int64).cuda()
    e0 = engs[0]
    _lib.check(e0.lib.fzb_shard_merge_dev(e0.h, gathered.data_ptr(), W, n, lmap.data_ptr(), levid.data_ptr(),
                                          best.data_ptr()))
    _lib.check(e0.lib.fzb_synchronize(e0.h))
    tl, te, tb = merge_gathered(gathered.cpu())
    assert np.array_equal(lmap.cpu().numpy(), tl.numpy(), equal_nan=True)
    assert np.allclose(levid.cpu().numpy(), te.numpy(), rtol=1e-13, atol=0, equal_nan=True)
    ok = np.isfinite(tl.numpy())
    assert np.array_equal(best.cpu().numpy()[ok], tb.numpy()[ok])
    tot = torch.zeros((n, 701), dtype=torch.float32).cuda()
    for e in engs:
        part = torch.full((n, 701), float("nan"), dtype=torch.float32).cuda()
        torch.cuda.synchronize()
        _lib.check(e.lib.fzb_shard_pass2_f32_dev(e.h, tx[0].data_ptr(), tx[1].data_ptr(), tx[2].data_ptr(), n,
                                                 C.byref(cfg), lmap.data_ptr(), levid.data_ptr(), part.data_ptr()))
        tot += part
    torch.cuda.synchronize()
    pdfs = torch.empty((n, 701), dtype=torch.float64).cuda()
    _lib.check(e0.lib.fzb_shard_normalise_dev(e0.h, tot.data_ptr(), n, 701, pdfs.data_ptr()))
    _lib.check(e0.lib.fzb_synchronize(e0.h))
    bf = fz.BruteForce(m, me, mm)
    p1, (lm1, le1) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), z, labe, label_dict=rdict, return_gof=True,
                                    verbose=False, save_fits=False, lprob_kwargs=dict(kw, precision="fp64"))
    good = np.isfinite(lm1)
    p = pdfs.cpu().numpy()
    assert np.array_equal(np.isfinite(lmap.cpu().numpy()), good)
    assert np.max(np.sum(np.abs(p[good] - p1[good]), axis=1)) <= 1e-5
    assert np.all(np.abs(lmap.cpu().numpy()[good] - lm1[good]) <= 1e-5 * np.maximum(1, np.abs(lm1[good])))
    assert np.all(np.abs(levid.cpu().numpy()[good] - le1[good]) <= 1e-5 * np.maximum(1, np.abs(le1[good])))
    # the fp32 sweep decides the arg-max; models within its rounding of the maximum may swap (lmap agrees regardless)
    assert np.mean(best.cpu().numpy()[good] == bf.best_idx[good]) > 0.99


def test_single_rank_driver_chunks_and_ragged_tail():
    """ModelShardedBruteForce without a process group, several chunks with a ragged tail: the chunk loop, the event
    ordering between the library stream and the communication stream, the fp32 partials."""
    import frankenz_b200 as fz
    from frankenz_b200.distributed import ModelShardedBruteForce
    m, me, mm, z, labe, x, xe, xm, zgrid, sig = _problem()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    sb = ModelShardedBruteForce(m, me, mm, chunk=300)
    p, (lm, le), best = sb.fit_predict(x.copy(), xe.copy(), xm.copy(), z, labe, label_dict=rdict, return_best=True)
    assert sb.last["chunks"] == 3 and np.array_equal(sb.owned_indices(len(x)), np.arange(len(x)))
    bf = fz.BruteForce(m, me, mm)
    p1, (lm1, le1) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), z, labe, label_dict=rdict, return_gof=True,
                                    verbose=False, save_fits=False)
    good = np.isfinite(lm1)
    assert np.array_equal(np.isfinite(lm), good)
    assert np.max(np.sum(np.abs(p[good] - p1[good]), axis=1)) <= 1e-5
    assert np.allclose(lm[good], lm1[good], rtol=0, atol=1e-6) and np.array_equal(best[good], bf.best_idx[good])
    sb.close()


def _knn_problem():
    rs = np.random.RandomState(21)
    base = rs.normal(size=(30000, 5)) * np.array([1.0, 0.7, 0.5, 0.9, 1.3]) + 20.0
    feats = np.stack([base + rs.normal(size=base.shape) * 0.02 for _ in range(3)]).astype(np.float32)
    feats[2, 29000] = feats[2, 40]                       # an exact tie between the first and the last shard
    q = base[rs.choice(len(base), 600)] + rs.normal(size=(600, 5)) * 0.05
    q[3] = feats[2, 40].astype(np.float64)
    return feats, q


def test_knn_row_shards_merge_like_the_whole_set():
    """Row-sharded kNN (distributed.knn_query_row_sharded) with the all-gather replaced by a concatenation: three
    shards searched on one GPU and merged by `topk_lex` give the neighbours of the unsharded search, bit for bit."""
    import torch
    from frankenz_b200._engine import Engine
    from frankenz_b200.distributed import shard_bounds, topk_lex
    feats, q = _knn_problem()
    k = 25
    ones = np.ones((feats.shape[1], 5))
    e = Engine(ones, ones, ones)
    e.knn_build(feats)
    idx_all, dist_all = e.knn_query(q, k, p=2)
    e.close()
    ds, ix = [], []
    for r in range(3):
        lo, hi = shard_bounds(feats.shape[1], 3, r)
        es = Engine(ones[lo:hi], ones[lo:hi], ones[lo:hi])
        es.knn_build(np.ascontiguousarray(feats[:, lo:hi]))
        i, d = es.knn_query(q, k, p=2)
        es.close()
        ds.append(torch.from_numpy(d).cuda())
        ix.append(torch.from_numpy(i + lo).cuda())
    dm, im = topk_lex(torch.cat(ds, dim=-1), torch.cat(ix, dim=-1), k)
    assert np.array_equal(im.cpu().numpy(), idx_all) and np.array_equal(dm.cpu().numpy(), dist_all)


def _knn_nccl_worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), FZB_DEVICE=str(rank))
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        try:
            from frankenz_b200.distributed import knn_query_row_sharded
            feats, qf = _knn_problem()
            idx, dd = knn_query_row_sharded(feats, qf, 25, device=rank)
            q.put((rank, (idx, dd)))
        finally:
            dist.destroy_process_group()
    except BaseException:
        import traceback
        q.put((rank, "ERROR: " + traceback.format_exc()))


@pytest.mark.timeout(600)
def test_knn_row_sharded_over_two_nccl_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from frankenz_b200._engine import Engine
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_knn_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        items = [q.get(timeout=300), q.get(timeout=300)]
    finally:
        for p in procs:
            p.join(30)
            if p.is_alive():
                p.kill()
    for rank, payload in items:
        assert not isinstance(payload, str), "rank %d failed:\n%s" % (rank, payload)
    feats, qf = _knn_problem()
    ones = np.ones((feats.shape[1], 5))
    e = Engine(ones, ones, ones)
    e.knn_build(feats)
    idx_all, dist_all = e.knn_query(qf, 25, p=2)
    for rank, (idx, dd) in items:
        assert np.array_equal(idx, idx_all) and np.array_equal(dd, dist_all)
