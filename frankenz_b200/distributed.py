"""Multi-GPU drivers of the brute-force path: one process per GPU, `torch.distributed` plumbing.

Two partitions:

* objects sharded, models replicated (`fit_predict_object_sharded`): every rank fits a contiguous
  slice of the objects against all models.  Objects are independent (bruteforce.py:192), so there
  is NO data-path collective; an optional all-gather assembles the outputs.
* kNN with the training rows sharded (`knn_query_row_sharded`): every rank searches its slice of every tree exactly,
  one all-gather of the per-shard top-k (distance, global row) and a k-way merge by (distance, index) give the
  neighbours of the whole set - the reference's `KDTree.query` call site (knn.py:362-365) for trees too large for one GPU.
* models sharded (`fit_predict_model_sharded`): every rank holds a slice of the models and sees all
  objects.  Per object the three associative reductions of the path are merged across ranks:
      lmap  = max_g pmax_g                                         one all-gather of the packed partials
      levid = lmap + ln sum_g psum_g * exp(pmax_g - lmap)          (24 B / object / rank), merged on every rank
      pdf   = sum_g pdf_g / sum(...)                               reduce-scatter(SUM) of fp32 partials to the owner
  between pass 1 (`fzb_shard_pass1_packed_dev`) and pass 2 (`fzb_shard_pass2_f32_dev`, which needs the GLOBAL
  lmap / levid for the wt_thresh selection, pdf.py:589-591), chunk by chunk, the reduce-scatter of chunk c running
  beside pass 1 of chunk c+1 (`ModelShardedBruteForce`).

The merge arithmetic is written on torch tensors of any device so that the world_size-2 gloo tests
in tests/test_distributed_cpu.py exercise exactly the code that runs over NCCL.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._engine import Engine, clean_inplace, make_config

try:
    import torch.distributed as dist
except Exception:  # pragma: no cover
    dist = None


def _world(group=None):
    if dist is not None and dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_bounds(n, world, rank):
    """Contiguous, balanced slice [lo, hi) of `n` items for `rank` of `world`."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _all_reduce(t, op, group=None):
    if _world(group)[1] > 1:
        dist.all_reduce(t, op=op, group=group)
    return t


def merge_pass1(pmax, psum, pbest, best_offset, group=None):
    """Merge per-rank partial (max, sum exp(l - max), argmax) into global (lmap, levid, best).

    pmax/psum float64 [No], pbest int64 [No] (index local to this rank's model slice),
    `best_offset` = first global model index of this rank.  NaN partials poison the object
    (numpy max / logsumexp semantics of bruteforce.py:359).
    """
    rank, world = _world(group)
    nan = torch.isnan(pmax) | torch.isnan(psum)
    nanflag = _all_reduce(nan.to(torch.float64), dist.ReduceOp.MAX if world > 1 else None, group)
    clean = torch.where(nan, torch.full_like(pmax, -float("inf")), pmax)
    gmax = _all_reduce(clean.clone(), dist.ReduceOp.MAX if world > 1 else None, group)
    # sum_g psum_g * exp(pmax_g - gmax); a rank whose partial max is -inf contributes nothing
    shift = torch.where(torch.isinf(clean) & (clean < 0), torch.zeros_like(clean), torch.exp(clean - gmax))
    s = torch.where(torch.isfinite(clean), psum * shift, torch.zeros_like(psum))
    s = _all_reduce(s, dist.ReduceOp.SUM if world > 1 else None, group)
    levid = gmax + torch.log(s)
    levid = torch.where(torch.isinf(gmax), gmax, levid)
    # argmax: lowest global index among the ranks that hold the maximum
    big = torch.iinfo(torch.int64).max
    cand = torch.where(clean == gmax, pbest + int(best_offset), torch.full_like(pbest, big))
    best = _all_reduce(cand, dist.ReduceOp.MIN if world > 1 else None, group)
    lmap = torch.where(nanflag > 0, torch.full_like(gmax, float("nan")), gmax)
    levid = torch.where(nanflag > 0, torch.full_like(levid, float("nan")), levid)
    return lmap, levid, best


def merge_pdfs(pdf_partial, group=None):
    """Sum the un-normalised PDF partials over ranks and normalise each row (bruteforce.py:370)."""
    world = _world(group)[1]
    total = _all_reduce(pdf_partial, dist.ReduceOp.SUM if world > 1 else None, group)
    return total / total.sum(dim=1, keepdim=True)


def _dev_f64(a, device):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)


def merge_gathered(gathered, group=None):
    """Global (lmap, levid, best) from the all-gathered packed partials [world, 3, No] (max, sum, int64 bits of the
    global arg-max): the torch restatement of `k_shard_merge` (csrc/fzb_shard.cu), used on CPU tensors by the gloo
    tests and as the checker of the kernel."""
    pmax, psum = gathered[:, 0], gathered[:, 1]
    pbest = gathered[:, 2].contiguous().view(torch.int64)
    nan = torch.isnan(pmax) | torch.isnan(psum)
    clean = torch.where(nan, torch.full_like(pmax, -float("inf")), pmax)
    gmax = clean.max(dim=0).values
    s = torch.where(torch.isfinite(clean), psum * torch.exp(clean - gmax), torch.zeros_like(psum)).sum(dim=0)
    levid = torch.where(torch.isinf(gmax), gmax, gmax + torch.log(s))
    big = torch.iinfo(torch.int64).max
    best = torch.where(clean == gmax, pbest, torch.full_like(pbest, big)).min(dim=0).values
    poisoned = nan.any(dim=0)
    lmap = torch.where(poisoned, torch.full_like(gmax, float("nan")), gmax)
    levid = torch.where(poisoned, torch.full_like(levid, float("nan")), levid)
    return lmap, levid, best


def merge_topk(dist_local, idx_global, k, group=None):
    """k-way merge of per-rank nearest-neighbour lists.

    dist_local float64 [No, K, k] (ascending per list, inf = missing), idx_global int64 [No, K, k] (row indices of the WHOLE
    training set, i.e. already offset by the rank's first row).  Every rank gets the k smallest of the union ordered by
    (distance, index), the order of the single-GPU search (exact ties go to the lowest row index)."""
    rank, world = _world(group)
    d, i = dist_local.contiguous(), idx_global.contiguous()
    if world > 1:
        gd = [torch.empty_like(d) for _ in range(world)]
        gi = [torch.empty_like(i) for _ in range(world)]
        dist.all_gather(gd, d, group=group)
        dist.all_gather(gi, i, group=group)
        d, i = torch.cat(gd, dim=-1), torch.cat(gi, dim=-1)
    return topk_lex(d, i, k)


def topk_lex(d, i, k):
    """The k smallest (distance, index) pairs along the last axis, lexicographic: stable sort by index, then by distance."""
    o1 = torch.argsort(i, dim=-1, stable=True)
    d, i = torch.gather(d, -1, o1), torch.gather(i, -1, o1)
    o2 = torch.argsort(d, dim=-1, stable=True)
    d, i = torch.gather(d, -1, o2), torch.gather(i, -1, o2)
    return d[..., :k].contiguous(), i[..., :k].contiguous()


def knn_query_row_sharded(features, qfeats, k, p=2, group=None, device=None, engine=None):
    """Exact k nearest rows per tree with the training rows sharded over the ranks.

    `features` float32 [K, Nrows, Nf]: the WHOLE feature set (every rank keeps its `shard_bounds` slice of the rows of
    every tree on its GPU); `qfeats` float64 [No, Nf] on every rank.  Returns (idx int64 [No, K, k], dist float64 [No, K, k])
    as `Engine.knn_query` would for the unsharded set.  Rows per rank must be >= k."""
    rank, world = _world(group)
    feats = np.asarray(features, dtype=np.float32)
    lo, hi = shard_bounds(feats.shape[1], world, rank)
    if hi - lo < k:
        raise ValueError("every rank needs at least k=%d training rows, rank %d has %d" % (k, rank, hi - lo))
    own = engine is None
    if own:
        nf = feats.shape[2]
        ones = np.ones((hi - lo, nf))
        engine = Engine(ones, ones, ones, device=device)
        engine.knn_build(np.ascontiguousarray(feats[:, lo:hi]))
    try:
        idx, dd = engine.knn_query(qfeats, k, p=p)
    finally:
        if own:
            engine.close()
    dev = torch.device("cuda", engine.device) if torch.cuda.is_available() else torch.device("cpu")
    dt, it = merge_topk(torch.from_numpy(dd).to(dev), torch.from_numpy(idx + lo).to(dev), k, group=group)
    return it.cpu().numpy(), dt.cpu().numpy()


def owned_rows(n, world, rank):
    """Rows of a chunk of `n` objects that `rank` owns after the reduce-scatter: the chunk is padded to a multiple of
    `world` rows and split evenly.  Returns (lo, hi, rows_per_rank) with hi clipped to n."""
    q = (int(n) + world - 1) // world
    lo = min(int(n), rank * q)
    return lo, min(int(n), lo + q), q


class ModelShardedBruteForce(object):
    """Model-sharded `BruteForce.fit_predict(save_fits=False)` with the shard resident on this rank's GPU
    (the C5 configuration).

    Rank g of `group` keeps models [lo_g, hi_g) (records, tiles and KDE tables are built once and reused from call to
    call).  Every call takes ALL objects of the batch on every rank and walks them in chunks of `chunk` objects:

        pass 1 (this rank's models)      fzb_shard_pass1_packed_dev -> (max, sum, global arg-max), 24 B / object
        all-gather of the packed partials                              ONE collective, communication stream
        merge -> global lmap / levid / best (k_shard_merge)           library stream
        pass 2 with the global lmap      fzb_shard_pass2_f32_dev   -> un-normalised PDF partial, fp32
        reduce-scatter (sum) of the partials to the owning rank        ONE collective, communication stream,
                                                                       overlapped with pass 1 of the NEXT chunk
        normalise the owned rows (k_shard_normalise)                   library stream, next iteration

    The library stream and the communication stream are ordered with events only (`Stream.wait_stream`): there is no
    host synchronisation around the collectives.  With `gather=False` every rank returns the PDFs of the objects it
    owns (and their indices); with `gather=True` the owned rows are all-gathered so that all ranks return the full
    arrays, identical everywhere."""

    def __init__(self, models, models_err, models_mask, group=None, device=None, chunk=65536):
        self.group = group
        self.rank, self.world = _world(group)
        self.lo, self.hi = shard_bounds(len(models), self.world, self.rank)
        self.eng = Engine(models[self.lo:self.hi], models_err[self.lo:self.hi], models_mask[self.lo:self.hi], device=device)
        self.dev = torch.device("cuda", self.eng.device)
        self.chunk = max(self.world, int(chunk) // self.world * self.world)
        self._buf = {}
        sp = C.c_void_p()
        _lib.check(self.eng.lib.fzb_get_stream(self.eng.h, C.byref(sp)))
        self.lib_stream = torch.cuda.ExternalStream(sp.value, device=self.dev)
        self.comm_stream = torch.cuda.Stream(self.dev)
        self.last = {}

    def _tensor(self, name, shape, dtype):
        t = self._buf.get(name)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.dev)
            self._buf[name] = t
        return t

    def owned_indices(self, no):
        """Global indices of the objects whose PDFs this rank returns with gather=False, in output order."""
        out = []
        for c0 in range(0, no, self.chunk):
            lo, hi, _ = owned_rows(min(self.chunk, no - c0), self.world, self.rank)
            out.append(np.arange(c0 + lo, c0 + hi))
        return np.concatenate(out) if out else np.zeros(0, dtype=np.int64)

    def fit_predict(self, data, data_err, data_mask, model_labels, model_label_errs, label_dict=None, label_grid=None,
                    lprob_kwargs=None, kde_kwargs=None, return_best=False, as_torch=False, gather=True):
        """Returns (pdfs, (lmap, levid)[, best]): lmap / levid / best for ALL objects on every rank; pdfs for all
        objects (gather=True) or for `owned_indices(len(data))` (gather=False).  numpy arrays, or device tensors with
        as_torch=True (no device-to-host copy)."""
        eng, lo, hi, W, rank = self.eng, self.lo, self.hi, self.world, self.rank
        lk = dict(lprob_kwargs or {})
        eng.set_lnprior(None if lk.get("lnprior", None) is None else np.asarray(lk["lnprior"])[lo:hi])
        lk.pop("lnprior", None)
        eng.set_kde(np.asarray(model_labels)[lo:hi], np.asarray(model_label_errs)[lo:hi], label_dict=label_dict,
                    label_grid=label_grid, kde_kwargs=kde_kwargs)
        cfg = make_config(lk, kde_kwargs)
        if torch.is_tensor(data):
            d_x, d_xe, d_xm = (t.to(self.dev, torch.float64).contiguous() for t in (data, data_err, data_mask))
        else:
            clean_inplace(data, data_err, data_mask)
            d_x, d_xe, d_xm = _dev_f64(data, self.dev), _dev_f64(data_err, self.dev), _dev_f64(data_mask, self.dev)
        no, nf, ng, Cn = len(d_x), d_x.shape[1], eng.Ng, self.chunk
        q_max = Cn // W
        lib, L, Cm = eng.lib, self.lib_stream, self.comm_stream
        lmap = self._tensor("lmap", (no,), torch.float64)
        levid = self._tensor("levid", (no,), torch.float64)
        best = self._tensor("best", (no,), torch.int64)
        packed = self._tensor("packed", (3, Cn), torch.float64)
        gathered = self._tensor("gathered", (W * 3, Cn), torch.float64)
        partial = [self._tensor("partial%d" % b, (Cn, ng), torch.float32) for b in range(2)]
        owned = [self._tensor("owned%d" % b, (q_max, ng), torch.float32) for b in range(2)] if W > 1 else partial
        n_own = len(self.owned_indices(no))
        pdf_own = self._tensor("pdf_own", (max(n_own, 1), ng), torch.float64)
        ev = []          # (start, end, bytes) of every collective, on the communication stream
        use_nccl = W > 1

        def collective(fn, nbytes):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(Cm)
            fn()
            e1.record(Cm)
            ev.append((e0, e1, nbytes))

        # uploads / earlier torch work on the current stream must be visible to the library stream
        L.wait_stream(torch.cuda.current_stream(self.dev))
        t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_start.record(L)
        ms_p1 = ms_p2 = 0.0
        pending = None     # (buffer, rows, output offset) of the chunk whose reduce-scatter is in flight
        out_off = 0
        for ci, c0 in enumerate(range(0, no, Cn)):
            nc = min(Cn, no - c0)
            b = ci & 1
            xs = (d_x[c0:c0 + nc], d_xe[c0:c0 + nc], d_xm[c0:c0 + nc])
            pk = packed[:, :nc] if nc == Cn else self._tensor("packed_tail", (3, nc), torch.float64)
            _lib.check(lib.fzb_shard_pass1_packed_dev(eng.h, xs[0].data_ptr(), xs[1].data_ptr(), xs[2].data_ptr(), nc,
                                                      C.byref(cfg), lo, pk.data_ptr()))
            ms_p1 += eng.stats()["ms_total"]
            if use_nccl:
                ga = gathered if nc == Cn else self._tensor("gathered_tail", (W * 3, nc), torch.float64)
                Cm.wait_stream(L)
                with torch.cuda.stream(Cm):
                    collective(lambda: dist.all_gather_into_tensor(ga, pk, group=self.group), pk.numel() * 8 * W)
                L.wait_stream(Cm)          # the partials of all ranks are here (and the previous reduce-scatter is done)
            else:
                ga = pk
            if pending is not None:
                pb, rows, off = pending
                _lib.check(lib.fzb_shard_normalise_dev(eng.h, pb.data_ptr(), rows, ng, pdf_own[off:].data_ptr()))
                pending = None
            _lib.check(lib.fzb_shard_merge_dev(eng.h, ga.data_ptr(), W, nc, lmap[c0:].data_ptr(), levid[c0:].data_ptr(),
                                               best[c0:].data_ptr()))
            olo, ohi, q = owned_rows(nc, W, rank)
            part = partial[b][:q * W]
            _lib.check(lib.fzb_shard_pass2_f32_dev(eng.h, xs[0].data_ptr(), xs[1].data_ptr(), xs[2].data_ptr(), nc,
                                                   C.byref(cfg), lmap[c0:].data_ptr(), levid[c0:].data_ptr(),
                                                   part.data_ptr()))
            ms_p2 += eng.stats()["ms_total"]
            if use_nccl:
                own = owned[b][:q]
                Cm.wait_stream(L)
                with torch.cuda.stream(Cm):
                    collective(lambda: dist.reduce_scatter_tensor(own, part, op=dist.ReduceOp.SUM, group=self.group),
                               part.numel() * 4)
            else:
                own = part
            pending = (own, ohi - olo, out_off)
            out_off += ohi - olo
        if pending is not None:
            if use_nccl:
                L.wait_stream(Cm)
            pb, rows, off = pending
            _lib.check(lib.fzb_shard_normalise_dev(eng.h, pb.data_ptr(), rows, ng, pdf_own[off:].data_ptr()))
        t_end.record(L)
        torch.cuda.current_stream(self.dev).wait_stream(L)
        L.synchronize()
        self.last = {"ms_total": t_start.elapsed_time(t_end), "ms_pass1": ms_p1, "ms_pass2": ms_p2,
                     "ms_nccl": float(sum(a.elapsed_time(b_) for a, b_, _ in ev)),
                     "nccl_bytes": int(sum(n for _, _, n in ev)), "collectives": len(ev), "chunks": (no + Cn - 1) // Cn,
                     "owned_objects": n_own}
        pdfs = pdf_own[:n_own]
        if gather and W > 1:
            pdfs = self._gather_owned(pdfs, no, ng)
        if as_torch:
            out = (pdfs, (lmap, levid))
            return out + (best,) if return_best else out
        out = (pdfs.cpu().numpy(), (lmap.cpu().numpy(), levid.cpu().numpy()))
        return out + (best.cpu().numpy(),) if return_best else out

    def _gather_owned(self, pdf_own, no, ng):
        """All-gather the owned rows and put them back in object order (every rank ends with the full array)."""
        W = self.world
        counts = []
        for r in range(W):
            n = 0
            for c0 in range(0, no, self.chunk):
                lo, hi, _ = owned_rows(min(self.chunk, no - c0), W, r)
                n += hi - lo
            counts.append(n)
        nmax = max(counts)
        send = torch.zeros((nmax, ng), dtype=torch.float64, device=self.dev)
        send[:len(pdf_own)] = pdf_own
        recv = torch.empty((W * nmax, ng), dtype=torch.float64, device=self.dev)
        dist.all_gather_into_tensor(recv, send, group=self.group)
        full = torch.empty((no, ng), dtype=torch.float64, device=self.dev)
        for r in range(W):
            off = 0
            for c0 in range(0, no, self.chunk):
                lo, hi, _ = owned_rows(min(self.chunk, no - c0), W, r)
                full[c0 + lo:c0 + hi] = recv[r * nmax + off:r * nmax + off + hi - lo]
                off += hi - lo
        return full

    def close(self):
        self.eng.close()


def fit_predict_model_sharded(models, models_err, models_mask, data, data_err, data_mask, model_labels,
                              model_label_errs, label_dict=None, label_grid=None, lprob_kwargs=None,
                              kde_kwargs=None, group=None, device=None, return_best=False):
    """One-shot form of `ModelShardedBruteForce` (builds the shard, runs one batch, frees it).

    Every rank passes the FULL model arrays (or arrays of which it only needs its own slice) and all
    objects; rank g keeps models [lo_g, hi_g).  Returns (pdfs, (lmap, levid)) as numpy arrays, identical
    on every rank.
    """
    sb = ModelShardedBruteForce(models, models_err, models_mask, group=group, device=device)
    try:
        return sb.fit_predict(data, data_err, data_mask, model_labels, model_label_errs, label_dict=label_dict,
                              label_grid=label_grid, lprob_kwargs=lprob_kwargs, kde_kwargs=kde_kwargs,
                              return_best=return_best)
    finally:
        sb.close()


def fit_predict_object_sharded(bf, data, data_err, data_mask, model_labels, model_label_errs, gather=True,
                               group=None, **kwargs):
    """Object-sharded `BruteForce.fit_predict`: rank g handles objects [lo_g, hi_g) on its own GPU.

    `bf` is this rank's `BruteForce` (models replicated).  With gather=True the per-rank slices are
    all-gathered so every rank returns the full (Ndata, Ngrid) array; otherwise each rank returns its slice
    and (lo, hi).  No collective touches the likelihood data path.
    """
    rank, world = _world(group)
    lo, hi = shard_bounds(len(data), world, rank)
    kwargs.setdefault("verbose", False)
    kwargs["return_gof"] = True
    kwargs.setdefault("save_fits", False)
    p, (lm, le) = bf.fit_predict(data[lo:hi], data_err[lo:hi], data_mask[lo:hi], model_labels, model_label_errs,
                                 **kwargs)
    if not gather or world == 1:
        return p, (lm, le), (lo, hi)
    outs = []
    for arr in (p, lm, le):
        parts = [None] * world
        dist.all_gather_object(parts, arr, group=group)
        outs.append(np.concatenate(parts, axis=0))
    return outs[0], (outs[1], outs[2]), (0, len(data))
