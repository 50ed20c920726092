"""Multi-GPU drivers of the brute-force path: one process per GPU, `torch.distributed` plumbing.

Two partitions:

* objects sharded, models replicated (`fit_predict_object_sharded`): every rank fits a contiguous
  slice of the objects against all models.  Objects are independent (bruteforce.py:192), so there
  is NO data-path collective; an optional all-gather assembles the outputs.
* models sharded (`fit_predict_model_sharded`): every rank holds a slice of the models and sees all
  objects.  Per object the three associative reductions of the path are merged across ranks:
      lmap  = max_g pmax_g                                         all-reduce(MAX)
      levid = lmap + ln sum_g psum_g * exp(pmax_g - lmap)          all-reduce(SUM)
      pdf   = sum_g pdf_g / sum(...)                               all-reduce(SUM), then normalise
  between pass 1 (`fzb_shard_pass1_dev`) and pass 2 (`fzb_shard_pass2_dev`, which needs the GLOBAL
  lmap / levid for the wt_thresh selection, pdf.py:589-591).

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


class ModelShardedBruteForce(object):
    """Model-sharded `BruteForce.fit_predict(save_fits=False)` with the shard resident on this rank's GPU.

    Rank g of `group` keeps models [lo_g, hi_g) (records, tiles and KDE tables are built once and reused from call to
    call); every call takes ALL objects of the batch on every rank and exchanges, per object, the partial
    (max, sum, arg-max) of pass 1 (all-reduce MAX / SUM / MIN: 24 B) and the un-normalised PDF partial of pass 2
    (all-reduce SUM: Ngrid x 8 B) over NCCL.  SURVEY.md section 8e."""

    def __init__(self, models, models_err, models_mask, group=None, device=None):
        self.group = group
        self.rank, self.world = _world(group)
        self.lo, self.hi = shard_bounds(len(models), self.world, self.rank)
        self.eng = Engine(models[self.lo:self.hi], models_err[self.lo:self.hi], models_mask[self.lo:self.hi], device=device)
        self.dev = torch.device("cuda", self.eng.device)
        self._buf = {}

    def _tensor(self, name, shape, dtype):
        t = self._buf.get(name)
        if t is None or t.shape != torch.Size(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.dev)
            self._buf[name] = t
        return t

    def fit_predict(self, data, data_err, data_mask, model_labels, model_label_errs, label_dict=None, label_grid=None,
                    lprob_kwargs=None, kde_kwargs=None, return_best=False, as_torch=False):
        """Returns (pdfs, (lmap, levid)[, best]) identical on every rank: numpy arrays, or device tensors with
        as_torch=True (no device-to-host copy)."""
        eng, lo, hi = self.eng, self.lo, self.hi
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
        no = len(d_x)
        pmax = self._tensor("pmax", (no,), torch.float64)
        psum = self._tensor("psum", (no,), torch.float64)
        pbest = self._tensor("pbest", (no,), torch.int64)
        lib = eng.lib
        # the library works on its own stream (and synchronises it before returning): what torch has enqueued on its
        # stream - uploads, collectives - must be complete before the library reads it
        torch.cuda.current_stream(self.dev).synchronize()
        _lib.check(lib.fzb_shard_pass1_dev(eng.h, d_x.data_ptr(), d_xe.data_ptr(), d_xm.data_ptr(), no, C.byref(cfg),
                                           pmax.data_ptr(), psum.data_ptr(), pbest.data_ptr()))
        lmap, levid, best = merge_pass1(pmax, psum, pbest, lo, self.group)
        part = self._tensor("part", (no, eng.Ng), torch.float64)
        torch.cuda.current_stream(self.dev).synchronize()      # lmap / levid come out of the all-reduces on torch's stream
        _lib.check(lib.fzb_shard_pass2_dev(eng.h, d_x.data_ptr(), d_xe.data_ptr(), d_xm.data_ptr(), no, C.byref(cfg),
                                           lmap.data_ptr(), levid.data_ptr(), part.data_ptr()))
        pdfs = merge_pdfs(part, self.group)
        if as_torch:
            out = (pdfs, (lmap, levid))
            return out + (best,) if return_best else out
        out = (pdfs.cpu().numpy(), (lmap.cpu().numpy(), levid.cpu().numpy()))
        return out + (best.cpu().numpy(),) if return_best else out

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
