"""Where the arg-max of a model-sharded run differs from the unsharded run of the same problem (bench.py's
`best_index_mismatch_fraction`): W shards emulated on ONE GPU (the calls ModelShardedBruteForce makes, with the
merge kernel in place of the all-gather), the unsharded run on the same GPU, and for every object whose index differs
the float64 log-posterior of BOTH picks computed here in numpy.  Prints, per route (fp32 sweep / float64 sweep), how many
differ and by how much the two picks differ in float64.

    python tools/diag_shard_best.py [n_train] [n_obj] [W]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    import torch
    import bench_data
    import frankenz_b200 as fz
    from frankenz_b200 import _lib
    from frankenz_b200._engine import Engine, make_config
    from frankenz_b200.distributed import shard_bounds
    n_train = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
    W = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    tr, tre, trm, ztr, x, xe, xm = bench_data.c5_dataset(n_train, n)
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(n_train, 0.05)
    tx = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (x, xe, xm)]
    cfg = make_config({}, None)
    gathered = torch.empty((W, 3, n), dtype=torch.float64).cuda()
    route = np.zeros(n, dtype=np.int64)
    for r in range(W):
        lo, hi = shard_bounds(n_train, W, r)
        e = Engine(tr[lo:hi], tre[lo:hi], trm[lo:hi])
        e.set_kde(ztr[lo:hi], labe[lo:hi], label_dict=rdict)
        _lib.check(e.lib.fzb_shard_pass1_packed_dev(e.h, tx[0].data_ptr(), tx[1].data_ptr(), tx[2].data_ptr(), n,
                                                    C.byref(cfg), lo, gathered[r].data_ptr()))
        torch.cuda.synchronize()
        st = e.stats()
        print("shard %d: objects_fp64 %d pairs_fp32 %.3g pairs_fp64 %.3g" % (r, st["objects_fp64"], st["pairs_fp32"],
                                                                               st["pairs_fp64"]))
        if r == W - 1:
            lmap = torch.empty(n, dtype=torch.float64).cuda()
            levid = torch.empty(n, dtype=torch.float64).cuda()
            best = torch.empty(n, dtype=torch.int64).cuda()
            _lib.check(e.lib.fzb_shard_merge_dev(e.h, gathered.data_ptr(), W, n, lmap.data_ptr(), levid.data_ptr(),
                                                 best.data_ptr()))
            _lib.check(e.lib.fzb_synchronize(e.h))
        del e
    best = best.cpu().numpy()
    lmap = lmap.cpu().numpy()
    res = {}
    for prec in ("mixed", "fp64"):
        bf = fz.BruteForce(tr, tre, trm)
        kw = dict(precision="fp64") if prec == "fp64" else {}
        p1, (lm1, le1) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), ztr, labe, label_dict=rdict, return_gof=True,
                                        verbose=False, save_fits=False, lprob_kwargs=kw)
        st = bf._eng().stats()
        print("unsharded %s: objects_fp64 %d objects_fused %d" % (prec, st["objects_fp64"], st["objects_fused"]))
        res[prec] = (lm1.copy(), bf.best_idx.copy())
        del bf
    snr = np.sqrt(np.sum((x / xe) ** 2, axis=1))
    for name, (lm1, b1) in res.items():
        for tag, b in (("sharded", best), ("unsharded mixed", res["mixed"][1])):
            if name == "mixed" and tag != "sharded":
                continue
            bad = np.flatnonzero(b != b1)
            print("\n%s vs unsharded %s: %d of %d indices differ (%.3f)" % (tag, name, len(bad), n, len(bad) / n))
            if len(bad) == 0:
                continue
            la = lnpost_pairs(x[bad], xe[bad], tr[b[bad]], tre[b[bad]])
            lb = lnpost_pairs(x[bad], xe[bad], tr[b1[bad]], tre[b1[bad]])
            d = la - lb
            print("   float64 lnpost(%s pick) - lnpost(unsharded %s pick): min %.3g median %.3g max %.3g" %
                  (tag, name, d.min(), np.median(d), d.max()))
            print("   relative to |lmap|: median %.3g max %.3g" % (np.median(np.abs(d) / np.maximum(1, np.abs(lb))),
                                                                   np.max(np.abs(d) / np.maximum(1, np.abs(lb)))))
            print("   total S/N of the differing objects: median %.3g (all objects: %.3g); |lmap| median %.3g" %
                  (np.median(snr[bad]), np.median(snr), np.median(np.abs(lb))))
            print("   index distance |i-j|: median %d; duplicates (identical rows): %d" %
                  (np.median(np.abs(b[bad] - b1[bad])), int(np.sum(np.all(tr[b[bad]] == tr[b1[bad]], axis=1)))))


def lnpost_pairs(x, xe, m, me):
    """default likelihood with dim_prior (frankenz/pdf.py:79-93), one (object, model) pair per row, float64"""
    from scipy.special import gammaln, xlogy
    var = xe * xe + me * me
    r = x - m
    chi2 = np.sum(r * r / var, axis=1)
    a = 0.5 * x.shape[1]
    return xlogy(a - 1.0, chi2) - 0.5 * chi2 - gammaln(a) - np.log(2.0) * a


if __name__ == "__main__":
    main()
