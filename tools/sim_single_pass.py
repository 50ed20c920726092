"""CPU simulation behind the fused single pass of the tensor-core sweep: for C3 objects by S/N class, how
far the maximum over every r-th model (M0) lies below the true maximum, how many weights fall in the recorded band above the
running cut, and how many objects would have to take pass 2 (M_f - M0 > band).  numpy only; run with PYTHONPATH=.  """
import numpy as np, bench_data
models, labels, depth = bench_data.c3_models()
x, xe, xm, j, mag = bench_data.c3_objects(30000, models, depth, seed=3)
snr = np.sqrt(np.sum((x/xe)**2, axis=1))
lnthr = np.log(1e-3)
def lnprob(xo, eo):
    w = 1/eo**2
    B = models @ (w*xo); C = (models**2) @ w
    chi2 = np.maximum(np.sum(w*xo*xo) - B*B/C, 1e-12)
    return np.log(chi2) - 0.5*chi2
for lo_, hi_ in ((0,10),(10,30),(30,100),(100,1000)):
    sel = np.where((snr>lo_)&(snr<=hi_))[0][:80]
    for r in (16,64):
      for g in (0.01,0.03,0.1):
        nrec=[]; redo=[]
        for o in sel:
            l = lnprob(x[o], xe[o]); Mf=l.max(); M0=l[::r].max()
            run = np.maximum.accumulate(np.maximum(l, M0)); runp = np.concatenate([[M0], run[:-1]])
            cut = runp + lnthr
            rec = (l > cut - 3e-5) & (l <= cut + g)
            nrec.append(rec.sum()); redo.append(Mf-M0 > g)
        nrec=np.array(nrec); redo=np.array(redo)
        print('S/N(%d,%d] r=%d g=%.2f: records/object mean %.0f p90 %.0f max %d; redo(M_f-M0>g) %.2f; records of non-redo mean %.0f max %d' % (lo_,hi_,r,g,nrec.mean(),np.percentile(nrec,90),nrec.max(),redo.mean(), nrec[~redo].mean() if (~redo).any() else -1, nrec[~redo].max() if (~redo).any() else -1))
