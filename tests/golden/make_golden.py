#!/usr/bin/env python
"""Golden vectors produced by the UNMODIFIED reference (frankenz v0.3.5).

Run in the build container only (imports /root/reference read-only):

    python tests/golden/make_mock_inputs.py      # mock photometry (slow, once)
    python tests/golden/make_golden.py           # reference outputs -> *.npz

The reference has no tests of its own, so these files are
what pins the oracle (`oracle/fz_oracle.py`) and, through it, the CUDA path.
Cases follow SURVEY.md section 8c (i)-(vi).
"""
import os
import sys
import warnings

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
from frankenz import pdf as rpdf  # noqa: E402
from frankenz.bruteforce import BruteForce  # noqa: E402
from frankenz.knn import NearestNeighbors  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
COMBOS = [(fs, ime, dp) for fs in (False, True) for ime in (False, True) for dp in (False, True)]


def tag(fs, ime, dp):
    return "fs%d_ime%d_dp%d" % (fs, ime, dp)


def load_mock():
    d = np.load(os.path.join(HERE, "sdss_cww_mock.npz"))
    return d["phot_obs"], d["phot_err"], d["redshifts"]


def case_loglike():
    """(i)+(ii): all 8 flag combos, random band dropouts, dirty entries."""
    phot, err, z = load_mock()
    rs = np.random.RandomState(101)
    nm, nd = 300, 24
    m, me = phot[:nm].copy(), err[:nm].copy()
    mm = (rs.uniform(size=m.shape) > 0.08).astype(float)
    x, xe = phot[2000:2000 + nd].copy(), err[2000:2000 + nd].copy()
    xm = (rs.uniform(size=x.shape) > 0.10).astype(float)
    # dirty entries: NaN flux, inf flux, zero / negative / NaN error
    x[1, 0], x[2, 3], xe[3, 1], xe[4, 2], xe[5, 4] = np.nan, np.inf, 0.0, -1.0, np.nan
    out = dict(models=m, models_err=me, models_mask=mm, data=x, data_err=xe, data_mask=xm)
    for fs, ime, dp in COMBOS:
        res = []
        xc, xec, xmc = x.copy(), xe.copy(), xm.copy()
        for i in range(nd):
            r = rpdf.loglike(xc[i], xec[i], xmc[i], m, me, mm, free_scale=fs, ignore_model_err=ime,
                             dim_prior=dp, ltol=1e-4, return_scale=True)
            res.append(r)
        t = tag(fs, ime, dp)
        out[t + "_lnl"] = np.array([r[0] for r in res])
        out[t + "_ndim"] = np.array([r[1] for r in res])
        out[t + "_chi2"] = np.array([r[2] for r in res])
        if fs:
            out[t + "_scale"] = np.array([r[3] for r in res])
            out[t + "_scale_err"] = np.array([r[4] for r in res])
        out["cleaned_data"], out["cleaned_err"], out["cleaned_mask"] = xc, xec, xmc
    np.savez_compressed(os.path.join(HERE, "loglike_combos.npz"), **out)
    print("loglike_combos ok")


def case_degenerate():
    """(iii): self-match chi2=0, Ndim in {0,1,2}, all-masked model, zero-flux model."""
    phot, err, z = load_mock()
    nm = 40
    m, me = phot[:nm].copy(), err[:nm].copy()
    mm = np.ones_like(m)
    mm[1] = 0.0                       # all-masked model
    mm[2, 1:] = 0.0                   # one band
    mm[3, 2:] = 0.0                   # two bands
    m[4] = 0.0                        # zero-flux model (shape = 0 when errors ignored)
    x = np.stack([phot[0], phot[5], phot[6], phot[7]]).copy()   # row 0 == model 0 (self match)
    xe = np.stack([err[0], err[5], err[6], err[7]]).copy()
    xm = np.ones_like(x)
    xm[2, :3] = 0.0                   # object with 2 usable bands
    xm[3, :] = 0.0                    # object with nothing
    out = dict(models=m, models_err=me, models_mask=mm, data=x, data_err=xe, data_mask=xm)
    for fs, ime, dp in COMBOS:
        res = []
        for i in range(len(x)):
            r = rpdf.loglike(x[i].copy(), xe[i].copy(), xm[i].copy(), m, me, mm, free_scale=fs,
                             ignore_model_err=ime, dim_prior=dp, return_scale=True)
            res.append(r)
        t = tag(fs, ime, dp)
        out[t + "_lnl"] = np.array([r[0] for r in res])
        out[t + "_ndim"] = np.array([r[1] for r in res])
        out[t + "_chi2"] = np.array([r[2] for r in res])
        if fs:
            out[t + "_scale"] = np.array([r[3] for r in res])
            out[t + "_scale_err"] = np.array([r[4] for r in res])
    np.savez_compressed(os.path.join(HERE, "loglike_degenerate.npz"), **out)
    print("loglike_degenerate ok")


def case_bruteforce():
    """(vi) scaled-down config C1: fit + predict(dict / grid / logwt=lnlike) + fit_predict."""
    phot, err, z = load_mock()
    ntr, nte = 1200, 32
    m, me, mm = phot[:ntr].copy(), err[:ntr].copy(), np.ones((ntr, 5))
    x, xe, xm = phot[3000:3000 + nte].copy(), err[3000:3000 + nte].copy(), np.ones((nte, 5))
    lab, labe = z[:ntr].copy(), np.full(ntr, 0.05)
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    rdict = rpdf.PDFDict(zgrid, np.linspace(0.005, 2, 500))
    out = dict(models=m, models_err=me, models_mask=mm, data=x, data_err=xe, data_mask=xm,
               labels=lab, label_errs=labe, zgrid=zgrid)
    bf = BruteForce(m, me, mm)
    bf.fit(x.copy(), xe.copy(), xm.copy(), verbose=False)
    out["fit_lnprob"], out["fit_chi2"], out["fit_Ndim"] = bf.fit_lnprob, bf.fit_chi2, bf.fit_Ndim
    p, (lm, le) = bf.predict(lab, labe, label_dict=rdict, return_gof=True, verbose=False)
    out["pdf_dict"], out["lmap"], out["levid"] = p, lm, le
    p = bf.predict(lab, labe, label_grid=zgrid, verbose=False)
    out["pdf_grid"] = p
    # heteroscedastic label errors -> several dictionary widths
    labe2 = 0.01 * (1.0 + lab)
    out["label_errs2"] = labe2
    out["pdf_dict_mixed"] = bf.predict(lab, labe2, label_dict=rdict, verbose=False)
    out["pdf_grid_mixed"] = bf.predict(lab, labe2, label_grid=zgrid, verbose=False)
    # thresholds: none, cdf rule
    out["pdf_dict_nothresh"] = bf.predict(lab, labe, label_dict=rdict, verbose=False,
                                          kde_kwargs=dict(wt_thresh=None, cdf_thresh=None))
    out["pdf_dict_cdf"] = bf.predict(lab, labe, label_dict=rdict, verbose=False,
                                     kde_kwargs=dict(wt_thresh=None, cdf_thresh=2e-4))
    out["pdf_grid_cdf"] = bf.predict(lab, labe, label_grid=zgrid, verbose=False,
                                     kde_kwargs=dict(wt_thresh=None, cdf_thresh=2e-4))
    # fit_predict for each likelihood flavour (save_fits=False)
    for fs, ime, dp in COMBOS:
        bf2 = BruteForce(m, me, mm)
        p, (lm, le) = bf2.fit_predict(x.copy(), xe.copy(), xm.copy(), lab, labe, label_dict=rdict,
                                      lprob_kwargs=dict(free_scale=fs, ignore_model_err=ime, dim_prior=dp),
                                      return_gof=True, verbose=False, save_fits=False)
        t = tag(fs, ime, dp)
        out[t + "_pdf"], out[t + "_lmap"], out[t + "_levid"] = p, lm, le
    np.savez_compressed(os.path.join(HERE, "bruteforce_c1small.npz"), **out)
    print("bruteforce_c1small ok")


def case_kde_edges():
    """(iv): labels on / beyond the grid edges, sigma=0, clipped sigma, tiny grids."""
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    rdict = rpdf.PDFDict(zgrid, np.linspace(0.005, 2, 500))
    lab = np.array([0.0, 0.02, 0.004, 0.005, 0.015, 6.99, 7.0, 3.0, 3.004999, 3.005, 0.1, 6.9, 1.234, 5.5])
    labe = np.array([0.05, 0.05, 0.0, 0.007, 0.003, 0.05, 0.05, 0.0, 0.2, 0.05, 0.3, 0.12, 0.05, 0.6])
    rs = np.random.RandomState(5)
    lw = rs.normal(size=(6, len(lab))) * 2.0
    lw[1, :] = -np.inf
    lw[1, 3] = 0.0                    # single finite weight
    lw[2, 7] = 50.0                   # dominant weight
    out = dict(labels=lab, label_errs=labe, logwt=lw, zgrid=zgrid)
    bf = BruteForce(np.ones((len(lab), 5)), np.ones((len(lab), 5)), np.ones((len(lab), 5)))
    bf.NDATA = len(lw)
    yi, si = rdict.fit(lab, labe)
    out["y_idx"], out["y_std_idx"] = yi, si
    p, (lm, le) = bf.predict(lab, labe, label_dict=rdict, logwt=lw, return_gof=True, verbose=False)
    out["pdf_dict"], out["lmap"], out["levid"] = p, lm, le
    out["pdf_grid"] = bf.predict(lab, labe + 1e-3, label_grid=zgrid, logwt=lw, verbose=False)
    out["label_errs_grid"] = labe + 1e-3
    # kernel dictionary internals for a coarse dict (pins widths / cdf / wrap-around bug region)
    d2 = rpdf.PDFDict(np.linspace(-1, 1, 41), np.linspace(0.01, 0.6, 12), sigma_trunc=4.0)
    out["d2_grid"], out["d2_sig"] = d2.grid, d2.sigma_grid
    out["d2_width"] = d2.sigma_width
    for i in (0, 3, 5):
        out["d2_kernel%d" % i] = d2.sigma_dict[i]
        out["d2_cdf%d" % i] = d2.sigma_dict_cdf[i]
    np.savez_compressed(os.path.join(HERE, "kde_edges.npz"), **out)
    print("kde_edges ok")


def case_knn():
    """(v): exact (eps=0) KMCkNN with fixed RandomStates."""
    phot, err, z = load_mock()
    ntr, nte = 1500, 48
    m, me, mm = phot[:ntr].copy(), err[:ntr].copy(), np.ones((ntr, 5))
    x, xe, xm = phot[3200:3200 + nte].copy(), err[3200:3200 + nte].copy(), np.ones((nte, 5))
    lab, labe = z[:ntr].copy(), np.full(ntr, 0.05)
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    rdict = rpdf.PDFDict(zgrid, np.linspace(0.005, 2, 500))
    depth = np.load(os.path.join(HERE, "sdss_cww_mock.npz"))["depth_flux1sig"]
    fkw = dict(skynoise=depth, zeropoints=10 ** (-0.4 * -23.9))
    out = dict(models=m, models_err=me, models_mask=mm, data=x, data_err=xe, data_mask=xm,
               labels=lab, label_errs=labe, zgrid=zgrid, skynoise=depth,
               zeropoints=np.float64(10 ** (-0.4 * -23.9)))
    for name, K, k, fmap in (("a", 5, 20, "luptitude"), ("b", 1, 1, "luptitude"), ("c", 8, 7, "identity"),
                             ("d", 3, 25, "magnitude")):
        kw = fkw if fmap == "luptitude" else (dict(zeropoints=fkw["zeropoints"]) if fmap == "magnitude" else {})
        if fmap == "magnitude":
            # magnitudes need positive fluxes: use noiseless-ish bright subset
            mu, meu = np.abs(m) + 5 * me, me * 1e-3
            xu = np.abs(x) + 5 * xe
            xeu = xe * 1e-3
        else:
            mu, meu, xu, xeu = m, me, x, xe
        nn = NearestNeighbors(mu, meu, mm, K=K, feature_map=fmap, fmap_kwargs=kw,
                              rstate=np.random.RandomState(1), verbose=False)
        out[name + "_feats"] = np.array([np.asarray(T.data, dtype=np.float32) for T in nn.KDTrees])
        p, (lm, le) = nn.fit_predict(xu.copy(), xeu.copy(), xm.copy(), lab, labe, label_dict=rdict, k=k, eps=0.0,
                                     rstate=np.random.RandomState(2), return_gof=True, verbose=False)
        out[name + "_pdf"], out[name + "_lmap"], out[name + "_levid"] = p, lm, le
        out[name + "_neighbors"], out[name + "_Nneighbors"] = nn.neighbors, nn.Nneighbors
        out[name + "_lnprob"], out[name + "_chi2"] = nn.fit_lnprob, nn.fit_chi2
        p2 = nn.predict(lab, labe, label_grid=zgrid, verbose=False)
        out[name + "_pdf_grid"] = p2
    np.savez_compressed(os.path.join(HERE, "knn_exact.npz"), **out)
    print("knn_exact ok")


def case_fs1_iters():
    """Iterated free-scale mode: iteration counts the reference actually takes (hard part 3)."""
    phot, err, z = load_mock()
    m, me, mm = phot[:800].copy(), err[:800].copy(), np.ones((800, 5))
    x, xe, xm = phot[3500:3516].copy(), err[3500:3516].copy(), np.ones((16, 5))
    out = dict(models=m, models_err=me, models_mask=mm, data=x, data_err=xe, data_mask=xm)
    for ltol in (1e-2, 1e-4, 1e-7):
        lnl = np.array([rpdf.loglike(x[i].copy(), xe[i].copy(), xm[i].copy(), m, me, mm, free_scale=True,
                                     dim_prior=False, ltol=ltol)[0] for i in range(len(x))])
        out["lnl_ltol%g" % ltol] = lnl
    np.savez_compressed(os.path.join(HERE, "fs1_ltol.npz"), **out)
    print("fs1_ltol ok")



def mock_pdfs(n, zgrid, seed):
    """Seeded PDFs on `zgrid`: Gaussian mixtures (narrow, broad, multi-modal), exact zeros between the peaks
    (CDF plateaus), a single-bin spike, a flat row, un-normalised."""
    rs = np.random.RandomState(seed)
    p = np.zeros((n, len(zgrid)))
    for i in range(n):
        for _ in range(rs.randint(1, 4)):
            mu, sg, a = rs.uniform(0.05, 6.5), 10 ** rs.uniform(-2, -0.2), rs.uniform(0.2, 1.0)
            g = a * np.exp(-0.5 * ((zgrid - mu) / sg) ** 2)
            g[g < 1e-6 * g.max()] = 0.0
            p[i] += g
    p[0] = 0.0
    p[0, min(137, len(zgrid) - 3)] = 3.0
    p[1] = 1.0
    p[2, :min(50, len(zgrid) // 3)] = 0.0
    return p * rs.uniform(0.5, 20.0, size=(n, 1))


def case_summarize():
    """pdf.pdfs_summarize with the three built-in loss kernels, a custom kernel, a custom
    confidence width, with / without renormalisation; pdf.pdfs_resample."""
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    out = dict(zgrid=zgrid)
    names = ["mean", "med", "mode", "best"]

    def store(tag, res, pdfs_after):
        for k, nme in enumerate(names):
            for j, q in enumerate(("", "_std", "_conf", "_risk")):
                out["%s_%s%s" % (tag, nme, q)] = res[k][j]
        for j, q in enumerate(("low95", "low68", "high68", "high95")):
            out["%s_%s" % (tag, q)] = res[4][j]
        out[tag + "_mc"] = res[5]
        out[tag + "_pdfs_after"] = pdfs_after

    p0 = mock_pdfs(96, zgrid, 11)
    out["pdfs"] = p0
    for kern in ("lorentz", "gaussian", "tophat"):
        p = p0.copy()
        store(kern, rpdf.pdfs_summarize(p, zgrid, rstate=np.random.RandomState(5), pkern=kern), p)
    p = p0 / p0.sum(axis=1)[:, None]
    out["pdfs_normed"] = p.copy()
    store("noren", rpdf.pdfs_summarize(p, zgrid, renormalize=False, rstate=np.random.RandomState(6)), p)
    p = p0.copy()
    store("custom", rpdf.pdfs_summarize(p, zgrid, rstate=np.random.RandomState(7),
                                        pkern=lambda x: np.exp(-np.abs(x)), wconf_func=lambda z: 0.02 + 0.05 * z * z), p)
    # a coarse grid (Ngrid not a multiple of anything) and a user kernel argument grid
    g2 = np.linspace(0.0, 3.0, 61)
    q0 = mock_pdfs(40, g2 * 7 / 3, 12)
    out["grid2"], out["pdfs2"] = g2, q0
    kg = (g2.reshape(-1, 1) - g2.reshape(1, -1)) / 0.1
    out["kgrid2"] = kg
    q = q0.copy()
    store("grid2", rpdf.pdfs_summarize(q, g2, rstate=np.random.RandomState(8), pkern="gaussian", pkern_grid=kg), q)
    new_grid = np.linspace(-0.5, 7.5, 333)
    out["resample_grid"] = new_grid
    out["resampled"] = rpdf.pdfs_resample(p0.copy(), zgrid, new_grid)
    out["resampled_noren"] = rpdf.pdfs_resample(p0.copy(), zgrid, new_grid, renormalize=False, left=0.5, right=0.25)
    np.savez_compressed(os.path.join(HERE, "pdfs_summarize.npz"), **out)
    print("pdfs_summarize.npz", len(out), "arrays")

def case_loglike_nz():
    """samplers.loglike_nz (samplers.py:24-76) on seeded PDFs: plain, with a perturbed pair of bins,
    with overlaps returned, and for an invalid N(z)."""
    from frankenz import samplers
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    p = mock_pdfs(700, zgrid, 91)
    p[0] = p[3]                                   # the single-bin spike of row 0 would give overlap 0 with most N(z)
    p /= p.sum(axis=1)[:, None]
    rs = np.random.RandomState(92)
    out = dict(zgrid=zgrid, pdfs=p)
    nzs = []
    for k in range(4):
        nz = rs.gamma(2.0, size=len(zgrid)) * np.exp(-0.5 * ((zgrid - 1.0 - 0.4 * k) / (0.7 + 0.2 * k)) ** 2) + 1e-4
        nzs.append(nz / nz.sum())
    nzs = np.array(nzs)
    out["nz"] = nzs
    out["lnlike"] = np.array([samplers.loglike_nz(nz, p) for nz in nzs])
    ll, ov = samplers.loglike_nz(nzs[1], p, return_overlap=True)
    out["lnlike_ov"], out["overlap"] = ll, ov
    ll, ov = samplers.loglike_nz(nzs[2], p, return_overlap=True, pair=(120, 260), pair_step=3e-4)
    out["lnlike_pair"], out["overlap_pair"] = ll, ov
    bad = nzs[0].copy()
    bad[5] = -1e-3
    ll, ov = samplers.loglike_nz(bad, p, return_overlap=True)
    out["lnlike_bad"], out["overlap_bad"], out["nz_bad"] = ll, ov, bad
    np.savez_compressed(os.path.join(HERE, "loglike_nz.npz"), **out)
    print("loglike_nz ok", out["lnlike"])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "nz":
        case_loglike_nz()
        sys.exit(0)
    case_loglike()
    case_degenerate()
    case_bruteforce()
    case_kde_edges()
    case_knn()
    case_fs1_iters()
    case_summarize()
    case_loglike_nz()
