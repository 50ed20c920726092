"""Host-side logic of the drop-in (no GPU): keyword translation, in-place cleaning, PDFDict tables, feature maps,
kNN Monte-Carlo draw order, error behaviour that mirrors the reference."""
import numpy as np
import pytest

from conftest import golden
from oracle import fz_oracle as fo


def test_make_config_translates_reference_kwargs():
    from frankenz_b200._engine import make_config
    c = make_config(None, None)
    assert (c.free_scale, c.ignore_model_err, c.dim_prior, c.track_scale) == (0, 0, 1, 0)
    assert c.ltol == 1e-4 and c.use_wt_thresh == 1 and c.wt_thresh == 1e-3 and c.cdf_thresh == 2e-4
    c = make_config(dict(free_scale=True, ignore_model_err=True, dim_prior=False, ltol=1e-7, return_scale=True),
                    dict(wt_thresh=None, cdf_thresh=1e-3), track_scale=True)
    assert (c.free_scale, c.ignore_model_err, c.dim_prior, c.track_scale) == (1, 1, 0, 1)
    assert c.use_wt_thresh == 0 and c.use_cdf_thresh == 1 and c.cdf_thresh == 1e-3 and c.ltol == 1e-7
    # pdf.py:197 tests `ignore_model_err is not True`: a truthy non-bool still iterates
    assert make_config(dict(ignore_model_err=1)).ignore_model_err == 2
    assert make_config(None, dict(wt_thresh=None, cdf_thresh=None)).use_cdf_thresh == 0
    with pytest.raises(TypeError):
        make_config(dict(bogus=1))


def test_clean_inplace_matches_reference_rows():
    from frankenz_b200._engine import clean_inplace
    g = golden("loglike_combos.npz")
    x, xe, xm = g["data"].copy(), g["data_err"].copy(), g["data_mask"].copy()
    clean_inplace(x, xe, xm)
    assert np.array_equal(x, g["cleaned_data"]) and np.array_equal(xe, g["cleaned_err"])
    assert np.array_equal(xm, g["cleaned_mask"])


def test_pdfdict_tables_match_reference():
    from frankenz_b200.pdf import PDFDict
    g = golden("kde_edges.npz")
    d2 = PDFDict(g["d2_grid"], g["d2_sig"], sigma_trunc=4.0)
    assert np.array_equal(d2.sigma_width, g["d2_width"])
    for i in (0, 3, 5):
        assert np.array_equal(d2.sigma_dict[i], g["d2_kernel%d" % i])
        assert np.array_equal(d2.sigma_dict_cdf[i], g["d2_cdf%d" % i])
    zgrid = np.arange(0, 7 + 1e-5, 0.01)
    rd = PDFDict(zgrid, np.linspace(0.005, 2, 500))
    yi, si = rd.fit(g["labels"], g["label_errs"])
    assert np.array_equal(yi, g["y_idx"]) and np.array_equal(si, g["y_std_idx"])
    kd = fo.KernelDict(zgrid, np.linspace(0.005, 2, 500))
    assert rd.Ngrid == kd.Ngrid and rd.delta == kd.delta and np.array_equal(rd.sigma_width, kd.sigma_width)
    assert all(np.array_equal(a, b) for a, b in zip(rd.sigma_dict, kd.sigma_dict))


def test_feature_maps_match_oracle():
    from frankenz_b200 import pdf
    rs = np.random.RandomState(0)
    phot, err = rs.normal(size=(50, 5)) * 3, rs.uniform(0.1, 1, size=(50, 5))
    sky = rs.uniform(0.1, 1, size=5)
    for a, b in zip(pdf.luptitude(phot, err, skynoise=sky, zeropoints=3631.0),
                    fo.luptitude(phot, err, skynoise=sky, zeropoints=3631.0)):
        assert np.array_equal(a, b)
    for a, b in zip(pdf.magnitude(np.abs(phot) + 1, err, zeropoints=10.0), fo.magnitude(np.abs(phot) + 1, err, 10.0)):
        assert np.array_equal(a, b)
    m, me = pdf.luptitude(phot, err, skynoise=sky)
    p2, e2 = pdf.inv_luptitude(m, me, skynoise=sky)
    assert np.allclose(p2, phot, rtol=1e-10) and np.allclose(e2, err, rtol=1e-10)
    m, me = pdf.magnitude(np.abs(phot) + 1, err)
    p2, e2 = pdf.inv_magnitude(m, me)
    assert np.allclose(p2, np.abs(phot) + 1, rtol=1e-12) and np.allclose(e2, err, rtol=1e-12)


def test_vectorised_mc_draw_equals_per_object_draws():
    """NearestNeighbors draws all objects with one RandomState.normal call; the reference draws per object
    (knn.py:358).  RandomState fills in C order, so the streams are identical."""
    rs1, rs2 = np.random.RandomState(42), np.random.RandomState(42)
    x, xe = np.arange(35.0).reshape(7, 5), np.linspace(0.1, 2, 35).reshape(7, 5)
    a = rs1.normal(x, xe)
    b = np.array([rs2.normal(x[i], xe[i]) for i in range(len(x))])
    assert np.array_equal(a, b)
    assert rs1.normal() == rs2.normal()


def test_estimators_reject_python_lprob_func_and_missing_labels():
    import frankenz_b200 as fz
    bf = fz.BruteForce(np.ones((4, 5)), np.ones((4, 5)), np.ones((4, 5)))
    with pytest.raises(NotImplementedError):
        bf.fit(np.ones((2, 5)), np.ones((2, 5)), np.ones((2, 5)), lprob_func=lambda *a, **k: None, verbose=False)
    with pytest.raises(ValueError):   # bruteforce.py:265-266
        bf.predict(np.ones(4), np.ones(4), verbose=False)
    with pytest.raises(ValueError):   # bruteforce.py:267-269 (no fits, no weights)
        bf.predict(np.ones(4), np.ones(4), label_grid=np.linspace(0, 1, 11), verbose=False)
    with pytest.raises(IndexError):   # track_scale without return_scale: results[5] of a 5-tuple
        bf.fit(np.ones((2, 5)), np.ones((2, 5)), np.ones((2, 5)), track_scale=True, verbose=False)
    assert fz.fitting.BruteForce is fz.BruteForce and fz.fitting.NearestNeighbors is fz.NearestNeighbors


def test_ordered_union_and_exact_knn_oracle():
    idx = np.array([[5, 3, 9], [3, 7, 5], [11, 9, 0]])
    assert fo.ordered_unique(idx).tolist() == [5, 3, 9, 7, 11, 0]
    feats = np.array([[[0., 0.], [1., 0.], [0., 2.], [1., 0.]]], dtype=np.float32)
    i, d = fo.knn_query_exact(feats, np.array([0.9, 0.0]), 3, 2)
    assert i.tolist() == [[1, 3, 0]] and np.allclose(d[0], [0.1, 0.1, 0.9])     # tie -> lowest index first


def test_clean_inplace_threaded_helper_matches_numpy_rule():
    """pdf.py:310-311 on large float64 batches runs in libfzb200 (host threads, no device): same arrays afterwards as the
    numpy form used for small / non-float64 inputs."""
    from frankenz_b200._engine import clean_inplace
    rs = np.random.RandomState(0)
    x = rs.normal(size=(80000, 5))
    xe = np.abs(rs.normal(size=x.shape)) + 0.1
    xm = np.ones_like(x)
    x[rs.uniform(size=x.shape) < 0.01] = np.nan
    xe[rs.uniform(size=x.shape) < 0.01] = -1.0
    x[3, 2], xe[4, 1], xe[5, 0] = np.inf, np.nan, 0.0
    a, b, c = x.copy(), xe.copy(), xm.copy()
    clean_inplace(a, b, c)                                   # 400k entries: threaded helper
    bad = ~(np.isfinite(x) & np.isfinite(xe) & (xe > 0))
    assert np.all(a[bad] == 0) and np.all(b[bad] == 1) and np.all(c[bad] == 0)
    assert np.array_equal(a[~bad], x[~bad]) and np.array_equal(b[~bad], xe[~bad]) and np.all(c[~bad] == 1)
    a2, b2, c2 = x[:100].copy(), xe[:100].copy(), xm[:100].astype(bool)       # small + bool mask: numpy form
    clean_inplace(a2, b2, c2)
    assert np.array_equal(a2, a[:100]) and np.array_equal(b2, b[:100]) and np.array_equal(c2, c[:100].astype(bool))


def test_page_pool_recycles_buffers_only_when_every_view_is_gone():
    """PagePool (_engine.py): the large pageable output arrays are handed out again once the array AND its views have been
    garbage collected; small requests and a zero budget fall back to numpy.empty."""
    import gc
    from frankenz_b200._engine import PagePool
    pool = PagePool(max_bytes=1 << 30)
    a = pool.empty((100000, 120))                    # 96 MB: pooled
    addr = a.ctypes.data
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"] and pool.total == a.nbytes
    a[:] = 3.0
    view = a[10:20]
    del a
    gc.collect()
    assert not pool.free and view[0, 0] == 3.0       # a view keeps the block out of the pool
    del view
    gc.collect()
    assert len(pool.free) == 1
    b = pool.empty((100000, 120))
    assert b.ctypes.data == addr and not pool.free   # same memory, already mapped
    c = pool.empty((100000, 120))                    # a second live array gets its own block
    assert c.ctypes.data != addr and pool.total == 2 * b.nbytes
    small = pool.empty((10, 10))
    assert small.base is None and pool.total == 2 * b.nbytes
    assert PagePool(max_bytes=0).empty((100000, 120)).base is None


def test_loss_matrix_cache_is_keyed_by_kernel_and_grid():
    """pdf._loss_matrix: named kernels on the default argument are cached per grid (read-only, most recent first, at
    most 4); callables and user kernel grids are evaluated on every call like the reference does (pdf.py:1003-1023)."""
    from frankenz_b200 import pdf
    del pdf._LOSS_CACHE[:]
    g = np.linspace(0., 6., 61)
    a = pdf._loss_matrix(g, "lorentz", None)
    assert np.array_equal(a, 1.0 - pdf._loss_kernel(g, "lorentz", None)) and not a.flags.writeable
    assert pdf._loss_matrix(g.copy(), "lorentz", None) is a                   # same values, another array object
    assert pdf._loss_matrix(g, "gaussian", None) is not a
    g2 = g.copy()
    g2[7] += 1e-9
    b = pdf._loss_matrix(g2, "lorentz", None)
    assert b is not a and not np.array_equal(a, b)
    calls = []

    def kern(x):
        calls.append(1)
        return np.exp(-np.abs(x))
    pdf._loss_matrix(g, kern, None)
    pdf._loss_matrix(g, kern, None)
    assert len(calls) == 2
    kg = np.subtract.outer(g, g)
    assert pdf._loss_matrix(g, "tophat", kg) is not pdf._loss_matrix(g, "tophat", kg)
    for i in range(6):
        pdf._loss_matrix(g + i, "lorentz", None)
    assert len(pdf._LOSS_CACHE) == 4
    with pytest.raises(RuntimeError):
        pdf._loss_matrix(g, "no such kernel", None)
