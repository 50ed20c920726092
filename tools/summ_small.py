"""Small pdfs_summarize problems for compute-sanitizer (memcheck / racecheck): the 701-point grid with a ragged last tile,
a grid above 704 points (16 objects per CTA) and a tiny grid, checked against the committed reference outputs where they exist."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import frankenz_b200 as fz  # noqa: E402

g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "pdfs_summarize.npz"))
res = fz.pdf.pdfs_summarize(g["pdfs"].copy(), g["zgrid"], rstate=np.random.RandomState(5), pkern="lorentz")
assert np.array_equal(res[1][0], g["lorentz_med"]) and np.array_equal(res[5], g["lorentz_mc"])
rs = np.random.RandomState(2)
for n, ng in ((77, 701), (45, 900), (37, 40)):
    zg = np.linspace(0, 6, ng)
    mu, sg = rs.uniform(0.5, 5, n), rs.uniform(0.05, 0.6, n)
    p = np.exp(-0.5 * ((zg[None, :] - mu[:, None]) / sg[:, None]) ** 2)
    out = fz.pdf.pdfs_summarize(p, zg, rstate=np.random.RandomState(1))
    mean = (p * zg[None, :]).sum(axis=1)
    assert np.allclose(out[0][0], mean, rtol=1e-12, atol=1e-12), (n, ng)
    assert np.all(np.isfinite(out[3][3]))
print("summ_small ok")
