#!/usr/bin/env python
"""Golden vectors for the SOM / GNG node-fit stage, produced by the UNMODIFIED reference:
a SelfOrganizingMap trained on the committed SDSS mock (networks.py:1517-1867; the sequential training itself is out of
scope and only supplies `nodes`), then the two stages that score photometry against the nodes through `logprob`:

  populate_network (networks.py:246-356)  models x nodes free-scale fit, BMU, weight threshold, normalised ln-weights
  fit              (networks.py:782-936)  objects x nodes fit, threshold, node selection (`nodes_only=True`) and the
                                          union of the models mapped to the selected nodes + their fits

Runs only in the build container (needs /root/reference)."""
import os
import sys
import warnings

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
sys.path.insert(0, "/root/reference")
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def ragged(lists, dtype):
    off = np.zeros(len(lists) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(v) for v in lists])
    flat = np.concatenate([np.asarray(v, dtype=dtype) for v in lists]) if off[-1] else np.zeros(0, dtype=dtype)
    return off, flat


def main():
    from frankenz.networks import SelfOrganizingMap
    d = np.load(os.path.join(HERE, "sdss_cww_mock.npz"))
    phot, err = d["phot_obs"], d["phot_err"]
    nm, no = 1500, 60
    m, me, mm = phot[:nm].copy(), err[:nm].copy(), np.ones((nm, 5))
    x, xe, xm = phot[3000:3000 + no].copy(), err[3000:3000 + no].copy(), np.ones((no, 5))
    xm[3, 1] = 0.0
    x[7, 4] = np.nan
    som = SelfOrganizingMap(m, me, mm)
    som.train_network(nside=8, niter=300, nbatch=20, rstate=np.random.RandomState(11), verbose=False)
    out = dict(models=m, models_err=me, models_mask=mm, data=x, data_err=xe, data_mask=xm, nodes=som.nodes.copy())
    som.populate_network(verbose=False)
    for name in ("nodes_idxs", "nodes_bmus"):
        out[name + "_off"], out[name] = ragged(getattr(som, name), np.int64)
    for name in ("nodes_logwts", "nodes_scales", "nodes_scales_err"):
        out[name + "_off"], out[name] = ragged(getattr(som, name), np.float64)
    out["nodes_Nmatch"] = som.nodes_Nmatch.copy()
    out["models_lmap"], out["models_levid"] = som.models_lmap.copy(), som.models_levid.copy()
    kw = dict(free_scale=True, ignore_model_err=True, dim_prior=True)
    for tag, nodes_only in (("nodes", True), ("full", False)):
        som.fit(x.copy(), xe.copy(), xm.copy(), nodes_only=nodes_only, lprob_kwargs=kw, verbose=False)
        out[tag + "_Nneighbors"] = np.array(som.Nneighbors)
        out[tag + "_neighbors_off"], out[tag + "_neighbors"] = ragged(som.neighbors, np.int64)
        out[tag + "_lnprob_off"], out[tag + "_lnprob"] = ragged(som.fit_lnprob, np.float64)
        out[tag + "_chi2_off"], out[tag + "_chi2"] = ragged(som.fit_chi2, np.float64)
    np.savez_compressed(os.path.join(HERE, "som_nodefit.npz"), **out)
    print("som_nodefit ok: nodes", som.nodes.shape, "matches", int(som.nodes_Nmatch.sum()),
          "Nneighbors", out["full_Nneighbors"].min(), out["full_Nneighbors"].max())


if __name__ == "__main__":
    main()
