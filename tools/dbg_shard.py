import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data, frankenz_b200 as fz
from frankenz_b200 import _lib
from frankenz_b200._engine import Engine, make_config
from frankenz_b200.distributed import shard_bounds
n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
models, labels, depth = bench_data.c3_models()
x, xe, xm, _, _ = bench_data.c3_objects(n, models, depth, seed=99)
zgrid, sig = bench_data.c3_kde(); rdict = fz.pdf.PDFDict(zgrid, sig)
labe = np.full(len(models), 0.05)
kw = dict(free_scale=True, ignore_model_err=True, dim_prior=True)
cfg = make_config(kw, None)
tx = [torch.from_numpy(a).cuda() for a in (x, xe, xm)]
W = 2
engs, parts = [], []
for r in range(W):
    lo, hi = shard_bounds(len(models), W, r)
    e = Engine(models[lo:hi], np.zeros((hi - lo, 5)), np.ones((hi - lo, 5)))
    e.set_kde(labels[lo:hi], labe[lo:hi], label_dict=rdict)
    pm = torch.empty(n, dtype=torch.float64).cuda(); ps = torch.empty(n, dtype=torch.float64).cuda(); pb = torch.empty(n, dtype=torch.int64).cuda()
    _lib.check(e.lib.fzb_shard_pass1_dev(e.h, tx[0].data_ptr(), tx[1].data_ptr(), tx[2].data_ptr(), n, C.byref(cfg), pm.data_ptr(), ps.data_ptr(), pb.data_ptr()))
    print("shard", r, "pass1", e.stats())
    engs.append(e); parts.append((pm, ps, pb + lo))
gmax = torch.stack([p[0] for p in parts]).max(dim=0).values
s = sum(ps * torch.exp(pm - gmax) for pm, ps, _ in parts)
levid = gmax + torch.log(s)
pp = []
for r, e in enumerate(engs):
    part = torch.empty((n, 701), dtype=torch.float64).cuda()
    _lib.check(e.lib.fzb_shard_pass2_dev(e.h, tx[0].data_ptr(), tx[1].data_ptr(), tx[2].data_ptr(), n, C.byref(cfg), gmax.data_ptr(), levid.data_ptr(), part.data_ptr()))
    print("shard", r, "pass2", e.stats())
    pp.append(part)
tot = sum(pp)
rs = tot.sum(dim=1)
bad = torch.where(~torch.isfinite(rs) | (rs <= 0))[0].cpu().numpy()
print("bad rows", len(bad), bad[:10])
for i in bad[:4]:
    print(i, "gmax", float(gmax[i]), "levid", float(levid[i]), [(float(p[0][i]), float(p[1][i]), int(p[2][i])) for p in parts],
          "partial sums", [float(q[i].sum()) for q in pp], "nan in partial", [bool(torch.isnan(q[i]).any()) for q in pp])
