"""Model-sharded fit_predict over NCCL (one process per GPU): every rank holds Nm/world models and all objects.
torchrun --nproc-per-node N tools/run_model_sharded.py [Nobjects]
Rank 0 also runs the unsharded problem on its own GPU and reports the difference."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench_data  # noqa: E402
import frankenz_b200 as fz  # noqa: E402
from frankenz_b200.distributed import fit_predict_model_sharded  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
os.environ["FZB_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
models, labels, depth = bench_data.c3_models()
x, xe, xm, _, _ = bench_data.c3_objects(n, models, depth, seed=99)
zgrid, sig = bench_data.c3_kde()
rdict = fz.pdf.PDFDict(zgrid, sig)
labe = np.full(len(models), 0.05)
kw = dict(free_scale=True, ignore_model_err=True, dim_prior=True)
me, mm = np.zeros_like(models), np.ones_like(models)
from frankenz_b200.distributed import ModelShardedBruteForce  # noqa: E402
sb = ModelShardedBruteForce(models, me, mm, device=local)
dev = torch.device("cuda", local)
tx, txe, txm = (torch.from_numpy(a).to(dev) for a in (x, xe, xm))
for rep in range(4):
    dist.barrier()
    torch.cuda.synchronize()
    t = time.time()
    p, (lm, le) = sb.fit_predict(tx, txe, txm, labels, labe, label_dict=rdict, lprob_kwargs=kw, as_torch=True)
    torch.cuda.synchronize()
    dt = time.time() - t
    if rank == 0:
        print("rep %d: %d objects x %d models over %d GPUs (models sharded, shard and objects resident): %.3f s -> %.3e pairs/s"
              % (rep, n, len(models), world, dt, n * len(models) / dt))
p, lm, le = p.cpu().numpy(), lm.cpu().numpy(), le.cpu().numpy()
if rank == 0:
    bf = fz.BruteForce(models, me, mm)
    p1, (lm1, le1) = bf.fit_predict(x.copy(), xe.copy(), xm.copy(), labels, labe, label_dict=rdict, return_gof=True,
                                    verbose=False, save_fits=False, lprob_kwargs=kw)
    bad = np.where(~np.isfinite(p).all(axis=1) | ~np.isfinite(p1).all(axis=1))[0]
    print("rows with non-finite PDFs: sharded %d, unsharded %d; first: %s" % ((~np.isfinite(p).all(axis=1)).sum(),
          (~np.isfinite(p1).all(axis=1)).sum(), [(int(i), float(lm[i]), float(lm1[i]), float(le[i]), float(le1[i]),
          float(np.nansum(p[i])), float(np.nansum(p1[i])), x[i].tolist(), xe[i].tolist()) for i in bad[:3]]))
    print("vs unsharded: PDF L1 max %.3g  |dlmap| max %.3g  |dlevid| max %.3g"
          % (np.nanmax(np.sum(np.abs(p - p1), axis=1)), np.nanmax(np.abs(lm - lm1)), np.nanmax(np.abs(le - le1))))
dist.destroy_process_group()
