#!/usr/bin/env python
"""Benchmark of the brute-force photometric likelihood path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--objects No]

Workload (SURVEY.md section 8d, config C3 = BASELINE.json configs[2]): template fitting with a free
scale, 1M synthetic HSC grizy objects x 199,950 template x redshift models, dictionary-KDE redshift
PDFs on a 701-point grid; BruteForce.fit_predict(save_fits=False).  One step = one pass of the hot
path over the whole object batch.  With N GPUs every rank processes its own batch of that size
(objects sharded, models replicated, no data-path collective): weak scaling.

`value`  = object-model pairs/s, inputs resident in HBM (fzb_fit_predict_dev), CUDA-event time, max
           over ranks.
`e2e`    = the same metric through the reference-shaped Python API (numpy in / numpy out, H2D and
           D2H copies inside the timed region).
`--impl reference` times the CPU restatement of the reference (oracle/fz_oracle.py, a port: the
reference is pure Python and cannot travel to the GPU box) on all host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_data  # noqa: E402

FLOPS_PER_PAIR = 54.0   # FS0 at Nf=5, 9*Nf+9 (FMA = 2 flops), + 3 MUFU per pair
MUFU_PER_PAIR = 3.0
LPROB = dict(free_scale=True, ignore_model_err=True, dim_prior=True)


def workload(n_obj, seed):
    models, labels, depth = bench_data.c3_models()
    x, xe, xm, _, _ = bench_data.c3_objects(n_obj, models, depth, seed=seed)
    return models, labels, x, xe, xm


# ---- clocks ------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.rows, self.proc = gpu, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU arms ------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def reference_available():
    """The unmodified reference, installed by oracle/build_ref.sh (travels to the GPU box with the snapshot)."""
    return os.path.isdir(os.path.join(REF_DIR, "frankenz"))


def _cpu_chunk(args):
    import warnings
    warnings.filterwarnings("ignore")
    models, labels, x, xe, xm, use_ref = args
    zgrid, sig = bench_data.c3_kde()
    t = time.time()
    if use_ref:
        # the reference's own public API and stock code path: frankenz.fitting.BruteForce.fit_predict
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        from frankenz.fitting import BruteForce
        from frankenz.pdf import PDFDict
        with np.errstate(all="ignore"):
            bf = BruteForce(models, np.zeros_like(models), np.ones_like(models))
            bf.fit_predict(x, xe, xm, labels, np.full(len(models), 0.05), label_dict=PDFDict(zgrid, sig),
                           lprob_kwargs=dict(LPROB), return_gof=True, verbose=False, save_fits=False)
        return time.time() - t
    from oracle import fz_oracle as fo
    kd = fo.KernelDict(zgrid, sig)
    with np.errstate(all="ignore"):
        fo.bruteforce_fit_predict(models, np.zeros_like(models), np.ones_like(models), x, xe, xm, labels,
                                  np.full(len(models), 0.05), label_dict=kd, **LPROB)
    return time.time() - t


def cpu_baseline(models, labels, x, xe, xm, per_core, cores=None, use_ref=False):
    """The reference (oracle/_ref, `use_ref`) or the oracle port on `cores` processes, `per_core` objects each (a
    bounded sample of the workload; the reference is single-threaded, so one process per core)."""
    import multiprocessing as mp
    if use_ref:      # import once in the parent: the forked workers inherit the loaded package
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        import frankenz.fitting  # noqa: F401
    cores = cores or os.cpu_count() or 1
    n = min(len(x), per_core * cores)
    per = max(1, n // cores)
    chunks = [(models, labels, x[i * per:(i + 1) * per].copy(), xe[i * per:(i + 1) * per].copy(),
               xm[i * per:(i + 1) * per].copy(), use_ref) for i in range(cores)]
    nobj = sum(len(c[2]) for c in chunks)
    ctx = mp.get_context("fork")
    t = time.time()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_chunk, chunks)
    dt = time.time() - t
    return nobj * len(models) / dt, nobj, dt, cores


# ---- distributed plumbing --------------------------------------------------------------------------
def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---- legs beyond the headline: C5-shaped model-sharded run and C4-shaped kNN ---------------
def leg_model_sharded(args, rank, world, local, dist, torch, fz):
    """C5-shaped: 6-band LSST training rows with errors (float64), default likelihood (fixed scale, model errors,
    dim_prior), models sharded over the ranks, objects replicated; merged with ONE all-gather (24 B/object/rank) and ONE
    reduce-scatter of fp32 PDF partials per chunk (frankenz_b200/distributed.py).  Rank 0 checks the result against the
    unsharded fit_predict of the same problem on its own GPU in the same run."""
    from frankenz_b200.distributed import ModelShardedBruteForce
    n_train, n_obj = args.c5_models, args.c5_objects
    tr, tre, trm, ztr, x, xe, xm = bench_data.c5_dataset(n_train, n_obj)
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(n_train, 0.05)
    dev = torch.device("cuda", local)
    sb = ModelShardedBruteForce(tr, tre, trm, device=local, chunk=args.c5_chunk)
    tx, txe, txm = (torch.from_numpy(a).to(dev) for a in (x, xe, xm))
    times, last = [], None
    for rep in range(1 + max(1, min(args.steps, 3))):
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        p, (lm, le), best = sb.fit_predict(tx, txe, txm, ztr, labe, label_dict=rdict, as_torch=True, gather=False,
                                           return_best=True)
        if rep > 0:
            times.append(sb.last["ms_total"])
            last = dict(sb.last)
    tt = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
    # every rank: the same objects against ALL models on one GPU (a sub-batch) = what an independent GPU delivers
    nsub = min(n_obj, args.c5_check_objects)
    bf = fz.BruteForce(tr, tre, trm)
    ref_times = []
    for rep in range(2):
        p1, (lm1, le1) = bf.fit_predict(x[:nsub].copy(), xe[:nsub].copy(), xm[:nsub].copy(), ztr, labe, label_dict=rdict,
                                        return_gof=True, verbose=False, save_fits=False)
        ref_times.append(bf._eng().stats()["ms_total"])
    single = float(nsub) * n_train / (min(ref_times) * 1e-3)
    ts = torch.tensor([single], dtype=torch.float64, device=dev)
    dist.all_reduce(ts, op=dist.ReduceOp.SUM)
    # parity: the rows this rank owns among the first nsub objects, against its own unsharded run
    own = sb.owned_indices(n_obj)
    sel = own < nsub
    pd = p[torch.from_numpy(sel).to(dev)].cpu().numpy()
    l1 = float(np.max(np.sum(np.abs(pd - p1[own[sel]]), axis=1))) if sel.any() else 0.0
    lmh, leh, bh = lm[:nsub].cpu().numpy(), le[:nsub].cpu().numpy(), best[:nsub].cpu().numpy()
    dl = float(np.max(np.abs(lmh - lm1) / np.maximum(1.0, np.abs(lm1))))
    de = float(np.max(np.abs(leh - le1) / np.maximum(1.0, np.abs(le1))))
    stat = torch.tensor([l1, dl, de, float(np.mean(bh != bf.best_idx))], dtype=torch.float64, device=dev)
    dist.all_reduce(stat, op=dist.ReduceOp.MAX)
    sb.close()
    value = float(n_obj) * n_train / (ms * 1e-3)
    return {"workload": "C5-shaped: %d objects x %d training rows, 6 bands (LSST ugrizY), default likelihood with model "
                        "errors, float64 rows, models sharded x%d, dictionary KDE" % (n_obj, n_train, world),
            "value": value, "unit": "pairs/s", "ms_per_step": ms, "chunk_objects": sb.chunk,
            "collectives_per_step": last["collectives"], "nccl_bytes_per_rank_per_step": last["nccl_bytes"],
            "nccl_ms_rank0": last["ms_nccl"], "nccl_share_of_step_rank0": last["ms_nccl"] / last["ms_total"],
            "pass1_ms_rank0": last["ms_pass1"], "pass2_ms_rank0": last["ms_pass2"],
            "exposed_ms_rank0": last["ms_total"] - last["ms_pass1"] - last["ms_pass2"],
            "independent_gpus_pairs_per_s": float(ts.item()), "efficiency_vs_independent_gpus": value / float(ts.item()),
            "parity_vs_unsharded": {"objects_checked": int(nsub), "pdf_l1_max": float(stat[0].item()),
                                    "lmap_rel_max": float(stat[1].item()), "levid_rel_max": float(stat[2].item()),
                                    "best_index_mismatch_fraction": float(stat[3].item()),
                                    "best_index_note": "not a reference output; with dim_prior the log-posterior is flat at chi2 = "
                                                       "Nband-2, so among 1M rows several lie within fp32 rounding of the maximum; the "
                                                       "picks differ by <= 2e-7 in float64 log-posterior (tools/diag_shard_best.py)",
                                    "bound": 1e-5,
                                    "ok": bool(stat[0].item() <= 1e-5 and stat[1].item() <= 1e-5 and stat[2].item() <= 1e-5)}}


def leg_default_likelihood(args, rank, world, local, dist, torch, fz):
    """The reference's DEFAULT likelihood (fixed scale, model errors, dim_prior: pdf.py:27-100), the mode of configs C1,
    C2 and C5: 6-band LSST training rows with errors (float64), device-resident inputs, one GPU's share.  Per pair
    6 Nf + 8 = 44 flop and Nf + 2 = 8 MUFU results (one reciprocal per band): MUFU-bound."""
    from frankenz_b200._engine import make_config
    n_train, n_obj = args.fx1_models, args.fx1_objects
    tr, tre, trm, ztr, x, xe, xm = bench_data.c5_dataset(n_train, n_obj, seed=20260107 + rank)
    zgrid, sig = bench_data.c3_kde()
    bf = fz.BruteForce(tr, tre, trm)
    eng = bf._eng()
    eng.set_kde(ztr, np.full(n_train, 0.05), label_dict=fz.pdf.PDFDict(zgrid, sig))
    cfg = make_config(dict(), None)
    dev = torch.device("cuda", local)
    d = [torch.from_numpy(a).to(dev) for a in (x, xe, xm)]
    out = [torch.empty((n_obj, eng.Ng), dtype=torch.float64, device=dev)] + \
          [torch.empty(n_obj, dtype=torch.float64, device=dev) for _ in range(2)] + \
          [torch.empty(n_obj, dtype=torch.int64, device=dev)] + \
          [torch.empty(n_obj, dtype=torch.float64, device=dev) for _ in range(2)]
    sts = []
    for rep in range(4):
        torch.cuda.synchronize()
        eng.fit_predict_dev(d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n_obj, cfg, *[t.data_ptr() for t in out])
        if rep > 0:
            sts.append(eng.stats())
    ms = float(np.mean([s["ms_total"] for s in sts]))
    ms_scan = float(np.mean([s["ms_scan"] for s in sts]))
    n64 = float(np.mean([s["objects_fp64"] for s in sts]))
    fp32_peak, mufu_peak = eng.measure_peaks(3)
    vals = torch.tensor([ms, ms_scan], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms, ms_scan = float(vals[0].item()), float(vals[1].item())
    pairs = float(n_obj) * n_train
    eng.close()
    return {"workload": "default likelihood (free_scale=False, model errors, dim_prior=True), %d objects x %d training "
                        "rows per GPU, 6 bands (LSST ugrizY), float64 rows, fit_predict(save_fits=False) on device-resident "
                        "inputs" % (n_obj, n_train),
            "value": pairs * max(1, world) / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms,
            "objects_routed_to_fp64": n64,
            "objects_completed_by_the_fused_pass_frac": float(np.mean([s.get("objects_fused", 0) for s in sts])) / float(n_obj),
            "ms": {"sweep_phase": ms_scan, "pass2_and_redecisions": float(np.mean([s["ms_accum"] for s in sts])),
                   "finish": float(np.mean([s["ms_finish"] for s in sts]))},
            "roofline": {"bound": "mufu", "kernel": "k_sweep2<FX1, fused single pass> incl. the 1/16 pre-pass (the plain pass 1 of "
                                                     "round 1 ran at 0.92 of the MUFU peak; the fused pass also fills the KDE "
                                                     "histogram and is issue bound)", "mufu_per_pair": 8.0,
                         "achieved_gops": 8.0 * pairs / (ms_scan * 1e-3) / 1e9, "peak_gops": mufu_peak,
                         "frac": 8.0 * pairs / (ms_scan * 1e-3) / 1e9 / mufu_peak, "pass1_ms": ms_scan,
                         "flops_per_pair": 44.0, "fp32_tflops": 44.0 * pairs / (ms_scan * 1e-3) / 1e12,
                         "fp32_peak_tflops": fp32_peak,
                         "peak_source": "fzb_measure_peaks (dependency-free MUFU.EX2 loop, this run)"}}


def leg_knn(args, rank, world, local, dist, torch, fz):
    """C4-shaped: NearestNeighbors over K Monte-Carlo realisations of the training set, queries sharded over the ranks
    (training features replicated), k neighbours per tree; the search is checked against the all-float64 kernel
    (FZB_KNN_EXACT_ONLY) on a sub-sample in the same run."""
    ntr, nq, K, k = args.knn_train, args.knn_queries, args.knn_K, args.knn_k
    (tr, tre, trm, ztr), (qx, qe, qm), fmap = bench_data.c4_dataset(ntr, nq * max(1, world))
    qx, qe, qm = (a[rank * nq:(rank + 1) * nq] for a in (qx, qe, qm))
    t0 = time.perf_counter()
    nn = fz.NearestNeighbors(tr, tre, trm, K=K, fmap_kwargs=fmap, rstate=np.random.RandomState(1), verbose=False)
    t_build = time.perf_counter() - t0
    eng = nn._engine
    q = nn._query_features(qx, qe, np.random.RandomState(2 + rank))
    dev = torch.device("cuda", local)
    ms, redo = [], 0
    for rep in range(3):
        idx, _ = eng.knn_query(q, k, p=2, return_dist=False)
        st = eng.stats()
        if rep > 0:
            ms.append(st["ms_scan"])          # CUDA-event time of the search (the call's total adds the index download)
            redo = int(st["knn_redo"])
    ncheck = min(nq, args.knn_check)
    os.environ["FZB_KNN_EXACT_ONLY"] = "1"
    try:
        idx_exact, _ = eng.knn_query(q[:ncheck], k, p=2, return_dist=False)
    finally:
        del os.environ["FZB_KNN_EXACT_ONLY"]
    mism = int(np.count_nonzero(idx[:ncheck] != idx_exact))
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(ntr, 0.05)
    te = []
    for rep in range(2):
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
        t0 = time.perf_counter()
        p = nn.fit_predict(qx.copy(), qe.copy(), qm.copy(), ztr, labe, label_dict=rdict, k=k, eps=0,
                           rstate=np.random.RandomState(2 + rank), verbose=False, save_fits=False)
        te.append(time.perf_counter() - t0)
    assert np.max(np.abs(p.sum(axis=1) - 1.0)) < 1e-9
    vals = torch.tensor([float(np.mean(ms)), te[-1], float(mism), float(redo)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_scan, t_e2e = float(vals[0].item()), float(vals[1].item())
    evals = float(nq) * K * ntr * max(1, world)
    return {"workload": "C4-shaped: %d-row training set x K=%d Monte-Carlo realisations (luptitude features), k=%d, "
                        "%d queries per GPU, queries sharded x%d" % (ntr, K, k, nq, max(1, world)),
            "distance_evaluations_per_s": evals / (ms_scan * 1e-3), "search_ms": ms_scan,
            "queries_per_s_search": float(nq) * max(1, world) / (ms_scan * 1e-3),
            "e2e_queries_per_s": float(nq) * max(1, world) / t_e2e, "e2e_seconds": t_e2e,
            "build_seconds_host_mc_featuremap_h2d": t_build, "redo_searches": int(vals[3].item()),
            "redo_fraction": float(vals[3].item()) / (float(nq) * K),
            "index_check": {"queries": int(ncheck), "against": "all-float64 kernel (FZB_KNN_EXACT_ONLY)",
                            "mismatching_indices": int(vals[2].item()), "ok": bool(vals[2].item() == 0)},
            "roofline": {"bound": "fp32_fma", "flops_per_distance": 15.0,
                         "achieved_tflops": 15.0 * evals / (ms_scan * 1e-3) / 1e12 / max(1, world)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--objects", type=int, default=int(os.environ.get("FZB_BENCH_OBJECTS", 1000000)))
    ap.add_argument("--e2e-objects", type=int, default=0, help="objects per e2e step (default: same as --objects)")
    ap.add_argument("--cpu-objects-per-core", type=int, default=24)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--lprob", default="", help="JSON overriding the likelihood flags (experiments only)")
    ap.add_argument("--grid", default="both", choices=["fp32", "float64", "both"],
                    help="model grid of the headline: float32-rounded fluxes, the reference's float64 grid, or both")
    ap.add_argument("--no-legs", action="store_true", help="skip the model-sharded (N > 1) and kNN legs")
    ap.add_argument("--c5-models", type=int, default=1048576)
    ap.add_argument("--c5-objects", type=int, default=262144)
    ap.add_argument("--c5-chunk", type=int, default=65536)
    ap.add_argument("--c5-check-objects", type=int, default=16384)
    ap.add_argument("--fx1-models", type=int, default=262144)
    ap.add_argument("--fx1-objects", type=int, default=262144)
    ap.add_argument("--knn-train", type=int, default=1000000)
    ap.add_argument("--knn-queries", type=int, default=65536)
    ap.add_argument("--knn-K", type=int, default=20)
    ap.add_argument("--knn-k", type=int, default=25)
    ap.add_argument("--knn-check", type=int, default=2048)
    args = ap.parse_args()
    if args.lprob:
        LPROB.clear()
        LPROB.update(json.loads(args.lprob))
    rank, world, local = dist_env()
    if world != max(1, args.gpus) and world > 1:
        args.gpus = world
    warm = max(3, args.warmup)
    cfg_json = {"workload": "C3 template fitting free_scale=True ignore_model_err=True dim_prior=True, HSC grizy, "
                            "BruteForce.fit_predict(save_fits=False) + dictionary KDE (701-point zgrid)",
                "objects_per_gpu": args.objects, "models": 199950, "filters": 5, "parallelism": "objects sharded x%d, "
                "models replicated" % max(1, args.gpus),
                "cache": "per-step working set (2.8 GB histogram + 5.6 GB PDFs + 126 MB inputs) exceeds the 126 MB L2; "
                         "an extra 512 MB buffer is overwritten between timed steps"}

    # ---------------- reference arm: oracle port on the host cores ------------------------------------
    if args.impl == "reference":
        if rank != 0:
            return
        models, labels, x, xe, xm = workload(max(256, args.cpu_objects_per_core * (os.cpu_count() or 1)), 20260103)
        vals = []
        use_ref = reference_available()
        for i in range(max(1, args.warmup and 1) + args.steps):
            v, nobj, dt, cores = cpu_baseline(models, labels, x, xe, xm, args.cpu_objects_per_core, use_ref=use_ref)
            if i >= 1:
                vals.append((v, dt))
        v = float(np.mean([a for a, _ in vals]))
        ms = float(np.mean([b for _, b in vals])) * 1e3
        sample = "%d objects x %d models per step on %d processes (%s)" % (
            nobj, len(models), cores, "unmodified frankenz.fitting.BruteForce.fit_predict from oracle/_ref" if use_ref
            else "oracle/fz_oracle.py")
        print(json.dumps({"impl": "reference", "metric": "object-model likelihood pairs/sec", "value": v,
                          "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": 1,
                          "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": cfg_json,
                          "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores,
                                           "kind": "reference" if use_ref else "port", "sample": sample},
                          "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}))
        return

    # ---------------- our arm ------------------------------------------------------------------------
    import torch
    import frankenz_b200 as fz

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    os.environ["FZB_DEVICE"] = str(local)
    use_dist = world > 1
    dist = None
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = dict(args=args, rank=rank, world=world, local=local, dist=dist, torch=torch, fz=fz, warm=warm,
               use_dist=use_dist)

    # headline: the reference's own float64 model grid (simulate.py:954-1021 returns float64); the float32-rounded
    # grid of round 1 is measured beside it (device-timed only)
    head_grid = "float64" if args.grid in ("float64", "both") else "fp32"
    head = run_c3(ctx, head_grid, steps=args.steps, e2e=not args.no_e2e, cpu=not args.no_cpu)
    other = None
    if args.grid == "both":
        other = run_c3(ctx, "fp32", steps=max(1, min(args.steps, 3)), e2e=False, cpu=False)

    legs = {}
    if not args.no_legs:
        if use_dist:
            legs["model_sharded"] = leg_model_sharded(args, rank, world, local, dist, torch, fz)
        legs["default_likelihood"] = leg_default_likelihood(args, rank, world, local, dist, torch, fz)
        legs["knn"] = leg_knn(args, rank, world, local, dist, torch, fz)

    if rank == 0:
        cfg_json["model_grid"] = head["grid_note"]
        out = {"metric": "object-model likelihood pairs/sec", "value": head["value"], "unit": "pairs/s",
               "n_gpus": max(1, world), "steps": args.steps, "warmup": warm, "ms_per_step": head["ms_per_step"],
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32 sweep (tf32x3 tensor-core dot products) + f64 best-fit/PDF (f64 fallback per object)",
               "data": "synthetic", "config": cfg_json, "objects_per_s": head["value"] / head["nm"],
               "clocks": head["clocks"], "e2e": head["e2e"], "e2e_summaries": head.get("e2e_summaries"),
               "gpu_launches": head["gpu_launches"], "roofline": head["roofline"], "cpu_baseline": head["cpu"]}
        if other is not None:
            out["fp32_rounded_grid"] = {"value": other["value"], "unit": "pairs/s", "ms_per_step": other["ms_per_step"],
                                        "grid": other["grid_note"], "roofline_frac": other["roofline"]["frac"],
                                        "ms": other["roofline"]["ms"], "kernel": other["roofline"]["kernel"]}
        out.update(legs)
        print(json.dumps(out))
    if use_dist:
        dist.destroy_process_group()


def run_c3(ctx, grid, steps, e2e, cpu):
    """The headline workload on one model grid: K device-timed steps (`value`), then the numpy-in / numpy-out call
    (`e2e`).  Returns a dict on every rank (rank 0's is printed)."""
    args, rank, world, local = ctx["args"], ctx["rank"], ctx["world"], ctx["local"]
    dist, torch, fz, warm, use_dist = ctx["dist"], ctx["torch"], ctx["fz"], ctx["warm"], ctx["use_dist"]
    from frankenz_b200._engine import make_config
    models, labels, depth = bench_data.c3_models(float64_grid=(grid == "float64"))
    x, xe, xm, _, _ = bench_data.c3_objects(args.objects, models, depth, seed=20260103 + rank)
    no, nm = len(x), len(models)
    f32_exact = bool(np.array_equal(models.astype(np.float32).astype(np.float64), models))
    grid_note = ("float64 fluxes as frankenz.simulate.make_model_grid returns them (not fp32-representable)"
                 if not f32_exact else "model fluxes fp32-representable (the float64 grid rounded to float32)")
    zgrid, sig = bench_data.c3_kde()
    rdict = fz.pdf.PDFDict(zgrid, sig)
    labe = np.full(nm, 0.05)
    bf = fz.BruteForce(models, np.zeros_like(models), np.ones_like(models))
    eng = bf._eng()
    eng.set_kde(labels, labe, label_dict=rdict)
    cfg = make_config(LPROB, None)
    dev = torch.device("cuda", local)
    d_x, d_xe, d_xm = (torch.from_numpy(a).to(dev) for a in (x, xe, xm))
    d_pdf = torch.empty((no, eng.Ng), dtype=torch.float64, device=dev)
    d_lmap = torch.empty(no, dtype=torch.float64, device=dev)
    d_levid = torch.empty(no, dtype=torch.float64, device=dev)
    d_best = torch.empty(no, dtype=torch.int64, device=dev)
    d_bchi2 = torch.empty(no, dtype=torch.float64, device=dev)
    d_bscale = torch.empty(no, dtype=torch.float64, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def step():
        eng.fit_predict_dev(d_x.data_ptr(), d_xe.data_ptr(), d_xm.data_ptr(), no, cfg, d_pdf.data_ptr(),
                            d_lmap.data_ptr(), d_levid.data_ptr(), d_best.data_ptr(), d_bchi2.data_ptr(),
                            d_bscale.data_ptr())
        return eng.stats()

    def barrier():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ms_steps, st_acc = [], []
    for _ in range(steps):
        flush.fill_(1)
        torch.cuda.synchronize()
        st = step()            # the library times its own stream with CUDA events (ms_total)
        ms_steps.append(st["ms_total"])
        st_acc.append(st)
    barrier()
    clocks = sampler.stop()
    t_rank = float(sum(ms_steps))
    if use_dist:
        tt = torch.tensor([t_rank], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_all = float(tt.item())
    else:
        t_all = t_rank
    pairs_step = float(no) * nm * max(1, world)
    value = pairs_step * steps / (t_all * 1e-3)

    # parity spot-check of the timed configuration's outputs (cheap invariants; full parity is in tests/)
    psum = d_pdf[:4096].sum(dim=1)
    assert bool(torch.all(torch.abs(psum[torch.isfinite(psum)] - 1.0) < 1e-9)), "PDFs are not normalised"
    del d_pdf, flush

    # ---- end to end through the public API (numpy in / numpy out) ------------------------------------
    res_e2e = None
    if e2e:
        ne = args.e2e_objects or no
        xs, xes, xms = x[:ne], xe[:ne], xm[:ne]
        times = []
        for i in range(1 + max(1, min(steps, 3))):
            barrier()
            t0 = time.perf_counter()
            p, (lm, le) = bf.fit_predict(xs, xes, xms, labels, labe, label_dict=rdict, return_gof=True,
                                         verbose=False, save_fits=False, lprob_kwargs=LPROB)
            t1 = time.perf_counter()
            if i > 0:
                times.append(t1 - t0)
            h2d = 3 * xs.nbytes
            d2h = p.nbytes + lm.nbytes + le.nbytes + 3 * lm.nbytes
            del p
        te = float(np.mean(times))
        if use_dist:
            tt = torch.tensor([te], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            te = float(tt.item())
        res_e2e = {"value": float(ne) * nm * max(1, world) / te, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
                   "d2h_bytes_per_step": int(d2h), "objects_per_gpu": int(ne), "seconds_per_step": te,
                   "objects_per_s": float(ne) * max(1, world) / te,
                   "call": "BruteForce.fit_predict(numpy..., save_fits=False, return_gof=True) -> (No x 701) float64 PDFs"}

    # ---- end to end with the summaries fused behind fit_predict (the PDFs stay on the device) ---------------------
    res_summ = None
    if e2e:
        times, ms_s = [], []
        for i in range(1 + max(1, min(steps, 3))):
            barrier()
            t0 = time.perf_counter()
            summ, (lm, le) = bf.fit_predict_summarize(xs, xes, xms, labels, labe, label_dict=rdict, return_gof=True,
                                                      verbose=False, lprob_kwargs=LPROB, return_pdfs=False,
                                                      rstate=np.random.RandomState(1))
            t1 = time.perf_counter()
            if i > 0:
                times.append(t1 - t0)
                ms_s.append(eng.stats()["ms_summarize"])
        ts = float(np.mean(times))
        if use_dist:
            tt = torch.tensor([ts], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ts = float(tt.item())
        d2h_s = int(sum(a.nbytes for est in summ[:4] for a in est) + sum(a.nbytes for a in summ[4]) + summ[5].nbytes
                    + lm.nbytes + le.nbytes + 3 * lm.nbytes)
        assert np.all(np.isfinite(summ[0][0][np.isfinite(lm)]))
        res_summ = {"value": float(ne) * nm * max(1, world) / ts, "unit": "pairs/s", "seconds_per_step": ts,
                    "h2d_bytes_per_step": int(3 * xs.nbytes + ne * 8 + eng.Ng * eng.Ng * 8), "d2h_bytes_per_step": d2h_s,
                    "objects_per_s": float(ne) * max(1, world) / ts, "summarize_ms_device": float(np.mean(ms_s)),
                    "call": "BruteForce.fit_predict(..., summarize=True, return_pdfs=False): fit + PDF + pdfs_summarize "
                            "on the device, point estimates / intervals / risks out"}
        ms_dev = float(np.mean(ms_s))
        if ms_dev > 0:
            # k_summarize: the risk product pdf . (1 - kernel) is 2 Ng^2 flop per object in float64 on the FP64 tensor cores
            tf = 2.0 * eng.Ng * eng.Ng * ne / (ms_dev * 1e-3) * 1e-12
            res_summ["summarize_roofline"] = {
                "kernel": "k_summarize (DMMA m8n8k4 risk product + numpy-order CDF / quantiles / estimators)", "bound": "fp64 tensor",
                "achieved": tf, "peak": 37.1, "unit": "TFLOP/s", "frac": tf / 37.1,
                "peak_source": "mma.sync.m8n8k4.f64 microbenchmark on B200 (tools/fp64_rate.cu: DMMA 37.1, DFMA 33.9 TFLOP/s)"}

    # ---- roofline of the dominant kernel (the fp32 sweep), measured live ------------------------------
    fp32_peak, mufu_peak = eng.measure_peaks(5)
    ms_scan = float(np.mean([s["ms_scan"] for s in st_acc]))
    ms_acc = float(np.mean([s["ms_accum"] for s in st_acc]))
    ms_fin = float(np.mean([s["ms_finish"] for s in st_acc]))
    n64 = float(np.mean([s["objects_fp64"] for s in st_acc]))
    dom_ms = max(ms_scan, ms_acc)
    dom_pairs = float(no) * nm if ms_scan >= ms_acc else float(no - n64) * nm
    achieved = FLOPS_PER_PAIR * dom_pairs / (dom_ms * 1e-3) / 1e12
    kind = int(st_acc[-1].get("sweep_kind", 1))
    kname = {1: "k_sweep2", 2: "k_sweep_tc", 3: "k_sweep_tc<LIN>"}.get(kind, "k_sweep2")
    n_fused = float(np.mean([s.get("objects_fused", 0) for s in st_acc]))
    if n_fused > 0:
        kname = "k_sweep_tc<LIN, fused single pass / seeded pass 1 by S/N, incl. the 1/16 pre-pass>"
    tc = kind >= 2
    mufu_per_pair = 2.0 if kind == 3 else MUFU_PER_PAIR
    traffic = ncu_traffic_per_object(tc)
    roofline = {"bound": "fp32_fma", "kernel": "%s<pass %d>" % (kname, 1 if ms_scan >= ms_acc else 2),
                "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                "peak_source": "fzb_measure_peaks (dependency-free FFMA loop, this run; MEASURED_PEAKS.json has no "
                               "fp32 entry)",
                "step_structure": ("pass1_scan = coarse pre-pass (every 16th model) + ONE sweep over all pairs that also fills the "
                                   "KDE histogram for the faint objects (S/N <= 32); pass2_accumulate = pruned second sweep of "
                                   "the bright objects + float64 re-decision of the weights recorded at the cut") if n_fused > 0
                                  else "pass 1 (max / evidence / arg-max) + pass 2 (weights above the cut -> histogram)",
                "note": ("algorithmic 54 flop/pair (SURVEY 8d) against the FP32 FMA peak; the tensor-core sweep executes "
                         "30 of them (the three K=Nf dot products, as tf32x3 tcgen05 MMAs) on the tensor pipe and ~40 "
                         "on the FMA pipe, so the fraction is a figure of merit of the whole SM, not an FMA-pipe "
                         "utilisation") if tc else "algorithmic 54 flop/pair (SURVEY 8d) against the FP32 FMA peak",
                "mufu": {"per_pair": mufu_per_pair,
                         "achieved_gops": mufu_per_pair * dom_pairs / (dom_ms * 1e-3) / 1e9, "peak_gops": mufu_peak},
                "traffic": None if traffic is None else traffic[0] * float(no),
                "traffic_note": None if traffic is None else traffic[1],
                "algorithmic_flops_per_pair": FLOPS_PER_PAIR,
                "pairs_per_s_kernel": dom_pairs / (dom_ms * 1e-3),
                "fit_only_pairs_per_s": float(no) * nm / (ms_scan * 1e-3),
                "whole_step_frac": FLOPS_PER_PAIR * float(no) * nm / (float(np.mean(ms_steps)) * 1e-3) / 1e12 / fp32_peak,
                "pass2_pairs_evaluated_frac": float(np.mean([s.get("pairs_pass2", 0) for s in st_acc])) / (float(no) * nm),
                "ms": {"pass1_scan": ms_scan, "pass2_accumulate": ms_acc, "finish": ms_fin,
                       "step_total": float(np.mean(ms_steps))},
                "objects_routed_to_fp64": n64, "objects_completed_by_the_fused_pass_frac": n_fused / float(no),
                "weights_redecided_in_float64_per_object": float(np.mean([s.get("cut_recorded", 0) for s in st_acc])) / float(no)}
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        try:
            roofline["hbm_peak_gbs_measured"] = json.load(open(peaks_file)).get("hbm_gbs")
        except Exception:
            pass

    res_cpu = None
    if cpu and world == 1 and rank == 0:
        use_ref = reference_available()
        v, nobj, dt, cores = cpu_baseline(models, labels, x, xe, xm, args.cpu_objects_per_core, use_ref=use_ref)
        res_cpu = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "reference" if use_ref else "port",
                   "sample": "%d objects x %d models of the same workload, %.1f s on %d processes (%s, numpy float64)"
                             % (nobj, nm, dt, cores, "unmodified frankenz BruteForce.fit_predict, oracle/_ref" if use_ref
                                else "oracle/fz_oracle.py")}
    launches = int(sum(s["kernel_launches"] for s in st_acc))
    eng.close()
    bf._engine = None
    del d_x, d_xe, d_xm
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": t_all / steps, "nm": nm, "clocks": clocks, "e2e": res_e2e,
            "e2e_summaries": res_summ,
            "gpu_launches": launches, "roofline": roofline, "cpu": res_cpu, "grid_note": grid_note}


def ncu_traffic_per_object(tc):
    """DRAM bytes per object of one pass-1 launch of the dominant kernel, parsed from the committed `ncu --set full`
    summary (dram__bytes_read.sum + dram__bytes_write.sum over the 196,608 objects of the profiled launch)."""
    import re
    names = ["r2_sweep_tc_ncu.md", "r1_sweep_tc_ncu.md"] if tc else ["r1_sweep_ncu.md"]
    for name in names:
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        txt = open(path).read()
        mark = txt.find("Objects in this launch:")
        sec = txt[mark:] if mark >= 0 else (txt.split("## ")[1] if "## " in txt else txt)      # the dominant sweep launch
        rd = re.search(r"dram__bytes_read\.sum`\) \| ([0-9.]+) (\w+)", sec)
        wr = re.search(r"dram__bytes_write\.sum`\) \| ([0-9.]+) (\w+)", sec)
        ob = re.search(r"Objects in this launch: (\d+)", txt) or re.search(r"--objects (\d+)", txt)
        if not (rd and wr and ob):
            continue
        unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        tot = float(rd.group(1)) * unit.get(rd.group(2), 1.0) + float(wr.group(1)) * unit.get(wr.group(2), 1.0)
        return tot / float(ob.group(1)), ("dram__bytes_read.sum + dram__bytes_write.sum of the dominant sweep launch in "
                                          "profiles/%s, per object x objects of this run (photometry planes in; "
                                          "partials, live bits, records and histogram REDs out; the kernel is "
                                          "compute-bound)" % name)
    return None


if __name__ == "__main__":
    main()
