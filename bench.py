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
# DRAM traffic of one k_sweep2 launch from the `ncu --set full` capture in profiles/r1_sweep_ncu.md
# (dram__bytes_read.sum + dram__bytes_write.sum = 85.9 MB for 196,608 objects): bytes per object of the launch
NCU_DRAM_BYTES_PER_OBJECT = 85.9e6 / 196608
# the same for one k_sweep_tc pass-1 launch (profiles/r1_sweep_tc_ncu.md): 50.7 + 15.0 MB for 196,608 objects
NCU_TC_DRAM_BYTES_PER_OBJECT = 65.7e6 / 196608
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
    from frankenz_b200._engine import make_config

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    os.environ["FZB_DEVICE"] = str(local)
    use_dist = world > 1
    if use_dist:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    models, labels, x, xe, xm = workload(args.objects, 20260103 + rank)
    no, nm = len(x), len(models)
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
    for _ in range(args.steps):
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
    value = pairs_step * args.steps / (t_all * 1e-3)

    # parity spot-check of the timed configuration's outputs (cheap invariants; full parity is in tests/)
    psum = d_pdf[:4096].sum(dim=1)
    assert bool(torch.all(torch.abs(psum[torch.isfinite(psum)] - 1.0) < 1e-9)), "PDFs are not normalised"

    # ---- end to end through the public API (numpy in / numpy out) ------------------------------------
    e2e = None
    if not args.no_e2e:
        ne = args.e2e_objects or no
        xs, xes, xms = x[:ne], xe[:ne], xm[:ne]
        times = []
        for i in range(1 + max(1, min(args.steps, 3))):
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
        e2e = {"value": float(ne) * nm * max(1, world) / te, "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "objects_per_gpu": int(ne), "seconds_per_step": te,
               "objects_per_s": float(ne) * max(1, world) / te}

    if rank != 0:
        if use_dist:
            dist.destroy_process_group()
        return

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
    tc = kind >= 2
    mufu_per_pair = 2.0 if kind == 3 else MUFU_PER_PAIR
    roofline = {"bound": "fp32_fma", "kernel": "%s<pass %d>" % (kname, 1 if ms_scan >= ms_acc else 2),
                "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s", "frac": achieved / fp32_peak,
                "peak_source": "fzb_measure_peaks (dependency-free FFMA loop, this run; MEASURED_PEAKS.json has no "
                               "fp32 entry)",
                "note": ("algorithmic 54 flop/pair (SURVEY 8d) against the FP32 FMA peak; the tensor-core sweep executes "
                         "30 of them (the three K=Nf dot products, as tf32x3 tcgen05 MMAs: 12.3 kflop of tensor work "
                         "per pair incl. the split and padding) on the tensor pipe and ~40 on the FMA pipe, so the "
                         "fraction is a figure of merit of the whole SM, not an FMA-pipe utilisation")
                if tc else "algorithmic 54 flop/pair (SURVEY 8d) against the FP32 FMA peak",
                "mufu": {"per_pair": mufu_per_pair,
                         "achieved_gops": mufu_per_pair * dom_pairs / (dom_ms * 1e-3) / 1e9, "peak_gops": mufu_peak},
                "traffic": (NCU_TC_DRAM_BYTES_PER_OBJECT if tc else NCU_DRAM_BYTES_PER_OBJECT) * float(no),
                "traffic_note": "bytes per launch scaled from the ncu --set full capture in profiles/ (%s: photometry "
                                "planes in, per-split partials out, pass 2 adds the histogram atomics); the kernel is "
                                "compute-bound, HBM carries ~0.002 B per pair"
                                % ("r1_sweep_tc_ncu.md" if tc else "r1_sweep_ncu.md"),
                "algorithmic_flops_per_pair": FLOPS_PER_PAIR,
                "pairs_per_s_kernel": dom_pairs / (dom_ms * 1e-3),
                "fit_only_pairs_per_s": float(no) * nm / (ms_scan * 1e-3),
                "ms": {"pass1_scan": ms_scan, "pass2_accumulate": ms_acc, "finish": ms_fin,
                       "step_total": float(np.mean(ms_steps))},
                "objects_routed_to_fp64": n64}
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_file):
        try:
            roofline["hbm_peak_gbs_measured"] = json.load(open(peaks_file)).get("hbm_gbs")
        except Exception:
            pass

    cpu = None
    if not args.no_cpu and world == 1:
        use_ref = reference_available()
        v, nobj, dt, cores = cpu_baseline(models, labels, x, xe, xm, args.cpu_objects_per_core, use_ref=use_ref)
        cpu = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "reference" if use_ref else "port",
               "sample": "%d objects x %d models of the same workload, %.1f s on %d processes (%s, numpy float64)"
                         % (nobj, nm, dt, cores, "unmodified frankenz BruteForce.fit_predict, oracle/_ref" if use_ref
                            else "oracle/fz_oracle.py")}

    out = {"metric": "object-model likelihood pairs/sec", "value": value, "unit": "pairs/s", "n_gpus": max(1, world),
           "steps": args.steps, "warmup": warm, "ms_per_step": t_all / args.steps, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32 sweep (tf32x3 tensor-core dot products) + f64 best-fit/PDF (f64 fallback per object)",
           "data": "synthetic", "config": cfg_json, "objects_per_s": value / nm, "clocks": clocks, "e2e": e2e,
           "gpu_launches": int(sum(s["kernel_launches"] for s in st_acc)), "roofline": roofline, "cpu_baseline": cpu}
    print(json.dumps(out))
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
