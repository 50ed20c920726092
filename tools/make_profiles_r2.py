"""Rebuild the tracked round-2 profile documents under profiles/ from the scratch outputs of tools/gpu_r2_ncu.sh (tensor-core sweep
report) and tools/gpu_r2_final.sh (kNN report, launch lists, bench lines) in gpurun_out/.  Reads reports, profiles nothing."""
import csv
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.chdir(ROOT)


def run(cmd):
    return subprocess.run(cmd, shell=True, capture_output=True, text=True).stdout


def launch_table(path, key, title, note):
    rows = [r for r in csv.reader(open(path))]
    hdr = [r for r in rows if "Kernel Name" in r][0]
    ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    data = [r for r in rows if len(r) > 10 and r[0].isdigit()]
    marks = [i for i, r in enumerate(data) if key in r[ki]]
    step = data[marks[1]:marks[2]] if len(marks) > 2 else data[marks[-1]:]
    tot = sum(float(r[vi].replace(",", "")) for r in step) / 1e6
    o = "# %s\n\n%s\n\n| launch | grid | ms | share |\n|---|---|---:|---:|\n" % (title, note)
    for r in step:
        v = float(r[vi].replace(",", "")) / 1e6
        name = r[ki].replace("<unnamed>::", "").replace("void ", "").split("(")[0][:80]
        if v / tot >= 2e-4:
            o += "| `%s` | %s | %.3f | %.1f%% |\n" % (name, r[gi], v, 100 * v / tot)
    return o + "| total | | %.3f | |\n" % tot


if os.path.exists("gpurun_out/r2f_launches.csv"):
    shutil.copy("gpurun_out/r2f_launches.csv", "profiles/r2_launches.csv")
    md = launch_table("gpurun_out/r2f_launches.csv", "k_prep_objects",
                      "Launch list of one step of the C3 workload at the end of round 2 (262,144 objects x 199,950 models, float64 grid)",
                      "`ncu --metrics gpu__time_duration.sum --clock-control none -c 170 --csv python bench.py --objects 262144 --steps 2 "
                      "--warmup 1 --no-e2e --no-cpu --no-legs --grid float64` (tools/gpu_r2_final.sh; raw list: profiles/r2_launches.csv).  One "
                      "step = the launches between two `k_prep_objects`; durations are cold-cache and serialised, so the SHARES are what "
                      "compares with the live step (bench.py: pass1_scan 82 %, pass2_accumulate 15 %, finish 3 %).  `k_sweep_tc<5,1,0,1,1,0,1>` is "
                      "the fused single pass (faint objects), `<...,1,2>` the seeded pass 1 (bright objects), `<...,0>` with one column of "
                      "CTAs the coarse pre-passes, `<5,1,0,2,...>` the pruned pass 2.")
    open("profiles/r2_launches.md", "w").write(md)
if os.path.exists("gpurun_out/r2f_knn_launches.csv"):
    md = launch_table("gpurun_out/r2f_knn_launches.csv", "k_knn_rerank",
                      "Launch list of one kNN search (C4-shaped: 761k rows x 20 trees, k = 25, 65,536 queries)",
                      "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_knn -c 40 --csv python tools/bench_knn.py 1000000 "
                      "65536 20 25` (the tool keeps the rows with S/N > 5: 761k of 1M).  From one `k_knn_rerank` to the next: staged filter / "
                      "select of tree 0 (grid y = 1), the single sweep of the other 19 trees with the borrowed threshold, select, re-rank.")
    open("profiles/r2_knn_launches.md", "w").write(md)
if os.path.exists("gpurun_out/prof_knn_r2b.ncu-rep"):
    old = run("python tools/ncu_summary.py gpurun_out/prof_knn_r2.ncu-rep")
    new = run("python tools/ncu_summary.py gpurun_out/prof_knn_r2b.ncu-rep")
    open("profiles/r2_knn_scan_ncu.md", "w").write("""# ncu `--set full`: the kNN candidate scan before and after the round-2 redesign

Before (`k_knn_scan`, per-thread top-(k+8) lists in local memory; `ncu ... -k regex:k_knn_scan -c 1 python tools/bench_knn.py 1000000
8192 4 25`, tools/gpu_r2_profile.sh): 12.4 active threads per instruction, 1.1e9 local-memory loads, long-scoreboard stall 4.6 warps per
issue, 6 GB of DRAM writes - 85 % of the issued warp instructions were the divergent insert path.

""" + old + """
After (`k_knn_filter<5, DOT, 8>`: no list, rows below a threshold are appended, `k_knn_select` picks the k+8 smallest; the launch is
the sweep over all 761k rows of 4 trees for 16,384 queries; `ncu ... -k regex:k_knn_filter -s 3 -c 1 python tools/bench_knn.py 1000000
16384 4 25`, tools/gpu_r2_ncu.sh - taken before the threshold of tree 0 was shared with the other trees, which changed the launch
plan, not the kernel's loop): 6.55e10 distances in 14.4 ms = 4.5e12 distances/s in this launch, 3.45 thread instructions per distance,
28.4 active threads per instruction, no local memory.  Launch list of a whole search at HEAD: profiles/r2_knn_launches.md.

""" + new)
for src, dst in (("r2f_bench.json", "r2_bench_1gpu.json"), ("r2f_ref.json", "r2_bench_reference_arm.json")):
    if os.path.exists("gpurun_out/" + src):
        shutil.copy("gpurun_out/" + src, "profiles/" + dst)
print(open("profiles/r2_launches.md").read()[-1800:])
print(open("profiles/r2_knn_launches.md").read()[-1500:])
