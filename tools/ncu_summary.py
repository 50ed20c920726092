"""Markdown summary of an `ncu --set full` report: one table of the metrics the roofline argument uses per profiled launch.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md   (needs ncu on PATH; reads, does not profile)"""
import csv
import io
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "kernel duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory / block"),
    ("launch__waves_per_multiprocessor", "waves per SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe busy % (of active cycles)"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor-memory pipe active % (tcgen05.ld / MMA accumulators)"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
    ("sm__cycles_active.avg", "SM active cycles (avg)"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("smsp__inst_executed_op_tma_ld.sum", "TMA bulk loads (UBLKCP)"),
    ("smsp__sass_inst_executed_op_local_ld.sum", "local-memory loads"),
    ("smsp__sass_inst_executed_op_local_st.sum", "local-memory stores"),
]
STALLS = ["wait", "not_selected", "math_pipe_throttle", "selected", "long_scoreboard", "short_scoreboard", "mio_throttle",
          "barrier", "no_instruction", "sleeping", "branch_resolving", "lg_throttle"]

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
name_col = col.get("Kernel Name", 4)
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("## `%s`\n" % r[name_col].replace("<unnamed>::", "")[:160])
    print("| metric | value |\n|---|---|")
    for key, label in METRICS:
        if key in col:
            print("| %s (`%s`) | %s %s |" % (label, key, r[col[key]], units[col[key]]))
    for st in STALLS:
        key = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % st
        if key in col:
            try:
                v = float(r[col[key]].replace(",", ""))
            except ValueError:
                continue
            if v >= 0.2:
                print("| stall: %s (warps / issue) (`%s`) | %.3f |" % (st.replace("_", " "), key, v))
    print()
