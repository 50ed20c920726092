"""Roofline denominators of this pool's B200 that MEASURED_PEAKS.json does not carry: FP32 FMA and MUFU throughput from
the library's dependency-free loops (fzb_measure_peaks), with the SM clock sampled while they run.
Usage: python tools/measure_peaks.py > profiles/peaks_r2.json"""
import json
import os
import subprocess
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from frankenz_b200._engine import Engine  # noqa: E402

eng = Engine(np.ones((8, 5)), np.ones((8, 5)), np.ones((8, 5)))
clk = []
stop = False


def sample():
    while not stop:
        try:
            out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True).stdout
            clk.append([float(v) for v in out.strip().split(",")])
        except Exception:
            pass
        time.sleep(0.05)


th = threading.Thread(target=sample, daemon=True)
th.start()
runs = [eng.measure_peaks(20) for _ in range(5)]
stop = True
th.join()
name = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=name", "--format=csv,noheader"], capture_output=True,
                      text=True).stdout.strip()
sm = float(np.median([c[0] for c in clk])) if clk else None
print(json.dumps({"gpu": name, "fp32_fma_tflops": max(r[0] for r in runs), "mufu_gops": max(r[1] for r in runs),
                  "runs": runs, "sm_mhz_median_under_load": sm, "sm_max_mhz": max(c[1] for c in clk) if clk else None,
                  "theoretical_fp32_tflops_at_median_clock": None if sm is None else 148 * 128 * 2 * sm * 1e6 / 1e12,
                  "theoretical_mufu_gops_at_median_clock": None if sm is None else 148 * 16 * sm * 1e6 / 1e9,
                  "method": "fzb_measure_peaks: 8 independent FFMA / MUFU.EX2 chains per thread, 2048 resident threads "
                            "per SM, best of 20 launches, CUDA events"}, indent=1))
