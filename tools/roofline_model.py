#!/usr/bin/env python
"""First-principles bounds of the C3 kernels next to their measured times (no GPU needed).

For every large kernel of a Block forward + backward: executed MACs, the time the FP32 FMA pipe needs
for them at 100 % (148 SMs x 128 lanes at the measured SM clock), the algorithmic HBM bytes and the time the measured copy
bandwidth needs for them, the measured CUDA-event time (profiles/r02_bench_c3_n1.json or the file given) and the two
fractions.  `python tools/roofline_model.py [bench.json]`
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# kernel: (MAC per edge in F^2, MAC per fibre in F^2, rows per edge, rows per fibre)   -- rows of F floats
KERNELS = {
    "k_edge_fwd": (8, 0, 2, 4),
    "k_source_edge_fwd": (6, 0, 2, 10),
    "k_source_node_fwd_mma": (0, 100, 0, 22),
    "k_target_edge_fwd": (2, 0, 1, 2),
    "k_target_edge_bwd": (6, 0, 2, 4),
    "k_source_node_bwd_mma": (0, 200, 0, 32),
    "k_source_edge_bwd": (18, 0, 2, 18),
    "k_edge_bwd2": (20, 0, 3, 8),
}


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_bench_c3_n1.json")
    line = json.loads(open(path).read().strip().splitlines()[-1])
    cfg = line["config"]
    G, S, T, F = cfg["graphs_per_gpu"], cfg["fibres"], cfg["classes"], cfg["fdim"]
    E = S * T
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        hbm = float(peaks["hbm_gbs"])
    except Exception:
        hbm = 6650.0
    clock = (line.get("clocks") or {}).get("sm_mhz", 1965) * 1e6
    fma_per_s = 148 * 128 * clock
    print("C3 at N = %d: %d graphs of %d x %d per GPU, Fdim %d; FMA pipe %.1f TMAC/s at %.0f MHz, HBM %.0f GB/s (measured copy)"
          % (line["n_gpus"], G, S, T, F, fma_per_s / 1e12, clock / 1e6, hbm))
    print("%-26s %9s %9s %9s %9s %8s %8s" % ("kernel", "GMAC", "FMA ms", "HBM MB", "HBM ms", "meas ms", "FMA frac"))
    tot = [0.0, 0.0, 0.0, 0.0]
    for k, (me, mf, re_, rf) in KERNELS.items():
        macs = (me * E + mf * S) * F * F * G
        byts = (re_ * E + rf * S) * F * 4.0 * G
        t_fma, t_hbm = macs / fma_per_s * 1e3, byts / (hbm * 1e9) * 1e3
        meas = (line.get("kernels") or {}).get(k, {}).get("ms_per_step")
        print("%-26s %9.2f %9.3f %9.1f %9.3f %8s %8s" % (k, macs / 1e9, t_fma, byts / 1e6, t_hbm,
              "%.3f" % meas if meas else "-", "%.2f" % (t_fma / meas) if meas else "-"))
        tot[0] += macs; tot[1] += t_fma; tot[2] += byts; tot[3] += meas or 0.0
    print("%-26s %9.2f %9.3f %9.1f %9.3f %8.3f %8.2f" % ("sum of the eight", tot[0] / 1e9, tot[1], tot[2] / 1e6,
          tot[2] / (hbm * 1e9) * 1e3, tot[3], tot[1] / tot[3] if tot[3] else 0))
    print("step: %.3f ms measured; %.3f ms if the FMA pipe ran at 100 %%; %.3f ms if HBM were the bound (algorithmic bytes)"
          % (line["ms_per_step"], tot[1], tot[2] / (hbm * 1e9) * 1e3))
    print("note: the fibre-MLP kernels (k_source_node_*_mma) run 90 % of their MACs on the tensor pipe (3xTF32); their FMA "
          "column is what the FMA pipe WOULD need and is shown for the step sum only")


if __name__ == "__main__":
    main()
