"""Summarise an `ncu --set full` report (read on the CPU box with `ncu -i ... --page raw --csv`).
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
    "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    for r in data:
        print("== %s" % r[idx["Kernel Name"]])
        for w in WANT:
            if w in idx:
                print("   %-66s %16s %s" % (w, r[idx[w]], units[idx[w]]))
        mb = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}     # the two columns may differ in unit
        rd = float(r[idx["dram__bytes_read.sum"]] or 0) * mb.get(units[idx["dram__bytes_read.sum"]], 1.0)
        wr = float(r[idx["dram__bytes_write.sum"]] or 0) * mb.get(units[idx["dram__bytes_write.sum"]], 1.0)
        print("   %-66s %16.3f %s (read+write)" % ("dram traffic", rd + wr, "Mbyte"))
        top = sorted(((float(r[idx[k]] or 0), k) for k in stall if r[idx[k]] not in ("", "n/a")), reverse=True)[:5]
        print("   top stalls (warps per issue): " + ", ".join(
            "%s %.2f" % (k.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), v) for v, k in top))


if __name__ == "__main__":
    main(sys.argv[1])
