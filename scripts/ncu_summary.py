"""Print the key metrics of every kernel in an .ncu-rep (run in the build container: ncu -i needs no GPU)."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor_op_gmma.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def main(path, extra=()):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")][:110])
        for k in list(KEYS) + list(extra):
            for i, h in enumerate(hdr):
                if h == k or (k.endswith("*") and h.startswith(k[:-1])):
                    print(f"   {h:80s} {r[i]:>16s} {units[i]}")
        stalls = [(float(r[i].replace(",", "")), h) for i, h in enumerate(hdr)
                  if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i]]
        tot = sum(v for v, _ in stalls) or 1.0
        top = sorted(stalls, reverse=True)[:5]
        print("   stalls: " + ", ".join(f"{h.split('stalled_')[1]} {100 * v / tot:.0f}%" for v, h in top))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
