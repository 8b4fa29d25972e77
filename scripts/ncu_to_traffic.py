"""profiles/<name>.ncu-rep (ncu --set full capture of representative GEMM launches) -> profiles/gemm_traffic.json,
which bench.py reads for roofline.traffic (dram__bytes_read.sum + dram__bytes_write.sum per launch)."""
import csv
import json
import subprocess
import sys


def main(rep, out, note="representative ViT-g bs64 launches: fc1 fwd (+GELU), fc2 dgrad, fc2 wgrad, fc2 fwd (+residual)", first=4):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]

    def val(r, k):
        i = hdr.index(k)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "%": 1, "": 1}.get(u, 1)

    launches = []
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        if "gemm_bf16_kernel" not in name:
            continue
        launches.append(dict(kernel=name.split("(")[0].split("::")[-1][:60],
                             dram_bytes=val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum"),
                             duration_s=val(r, "gpu__time_duration.sum"),
                             tensor_pipe_pct=val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")))
    rep4 = launches[:int(first)]       # the four representative launches; the rest of the capture is listed for reference
    avg = sum(l["dram_bytes"] for l in rep4) / max(len(rep4), 1)
    json.dump(dict(source=rep, note=note, launches=launches, avg_dram_bytes_per_launch=avg), open(out, "w"), indent=1)
    print(f"{len(launches)} launches, average {avg / 1e6:.1f} MB/launch -> {out}")


if __name__ == "__main__":
    main(*sys.argv[1:])
