"""ncu launch-list CSV (`ncu --metrics gpu__time_duration.sum --csv --log-file X`) -> per-kernel share table.
usage: python scripts/launch_summary.py gpurun_out/launches.csv "<header note>" > profiles/<round>_launches_summary.txt"""
import csv
import re
import sys


def family(name):
    if "gemm_bf16_kernel" in name: return "gemm"
    if "attn_fwd" in name or "attn_tail_q_kernel<0>" in name: return "attention_fwd"
    if "attn_" in name: return "attention_bwd"
    if "ln_fwd" in name: return "layernorm_fwd"
    if "ln_bwd" in name: return "layernorm_bwd"
    if "mico::" in name: return "other(mico)"
    return "torch glue"


def main(path, note=""):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    hdr = rows[0]
    iname, ival, imetric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = {}
    for r in rows[1:]:
        if len(r) <= ival or r[imetric] != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", r[iname])[:110]
        t = float(r[ival].replace(",", "")) / 1e6     # ns -> ms
        a = agg.setdefault(name, [0.0, 0])
        a[0] += t
        a[1] += 1
    total = sum(a[0] for a in agg.values())
    n = sum(a[1] for a in agg.values())
    print(f"# {note}")
    print("# share%   total_ms   launches  kernel")
    for name, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        print(f"{100 * t / total:6.2f} {t:10.3f} {c:7d}  {name}")
    print(f"# total {total:.2f} ms over {n} launches")
    fam = {}
    for name, (t, c) in agg.items():
        fam[family(name)] = fam.get(family(name), 0.0) + t
    print("# family shares: " + ", ".join(f"{k} {100 * v / total:.1f}%" for k, v in sorted(fam.items(), key=lambda kv: -kv[1])))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
