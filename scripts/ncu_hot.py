"""Top SASS instructions by stall samples from an .ncu-rep source page (run in the build container).
usage: python scripts/ncu_hot.py report.ncu-rep [kernel_index=0] [top=40]"""
import csv, subprocess, sys, io

def main(path, kidx=0, top=40):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    # split per kernel: each kernel section starts with a "Kernel Name" row
    sections, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}
            sections.append(cur)
        elif cur is not None:
            cur["rows"].append(row)
    s = sections[kidx]
    hdr = s["rows"][0]
    rows = s["rows"][1:]
    iS, iSamp, iExec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    total = sum(int(r[iSamp] or 0) for r in rows)
    print(s["name"][:120]); print("instructions", len(rows), "samples", total)
    order = sorted(range(len(rows)), key=lambda i: -int(rows[i][iSamp] or 0))[:top]
    for i in sorted(order):
        r = rows[i]
        st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
        print(f"{i:5d} {100*int(r[iSamp] or 0)/total:5.1f}% exec={r[iExec]:>9s} {r[iS].strip()[:70]:70s} " +
              " ".join(f"{n}:{v}" for v, n in st if v))
    # opcode histogram weighted by executed count
    hist = {}
    for r in rows:
        op = r[iS].strip().split()[0] if r[iS].strip() else "?"
        if op.startswith("@"): op = r[iS].strip().split()[1]
        op = op.split(".")[0]
        hist[op] = hist.get(op, 0) + int(r[iExec] or 0)
    tot = sum(hist.values())
    print("executed warp-instructions by opcode (total %d):" % tot)
    print("  " + ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(hist.items(), key=lambda kv: -kv[1])[:25]))

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40)


def regions(path, kidx, bounds):
    """sum samples / executed instructions over instruction-index ranges"""
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    sections, cur = [], None
    for row in csv.reader(io.StringIO(out)):
        if row and row[0] == "Kernel Name":
            cur = []
            sections.append(cur)
        elif cur is not None:
            cur.append(row)
    hdr, rows = sections[kidx][0], sections[kidx][1:]
    iSamp, iExec = hdr.index("# Samples"), hdr.index("Instructions Executed")
    tot = sum(int(r[iSamp] or 0) for r in rows)
    b = [0] + list(bounds) + [len(rows)]
    for lo, hi in zip(b[:-1], b[1:]):
        s = sum(int(r[iSamp] or 0) for r in rows[lo:hi])
        e = sum(int(r[iExec] or 0) for r in rows[lo:hi])
        print(f"[{lo:5d},{hi:5d}) samples {100*s/tot:5.1f}%  executed {e/1e6:8.2f} M")
