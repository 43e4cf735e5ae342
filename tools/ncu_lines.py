"""Per-source-line stall samples of the first kernel of an .ncu-rep (needs -lineinfo and --import-source on):
   python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname, hdr, out, total = "", None, [], 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples")
        i_inst = hdr.index("Instructions Executed")
        stall_cols = [(i, n) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
        continue
    if hdr is None or r[0] in ("Function Name",) or not r[0].isdigit():
        continue
    try:
        ns, ni = int(r[i_samp]), int(r[i_inst])
    except ValueError:
        continue
    st = sorted(((int(r[i]) if r[i].isdigit() else 0, n[6:]) for i, n in stall_cols), reverse=True)[:3]
    out.append((ns, ni, fname, r[0], r[1].strip()[:90], st))
    total += ns
out.sort(reverse=True)
print("total samples", total)
for ns, ni, f, ln, src, st in out[:top]:
    print("%6d %5.1f%% inst %10d  %s:%s  %s   %s" % (ns, 100.0 * ns / max(1, total), ni, f, ln, src,
                                                     " ".join("%s=%d" % (n, v) for v, n in st if v)))
