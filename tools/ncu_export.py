"""On the GPU box: export the raw page and the hottest source lines of an .ncu-rep as small text files (the report
itself can exceed what gpurun copies back).  usage: python tools/ncu_export.py report.ncu-rep out_prefix"""
import csv
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
open(out + "_raw.csv", "w").write(raw)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(src.splitlines()))
idx = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
with open(out + "_samples_by_line.txt", "w") as f:
    for n, i in enumerate(idx):
        hdr = rows[i]
        seg = rows[i + 1: idx[n + 1] if n + 1 < len(idx) else None]
        si, li, wi, ii = (hdr.index(k) for k in ("# Samples", "stall_long_sb", "stall_wait", "Instructions Executed"))
        lines = [r for r in seg if len(r) == len(hdr) and r[0].isdigit()]
        tot = sum(int(r[si]) for r in lines)
        f.write("launch %d: %d samples\n" % (n, tot))
        if tot < 5000:
            continue
        for r in sorted(lines, key=lambda r: -int(r[si]))[:28]:
            f.write("%5.1f%% line %4s long_sb %6s wait %6s inst %10s | %s\n" %
                    (100 * int(r[si]) / tot, r[0], r[li], r[wi], r[ii], r[1][:110]))
