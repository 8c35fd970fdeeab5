"""Summarise an `ncu --page source --csv --print-source cuda,sass` export: executed instructions and stall samples
per CUDA source line (top N).  Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > x.csv;
python profiles/ncu_lines.py x.csv [N]"""
import collections
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur, ix, line, text = None, None, None, {}
inst, samp = collections.Counter(), collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        ix = {h: i for i, h in enumerate(r)}
        i_inst, i_samp = r.index("Instructions Executed"), r.index("# Samples")
    elif cur and ix and len(r) > i_inst:
        if r[0].strip().isdigit():
            line = (cur, int(r[0]))
            text[line] = r[1].strip()
        if line and r[2].startswith("0x"):
            try:
                inst[line] += int(r[i_inst])
                samp[line] += int(r[i_samp])
            except ValueError:
                pass
ti, ts = sum(inst.values()), sum(samp.values())
print("total warp-instructions %d, samples %d" % (ti, ts))
for k, v in inst.most_common(top):
    print("%5.2f%% inst %5.2f%% smp  %s:%d  %s" % (100 * v / ti, 100 * samp[k] / max(ts, 1), k[0], k[1], text[k][:100]))
