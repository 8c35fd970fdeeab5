"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: python tools/launch_summary.py launches.csv"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
acc = collections.defaultdict(list)
for r in rows:
    name = r[4]
    if not re.search(r"rvdd::|interleave_kernel|poison_on_failure", name):
        continue                                   # torch's own kernels (input synthesis, copies) are not ours
    acc[re.sub(r"\(.*", "", name)].append(float(r[14]))
tot = sum(sum(v) for v in acc.values())
print('# ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:solver_kernel|warp_|gauss_|resample_|gray_|minmax_|setup_|interleave_|poison_|demosaic_|upsample2_|remosaick" -c 300 --csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline')
print("# (warm-up, trace step, 2 timed steps and the end-to-end steps of one bench run; torch input-synthesis kernels filtered out)")
print("# per-launch times are cold-cache and serialised: compare SHARES (unit: ns)")
for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
    print("%-50s launches=%3d  avg=%12.0f  share=%.4f" % (k, len(v), sum(v) / len(v), sum(v) / tot))
