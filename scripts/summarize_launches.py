"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel share table."""
import csv, sys, re, collections
src, dst = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "")
    t = float(r[vi].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t
tot = sum(a[1] for a in agg.values())
with open(dst, "w") as f:
    f.write(f"# ncu launch list summary ({src.split('/')[-1]})\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 1 --warmup 1 --ni-total 16384 --no-cpu-baseline`\n")
    f.write("(per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes)\n\n")
    f.write("| kernel | launches | total ms | mean us | share |\n|---|---|---|---|---|\n")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"| {k} | {n} | {t*1e-6:.3f} | {t/n*1e-3:.1f} | {100*t/tot:.2f} % |\n")
print(open(dst).read())
