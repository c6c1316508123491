"""ncu launch list (csv, --metrics gpu__time_duration.sum) -> markdown table: python scripts/summarize_launches.py launches.csv > summary.md"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
iu = hdr.index("Metric Unit")
t = collections.defaultdict(float); n = collections.Counter()
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    v = float(r[iv].replace(",", ""))
    v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r[iu], 1.0)
    name = re.sub(r"<.*", "", r[ik]).strip()
    name = re.sub(r"\(.*", "", name)
    t[name] += v; n[name] += 1
tot = sum(t.values())
print(f"launches: {sum(n.values())}; summed device time {tot / 1e3:.1f} ms (cold-cache, serialised under ncu: compare SHARES, not absolutes)\n")
gd = {k: v for k, v in t.items() if "gd::" in k}
print(f"hand-written `gd::` kernels: {sum(gd.values()) / 1e3:.1f} ms = {100 * sum(gd.values()) / tot:.1f} % of device time\n")
print("| device ms | share | launches | avg us | kernel |\n|---:|---:|---:|---:|---|")
for k, v in sorted(t.items(), key=lambda x: -x[1])[:40]:
    print(f"| {v / 1e3:.2f} | {100 * v / tot:.1f}% | {n[k]} | {v / n[k]:.1f} | `{k[:70]}` |")
print("\n## gd:: kernels only\n\n| device ms | share of gd | launches | avg us | kernel |\n|---:|---:|---:|---:|---|")
for k, v in sorted(gd.items(), key=lambda x: -x[1]):
    print(f"| {v / 1e3:.2f} | {100 * v / sum(gd.values()):.1f}% | {n[k]} | {v / n[k]:.1f} | `{k[:70]}` |")
