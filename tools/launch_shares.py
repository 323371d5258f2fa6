"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total time, share."""
import csv, collections, re, sys
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr = rows[0]; ik = hdr.index("Kernel Name"); iv = hdr.index("Metric Value"); iu = hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik]
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"^void\s+", "", name)
    if "at::" in name or "elementwise" in name or "nccl" in name.lower():
        name = "[torch] " + name.split("<")[0][-60:]
    v = float(r[iv].replace(",", "")); u = r[iu]
    ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
ours = sum(a[1] for n, a in agg.items() if not n.startswith("[torch]"))
print(f"| kernel | launches | total ms | share of all | share of bgdebias kernels |\n|---|---:|---:|---:|---:|")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    own = "" if n.startswith("[torch]") else f"{100 * a[1] / ours:.1f}%"
    print(f"| `{n}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {own} |")
