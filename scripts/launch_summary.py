"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt, mx = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float)
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("mprg::", "")
    v = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1.0)
    tot[name] += v
    cnt[name] += 1
    mx[name] = max(mx[name], v)
T = sum(tot.values())
print(f"{sys.argv[1]}: {sum(cnt.values())} launches, {T:.1f} us of kernel time")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    print(f"{v:10.1f} us {100 * v / T:5.1f}%  n={cnt[k]:4d}  avg {v / cnt[k]:8.1f}  max {mx[k]:8.1f}  {k}")
