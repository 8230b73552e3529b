"""Per-kernel totals and shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 14 and r[0].isdigit()]
tot = collections.Counter(); cnt = collections.Counter()
for r in rows:
    name = re.split(r"[<(]", r[4])[0].replace("void ", "").split("::")[-1].strip()
    ns = float(r[14].replace(",", ""))
    tot[name] += ns; cnt[name] += 1
skip = {"synth_fill_kernel"}
total = sum(v for k, v in tot.items() if k not in skip)
print("launches %d (synthetic-input generator excluded from the shares); total %.2f ms" % (len(rows), total / 1e6))
for k, v in tot.most_common():
    if k in skip:
        continue
    print("%-34s n=%5d  %9.3f ms  %5.1f %%  avg %8.1f us" % (k, cnt[k], v / 1e6, 100 * v / total, v / cnt[k] / 1e3))
