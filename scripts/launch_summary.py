"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (one step of the bench workload).
usage: launch_summary.py launches.csv"""
import collections, csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
hdr = next(r for r in csv.reader(open(sys.argv[1])) if r and r[0] == "ID")
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows:
    name = r[ki]
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"rbfmap::<?unnamed>?::", "", name).replace("rbfmap::", "")
    name = re.sub(r"\(.*$", "", name)
    name = re.sub(r"^(pumMapConsistentKernel|pumFactorWarpKernel|pumFactorKernel|pumMapConservativeKernel)<.*", r"\1<all buckets>", name)
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(r[ui], 1e-6)
    tot[name] += v
    cnt[name] += 1
total = sum(tot.values())
print(f"{'kernel':60s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
for k, v in tot.most_common():
    print(f"{k[:60]:60s} {cnt[k]:8d} {v:9.3f} {100 * v / total:6.1f}%")
print(f"{'total':60s} {'':8s} {total:9.3f}")
