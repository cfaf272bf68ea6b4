"""Per-kernel share of one step from an ncu launch list (gpu__time_duration.sum per launch).

    python tools/launch_shares.py gpurun_out/launches.csv > profiles/rNx_launch_shares.txt
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
mu = hdr.index("Metric Unit")
tot, cnt = collections.OrderedDict(), collections.Counter()
for r in rows[1:]:
    if r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    v = float(r[mv].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[mu], 1.0)
    name = r[kn].split("(")[0].replace("void ", "").replace("mt::<unnamed>::", "")
    tot[name] = tot.get(name, 0.0) + v
    cnt[name] += 1
s = sum(tot.values())
print("ncu launch list (cold caches, serialised launches): average us per launch and share of the listed time")
for k, v in tot.items():
    print("  %-70s n=%3d  avg %7.2f us  share %5.1f %%" % (k[:70], cnt[k], v / cnt[k], 100.0 * v / s))
