"""Where does the cfg3 step go when the launches run back to back?  Times the un-instrumented loop over PREFIXES of the
recorded plan (entries 0..j): the increments are the in-context cost of every C-ABI call (no events between calls)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench                                          # noqa: E402
import master_thesis_b200 as mtb                      # noqa: E402
from master_thesis_b200 import ops                    # noqa: E402

wl = bench.WORKLOADS["cfg3"]()
dev = torch.device("cuda", 0)
dsets = [{k: torch.from_numpy(v).to(dev) for k, v in wl.host_inputs(17 * i).items()} for i in range(3)]
plans, outs = [], []
for d in dsets:
    with ops.record() as plan:
        outs.append(wl.gpu_step(mtb, d))
    plans.append(plan)
torch.cuda.synchronize()
names = plans[0].names()
K = 50


def loop(upto):
    ts = []
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            p = plans[i % 3]
            for j in range(upto + 1):
                p.run_entry(j)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / K)
    return sorted(ts)[1]


def loop_subset(idx):
    ts = []
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            p = plans[i % 3]
            for j in idx:
                p.run_entry(j)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / K)
    return sorted(ts)[1]


if len(sys.argv) > 1 and sys.argv[1] == "subsets":
    print("lib", os.environ.get("MT_B200_LIB", "default"), "MT_PDL", os.environ.get("MT_PDL", "1"))
    for sub in ([6], [5, 6], [1, 6], [0, 6], [2, 6], [4, 6], [1, 2, 6], [1, 5, 6], [1, 2, 3, 4, 5, 6], [0, 1, 2, 3, 4, 5, 6],
                [7], [6, 7], [1, 7], [1, 6, 7]):
        print("  %-26s %7.1f us" % (sub, loop_subset(sub)))
    sys.exit(0)

prev = 0.0
print("lib", os.environ.get("MT_B200_LIB", "default"), "MT_PDL", os.environ.get("MT_PDL", "1"))
for j, n in enumerate(names):
    t = loop(j)
    print("  0..%-2d %-24s prefix %7.1f us   +%6.1f us" % (j, n, t, t - prev))
    prev = t
