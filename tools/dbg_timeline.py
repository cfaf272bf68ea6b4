"""Developer probe: where the time of one staged-warp launch goes (globaltimer stamps per CTA)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["MT_WARP_DBG"] = "8"
# needs a probe build: python -m master_thesis_b200.build -DMT_DEV_PROBES --out=tools/libmt_dev.so
os.environ.setdefault("MT_B200_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmt_dev.so"))
import master_thesis_b200 as mtb
from master_thesis_b200 import synth, _lib
b, f, h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 4, 256, 256
x, m, _ = synth.frames(3, b, f + 1, h, w)
theta = synth.thetas(4, b * f, 0.1)
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
args = [(d(x[:, :, 1:]), d(m[:, :, 1:]), d(m[:, :, 0]), d(theta)) for _ in range(4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lib = _lib.load()
for it in range(4):
    flush.fill_(it)                      # evict L2
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mtb.cpn_align_tail(*args[it])
    e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 2048)()
    lib.mt_debug_warp_timeline(buf, 2048)
    t = np.array(buf[:148 * 8], dtype=np.int64).reshape(148, 8)
    t0 = t[:, 0].min()
    rel = (t[:, :5] - t0) / 1e3
    print("run %d: event %.1f us | entry spread %.1f | setup+pdl done: med %.1f max %.1f | first TMA issued: med %.1f | "
          "first data: med %.1f max %.1f | done: min %.1f med %.1f max %.1f | tiles/CTA %d..%d"
          % (it, e0.elapsed_time(e1) * 1e3, rel[:, 0].max(), np.median(rel[:, 1]), rel[:, 1].max(), np.median(rel[:, 2]),
             np.median(rel[:, 3]), rel[:, 3].max(), rel[:, 4].min(), np.median(rel[:, 4]), rel[:, 4].max(),
             t[:, 5].min(), t[:, 5].max()))
