"""Developer probe: where the time of one pipelined CM launch goes (needs a -DMT_DEV_PROBES build:
python -m master_thesis_b200.build -DMT_DEV_PROBES --out=tools/libmt_dev.so)."""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MT_B200_LIB", os.path.join(ROOT, "tools", "libmt_dev.so"))
from master_thesis_b200 import ops, synth, _lib
b, c, f, h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 128, 5, 64, 64
rng = synth.rng(7)
sets = []
for i in range(3):
    cf = torch.from_numpy(rng.standard_normal((b, c, f, h, w)).astype(np.float32)).cuda()
    vt = torch.from_numpy((rng.random_sample((b, 1, 4 * h, 4 * w)) > 0.2).astype(np.float32)).cuda()
    va = torch.from_numpy((rng.random_sample((b, 1, f - 1, 4 * h, 4 * w)) > 0.2).astype(np.float32)).cuda()
    sets.append((cf, vt, va))
lib = _lib.load()
for it in range(5):
    cf, vt, va = sets[it % 3]
    torch.cuda.synchronize()
    if hasattr(lib, "mt_debug_cm_reset"):
        lib.mt_debug_cm_reset()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.cm_match(cf, vt, va)
    e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 2048)()
    lib.mt_debug_cm_probe(buf, 2048)
    t = np.array(buf[:148 * 8], dtype=np.float64).reshape(148, 8)
    buf2 = (ctypes.c_ulonglong * 2048)()
    lib.mt_debug_cm_probe(buf2, -1)
    t2 = np.array(buf2[:148 * 8], dtype=np.float64).reshape(148, 8)
    full = np.array(buf2[:2048], dtype=np.float64)
    t00 = full[1200]
    print("   per sample: last S arrival %s | published %s | first slow C poll %s (us after CTA 0 start)" % (
        np.round((full[1152:1152 + b] - t00) / 1e3, 1), np.round((full[1024:1024 + b] - t00) / 1e3, 1),
        np.round((full[1088:1088 + b] - t00) / 1e3, 1)))
    print("   slow items with flag already set %.1f | ns in slow path %.0f | ns in decode %.0f" % (t2[:, 0].mean(), np.median(t2[:, 1]), np.median(t2[:, 2])))
    print("run %d: event %.1f us | kernel ns med %.0f max %.0f | wait %.0f | S %.0f | C %.0f | slow C items %.1f of items %.1f | "
          "publisher busy %.0f ns over %.1f publishes" % (it, e0.elapsed_time(e1) * 1e3, np.median(t[:, 0]), t[:, 0].max(),
          np.median(t[:, 1]), np.median(t[:, 2]), np.median(t[:, 3]), t[:, 4].mean(), t[:, 5].mean(), np.median(t[:, 6]), t[:, 7].mean()))
