"""Developer probe: phases of one grouped CM launch (globaltimer stamps per CTA of cm_group_kernel).
Needs a probe build: python -m master_thesis_b200.build -DMT_DEV_PROBES --out=tools/libmt_dev.so"""
import ctypes, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("MT_B200_LIB", os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmt_dev.so"))
from master_thesis_b200 import ops, synth, _lib
b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
sets = [tuple(d(t) for t in synth.cm_inputs(50 + i, b, 5, 128, 64, 64)) for i in range(3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
lib = _lib.load()
lib.mt_debug_cm_timeline.argtypes = [ctypes.c_void_p, ctypes.c_int]
for it in range(5):
    flush.fill_(it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.cm_match(*sets[it % 3])
    e1.record()
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * 8192)()
    lib.mt_debug_cm_timeline(buf, 8192)
    n = 296
    t = np.array(buf[:n * 8], dtype=np.int64).reshape(n, 8)
    t0 = t[:, 0].min()
    rel = (t[:, :5] - t0) / 1e3
    p = lambda v: "min %.1f med %.1f p90 %.1f max %.1f" % (v.min(), np.median(v), np.percentile(v, 90), v.max())
    print("run %d: call %.1f us | entry %s | pass1 done %s | hand-off %s | table %s | pass2 done %s"
          % (it, e0.elapsed_time(e1) * 1e3, p(rel[:, 0]), p(rel[:, 1]), p(rel[:, 2]), p(rel[:, 3]), p(rel[:, 4])))
    print("        durations: pass1 %s | wait %s | fold+table %s | pass2 %s"
          % (p(rel[:, 1] - rel[:, 0]), p(rel[:, 2] - rel[:, 1]), p(rel[:, 3] - rel[:, 2]), p(rel[:, 4] - rel[:, 3])))
