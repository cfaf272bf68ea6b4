"""Event-timed mt_corr4d_vgg_l1_fwd / mt_corr4d_l1_bwd / store-mode correlation at B x 4 frames for the tile knobs given
in the environment (MT_CORR_TM, MT_CORR_TN, MT_CORR_2CTA); rotating inputs (3 sets > L2)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from master_thesis_b200 import ops, synth            # noqa: E402

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
sets = []
for i in range(3):
    ft, vt, fr, vr = synth.vgg_feats(60 + i, B, 4)
    sets.append((dev(synth.rng(70 + i).random_sample((B, 4, 16, 16, 16, 16)).astype(np.float32)), dev(ft), dev(fr)))


def timed(fn, reps=30):
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    ts = []
    for r in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s = sets[r % 3]
        e0.record(); fn(s); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


with torch.no_grad():
    t_store = timed(lambda s: ops.corr4d_vgg(s[1], None, s[2], None))
    t_l1 = timed(lambda s: ops.corr4d_l1_fwd_raw(s[0], s[1], s[2], want_sign=True))
    t_l1ns = timed(lambda s: ops.corr4d_l1_fwd_raw(s[0], s[1], s[2], want_sign=False))
sign = ops.corr4d_l1_fwd_raw(sets[0][0], sets[0][1], sets[0][2])[1]
g = torch.ones((), device="cuda")
gp = torch.empty(sign.shape, dtype=torch.float32, device="cuda")
from master_thesis_b200 import _lib                   # noqa: E402
import ctypes                                          # noqa: E402
p = lambda t: ctypes.c_void_p(t.data_ptr())            # noqa: E731
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
t_bwd = timed(lambda s: _lib.call("mt_corr4d_l1_bwd", p(sign), p(g), p(gp), sign.numel(), st))
print("frames %d  TM %s TN %s 2CTA %s : store %.1f us  l1 %.1f us  l1 (no signs) %.1f us  bwd %.1f us" % (
    B * 4, os.environ.get("MT_CORR_TM", "-"), os.environ.get("MT_CORR_TN", "-"), os.environ.get("MT_CORR_2CTA", "-"),
    t_store, t_l1, t_l1ns, t_bwd))
