"""Small invocations of the kernels that changed in round 2, for `compute-sanitizer --tool memcheck|racecheck`:
grouped CM (one and several samples per group, ragged shapes), correlation (single-CTA tiles and CTA pairs),
dense / low-resolution / staged warp, fused loss forward + backward."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import master_thesis_b200 as mtb                     # noqa: E402
from master_thesis_b200 import _lib, ops, synth      # noqa: E402

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731


def tune(name, v):
    _lib.call("mt_set_tuning", name.encode(), v)


for shape in ((5, 4, 9, 24, 48), (3, 5, 16, 32, 32), (1, 8, 6, 16, 16)):
    b, f, c, h, w = shape
    cf, vt, va = synth.cm_inputs(61 + b, b, f, c, h, w, 4)
    for groups in (0, 1):
        tune("MT_CM_GROUPS", groups)
        ops.cm_match(dev(cf), dev(vt), dev(va))
tune("MT_CM_GROUPS", 0)
ft, vt, fr, vr = synth.vgg_feats(5, 1, 2)
for pair, tm, tn in ((0, 0, 0), (0, 256, 256), (0, 128, 64), (1, 0, 256), (1, 0, 128)):
    tune("MT_CORR_2CTA", pair); tune("MT_CORR_TM", tm); tune("MT_CORR_TN", tn)
    ops.corr4d(dev(ft), dev(vt), dev(fr), dev(vr))
tune("MT_CORR_2CTA", -1); tune("MT_CORR_TM", 0); tune("MT_CORR_TN", 0)
b, f, h, w = 2, 2, 96, 160
x, m, _ = synth.frames(1, b, f + 1, h, w)
xr, mr, mt_ = dev(x[:, :, 1:]), dev(m[:, :, 1:]), dev(m[:, :, 0])
flow = dev(synth.dense_flow(3, b, f, h, w, 0.05, True))
mtb.dfpn_align_tail(xr, mr, mt_, flow)
mtb.cpn_align_tail(xr, mr, mt_, dev(synth.thetas(2, b * f, 0.1)))
x, m, _ = synth.frames(2, 8, 5, 128, 128)          # 8 * 4 * 16 = 512 tiles: the staged kernel runs
mtb.cpn_align_tail(dev(x[:, :, 1:]), dev(m[:, :, 1:]), dev(m[:, :, 0]), dev(synth.thetas(4, 32, 0.1)))
fl = flow.clone().requires_grad_(True)
x2, m2, _ = synth.frames(1, b, f + 1, h, w)
loss = mtb.LossesUtils.alignment_recons(dev(x2[:, :, 0]), dev(1 - m2[:, :, 0]), xr, 1 - mr, fl)
loss.backward()
torch.cuda.synchronize()
print("sanitize_small ok")
