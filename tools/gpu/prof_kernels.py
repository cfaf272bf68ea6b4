"""Launches the three headline kernels a few times on rotating inputs (for ncu -k captures):
corr_tc_kernel at 128 frames, the DFPN direct-gather warp and the staged CPN warp at 32 frames, CM_Module at B = 8;
optionally ("corrl1", "flowpack") the L1-mode correlation with its backward at 128 frames and the FlowEstimator input
pack at 128 frames of 256 x 256."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import master_thesis_b200 as mtb                     # noqa: E402
from master_thesis_b200 import ops, synth            # noqa: E402

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731
which = sys.argv[1:] or ["corr", "dfpn", "cpn", "cm"]
sets = []
for i in range(3):
    d = {}
    if "corr" in which:
        ft, vt, fr, vr = synth.vgg_feats(10 + i, 32, 4)
        d["corr"] = tuple(dev(a) for a in (ft, vt, fr, vr))
    x, m, _ = synth.frames(20 + i, 8, 5, 256, 256)
    d["x"], d["m"], d["mt"] = dev(x[:, :, 1:]), dev(m[:, :, 1:]), dev(m[:, :, 0])
    d["flow"] = dev(synth.dense_flow(30 + i, 8, 4, 256, 256, 0.05, True))
    d["theta"] = dev(synth.thetas(40 + i, 32, 0.1))
    if "cm" in which:
        cf, vt, va = synth.cm_inputs(50 + i, 8, 5, 128, 64, 64)
        d["cm"] = (dev(cf), dev(vt), dev(va))
    if "corrl1" in which:
        ft, vt, fr, vr = synth.vgg_feats(60 + i, 32, 4)
        d["corrl1"] = (dev(synth.rng(70 + i).random_sample((32, 4, 16, 16, 16, 16)).astype(np.float32)), dev(ft), dev(fr))
    if "flowpack" in which:
        xb, mb, _ = synth.frames(80 + i, 32, 5, 256, 256)
        d["flowpack"] = (dev(xb[:, :, 0]), dev(mb[:, :, 0]), dev(xb[:, :, 1:]), dev(mb[:, :, 1:]),
                         dev(synth.dense_flow(90 + i, 32, 4, 256, 256, 0.05, True)))
    sets.append(d)
torch.cuda.synchronize()
for rep in range(3):
    for d in sets:
        if "corr" in which:
            ops.corr4d(*d["corr"])
        if "dfpn" in which:
            mtb.dfpn_align_tail(d["x"], d["m"], d["mt"], d["flow"])
        if "cpn" in which:
            mtb.cpn_align_tail(d["x"], d["m"], d["mt"], d["theta"])
        if "cm" in which:
            ops.cm_match(*d["cm"])
        if "corrl1" in which:
            p = d["corrl1"][0].detach().requires_grad_(True)
            ops.corr4d_l1(p, d["corrl1"][1], d["corrl1"][2]).backward()
        if "flowpack" in which:
            ops.flow_pack(*d["flowpack"])
torch.cuda.synchronize()
print("ok")
