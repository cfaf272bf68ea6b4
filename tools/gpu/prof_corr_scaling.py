"""Correlation at 128 / 256 / 512 frames (F = 4), single-CTA tiles and CTA pairs, for an ncu metrics pass:
how the tensor-pipe-active figure develops once the per-launch fixed costs are amortised over more tiles per CTA."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from master_thesis_b200 import _lib, ops, synth      # noqa: E402

dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731
for frames in (128, 256, 512):
    ft, vt, fr, vr = (dev(t) for t in synth.vgg_feats(7, frames // 4, 4))
    for pair in (0, 1):
        _lib.call("mt_set_tuning", b"MT_CORR_2CTA", pair)
        for rep in range(3):
            ops.corr4d(ft, vt, fr, vr)
    torch.cuda.synchronize()
    del ft, vt, fr, vr
_lib.call("mt_set_tuning", b"MT_CORR_2CTA", -1)
print("ok")
