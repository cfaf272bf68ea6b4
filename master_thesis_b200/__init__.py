"""master_thesis_b200 - B200-native (sm_100a) kernels for the frame-alignment /
temporal-copying hot path of davidalvarezdlt/master_thesis, behind the
reference's own Python plug points.

    import master_thesis as mt            # the unmodified reference
    import master_thesis_b200 as mtb
    mtb.patch(mt)                         # rebinds align_set, masked_l1,
                                          # correlation_masked_4d, DFPN.align, CPN.align,
                                          # CM_Module.forward, CHN.forward / inpaint_*

Importing this package does not load the CUDA library; the first operator
call does, and raises if libmt_b200.so is missing (no CPU fallback).
"""
from . import ops, synth  # noqa: F401
from .plug import (CM_Module, CorrelationVGG, FlowsUtils, LossesUtils,  # noqa: F401
                   chn_compute_loss, chn_forward, chn_inpaint_cp, chn_inpaint_ff, chn_inpaint_ip,
                   cpn_align, cpn_align_tail, dfpn_align, dfpn_align_tail, patch,
                   trivial_copy, unpatch)

__version__ = "0.1.0"
