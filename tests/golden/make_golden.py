"""Generates tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference has no tests, fixtures or golden vectors,
so these files are the pin for the oracle (tests/test_oracle_golden.py) and
for the CUDA path (tests/test_gpu_*.py).  Inputs are NOT stored: they are
regenerated bit-identically from ``master_thesis_b200.synth`` (numpy
RandomState).  Every case calls a reference function as is; where the
function needs a CNN (``DFPN.align``, ``CPN.align``, ``CHN.forward``) the
network is replaced by a stand-in object returning preset tensors, so the
reference lines under test (cited per case) run unmodified.

torch 2.11.0+cu128 CPU, fp32, generated on the build container.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from master_thesis_b200 import synth  # noqa: E402
from oracle.ref_import import import_reference  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, OUT)
T = torch.from_numpy


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("%-28s %8.1f KB" % (name, os.path.getsize(path) / 1024.0))


# --------------------------------------------------------------------------
# case tables (shared with the tests through cases.py)
# --------------------------------------------------------------------------
from cases import WARP_CASES, CPN_CASES, CORR_CASES, CM_CASES, CHN_CASES, LOSS_CASES  # noqa: E402
from cases import (warp_inputs, cpn_inputs, corr_inputs, cm_inputs, chn_inputs, loss_inputs,  # noqa: E402
                   chnloss_inputs, CHNLOSS_CASES, inpaint_inputs, INPAINT_CASES, l1_broadcast_inputs,
                   LOWRES_CASES, lowres_inputs, DFPNLOSS_CASES, dfpnloss_inputs, CORRVGG_CASES, corrvgg_inputs,
                   FLOWPACK_CASES, flowpack_inputs, flowpack_view)


def flowpack_goldens(mt):
    """8f-4: the unmodified FlowEstimator.forward (model_dfpn.py:714-744) with its conv stack replaced by a stand-in
    that records the 10-channel input and returns a differentiable function of it (so that the gradient the
    reference's `cat` sends to flow_pre is pinned as well)."""
    FlowEstimator = mt.model_dfpn.FlowEstimator
    for name, spec in FLOWPACK_CASES.items():
        x_t, m_t, x_r, m_r, base, gain, up = flowpack_inputs(spec)
        seen = {}

        class FakeEstimator(object):
            def nn(self, inp):
                seen['nn_input'] = inp.detach().clone()
                return inp[:, 0:2] * 0.5 + inp[:, 8:10] * T(gain)

        tb = T(base).clone().requires_grad_(True)
        out = FlowEstimator.forward(FakeEstimator(), T(x_t), T(m_t), T(x_r), T(m_r), flowpack_view(tb, spec))
        (out * T(up)).sum().backward()
        save("flowpack_" + name, nn_input=seen['nn_input'].numpy(), flow_out=out.detach().contiguous().numpy(),
             g_base=tb.grad.numpy())


def main():
    torch.set_num_threads(1)
    mt = import_reference()
    if "--only-flowpack" in sys.argv:
        return flowpack_goldens(mt)
    flowpack_goldens(mt)
    DFPN = mt.model_dfpn.DFPN
    CorrelationVGG = mt.model_dfpn.CorrelationVGG
    CPN = mt.model_cpn.CPN
    CM_Module = mt.model_cpn.CM_Module
    CHN = mt.model_chn.CHN

    # ---- a1 + a2: FlowsUtils.align_set (utils.py:78-104) and the DFPN.align
    # tail (model_dfpn.py:125-133) with the flow network replaced.
    for name, spec in WARP_CASES.items():
        x, m, mt_, flow = warp_inputs(spec)

        class FakeDFPN(object):
            def __call__(self, *a):
                return None, None, None, T(flow)

        xa, va, vm = DFPN.align(FakeDFPN(), T(x[:, :, 0] * 0), T(mt_), T(x), T(m))
        xa2, va2 = mt.FlowsUtils.align_set(T(x), T(1 - m), T(flow))
        assert torch.equal(xa, xa2) and torch.equal(va, va2)
        save("warp_" + name, x_aligned=xa.contiguous().numpy(), v_aligned=va.contiguous().numpy(),
             v_map=vm.contiguous().numpy())

    # ---- a3: CPN.align (model_cpn.py:31-91) with encoder/regressor replaced.
    for name, spec in CPN_CASES.items():
        x, m, mt_, theta = cpn_inputs(spec)
        b, _, f, h, w = x.shape

        class FakeCPN(object):
            def A_Encoder(self, xx, mm):
                return torch.zeros(xx.size(0), 1, 1, 1)

            def A_Regressor(self, a, bb):
                return T(theta)

        xa, va, vm = CPN.align(FakeCPN(), T(x[:, :, 0] * 0), T(mt_), T(x), T(m))
        grid = torch.nn.functional.affine_grid(T(theta), [b * f, 3, h, w], align_corners=False)
        # the pre-threshold bilinear visibility, to classify near-threshold pixels
        vs = torch.nn.functional.grid_sample(
            1 - T(m).transpose(1, 2).reshape(-1, 1, h, w), grid, align_corners=False)
        save("cpn_" + name, x_aligned=xa.contiguous().numpy(), v_aligned=va.contiguous().numpy(),
             v_map=vm.contiguous().numpy(), grid=grid.numpy().reshape(b, f, h, w, 2),
             v_soft=vs.reshape(b, f, 1, h, w).transpose(1, 2).contiguous().numpy())

    # ---- a4 + a5 + a6: mask_out (model_dfpn.py:269-272), masked_l1
    # (utils.py:139-169) as DFPN.compute_loss builds its arguments
    # (model_dfpn.py:259-287), and autograd of align_set -> masked_l1.
    for name, spec in LOSS_CASES.items():
        x, m, flow, flow_gt, flows_use, t, r_list = loss_inputs(spec)
        xt = T(x)
        vt = 1 - T(m)
        fl = T(flow).clone().requires_grad_(True)
        xa, va = mt.FlowsUtils.align_set(xt[:, :, r_list], vt[:, :, r_list], fl)
        mask_out = ((fl < -1).float() + (fl > 1).float()).sum(4).clamp(0, 1).unsqueeze(1)
        nref = len(r_list)
        y_hat = xt[:, :, t].unsqueeze(2).repeat(1, 1, nref, 1, 1)
        mask = vt[:, :, t].unsqueeze(2).repeat(1, 1, nref, 1, 1) * (1 - mask_out)
        recons = mt.LossesUtils.masked_l1(y_hat, xa, mask, reduction='sum')
        g_xa, = torch.autograd.grad(recons, xa, retain_graph=True)
        g_flow, = torch.autograd.grad(recons, fl, retain_graph=True)
        flow_l1 = mt.LossesUtils.masked_l1(fl, T(flow_gt), torch.ones_like(fl),
                                           T(flows_use))
        g_flow_l1, = torch.autograd.grad(flow_l1, fl, allow_unused=True) \
            if flow_l1.requires_grad else (torch.zeros_like(fl),)
        none_sel = mt.LossesUtils.masked_l1(fl, T(flow_gt), torch.ones_like(fl),
                                            torch.zeros(x.shape[0], dtype=torch.bool))
        mean_l1 = mt.LossesUtils.masked_l1(y_hat, xa, mask, reduction='mean', weight=2)
        save("loss_" + name, mask_out=mask_out.detach().numpy(), recons=recons.detach().numpy(),
             g_x_aligned=g_xa.contiguous().numpy(), g_flow=g_flow.numpy(),
             flow_l1=flow_l1.detach().numpy(), g_flow_l1=g_flow_l1.numpy(),
             none_selected=none_sel.numpy(), mean_l1=mean_l1.detach().numpy())

    # ---- a5 with masks smaller than y_hat (utils.py:139-169: the denominator is torch.sum(mask) of the
    # mask as given, the numerator runs over the broadcast product)
    y4a, y4b, m4, y5a, y5b, m5f, m5b, m5p = l1_broadcast_inputs()
    out = {}
    for key, (ya, yb, mk) in dict(l4=(y4a, y4b, m4), l5f=(y5a, y5b, m5f), l5b=(y5a, y5b, m5b),
                                  l5p=(y5a, y5b, m5p)).items():
        a = T(ya).clone().requires_grad_(True)
        l = mt.LossesUtils.masked_l1(a, T(yb), T(mk), reduction='sum', weight=1.5)
        l.backward()
        out[key] = l.detach().numpy()
        out["g_" + key] = a.grad.numpy()
        out[key + "_mean"] = mt.LossesUtils.masked_l1(T(ya), T(yb), T(mk), reduction='mean').numpy()
    save("l1_broadcast", **out)

    # ---- f1: the DFPN.align tail on the flow the DFPN forward returns for a frame size other than 256 x 256:
    # FlowsUtils.resize_flow(flow_256, (h, w), mode='bilinear') (model_dfpn.py:100-101, utils.py:107-126), then
    # DFPN.align (model_dfpn.py:125-133), both unmodified.
    for name, spec in LOWRES_CASES.items():
        x, m, mt_, flow256 = lowres_inputs(spec)
        h, w = x.shape[-2:]
        flow_hw = mt.FlowsUtils.resize_flow(T(flow256), (h, w), mode='bilinear')

        class FakeDFPN4(object):
            def __call__(self, *a):
                return None, None, None, flow_hw

        xa, va, vm = DFPN.align(FakeDFPN4(), T(x[:, :, 0] * 0), T(mt_), T(x), T(m))
        xa = xa.contiguous().numpy()
        save("lowres_" + name, x_sample=xa.reshape(-1)[::5].copy(), x_total=np.float64(xa.astype(np.float64).sum()),
             v_aligned=va.contiguous().numpy().astype(np.uint8), v_map=vm.contiguous().numpy().astype(np.uint8),
             flow_sample=flow_hw.contiguous().numpy().reshape(-1)[::3].copy())

    # ---- a1 + a4 + a5 + a6 in context: the unmodified DFPN._train_val_wrapper (model_dfpn.py:310-394) and
    # DFPN.compute_loss (:210-293); the DFPN forward and the VGG hand out preset tensors.
    for name, spec in DFPNLOSS_CASES.items():
        x, m, y, flow_gt, use, corr, f16, f64, fhw, feats = dfpnloss_inputs(spec)
        t, r_list = DFPN.get_indexes(x.shape[2])
        leaves = [T(a).clone().requires_grad_(True) for a in (corr, f16, f64, fhw)]

        class FakeDFPN5(object):
            def __call__(self, *a):
                return tuple(leaves)

            def model_vgg(self, inp):
                return [None, None, None, T(feats)]

        fake = FakeDFPN5()
        res = DFPN._train_val_wrapper(fake, T(x), T(m), T(y), T(flow_gt), T(use), t, r_list)
        loss, items = DFPN.compute_loss(fake, *res, t, r_list)
        grads = torch.autograd.grad(loss, leaves)
        xs_al = res[4]
        save("dfpnloss_" + name, loss=loss.detach().numpy(), items=np.array([float(i) for i in items], np.float32),
             g_corr_sample=grads[0].numpy().reshape(-1)[::101].copy(),
             g_corr_abs=np.float64(grads[0].abs().double().sum()),
             g_flow16_abs=np.float64(grads[1].abs().double().sum()),
             g_flow64=grads[2].numpy(), g_flowhw=grads[3].numpy(),
             x16_al_sum=np.float64(xs_al[0].double().sum()), x64_al_sum=np.float64(xs_al[1].double().sum()),
             xhw_al_sum=np.float64(xs_al[2].double().sum()))

    # ---- a7: CorrelationVGG.correlation_masked_4d (model_dfpn.py:534-565)
    for name, spec in CORR_CASES.items():
        ft, vt, fr, vr = corr_inputs(spec)
        c = CorrelationVGG.correlation_masked_4d(
            T(ft), None if vt is None else T(vt), T(fr), None if vr is None else T(vr))
        c = c.numpy()
        if c.size > 300000:   # keep the fixture small: strided sample + checksum
            save("corr_" + name, sample=c.reshape(-1)[::spec.get("stride", 37)].copy(),
                 total=np.float64(c.astype(np.float64).sum()),
                 abs_total=np.float64(np.abs(c).astype(np.float64).sum()))
        else:
            save("corr_" + name, corr=c)

    # ---- f3: the unmodified CorrelationVGG.forward (model_dfpn.py:491-532) with the VGG and the 4-D
    # convolution replaced (preset features, identity): feature permute, nearest mask down-sample, correlation.
    for name, spec in CORRVGG_CASES.items():
        x_t, m_t, x_r, m_r, ft, fr = corrvgg_inputs(spec)
        b = x_t.shape[0]

        class FakeCorrVGG(object):
            use_softmax = False
            conv = staticmethod(lambda c: c)

            def model_vgg(self, inp, normalize_input=True):     # first call: the target, second: the references
                assert normalize_input is False
                self.n = getattr(self, "n", 0) + 1
                return [None, None, None, T(ft) if self.n == 1 else T(fr)]

        c = CorrelationVGG.forward(FakeCorrVGG(), T(x_t), T(m_t), T(x_r), T(m_r)).numpy()
        save("corrvgg_" + name, sample=c.reshape(-1)[::spec["stride"]].copy(),
             total=np.float64(c.astype(np.float64).sum()), abs_total=np.float64(np.abs(c).astype(np.float64).sum()),
             zero_rows=np.array([(c.reshape(c.shape[0], c.shape[1], 256, 256) == 0).all(-1).sum()], np.int64))

    # ---- a8: CM_Module.forward (model_cpn.py:206-254)
    for name, spec in CM_CASES.items():
        cf, vt, va = cm_inputs(spec)
        out, cmask = CM_Module()(T(cf), T(vt), T(va))
        out = out.numpy()
        if out.size > 300000:
            save("cm_" + name, sample=out.reshape(-1)[::53].copy(), c_mask=cmask.numpy(),
                 total=np.float64(out.astype(np.float64).sum()))
        else:
            save("cm_" + name, out=out, c_mask=cmask.numpy())

    # ---- a9 + a10 (+ backward) : CHN.forward (model_chn.py:44-85) with the
    # RRDBNet replaced; a11: hole update via CHN.inpaint_ff (model_chn.py:87-133);
    # a12: trivial copy (model_dfpn.py:427-429).
    for name, spec in CHN_CASES.items():
        x_t, v_t, x_al, v_al, v_map, nn_out = chn_inputs(spec)
        b, _, f, h, w = x_al.shape
        seen = {}

        class FakeCHN(object):
            mean = torch.as_tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)
            std = torch.as_tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)

            def nn(self, inp):
                seen['nn_input'] = inp.detach().clone()
                return seen['nn_out']

        seen['nn_out'] = T(nn_out).clone().requires_grad_(True)
        y_hat, y_comp = CHN.forward(FakeCHN(), T(x_t), T(v_t), T(x_al), T(v_al), T(v_map))
        r = synth.rng(spec['seed'] + 7)
        gy = r.standard_normal(y_hat.shape).astype(np.float32)
        gc = r.standard_normal(y_hat.shape).astype(np.float32)
        (y_hat * T(gy)).sum().add((y_comp * T(gc)).sum()).backward()
        # a11: the three reference lines, model_chn.py:128-131
        m_target = 1 - T(v_t)
        fill_color = torch.as_tensor([0.485, 0.456, 0.406], dtype=torch.float32).view(1, 3, 1, 1)
        m_new = m_target - T(v_map)[:, :, 0]
        x_new = (1 - m_new) * y_comp[:, :, 0] + m_new.repeat(1, 3, 1, 1) * fill_color
        inp_per = torch.sum(m_new) * 100 / m_new.numel()
        # a12: model_dfpn.py:427-429
        triv = T(x_t).unsqueeze(2).repeat(1, 1, f, 1, 1) * (1 - T(v_map)) + T(x_al) * T(v_map)
        save("chn_" + name, nn_input=seen['nn_input'].numpy(), y_hat=y_hat.detach().contiguous().numpy(),
             y_hat_comp=y_comp.detach().contiguous().numpy(), g_nn_out=seen['nn_out'].grad.numpy(),
             m_new=m_new.numpy(), x_new=x_new.detach().numpy(), inp_per=inp_per.numpy(),
             trivial=triv.numpy())

    # ---- a5 inside CHN.compute_loss (model_chn.py:324-375), unmodified; only the two terms outside
    # the hot path are replaced by zeros: perceptual (VGG) and the image-gradient loss.
    zero = lambda *a, **k: (torch.zeros(()), None, None)  # noqa: E731
    orig_p, orig_g = mt.LossesUtils.perceptual, mt.LossesUtils.grad
    mt.LossesUtils.perceptual = staticmethod(zero)
    mt.LossesUtils.grad = staticmethod(lambda *a, **k: torch.zeros(()))
    try:
        for name, spec in CHNLOSS_CASES.items():
            y_target, v_target, y_hat, y_comp, v_map = chnloss_inputs(spec)

            class FakeCHN2(object):
                model_vgg = None

            yh = T(y_hat).clone().requires_grad_(True)
            yc = T(y_comp).clone().requires_grad_(True)
            loss, items = CHN.compute_loss(FakeCHN2(), T(y_target), T(v_target), yh, yc, T(v_map))
            loss.backward()
            save("chnloss_" + name, losses=np.array([float(items[0]), float(items[1]), float(items[2])], np.float32),
                 g_y_hat=yh.grad.numpy(), g_y_hat_comp=yc.grad.numpy())
    finally:
        mt.LossesUtils.perceptual, mt.LossesUtils.grad = orig_p, orig_g

    # ---- a9-a11 in context: the unmodified CHN.inpaint_ff (model_chn.py:87-133) with the unmodified
    # DFPN.align on a short clip; the DFPN forward and the RRDBNet hand out preset tensors in call order.
    for name, spec in INPAINT_CASES.items():
        x, m, flows, nn_outs = inpaint_inputs(spec)

        class FakeDFPN2(object):
            align = DFPN.align

            def __init__(self):
                self.n = 0

            def __call__(self, *a):
                i = self.n
                self.n += 1
                return None, None, None, T(flows[i % len(flows)])

        class FakeCHN3(object):
            mean = torch.as_tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)
            std = torch.as_tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)
            forward = CHN.forward

            def __init__(self):
                self.model_aligner = FakeDFPN2()
                self.k = 0

            def nn(self, inp):
                i = self.k
                self.k += 1
                return T(nn_outs[i % len(nn_outs)])

            def __call__(self, *a):
                return self.forward(*a)

        fake = FakeCHN3()
        if spec.get("algo") == "ip":
            y = CHN.inpaint_ip(fake, T(x).clone(), T(m).clone(), s=1, D=20, e=spec.get("e", 1))
        else:
            y = CHN.inpaint_ff(fake, T(x), T(m), s=1, D=20, e=spec.get("e", 1))
        save("inpaint_" + name, y=y.numpy(), steps=np.array([fake.k], np.int32))


if __name__ == "__main__":
    main()
