"""Pins the CPU oracle (oracle/mt_oracle.c) to the reference's golden vectors.

The reference has no tests; the golden vectors were produced by running the
unmodified reference in the build container (tests/golden/make_golden.py).
Tolerances: bit-exact for index / mask logic; <= 1e-6 abs for fp32 values
that go through the same operation order; reductions <= 1e-5 relative
(the oracle accumulates in double, torch in blocked fp32).
"""
import numpy as np
import pytest

import cases
import oracle
from conftest import load_golden


@pytest.mark.parametrize("name", sorted(cases.WARP_CASES))
def test_align_set_and_dfpn_tail(name):
    x, m, m_t, flow = cases.warp_inputs(cases.WARP_CASES[name])
    g = load_golden("warp_" + name)
    xa, va, vm = oracle.dfpn_align_tail(x, m, m_t, flow)
    assert np.array_equal(va, g["v_aligned"])          # nearest: bit-exact
    assert np.array_equal(vm, g["v_map"])
    assert np.array_equal(xa, g["x_aligned"])          # same op order + FMA: bit-exact
    xa2, va2 = oracle.align_set(x, 1 - m, flow)
    assert np.array_equal(xa2, xa) and np.array_equal(va2, va)


@pytest.mark.parametrize("name", sorted(cases.CPN_CASES))
def test_cpn_align_tail(name):
    x, m, m_t, theta = cases.cpn_inputs(cases.CPN_CASES[name])
    g = load_golden("cpn_" + name)
    b, _, f, h, w = x.shape
    # (i) with the reference's own dense grid everything is bit-exact
    xa, va, vm = oracle.cpn_align_tail(x, m, m_t, grid=g["grid"])
    assert np.array_equal(xa, g["x_aligned"])
    assert np.array_equal(va, g["v_aligned"])
    assert np.array_equal(vm, g["v_map"])
    # (ii) with theta: the oracle's affine grid follows torch's scalar
    # linspace, ATen's vectorised linspace differs in the last ulp
    grid = oracle.affine_grid(theta, h, w, False).reshape(b, f, h, w, 2)
    assert np.abs(grid - g["grid"]).max() <= 2.5e-7
    xa, va, vm = oracle.cpn_align_tail(x, m, m_t, theta=theta)
    assert np.abs(xa - g["x_aligned"]).max() <= 1e-5
    bad = va != g["v_aligned"]
    # a threshold flip is only legitimate where the soft visibility is within
    # float noise of the 0.5 threshold
    assert np.all(np.abs(g["v_soft"][bad] - 0.5) <= 1e-5)
    assert bad.mean() <= 2e-3
    assert np.array_equal(vm != g["v_map"], bad & (vm != g["v_map"]))


@pytest.mark.parametrize("name", sorted(cases.LOSS_CASES))
def test_losses_and_backward(name):
    x, m, flow, flow_gt, use, t, r_list = cases.loss_inputs(cases.LOSS_CASES[name])
    g = load_golden("loss_" + name)
    f = len(r_list)
    mo = oracle.mask_out(flow)
    assert np.array_equal(mo, g["mask_out"])
    xa, _ = oracle.align_set(x[:, :, r_list], 1 - m[:, :, r_list], flow)
    y_hat = np.repeat(x[:, :, t][:, :, None], f, axis=2)
    mask = np.repeat((1 - m)[:, :, t][:, :, None], f, axis=2) * (1 - mo)
    rec = oracle.masked_l1(y_hat, xa, mask, reduction="sum")
    assert rec == pytest.approx(float(g["recons"]), rel=1e-5)
    mean = oracle.masked_l1(y_hat, xa, mask, reduction="mean", weight=2.0)
    assert mean == pytest.approx(float(g["mean_l1"]), rel=1e-5)
    fl1 = oracle.masked_l1(flow, flow_gt, np.ones_like(flow), batch_mask=use)
    assert fl1 == pytest.approx(float(g["flow_l1"]), rel=1e-5)
    none = oracle.masked_l1(flow, flow_gt, np.ones_like(flow), batch_mask=np.zeros(len(use), bool))
    assert none == 0.0 and g["none_selected"].shape == (1,) and float(g["none_selected"][0]) == 0.0
    gxa = oracle.masked_l1_bwd(y_hat, xa, mask, reduction="sum")
    assert np.abs(gxa - g["g_x_aligned"]).max() <= 1e-6 * max(1.0, np.abs(g["g_x_aligned"]).max())
    gfl = oracle.align_set_bwd_flow(x[:, :, r_list], flow, gxa)
    scale = np.abs(g["g_flow"]).max()
    assert np.abs(gfl - g["g_flow"]).max() <= 2e-5 * scale
    gfl1 = oracle.masked_l1_bwd(flow, flow_gt, np.ones_like(flow), batch_mask=use)
    assert np.abs(-gfl1 - g["g_flow_l1"]).max() <= 1e-6 * np.abs(g["g_flow_l1"]).max() + 1e-12


@pytest.mark.parametrize("name", sorted(cases.CHNLOSS_CASES))
def test_chn_l1_terms(name):
    """The L1 terms of the unmodified CHN.compute_loss (model_chn.py:347-362) and their autograd."""
    y_target, v_target, y_hat, y_comp, v_map = cases.chnloss_inputs(cases.CHNLOSS_CASES[name])
    g = load_golden("chnloss_" + name)
    losses, g_yh, g_yc = oracle.chn_l1_terms(y_target, v_target, y_hat, y_comp, v_map, grads=True)
    assert np.allclose(losses, g["losses"], rtol=1e-5, atol=0)
    for got, ref in ((g_yh, g["g_y_hat"]), (g_yc, g["g_y_hat_comp"])):
        assert np.abs(got - ref).max() <= 1e-6 * max(1e-12, np.abs(ref).max())


@pytest.mark.parametrize("name", sorted(cases.INPAINT_CASES))
def test_inpaint_ff(name):
    """a2 + a9-a11 chained as the unmodified CHN.inpaint_ff chains them (model_chn.py:87-133)."""
    x, m, flows, nn_outs = cases.inpaint_inputs(cases.INPAINT_CASES[name])
    g = load_golden("inpaint_" + name)
    algo = oracle.inpaint_ip if cases.INPAINT_CASES[name].get("algo") == "ip" else oracle.inpaint_ff
    y, steps = algo(x, m, flows, nn_outs, e=cases.INPAINT_CASES[name].get("e", 1))
    assert steps == int(g["steps"][0])
    assert np.array_equal(y, g["y"])


@pytest.mark.parametrize("name", sorted(cases.CORR_CASES))
def test_corr4d(name):
    ft, vt, fr, vr = cases.corr_inputs(cases.CORR_CASES[name])
    g = load_golden("corr_" + name)
    c = oracle.corr4d(ft, vt, fr, vr)
    cases.corr_check(c, g, cases.CORR_CASES[name], 2e-6)      # cosine values are <= 1
    if vt is not None:                            # masked rows / cols are exactly 0
        assert np.all(c[0, :, 0, :] == 0.0)


def test_masked_l1_broadcast_masks():
    """utils.py:166-169 with masks smaller than y_hat: (B,1,H,W) against (B,C,H,W) - the documented shapes -
    and 5-D masks broadcast over F, over B, and inside the plane."""
    y4a, y4b, m4, y5a, y5b, m5f, m5b, m5p = cases.l1_broadcast_inputs()
    g = load_golden("l1_broadcast")
    for key, (ya, yb, mk) in dict(l4=(y4a, y4b, m4), l5f=(y5a, y5b, m5f), l5b=(y5a, y5b, m5b),
                                  l5p=(y5a, y5b, m5p)).items():
        loss, grad = oracle.masked_l1_bcast(ya, yb, mk, "sum", 1.5, grad=True)
        assert loss == pytest.approx(float(g[key]), rel=1e-5)
        assert np.abs(grad - g["g_" + key]).max() <= 1e-6 * np.abs(g["g_" + key]).max()
        assert oracle.masked_l1_bcast(ya, yb, mk, "mean") == pytest.approx(float(g[key + "_mean"]), rel=1e-5)


@pytest.mark.parametrize("name", sorted(cases.LOWRES_CASES))
def test_lowres_flow_align(name):
    """f1: resize_flow(flow_256, (h, w), 'bilinear') + DFPN.align (model_dfpn.py:100-101, 125-133): the
    oracle's resize restates ATen's CPU operation order bit for bit, so everything downstream is exact."""
    x, m, m_t, flow256 = cases.lowres_inputs(cases.LOWRES_CASES[name])
    g = load_golden("lowres_" + name)
    flow = oracle.resize_flow(flow256, x.shape[-2:])
    assert np.array_equal(flow.reshape(-1)[::3], g["flow_sample"])
    xa, va, vm = oracle.dfpn_align_tail(x, m, m_t, flow)
    cases.lowres_check(xa, va, vm, g)


@pytest.mark.parametrize("name", sorted(cases.DFPNLOSS_CASES))
def test_dfpn_training_step(name):
    """DFPN._train_val_wrapper + DFPN.compute_loss (model_dfpn.py:310-394, 210-293): the torch port of the
    call sequence reproduces the unmodified reference, and the C oracle's a1 / a4 / a5 / a6 chained the same
    way reproduce the two reconstruction terms and their flow gradients."""
    import torch
    from oracle import torch_port as tp
    T = torch.from_numpy
    torch.set_num_threads(2)
    spec = cases.DFPNLOSS_CASES[name]
    x, m, y, flow_gt, use, corr, f16, f64, fhw, feats = cases.dfpnloss_inputs(spec)
    g = load_golden("dfpnloss_" + name)
    n = x.shape[2]
    t, r_list = n // 2, [i for i in range(n) if i != n // 2]
    leaves = [T(a).clone().requires_grad_(True) for a in (corr, f16, f64, fhw)]
    res = tp.dfpn_train_val_wrapper(lambda *a: tuple(leaves), T(x), T(m), T(y), T(flow_gt), T(use), t, r_list)
    loss, items = tp.dfpn_compute_loss(lambda inp: [None, None, None, T(feats)], *res, t, r_list)
    assert float(loss) == pytest.approx(float(g["loss"]), rel=1e-6)
    assert np.allclose([float(i) for i in items], g["items"], rtol=1e-6, atol=0)
    grads = torch.autograd.grad(loss, leaves)
    assert np.abs(grads[2].numpy() - g["g_flow64"]).max() <= 1e-6 * np.abs(g["g_flow64"]).max()
    assert np.abs(grads[3].numpy() - g["g_flowhw"]).max() <= 1e-6 * np.abs(g["g_flowhw"]).max()
    assert float(grads[0].abs().double().sum()) == pytest.approx(float(g["g_corr_abs"]), rel=1e-6)
    # the C oracle on the same resized sets
    xs, vs = res[1], res[2]
    for i, (fl, gkey) in ((1, (f64, "g_flow64")), (2, (fhw, "g_flowhw"))):
        xr = xs[i][:, :, r_list].contiguous().numpy()
        vr = vs[i][:, :, r_list].contiguous().numpy()
        xa, _ = oracle.align_set(xr, vr, fl)
        f = len(r_list)
        y_hat = np.repeat(xs[i][:, :, t].numpy()[:, :, None], f, axis=2)
        mask = np.repeat(vs[i][:, :, t].numpy()[:, :, None], f, axis=2) * (1 - oracle.mask_out(fl))
        rec = oracle.masked_l1(y_hat, xa, mask, reduction="sum")
        assert rec == pytest.approx(float(g["items"][3 + i]), rel=1e-5)
        gfl = oracle.align_set_bwd_flow(xr, fl, oracle.masked_l1_bwd(y_hat, xa, mask, reduction="sum"))
        # the golden gradient of this flow also carries its flow-L1 term: remove it (sign / numel per selected item)
        sel = np.asarray(spec["use"], bool)
        fgt = res[6][i].contiguous().numpy()
        gl1 = np.sign(fl - fgt) / (sel.sum() * fl[0].size) * sel[:, None, None, None, None]
        assert np.abs(gfl + gl1.astype(np.float32) - g[gkey]).max() <= 2e-5 * np.abs(g[gkey]).max()


@pytest.mark.parametrize("name", sorted(cases.CM_CASES))
def test_cm_module(name):
    cf, vt, va = cases.cm_inputs(cases.CM_CASES[name])
    g = load_golden("cm_" + name)
    out, cmask = oracle.cm_module(cf, vt, va)
    assert np.abs(cmask - g["c_mask"]).max() <= 2e-6
    if "out" in g:
        assert np.abs(out - g["out"]).max() <= 1e-5
    else:
        assert np.abs(out.reshape(-1)[::53] - g["sample"]).max() <= 1e-5
        assert out.astype(np.float64).sum() == pytest.approx(float(g["total"]), rel=1e-5, abs=1e-2)


@pytest.mark.parametrize("name", sorted(cases.CHN_CASES))
def test_chn(name):
    x_t, v_t, x_al, v_al, v_map, nn_out = cases.chn_inputs(cases.CHN_CASES[name])
    g = load_golden("chn_" + name)
    b, _, f, h, w = x_al.shape
    assert np.array_equal(oracle.chn_pack(x_t, v_t, x_al, v_al, v_map), g["nn_input"])
    yh, yc = oracle.chn_composite(nn_out, x_t, v_t, b, f)
    assert np.array_equal(yh, g["y_hat"])
    assert np.array_equal(yc, g["y_hat_comp"])
    r = cases.synth.rng(cases.CHN_CASES[name]["seed"] + 7)
    gy = r.standard_normal(yh.shape).astype(np.float32)
    gc = r.standard_normal(yh.shape).astype(np.float32)
    gn = oracle.chn_composite_bwd(nn_out, v_t, gy, gc, b, f)
    assert np.abs(gn - g["g_nn_out"]).max() <= 1e-6
    m_new, x_new, per = oracle.hole_update(1 - v_t, v_map[:, :, 0], yc[:, :, 0])
    assert np.array_equal(m_new, g["m_new"])
    assert np.array_equal(x_new, g["x_new"])
    assert per == pytest.approx(float(g["inp_per"]), rel=1e-5)
    assert np.array_equal(oracle.trivial_copy(x_t, x_al, v_map), g["trivial"])


# --------------------------------------------------------------------------
# the torch-CPU port (bench.py's cpu_baseline / --impl reference) is pinned too
# --------------------------------------------------------------------------
def test_torch_port_matches_golden():
    import torch
    from oracle import torch_port as tp
    T = torch.from_numpy
    torch.set_num_threads(2)
    x, m, m_t, flow = cases.warp_inputs(cases.WARP_CASES["noisy_oob"])
    g = load_golden("warp_noisy_oob")
    xa, va, vm = tp.dfpn_align_tail(T(x), T(m), T(m_t), T(flow))
    assert np.array_equal(xa.numpy(), g["x_aligned"]) and np.array_equal(vm.numpy(), g["v_map"])
    x, m, m_t, theta = cases.cpn_inputs(cases.CPN_CASES["rand_f4"])
    g = load_golden("cpn_rand_f4")
    xa, va, vm = tp.cpn_align_tail(T(x), T(m), T(m_t), T(theta))
    assert np.array_equal(xa.numpy(), g["x_aligned"]) and np.array_equal(va.numpy(), g["v_aligned"])
    ft, vt, fr, vr = cases.corr_inputs(cases.CORR_CASES["small_masked"])
    assert np.array_equal(tp.corr4d(T(ft), T(vt), T(fr), T(vr)).numpy(), load_golden("corr_small_masked")["corr"])
    cf, vt, va = cases.cm_inputs(cases.CM_CASES["edge"])
    out, cmask = tp.cm_module(T(cf), T(vt), T(va))
    assert np.array_equal(out.numpy(), load_golden("cm_edge")["out"])
    x_t, v_t, x_al, v_al, v_map, nn_out = cases.chn_inputs(cases.CHN_CASES["f4"])
    g = load_golden("chn_f4")
    assert np.array_equal(tp.chn_pack(T(x_t), T(v_t), T(x_al), T(v_al), T(v_map)).numpy(), g["nn_input"])
    yh, yc = tp.chn_composite(T(nn_out), T(x_t), T(v_t), 2, 4)
    assert np.array_equal(yc.numpy(), g["y_hat_comp"])
    x, m, flow, flow_gt, use, t, r_list = cases.loss_inputs(cases.LOSS_CASES["f4"])
    rec = tp.alignment_recons(T(x)[:, :, t], T(1 - m)[:, :, t], T(x)[:, :, r_list], T(1 - m)[:, :, r_list], T(flow))
    assert float(rec) == float(load_golden("loss_f4")["recons"])


@pytest.mark.parametrize("name", sorted(cases.CORRVGG_CASES))
def test_corr_vgg_forward_neighbours(name):
    """f3: CorrelationVGG.forward between its two networks (model_dfpn.py:516-528): feature permute, nearest
    down-sample of 1 - m to the feature resolution, masked correlation."""
    spec = cases.CORRVGG_CASES[name]
    x_t, m_t, x_r, m_r, ft, fr = cases.corrvgg_inputs(spec)
    g = load_golden("corrvgg_" + name)
    b, f = spec["b"], spec["f"]
    feats_r = fr.reshape(b, f, 512, 16, 16).transpose(0, 2, 1, 3, 4)
    c = oracle.corr4d(ft, oracle.vis_nearest(m_t, (16, 16)), feats_r, oracle.vis_nearest(m_r, (16, 16)))
    cases.corr_check(c, g, spec, 2e-6)
    assert int((c.reshape(b, f, 256, 256) == 0).all(-1).sum()) == int(g["zero_rows"][0])


@pytest.mark.parametrize("name", sorted(cases.FLOWPACK_CASES))
def test_flow_estimator_input_pack(name):
    """8f-4: FlowEstimator.forward's 10-channel `cat` (model_dfpn.py:733-741), C oracle and torch port, every
    flow layout the reference hands over (resize_flow's permuted view, contiguous, a strided window)."""
    import torch
    from oracle import torch_port as tp
    spec = cases.FLOWPACK_CASES[name]
    x_t, m_t, x_r, m_r, base, gain, up = cases.flowpack_inputs(spec)
    g = load_golden("flowpack_" + name)
    flow = cases.flowpack_view(base, spec)
    assert np.array_equal(oracle.flow_pack(x_t, m_t, x_r, m_r, flow), g["nn_input"])
    T = torch.from_numpy
    got = tp.flow_pack(T(x_t), T(m_t), T(x_r), T(m_r), cases.flowpack_view(T(base), spec))
    assert np.array_equal(got.numpy(), g["nn_input"])


@pytest.mark.parametrize("name", sorted(cases.DFPNLOSS_CASES))
def test_corr_l1_of_the_ground_truth_volume(name):
    """8f-3: F.l1_loss(corr, corr_y) (model_dfpn.py:254-257) against the corr_loss item and the gradient the
    unmodified DFPN.compute_loss produced."""
    spec = cases.DFPNLOSS_CASES[name]
    *_, corr, _, _, _, feats = cases.dfpnloss_inputs(spec)
    g = load_golden("dfpnloss_" + name)
    b, n = spec["b"], spec["n"]
    t, r_list = n // 2, [i for i in range(n) if i != n // 2]
    fy = feats.reshape(b, n, -1, 16, 16).transpose(0, 2, 1, 3, 4)
    loss, grad = oracle.corr4d_l1(corr, fy[:, :, t], fy[:, :, r_list])
    assert loss == pytest.approx(float(g["items"][0]), rel=1e-5)
    assert float(np.abs(grad).astype(np.float64).sum()) == pytest.approx(float(g["g_corr_abs"]), rel=1e-5)
