"""Case tables + input builders shared by make_golden.py and the tests.

Inputs are regenerated from ``master_thesis_b200.synth`` (numpy RandomState:
bit-identical on every machine); only reference OUTPUTS are stored in the
.npz fixtures.  Edge cases follow SURVEY.md section 8(c): half-pixel ties,
-0.5 / size-0.5 borders, |g| > 1, all-masked target, all-refs-invisible,
batch_mask all-False, v=None in the correlation, F=1 and F=4, non-square.
"""
import numpy as np

from master_thesis_b200 import synth

# a1/a2 -------------------------------------------------------------------
WARP_CASES = {
    # name: b, f, h, w, flow kind
    "smooth_f4": dict(seed=11, b=2, f=4, h=32, w=48, flow="smooth", sigma=0.05),
    "noisy_oob": dict(seed=12, b=2, f=2, h=24, w=40, flow="white", sigma=0.5),
    "ties": dict(seed=13, b=1, f=2, h=20, w=28, flow="ties"),
    "f1_odd": dict(seed=14, b=3, f=1, h=17, w=23, flow="smooth", sigma=0.1),
    "identity": dict(seed=15, b=1, f=1, h=16, w=16, flow="identity"),
}


def warp_inputs(spec):
    b, f, h, w = spec["b"], spec["f"], spec["h"], spec["w"]
    x, m, _ = synth.frames(spec["seed"], b, f + 1, h, w)
    x_refs, m_refs = x[:, :, 1:].copy(), m[:, :, 1:].copy()
    m_target = m[:, :, 0].copy()
    kind = spec["flow"]
    if kind == "smooth":
        flow = synth.dense_flow(spec["seed"] + 1, b, f, h, w, spec["sigma"], True)
    elif kind == "white":
        flow = synth.dense_flow(spec["seed"] + 1, b, f, h, w, spec["sigma"], False)
    elif kind == "ties":
        flow = synth.tie_flow(spec["seed"] + 1, b, f, h, w)
    else:
        flow = np.broadcast_to(synth.identity_grid(h, w, True), (b, f, h, w, 2)).copy()
    return x_refs, m_refs, m_target, flow


# a3 ----------------------------------------------------------------------
CPN_CASES = {
    "rand_f4": dict(seed=21, b=2, f=4, h=32, w=48, sigma=0.1),
    "big_f1": dict(seed=22, b=2, f=1, h=24, w=24, sigma=0.6),
    "halfpix": dict(seed=23, b=1, f=3, h=16, w=20, sigma=0.0),
}


def cpn_inputs(spec):
    b, f, h, w = spec["b"], spec["f"], spec["h"], spec["w"]
    x, m, _ = synth.frames(spec["seed"], b, f + 1, h, w)
    theta = synth.thetas(spec["seed"] + 1, b * f, spec["sigma"])
    if spec["sigma"] == 0.0:
        # pure translations by exactly half / one / 1.5 pixels: the bilinear
        # visibility lands exactly on 0.5 (strict > 0.5 must say 0)
        for i in range(b * f):
            theta[i, 0, 2] = (i + 1) * 1.0 / w
            theta[i, 1, 2] = (i % 2) * 1.0 / h
    return x[:, :, 1:].copy(), m[:, :, 1:].copy(), m[:, :, 0].copy(), theta


# a4/a5/a6 ----------------------------------------------------------------
LOSS_CASES = {
    "f4": dict(seed=31, b=3, frames=5, h=24, w=32, sigma=0.08, use=[1, 0, 1]),
    "f1_oob": dict(seed=32, b=2, frames=2, h=16, w=24, sigma=0.6, use=[1, 1]),
}


def loss_inputs(spec):
    b, n, h, w = spec["b"], spec["frames"], spec["h"], spec["w"]
    x, m, _ = synth.frames(spec["seed"], b, n, h, w)
    t = n // 2                         # model_dfpn.py:471-473
    r_list = [i for i in range(n) if i != t]
    f = len(r_list)
    flow = synth.dense_flow(spec["seed"] + 1, b, f, h, w, spec["sigma"], True)
    flow_gt = synth.dense_flow(spec["seed"] + 2, b, f, h, w, spec["sigma"], True)
    flows_use = np.array(spec["use"], dtype=bool)
    return x, m, flow, flow_gt, flows_use, t, r_list


# a7 ----------------------------------------------------------------------
CORR_CASES = {
    "small_masked": dict(seed=41, b=2, f=2, c=32, h=4, w=4, masked=True),
    "small_nomask": dict(seed=42, b=1, f=3, c=16, h=4, w=4, masked=False),
    "real_masked": dict(seed=43, b=1, f=2, c=512, h=16, w=16, masked=True),
    "real_nomask": dict(seed=44, b=2, f=1, c=512, h=16, w=16, masked=False),
    # batch sizes at which mt_corr4d_fwd picks its wider tiles by itself (corr_tc.cu: TN = 128 from 37 frames,
    # TN = 256 from 74 frames on a 148-SM part); the fixtures hold a strided sample + checksums
    "mid_masked": dict(seed=45, b=10, f=4, c=512, h=16, w=16, masked=True, stride=211),
    "big_masked": dict(seed=46, b=32, f=4, c=512, h=16, w=16, masked=True, stride=997),
}


def corr_inputs(spec):
    ft, vt, fr, vr = synth.vgg_feats(spec["seed"], spec["b"], spec["f"], spec["c"],
                                     spec["h"], spec["w"])
    if not spec["masked"]:
        return ft, None, fr, None
    vt[0, 0, 0, :] = 0.0          # a fully masked row of target pixels
    return ft, vt, fr, vr


def corr_check(c, g, spec, tol):
    """Compares a correlation volume with its golden file: the whole volume, or - for the large cases - the
    strided sample and the two checksums the fixture holds."""
    if "corr" in g:
        assert c.shape == g["corr"].shape
        assert np.abs(c - g["corr"]).max() <= tol
        return
    sample = c.reshape(-1)[::spec.get("stride", 37)]
    assert sample.shape == g["sample"].shape
    assert np.abs(sample - g["sample"]).max() <= tol
    total, abs_total = c.astype(np.float64).sum(), np.abs(c).astype(np.float64).sum()
    assert abs(total - float(g["total"])) <= tol * c.size * 0.5 + 1e-6 * abs(float(g["total"]))
    assert abs(abs_total - float(g["abs_total"])) <= tol * c.size * 0.5 + 1e-6 * abs(float(g["abs_total"]))


# a8 ----------------------------------------------------------------------
CM_CASES = {
    "small": dict(seed=51, b=2, f=3, c=8, h=8, w=8, up=4),
    "edge": dict(seed=52, b=3, f=4, c=4, h=6, w=10, up=4, edge=True),
    "real": dict(seed=53, b=1, f=5, c=128, h=64, w=64, up=4),
}


def cm_inputs(spec):
    cf, vt, va = synth.cm_inputs(spec["seed"], spec["b"], spec["f"], spec["c"], spec["h"],
                                 spec["w"], spec["up"])
    if spec.get("edge"):
        vt[0] = 0.0               # all-masked target: v_sum guard (:221-228)
        va[1] = 0.0               # all refs invisible: masked_sums guard (:251-253)
        va[2, 0, 0] = 0.0         # one invisible ref among visible ones
    return cf, vt, va


# a9..a12 -----------------------------------------------------------------
CHN_CASES = {
    "f4": dict(seed=61, b=2, f=4, h=16, w=24),
    "f1_odd": dict(seed=62, b=1, f=1, h=15, w=21),
}


def chn_inputs(spec):
    b, f, h, w = spec["b"], spec["f"], spec["h"], spec["w"]
    x, m, _ = synth.frames(spec["seed"], b, f + 1, h, w)
    x_t, m_t = x[:, :, 0].copy(), m[:, :, 0].copy()
    flow = synth.dense_flow(spec["seed"] + 1, b, f, h, w, 0.08, True)
    r = synth.rng(spec["seed"] + 2)
    x_al = r.random_sample((b, 3, f, h, w)).astype(np.float32)
    v_al = (r.random_sample((b, 1, f, h, w)) < 0.8).astype(np.float32)
    v_map = np.clip(v_al - (1 - m_t[:, :, None]), 0, 1).astype(np.float32)
    nn_out = synth.nn_output(spec["seed"] + 3, b * f, h, w)
    # put some pre-clamp values exactly on the clamp edges 0 and 1
    nn_out.reshape(-1)[::97] = ((0.0 - 0.485) / 0.229)
    del flow
    return x_t, (1 - m_t).astype(np.float32), x_al, v_al, v_map, nn_out


# a5 in CHN.compute_loss --------------------------------------------------------
CHNLOSS_CASES = {
    "f4": dict(seed=81, b=2, f=4, h=16, w=24),
    "f1_odd": dict(seed=82, b=3, f=1, h=15, w=21),
}


def chnloss_inputs(spec):
    """y_target (b,3,h,w), v_target (b,1,h,w), y_hat, y_hat_comp (b,3,f,h,w), v_map (b,1,f,h,w)."""
    b, f, h, w = spec["b"], spec["f"], spec["h"], spec["w"]
    r = synth.rng(spec["seed"])
    y_target = r.random_sample((b, 3, h, w)).astype(np.float32)
    v_target = (r.random_sample((b, 1, h, w)) < 0.85).astype(np.float32)
    y_hat = r.random_sample((b, 3, f, h, w)).astype(np.float32)
    v_al = (r.random_sample((b, 1, f, h, w)) < 0.8).astype(np.float32)
    v_map = np.clip(v_al - v_target[:, :, None], 0, 1).astype(np.float32)
    y_hat_comp = (v_target[:, :, None] * y_target[:, :, None] + (1 - v_target[:, :, None]) * y_hat).astype(np.float32)
    y_hat.reshape(-1)[::89] = np.repeat(y_target[:, :, None], f, axis=2).reshape(-1)[::89]   # exact ties: sign 0
    return y_target, v_target, y_hat, y_hat_comp, v_map


# a9-a11 in context: CHN.inpaint_ff --------------------------------------------
INPAINT_CASES = {
    "ff_n4": dict(seed=91, n=4, h=15, w=21, k=8),
    "ip_n5": dict(seed=92, n=5, h=18, w=22, k=11, algo="ip"),
    # a threshold that the sprinkled single-pixel holes keep the loops from reaching: several steps per frame
    "ff_n6_long": dict(seed=93, n=6, h=16, w=20, k=13, e=0.05),
    "ip_n6_long": dict(seed=94, n=6, h=16, w=20, k=13, e=0.05, algo="ip"),
}


def inpaint_inputs(spec):
    """x (3,n,h,w), m (1,n,h,w), k dense flows (1,1,h,w,2) and k CNN outputs (1,3,h,w) handed out in
    call order by the stand-ins for the DFPN forward and the RRDBNet."""
    n, h, w, k = spec["n"], spec["h"], spec["w"], spec["k"]
    x, m, _ = synth.frames(spec["seed"], 1, n, h, w)
    flows = [synth.dense_flow(spec["seed"] + 10 + i, 1, 1, h, w, 0.05, True) for i in range(k)]
    nn_outs = [synth.nn_output(spec["seed"] + 40 + i, 1, h, w) for i in range(k)]
    return x[0].copy(), m[0].copy(), flows, nn_outs


def get_indexes_ff(t, max_t, s, D):
    """Reference-frame order of the frame-by-frame algorithm (model_chn.py:460-482)."""
    cand = [r for r in range(max_t) if r != t]
    cand = [r for _, r in sorted((abs(r - t), r) for r in cand)]
    return [r for r in cand if abs(r - t) <= D and abs(r - t) % s == 0]


def get_indexes_ip(t, t_list, s, D):
    """Reference-frame order of the inpaint-and-propagate algorithm (model_chn.py:484-503)."""
    done = list(reversed(t_list[:t_list.index(t)]))
    return done + [r for r in get_indexes_ff(t, len(t_list), s, D) if r not in done]


# a5, broadcast masks (the denominator is torch.sum(mask) of the mask as given) -------------
def l1_broadcast_inputs(seed=35, b=3, c=3, f=4, h=12, w=20):
    """y_hat, y (b,c,h,w) with a (b,1,h,w) mask - the shapes LossesUtils.masked_l1 documents (utils.py:139-157) -
    and 5-D y_hat, y (b,c,f,h,w) with masks broadcast over F, over B, and inside the plane."""
    r = synth.rng(seed)
    y4a = r.random_sample((b, c, h, w)).astype(np.float32)
    y4b = r.random_sample((b, c, h, w)).astype(np.float32)
    m4 = (r.random_sample((b, 1, h, w)) < 0.7).astype(np.float32)
    y5a = r.random_sample((b, c, f, h, w)).astype(np.float32)
    y5b = r.random_sample((b, c, f, h, w)).astype(np.float32)
    m5f = (r.random_sample((b, 1, 1, h, w)) < 0.7).astype(np.float32)       # broadcast over C and F
    m5b = (r.random_sample((1, c, f, h, w)) < 0.7).astype(np.float32)       # broadcast over B
    m5p = (r.random_sample((b, 1, f, 1, 1)) < 0.7).astype(np.float32)       # broadcast inside the plane
    return y4a, y4b, m4, y5a, y5b, m5f, m5b, m5p


# f1: resize_flow fused into the DFPN.align tail ---------------------------------------------------------
LOWRES_CASES = {
    # the DFPN predicts its last flow at 256 x 256 and resizes it to the frame size (model_dfpn.py:100-101)
    "davis4": dict(seed=95, b=1, f=2, h=120, w=214, sigma=0.04),
    "tall": dict(seed=96, b=2, f=1, h=181, w=104, sigma=0.3),
}


def lowres_inputs(spec):
    """x_refs (b,3,f,h,w), m_refs (b,1,f,h,w), m_target (b,1,h,w), flow_256 (b,f,256,256,2)."""
    b, f, h, w = spec["b"], spec["f"], spec["h"], spec["w"]
    x, m, _ = synth.frames(spec["seed"], b, f + 1, h, w)
    flow = synth.dense_flow(spec["seed"] + 1, b, f, 256, 256, spec["sigma"], True)
    return x[:, :, 1:].copy(), m[:, :, 1:].copy(), m[:, :, 0].copy(), flow


def lowres_check(xa, va, vm, g):
    """Bit-exact comparison with a lowres_* golden file (strided sample of x_aligned + its sum, full masks)."""
    assert np.array_equal(va, g["v_aligned"].astype(np.float32))
    assert np.array_equal(vm, g["v_map"].astype(np.float32))
    assert np.array_equal(xa.reshape(-1)[::5], g["x_sample"])
    assert xa.astype(np.float64).sum() == float(g["x_total"])


# a1 + a4 + a5 + a6 in context: DFPN._train_val_wrapper + DFPN.compute_loss ------------------------------
DFPNLOSS_CASES = {
    "n5": dict(seed=101, b=2, n=5, h=40, w=56, sigma=0.3, use=[1, 0]),
    "n2": dict(seed=102, b=3, n=2, h=24, w=36, sigma=0.1, use=[1, 1, 1]),
}


def dfpnloss_inputs(spec):
    """Inputs of DFPN._train_val_wrapper (model_dfpn.py:310-394) and the preset outputs of the networks it
    calls: x, m, y (b,·,n,h,w), flow_gt (b,n,h,w,2), flows_use (b) bool, the DFPN forward's
    (corr (b,f,16,16,16,16), flow_16, flow_64, flow_hw) and the VGG pool-4 features of y (b*n,512,16,16)."""
    b, n, h, w = spec["b"], spec["n"], spec["h"], spec["w"]
    x, m, y = synth.frames(spec["seed"], b, n, h, w)
    f = n - 1
    r = synth.rng(spec["seed"] + 1)
    flow_gt = synth.dense_flow(spec["seed"] + 2, b, n, h, w, 0.05, True)
    corr = r.random_sample((b, f, 16, 16, 16, 16)).astype(np.float32)
    flow_16 = synth.dense_flow(spec["seed"] + 3, b, f, 16, 16, spec["sigma"], True)
    flow_64 = synth.dense_flow(spec["seed"] + 4, b, f, 64, 64, spec["sigma"], True)
    flow_hw = synth.dense_flow(spec["seed"] + 5, b, f, h, w, spec["sigma"], True)
    feats = np.maximum(r.standard_normal((b * n, 512, 16, 16)), 0).astype(np.float32)
    return x, m, y, flow_gt, np.array(spec["use"], dtype=bool), corr, flow_16, flow_64, flow_hw, feats


# f3: CorrelationVGG.forward around the correlation (model_dfpn.py:491-532) ---------------------------
CORRVGG_CASES = {
    "f3": dict(seed=111, b=2, f=3, h=64, w=96, stride=53),
    "f1_odd": dict(seed=112, b=1, f=1, h=50, w=70, stride=17),
}


def corrvgg_inputs(spec):
    """x_target (b,3,h,w), m_target (b,1,h,w), x_refs (b,3,f,h,w), m_refs (b,1,f,h,w) and the VGG pool-4
    features the stand-in network hands out: target (b,512,16,16), references (b*f,512,16,16)."""
    b, f, h, w = spec["b"], spec["f"], spec["h"], spec["w"]
    x, m, _ = synth.frames(spec["seed"], b, f + 1, h, w)
    r = synth.rng(spec["seed"] + 1)
    ft = np.maximum(r.standard_normal((b, 512, 16, 16)), 0).astype(np.float32)
    fr = np.maximum(r.standard_normal((b * f, 512, 16, 16)), 0).astype(np.float32)
    return x[:, :, 0].copy(), m[:, :, 0].copy(), x[:, :, 1:].copy(), m[:, :, 1:].copy(), ft, fr


# 8f-4: FlowEstimator.forward input pack (model_dfpn.py:714-744) ------------------------------------------
FLOWPACK_CASES = {
    # layout of flow_pre: "planar" = what FlowsUtils.resize_flow returns (a permuted view of (b,f,2,h,w)),
    # "interleaved" = contiguous (b,f,h,w,2), "sliced" = a window of a wider interleaved tensor (generic strides)
    "planar_f4": dict(seed=121, b=2, f=4, h=16, w=24, layout="planar"),
    "inter_f2": dict(seed=122, b=3, f=2, h=12, w=20, layout="interleaved"),
    "odd_f1": dict(seed=123, b=1, f=1, h=15, w=21, layout="planar"),
    "sliced_f3": dict(seed=124, b=2, f=3, h=8, w=12, layout="sliced"),
}


def flowpack_inputs(spec):
    """x_target (b,3,h,w), m_target (b,1,h,w), x_refs (b,3,f,h,w), m_refs (b,1,f,h,w), flow_pre (b,f,h,w,2) in the
    case's memory layout (a numpy view; `flowpack_torch_flow` rebuilds the same view on a torch tensor), and the
    two tensors of the conv-stack stand-in: gain (b*f,2,h,w) and the upstream gradient (b,f,h,w,2)."""
    b, f, h, w = spec["b"], spec["f"], spec["h"], spec["w"]
    x, m, _ = synth.frames(spec["seed"], b, f + 1, h, w)
    r = synth.rng(spec["seed"] + 1)
    if spec["layout"] == "planar":
        base = r.standard_normal((b, f, 2, h, w)).astype(np.float32)
    elif spec["layout"] == "interleaved":
        base = r.standard_normal((b, f, h, w, 2)).astype(np.float32)
    else:
        base = r.standard_normal((b, f, h + 1, w + 3, 2)).astype(np.float32)
    gain = r.standard_normal((b * f, 2, h, w)).astype(np.float32)
    up = r.standard_normal((b, f, h, w, 2)).astype(np.float32)
    return x[:, :, 0].copy(), m[:, :, 0].copy(), x[:, :, 1:].copy(), m[:, :, 1:].copy(), base, gain, up


def flowpack_view(base, spec):
    """The (b,f,h,w,2) view of ``base`` (numpy array or torch tensor) the case hands to FlowEstimator.forward."""
    if spec["layout"] == "planar":
        return base.transpose(0, 1, 3, 4, 2) if isinstance(base, np.ndarray) else base.permute(0, 1, 3, 4, 2)
    if spec["layout"] == "interleaved":
        return base
    return base[:, :, 1:, 2:2 + spec["w"]]
