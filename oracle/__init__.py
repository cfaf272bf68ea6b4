"""CPU oracle for the frame-alignment hot path - TEST INFRASTRUCTURE ONLY.

``oracle/mt_oracle.c`` is the restatement (each function cites the reference
file:line it follows); this module is its numpy/ctypes face.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
``master_thesis_b200`` never does.

Parity pinning: the reference has no tests or golden vectors of its own
(SURVEY.md section 4), so the oracle is pinned against outputs of the
UNMODIFIED reference run in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``,
checked by ``tests/test_oracle_golden.py``).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmt_oracle.so")
_lib = None

_f = ctypes.POINTER(ctypes.c_float)
_u8 = ctypes.POINTER(ctypes.c_uint8)
_d = ctypes.POINTER(ctypes.c_double)
_i = ctypes.c_int
_l = ctypes.c_int64


def build(force=False):
    """Compiles oracle/libmt_oracle.so with the committed Makefile."""
    src = os.path.join(_HERE, "mt_oracle.c")
    if (not force and os.path.exists(_LIB_PATH)
            and os.path.getmtime(_LIB_PATH) >= os.path.getmtime(src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-B", "libmt_oracle.so"],
                          stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = ctypes.CDLL(_LIB_PATH)
        L.mto_masked_l1.restype = ctypes.c_float
        L.mto_hole_update.restype = ctypes.c_float
        L.mto_num_threads.restype = _i
        _lib = L
    return _lib


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return None if a is None else a.ctypes.data_as(_f)


def num_threads():
    return lib().mto_num_threads()


def set_num_threads(n):
    lib().mto_set_num_threads(_i(n))


def grid_sample(inp, grid, mode="bilinear", align_corners=True):
    inp, grid = _c(inp), _c(grid)
    n, c, h, w = inp.shape
    _, ho, wo, _ = grid.shape
    out = np.empty((n, c, ho, wo), np.float32)
    lib().mto_grid_sample(_p(inp), _p(grid), _i(n), _i(c), _i(h), _i(w), _i(ho), _i(wo),
                          _i(0 if mode == "bilinear" else 1), _i(int(align_corners)), _p(out))
    return out


def affine_grid(theta, h, w, align_corners=False):
    theta = _c(theta)
    n = theta.shape[0]
    out = np.empty((n, h, w, 2), np.float32)
    lib().mto_affine_grid(_p(theta), _i(n), _i(h), _i(w), _i(int(align_corners)), _p(out))
    return out


def align_set(x, v, flow):
    """a1 - utils.py:78-104."""
    x, v, flow = _c(x), _c(v), _c(flow)
    b, c, f, h, w = x.shape
    xa = np.empty_like(x)
    va = np.empty((b, 1, f, h, w), np.float32)
    lib().mto_align_set(_p(x), _p(v), _p(flow), _i(b), _i(c), _i(f), _i(h), _i(w), _p(xa), _p(va))
    return xa, va


def dfpn_align_tail(x_refs, m_refs, m_target, flow):
    """a2 - model_dfpn.py:128-133."""
    x_refs, m_refs, m_target, flow = _c(x_refs), _c(m_refs), _c(m_target), _c(flow)
    b, c, f, h, w = x_refs.shape
    xa = np.empty_like(x_refs)
    va = np.empty((b, 1, f, h, w), np.float32)
    vm = np.empty((b, 1, f, h, w), np.float32)
    lib().mto_dfpn_align_tail(_p(x_refs), _p(m_refs), _p(m_target), _p(flow), _i(b), _i(c), _i(f),
                              _i(h), _i(w), _p(xa), _p(va), _p(vm))
    return xa, va, vm


def cpn_align_tail(x_refs, m_refs, m_target, theta=None, grid=None):
    """a3 - model_cpn.py:75-89 (theta (b*f,2,3), or an explicit dense grid)."""
    x_refs, m_refs, m_target = _c(x_refs), _c(m_refs), _c(m_target)
    theta = None if theta is None else _c(theta)
    grid = None if grid is None else _c(grid)
    b, c, f, h, w = x_refs.shape
    xa = np.empty_like(x_refs)
    va = np.empty((b, 1, f, h, w), np.float32)
    vm = np.empty((b, 1, f, h, w), np.float32)
    lib().mto_cpn_align_tail(_p(x_refs), _p(m_refs), _p(m_target), _p(theta), _p(grid), _i(b),
                             _i(c), _i(f), _i(h), _i(w), _p(xa), _p(va), _p(vm))
    return xa, va, vm


def mask_out(flow):
    """a4 - model_dfpn.py:269-272; flow (b,f,h,w,2) -> (b,1,f,h,w)."""
    flow = _c(flow)
    b, f, h, w, _ = flow.shape
    out = np.empty((b, 1, f, h, w), np.float32)
    lib().mto_mask_out(_p(flow), _l(flow.size // 2), _p(out))
    return out


def _l1_args(y_hat, y, mask, batch_mask):
    y_hat, y, mask = _c(y_hat), _c(y), _c(mask)
    b, c = y_hat.shape[0], y_hat.shape[1]
    inner = int(np.prod(y_hat.shape[2:]))
    if mask.shape == y_hat.shape:
        mask_c = c
    else:
        assert mask.shape[0] == b and mask.shape[1] == 1 and mask.shape[2:] == y_hat.shape[2:]
        mask_c = 1
    bm = None
    if batch_mask is not None:
        bm = np.ascontiguousarray(np.asarray(batch_mask).astype(np.uint8))
    return y_hat, y, mask, b, c, inner, mask_c, bm


def masked_l1(y_hat, y, mask, batch_mask=None, reduction="mean", weight=1.0):
    """a5 - utils.py:139-169.  Returns a python float."""
    y_hat, y, mask, b, c, inner, mask_c, bm = _l1_args(y_hat, y, mask, batch_mask)
    return float(lib().mto_masked_l1(
        _p(y_hat), _p(y), _p(mask), _i(b), _i(c), _l(inner), _i(mask_c),
        None if bm is None else bm.ctypes.data_as(_u8), _i(1 if reduction == "sum" else 0),
        ctypes.c_float(weight), None))


def masked_l1_bcast(y_hat, y, mask, reduction="mean", weight=1.0, grad=False):
    """a5 with a mask that broadcasts against y_hat (utils.py:166-169): the numerator runs over the broadcast
    product, the 'sum' denominator is sum(mask) of the mask AS GIVEN.  Returns the loss (and d loss / d y_hat)."""
    y_hat, y, mask = _c(y_hat), _c(y), _c(mask)
    full = np.ascontiguousarray(np.broadcast_to(mask, y_hat.shape))
    b = y_hat.shape[0]
    sums = (ctypes.c_double * 3)()
    inner = int(np.prod(y_hat.shape[1:]))
    lib().mto_masked_l1(_p(y_hat), _p(y), _p(full), _i(b), _i(1), _l(inner), _i(1), None, _i(1),
                        ctypes.c_float(1.0), sums)
    num = np.float32(sums[0])
    if reduction == "sum":
        den = np.float32(np.float32(mask.astype(np.float64).sum()) + np.float32(1e-9))
        loss, scale = np.float32(weight) * num / den, np.float32(weight) / den
    else:
        loss, scale = np.float32(weight) * np.float32(sums[0] / y_hat.size), np.float32(weight) / np.float32(y_hat.size)
    if not grad:
        return float(loss)
    d = y_hat * full - y * full
    return float(loss), (np.sign(d) * full * scale).astype(np.float32)


def masked_l1_bwd(y_hat, y, mask, batch_mask=None, reduction="mean", weight=1.0, grad_out=1.0):
    """Gradient of a5 w.r.t. ``y`` (grad w.r.t. ``y_hat`` is its negation)."""
    y_hat, y, mask, b, c, inner, mask_c, bm = _l1_args(y_hat, y, mask, batch_mask)
    g = np.empty_like(y)
    lib().mto_masked_l1_bwd(
        _p(y_hat), _p(y), _p(mask), _i(b), _i(c), _l(inner), _i(mask_c),
        None if bm is None else bm.ctypes.data_as(_u8), _i(1 if reduction == "sum" else 0),
        ctypes.c_float(weight), ctypes.c_float(grad_out), _p(g))
    return g


def align_set_bwd_flow(x, flow, gout, align_corners=True):
    """a6 - gradient of a1's bilinear warp w.r.t. the flow."""
    x, flow, gout = _c(x), _c(flow), _c(gout)
    b, c, f, h, w = x.shape
    g = np.empty_like(flow)
    lib().mto_align_set_bwd_flow(_p(x), _p(flow), _p(gout), _i(b), _i(c), _i(f), _i(h), _i(w),
                                 _i(int(align_corners)), _p(g))
    return g


def corr4d(ft, vt, fr, vr):
    """a7 - model_dfpn.py:534-565."""
    ft, fr = _c(ft), _c(fr)
    vt = None if vt is None else _c(vt)
    vr = None if vr is None else _c(vr)
    b, c, f, h, w = fr.shape
    out = np.empty((b, f, h, w, h, w), np.float32)
    lib().mto_corr4d(_p(ft), _p(vt), _p(fr), _p(vr), _i(b), _i(c), _i(f), _i(h * w), _p(out))
    return out


def cm_module(c_feats, v_t, v_aligned, return_gs=False):
    """a8 - model_cpn.py:206-254."""
    c_feats, v_t, v_aligned = _c(c_feats), _c(v_t), _c(v_aligned)
    b, c, f, h, w = c_feats.shape
    H, W = v_t.shape[-2:]
    out = np.empty((b, 2 * c + 1, h, w), np.float32)
    cmask = np.empty((b, 1, h, w), np.float32)
    gs = np.empty((b, f - 1), np.float32)
    lib().mto_cm_module(_p(c_feats), _p(v_t), _p(v_aligned), _i(b), _i(c), _i(f), _i(h), _i(w),
                        _i(H), _i(W), _p(out), _p(cmask), _p(gs))
    return (out, cmask, gs) if return_gs else (out, cmask)


def chn_pack(x_t, v_t, x_al, v_al, v_map):
    """a9 - model_chn.py:68-80."""
    x_t, v_t, x_al, v_al, v_map = _c(x_t), _c(v_t), _c(x_al), _c(v_al), _c(v_map)
    b, _, f, h, w = x_al.shape
    out = np.empty((b * f, 9, h, w), np.float32)
    lib().mto_chn_pack(_p(x_t), _p(v_t), _p(x_al), _p(v_al), _p(v_map), _i(b), _i(f), _l(h * w),
                       _p(out))
    return out


def corr4d_l1(pred, feats_t, feats_r):
    """8f-3 - F.l1_loss(corr, corr_y) with corr_y the unmasked correlation of the ground-truth features,
    model_dfpn.py:254-257.  -> (loss as float64, d loss / d pred as float32)."""
    corr_y = corr4d(feats_t, None, feats_r, None).astype(np.float64)
    d = np.asarray(pred, np.float64) - corr_y
    return float(np.abs(d).mean()), (np.sign(d) / d.size).astype(np.float32)


def flow_pack(x_target, m_target, x_refs, m_refs, flow_pre):
    """8f-4 - FlowEstimator.forward's nn_input, model_dfpn.py:733-741."""
    x_target, m_target, x_refs, m_refs, flow_pre = _c(x_target), _c(m_target), _c(x_refs), _c(m_refs), _c(flow_pre)
    b, _, f, h, w = x_refs.shape
    out = np.empty((b * f, 10, h, w), np.float32)
    lib().mto_flow_pack(_p(x_refs), _p(x_target), _p(m_refs), _p(m_target), _p(flow_pre), _i(b), _i(f), _l(h * w),
                        _p(out))
    return out


def chn_composite(nn_out, x_t, v_t, b, f):
    """a10 - model_chn.py:80-85."""
    nn_out, x_t, v_t = _c(nn_out), _c(x_t), _c(v_t)
    h, w = nn_out.shape[-2:]
    yh = np.empty((b, 3, f, h, w), np.float32)
    yc = np.empty((b, 3, f, h, w), np.float32)
    lib().mto_chn_composite(_p(nn_out), _p(x_t), _p(v_t), _i(b), _i(f), _l(h * w), _p(yh), _p(yc))
    return yh, yc


def chn_composite_bwd(nn_out, v_t, g_yhat, g_comp, b, f):
    nn_out, v_t = _c(nn_out), _c(v_t)
    g_yhat = None if g_yhat is None else _c(g_yhat)
    g_comp = None if g_comp is None else _c(g_comp)
    h, w = nn_out.shape[-2:]
    g = np.empty_like(nn_out)
    lib().mto_chn_composite_bwd(_p(nn_out), _p(v_t), _p(g_yhat), _p(g_comp), _i(b), _i(f),
                                _l(h * w), _p(g))
    return g


def hole_update(m_t, v_map0, y_comp0):
    """a11 - model_chn.py:128-131."""
    m_t, v_map0, y_comp0 = _c(m_t), _c(v_map0), _c(y_comp0)
    b = m_t.shape[0]
    P = int(np.prod(m_t.shape[2:]))
    m_new = np.empty_like(m_t)
    x_new = np.empty_like(y_comp0)
    per = lib().mto_hole_update(_p(m_t), _p(v_map0), _p(y_comp0), _i(b), _l(P), _p(m_new),
                                _p(x_new))
    return m_new, x_new, float(per)


def trivial_copy(x_t, x_al, v_map):
    """a12 - model_dfpn.py:427-429."""
    x_t, x_al, v_map = _c(x_t), _c(x_al), _c(v_map)
    b, _, f, h, w = x_al.shape
    y = np.empty_like(x_al)
    lib().mto_trivial_copy(_p(x_t), _p(x_al), _p(v_map), _i(b), _i(f), _l(h * w), _p(y))
    return y


def resize_flow(flow, size):
    """f1 - FlowsUtils.resize_flow(flow, size, mode='bilinear') (utils.py:107-126); flow (b,f,h,w,2)."""
    flow = _c(flow)
    b, f, h, w, _ = flow.shape
    H, W = size
    out = np.empty((b, f, H, W, 2), np.float32)
    lib().mto_resize_flow(_p(flow), _i(b * f), _i(h), _i(w), _i(H), _i(W), _p(out))
    return out


def vis_nearest(m, size):
    """f3 - F.interpolate(1 - m, size, mode='nearest') (model_dfpn.py:521-526); m (..., H, W)."""
    m = _c(m)
    H, W = m.shape[-2:]
    h, w = size
    out = np.empty(m.shape[:-2] + (h, w), np.float32)
    lib().mto_vis_nearest(_p(m), _i(int(np.prod(m.shape[:-2]))), _i(H), _i(W), _i(h), _i(w), _p(out))
    return out


def chn_l1_terms(y_target, v_target, y_hat, y_hat_comp, v_map, weights=(0.5, 2.0, 1.0), grads=False):
    """The three masked-L1 terms of CHN.compute_loss, model_chn.py:347-362, as three a5 calls.
    Returns [loss_nh, loss_vh, loss_nvh] (+ grads w.r.t. y_hat and y_hat_comp for upstream grads of 1)."""
    f = y_hat.shape[2]
    target_img = np.repeat(y_target[:, :, None], f, axis=2)
    nh = np.repeat(v_target[:, :, None], f, axis=2).astype(np.float32)
    vh = np.ascontiguousarray(v_map, dtype=np.float32)
    nvh = ((1 - nh) - vh).astype(np.float32)
    losses = [masked_l1(y_hat, target_img, nh, reduction="sum", weight=weights[0]),
              masked_l1(y_hat, target_img, vh, reduction="sum", weight=weights[1]),
              masked_l1(y_hat_comp, target_img, nvh, reduction="sum", weight=weights[2])]
    if not grads:
        return losses
    # masked_l1_bwd returns the gradient w.r.t. y; the one w.r.t. y_hat is its negation
    g_yh = -(masked_l1_bwd(y_hat, target_img, nh, reduction="sum", weight=weights[0]) +
             masked_l1_bwd(y_hat, target_img, vh, reduction="sum", weight=weights[1]))
    g_yc = -masked_l1_bwd(y_hat_comp, target_img, nvh, reduction="sum", weight=weights[2])
    return losses, g_yh, g_yc


def inpaint_ff(x, m, flows, nn_outs, s=1, D=20, e=1):
    """CHN.inpaint_ff (model_chn.py:87-133) with the DFPN aligner, as a loop over the oracle's a2, a9-a11;
    the DFPN forward and the RRDBNet are the preset tensors ``flows`` / ``nn_outs`` in call order.
    x (3,n,h,w), m (1,n,h,w).  Returns (y_inpainted (3,n,h,w), number of steps)."""
    n = x.shape[1]
    y = np.zeros_like(x)
    k = 0
    for t in range(n):
        x_t, m_t = x[None, :, t].copy(), m[None, :, t].copy()
        cand = [r for r in range(n) if r != t]
        cand = [r for _, r in sorted((abs(r - t), r) for r in cand)]
        cand = [r for r in cand if abs(r - t) <= D and abs(r - t) % s == 0]      # :460-482
        y_comp, per = None, 0.0
        while y_comp is None or (len(cand) > 0 and per > e):                     # :111-112
            r = cand.pop(0)
            x_al, v_al, v_map = dfpn_align_tail(x[None, :, r:r + 1], m[None, :, r:r + 1], m_t, flows[k % len(flows)])
            nn_in = chn_pack(x_t, 1 - m_t, x_al, v_al, v_map)
            assert nn_in.shape[1] == 9
            _, yc = chn_composite(nn_outs[k % len(nn_outs)], x_t, 1 - m_t, 1, 1)
            k += 1
            m_t, x_t, per = hole_update(m_t, v_map[:, :, 0], yc[:, :, 0])         # :128-131
            y_comp = yc[:, :, 0]
        y[:, t] = y_comp[0]
    return y, k


def inpaint_ip(x, m, flows, nn_outs, s=1, D=20, e=1):
    """CHN.inpaint_ip (model_chn.py:135-189) with the DFPN aligner as a loop over the oracle's a2, a9-a11 (see
    inpaint_ff).  x (3,n,h,w), m (1,n,h,w) are not modified.  Returns (y (3,n,h,w), number of steps)."""
    y_inp, m_inp = x.copy(), m.copy()
    n = x.shape[1]
    order = sorted(range(n), key=lambda i: abs(i - n // 2))
    k = 0
    for t in order:
        done = list(reversed(order[:order.index(t)]))
        cand = [r for r in range(n) if r != t]
        cand = [r for _, r in sorted((abs(r - t), r) for r in cand)]
        cand = done + [r for r in cand if abs(r - t) <= D and abs(r - t) % s == 0 and r not in done]   # :484-503
        y_comp, per = None, 0.0
        while y_comp is None or (len(cand) > 0 and per > e):
            r = cand.pop(0)
            x_t, m_t = y_inp[None, :, t].copy(), m_inp[None, :, t].copy()
            x_al, v_al, v_map = dfpn_align_tail(y_inp[None, :, r:r + 1], m_inp[None, :, r:r + 1], m_t, flows[k % len(flows)])
            chn_pack(x_t, 1 - m_t, x_al, v_al, v_map)
            _, yc = chn_composite(nn_outs[k % len(nn_outs)], x_t, 1 - m_t, 1, 1)
            k += 1
            m_new, x_new, per = hole_update(m_t, v_map[:, :, 0], yc[:, :, 0])                           # :181-186
            m_inp[:, t], y_inp[:, t] = m_new[0], x_new[0]
            y_comp = yc[:, :, 0]
        m_inp[:, t] = 0
        y_inp[:, t] = y_comp[0]
    return y_inp, k
