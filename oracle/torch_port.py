"""CPU baseline "port": the reference's call sequences for the hot path restated on
torch CPU ops - TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference cannot travel to the GPU box (/root/reference does not exist
there), so ``bench.py``'s ``cpu_baseline`` and ``--impl reference`` legs time
this module instead: it issues the SAME ATen calls in the same order as the
reference (F.grid_sample / F.affine_grid / F.interpolate / torch.matmul /
F.l1_loss ...), so it has the reference's CPU performance characteristics
(multi-threaded ATen, mkldnn/MKL).  tests/test_oracle_golden.py pins it
bit-exactly to the golden vectors produced by the unmodified reference.
Each function cites the reference lines it restates.
"""
import torch
import torch.nn.functional as F

MEAN = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1, 1)   # model_chn.py:32-37
STD = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1, 1)


def _frames(t):
    """(B,C,F,H,W) -> (B*F,C,H,W), the reference's transpose+reshape (utils.py:94)."""
    b, c, f, h, w = t.shape
    return t.transpose(1, 2).reshape(b * f, c, h, w)


def _clip(t, b):
    """(B*F,C,H,W) -> (B,C,F,H,W) view (utils.py:97)."""
    n, c, h, w = t.shape
    return t.reshape(b, n // b, c, h, w).transpose(1, 2)


def align_set(x, v, flow):
    """utils.py:78-104."""
    b, _, _, h, w = x.shape
    grid = flow.reshape(-1, h, w, 2)
    xa = F.grid_sample(_frames(x), grid, mode='bilinear', padding_mode='zeros', align_corners=True)
    va = F.grid_sample(_frames(v), grid, mode='nearest', padding_mode='zeros', align_corners=True)
    return _clip(xa, b), _clip(va, b)


def dfpn_align_tail(x_refs, m_refs, m_target, flow):
    """model_dfpn.py:128-133."""
    xa, va = align_set(x_refs, 1 - m_refs, flow)
    return xa, va, (va - (1 - m_target).unsqueeze(2)).clamp(0, 1)


def cpn_align_tail(x_refs, m_refs, m_target, theta):
    """model_cpn.py:75-89."""
    b, c, f, h, w = x_refs.shape
    grid = F.affine_grid(theta, [theta.size(0), c, h, w], align_corners=False)
    xa = _clip(F.grid_sample(_frames(x_refs), grid, align_corners=False), b)
    va = (_clip(F.grid_sample(1 - _frames(m_refs), grid, align_corners=False), b) > 0.5).float()
    return xa, va, (va - (1 - m_target.unsqueeze(2))).clamp(0, 1)


def mask_out(flow):
    """model_dfpn.py:269-272."""
    return ((flow < -1).float() + (flow > 1).float()).sum(4).clamp(0, 1).unsqueeze(1)


def masked_l1(y_hat, y, mask, batch_mask=None, reduction='mean', weight=1):
    """utils.py:139-169."""
    if batch_mask is not None:
        if not bool(batch_mask.any()):
            return torch.zeros(1)
        y_hat, y, mask = y_hat[batch_mask], y[batch_mask], mask[batch_mask]
    loss = F.l1_loss(y_hat * mask, y * mask, reduction=reduction)
    return weight * loss / (torch.sum(mask) + 1e-9 if reduction == 'sum' else 1)


def alignment_recons(x_target, v_target, x_refs, v_refs, flow):
    """model_dfpn.py:377-383 + 269-287: warp, mask_out, masked L1 ('sum')."""
    f = x_refs.size(2)
    xa, _ = align_set(x_refs, v_refs, flow)
    mask = v_target.unsqueeze(2).repeat(1, 1, f, 1, 1) * (1 - mask_out(flow))
    return masked_l1(x_target.unsqueeze(2).repeat(1, 1, f, 1, 1), xa, mask, reduction='sum')


def corr4d(ft, vt, fr, vr):
    """model_dfpn.py:534-565."""
    b, c, f, h, w = fr.shape
    if vt is not None:
        ft = ft * vt
    if vr is not None:
        fr = fr * vr
    a = ft.reshape(b, c, -1).transpose(-1, -2).unsqueeze(1)
    a_n = torch.norm(a, dim=3).unsqueeze(3) + 1e-9
    bm = fr.reshape(b, c, f, -1).permute(0, 2, 1, 3)
    bm_n = torch.norm(bm, dim=2).unsqueeze(2) + 1e-9
    return torch.matmul(a / a_n, bm / bm_n).reshape(b, f, h, w, h, w)


def masked_softmax(vec, mask, dim):
    """model_cpn.py:245-254."""
    mv = vec * mask
    e = torch.exp(mv - mv.max(dim=dim, keepdim=True)[0]) * mask
    s = e.sum(dim, keepdim=True)
    s = s + (s < 1e-4).float()
    return e / s


def cm_module(c_feats, v_t, v_aligned):
    """model_cpn.py:206-243."""
    b, c, f, h, w = c_feats.shape
    vt = (F.interpolate(v_t, size=(h, w), mode='bilinear', align_corners=False) > 0.5).float()
    sims, vrs = [], []
    for r in range(f - 1):
        vr = (F.interpolate(v_aligned[:, :, r], size=(h, w), mode='bilinear',
                            align_corners=False) > 0.5).float()
        vrs.append(vr)
        vmap = vt * vr
        v_sum = vmap[:, 0].sum(-1).sum(-1)
        zeros = v_sum < 1e-4
        v_sum = v_sum + zeros.float()
        gs = (vmap * c_feats[:, :, 0] * c_feats[:, :, r + 1]).sum(-1).sum(-1).sum(-1) / (v_sum * c)
        gs[zeros] = 0
        sims.append(torch.ones((b, c, h, w)) * gs.view(b, 1, 1, 1))
    sims, vrs = torch.stack(sims, dim=2), torch.stack(vrs, dim=2)
    match = masked_softmax(sims, vrs, dim=2)
    c_out = torch.sum(c_feats[:, :, 1:] * match, dim=2)
    c_mask = 1 - torch.mean(torch.sum(match * vrs, 2), 1, keepdim=True)
    return torch.cat([c_feats[:, :, 0], c_out, c_mask], dim=1), c_mask


def chn_pack(x_t, v_t, x_al, v_al, v_map):
    """model_chn.py:68-80."""
    b, c, f, h, w = x_al.shape
    xt = x_t.unsqueeze(2).repeat(1, 1, f, 1, 1)
    vt = v_t.unsqueeze(2).repeat(1, 1, f, 1, 1)
    nn_in = torch.cat([(xt - MEAN) / STD, (x_al - MEAN) / STD, vt, v_al, v_map], dim=1)
    return nn_in.transpose(1, 2).reshape(b * f, 9, h, w)


def flow_pack(x_target, m_target, x_refs, m_refs, flow_pre):
    """model_dfpn.py:733-741 (FlowEstimator.forward's nn_input)."""
    b, c, f, h, w = x_refs.shape
    return torch.cat([
        x_refs.transpose(1, 2).reshape(b * f, c, h, w),
        x_target.unsqueeze(1).repeat(1, f, 1, 1, 1).reshape(b * f, c, h, w),
        m_refs.transpose(1, 2).reshape(b * f, 1, h, w),
        m_target.unsqueeze(1).repeat(1, f, 1, 1, 1).reshape(b * f, 1, h, w),
        flow_pre.reshape(b * f, h, w, 2).permute(0, 3, 1, 2),
    ], dim=1)


def chn_composite(nn_out, x_t, v_t, b, f):
    """model_chn.py:80-85."""
    _, c, h, w = nn_out.shape
    out = nn_out.reshape(b, f, c, h, w).transpose(1, 2)
    xt = x_t.unsqueeze(2).repeat(1, 1, f, 1, 1)
    vt = v_t.unsqueeze(2).repeat(1, 1, f, 1, 1)
    y_hat = torch.clamp(out * STD + MEAN, 0, 1)
    return y_hat, vt * xt + (1 - vt) * y_hat


def hole_update(m_t, v_map0, y_comp0):
    """model_chn.py:128-131."""
    fill = MEAN.view(1, 3, 1, 1)
    m_new = m_t - v_map0
    x_new = (1 - m_new) * y_comp0 + m_new.repeat(1, 3, 1, 1) * fill
    return m_new, x_new, torch.sum(m_new) * 100 / m_new.numel()


# --------------------------------------------------------------------------------------------------
# DFPN training step around the hot path (model_dfpn.py:310-394, 210-293): needed to pin the patched
# DFPN._train_val_wrapper / DFPN.compute_loss on the GPU box, where the reference is absent.
# --------------------------------------------------------------------------------------------------
def resize_set(x, v, y, size):
    """TransformsUtils.resize_set, utils.py:522-560."""
    b, c, f, h, w = x.size()

    def rs(t, ch, **kw):
        return F.interpolate(t.transpose(1, 2).reshape(-1, ch, h, w), (size, size), **kw) \
            .reshape(b, f, ch, size, size).transpose(1, 2)
    return rs(x, c, mode='bilinear'), rs(v, 1), rs(y, c, mode='bilinear')


def resize_flow(flow, size, mode='nearest'):
    """FlowsUtils.resize_flow, utils.py:107-126."""
    b, f, h, w, _ = flow.size()
    r = F.interpolate(flow.reshape(b * f, h, w, 2).permute(0, 3, 1, 2), size, mode=mode)
    return r.reshape(b, f, 2, size[0], size[1]).permute(0, 1, 3, 4, 2)


def dfpn_train_val_wrapper(forward, x, m, y, flow_gt, flows_use, t, r_list):
    """model_dfpn.py:346-394 (the ground-truth alignments :358-375 feed nothing and are omitted)."""
    corr, flow_16, flow_64, flow_256 = forward(x[:, :, t], m[:, :, t], x[:, :, r_list], m[:, :, r_list])
    x_16, v_16, y_16 = resize_set(x, 1 - m, y, 16)
    x_64, v_64, y_64 = resize_set(x, 1 - m, y, 64)
    x_256, v_256, y_256 = x, 1 - m, y
    flows_gt = (resize_flow(flow_gt[:, r_list], (16, 16)), resize_flow(flow_gt[:, r_list], (64, 64)), flow_gt[:, r_list])
    xs_aligned = tuple(align_set(xx[:, :, r_list], vv[:, :, r_list], fl)[0]
                       for xx, vv, fl in ((x_16, v_16, flow_16), (x_64, v_64, flow_64), (x_256, v_256, flow_256)))
    return corr, (x_16, x_64, x_256), (v_16, v_64, v_256), (y_16, y_64, y_256), xs_aligned, \
        (flow_16, flow_64, flow_256), flows_gt, flows_use


def dfpn_compute_loss(model_vgg, corr, xs, vs, ys, xs_aligned, flows, flows_gt, flows_use, t, r_list):
    """model_dfpn.py:210-293."""
    b, c, f, h, w = ys[2].size()
    with torch.no_grad():
        y_in = ys[2].transpose(1, 2).reshape(b * f, c, h, w)
        if not (h == 256 and w == 256):
            y_in = F.interpolate(y_in, (256, 256), mode='bilinear')
        feats = model_vgg(y_in)
    feats = feats[3].reshape(b, f, -1, 16, 16).transpose(1, 2)
    corr_loss = F.l1_loss(corr, corr4d(feats[:, :, t], None, feats[:, :, r_list], None))
    items = [corr_loss] + [masked_l1(flows[i], flows_gt[i], torch.ones_like(flows[i]), flows_use) for i in range(3)]
    for i in (1, 2):
        n = len(r_list)
        items.append(masked_l1(xs[i][:, :, t].unsqueeze(2).repeat(1, 1, n, 1, 1), xs_aligned[i],
                               vs[i][:, :, t].unsqueeze(2).repeat(1, 1, n, 1, 1) * (1 - mask_out(flows[i])),
                               reduction='sum'))
    return sum(items[1:], items[0]), items
