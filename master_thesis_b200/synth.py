"""Deterministic synthetic inputs for the frame-alignment hot path.

Everything is drawn from ``numpy.random.RandomState`` (the legacy MT19937
stream, bit-stable across numpy versions and machines), so the build
container, the GPU box and the committed golden vectors all see identical
arrays.  Shapes and distributions follow SURVEY.md section 8(d).

All arrays are float32; masks are float32 {0, 1} as in the reference
(dataset.py:168-169 builds ``x = (1 - m) * y + m * fill``).
"""
import numpy as np

FILL = np.array([0.485, 0.456, 0.406], dtype=np.float32)  # dataset.py:36


def rng(seed):
    return np.random.RandomState(seed)


def identity_grid(h, w, align_corners=True):
    """Absolute identity flow in [-1, 1], (h, w, 2), last dim = (x, y).

    Same values as the reference's ``affine_grid`` identity (utils.py:27-31)
    up to 1 ulp; exactness is irrelevant here, it is only a generator.
    """
    if align_corners:
        xs = np.linspace(-1.0, 1.0, w)
        ys = np.linspace(-1.0, 1.0, h)
    else:
        xs = (2.0 * np.arange(w) + 1.0) / w - 1.0
        ys = (2.0 * np.arange(h) + 1.0) / h - 1.0
    g = np.stack(np.meshgrid(xs, ys), axis=-1)
    return g.astype(np.float32)


def frames(seed, b, f, h, w, hole=0.1, blobs=True):
    """Masked frames ``x`` (b,3,f,h,w) and masks ``m`` (b,1,f,h,w).

    ``blobs`` makes the holes rectangles (like the reference's object masks)
    with a sprinkle of isolated pixels so every mask edge case is hit.
    """
    r = rng(seed)
    y = r.random_sample((b, 3, f, h, w)).astype(np.float32)
    m = (r.random_sample((b, 1, f, h, w)) > (1.0 - hole * 0.2)).astype(np.float32)
    if blobs:
        for bi in range(b):
            for fi in range(f):
                hh = max(1, int(h * (0.2 + 0.2 * r.random_sample())))
                ww = max(1, int(w * (0.2 + 0.2 * r.random_sample())))
                y0 = r.randint(0, h - hh + 1)
                x0 = r.randint(0, w - ww + 1)
                m[bi, 0, fi, y0:y0 + hh, x0:x0 + ww] = 1.0
    else:
        m = (r.random_sample((b, 1, f, h, w)) > (1.0 - hole)).astype(np.float32)
    x = (1.0 - m) * y + m * FILL.reshape(1, 3, 1, 1, 1)
    return x.astype(np.float32), m, y


def dense_flow(seed, b, f, h, w, sigma=0.05, smooth=True):
    """Absolute flow (b,f,h,w,2): identity + noise.

    ``smooth`` draws the noise at 1/8 resolution and repeats it, which is
    what a predicted optical flow looks like (locally coherent); ``False``
    gives per-pixel white noise (worst case for gather locality).
    """
    r = rng(seed)
    if smooth:
        hs, ws = (h + 7) // 8, (w + 7) // 8
        n = r.standard_normal((b, f, hs, ws, 2)).astype(np.float32)
        n = np.repeat(np.repeat(n, 8, axis=2), 8, axis=3)[:, :, :h, :w]
        n = n + 0.02 * r.standard_normal((b, f, h, w, 2)).astype(np.float32)
    else:
        n = r.standard_normal((b, f, h, w, 2)).astype(np.float32)
    g = identity_grid(h, w, True)[None, None]
    return (g + np.float32(sigma) * n).astype(np.float32)


def tie_flow(seed, b, f, h, w):
    """Flow whose pixel coordinates sit on / next to k+0.5 and integer k.

    Exercises round-half-to-even of the nearest sampler and the zero padding
    borders (-0.5, size-0.5, < -1, > 1) - SURVEY 8(c) edge cases.
    """
    r = rng(seed)
    kx = r.randint(-2, w + 2, size=(b, f, h, w)).astype(np.float64)
    ky = r.randint(-2, h + 2, size=(b, f, h, w)).astype(np.float64)
    half = r.randint(0, 3, size=(b, f, h, w)) * 0.5  # 0, .5, 1.0
    gx = ((kx + half) / ((w - 1) / 2.0) - 1.0).astype(np.float32)
    gy = ((ky + half) / ((h - 1) / 2.0) - 1.0).astype(np.float32)
    for arr in (gx, gy):
        step = r.randint(-2, 3, size=arr.shape)
        for k in (1, 2):
            up = step >= k
            dn = step <= -k
            arr[up] = np.nextafter(arr[up], np.float32(np.inf))
            arr[dn] = np.nextafter(arr[dn], np.float32(-np.inf))
    return np.stack([gx, gy], axis=-1).astype(np.float32)


def thetas(seed, n, sigma=0.1):
    """Affine parameters (n,2,3) = identity + sigma * randn."""
    r = rng(seed)
    t = np.zeros((n, 2, 3), dtype=np.float32)
    t[:, 0, 0] = 1.0
    t[:, 1, 1] = 1.0
    return (t + np.float32(sigma) * r.standard_normal((n, 2, 3))).astype(np.float32)


def vgg_feats(seed, b, f, c=512, h=16, w=16, keep=0.9):
    """Post-ReLU feature maps and visibility maps for the correlation."""
    r = rng(seed)
    ft = np.maximum(r.standard_normal((b, c, h, w)), 0).astype(np.float32)
    fr = np.maximum(r.standard_normal((b, c, f, h, w)), 0).astype(np.float32)
    vt = (r.random_sample((b, 1, h, w)) < keep).astype(np.float32)
    vr = (r.random_sample((b, 1, f, h, w)) < keep).astype(np.float32)
    return ft, vt, fr, vr


def cm_inputs(seed, b, f, c=128, h=64, w=64, up=4, keep=0.8):
    """CM_Module inputs: c_feats (b,c,f,h,w), v_t (b,1,H,W), v_aligned (b,1,f-1,H,W)."""
    r = rng(seed)
    cf = r.standard_normal((b, c, f, h, w)).astype(np.float32)
    vt = (r.random_sample((b, 1, h * up, w * up)) < keep).astype(np.float32)
    va = (r.random_sample((b, 1, f - 1, h * up, w * up)) < keep).astype(np.float32)
    # coherent blobs so that the >0.5 threshold of the 4x down-sample is not
    # all-or-nothing noise
    for bi in range(b):
        hh, ww = (h * up) // 3, (w * up) // 3
        y0 = r.randint(0, h * up - hh + 1)
        x0 = r.randint(0, w * up - ww + 1)
        vt[bi, 0, y0:y0 + hh, x0:x0 + ww] = 0.0
        for fi in range(f - 1):
            y0 = r.randint(0, h * up - hh + 1)
            x0 = r.randint(0, w * up - ww + 1)
            va[bi, 0, fi, y0:y0 + hh, x0:x0 + ww] = 0.0
    return cf, vt, va


def nn_output(seed, n, h, w):
    """Stand-in for the hallucination CNN's output (n,3,h,w), normalised space."""
    r = rng(seed)
    return (2.5 * r.standard_normal((n, 3, h, w))).astype(np.float32)
