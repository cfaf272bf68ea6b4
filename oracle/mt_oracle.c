/*
 * mt_oracle.c - CPU ORACLE for the frame-alignment / temporal-copying hot path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.  The
 * product (master_thesis_b200/, libmt_b200.so) never links or calls it.
 *
 * What it restates.  The reference (davidalvarezdlt/master_thesis) is pure
 * Python; the arithmetic of this path lives in a third-party dependency that
 * is NOT under /root/reference: torch (pinned torch==1.10.0 in
 * requirements.txt:54; this image ships torch 2.11.0+cu128).  The functions
 * below restate (i) the reference's own call sequences, cited file:line per
 * function, and (ii) the published ATen CPU algorithms those calls resolve to
 * (aten/src/ATen/native/cpu/GridSamplerKernel.cpp, GridSampler.h:27-36,
 * 205-243, AffineGridGenerator.cpp, UpSample.h:289-323), including the exact
 * operation order and FMA contraction of the CPU build, which were pinned by
 * probing torch 2.11 CPU in the build container:
 *     align_corners=True :  ix = (gx + 1) * ((W - 1) / 2)
 *     align_corners=False:  ix = fma(gx + 1, W / 2, -0.5)
 *     w = ix - floor(ix); e = 1 - w; n = iy - floor(iy); s = 1 - n
 *     out = fma(se_v, n*w, fma(sw_v, n*e, fma(ne_v, s*w, nw_v * (s*e))))
 *     nearest: rint() (half-to-even) of the same ix/iy, then the bounds test
 *     affine_grid: g = fma(base_y, th[1], base_x * th[0]) + th[2]
 * Parity pinning: tests/test_oracle_golden.py checks every function here
 * against golden vectors produced by the unmodified reference
 * (tests/golden/make_golden.py).  The reference has no tests of its own.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp).
 * -ffp-contract=off matters: every fused multiply-add below is an explicit
 * fmaf(); nothing else may be contracted.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define MTO_API __attribute__((visibility("default")))

MTO_API int mto_version(void) { return 1; }

MTO_API int mto_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

MTO_API void mto_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------------ */
/* Sampling primitives (ATen GridSamplerKernel.cpp, CPU vectorised path)     */
/* ------------------------------------------------------------------------ */

static inline float unnormalize(float g, int size, int align_corners) {
    /* ComputeLocationBase<float, align_corners>::unnormalize */
    if (align_corners) {
        float sf = (float)(size - 1) / 2.0f;
        return (g + 1.0f) * sf;
    } else {
        float sf = (float)size / 2.0f;
        return fmaf(g + 1.0f, sf, -0.5f);
    }
}

static inline float fetch(const float *plane, int h, int w, float fy, float fx) {
    /* zero padding: a corner contributes only if it lies inside the frame.
       The bounds test is done on the float so that NaN / +-inf / |v| >= 2^31
       are out of bounds, as cvttps2dq -> INT_MIN makes them on the CPU. */
    if (!(fx >= 0.0f && fx <= (float)(w - 1) && fy >= 0.0f && fy <= (float)(h - 1)))
        return 0.0f;
    return plane[(int64_t)(int)fy * w + (int)fx];
}

typedef struct {
    float xw, yn;       /* floor(ix), floor(iy) */
    float nw, ne, sw, se;
    float w, e, n, s;
} bil_t;

static inline bil_t bilinear_params(float ix, float iy) {
    bil_t p;
    p.xw = floorf(ix);
    p.yn = floorf(iy);
    p.w = ix - p.xw;
    p.e = 1.0f - p.w;
    p.n = iy - p.yn;
    p.s = 1.0f - p.n;
    p.nw = p.s * p.e;
    p.ne = p.s * p.w;
    p.sw = p.n * p.e;
    p.se = p.n * p.w;
    return p;
}

static inline float bilinear_sample(const float *plane, int h, int w, const bil_t *p) {
    float vnw = fetch(plane, h, w, p->yn, p->xw);
    float vne = fetch(plane, h, w, p->yn, p->xw + 1.0f);
    float vsw = fetch(plane, h, w, p->yn + 1.0f, p->xw);
    float vse = fetch(plane, h, w, p->yn + 1.0f, p->xw + 1.0f);
    return fmaf(vse, p->se, fmaf(vsw, p->sw, fmaf(vne, p->ne, vnw * p->nw)));
}

static inline float nearest_sample(const float *plane, int h, int w, float ix, float iy) {
    return fetch(plane, h, w, rintf(iy), rintf(ix));
}

/* F.grid_sample(input (n,c,h,w), grid (n,ho,wo,2), mode, zeros, align_corners)
   -> out (n,c,ho,wo).  mode: 0 = bilinear, 1 = nearest. */
MTO_API void mto_grid_sample(const float *in, const float *grid, int n, int c, int h, int w,
                             int ho, int wo, int mode, int align_corners, float *out) {
    int64_t po = (int64_t)ho * wo, pi = (int64_t)h * w;
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < n; ++i) {
        for (int64_t p = 0; p < po; ++p) {
            const float *g = grid + ((int64_t)i * po + p) * 2;
            float ix = unnormalize(g[0], w, align_corners);
            float iy = unnormalize(g[1], h, align_corners);
            if (mode == 0) {
                bil_t bp = bilinear_params(ix, iy);
                for (int k = 0; k < c; ++k)
                    out[((int64_t)i * c + k) * po + p] =
                        bilinear_sample(in + ((int64_t)i * c + k) * pi, h, w, &bp);
            } else {
                for (int k = 0; k < c; ++k)
                    out[((int64_t)i * c + k) * po + p] =
                        nearest_sample(in + ((int64_t)i * c + k) * pi, h, w, ix, iy);
            }
        }
    }
}

/* torch.linspace(-1, 1, steps) CPU scalar algorithm (RangeFactoriesKernel.cpp):
   step = (end - start) / (steps - 1); first half counts up from start, second
   half counts down from end.  (The vectorised ATen path differs from this in
   the last ulp depending on the host's SIMD width; see DESIGN.md.) */
static inline float linspace_m1_p1(int idx, int steps) {
    if (steps <= 1) return -1.0f;
    float step = (1.0f - (-1.0f)) / (float)(steps - 1);
    if (idx < steps / 2) return -1.0f + step * (float)idx;
    return 1.0f - step * (float)(steps - idx - 1);
}

static inline float base_coord(int idx, int size, int align_corners) {
    /* AffineGridGenerator.cpp linspace_from_neg_one: for align_corners=False
       the linspace is scaled by (size - 1) / size. */
    float v = linspace_m1_p1(idx, size);
    if (!align_corners) v = v * (float)(size - 1) / (float)size;
    return v;
}

static inline void affine_point(const float *th, float bx, float by, float *gx, float *gy) {
    /* base_grid (x, y, 1) @ theta^T, CPU sgemm order with FMA as pinned. */
    *gx = fmaf(by, th[1], bx * th[0]) + th[2];
    *gy = fmaf(by, th[4], bx * th[3]) + th[5];
}

/* F.affine_grid(theta (n,2,3), [n,*,h,w], align_corners) -> grid (n,h,w,2). */
MTO_API void mto_affine_grid(const float *theta, int n, int h, int w, int align_corners,
                             float *grid) {
    for (int i = 0; i < n; ++i)
        for (int y = 0; y < h; ++y)
            for (int x = 0; x < w; ++x) {
                float *g = grid + (((int64_t)i * h + y) * w + x) * 2;
                affine_point(theta + (int64_t)i * 6, base_coord(x, w, align_corners),
                             base_coord(y, h, align_corners), g, g + 1);
            }
}

/* ------------------------------------------------------------------------ */
/* a1: FlowsUtils.align_set                      master_thesis/utils.py:78-104 */
/* ------------------------------------------------------------------------ */
/* x (b,c,f,h,w), v (b,1,f,h,w), flow (b,f,h,w,2) absolute coords.
   x_al (b,c,f,h,w) = grid_sample(bilinear, zeros, align_corners=True)  :93-97
   v_al (b,1,f,h,w) = grid_sample(nearest,  zeros, align_corners=True)  :98-103 */
MTO_API void mto_align_set(const float *x, const float *v, const float *flow, int b, int c,
                           int f, int h, int w, float *x_al, float *v_al) {
    int64_t P = (int64_t)h * w;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        for (int fi = 0; fi < f; ++fi) {
            const float *g = flow + ((int64_t)bi * f + fi) * P * 2;
            const float *vp = v + ((int64_t)bi * f + fi) * P;
            for (int64_t p = 0; p < P; ++p) {
                float ix = unnormalize(g[2 * p], w, 1);
                float iy = unnormalize(g[2 * p + 1], h, 1);
                bil_t bp = bilinear_params(ix, iy);
                for (int k = 0; k < c; ++k) {
                    int64_t o = (((int64_t)bi * c + k) * f + fi) * P;
                    x_al[o + p] = bilinear_sample(x + o, h, w, &bp);
                }
                v_al[((int64_t)bi * f + fi) * P + p] = nearest_sample(vp, h, w, ix, iy);
            }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* a2: DFPN.align tail                     master_thesis/model_dfpn.py:128-133 */
/* ------------------------------------------------------------------------ */
/* v = 1 - m_refs (:129); a1; v_map = clamp(v_al - (1 - m_target)[:, :, None], 0, 1) (:131) */
static inline float clamp01(float a) { return fminf(fmaxf(a, 0.0f), 1.0f); }

MTO_API void mto_dfpn_align_tail(const float *x_refs, const float *m_refs, const float *m_target,
                                 const float *flow, int b, int c, int f, int h, int w,
                                 float *x_al, float *v_al, float *v_map) {
    int64_t P = (int64_t)h * w, nv = (int64_t)b * f * P;
    float *v = (float *)malloc(sizeof(float) * nv);
    for (int64_t i = 0; i < nv; ++i) v[i] = 1.0f - m_refs[i];
    mto_align_set(x_refs, v, flow, b, c, f, h, w, x_al, v_al);
    free(v);
    for (int bi = 0; bi < b; ++bi)
        for (int fi = 0; fi < f; ++fi)
            for (int64_t p = 0; p < P; ++p) {
                int64_t o = ((int64_t)bi * f + fi) * P + p;
                v_map[o] = clamp01(v_al[o] - (1.0f - m_target[(int64_t)bi * P + p]));
            }
}

/* ------------------------------------------------------------------------ */
/* a3: CPN.align tail                        master_thesis/model_cpn.py:75-89 */
/* ------------------------------------------------------------------------ */
/* grid = affine_grid(theta (b*f,2,3), align_corners=False)            :75-77
   x_al = grid_sample(x_refs, bilinear, zeros, align_corners=False)     :79-83
   v_al = (grid_sample(1 - m_refs, bilinear, ...) > 0.5).float()        :84-88
   v_maps = clamp(v_al - (1 - m_target[:, :, None]), 0, 1)              :89
   If `grid` is non-NULL it is used instead of theta (dense (b,f,h,w,2)). */
MTO_API void mto_cpn_align_tail(const float *x_refs, const float *m_refs, const float *m_target,
                                const float *theta, const float *grid, int b, int c, int f,
                                int h, int w, float *x_al, float *v_al, float *v_map) {
    int64_t P = (int64_t)h * w;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        for (int fi = 0; fi < f; ++fi) {
            int64_t n = (int64_t)bi * f + fi;
            const float *mp = m_refs + n * P;
            for (int y = 0; y < h; ++y)
                for (int xx = 0; xx < w; ++xx) {
                    int64_t p = (int64_t)y * w + xx;
                    float gx, gy;
                    if (grid) {
                        gx = grid[(n * P + p) * 2];
                        gy = grid[(n * P + p) * 2 + 1];
                    } else {
                        affine_point(theta + n * 6, base_coord(xx, w, 0), base_coord(y, h, 0),
                                     &gx, &gy);
                    }
                    float ix = unnormalize(gx, w, 0), iy = unnormalize(gy, h, 0);
                    bil_t bp = bilinear_params(ix, iy);
                    for (int k = 0; k < c; ++k) {
                        int64_t o = (((int64_t)bi * c + k) * f + fi) * P;
                        x_al[o + p] = bilinear_sample(x_refs + o, h, w, &bp);
                    }
                    /* bilinear sample of v = 1 - m, corner by corner */
                    float c4[4];
                    float cy[4] = {bp.yn, bp.yn, bp.yn + 1.0f, bp.yn + 1.0f};
                    float cx[4] = {bp.xw, bp.xw + 1.0f, bp.xw, bp.xw + 1.0f};
                    for (int q = 0; q < 4; ++q) {
                        int inb = (cx[q] >= 0.0f && cx[q] <= (float)(w - 1) && cy[q] >= 0.0f &&
                                   cy[q] <= (float)(h - 1));
                        c4[q] = inb ? 1.0f - mp[(int64_t)(int)cy[q] * w + (int)cx[q]] : 0.0f;
                    }
                    float vs = fmaf(c4[3], bp.se, fmaf(c4[2], bp.sw, fmaf(c4[1], bp.ne, c4[0] * bp.nw)));
                    float va = vs > 0.5f ? 1.0f : 0.0f;
                    v_al[n * P + p] = va;
                    v_map[n * P + p] = clamp01(va - (1.0f - m_target[(int64_t)bi * P + p]));
                }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* a4: mask_out                            master_thesis/model_dfpn.py:269-272 */
/* ------------------------------------------------------------------------ */
/* clamp(sum_k[(flow_k < -1) + (flow_k > 1)], 0, 1): flow (n,2) -> out (n) */
MTO_API void mto_mask_out(const float *flow, int64_t n, float *out) {
    for (int64_t i = 0; i < n; ++i) {
        float gx = flow[2 * i], gy = flow[2 * i + 1];
        float s = (float)(gx < -1.0f) + (float)(gx > 1.0f) + (float)(gy < -1.0f) + (float)(gy > 1.0f);
        out[i] = clamp01(s);
    }
}

/* ------------------------------------------------------------------------ */
/* a5: LossesUtils.masked_l1                    master_thesis/utils.py:139-169 */
/* ------------------------------------------------------------------------ */
/* y_hat, y: (b, c, inner); mask: (b, mask_c, inner), mask_c in {1, c}.
   batch_mask: b bytes or NULL (:158-165).  reduction: 0 = 'mean', 1 = 'sum'.
   Returns weight * l1(y_hat*mask, y*mask) / (sum(mask) + 1e-9 if 'sum' else 1)
   (:166-169); 0 if batch_mask selects nothing (:158-159).  NB the denominator
   counts the mask once, the numerator once per channel.
   If sums != NULL: sums[0] = sum |.|, sums[1] = sum(mask), sums[2] = numel. */
MTO_API float mto_masked_l1(const float *y_hat, const float *y, const float *mask, int b, int c,
                            int64_t inner, int mask_c, const uint8_t *batch_mask, int reduction,
                            float weight, double *sums) {
    double num = 0.0, den = 0.0, cnt = 0.0;
    int any = 0;
    for (int bi = 0; bi < b; ++bi) {
        if (batch_mask && !batch_mask[bi]) continue;
        any = 1;
        for (int k = 0; k < c; ++k) {
            const float *mp = mask + ((int64_t)bi * mask_c + (mask_c == 1 ? 0 : k)) * inner;
            const float *a = y_hat + ((int64_t)bi * c + k) * inner;
            const float *d = y + ((int64_t)bi * c + k) * inner;
            double acc = 0.0;
#pragma omp parallel for reduction(+ : acc) schedule(static)
            for (int64_t i = 0; i < inner; ++i) acc += (double)fabsf(a[i] * mp[i] - d[i] * mp[i]);
            num += acc;
        }
        for (int k = 0; k < mask_c; ++k) {
            const float *mp = mask + ((int64_t)bi * mask_c + k) * inner;
            double acc = 0.0;
            for (int64_t i = 0; i < inner; ++i) acc += (double)mp[i];
            den += acc;
        }
        cnt += (double)c * (double)inner;
    }
    if (sums) { sums[0] = num; sums[1] = den; sums[2] = cnt; }
    if (batch_mask && !any) return 0.0f;
    if (reduction == 1) return weight * (float)num / ((float)den + 1e-9f);
    return weight * (float)(num / cnt);
}

/* Backward of a5 w.r.t. y (the second argument), grad_out = upstream scalar.
   d/dy |yh*m - y*m| = -sign(yh*m - y*m) * m, scaled by weight / (sum(mask)+1e-9)
   ('sum') or weight / numel ('mean').  grad_yhat is the negation.  (autograd of
   F.l1_loss and the two mask multiplications, utils.py:166-169.) */
MTO_API void mto_masked_l1_bwd(const float *y_hat, const float *y, const float *mask, int b, int c,
                               int64_t inner, int mask_c, const uint8_t *batch_mask,
                               int reduction, float weight, float grad_out, float *grad_y) {
    double sums[3];
    (void)mto_masked_l1(y_hat, y, mask, b, c, inner, mask_c, batch_mask, reduction, weight, sums);
    float scale = reduction == 1 ? weight / ((float)sums[1] + 1e-9f) : weight / (float)sums[2];
    scale *= grad_out;
    for (int bi = 0; bi < b; ++bi)
        for (int k = 0; k < c; ++k) {
            const float *mp = mask + ((int64_t)bi * mask_c + (mask_c == 1 ? 0 : k)) * inner;
            int64_t o = ((int64_t)bi * c + k) * inner;
            for (int64_t i = 0; i < inner; ++i) {
                if (batch_mask && !batch_mask[bi]) { grad_y[o + i] = 0.0f; continue; }
                float d = y_hat[o + i] * mp[i] - y[o + i] * mp[i];
                float sg = (d > 0.0f) - (d < 0.0f);
                grad_y[o + i] = -sg * mp[i] * scale;
            }
        }
}

/* ------------------------------------------------------------------------ */
/* a6: backward of a1 w.r.t. the flow (autograd of F.grid_sample, bilinear,    */
/*     zeros, align_corners; ATen GridSamplerKernel.cpp backward, GridSampler.h */
/*     :43-56 for the (size-1)/2 multiplier).  No gradient to x or v; nearest  */
/*     has zero grid gradient.                                                */
/* ------------------------------------------------------------------------ */
/* x (b,c,f,h,w), flow (b,f,h,w,2), gout (b,c,f,h,w) -> gflow (b,f,h,w,2) */
MTO_API void mto_align_set_bwd_flow(const float *x, const float *flow, const float *gout, int b,
                                    int c, int f, int h, int w, int align_corners, float *gflow) {
    int64_t P = (int64_t)h * w;
    float mx = align_corners ? (float)(w - 1) / 2.0f : (float)w / 2.0f;
    float my = align_corners ? (float)(h - 1) / 2.0f : (float)h / 2.0f;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        for (int fi = 0; fi < f; ++fi) {
            const float *g = flow + ((int64_t)bi * f + fi) * P * 2;
            float *gg = gflow + ((int64_t)bi * f + fi) * P * 2;
            for (int64_t p = 0; p < P; ++p) {
                float ix = unnormalize(g[2 * p], w, align_corners);
                float iy = unnormalize(g[2 * p + 1], h, align_corners);
                bil_t bp = bilinear_params(ix, iy);
                float gx = 0.0f, gy = 0.0f;
                for (int k = 0; k < c; ++k) {
                    int64_t o = (((int64_t)bi * c + k) * f + fi) * P;
                    const float *pl = x + o;
                    float vnw = fetch(pl, h, w, bp.yn, bp.xw);
                    float vne = fetch(pl, h, w, bp.yn, bp.xw + 1.0f);
                    float vsw = fetch(pl, h, w, bp.yn + 1.0f, bp.xw);
                    float vse = fetch(pl, h, w, bp.yn + 1.0f, bp.xw + 1.0f);
                    float go = gout[o + p];
                    gx += ((vne - vnw) * bp.s + (vse - vsw) * bp.n) * go;
                    gy += ((vsw - vnw) * bp.e + (vse - vne) * bp.w) * go;
                }
                gg[2 * p] = gx * mx;
                gg[2 * p + 1] = gy * my;
            }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* a7: CorrelationVGG.correlation_masked_4d  master_thesis/model_dfpn.py:534-565 */
/* ------------------------------------------------------------------------ */
/* ft (b,c,P), vt (b,P) or NULL, fr (b,c,f,P), vr (b,f,P) or NULL, P = h*w.
   A = ft*vt as (b,P,c) rows; Bm = fr*vr as (b,f,c,P) cols          :551-556
   A /= (||A||_2 over c + 1e-9)   :558-560;  Bm /= (||Bm||_2 + 1e-9) :561-562
   out (b,f,P_t,P_r) = A @ Bm                                        :564-565 */
MTO_API void mto_corr4d(const float *ft, const float *vt, const float *fr, const float *vr, int b,
                        int c, int f, int P, float *out) {
    float *an = (float *)malloc(sizeof(float) * (size_t)c * P);
    float *bn = (float *)malloc(sizeof(float) * (size_t)c * P);
    for (int bi = 0; bi < b; ++bi) {
        for (int p = 0; p < P; ++p) {
            double ss = 0.0;
            float m = vt ? vt[(int64_t)bi * P + p] : 1.0f;
            for (int k = 0; k < c; ++k) {
                float a = ft[((int64_t)bi * c + k) * P + p] * m;
                ss += (double)a * a;
            }
            float nrm = (float)sqrt(ss) + 1e-9f;
            for (int k = 0; k < c; ++k)
                an[(int64_t)k * P + p] = (ft[((int64_t)bi * c + k) * P + p] * m) / nrm;
        }
        for (int fi = 0; fi < f; ++fi) {
            for (int p = 0; p < P; ++p) {
                double ss = 0.0;
                float m = vr ? vr[((int64_t)bi * f + fi) * P + p] : 1.0f;
                for (int k = 0; k < c; ++k) {
                    float a = fr[(((int64_t)bi * c + k) * f + fi) * P + p] * m;
                    ss += (double)a * a;
                }
                float nrm = (float)sqrt(ss) + 1e-9f;
                for (int k = 0; k < c; ++k)
                    bn[(int64_t)k * P + p] = (fr[(((int64_t)bi * c + k) * f + fi) * P + p] * m) / nrm;
            }
            float *o = out + ((int64_t)bi * f + fi) * P * P;
#pragma omp parallel for schedule(static)
            for (int pt = 0; pt < P; ++pt) {
                double *acc = (double *)calloc((size_t)P, sizeof(double));
                for (int k = 0; k < c; ++k) {
                    double a = an[(int64_t)k * P + pt];
                    const float *br = bn + (int64_t)k * P;
                    for (int pr = 0; pr < P; ++pr) acc[pr] += a * (double)br[pr];
                }
                for (int pr = 0; pr < P; ++pr) o[(int64_t)pt * P + pr] = (float)acc[pr];
                free(acc);
            }
        }
    }
    free(an);
    free(bn);
}

/* ------------------------------------------------------------------------ */
/* a8: CM_Module.forward + masked_softmax    master_thesis/model_cpn.py:206-254 */
/* ------------------------------------------------------------------------ */
/* F.interpolate(v (H,W) -> (h,w), bilinear, align_corners=False) > 0.5
   (UpSample.h area_pixel_compute_source_index: src = scale*(dst+0.5)-0.5,
   clamped at 0; i1 = i0 + (i0 < in-1); l1 = src - i0).                       */
static void resize_gt_half(const float *v, int H, int W, int h, int w, float *out) {
    float sy = (float)H / (float)h, sx = (float)W / (float)w;
    for (int y = 0; y < h; ++y) {
        float fy = sy * ((float)y + 0.5f) - 0.5f;
        if (fy < 0.0f) fy = 0.0f;
        int y0 = (int)fy, y1 = y0 + (y0 < H - 1);
        float ly1 = fy - (float)y0, ly0 = 1.0f - ly1;
        for (int x = 0; x < w; ++x) {
            float fx = sx * ((float)x + 0.5f) - 0.5f;
            if (fx < 0.0f) fx = 0.0f;
            int x0 = (int)fx, x1 = x0 + (x0 < W - 1);
            float lx1 = fx - (float)x0, lx0 = 1.0f - lx1;
            float val = ly0 * (lx0 * v[(int64_t)y0 * W + x0] + lx1 * v[(int64_t)y0 * W + x1]) +
                        ly1 * (lx0 * v[(int64_t)y1 * W + x0] + lx1 * v[(int64_t)y1 * W + x1]);
            out[(int64_t)y * w + x] = val > 0.5f ? 1.0f : 0.0f;
        }
    }
}

/* c_feats (b,c,f,h,w) (index 0 of f = target), v_t (b,1,H,W), v_aligned
   (b,1,f-1,H,W) -> out (b,2c+1,h,w) = cat[c_t, c_out, c_mask] and c_mask
   (b,1,h,w); optionally gs_out (b,f-1) = the per-reference similarities. */
MTO_API void mto_cm_module(const float *c_feats, const float *v_t, const float *v_aligned, int b,
                           int c, int f, int h, int w, int H, int W, float *out, float *c_mask,
                           float *gs_out) {
    int R = f - 1;
    int64_t P = (int64_t)h * w, PP = (int64_t)H * W;
#pragma omp parallel for schedule(static)
    for (int bi = 0; bi < b; ++bi) {
        float *vt = (float *)malloc(sizeof(float) * P);
        float *vr = (float *)malloc(sizeof(float) * P * R);
        float *gs = (float *)malloc(sizeof(float) * R);
        resize_gt_half(v_t + (int64_t)bi * PP, H, W, h, w, vt);                    /* :208-210 */
        for (int r = 0; r < R; ++r) {
            resize_gt_half(v_aligned + ((int64_t)bi * R + r) * PP, H, W, h, w, vr + r * P); /* :214-217 */
            double v_sum = 0.0;
            for (int64_t p = 0; p < P; ++p) v_sum += (double)(vt[p] * vr[r * P + p]);   /* :220-221 */
            int zero = v_sum < 1e-4;                                               /* :222 */
            float v_sum_f = (float)v_sum + (zero ? 1.0f : 0.0f);                    /* :223 */
            double acc = 0.0;
            for (int k = 0; k < c; ++k) {
                const float *ct = c_feats + (((int64_t)bi * c + k) * f + 0) * P;
                const float *cr = c_feats + (((int64_t)bi * c + k) * f + r + 1) * P;
                for (int64_t p = 0; p < P; ++p)
                    acc += (double)(vt[p] * vr[r * P + p] * ct[p] * cr[p]);           /* :226-227 */
            }
            float g = (float)acc / (v_sum_f * (float)c);                            /* :225-227 */
            if (zero) g = 0.0f;                                                    /* :228 */
            gs[r] = g;
            if (gs_out) gs_out[(int64_t)bi * R + r] = g;
        }
        float *o = out + (int64_t)bi * (2 * c + 1) * P;
        for (int k = 0; k < c; ++k)
            memcpy(o + (int64_t)k * P, c_feats + (((int64_t)bi * c + k) * f) * P, sizeof(float) * P);
        for (int64_t p = 0; p < P; ++p) {
            /* masked_softmax over refs, :245-254 */
            float wgt[64];
            float mx = -INFINITY;
            for (int r = 0; r < R; ++r) {
                float mv = gs[r] * vr[r * P + p];
                if (mv > mx) mx = mv;
            }
            float s = 0.0f;
            for (int r = 0; r < R; ++r) {
                float mv = gs[r] * vr[r * P + p];
                wgt[r] = expf(mv - mx) * vr[r * P + p];
                s += wgt[r];
            }
            if (s < 1e-4f) s += 1.0f;
            float cm = 0.0f;
            for (int r = 0; r < R; ++r) {
                wgt[r] = wgt[r] / s;
                cm += wgt[r] * vr[r * P + p];                                      /* :240 */
            }
            for (int k = 0; k < c; ++k) {                                          /* :238 */
                float acc = 0.0f;
                for (int r = 0; r < R; ++r)
                    acc += c_feats[(((int64_t)bi * c + k) * f + r + 1) * P + p] * wgt[r];
                o[((int64_t)c + k) * P + p] = acc;
            }
            /* mean over C of identical values = the value; 1 - mean  :241 */
            float cmv = 1.0f - cm;
            o[((int64_t)2 * c) * P + p] = cmv;
            c_mask[(int64_t)bi * P + p] = cmv;
        }
        free(vt); free(vr); free(gs);
    }
}

/* ------------------------------------------------------------------------ */
/* a9: CHN.forward pack                      master_thesis/model_chn.py:68-80 */
/* ------------------------------------------------------------------------ */
static const float MT_MEAN[3] = {0.485f, 0.456f, 0.406f}; /* model_chn.py:32-37 */
static const float MT_STD[3] = {0.229f, 0.224f, 0.225f};

/* x_t (b,3,P), v_t (b,1,P), x_ref_al (b,3,f,P), v_ref_al (b,1,f,P), v_map
   (b,1,f,P) -> nn_input (b*f, 9, P) = [x_t_norm(3), x_ref_norm(3), v_t, v_ref_al, v_map] */
MTO_API void mto_chn_pack(const float *x_t, const float *v_t, const float *x_al, const float *v_al,
                          const float *v_map, int b, int f, int64_t P, float *nn_in) {
    for (int bi = 0; bi < b; ++bi)
        for (int fi = 0; fi < f; ++fi) {
            float *o = nn_in + ((int64_t)bi * f + fi) * 9 * P;
            for (int k = 0; k < 3; ++k)
                for (int64_t p = 0; p < P; ++p) {
                    o[(int64_t)k * P + p] = (x_t[((int64_t)bi * 3 + k) * P + p] - MT_MEAN[k]) / MT_STD[k];
                    o[(int64_t)(3 + k) * P + p] =
                        (x_al[(((int64_t)bi * 3 + k) * f + fi) * P + p] - MT_MEAN[k]) / MT_STD[k];
                }
            for (int64_t p = 0; p < P; ++p) {
                o[6 * P + p] = v_t[(int64_t)bi * P + p];
                o[7 * P + p] = v_al[((int64_t)bi * f + fi) * P + p];
                o[8 * P + p] = v_map[((int64_t)bi * f + fi) * P + p];
            }
        }
}

/* ------------------------------------------------------------------------ */
/* 8f-4: FlowEstimator.forward input pack  master_thesis/model_dfpn.py:733-741 */
/* ------------------------------------------------------------------------ */
/* x_refs (b,3,f,P), x_t (b,3,P), m_refs (b,1,f,P), m_t (b,1,P), flow_pre (b,f,P,2) ->
   nn_input (b*f, 10, P) = cat[x_refs (:734), x_t repeated over f (:735-736), m_refs (:737),
   m_t repeated (:738-739), flow_pre as (2, P) planes (:740)] */
MTO_API void mto_flow_pack(const float *x_refs, const float *x_t, const float *m_refs, const float *m_t,
                           const float *flow, int b, int f, int64_t P, float *nn_in) {
    for (int bi = 0; bi < b; ++bi)
        for (int fi = 0; fi < f; ++fi) {
            float *o = nn_in + ((int64_t)bi * f + fi) * 10 * P;
            for (int k = 0; k < 3; ++k)
                for (int64_t p = 0; p < P; ++p) {
                    o[(int64_t)k * P + p] = x_refs[(((int64_t)bi * 3 + k) * f + fi) * P + p];
                    o[(int64_t)(3 + k) * P + p] = x_t[((int64_t)bi * 3 + k) * P + p];
                }
            for (int64_t p = 0; p < P; ++p) {
                o[6 * P + p] = m_refs[((int64_t)bi * f + fi) * P + p];
                o[7 * P + p] = m_t[(int64_t)bi * P + p];
                o[8 * P + p] = flow[(((int64_t)bi * f + fi) * P + p) * 2];
                o[9 * P + p] = flow[(((int64_t)bi * f + fi) * P + p) * 2 + 1];
            }
        }
}

/* ------------------------------------------------------------------------ */
/* a10: CHN.forward composite                master_thesis/model_chn.py:80-85 */
/* ------------------------------------------------------------------------ */
/* nn_out (b*f,3,P) -> y_hat (b,3,f,P) = clamp(nn_out*std + mean, 0, 1)   :83
                       y_hat_comp     = v_t*x_t + (1 - v_t)*y_hat        :84 */
MTO_API void mto_chn_composite(const float *nn_out, const float *x_t, const float *v_t, int b,
                               int f, int64_t P, float *y_hat, float *y_comp) {
    for (int bi = 0; bi < b; ++bi)
        for (int fi = 0; fi < f; ++fi)
            for (int k = 0; k < 3; ++k)
                for (int64_t p = 0; p < P; ++p) {
                    float o = nn_out[(((int64_t)bi * f + fi) * 3 + k) * P + p];
                    float yh = clamp01(o * MT_STD[k] + MT_MEAN[k]);
                    float vt = v_t[(int64_t)bi * P + p];
                    float xt = x_t[((int64_t)bi * 3 + k) * P + p];
                    int64_t d = (((int64_t)bi * 3 + k) * f + fi) * P + p;
                    y_hat[d] = yh;
                    y_comp[d] = vt * xt + (1.0f - vt) * yh;
                }
}

/* backward of a10 w.r.t. nn_out given grads of both outputs (b,3,f,P):
   g = (g_yhat + (1 - v_t) * g_comp) * 1[0 <= pre <= 1] * std   (torch.clamp
   passes the gradient where min <= x <= max). */
MTO_API void mto_chn_composite_bwd(const float *nn_out, const float *v_t, const float *g_yhat,
                                   const float *g_comp, int b, int f, int64_t P, float *g_nn) {
    for (int bi = 0; bi < b; ++bi)
        for (int fi = 0; fi < f; ++fi)
            for (int k = 0; k < 3; ++k)
                for (int64_t p = 0; p < P; ++p) {
                    int64_t s = (((int64_t)bi * f + fi) * 3 + k) * P + p;
                    int64_t d = (((int64_t)bi * 3 + k) * f + fi) * P + p;
                    float pre = nn_out[s] * MT_STD[k] + MT_MEAN[k];
                    float vt = v_t[(int64_t)bi * P + p];
                    float g = (g_yhat ? g_yhat[d] : 0.0f) + (g_comp ? (1.0f - vt) * g_comp[d] : 0.0f);
                    g_nn[s] = (pre >= 0.0f && pre <= 1.0f) ? g * MT_STD[k] : 0.0f;
                }
}

/* ------------------------------------------------------------------------ */
/* a11: hole update in CHN.inpaint_*   master_thesis/model_chn.py:128-131,     */
/*      181-186, 242-248                                                      */
/* ------------------------------------------------------------------------ */
/* m_t <- m_t - v_map[:, :, 0]; x_t <- (1 - m_t)*y_comp[:, :, 0] + m_t*fill;
   returns inp_per = sum(m_t) * 100 / numel.  m_t (b,1,P), v_map0 (b,1,P),
   y_comp0 (b,3,P) -> m_new (b,1,P), x_new (b,3,P). */
MTO_API float mto_hole_update(const float *m_t, const float *v_map0, const float *y_comp0, int b,
                              int64_t P, float *m_new, float *x_new) {
    double s = 0.0;
    for (int bi = 0; bi < b; ++bi)
        for (int64_t p = 0; p < P; ++p) {
            float m = m_t[(int64_t)bi * P + p] - v_map0[(int64_t)bi * P + p];
            m_new[(int64_t)bi * P + p] = m;
            s += (double)m;
            for (int k = 0; k < 3; ++k) {
                int64_t o = ((int64_t)bi * 3 + k) * P + p;
                x_new[o] = (1.0f - m) * y_comp0[o] + m * MT_MEAN[k];
            }
        }
    return (float)s * 100.0f / (float)((double)b * (double)P);
}

/* ------------------------------------------------------------------------ */
/* a12: trivial copy                      master_thesis/model_dfpn.py:427-429 */
/* ------------------------------------------------------------------------ */
/* y = x_t[:, :, None] * (1 - v_map) + x_ref_aligned * v_map, (b,3,f,P) */
MTO_API void mto_trivial_copy(const float *x_t, const float *x_al, const float *v_map, int b, int f,
                              int64_t P, float *y) {
    for (int bi = 0; bi < b; ++bi)
        for (int k = 0; k < 3; ++k)
            for (int fi = 0; fi < f; ++fi)
                for (int64_t p = 0; p < P; ++p) {
                    int64_t d = (((int64_t)bi * 3 + k) * f + fi) * P + p;
                    float vm = v_map[((int64_t)bi * f + fi) * P + p];
                    y[d] = x_t[((int64_t)bi * 3 + k) * P + p] * (1.0f - vm) + x_al[d] * vm;
                }
}

/* ------------------------------------------------------------------------ */
/* f1: FlowsUtils.resize_flow(flow, (H, W), mode='bilinear')                  */
/*     master_thesis/utils.py:107-126, called at model_dfpn.py:100-101        */
/* ------------------------------------------------------------------------ */
/* F.interpolate(bilinear, align_corners=False) of the two flow components
   (ATen UpSample.h:442-476 compute_source_index_and_lambda +
   UpSampleKernel.cpp cpu_upsample_generic).  Operation order and FMA
   contraction of the CPU build, pinned by probing torch 2.11 CPU with a
   256 x 256 source (the only source size the reference uses here) and
   outputs from 100 x 180 to 1080 x 1920:
       src = fma(scale, dst + 0.5, -0.5), clamped at 0; scale = in / out (fp32)
       i0 = min(floor(src), in - 1); i1 = i0 + (i0 < in - 1)
       l1 = clamp(src - i0, 0, 1); l0 = 1 - l1      (in == out: identity)
       row(y) = fma(v[y][x0], lx0, v[y][x1] * lx1)
       out    = fma(row(y0), ly0, row(y1) * ly1)
   (outputs of a few dozen pixels take another ATen code path whose last ulp
   differs: <= 2.4e-7.)                                                       */
typedef struct { int i0, i1; float l0, l1; } mto_lin_t;

static inline mto_lin_t lin_index(int dst, int in, int out) {
    mto_lin_t r;
    if (in == out) {
        r.i0 = r.i1 = dst; r.l0 = 1.0f; r.l1 = 0.0f;
        return r;
    }
    const float scale = (float)in / (float)out;
    float s = fmaf(scale, (float)dst + 0.5f, -0.5f);
    if (s < 0.0f) s = 0.0f;
    int i0 = (int)floorf(s);
    if (i0 > in - 1) i0 = in - 1;
    r.i0 = i0;
    r.i1 = i0 + (i0 < in - 1 ? 1 : 0);
    float l1 = s - (float)i0;
    l1 = l1 < 0.0f ? 0.0f : (l1 > 1.0f ? 1.0f : l1);
    r.l1 = l1;
    r.l0 = 1.0f - l1;
    return r;
}

/* flow (n, h, w, 2) -> out (n, H, W, 2) */
MTO_API void mto_resize_flow(const float *flow, int n, int h, int w, int H, int W, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int i = 0; i < n; ++i) {
        for (int y = 0; y < H; ++y) {
            const mto_lin_t ly = lin_index(y, h, H);
            const float *r0 = flow + ((int64_t)i * h + ly.i0) * w * 2;
            const float *r1 = flow + ((int64_t)i * h + ly.i1) * w * 2;
            float *o = out + ((int64_t)i * H + y) * W * 2;
            for (int x = 0; x < W; ++x) {
                const mto_lin_t lx = lin_index(x, w, W);
                for (int k = 0; k < 2; ++k) {
                    const float top = fmaf(r0[lx.i0 * 2 + k], lx.l0, r0[lx.i1 * 2 + k] * lx.l1);
                    const float bot = fmaf(r1[lx.i0 * 2 + k], lx.l0, r1[lx.i1 * 2 + k] * lx.l1);
                    o[x * 2 + k] = fmaf(top, ly.l0, bot * ly.l1);
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------ */
/* f3: nearest down-sample of the visibility maps in CorrelationVGG.forward  */
/*     master_thesis/model_dfpn.py:521-526:  v = F.interpolate(1 - m, (h, w),  */
/*     mode='nearest'); ATen UpSample.h nearest_neighbor_compute_source_index: */
/*     src = min(floor(dst * (in / out)), in - 1), scale in fp32.              */
/* ------------------------------------------------------------------------ */
/* m (n, H, W) -> v (n, h, w) = 1 - m[nearest] */
MTO_API void mto_vis_nearest(const float *m, int n, int H, int W, int h, int w, float *v) {
    const float sy = (float)H / (float)h, sx = (float)W / (float)w;
    for (int i = 0; i < n; ++i)
        for (int y = 0; y < h; ++y) {
            int ys = (int)floorf((float)y * sy);
            if (ys > H - 1) ys = H - 1;
            for (int x = 0; x < w; ++x) {
                int xs = (int)floorf((float)x * sx);
                if (xs > W - 1) xs = W - 1;
                v[((int64_t)i * h + y) * w + x] = 1.0f - m[((int64_t)i * H + ys) * W + xs];
            }
        }
}
