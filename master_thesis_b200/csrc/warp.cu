// warp.cu - K1: flow-guided warp + visibility + v_map, its backward w.r.t. the
// grid, and the fused warp + mask_out + masked-L1 forward / backward.
//
// Replaces (reference file:line):
//   FlowsUtils.align_set            master_thesis/utils.py:78-104        (a1)
//   DFPN.align tail                 master_thesis/model_dfpn.py:128-133  (a2)
//   CPN.align tail                  master_thesis/model_cpn.py:75-89     (a3)
//   mask_out                        master_thesis/model_dfpn.py:269-272  (a4)
//   masked_l1 'sum' on the warp     master_thesis/model_dfpn.py:274-287  (a5)
//   autograd of the above           (a6)
//
// HBM-bound gather kernels.  Layout: every tensor is a set of contiguous
// (H, W) planes.  Lanes of a warp own CONSECUTIVE pixels of the flattened plane,
// so a gather instruction of a coherent flow touches 1-2 cache lines (a
// 4-pixels-per-thread mapping was measured at 30% of HBM peak: every gather
// spanned 4 lines, see profiles/).  Each thread processes U = 4 pixels 256 apart
// for memory-level parallelism: all 4 * (4C + 1..4) gathers are issued before
// the first use.  Grid loads (8 B/lane) and all stores (4 B/lane) are fully
// coalesced 128 B-per-warp streaming accesses.
#include <math.h>

#include "mt_common.cuh"

namespace mt {
namespace {

struct Sampler {
    float sfx, sfy;      // (W-1)/2 | W/2, (H-1)/2 | H/2   (host-computed in fp32)
    float wmax, hmax;
    float stepx, stepy;  // 2/(W-1), 2/(H-1): torch.linspace step for affine grids
    int W, H;
    bool ac;
};

// ATen CPU ComputeLocationBase::unnormalize (pinned operation order: DESIGN.md "bit-exactness")
__device__ __forceinline__ float unnormalize(float g, float sf, bool ac) {
    const float t = __fadd_rn(g, 1.0f);
    return ac ? __fmul_rn(t, sf) : __fmaf_rn(t, sf, -0.5f);
}

struct Bil {
    float xw, yn, w, e, n, s, nw, ne, sw, se;
    bool x0, x1, y0, y1;  // corner column / row inside the frame
    int o00;              // offset of the (yn, xw) corner (valid only if y0 && x0 ...)
};

__device__ __forceinline__ Bil bil_params(float ix, float iy, const Sampler &sp) {
    Bil b;
    b.xw = floorf(ix);
    b.yn = floorf(iy);
    b.w = __fsub_rn(ix, b.xw);
    b.e = __fsub_rn(1.0f, b.w);
    b.n = __fsub_rn(iy, b.yn);
    b.s = __fsub_rn(1.0f, b.n);
    b.nw = __fmul_rn(b.s, b.e);
    b.ne = __fmul_rn(b.s, b.w);
    b.sw = __fmul_rn(b.n, b.e);
    b.se = __fmul_rn(b.n, b.w);
    const float xe = b.xw + 1.0f, ys = b.yn + 1.0f;
    // float-domain bounds tests: NaN / inf / |v| >= 2^31 are out of bounds
    b.x0 = (b.xw >= 0.0f) && (b.xw <= sp.wmax);
    b.x1 = (xe >= 0.0f) && (xe <= sp.wmax);
    b.y0 = (b.yn >= 0.0f) && (b.yn <= sp.hmax);
    b.y1 = (ys >= 0.0f) && (ys <= sp.hmax);
    // any in-bounds corner implies |xw|,|yn| small: the int conversion is exact
    const bool any = (b.x0 || b.x1) && (b.y0 || b.y1);
    b.o00 = any ? (int)b.yn * sp.W + (int)b.xw : 0;
    return b;
}

struct Corners {
    float nw, ne, sw, se;
};

__device__ __forceinline__ Corners gather(const float *__restrict__ plane, const Bil &b, int W) {
    Corners c;
    c.nw = (b.y0 && b.x0) ? __ldg(plane + b.o00) : 0.0f;
    c.ne = (b.y0 && b.x1) ? __ldg(plane + b.o00 + 1) : 0.0f;
    c.sw = (b.y1 && b.x0) ? __ldg(plane + b.o00 + W) : 0.0f;
    c.se = (b.y1 && b.x1) ? __ldg(plane + b.o00 + W + 1) : 0.0f;
    return c;
}

__device__ __forceinline__ float interp(const Corners &c, const Bil &b) {
    // fma(se_v, se, fma(sw_v, sw, fma(ne_v, ne, nw_v * nw)))  (pinned order)
    return __fmaf_rn(c.se, b.se, __fmaf_rn(c.sw, b.sw, __fmaf_rn(c.ne, b.ne, __fmul_rn(c.nw, b.nw))));
}

__device__ __forceinline__ float nearest(const float *__restrict__ plane, float ix, float iy,
                                         const Sampler &sp, bool from_mask) {
    const float xr = rintf(ix), yr = rintf(iy);  // half-to-even, like _mm256_round_ps
    const bool in = (xr >= 0.0f) && (xr <= sp.wmax) && (yr >= 0.0f) && (yr <= sp.hmax);
    if (!in) return 0.0f;
    const float v = __ldg(plane + (int)yr * sp.W + (int)xr);
    return from_mask ? __fsub_rn(1.0f, v) : v;
}

// torch.linspace(-1, 1, n)[i] (scalar CPU algorithm), scaled for align_corners=False
// step = 2 / (size - 1) in fp32, computed once on the host (same IEEE division)
__device__ __forceinline__ float base_coord(int idx, int size, float step, bool ac) {
    float v;
    if (size <= 1) {
        v = -1.0f;
    } else {
        v = (idx < size / 2) ? __fadd_rn(-1.0f, __fmul_rn(step, (float)idx))
                             : __fsub_rn(1.0f, __fmul_rn(step, (float)(size - idx - 1)));
    }
    if (!ac) v = __fdiv_rn(__fmul_rn(v, (float)(size - 1)), (float)size);
    return v;
}

constexpr int kThreads = 256;  // block size of the flat (reduction / backward) kernels
constexpr int kUnroll = 4;     // pixels per thread in those kernels, kThreads apart
constexpr int kCols = 128;     // forward kernel: threads per CTA = columns per CTA
constexpr int kRows = 2;       // forward kernel: rows per thread (swept on B200: 2 beats 4, profiles/)

struct WarpArgs {
    const float *x; int64_t x_sb, x_sc, x_sf;
    const float *vis; int64_t vis_sb, vis_sf;
    const float *grid;
    const float *m_target; int64_t mt_sb;
    float *x_al; int64_t xa_sb, xa_sc, xa_sf;
    float *v_al; float *v_map;
    int F; int P;
    Sampler sp;
    bool affine, from_mask;
};

// grid coordinate of pixel p of frame n; NaN for p >= P (=> every corner out of
// bounds, no gather is issued for the slot)
__device__ __forceinline__ void load_coord(const float *__restrict__ grid, bool affine, const Sampler &sp,
                                           int64_t n, int p, int P, float &gx, float &gy) {
    if (p >= P) {
        gx = gy = __int_as_float(0x7fc00000);
        return;
    }
    if (!affine) {
        const float2 t = __ldcs(reinterpret_cast<const float2 *>(grid) + n * P + p);
        gx = t.x;
        gy = t.y;
    } else {
        const float *th = grid + n * 6;
        const int y = p / sp.W, xx = p - y * sp.W;
        const float bx = base_coord(xx, sp.W, sp.stepx, sp.ac), by = base_coord(y, sp.H, sp.stepy, sp.ac);
        // base_grid (x, y, 1) @ theta^T: fma(by, t1, bx*t0) + t2  (pinned order)
        gx = __fadd_rn(__fmaf_rn(by, __ldg(th + 1), __fmul_rn(bx, __ldg(th))), __ldg(th + 2));
        gy = __fadd_rn(__fmaf_rn(by, __ldg(th + 4), __fmul_rn(bx, __ldg(th + 3))), __ldg(th + 5));
    }
}

// Lean bilinear state for the forward kernel: weights + the offset of the NW corner.
struct Tap {
    float nw, ne, sw, se;
    int o;  // yn * W + xw (meaningful when the pixel is interior)
};

// Forward-kernel arguments: every offset is a 32-bit ELEMENT offset (the launcher
// checks the ranges), because 64-bit stride arithmetic dominated the per-thread
// setup of the first versions.
struct WarpFwdArgs {
    const float *x, *vis, *grid, *m_target;
    float *x_al, *v_al, *v_map;
    int x_sb, x_sc, x_sf, vis_sb, vis_sf, mt_sb, xa_sb, xa_sc, xa_sf;
    int F, P;
    unsigned f_magic;  // ceil(2^32 / F): b = umulhi(n, f_magic) for n < 2^16; 0 when F == 1 (b = n)
    Sampler sp;
    bool from_mask;
};

// Forward kernel.  Thread = one column x, U consecutive rows; a warp = 32
// consecutive columns of one row, so grid loads and all stores are 128 B
// coalesced and no integer division is needed.  If every tap of every lane of
// the warp is interior (0 <= xw <= W-2, 0 <= yn <= H-2: the common case) the
// warp takes the FAST path: unpredicated loads through one 64-bit row pointer
// per (plane, row), no bounds logic.  Otherwise it takes the generic path.
// The kernel is instruction-issue-bound, not HBM-bound (ncu: profiles/), so the
// code below is written for instruction count: template flags instead of
// uniform branches, the affine row term computed once per CTA, 32-bit offsets.
// VIS: 1 = nearest (DFPN), 2 = bilinear > 0.5 (CPN).  FULL: x_al, v_al and v_map
// are all requested (no NULL checks).
template <int C, int U, int VIS, bool AFFINE, bool FULL>
__global__ void __launch_bounds__(kCols) warp_fwd_kernel(const WarpFwdArgs a) {
    __shared__ float s_by[U];
    const int W = a.sp.W, H = a.sp.H;
    // lanes past the last column stay alive (the warp vote below needs every lane)
    // on a clamped column and are masked at the stores
    const bool live = (int)(blockIdx.x * kCols + threadIdx.x) < W;
    const int x = min((int)(blockIdx.x * kCols + threadIdx.x), W - 1);
    const int y0 = blockIdx.y * U;
    const unsigned n = blockIdx.z;
    const int b = a.f_magic ? (int)__umulhi(n, a.f_magic) : (int)n, f = (int)n - b * a.F;
    const bool ac = a.sp.ac;
    if (AFFINE) {  // row term of the affine grid: U values per CTA
        if (threadIdx.x < U) s_by[threadIdx.x] = base_coord(min(y0 + (int)threadIdx.x, H - 1), H, a.sp.stepy, ac);
        __syncthreads();
    }
    const int xo = b * a.x_sb + f * a.x_sf;
    const float *__restrict__ vp = a.vis + (b * a.vis_sb + f * a.vis_sf);
    const int p0 = y0 * W + x;
    const int np0 = (int)n * a.P + p0;

    float t1 = 0.f, t2 = 0.f, t4 = 0.f, t5 = 0.f, bxt0 = 0.f, bxt3 = 0.f;
    if (AFFINE) {
        const float *th = a.grid + n * 6;
        const float bx = base_coord(x, W, a.sp.stepx, ac);
        bxt0 = __fmul_rn(bx, __ldg(th));
        bxt3 = __fmul_rn(bx, __ldg(th + 3));
        t1 = __ldg(th + 1); t2 = __ldg(th + 2); t4 = __ldg(th + 4); t5 = __ldg(th + 5);
    }
    const float wm2 = a.sp.wmax - 1.0f, hm2 = a.sp.hmax - 1.0f;

    // issue the target-mask loads first: they are only needed by the stores at the end
    float mtv[U];
    const int mto = b * a.mt_sb + p0;
    if (FULL || a.v_map) {
#pragma unroll
        for (int k = 0; k < U; ++k) mtv[k] = (y0 + k < H) ? __ldcs(a.m_target + (mto + k * W)) : 0.0f;
    }

    float ix[U], iy[U], xw[U], yn[U];
    float wnw[U], wne[U], wsw[U], wse[U];
    bool interior = true;
#pragma unroll
    for (int k = 0; k < U; ++k) {
        float gx, gy;
        if (AFFINE) {
            const float by = s_by[k];
            gx = __fadd_rn(__fmaf_rn(by, t1, bxt0), t2);  // fma(by, t1, bx*t0) + t2 (pinned order)
            gy = __fadd_rn(__fmaf_rn(by, t4, bxt3), t5);
        } else if (y0 + k < H) {
            const float2 g = __ldcs(reinterpret_cast<const float2 *>(a.grid) + (np0 + k * W));
            gx = g.x; gy = g.y;
        } else {
            gx = gy = 0.0f;
        }
        ix[k] = unnormalize(gx, a.sp.sfx, ac);
        iy[k] = unnormalize(gy, a.sp.sfy, ac);
        xw[k] = floorf(ix[k]);
        yn[k] = floorf(iy[k]);
        const float w = __fsub_rn(ix[k], xw[k]), e = __fsub_rn(1.0f, w);
        const float nn = __fsub_rn(iy[k], yn[k]), ss = __fsub_rn(1.0f, nn);
        wnw[k] = __fmul_rn(ss, e); wne[k] = __fmul_rn(ss, w);
        wsw[k] = __fmul_rn(nn, e); wse[k] = __fmul_rn(nn, w);
        // rows past the end of the frame are computed (harmlessly) and never stored
        interior = interior && (xw[k] >= 0.0f) && (xw[k] <= wm2) && (yn[k] >= 0.0f) && (yn[k] <= hm2);
    }
    float xa[C][U], va[U];
    if (__all_sync(0xffffffffu, interior)) {
        // ---------------- fast path ----------------
        float c00[C + 1][U], c01[C + 1][U], c10[C + 1][U], c11[C + 1][U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int o = (int)yn[k] * W + (int)xw[k];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float *r0 = a.x + (xo + c * a.x_sc + o);
                const float *r1 = a.x + (xo + c * a.x_sc + o + W);
                c00[c][k] = __ldg(r0); c01[c][k] = __ldg(r0 + 1);
                c10[c][k] = __ldg(r1); c11[c][k] = __ldg(r1 + 1);
            }
            if (VIS == 2) {
                const float *r0 = vp + o;
                const float *r1 = vp + (o + W);
                c00[C][k] = __ldg(r0); c01[C][k] = __ldg(r0 + 1);
                c10[C][k] = __ldg(r1); c11[C][k] = __ldg(r1 + 1);
            } else {
                // nearest: rint (half-to-even) lands on one of the 4 interior taps: in bounds
                c00[C][k] = __ldg(vp + ((int)rintf(iy[k]) * W + (int)rintf(ix[k])));
            }
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
#pragma unroll
            for (int c = 0; c < C; ++c)
                xa[c][k] = __fmaf_rn(c11[c][k], wse[k], __fmaf_rn(c10[c][k], wsw[k],
                           __fmaf_rn(c01[c][k], wne[k], __fmul_rn(c00[c][k], wnw[k]))));
            if (VIS == 2) {
                float v00 = c00[C][k], v01 = c01[C][k], v10 = c10[C][k], v11 = c11[C][k];
                if (a.from_mask) {
                    v00 = __fsub_rn(1.0f, v00); v01 = __fsub_rn(1.0f, v01);
                    v10 = __fsub_rn(1.0f, v10); v11 = __fsub_rn(1.0f, v11);
                }
                const float vs = __fmaf_rn(v11, wse[k], __fmaf_rn(v10, wsw[k],
                                 __fmaf_rn(v01, wne[k], __fmul_rn(v00, wnw[k]))));
                va[k] = vs > 0.5f ? 1.0f : 0.0f;  // strict, model_cpn.py:88
            } else {
                va[k] = a.from_mask ? __fsub_rn(1.0f, c00[C][k]) : c00[C][k];
            }
        }
    } else {
        // ---------------- generic path (border warps, out-of-frame flows) ----------------
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const Bil bl = bil_params(ix[k], iy[k], a.sp);
#pragma unroll
            for (int c = 0; c < C; ++c) xa[c][k] = interp(gather(a.x + (xo + c * a.x_sc), bl, W), bl);
            if (VIS == 2) {
                Corners cv = gather(vp, bl, W);
                if (a.from_mask) {  // v = 1 - m inside the frame, 0 outside (zero padding of v)
                    cv.nw = (bl.y0 && bl.x0) ? __fsub_rn(1.0f, cv.nw) : 0.0f;
                    cv.ne = (bl.y0 && bl.x1) ? __fsub_rn(1.0f, cv.ne) : 0.0f;
                    cv.sw = (bl.y1 && bl.x0) ? __fsub_rn(1.0f, cv.sw) : 0.0f;
                    cv.se = (bl.y1 && bl.x1) ? __fsub_rn(1.0f, cv.se) : 0.0f;
                }
                va[k] = interp(cv, bl) > 0.5f ? 1.0f : 0.0f;
            } else {
                va[k] = nearest(vp, ix[k], iy[k], a.sp, a.from_mask);
            }
        }
    }
    // ---------------- stores (coalesced, streaming) ----------------
    if (!live) return;
    const int xao = b * a.xa_sb + f * a.xa_sf + p0;
#pragma unroll
    for (int k = 0; k < U; ++k) {
        if (y0 + k >= H) break;
        if (FULL || a.x_al) {
#pragma unroll
            for (int c = 0; c < C; ++c) st_stream1(a.x_al + (xao + c * a.xa_sc + k * W), xa[c][k]);
        }
        if (FULL || a.v_al) st_stream1(a.v_al + (np0 + k * W), va[k]);
        if (FULL || a.v_map)  // clamp(v_al - (1 - m_t), 0, 1)
            st_stream1(a.v_map + (np0 + k * W),
                       clamp01(__fsub_rn(va[k], __fsub_rn(1.0f, mtv[k]))));
    }
}

// ---- backward w.r.t. the dense grid (generic upstream gradient) -------------
struct WarpBwdArgs {
    const float *x; int64_t x_sb, x_sc, x_sf;
    const float *grid;
    const float *gout; int64_t g_sb, g_sc, g_sf;
    float *ggrid;
    int F; int P;
    Sampler sp;
};

template <int C, int U>
__global__ void __launch_bounds__(kThreads) warp_bwd_grid_kernel(const WarpBwdArgs a) {
    const int p0 = blockIdx.x * (kThreads * U) + threadIdx.x;
    const int64_t n = blockIdx.y;
    const int b = (int)(n / a.F), f = (int)(n - (int64_t)b * a.F);
    const float *xb = a.x + b * a.x_sb + f * a.x_sf;
    const float *gb = a.gout + b * a.g_sb + f * a.g_sf;
    Bil bl[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
        float gx, gy;
        load_coord(a.grid, false, a.sp, n, p0 + k * kThreads, a.P, gx, gy);
        bl[k] = bil_params(unnormalize(gx, a.sp.sfx, a.sp.ac), unnormalize(gy, a.sp.sfy, a.sp.ac), a.sp);
    }
    Corners cx[C][U];
    float go[C][U];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int k = 0; k < U; ++k) {
            cx[c][k] = gather(xb + c * a.x_sc, bl[k], a.sp.W);
            go[c][k] = (p0 + k * kThreads < a.P) ? __ldcs(gb + c * a.g_sc + p0 + k * kThreads) : 0.0f;
        }
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const int p = p0 + k * kThreads;
        if (p >= a.P) continue;
        float ax = 0.0f, ay = 0.0f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const Corners &q = cx[c][k];
            ax += ((q.ne - q.nw) * bl[k].s + (q.se - q.sw) * bl[k].n) * go[c][k];
            ay += ((q.sw - q.nw) * bl[k].e + (q.se - q.ne) * bl[k].w) * go[c][k];
        }
        __stcs(reinterpret_cast<float2 *>(a.ggrid) + n * a.P + p, make_float2(ax * a.sp.sfx, ay * a.sp.sfy));
    }
}

// ---- fused warp + mask_out + masked L1 ('sum') ------------------------------
struct WarpL1Args {
    const float *x; int64_t x_sb, x_sc, x_sf;
    const float *vis; int64_t vis_sb, vis_sf;
    const float *flow;
    const float *xt; int64_t xt_sb, xt_sc;
    const float *vt; int64_t vt_sb;
    float *x_al; float *v_al;  // frame-major, may be NULL
    float *out3; void *ws;
    const float *out3_in; const float *grad_out; float *gflow;  // backward only
    int F; int P; int tiles;  // tiles = ceil(P / (kThreads*U))
    int64_t total_tiles;      // B*F*tiles
    float weight;
    Sampler sp;
    bool from_mask;
};

__device__ __forceinline__ float mask_out_of(float gx, float gy) {
    // clamp((gx<-1)+(gx>1)+(gy<-1)+(gy>1), 0, 1)   model_dfpn.py:269-272
    return (gx < -1.0f || gx > 1.0f || gy < -1.0f || gy > 1.0f) ? 1.0f : 0.0f;
}

template <int U>
__global__ void __launch_bounds__(kThreads) warp_l1_fwd_kernel(const WarpL1Args a) {
    __shared__ float red[2 * 32];
    float acc[2] = {0.0f, 0.0f};  // sum |x_t*M - x_al*M| over 3 channels, sum M
    for (int64_t t = blockIdx.x; t < a.total_tiles; t += gridDim.x) {
        const int64_t n = t / a.tiles;
        const int p0 = (int)(t - n * a.tiles) * (kThreads * U) + threadIdx.x;
        const int b = (int)(n / a.F), f = (int)(n - (int64_t)b * a.F);
        const float *xb = a.x + b * a.x_sb + f * a.x_sf;
        Bil bl[U];
        float ix[U], iy[U], M[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int p = p0 + k * kThreads;
            float gx, gy;
            load_coord(a.flow, false, a.sp, n, p, a.P, gx, gy);
            const float vt = p < a.P ? __ldg(a.vt + b * a.vt_sb + p) : 0.0f;
            ix[k] = unnormalize(gx, a.sp.sfx, a.sp.ac);
            iy[k] = unnormalize(gy, a.sp.sfy, a.sp.ac);
            bl[k] = bil_params(ix[k], iy[k], a.sp);
            M[k] = p < a.P ? __fmul_rn(vt, __fsub_rn(1.0f, mask_out_of(gx, gy))) : 0.0f;
            acc[1] += M[k];
        }
        Corners cx[3][U];
        float xt[3][U];
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int k = 0; k < U; ++k) {
                cx[c][k] = gather(xb + c * a.x_sc, bl[k], a.sp.W);
                xt[c][k] = (p0 + k * kThreads < a.P) ? __ldg(a.xt + b * a.xt_sb + c * a.xt_sc + p0 + k * kThreads) : 0.0f;
            }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int p = p0 + k * kThreads;
            if (p >= a.P) continue;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float r = interp(cx[c][k], bl[k]);
                acc[0] += fabsf(__fsub_rn(__fmul_rn(xt[c][k], M[k]), __fmul_rn(r, M[k])));
                if (a.x_al) st_stream1(a.x_al + (n * 3 + c) * a.P + p, r);
            }
            if (a.v_al) {
                const float *vp = a.vis + b * a.vis_sb + f * a.vis_sf;
                st_stream1(a.v_al + n * a.P + p, nearest(vp, ix[k], iy[k], a.sp, a.from_mask));
            }
        }
    }
    float *out3 = a.out3;
    const float weight = a.weight;
    grid_reduce_finish<2>(acc, a.ws, red, [out3, weight](const double *tot) {
        const float num = (float)tot[0], den = (float)tot[1];
        out3[0] = weight * num / (den + 1e-9f);  // utils.py:167-169
        out3[1] = num;
        out3[2] = den;
    });
}

// d loss / d flow in one pass: recomputes the sampling, never materialises
// x_aligned or its gradient.
template <int U>
__global__ void __launch_bounds__(kThreads) warp_l1_bwd_kernel(const WarpL1Args a) {
    const int p0 = blockIdx.x * (kThreads * U) + threadIdx.x;
    const int64_t n = blockIdx.y;
    const int b = (int)(n / a.F), f = (int)(n - (int64_t)b * a.F);
    const float scale = a.weight * __ldg(a.grad_out) / (__ldg(a.out3_in + 2) + 1e-9f);
    const float *xb = a.x + b * a.x_sb + f * a.x_sf;
    Bil bl[U];
    float M[U];
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const int p = p0 + k * kThreads;
        float gx, gy;
        load_coord(a.flow, false, a.sp, n, p, a.P, gx, gy);
        const float vt = p < a.P ? __ldg(a.vt + b * a.vt_sb + p) : 0.0f;
        bl[k] = bil_params(unnormalize(gx, a.sp.sfx, a.sp.ac), unnormalize(gy, a.sp.sfy, a.sp.ac), a.sp);
        M[k] = p < a.P ? __fmul_rn(vt, __fsub_rn(1.0f, mask_out_of(gx, gy))) : 0.0f;
    }
    Corners cx[3][U];
    float xt[3][U];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int k = 0; k < U; ++k) {
            cx[c][k] = gather(xb + c * a.x_sc, bl[k], a.sp.W);
            xt[c][k] = (p0 + k * kThreads < a.P) ? __ldg(a.xt + b * a.xt_sb + c * a.xt_sc + p0 + k * kThreads) : 0.0f;
        }
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const int p = p0 + k * kThreads;
        if (p >= a.P) continue;
        float ax = 0.0f, ay = 0.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const Corners &q = cx[c][k];
            const float d = __fsub_rn(__fmul_rn(xt[c][k], M[k]), __fmul_rn(interp(q, bl[k]), M[k]));
            const float sg = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
            const float go = -sg * M[k] * scale;  // d loss / d x_aligned
            ax += ((q.ne - q.nw) * bl[k].s + (q.se - q.sw) * bl[k].n) * go;
            ay += ((q.sw - q.nw) * bl[k].e + (q.se - q.ne) * bl[k].w) * go;
        }
        __stcs(reinterpret_cast<float2 *>(a.gflow) + n * a.P + p, make_float2(ax * a.sp.sfx, ay * a.sp.sfy));
    }
}

__global__ void __launch_bounds__(256) mask_out_kernel(const float *__restrict__ flow, int64_t n,
                                                       float *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float2 g = __ldcs(reinterpret_cast<const float2 *>(flow) + i);
        st_stream1(out + i, mask_out_of(g.x, g.y));
    }
}

Sampler make_sampler(int H, int W, bool ac) {
    Sampler s;
    s.sfx = ac ? (float)(W - 1) / 2.0f : (float)W / 2.0f;
    s.sfy = ac ? (float)(H - 1) / 2.0f : (float)H / 2.0f;
    s.wmax = (float)(W - 1);
    s.hmax = (float)(H - 1);
    s.stepx = W > 1 ? 2.0f / (float)(W - 1) : 0.0f;
    s.stepy = H > 1 ? 2.0f / (float)(H - 1) : 0.0f;
    s.W = W;
    s.H = H;
    s.ac = ac;
    return s;
}

bool aligned8(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 7u) == 0; }

}  // namespace
}  // namespace mt

using namespace mt;

extern "C" int mt_warp_fwd(const float *x, int64_t x_sb, int64_t x_sc, int64_t x_sf,
                           const float *vis, int64_t vis_sb, int64_t vis_sf, const float *grid,
                           const float *m_target, int64_t mt_sb, float *x_aligned, int64_t xa_sb,
                           int64_t xa_sc, int64_t xa_sf, float *v_aligned, float *v_map, int B,
                           int C, int F, int H, int W, int flags, mt_stream_t stream) {
    MT_REQUIRE(x && vis && grid, "mt_warp_fwd: NULL input");
    MT_REQUIRE(B > 0 && F > 0 && H > 0 && W > 0, "mt_warp_fwd: empty shape B=%d F=%d H=%d W=%d", B, F, H, W);
    MT_REQUIRE(C == 1 || C == 3, "mt_warp_fwd: C must be 1 or 3 (reference hard-codes 3, utils.py:97), got %d", C);
    MT_REQUIRE((int64_t)B * F <= 65535, "mt_warp_fwd: B*F > 65535");
    MT_REQUIRE(!v_map || m_target, "mt_warp_fwd: v_map needs m_target");
    MT_REQUIRE((flags & MT_GRID_AFFINE) || aligned8(grid), "mt_warp_fwd: dense grid must be 8 B aligned");
    MT_REQUIRE((H + kRows - 1) / kRows <= 65535, "mt_warp_fwd: H too large");
    const int64_t P = (int64_t)H * W, lim = (1ll << 31) - 1;
    auto span = [&](int64_t sb, int64_t sc, int64_t sf, int c) {
        return (B - 1) * sb + (c - 1) * sc + (F - 1) * sf + P + W + 1;
    };
    MT_REQUIRE(x_sb >= 0 && x_sc >= 0 && x_sf >= 0 && vis_sb >= 0 && vis_sf >= 0 && mt_sb >= 0 &&
               xa_sb >= 0 && xa_sc >= 0 && xa_sf >= 0, "mt_warp_fwd: negative strides are not supported");
    MT_REQUIRE(span(x_sb, x_sc, x_sf, C) < lim && span(vis_sb, 0, vis_sf, 1) < lim &&
               span(xa_sb, xa_sc, xa_sf, C) < lim && (int64_t)B * F * P * 2 < lim && (B - 1) * mt_sb + P < lim,
               "mt_warp_fwd: tensors beyond 2^31 elements are not supported (split the batch)");
    WarpFwdArgs a;
    a.x = x; a.vis = vis; a.grid = grid; a.m_target = m_target;
    a.x_al = x_aligned; a.v_al = v_aligned; a.v_map = v_map;
    a.x_sb = (int)x_sb; a.x_sc = (int)x_sc; a.x_sf = (int)x_sf;
    a.vis_sb = (int)vis_sb; a.vis_sf = (int)vis_sf; a.mt_sb = (int)mt_sb;
    a.xa_sb = (int)xa_sb; a.xa_sc = (int)xa_sc; a.xa_sf = (int)xa_sf;
    a.F = F; a.P = (int)P;
    a.f_magic = F == 1 ? 0u : (unsigned)(((1ull << 32) + (unsigned)F - 1) / (unsigned)F);
    a.sp = make_sampler(H, W, (flags & MT_ALIGN_CORNERS) != 0);
    a.from_mask = (flags & MT_VIS_FROM_MASK) != 0;
    const bool affine = (flags & MT_GRID_AFFINE) != 0;
    const bool vis_bil = (flags & MT_VIS_BILINEAR) != 0;
    const int rows = tuning("MT_WARP_ROWS", kRows) == 2 ? 2 : 4;
    dim3 block(kCols), gridd((W + kCols - 1) / kCols, (H + rows - 1) / rows, B * F);
    cudaStream_t st = (cudaStream_t)stream;
    const bool full = x_aligned && v_aligned && v_map;
#define MT_WARP_GO(CC, VV, AA, FF)                                                  \
    do {                                                                            \
        if (rows == 2) warp_fwd_kernel<CC, 2, VV, AA, FF><<<gridd, block, 0, st>>>(a); \
        else warp_fwd_kernel<CC, 4, VV, AA, FF><<<gridd, block, 0, st>>>(a);        \
    } while (0)
#define MT_WARP_PICK(CC)                                                            \
    do {                                                                            \
        if (vis_bil) {                                                              \
            if (affine) { if (full) MT_WARP_GO(CC, 2, true, true); else MT_WARP_GO(CC, 2, true, false); }   \
            else        { if (full) MT_WARP_GO(CC, 2, false, true); else MT_WARP_GO(CC, 2, false, false); } \
        } else {                                                                    \
            if (affine) { if (full) MT_WARP_GO(CC, 1, true, true); else MT_WARP_GO(CC, 1, true, false); }   \
            else        { if (full) MT_WARP_GO(CC, 1, false, true); else MT_WARP_GO(CC, 1, false, false); } \
        }                                                                           \
    } while (0)
    if (C == 3) MT_WARP_PICK(3); else MT_WARP_PICK(1);
#undef MT_WARP_PICK
#undef MT_WARP_GO
    return launch_status("mt_warp_fwd");
}

extern "C" int mt_warp_bwd_grid(const float *x, int64_t x_sb, int64_t x_sc, int64_t x_sf,
                                const float *grid, const float *gout, int64_t g_sb, int64_t g_sc,
                                int64_t g_sf, float *ggrid, int B, int C, int F, int H, int W,
                                int flags, mt_stream_t stream) {
    MT_REQUIRE(x && grid && gout && ggrid, "mt_warp_bwd_grid: NULL argument");
    MT_REQUIRE(B > 0 && F > 0 && H > 0 && W > 0, "mt_warp_bwd_grid: empty shape");
    MT_REQUIRE(C == 1 || C == 3, "mt_warp_bwd_grid: C must be 1 or 3, got %d", C);
    MT_REQUIRE(!(flags & MT_GRID_AFFINE), "mt_warp_bwd_grid: dense grids only");
    MT_REQUIRE((int64_t)H * W < (1ll << 30) && (int64_t)B * F <= 65535, "mt_warp_bwd_grid: too large");
    MT_REQUIRE(aligned8(grid) && aligned8(ggrid), "mt_warp_bwd_grid: grids must be 8 B aligned");
    WarpBwdArgs a;
    a.x = x; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sf = x_sf; a.grid = grid;
    a.gout = gout; a.g_sb = g_sb; a.g_sc = g_sc; a.g_sf = g_sf; a.ggrid = ggrid;
    a.F = F; a.P = H * W;
    a.sp = make_sampler(H, W, (flags & MT_ALIGN_CORNERS) != 0);
    dim3 block(kThreads), gridd((a.P + kThreads * kUnroll - 1) / (kThreads * kUnroll), B * F);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 3) warp_bwd_grid_kernel<3, kUnroll><<<gridd, block, 0, st>>>(a);
    else warp_bwd_grid_kernel<1, kUnroll><<<gridd, block, 0, st>>>(a);
    return launch_status("mt_warp_bwd_grid");
}

static int fill_l1_args(WarpL1Args &a, const float *x, int64_t x_sb, int64_t x_sc, int64_t x_sf,
                        const float *flow, const float *x_target, int64_t xt_sb, int64_t xt_sc,
                        const float *v_target, int64_t vt_sb, int B, int F, int H, int W,
                        float weight, int flags, const char *who) {
    MT_REQUIRE(x && flow && x_target && v_target, "%s: NULL input", who);
    MT_REQUIRE(B > 0 && F > 0 && H > 0 && W > 0, "%s: empty shape", who);
    MT_REQUIRE((int64_t)H * W < (1ll << 30) && (int64_t)B * F <= 65535, "%s: too large", who);
    MT_REQUIRE(aligned8(flow), "%s: flow must be 8 B aligned", who);
    a.x = x; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sf = x_sf; a.flow = flow;
    a.xt = x_target; a.xt_sb = xt_sb; a.xt_sc = xt_sc; a.vt = v_target; a.vt_sb = vt_sb;
    a.F = F; a.P = H * W; a.weight = weight;
    a.sp = make_sampler(H, W, (flags & MT_ALIGN_CORNERS) != 0);
    a.from_mask = (flags & MT_VIS_FROM_MASK) != 0;
    a.vis = nullptr; a.vis_sb = a.vis_sf = 0; a.x_al = a.v_al = nullptr; a.out3 = nullptr; a.ws = nullptr;
    a.out3_in = nullptr; a.grad_out = nullptr; a.gflow = nullptr;
    a.tiles = (a.P + kThreads * kUnroll - 1) / (kThreads * kUnroll);
    a.total_tiles = (int64_t)B * F * a.tiles;
    return MT_OK;
}

extern "C" int mt_warp_l1_fwd(const float *x, int64_t x_sb, int64_t x_sc, int64_t x_sf,
                              const float *vis, int64_t vis_sb, int64_t vis_sf, const float *flow,
                              const float *x_target, int64_t xt_sb, int64_t xt_sc,
                              const float *v_target, int64_t vt_sb, float *x_aligned,
                              float *v_aligned, float *out3, void *workspace, int B, int F, int H,
                              int W, float weight, int flags, mt_stream_t stream) {
    WarpL1Args a;
    int rc = fill_l1_args(a, x, x_sb, x_sc, x_sf, flow, x_target, xt_sb, xt_sc, v_target, vt_sb, B,
                          F, H, W, weight, flags, "mt_warp_l1_fwd");
    if (rc) return rc;
    MT_REQUIRE(out3 && workspace, "mt_warp_l1_fwd: NULL out3 / workspace");
    MT_REQUIRE(!v_aligned || vis, "mt_warp_l1_fwd: v_aligned needs vis");
    a.vis = vis; a.vis_sb = vis_sb; a.vis_sf = vis_sf; a.x_al = x_aligned; a.v_al = v_aligned;
    a.out3 = out3; a.ws = workspace;
    int64_t want = (int64_t)sm_count() * 8;
    int nblk = (int)(a.total_tiles < want ? a.total_tiles : want);
    if (nblk > kMaxReduceBlocks) nblk = kMaxReduceBlocks;
    warp_l1_fwd_kernel<kUnroll><<<nblk, kThreads, 0, (cudaStream_t)stream>>>(a);
    return launch_status("mt_warp_l1_fwd");
}

extern "C" int mt_warp_l1_bwd(const float *x, int64_t x_sb, int64_t x_sc, int64_t x_sf,
                              const float *flow, const float *x_target, int64_t xt_sb,
                              int64_t xt_sc, const float *v_target, int64_t vt_sb,
                              const float *out3, const float *grad_out, float *gflow, int B, int F,
                              int H, int W, float weight, int flags, mt_stream_t stream) {
    WarpL1Args a;
    int rc = fill_l1_args(a, x, x_sb, x_sc, x_sf, flow, x_target, xt_sb, xt_sc, v_target, vt_sb, B,
                          F, H, W, weight, flags, "mt_warp_l1_bwd");
    if (rc) return rc;
    MT_REQUIRE(out3 && grad_out && gflow, "mt_warp_l1_bwd: NULL out3 / grad_out / gflow");
    MT_REQUIRE(aligned8(gflow), "mt_warp_l1_bwd: gflow must be 8 B aligned");
    a.out3_in = out3; a.grad_out = grad_out; a.gflow = gflow;
    dim3 block(kThreads), gridd(a.tiles, B * F);
    warp_l1_bwd_kernel<kUnroll><<<gridd, block, 0, (cudaStream_t)stream>>>(a);
    return launch_status("mt_warp_l1_bwd");
}

extern "C" int mt_mask_out(const float *flow, int64_t n, float *out, mt_stream_t stream) {
    MT_REQUIRE(flow && out && n > 0, "mt_mask_out: bad argument");
    MT_REQUIRE((reinterpret_cast<uintptr_t>(flow) & 7u) == 0, "mt_mask_out: flow must be 8 B aligned");
    int64_t nb = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
    mask_out_kernel<<<(int)(nb < cap ? nb : cap), 256, 0, (cudaStream_t)stream>>>(flow, n, out);
    return launch_status("mt_mask_out");
}
