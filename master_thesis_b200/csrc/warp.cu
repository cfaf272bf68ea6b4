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
// (H, W) planes.  One thread owns VEC = 4 consecutive pixels of the flattened
// plane: the grid is read with two 16 B streaming loads, the outputs leave
// with one 16 B streaming store per plane, and the 4-corner gathers go through
// L1 (neighbouring pixels of a coherent flow share cache lines).  No shared
// memory staging in this version: see DESIGN.md for the ncu evidence.
#include <math.h>

#include "mt_common.cuh"

namespace mt {
namespace {

struct Sampler {
    float sfx, sfy;  // (W-1)/2 | W/2, (H-1)/2 | H/2   (host-computed in fp32)
    float wmax, hmax;
    int W;
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
__device__ __forceinline__ float base_coord(int idx, int size, bool ac) {
    float v;
    if (size <= 1) {
        v = -1.0f;
    } else {
        const float step = __fdiv_rn(2.0f, (float)(size - 1));
        v = (idx < size / 2) ? __fadd_rn(-1.0f, __fmul_rn(step, (float)idx))
                             : __fsub_rn(1.0f, __fmul_rn(step, (float)(size - idx - 1)));
    }
    if (!ac) v = __fdiv_rn(__fmul_rn(v, (float)(size - 1)), (float)size);
    return v;
}

struct WarpArgs {
    const float *x; int64_t x_sb, x_sc, x_sf;
    const float *vis; int64_t vis_sb, vis_sf;
    const float *grid;
    const float *m_target; int64_t mt_sb;
    float *x_al; int64_t xa_sb, xa_sc, xa_sf;
    float *v_al; float *v_map;
    int F, H, W; int P;
    Sampler sp;
    bool affine, from_mask;
};

// grid coordinates of VEC consecutive pixels starting at flat index p0
template <int VEC>
__device__ __forceinline__ void load_coords(const WarpArgs &a, int64_t n, int p0, float (&gx)[VEC],
                                            float (&gy)[VEC]) {
    if (!a.affine) {
        const float *g = a.grid + (n * a.P + p0) * 2;
        if (VEC == 4) {
            const float4 g0 = ld_stream4(g), g1 = ld_stream4(g + 4);
            gx[0] = g0.x; gy[0] = g0.y; gx[1 % VEC] = g0.z; gy[1 % VEC] = g0.w;
            gx[2 % VEC] = g1.x; gy[2 % VEC] = g1.y; gx[3 % VEC] = g1.z; gy[3 % VEC] = g1.w;
        } else {
            const float2 t = __ldcs(reinterpret_cast<const float2 *>(g));
            gx[0] = t.x; gy[0] = t.y;
        }
    } else {
        const float *th = a.grid + n * 6;
        const float t0 = __ldg(th), t1 = __ldg(th + 1), t2 = __ldg(th + 2);
        const float t3 = __ldg(th + 3), t4 = __ldg(th + 4), t5 = __ldg(th + 5);
        int y = p0 / a.W, xx = p0 - y * a.W;
        float by = base_coord(y, a.H, a.sp.ac);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float bx = base_coord(xx, a.W, a.sp.ac);
            // base_grid (x, y, 1) @ theta^T: fma(by, t1, bx*t0) + t2  (pinned order)
            gx[i] = __fadd_rn(__fmaf_rn(by, t1, __fmul_rn(bx, t0)), t2);
            gy[i] = __fadd_rn(__fmaf_rn(by, t4, __fmul_rn(bx, t3)), t5);
            if (++xx == a.W) { xx = 0; ++y; by = base_coord(y, a.H, a.sp.ac); }
        }
    }
}

template <int C, int VEC, bool VIS_BIL>
__global__ void __launch_bounds__(256) warp_fwd_kernel(const WarpArgs a) {
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (p0 >= a.P) return;
    const int64_t n = blockIdx.y;
    const int b = (int)(n / a.F), f = (int)(n - (int64_t)b * a.F);

    float gx[VEC], gy[VEC];
    load_coords<VEC>(a, n, p0, gx, gy);
    Vec<VEC> mt;
    if (a.v_map) mt.load_stream(a.m_target + b * a.mt_sb + p0);

    const float *xb = a.x + b * a.x_sb + f * a.x_sf;
    const float *vp = a.vis + b * a.vis_sb + f * a.vis_sf;

    Bil bl[VEC];
    float ix[VEC], iy[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        ix[i] = unnormalize(gx[i], a.sp.sfx, a.sp.ac);
        iy[i] = unnormalize(gy[i], a.sp.sfy, a.sp.ac);
        bl[i] = bil_params(ix[i], iy[i], a.sp);
    }
    // issue every gather before any use: VEC * (4C + 1..4) loads in flight
    Corners cx[C][VEC];
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int i = 0; i < VEC; ++i) cx[c][i] = gather(xb + c * a.x_sc, bl[i], a.W);
    Vec<VEC> va;
    if (VIS_BIL) {
        Corners cv[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            cv[i] = gather(vp, bl[i], a.W);
            if (a.from_mask) {  // v = 1 - m inside the frame, 0 outside (zero padding of v)
                cv[i].nw = (bl[i].y0 && bl[i].x0) ? __fsub_rn(1.0f, cv[i].nw) : 0.0f;
                cv[i].ne = (bl[i].y0 && bl[i].x1) ? __fsub_rn(1.0f, cv[i].ne) : 0.0f;
                cv[i].sw = (bl[i].y1 && bl[i].x0) ? __fsub_rn(1.0f, cv[i].sw) : 0.0f;
                cv[i].se = (bl[i].y1 && bl[i].x1) ? __fsub_rn(1.0f, cv[i].se) : 0.0f;
            }
            va.v[i] = interp(cv[i], bl[i]) > 0.5f ? 1.0f : 0.0f;  // strict, model_cpn.py:88
        }
    } else {
#pragma unroll
        for (int i = 0; i < VEC; ++i) va.v[i] = nearest(vp, ix[i], iy[i], a.sp, a.from_mask);
    }

    if (a.x_al) {
        float *o = a.x_al + b * a.xa_sb + f * a.xa_sf + p0;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            Vec<VEC> r;
#pragma unroll
            for (int i = 0; i < VEC; ++i) r.v[i] = interp(cx[c][i], bl[i]);
            r.store_stream(o + c * a.xa_sc);
        }
    }
    if (a.v_al) va.store_stream(a.v_al + n * a.P + p0);
    if (a.v_map) {
        Vec<VEC> vm;
#pragma unroll
        for (int i = 0; i < VEC; ++i)  // clamp(v_al - (1 - m_t), 0, 1)
            vm.v[i] = clamp01(__fsub_rn(va.v[i], __fsub_rn(1.0f, mt.v[i])));
        vm.store_stream(a.v_map + n * a.P + p0);
    }
}

// ---- backward w.r.t. the dense grid (generic upstream gradient) -------------
struct WarpBwdArgs {
    const float *x; int64_t x_sb, x_sc, x_sf;
    const float *grid;
    const float *gout; int64_t g_sb, g_sc, g_sf;
    float *ggrid;
    int F, H, W; int P;
    Sampler sp;
};

template <int C, int VEC>
__global__ void __launch_bounds__(256) warp_bwd_grid_kernel(const WarpBwdArgs a) {
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (p0 >= a.P) return;
    const int64_t n = blockIdx.y;
    const int b = (int)(n / a.F), f = (int)(n - (int64_t)b * a.F);
    WarpArgs wa;
    wa.grid = a.grid; wa.P = a.P; wa.affine = false;
    float gx[VEC], gy[VEC];
    load_coords<VEC>(wa, n, p0, gx, gy);
    const float *xb = a.x + b * a.x_sb + f * a.x_sf;
    const float *gb = a.gout + b * a.g_sb + f * a.g_sf + p0;
    float ax[VEC], ay[VEC];
    Bil bl[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        bl[i] = bil_params(unnormalize(gx[i], a.sp.sfx, a.sp.ac), unnormalize(gy[i], a.sp.sfy, a.sp.ac), a.sp);
        ax[i] = 0.0f; ay[i] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        Vec<VEC> go;
        go.load_stream(gb + c * a.g_sc);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const Corners k = gather(xb + c * a.x_sc, bl[i], a.W);
            ax[i] += ((k.ne - k.nw) * bl[i].s + (k.se - k.sw) * bl[i].n) * go.v[i];
            ay[i] += ((k.sw - k.nw) * bl[i].e + (k.se - k.ne) * bl[i].w) * go.v[i];
        }
    }
    float *o = a.ggrid + (n * a.P + p0) * 2;
    if (VEC == 4) {
        st_stream4(o, make_float4(ax[0] * a.sp.sfx, ay[0] * a.sp.sfy, ax[1 % VEC] * a.sp.sfx, ay[1 % VEC] * a.sp.sfy));
        st_stream4(o + 4, make_float4(ax[2 % VEC] * a.sp.sfx, ay[2 % VEC] * a.sp.sfy, ax[3 % VEC] * a.sp.sfx, ay[3 % VEC] * a.sp.sfy));
    } else {
        __stcs(reinterpret_cast<float2 *>(o), make_float2(ax[0] * a.sp.sfx, ay[0] * a.sp.sfy));
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
    int F, H, W; int P; int chunks;  // chunks = ceil(P / (256*VEC))
    int64_t total_chunks;            // B*F*chunks
    float weight;
    Sampler sp;
    bool from_mask;
};

__device__ __forceinline__ float mask_out_of(float gx, float gy) {
    // clamp((gx<-1)+(gx>1)+(gy<-1)+(gy>1), 0, 1)   model_dfpn.py:269-272
    return (gx < -1.0f || gx > 1.0f || gy < -1.0f || gy > 1.0f) ? 1.0f : 0.0f;
}

template <int VEC>
__global__ void __launch_bounds__(256) warp_l1_fwd_kernel(const WarpL1Args a) {
    __shared__ float red[2 * 32];
    float acc[2] = {0.0f, 0.0f};  // sum |x_t*M - x_al*M| over 3 channels, sum M
    WarpArgs wa;
    wa.grid = a.flow; wa.P = a.P; wa.affine = false;
    for (int64_t ch = blockIdx.x; ch < a.total_chunks; ch += gridDim.x) {
        const int64_t n = ch / a.chunks;
        const int p0 = ((int)(ch - n * a.chunks) * blockDim.x + threadIdx.x) * VEC;
        if (p0 >= a.P) continue;
        const int b = (int)(n / a.F), f = (int)(n - (int64_t)b * a.F);
        float gx[VEC], gy[VEC];
        load_coords<VEC>(wa, n, p0, gx, gy);
        const float *xb = a.x + b * a.x_sb + f * a.x_sf;
        Vec<VEC> vt;
        vt.load_cached(a.vt + b * a.vt_sb + p0);
        Bil bl[VEC];
        float ix[VEC], iy[VEC], M[VEC];
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            ix[i] = unnormalize(gx[i], a.sp.sfx, a.sp.ac);
            iy[i] = unnormalize(gy[i], a.sp.sfy, a.sp.ac);
            bl[i] = bil_params(ix[i], iy[i], a.sp);
            M[i] = __fmul_rn(vt.v[i], __fsub_rn(1.0f, mask_out_of(gx[i], gy[i])));
            acc[1] += M[i];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            Vec<VEC> xt, r;
            xt.load_cached(a.xt + b * a.xt_sb + c * a.xt_sc + p0);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                r.v[i] = interp(gather(xb + c * a.x_sc, bl[i], a.W), bl[i]);
                acc[0] += fabsf(__fsub_rn(__fmul_rn(xt.v[i], M[i]), __fmul_rn(r.v[i], M[i])));
            }
            if (a.x_al) r.store_stream(a.x_al + (n * 3 + c) * a.P + p0);
        }
        if (a.v_al) {
            const float *vp = a.vis + b * a.vis_sb + f * a.vis_sf;
            Vec<VEC> va;
#pragma unroll
            for (int i = 0; i < VEC; ++i) va.v[i] = nearest(vp, ix[i], iy[i], a.sp, a.from_mask);
            va.store_stream(a.v_al + n * a.P + p0);
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
template <int VEC>
__global__ void __launch_bounds__(256) warp_l1_bwd_kernel(const WarpL1Args a) {
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    if (p0 >= a.P) return;
    const int64_t n = blockIdx.y;
    const int b = (int)(n / a.F), f = (int)(n - (int64_t)b * a.F);
    WarpArgs wa;
    wa.grid = a.flow; wa.P = a.P; wa.affine = false;
    float gx[VEC], gy[VEC];
    load_coords<VEC>(wa, n, p0, gx, gy);
    const float scale = a.weight * __ldg(a.grad_out) / (__ldg(a.out3_in + 2) + 1e-9f);
    const float *xb = a.x + b * a.x_sb + f * a.x_sf;
    Vec<VEC> vt;
    vt.load_cached(a.vt + b * a.vt_sb + p0);
    Bil bl[VEC];
    float M[VEC], ax[VEC], ay[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        bl[i] = bil_params(unnormalize(gx[i], a.sp.sfx, a.sp.ac), unnormalize(gy[i], a.sp.sfy, a.sp.ac), a.sp);
        M[i] = __fmul_rn(vt.v[i], __fsub_rn(1.0f, mask_out_of(gx[i], gy[i])));
        ax[i] = 0.0f; ay[i] = 0.0f;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        Vec<VEC> xt;
        xt.load_cached(a.xt + b * a.xt_sb + c * a.xt_sc + p0);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const Corners k = gather(xb + c * a.x_sc, bl[i], a.W);
            const float d = __fsub_rn(__fmul_rn(xt.v[i], M[i]), __fmul_rn(interp(k, bl[i]), M[i]));
            const float sg = (d > 0.0f) ? 1.0f : ((d < 0.0f) ? -1.0f : 0.0f);
            const float go = -sg * M[i] * scale;  // d loss / d x_aligned
            ax[i] += ((k.ne - k.nw) * bl[i].s + (k.se - k.sw) * bl[i].n) * go;
            ay[i] += ((k.sw - k.nw) * bl[i].e + (k.se - k.ne) * bl[i].w) * go;
        }
    }
    float *o = a.gflow + (n * a.P + p0) * 2;
    if (VEC == 4) {
        st_stream4(o, make_float4(ax[0] * a.sp.sfx, ay[0] * a.sp.sfy, ax[1 % VEC] * a.sp.sfx, ay[1 % VEC] * a.sp.sfy));
        st_stream4(o + 4, make_float4(ax[2 % VEC] * a.sp.sfx, ay[2 % VEC] * a.sp.sfy, ax[3 % VEC] * a.sp.sfx, ay[3 % VEC] * a.sp.sfy));
    } else {
        __stcs(reinterpret_cast<float2 *>(o), make_float2(ax[0] * a.sp.sfx, ay[0] * a.sp.sfy));
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
    s.W = W;
    s.ac = ac;
    return s;
}

bool mult4(int64_t v) { return (v & 3) == 0; }

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
    MT_REQUIRE((int64_t)H * W < (1ll << 30), "mt_warp_fwd: plane too large");
    MT_REQUIRE((int64_t)B * F <= 65535, "mt_warp_fwd: B*F > 65535");
    MT_REQUIRE(!v_map || m_target, "mt_warp_fwd: v_map needs m_target");
    WarpArgs a;
    a.x = x; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sf = x_sf;
    a.vis = vis; a.vis_sb = vis_sb; a.vis_sf = vis_sf;
    a.grid = grid; a.m_target = m_target; a.mt_sb = mt_sb;
    a.x_al = x_aligned; a.xa_sb = xa_sb; a.xa_sc = xa_sc; a.xa_sf = xa_sf;
    a.v_al = v_aligned; a.v_map = v_map;
    a.F = F; a.H = H; a.W = W; a.P = H * W;
    a.sp = make_sampler(H, W, (flags & MT_ALIGN_CORNERS) != 0);
    a.affine = (flags & MT_GRID_AFFINE) != 0;
    a.from_mask = (flags & MT_VIS_FROM_MASK) != 0;
    const bool vis_bil = (flags & MT_VIS_BILINEAR) != 0;
    // VEC=4 needs every per-frame base 16 B aligned
    bool v4 = mult4(a.P) && aligned16(x_aligned) && aligned16(v_aligned) && aligned16(v_map) &&
              aligned16(m_target) && mult4(mt_sb) && mult4(xa_sb) && mult4(xa_sc) && mult4(xa_sf) &&
              (a.affine || aligned16(grid));
    const int vec = v4 ? 4 : 1;
    dim3 block(256), gridd((a.P + 256 * vec - 1) / (256 * vec), B * F);
    cudaStream_t st = (cudaStream_t)stream;
#define MT_LAUNCH_WARP(CC, VV, BB) warp_fwd_kernel<CC, VV, BB><<<gridd, block, 0, st>>>(a)
    if (C == 3) {
        if (v4) { if (vis_bil) MT_LAUNCH_WARP(3, 4, true); else MT_LAUNCH_WARP(3, 4, false); }
        else    { if (vis_bil) MT_LAUNCH_WARP(3, 1, true); else MT_LAUNCH_WARP(3, 1, false); }
    } else {
        if (v4) { if (vis_bil) MT_LAUNCH_WARP(1, 4, true); else MT_LAUNCH_WARP(1, 4, false); }
        else    { if (vis_bil) MT_LAUNCH_WARP(1, 1, true); else MT_LAUNCH_WARP(1, 1, false); }
    }
#undef MT_LAUNCH_WARP
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
    WarpBwdArgs a;
    a.x = x; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sf = x_sf; a.grid = grid;
    a.gout = gout; a.g_sb = g_sb; a.g_sc = g_sc; a.g_sf = g_sf; a.ggrid = ggrid;
    a.F = F; a.H = H; a.W = W; a.P = H * W;
    a.sp = make_sampler(H, W, (flags & MT_ALIGN_CORNERS) != 0);
    bool v4 = mult4(a.P) && aligned16(grid) && aligned16(ggrid) && aligned16(gout) && mult4(g_sb) &&
              mult4(g_sc) && mult4(g_sf);
    const int vec = v4 ? 4 : 1;
    dim3 block(256), gridd((a.P + 256 * vec - 1) / (256 * vec), B * F);
    cudaStream_t st = (cudaStream_t)stream;
    if (C == 3) { if (v4) warp_bwd_grid_kernel<3, 4><<<gridd, block, 0, st>>>(a); else warp_bwd_grid_kernel<3, 1><<<gridd, block, 0, st>>>(a); }
    else        { if (v4) warp_bwd_grid_kernel<1, 4><<<gridd, block, 0, st>>>(a); else warp_bwd_grid_kernel<1, 1><<<gridd, block, 0, st>>>(a); }
    return launch_status("mt_warp_bwd_grid");
}

static int fill_l1_args(WarpL1Args &a, const float *x, int64_t x_sb, int64_t x_sc, int64_t x_sf,
                        const float *flow, const float *x_target, int64_t xt_sb, int64_t xt_sc,
                        const float *v_target, int64_t vt_sb, int B, int F, int H, int W,
                        float weight, int flags, const char *who) {
    MT_REQUIRE(x && flow && x_target && v_target, "%s: NULL input", who);
    MT_REQUIRE(B > 0 && F > 0 && H > 0 && W > 0, "%s: empty shape", who);
    MT_REQUIRE((int64_t)H * W < (1ll << 30) && (int64_t)B * F <= 65535, "%s: too large", who);
    a.x = x; a.x_sb = x_sb; a.x_sc = x_sc; a.x_sf = x_sf; a.flow = flow;
    a.xt = x_target; a.xt_sb = xt_sb; a.xt_sc = xt_sc; a.vt = v_target; a.vt_sb = vt_sb;
    a.F = F; a.H = H; a.W = W; a.P = H * W; a.weight = weight;
    a.sp = make_sampler(H, W, (flags & MT_ALIGN_CORNERS) != 0);
    a.from_mask = (flags & MT_VIS_FROM_MASK) != 0;
    a.vis = nullptr; a.vis_sb = a.vis_sf = 0; a.x_al = a.v_al = nullptr; a.out3 = nullptr; a.ws = nullptr;
    a.out3_in = nullptr; a.grad_out = nullptr; a.gflow = nullptr; a.chunks = 0; a.total_chunks = 0;
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
    bool v4 = mult4(a.P) && aligned16(flow) && aligned16(x_target) && aligned16(v_target) &&
              mult4(xt_sb) && mult4(xt_sc) && mult4(vt_sb) && aligned16(x_aligned) && aligned16(v_aligned);
    const int vec = v4 ? 4 : 1;
    a.chunks = (a.P + 256 * vec - 1) / (256 * vec);
    a.total_chunks = (int64_t)B * F * a.chunks;
    int64_t want = (int64_t)sm_count() * 8;
    int nblk = (int)(a.total_chunks < want ? a.total_chunks : want);
    if (nblk > kMaxReduceBlocks) nblk = kMaxReduceBlocks;
    cudaStream_t st = (cudaStream_t)stream;
    if (v4) warp_l1_fwd_kernel<4><<<nblk, 256, 0, st>>>(a);
    else warp_l1_fwd_kernel<1><<<nblk, 256, 0, st>>>(a);
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
    a.out3_in = out3; a.grad_out = grad_out; a.gflow = gflow;
    bool v4 = mult4(a.P) && aligned16(flow) && aligned16(gflow) && aligned16(x_target) &&
              aligned16(v_target) && mult4(xt_sb) && mult4(xt_sc) && mult4(vt_sb);
    const int vec = v4 ? 4 : 1;
    dim3 block(256), gridd((a.P + 256 * vec - 1) / (256 * vec), B * F);
    cudaStream_t st = (cudaStream_t)stream;
    if (v4) warp_l1_bwd_kernel<4><<<gridd, block, 0, st>>>(a);
    else warp_l1_bwd_kernel<1><<<gridd, block, 0, st>>>(a);
    return launch_status("mt_warp_l1_bwd");
}

extern "C" int mt_mask_out(const float *flow, int64_t n, float *out, mt_stream_t stream) {
    MT_REQUIRE(flow && out && n > 0, "mt_mask_out: bad argument");
    MT_REQUIRE((reinterpret_cast<uintptr_t>(flow) & 7u) == 0, "mt_mask_out: flow must be 8 B aligned");
    int64_t nb = (n + 255) / 256, cap = (int64_t)sm_count() * 16;
    mask_out_kernel<<<(int)(nb < cap ? nb : cap), 256, 0, (cudaStream_t)stream>>>(flow, n, out);
    return launch_status("mt_mask_out");
}
