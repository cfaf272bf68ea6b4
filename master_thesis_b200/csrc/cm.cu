// cm.cu - K3: CPN context matching.
//
// Replaces CM_Module.forward + CM_Module.masked_softmax
//   master_thesis/model_cpn.py:206-254                                   (a8)
//
// The reference's "correlation" here is ONE masked global dot product per
// (sample, reference) - K = C*h*w = 524 288 terms, M = N = 1 - followed by a
// per-pixel softmax over the references and a weighted copy.  It is HBM-bound
// (AI ~ 0.6 flop/B), not a GEMM, so it stays on the SIMT pipes:
//   pass 0  cm_masks:  v' = bilinear_resize(v, (h,w)) > 0.5 for target + refs
//   pass 1  cm_sim:    partial sums of vt'*vr'*c_t*c_r (and of vt'*vr') per (b, r);
//                      reads c_feats exactly once from HBM (16 B loads)
//   pass 1b cm_weights: folds the partials into gs (fixed order, double) and
//                      computes the masked softmax over references ONCE per
//                      pixel -> weights (B,R,P), c_mask, c_mask channel of out
//   pass 2  cm_copy:   streams c_feats again (L2 where it still is: a sample is
//                      10.5 MB) and writes cat[c_t, sum_r c_r * w_r].
// Reductions are two-level and fixed-order (deterministic).  History (profiles/):
// a single-CTA-per-sample reduction kernel cost 14 us (latency chain); recomputing the
// softmax in every pass-2 CTA cost ~10 channels' worth of instructions per CTA; a
// per-sample ticket tail in pass 1 doubled pass 1 (fence + atomic + barrier per CTA).
#include <math.h>

#include "mt_common.cuh"

namespace mt {
namespace {

constexpr int kSimChannels = 2;   // default channels per CTA slab in pass 1 (swept on B200: profiles/)
constexpr int kCopyChannels = 4;  // default channels per CTA slab in pass 2 (template CC)
constexpr int kMaxRefs = 8;

struct CmArgs {
    const float *c_feats, *v_t, *v_al;
    float *out, *c_mask;
    float *masks;     // workspace: (B, f, P)   index 0 = target
    float *partials;  // workspace: (B, nparts, 2R): [dot_r ..., vsum_r ...]
    float *gs;        // workspace: (B, R)  (exported for tests)
    float *weights;   // workspace: (B, R, P) softmax weights over references
    int B, C, f, h, w, H, W, P, R, nparts, chunks, sim_ch, b_off;
};

// F.interpolate(bilinear, align_corners=False) source index (UpSample.h)
__device__ __forceinline__ void src_index(int dst, float scale, int in_size, int &i0, int &i1, float &l0,
                                          float &l1) {
    float s = __fsub_rn(__fmul_rn(scale, __fadd_rn((float)dst, 0.5f)), 0.5f);
    s = s < 0.0f ? 0.0f : s;
    i0 = (int)s;
    i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    l1 = __fsub_rn(s, (float)i0);
    l0 = __fsub_rn(1.0f, l1);
}

__global__ void __launch_bounds__(256) cm_masks_kernel(const CmArgs a) {
    pdl_sync();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y, b = blockIdx.z;  // j = 0: target, j >= 1: reference j-1
    if (p >= a.P) return;
    const float *src = (j == 0) ? a.v_t + (int64_t)b * a.H * a.W
                                : a.v_al + ((int64_t)b * a.R + (j - 1)) * a.H * a.W;
    const int y = p / a.w, x = p - y * a.w;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    src_index(y, __fdiv_rn((float)a.H, (float)a.h), a.H, y0, y1, ly0, ly1);
    src_index(x, __fdiv_rn((float)a.W, (float)a.w), a.W, x0, x1, lx0, lx1);
    const float v00 = __ldg(src + y0 * a.W + x0), v01 = __ldg(src + y0 * a.W + x1);
    const float v10 = __ldg(src + y1 * a.W + x0), v11 = __ldg(src + y1 * a.W + x1);
    const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
    const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
    const float val = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
    a.masks[((int64_t)b * a.f + j) * a.P + p] = val > 0.5f ? 1.0f : 0.0f;  // model_cpn.py:208-217
}

// gs[b, :] from the partials of sample b (fixed order, double).  Result in smem gs[R].
template <int R>
__device__ __forceinline__ void fold_gs(const CmArgs &a, int b, float *gs_smem) {
    __shared__ double dred[2 * R * 8];
    double v[2 * R];
#pragma unroll
    for (int r = 0; r < 2 * R; ++r) v[r] = 0.0;
    for (int i = threadIdx.x; i < a.nparts; i += blockDim.x) {
        const float *o = a.partials + ((int64_t)b * a.nparts + i) * (2 * R);
#pragma unroll
        for (int r = 0; r < 2 * R; ++r) v[r] += (double)__ldcg(o + r);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int r = 0; r < 2 * R; ++r) {
        v[r] = warp_sum(v[r]);
        if (lane == 0) dred[r * 8 + wid] = v[r];
    }
    __syncthreads();
    if (threadIdx.x < R) {
        const int r = threadIdx.x;
        double d = 0.0, vs = 0.0;
        for (int w = 0; w < nw; ++w) { d += dred[r * 8 + w]; vs += dred[(R + r) * 8 + w]; }
        const bool zero = vs < 1e-4;                                       // :222
        const float v_sum = (float)vs + (zero ? 1.0f : 0.0f);              // :223
        const float g = (float)d / (v_sum * (float)a.C);                   // :225-227
        gs_smem[r] = zero ? 0.0f : g;                                      // :228
    }
    __syncthreads();
}

// pass 1: grid (chunks, C / SC, B); thread = 4 pixels x SC channels
template <int R, int SC>
__global__ void __launch_bounds__(256) cm_sim_kernel(const CmArgs a) {
    pdl_sync();
    __shared__ float red[2 * R * 32];
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int slab = blockIdx.y, b = blockIdx.z + a.b_off;
    float acc[2 * R];  // [0, R): dot products, [R, 2R): sum of vt'*vr' (slab 0 only)
#pragma unroll
    for (int r = 0; r < 2 * R; ++r) acc[r] = 0.0f;
    if (p0 < a.P) {
        const float *mk = a.masks + (int64_t)b * a.f * a.P + p0;
        const float4 vt = *reinterpret_cast<const float4 *>(mk);
        float4 vm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float4 vr = *reinterpret_cast<const float4 *>(mk + (int64_t)(r + 1) * a.P);
            vm[r] = make_float4(vt.x * vr.x, vt.y * vr.y, vt.z * vr.z, vt.w * vr.w);  // :220
            if (slab == 0) acc[R + r] = (vm[r].x + vm[r].y) + (vm[r].z + vm[r].w);     // :221
        }
        const int c0 = slab * SC;
        float4 ct[SC], cr[SC][R];
#pragma unroll
        for (int k = 0; k < SC; ++k) {
            const int c = c0 + k;
            if (c < a.C) {
                const float *base = a.c_feats + ((int64_t)b * a.C + c) * a.f * a.P + p0;
                // default L2 policy (NOT evict-first): pass 2 re-reads these from L2
                ct[k] = __ldg(reinterpret_cast<const float4 *>(base));
#pragma unroll
                for (int r = 0; r < R; ++r)
                    cr[k][r] = __ldg(reinterpret_cast<const float4 *>(base + (int64_t)(r + 1) * a.P));
            }
        }
#pragma unroll
        for (int k = 0; k < SC; ++k) {
            if (c0 + k < a.C) {
#pragma unroll
                for (int r = 0; r < R; ++r) {  // vmap * c_t * c_r            :226
                    acc[r] += vm[r].x * ct[k].x * cr[k][r].x;
                    acc[r] += vm[r].y * ct[k].y * cr[k][r].y;
                    acc[r] += vm[r].z * ct[k].z * cr[k][r].z;
                    acc[r] += vm[r].w * ct[k].w * cr[k][r].w;
                }
            }
        }
    }
    block_sum<2 * R>(acc, red);
    if (threadIdx.x == 0) {
        float *o = a.partials + ((int64_t)b * a.nparts + slab * a.chunks + blockIdx.x) * (2 * R);
#pragma unroll
        for (int r = 0; r < 2 * R; ++r) o[r] = acc[r];
    }
}

// pass 1b: similarities -> per-pixel softmax weights, computed ONCE per pixel.
// grid (ceil(P / 1024), B), 256 threads, one 4-pixel group per thread.
template <int R>
__global__ void __launch_bounds__(256) cm_weights_kernel(const CmArgs a) {
    pdl_sync();
    __shared__ float gs_smem[R];
    const int b = blockIdx.y + a.b_off;
    fold_gs<R>(a, b, gs_smem);
    if (blockIdx.x == 0 && threadIdx.x < R) a.gs[(int64_t)b * R + threadIdx.x] = gs_smem[threadIdx.x];
    float gsr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) gsr[r] = gs_smem[r];
    const float *mk = a.masks + (int64_t)b * a.f * a.P;
    float *wout = a.weights + (int64_t)b * R * a.P;
    float *ocm = a.out + ((int64_t)b * (2 * a.C + 1) + 2 * a.C) * a.P;
    float *ocm2 = a.c_mask + (int64_t)b * a.P;
    // one 16 B group (4 pixels) per thread
    constexpr int kTail = 1;
    const int ngroups = a.P >> 2;
    for (int g0 = blockIdx.x * blockDim.x + threadIdx.x; g0 < ngroups; g0 += ngroups) {
        float4 vr4[kTail][R];
#pragma unroll
        for (int t = 0; t < kTail; ++t) {
            const int g = g0 + t * blockDim.x;
#pragma unroll
            for (int r = 0; r < R; ++r)
                vr4[t][r] = g < ngroups ? __ldcg(reinterpret_cast<const float4 *>(mk + (int64_t)(r + 1) * a.P) + g)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int t = 0; t < kTail; ++t) {
            const int g = g0 + t * blockDim.x;
            if (g >= ngroups) break;
            float w4[R][4], cm4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float vr[R];
#pragma unroll
                for (int r = 0; r < R; ++r)
                    vr[r] = i == 0 ? vr4[t][r].x : (i == 1 ? vr4[t][r].y : (i == 2 ? vr4[t][r].z : vr4[t][r].w));
                float mx = -INFINITY;  // masked_softmax over refs               :245-254
#pragma unroll
                for (int r = 0; r < R; ++r) mx = fmaxf(mx, __fmul_rn(gsr[r], vr[r]));
                float sum = 0.0f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    w4[r][i] = __fmul_rn(expf(__fsub_rn(__fmul_rn(gsr[r], vr[r]), mx)), vr[r]);
                    sum = __fadd_rn(sum, w4[r][i]);
                }
                if (sum < 1e-4f) sum = __fadd_rn(sum, 1.0f);
                float cm = 0.0f;
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    w4[r][i] = __fdiv_rn(w4[r][i], sum);
                    cm = __fadd_rn(cm, __fmul_rn(w4[r][i], vr[r]));              // :240
                }
                cm4[i] = __fsub_rn(1.0f, cm);                                    // :241
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
                reinterpret_cast<float4 *>(wout + (int64_t)r * a.P)[g] = make_float4(w4[r][0], w4[r][1], w4[r][2], w4[r][3]);
            const float4 c4 = make_float4(cm4[0], cm4[1], cm4[2], cm4[3]);
            reinterpret_cast<float4 *>(ocm)[g] = c4;
            reinterpret_cast<float4 *>(ocm2)[g] = c4;
        }
    }
}

// pass 2: grid (chunks, ceil(C / CC), B); thread = 4 pixels x CC channels.
template <int R, int CC>
__global__ void __launch_bounds__(256) cm_copy_kernel(const CmArgs a) {
    pdl_sync();
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p0 >= a.P) return;
    const int slab = blockIdx.y, b = blockIdx.z + a.b_off;
    const int c0 = slab * CC;
    float4 ct[CC], cr[CC][R];
#pragma unroll
    for (int k = 0; k < CC; ++k) {
        if (c0 + k < a.C) {
            const float *base = a.c_feats + ((int64_t)b * a.C + c0 + k) * a.f * a.P + p0;
            ct[k] = ld_stream4(base);
#pragma unroll
            for (int r = 0; r < R; ++r) cr[k][r] = ld_stream4(base + (int64_t)(r + 1) * a.P);
        }
    }
    float4 wg[R];
    const float *wp = a.weights + (int64_t)b * R * a.P + p0;
#pragma unroll
    for (int r = 0; r < R; ++r) wg[r] = __ldg(reinterpret_cast<const float4 *>(wp + (int64_t)r * a.P));
    float *ob = a.out + (int64_t)b * (2 * a.C + 1) * a.P + p0;
#pragma unroll
    for (int k = 0; k < CC; ++k) {
        const int c = c0 + k;
        if (c >= a.C) break;
        float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int r = 0; r < R; ++r) {  // sum_r c_r * w_r, sequential over r    :238
            o.x = __fadd_rn(o.x, __fmul_rn(cr[k][r].x, wg[r].x));
            o.y = __fadd_rn(o.y, __fmul_rn(cr[k][r].y, wg[r].y));
            o.z = __fadd_rn(o.z, __fmul_rn(cr[k][r].z, wg[r].z));
            o.w = __fadd_rn(o.w, __fmul_rn(cr[k][r].w, wg[r].w));
        }
        st_stream4(ob + (int64_t)c * a.P, ct[k]);                             // cat[c_t, ...]  :243
        st_stream4(ob + (int64_t)(a.C + c) * a.P, o);
    }
}

template <int R>
int launch_cm(CmArgs a, cudaStream_t st) {
    a.b_off = 0;
    dim3 g0((a.P + 255) / 256, a.f, a.B);
    launch(cm_masks_kernel, g0, 256, 0, st, a);
    // MT_CM_CHUNK > 0 processes the samples in chunks (sim -> weights -> copy per chunk) so that
    // pass 2 could re-read c_feats from L2.  Swept on B200 at B=8 (profiles/r1_sweep_cm.sh):
    // 1/2/4/8 samples per chunk -> 111/72/55/48 us: the extra small launches cost more than the
    // L2 reuse brings, so the default is the whole batch in one group.
    int chunk = tuning("MT_CM_CHUNK", 0);
    if (chunk < 1 || chunk > a.B) chunk = a.B;
    const int cc = tuning("MT_CM_COPY_CH", kCopyChannels);
    for (int b0 = 0; b0 < a.B; b0 += chunk) {
        const int nb = a.B - b0 < chunk ? a.B - b0 : chunk;
        a.b_off = b0;
        dim3 g1(a.chunks, (a.C + a.sim_ch - 1) / a.sim_ch, nb);
        if (a.sim_ch == 2) launch(cm_sim_kernel<R, 2>, g1, 256, 0, st, a);
        else launch(cm_sim_kernel<R, 4>, g1, 256, 0, st, a);
        dim3 gw(a.chunks, nb);
        launch(cm_weights_kernel<R>, gw, 256, 0, st, a);
        if (cc == 2) {
            dim3 g2(a.chunks, (a.C + 1) / 2, nb);
            launch(cm_copy_kernel<R, 2>, g2, 256, 0, st, a);
        } else {
            dim3 g2(a.chunks, (a.C + 3) / 4, nb);
            launch(cm_copy_kernel<R, 4>, g2, 256, 0, st, a);
        }
    }
    return launch_status("mt_cm_match_fwd");
}

int sim_channels() { return tuning("MT_CM_SIM_CH", kSimChannels) == 2 ? 2 : 4; }

int64_t align256(int64_t v) { return (v + 255) & ~int64_t(255); }

}  // namespace
}  // namespace mt

using namespace mt;

extern "C" int64_t mt_cm_workspace_bytes(int B, int C, int f, int h, int w) {
    if (B <= 0 || C <= 0 || f < 2 || h <= 0 || w <= 0) return 0;
    const int64_t P = (int64_t)h * w, R = f - 1;
    const int64_t chunks = (P + 1023) / 1024, nparts = chunks * ((C + 1) / 2);  // finest pass-1 split
    return align256(B * f * P * 4) + align256(B * nparts * 2 * R * 4) + align256(B * R * 4) +
           align256(B * R * P * 4);
}

extern "C" int mt_cm_match_fwd(const float *c_feats, const float *v_t, const float *v_aligned,
                               float *out, float *c_mask, void *workspace, int B, int C, int f,
                               int h, int w, int H, int W, mt_stream_t stream) {
    MT_REQUIRE(c_feats && v_t && v_aligned && out && c_mask && workspace, "mt_cm_match_fwd: NULL argument");
    MT_REQUIRE(B > 0 && C > 0 && h > 0 && w > 0 && H > 0 && W > 0, "mt_cm_match_fwd: empty shape");
    MT_REQUIRE(f >= 2 && f - 1 <= kMaxRefs, "mt_cm_match_fwd: needs 1..%d reference frames, got %d", kMaxRefs, f - 1);
    MT_REQUIRE(B <= 65535, "mt_cm_match_fwd: B > 65535");
    MT_REQUIRE(((int64_t)h * w) % 4 == 0, "mt_cm_match_fwd: h*w must be a multiple of 4 (got %dx%d)", h, w);
    MT_REQUIRE(aligned16(c_feats) && aligned16(out) && aligned16(c_mask) && aligned16(workspace),
               "mt_cm_match_fwd: pointers must be 16 B aligned");
    CmArgs a;
    a.c_feats = c_feats; a.v_t = v_t; a.v_al = v_aligned; a.out = out; a.c_mask = c_mask;
    a.B = B; a.C = C; a.f = f; a.h = h; a.w = w; a.H = H; a.W = W; a.P = h * w; a.R = f - 1;
    a.chunks = (a.P + 1023) / 1024;
    a.sim_ch = sim_channels();
    a.nparts = a.chunks * ((C + a.sim_ch - 1) / a.sim_ch);
    char *ws = reinterpret_cast<char *>(workspace);
    a.masks = reinterpret_cast<float *>(ws);
    ws += align256((int64_t)B * f * a.P * 4);
    a.partials = reinterpret_cast<float *>(ws);
    ws += align256((int64_t)B * a.chunks * ((C + 1) / 2) * 2 * a.R * 4);
    a.gs = reinterpret_cast<float *>(ws);
    ws += align256((int64_t)B * a.R * 4);
    a.weights = reinterpret_cast<float *>(ws);
    cudaStream_t st = (cudaStream_t)stream;
    switch (a.R) {
        case 1: return launch_cm<1>(a, st);
        case 2: return launch_cm<2>(a, st);
        case 3: return launch_cm<3>(a, st);
        case 4: return launch_cm<4>(a, st);
        case 5: return launch_cm<5>(a, st);
        case 6: return launch_cm<6>(a, st);
        case 7: return launch_cm<7>(a, st);
        default: return launch_cm<8>(a, st);
    }
}

// gs (B, f-1) as left in the workspace by the last mt_cm_match_fwd (for tests)
extern "C" const float *mt_cm_workspace_gs(const void *workspace, int B, int C, int f, int h, int w) {
    const int64_t P = (int64_t)h * w, R = f - 1;
    const int64_t chunks = (P + 1023) / 1024, nparts = chunks * ((C + 1) / 2);
    return reinterpret_cast<const float *>(reinterpret_cast<const char *>(workspace) +
                                           align256(B * f * P * 4) + align256(B * nparts * 2 * R * 4));
}
