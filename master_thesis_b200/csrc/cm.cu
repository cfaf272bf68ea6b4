// cm.cu - K3: CPN context matching.
//
// Replaces CM_Module.forward + CM_Module.masked_softmax
//   master_thesis/model_cpn.py:206-254                                   (a8)
//
// The reference's "correlation" here is ONE masked global dot product per
// (sample, reference) - K = C*h*w = 524 288 terms, M = N = 1 - followed by a
// per-pixel softmax over the references and a weighted copy.  It is HBM-bound
// (AI ~ 0.6 flop/B), not a GEMM, so it stays on the SIMT pipes.  Three launches:
//   pass 0  cm_masks:  v' = bilinear_resize(v, (h,w)) > 0.5 for target + refs, as floats and as
//                      one byte per pixel (bit 0 target, bit r+1 reference r)
//   pass 1  cm_sim:    partial sums of vt'*vr'*c_t*c_r (and of vt'*vr') per (b, r);
//                      reads c_feats exactly once from HBM (16 B loads); its feature loads are
//                      issued before griddepcontrol.wait, i.e. while cm_masks is still running
//   pass 2  cm_copy2:  streams c_feats again (reverse sample order: the tail of the batch is what
//                      pass 1 left in L2) and writes cat[c_t, sum_r c_r * w_r, c_mask]; every CTA
//                      folds the partials of its sample into gs (fixed order, double) and evaluates
//                      the masked softmax once per MASK PATTERN (2^R entries), weights per pixel are
//                      a lookup by the mask byte.  (With 8 references, or MT_CM_TABLE=0: the separate
//                      cm_weights kernel - softmax once per pixel -> weights (B,R,P) - and cm_copy.)
// Reductions are two-level and fixed-order (deterministic).  History (profiles/):
// a single-CTA-per-sample reduction kernel cost 14 us (latency chain); recomputing the per-PIXEL
// softmax in every pass-2 CTA cost ~10 channels' worth of instructions per CTA; a per-sample
// ticket tail in pass 1 doubled pass 1 (fence + atomic + barrier per CTA); per-group launches
// (MT_CM_CHUNK) and a persistent pipelined single launch (K3p below) are slower than this.
#include <math.h>

#include "mt_common.cuh"
#include "mt_tma.cuh"

namespace mt {
namespace {

constexpr int kSimChannels = 4;   // default channels per CTA slab in pass 1 (swept on B200: profiles/)
constexpr int kCopyChannels = 4;  // default channels per CTA slab in pass 2 (template CC)
constexpr int kMaxRefs = 8;

// exact unsigned division by a runtime constant (round-up method): q = x / d for every 32-bit x
struct FastDiv { uint32_t d, m, s1, s2; };
static inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t l = 0;
    while ((1ull << l) < d) ++l;
    f.d = d;
    f.m = (uint32_t)((((1ull << l) - d) << 32) / d + 1);
    f.s1 = l < 1 ? l : 1;
    f.s2 = l > 0 ? l - 1 : 0;
    return f;
}
__device__ __forceinline__ uint32_t fast_div(uint32_t x, const FastDiv &f) {
    const uint32_t t = __umulhi(f.m, x);
    return (t + ((x - t) >> f.s1)) >> f.s2;
}

struct CmArgs {
    const float *c_feats, *v_t, *v_al;
    float *out, *c_mask;
    float *masks;     // workspace: (B, f, P)   index 0 = target
    float *partials;  // workspace: (B, nparts, 2R): [dot_r ..., vsum_r ...]
    float *gs;        // workspace: (B, R)  (exported for tests)
    float *weights;   // workspace: (B, R, P) softmax weights over references
    int B, C, f, h, w, H, W, P, R, nparts, chunks, sim_ch, b_off;
    // fused path: per-sample count of finished similarity items (zeroed by cm_masks), items per
    // sample and pass, and the distance (in samples) between pass 1 and pass 2 of the same sample
    unsigned int *count, *flag;  // (B) finished S items / table published
    unsigned char *pmask;        // (B, P) bit 0: vt', bit r + 1: vr' of reference r
    float *table;                // (B, 2^R, R + 1) softmax weights and c_mask per mask pattern
    int n_items, lag, copy_reverse;
    int workers, rounds, head;  // schedule of the pipelined kernel (cm_decode)
    int n_copy, b_sim, n_sim;   // cm_copy_sim_kernel: copies samples [b_off, +n_copy), similarities of [b_sim, +n_sim)
    FastDiv dv_items, dv_chunks;
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

// pass 0: grid (ceil(P / 256), B); thread = one low-resolution pixel, all f masks.
// Also resets the per-sample state of the pipelined kernel (count, flag, table = NaN).
__global__ void __launch_bounds__(256) cm_masks_kernel(const CmArgs a) {
    pdl_sync();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (blockIdx.x == 0) {
        const int tabf = (1 << a.R) * (a.R + 1);
        for (int q = threadIdx.x; q < tabf; q += blockDim.x) a.table[(int64_t)b * tabf + q] = __int_as_float(0x7fc00000);
        if (threadIdx.x == 0) { a.count[b] = 0u; a.flag[b] = 0u; }
    }
    if (p >= a.P) return;
    const int y = p / a.w, x = p - y * a.w;
    int y0, y1, x0, x1;
    float ly0, ly1, lx0, lx1;
    src_index(y, __fdiv_rn((float)a.H, (float)a.h), a.H, y0, y1, ly0, ly1);
    src_index(x, __fdiv_rn((float)a.W, (float)a.w), a.W, x0, x1, lx0, lx1);
    unsigned int bits = 0u;
    for (int j = 0; j < a.f; ++j) {  // j = 0: target, j >= 1: reference j - 1
        const float *src = (j == 0) ? a.v_t + (int64_t)b * a.H * a.W
                                    : a.v_al + ((int64_t)b * a.R + (j - 1)) * a.H * a.W;
        const float v00 = __ldg(src + y0 * a.W + x0), v01 = __ldg(src + y0 * a.W + x1);
        const float v10 = __ldg(src + y1 * a.W + x0), v11 = __ldg(src + y1 * a.W + x1);
        const float top = __fadd_rn(__fmul_rn(lx0, v00), __fmul_rn(lx1, v01));
        const float bot = __fadd_rn(__fmul_rn(lx0, v10), __fmul_rn(lx1, v11));
        const float val = __fadd_rn(__fmul_rn(ly0, top), __fmul_rn(ly1, bot));
        const bool on = val > 0.5f;                                            // model_cpn.py:208-217
        a.masks[((int64_t)b * a.f + j) * a.P + p] = on ? 1.0f : 0.0f;
        bits |= on ? (1u << j) : 0u;
    }
    a.pmask[(int64_t)b * a.P + p] = (unsigned char)bits;
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

// pass 1 for one (1024-pixel chunk, SC-channel slab, sample): thread = 4 pixels x SC channels
template <int R, int SC, bool WAIT>
__device__ __forceinline__ void cm_sim_body(const CmArgs &a, int slab, int b, float *red) {
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    float acc[2 * R];  // [0, R): dot products, [R, 2R): sum of vt'*vr' (slab 0 only)
#pragma unroll
    for (int r = 0; r < 2 * R; ++r) acc[r] = 0.0f;
    const int c0 = slab * SC;
    float4 ct[SC], cr[SC][R];
    if (p0 < a.P) {
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
    }
    // WAIT: the features above were requested while the previous kernel of the stream (cm_masks) was
    // still running - they do not depend on it; the masks below do (pdl_wait)
    if (WAIT) pdl_wait();
    if (p0 < a.P) {
        float4 vm[R];
        if (R <= 7) {
            // the masks of 4 pixels as 4 bytes (bit 0 target, bit r + 1 reference r) instead of R + 1 float4
            const uint32_t mw = __ldcg(reinterpret_cast<const uint32_t *>(a.pmask + (int64_t)b * a.P + p0));
#pragma unroll
            for (int r = 0; r < R; ++r) {  // vt' * vr'                  :220
                const uint32_t m = mw & (mw >> (r + 1)) & 0x01010101u;
                vm[r] = make_float4((float)(m & 1u), (float)((m >> 8) & 1u), (float)((m >> 16) & 1u),
                                    (float)((m >> 24) & 1u));
            }
        } else {
            const float *mk = a.masks + (int64_t)b * a.f * a.P + p0;
            const float4 vt = __ldcg(reinterpret_cast<const float4 *>(mk));
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 vr = __ldcg(reinterpret_cast<const float4 *>(mk + (int64_t)(r + 1) * a.P));
                vm[r] = make_float4(vt.x * vr.x, vt.y * vr.y, vt.z * vr.z, vt.w * vr.w);  // :220
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (slab == 0) acc[R + r] = (vm[r].x + vm[r].y) + (vm[r].z + vm[r].w);     // :221
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

// pass 1: grid (chunks, C / SC, B)
template <int R, int SC>
__global__ void __launch_bounds__(256) cm_sim_kernel(const CmArgs a) {
    // Scheduled while cm_masks_kernel is still running (that kernel waited for ITS predecessor before it
    // let this one start, so c_feats is complete): the feature loads overlap it, the wait sits in the body.
    pdl_launch();
    __shared__ float red[2 * R * 32];
    cm_sim_body<R, SC, true>(a, (int)blockIdx.y, (int)blockIdx.z + a.b_off, red);
}

// pass 1b: similarities -> per-pixel softmax weights, computed ONCE per pixel.
// grid (ceil(P / 1024), B), 256 threads, one 4-pixel group per thread.
template <int R>
__global__ void __launch_bounds__(256) cm_weights_kernel(const CmArgs a) {
    // launch_dependents BEFORE the wait: the CTAs of cm_copy_kernel are scheduled into the slots that the
    // last wave of pass 1 frees and request their features (old data: c_feats) while pass 1 drains and
    // this kernel runs; they wait for this kernel before they touch the weights.
    pdl_launch();
    pdl_wait();
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

// pass 2 for one (1024-pixel chunk, CC-channel slab, sample): thread = 4 pixels x CC channels
template <int R, int CC, bool WAIT>
__device__ __forceinline__ void cm_copy_body(const CmArgs &a, int slab, int b) {
    const int p0 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (p0 >= a.P) {
        if (WAIT) pdl_wait();
        return;
    }
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
    // WAIT: the features above were requested while cm_weights_kernel was still running
    if (WAIT) pdl_wait();
    float4 wg[R];
    const float *wp = a.weights + (int64_t)b * R * a.P + p0;
#pragma unroll
    for (int r = 0; r < R; ++r) wg[r] = __ldcg(reinterpret_cast<const float4 *>(wp + (int64_t)r * a.P));
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

// pass 2: grid (chunks, ceil(C / CC), B)
template <int R, int CC>
__global__ void __launch_bounds__(256) cm_copy_kernel(const CmArgs a) {
    pdl_launch();
    // samples in REVERSE order: pass 1 streamed them 0 .. B-1, so the tail of the batch is what is
    // still in L2 (126 MB); walking forwards again evicts it just before it is needed (LRU: ncu
    // 4.6 % hit rate at B=8)
    const int b = (a.copy_reverse ? (int)gridDim.z - 1 - (int)blockIdx.z : (int)blockIdx.z) + a.b_off;
    cm_copy_body<R, CC, true>(a, (int)blockIdx.y, b);
}

// pass 1b + 2 in one launch (the default): every CTA folds the partials of its sample itself (4 KB from
// L2, one double per thread, fixed order) and evaluates the masked softmax once per MASK PATTERN - vr' is
// 0/1, so a sample has only 2^R distinct weight vectors; same operations in the same order as
// cm_weights_kernel, hence the same bits - while its feature loads, issued before the wait, are still in
// flight.  Per pixel the weights are a table lookup by the mask byte.  This removes cm_weights_kernel
// from the chain (7 us as a launch of 32 latency-bound CTAs), the weights array (B,R,P) and its reads.
template <int R, int CC>
__global__ void __launch_bounds__(256, 2) cm_copy2_kernel(const CmArgs a) {
    constexpr int G2 = 2 * R, TABF = (1 << R) * (R + 1);
    __shared__ double dred[256];
    __shared__ float gs_smem[R];
    __shared__ float tab[TABF];
    pdl_launch();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int p0 = (blockIdx.x * blockDim.x + tid) * 4;
    const bool live = p0 < a.P;
    const int slab = blockIdx.y, b = (a.copy_reverse ? (int)gridDim.z - 1 - (int)blockIdx.z : (int)blockIdx.z) + a.b_off;
    const int c0 = slab * CC;
    float4 ct[CC], cr[CC][R];
    if (live) {
#pragma unroll
        for (int k = 0; k < CC; ++k) {
            if (c0 + k < a.C) {
                const float *base = a.c_feats + ((int64_t)b * a.C + c0 + k) * a.f * a.P + p0;
                ct[k] = ld_stream4(base);
#pragma unroll
                for (int r = 0; r < R; ++r) cr[k][r] = ld_stream4(base + (int64_t)(r + 1) * a.P);
            }
        }
    }
    pdl_wait();  // pass 1 (and cm_masks before it) complete: partials and mask bytes are valid
    const uint32_t mw = live ? __ldcg(reinterpret_cast<const uint32_t *>(a.pmask + (int64_t)b * a.P + p0)) : 0u;
    // ---- gs[b, :]: thread t adds element t % 2R of rows t / 2R, t / 2R + kSlots, ... in increasing order ----
    constexpr int kSlots = 256 / G2;
    double v = 0.0;
    if (tid < kSlots * G2) {
        const int r = tid % G2, i0 = tid / G2;
        const float *o = a.partials + (int64_t)b * a.nparts * G2 + r;
#pragma unroll 4
        for (int i = i0; i < a.nparts; i += kSlots) v += (double)__ldcg(o + (int64_t)i * G2);
    }
    if constexpr ((G2 & (G2 - 1)) == 0 && G2 <= 16) {
        // the lanes l, l + 2R, l + 4R, ... of a warp hold the same element: xor tree, then 8 warps
#pragma unroll
        for (int o = G2; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < G2) dred[wid * G2 + lane] = v;
        __syncthreads();
        if (tid < R) {
            double d = 0.0, vs = 0.0;
#pragma unroll
            for (int w8 = 0; w8 < 8; ++w8) { d += dred[w8 * G2 + tid]; vs += dred[w8 * G2 + R + tid]; }
            const bool zero = vs < 1e-4;                                       // :222
            const float v_sum = (float)vs + (zero ? 1.0f : 0.0f);              // :223
            const float g = (float)d / (v_sum * (float)a.C);                   // :225-227
            gs_smem[tid] = zero ? 0.0f : g;                                    // :228
        }
    } else {
        dred[tid] = v;
        __syncthreads();
        if (tid < R) {
            double d = 0.0, vs = 0.0;
            for (int sl = 0; sl < kSlots; ++sl) { d += dred[sl * G2 + tid]; vs += dred[sl * G2 + R + tid]; }
            const bool zero = vs < 1e-4;
            const float v_sum = (float)vs + (zero ? 1.0f : 0.0f);
            const float g = (float)d / (v_sum * (float)a.C);
            gs_smem[tid] = zero ? 0.0f : g;
        }
    }
    __syncthreads();
    // ---- masked_softmax over the references, once per mask pattern t                :245-254 ----
    for (int t = tid; t < (1 << R); t += 256) {
        float vr[R], wv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) vr[r] = ((t >> r) & 1) ? 1.0f : 0.0f;
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < R; ++r) mx = fmaxf(mx, __fmul_rn(gs_smem[r], vr[r]));
        float sum = 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            wv[r] = __fmul_rn(expf(__fsub_rn(__fmul_rn(gs_smem[r], vr[r]), mx)), vr[r]);
            sum = __fadd_rn(sum, wv[r]);
        }
        if (sum < 1e-4f) sum = __fadd_rn(sum, 1.0f);
        float cm = 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            wv[r] = __fdiv_rn(wv[r], sum);
            cm = __fadd_rn(cm, __fmul_rn(wv[r], vr[r]));                          // :240
            tab[t * (R + 1) + r] = wv[r];
        }
        tab[t * (R + 1) + R] = __fsub_rn(1.0f, cm);                               // :241
    }
    __syncthreads();
    if (slab == 0 && blockIdx.x == 0 && tid < R) a.gs[(int64_t)b * R + tid] = gs_smem[tid];
    if (!live) return;
    int pat[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pat[j] = (int)((mw >> (8 * j + 1)) & ((1u << R) - 1u)) * (R + 1);
    float *ob = a.out + (int64_t)b * (2 * a.C + 1) * a.P + p0;
    if (slab == 0) {
        const float4 c4 = make_float4(tab[pat[0] + R], tab[pat[1] + R], tab[pat[2] + R], tab[pat[3] + R]);
        st_stream4(ob + (int64_t)(2 * a.C) * a.P, c4);
        st_stream4(a.c_mask + (int64_t)b * a.P + p0, c4);
    }
    float4 wg[R];
#pragma unroll
    for (int r = 0; r < R; ++r) wg[r] = make_float4(tab[pat[0] + r], tab[pat[1] + r], tab[pat[2] + r], tab[pat[3] + r]);
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

// pass 2 of one group of samples and pass 1 of the NEXT group in the same launch (slabs interleaved
// along blockIdx.y): the copy re-reads its group from L2, where pass 1 left it one launch ago, while
// the similarity of the next group streams from HBM - so c_feats crosses HBM once, without any
// synchronisation inside a kernel.  grid (chunks, 2 * ceil(C / CH), max(n_copy, n_sim)).
template <int R, int CH>
__global__ void __launch_bounds__(256) cm_copy_sim_kernel(const CmArgs a) {
    pdl_sync();
    __shared__ float red[2 * R * 32];
    const int slab = (int)blockIdx.y >> 1, z = (int)blockIdx.z;
    if (blockIdx.y & 1) {
        if (z < a.n_sim) cm_sim_body<R, CH, false>(a, slab, a.b_sim + z, red);
    } else {
        if (z < a.n_copy) cm_copy_body<R, CH, false>(a, slab, a.b_off + z);
    }
}


// ---------------------------------------------------------------------------------------------
// K3p (EXPERIMENTAL, MT_CM_FUSED=1, off by default): pass 1 + 1b + 2 as ONE persistent,
// software-pipelined launch in which pass 2 reads c_feats from L2.  Parity-green, but 1.3-1.4x slower
// than the three launches on B200 (55 vs 40 us at B=8, 175 vs 132 us at B=32); kept as the record of that design
// (measurements and the reasons: profiles/r1_experiments.md).
//
// The two passes over c_feats are inherent (the similarity is a global reduction over the sample),
// but as separate launches over the whole batch the second pass misses L2 (ncu: 4.6 % hit rate at
// B=8, 84 MB) and the op moves 84+10+84+34 MB through HBM for 128 MB algorithmic.  Here one CTA per SM
// (two groups of 8 compute warps = two workers, a publisher warp, a producer warp) walks a common
// round schedule (cm_decode): S rounds (similarity partial of a 1024-pixel x CH-channel slab) run
// `head` rounds ahead of the C rounds (weighted copy of such a slab), so a sample (10.5 MB) is
// re-read from L2 a few rounds after it was streamed from HBM.
//   * Memory pipeline: the producer warp fetches the operands of the next NST items with
//     cp.async.bulk (4 KB per slab and frame) into a shared-memory ring; slot layout = 16 B per
//     thread, so the compute warps read conflict-free LDS.128; full / empty mbarriers per stage.
//   * Masks travel as one byte per pixel (bit 0 target, bit r+1 reference r; cm_masks_kernel).
//   * S items hand their 2R sums to the publisher warp through a shared-memory mailbox; the
//     publisher stores the partial, fences and bumps the per-sample counter (off the compute warps'
//     path: with warp 0 publishing, every item cost 3-5 us).  The publisher that finishes the LAST
//     partial of a sample folds them in fixed order in double, evaluates the masked softmax once per
//     MASK PATTERN (vr' is 0/1: 2^R distinct weight vectors per sample; the same operations in the
//     same order as cm_weights_kernel: same bits) and publishes that table + a flag.
//   * C items read the table one item ahead through a register; cm_masks_kernel cleared it to NaN, so
//     a copy without NaN is complete (every word is written once); otherwise the group waits for the
//     sample's flag (acquire) and reloads.  Weights per pixel are a lookup by the mask byte.
// Progress: every S item of a sample precedes every C item of it in the common round sequence (the
// host picks `head` accordingly), S items never wait, and the grid is one resident wave.  A wait that
// does not end within 2 s traps (the launch fails loudly instead of hanging).
__device__ __forceinline__ unsigned int ld_acquire(const unsigned int *p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned int *p, unsigned int v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr int kMailbox = 16;

template <int R>
constexpr int cm_table_floats() { return (1 << R) * (R + 1); }  // per pattern: R weights, c_mask

struct CmItem { int b, idx, chunk, slab; bool copy, valid; };

#ifdef MT_DEV_PROBES
// developer probe (tools/dbg_cm.py): per CTA [0] kernel ns, [1] ns waiting for cp.async data, [2] ns in S
// items, [3] ns in C items, [4] C items through the slow path, [5] items, [6] publisher busy ns, [7] publishes
__device__ unsigned long long g_cm_probe[256 * 8];
__device__ unsigned long long g_cm_probe2[2048];  // [0] slow items whose flag was already set, [1] ns in the slow path, [2] ns in decode
#define CM_PROBE(...) __VA_ARGS__
#else
#define CM_PROBE(...)
#endif

// Schedule.  The S items of all samples form one stream (position u = b * n_items + idx), the C items
// another.  Every worker (a compute group of a CTA; NW workers in all) runs the SAME sequence of
// rounds and takes position round * NW + worker of the round's stream: `head` S rounds, then C and S
// rounds alternating, then the remaining C rounds.  Every worker therefore alternates between reads
// from HBM (S) and reads from L2 + writes (C) - a first version that interleaved the two streams in
// one list gave the even CTAs only S items and the odd ones only C items (2x slower) - and every S
// item of a sample is earlier in every worker's sequence than any C item of that sample (the host
// picks `head` accordingly), which is what makes waiting inside a C item safe.
__device__ __forceinline__ CmItem cm_decode(const CmArgs &a, int worker, int r) {
    CmItem d;
    const int R1 = a.rounds;  // rounds per stream
    int sr;                   // round within the stream
    if (r < a.head) { d.copy = false; sr = r; }
    else {
        const int t = r - a.head, pairs = R1 - a.head;  // alternating part: C first
        if (t < 2 * pairs) { d.copy = (t & 1) == 0; sr = d.copy ? (t >> 1) : a.head + (t >> 1); }
        else { d.copy = true; sr = pairs + (t - 2 * pairs); }
    }
    const int u = sr * a.workers + worker;
    d.valid = r < 2 * R1 && u < a.B * a.n_items;
    d.b = 0; d.idx = 0; d.chunk = 0; d.slab = 0;
    if (d.valid) {
        d.b = (int)fast_div((uint32_t)u, a.dv_items);
        d.idx = u - d.b * a.n_items;
        d.slab = (int)fast_div((uint32_t)d.idx, a.dv_chunks);
        d.chunk = d.idx - d.slab * a.chunks;
    }
    return d;
}
// a worker is done after round 2 * rounds - 1; rounds whose position is past the end of the stream are empty
__device__ __forceinline__ bool cm_done(const CmArgs &a, int r) { return r >= 2 * a.rounds; }

// executed by ONE warp: partials of sample b -> gs -> softmax table -> flag
template <int R>
__device__ __forceinline__ void cm_publish_table(const CmArgs &a, int b) {
    constexpr int G2 = 2 * R, TABF = cm_table_floats<R>();
    const int lane = threadIdx.x & 31;
    // lane l adds rows l, l + 32, ... in increasing order (4 rows of loads in flight), then an xor tree
    // over the lanes: fixed order, double
    double v[G2];
#pragma unroll
    for (int r = 0; r < G2; ++r) v[r] = 0.0;
    const float *rows = a.partials + (int64_t)b * a.nparts * G2;
    for (int i0 = lane; i0 < a.nparts; i0 += 4 * 32) {
        float t[4][G2];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u;
#pragma unroll
            for (int r = 0; r < G2; ++r) t[u][r] = i < a.nparts ? __ldcg(rows + (int64_t)i * G2 + r) : 0.0f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int r = 0; r < G2; ++r) v[r] += (double)t[u][r];
    }
#pragma unroll
    for (int r = 0; r < G2; ++r) v[r] = warp_sum(v[r]);
    float gs[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const double d = v[r], vs = v[R + r];
        const bool zero = vs < 1e-4;                                       // :222
        const float v_sum = (float)vs + (zero ? 1.0f : 0.0f);              // :223
        const float g0 = (float)d / (v_sum * (float)a.C);                  // :225-227
        gs[r] = zero ? 0.0f : g0;                                          // :228
        if (lane == 0) a.gs[(int64_t)b * R + r] = gs[r];
    }
    float *tab = a.table + (int64_t)b * TABF;
    for (int t = lane; t < (1 << R); t += 32) {
        // masked_softmax over the references for mask pattern t                 :245-254
        float vr[R], wv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) vr[r] = ((t >> r) & 1) ? 1.0f : 0.0f;
        float mx = -INFINITY;
#pragma unroll
        for (int r = 0; r < R; ++r) mx = fmaxf(mx, __fmul_rn(gs[r], vr[r]));
        float sum = 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            wv[r] = __fmul_rn(expf(__fsub_rn(__fmul_rn(gs[r], vr[r]), mx)), vr[r]);
            sum = __fadd_rn(sum, wv[r]);
        }
        if (sum < 1e-4f) sum = __fadd_rn(sum, 1.0f);
        float cm = 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            wv[r] = __fdiv_rn(wv[r], sum);
            cm = __fadd_rn(cm, __fmul_rn(wv[r], vr[r]));                          // :240
            __stcg(tab + t * (R + 1) + r, wv[r]);
        }
        __stcg(tab + t * (R + 1) + R, __fsub_rn(1.0f, cm));                       // :241
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(a.flag + b, 1u);
    CM_PROBE(if (lane == 0 && b < 64) g_cm_probe2[1024 + b] = global_ns();)
}

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

template <int R, int CH>
constexpr int cm_stage_bytes() { return CH * (R + 1) * 4096 + 1024; }  // CH x f slabs of 1024 px + 1024 mask bytes

// 8 values per lane -> lane l holds the warp total of value (l >> 2) & 7: 9 shuffles instead of 40
__device__ __forceinline__ float warp_sum8(const float (&v)[8]) {
    const int lane = threadIdx.x & 31;
    const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4;
    float w[4], x[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = h4 ? v[i] : v[i + 4], keep = h4 ? v[i + 4] : v[i];
        w[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = h3 ? w[i] : w[i + 2], keep = h3 ? w[i + 2] : w[i];
        x[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
    const float send = h2 ? x[0] : x[1], keep = h2 ? x[1] : x[0];
    float y = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    return y;
}
__device__ __forceinline__ void bar_group(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ bool bar_group_or(int id, bool p) {
    int r;
    asm volatile(
        "{\n\t.reg .pred pin, pout;\n\tsetp.ne.u32 pin, %2, 0;\n\tbar.red.or.pred pout, %1, 256, pin;\n\t"
        "selp.u32 %0, 1, 0, pout;\n\t}"
        : "=r"(r) : "r"(id), "r"((int)p) : "memory");
    return r != 0;
}

// NG groups of 8 compute warps (each group works on its own item: 4 warps per scheduler hide the
// LDS / FP latencies that 2 could not), one publisher warp, one producer warp.
template <int R, int CH, int NST, int NG>
__global__ void __launch_bounds__(NG * 256 + 64, 1) cm_pipe_kernel(const __grid_constant__ CUtensorMap map_c,
                                                                   const __grid_constant__ CmArgs a) {
    constexpr int TABF = cm_table_floats<R>();
    constexpr int kStageBytes = cm_stage_bytes<R, CH>();
    constexpr int G2 = 2 * R;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    float *tabs = reinterpret_cast<float *>(smem_raw + NST * kStageBytes);  // NG x 2 tables
    __shared__ float red[NG][2][G2 * 8];
    __shared__ uint64_t full[NST], empty[NST];
    // S items hand their sums to the publisher warp through these mailboxes: the global publication
    // (store, fence, returning atomic: 2-3 us of latency) is off the compute warps' path.  With the
    // publication done by warp 0 itself every S item cost 3-5 us (the next item's barrier waited).
    __shared__ float mbox[NG][kMailbox][G2];
    __shared__ int mbox_b[NG][kMailbox], mbox_idx[NG][kMailbox];
    __shared__ volatile int mb_ready[NG], mb_done[NG], mb_fin[NG];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) {
        for (int g = 0; g < NG; ++g) { mb_ready[g] = 0; mb_done[g] = 0; mb_fin[g] = 0; }
        for (int s = 0; s < NST; ++s) {
            mbar_init(smem_u32(full + s), 1);
            mbar_init(smem_u32(empty + s), 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_sync();
    CM_PROBE(unsigned long long pr_t0 = global_ns(); unsigned long long pr_wait = 0, pr_s = 0, pr_c = 0, pr_slow = 0, pr_n = 0;)
    CM_PROBE(if (tid == 0 && blockIdx.x < 256) { g_cm_probe2[blockIdx.x * 8 + 0] = 0; g_cm_probe2[blockIdx.x * 8 + 1] = 0; })
    CM_PROBE(if (tid == 0 && blockIdx.x == 0) g_cm_probe2[1200] = pr_t0;)
    if (wid == 8 * NG + 1) {
        // ===================== producer: bulk copies of the operands, NST items ahead =====================
        if (lane == 0) {
            int n = 0;  // valid items of this CTA so far (stage ring position)
            for (int r = 0; !cm_done(a, r); ++r)
            for (int g = 0; g < NG; ++g) {
                const CmItem d = cm_decode(a, blockIdx.x + g * gridDim.x, r);
                if (!d.valid) continue;
                const int s = n % NST;
                const uint32_t ph = (uint32_t)(n / NST) & 1u;
                ++n;
                mbar_wait(smem_u32(empty + s), ph ^ 1u);
                const int c0 = d.slab * CH;
                const int px = min(1024, a.P - d.chunk * 1024);
                const uint32_t fb = smem_u32(full + s), dst = smem_u32(smem_raw + s * kStageBytes);
                // ONE tensor load for the CH x (R + 1) slabs of 1024 pixels: c_feats as (256 px, P / 256, B*C*f
                // rows), box (256, 4, CH * (R + 1)).  As 4 KB bulk copies (one per slab and frame) an item took
                // ~3 us to arrive: ~0.3 us per copy, issued one after the other.  Rows past the tensor and
                // pixels past P are zero-filled and count towards the transaction bytes.
                mbar_expect_tx(fb, (uint32_t)(CH * (R + 1) * 4096 + px));
                tma_load_3d(dst, &map_c, fb, 0, d.chunk * 4, (d.b * a.C + c0) * a.f);
                bulk_load(dst + CH * (R + 1) * 4096, a.pmask + (int64_t)d.b * a.P + d.chunk * 1024, (uint32_t)px, fb);
            }
        }
        return;
    }
    if (wid == 8 * NG) {
        // ===================== publisher warp =====================
        // Batched: every ready mailbox entry of both groups is taken by its own lane - partial rows
        // stored, ONE fence, then the counter atomics side by side.  One entry at a time cost 1.4 us each
        // (fence + returning atomic), more than the compute warps need to produce one: the mailboxes ran
        // full and a sample's table appeared 20+ us after its last item.
        static_assert(NG * kMailbox <= 32, "one lane per mailbox entry");
        int done[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) done[g] = 0;
        CM_PROBE(unsigned long long pb_busy = 0, pb_n = 0;)
        for (;;) {
            int rdy[NG], total = 0;
            bool fin = true;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                rdy[g] = mb_ready[g];
                total += rdy[g] - done[g];
                fin = fin && mb_fin[g] && mb_ready[g] == done[g];
            }
            if (total == 0) {
                if (fin) break;
                __nanosleep(100);
                continue;
            }
            __threadfence_block();
            CM_PROBE(unsigned long long pb_t = global_ns();)
            // lane e < total: entry e, group by group
            int eg = -1, eslot = 0, e = lane;
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const int cnt = rdy[g] - done[g];
                if (eg < 0 && e < cnt) { eg = g; eslot = (done[g] + e) % kMailbox; }
                if (eg < 0) e -= cnt;
            }
            int b = 0;
            if (eg >= 0) {
                b = mbox_b[eg][eslot];
                float *o = a.partials + ((int64_t)b * a.nparts + mbox_idx[eg][eslot]) * G2;
#pragma unroll
                for (int r = 0; r < G2; ++r) __stcg(o + r, mbox[eg][eslot][r]);
            }
            __threadfence();  // release: the partials before the counts
            __syncwarp();
            if (lane == 0) {
#pragma unroll
                for (int g = 0; g < NG; ++g) mb_done[g] = rdy[g];  // the mailbox slots are free again
            }
            bool last = false;
            if (eg >= 0) last = atomicAdd(a.count + b, 1u) == (unsigned int)a.n_items - 1u;
            unsigned int lm = __ballot_sync(0xffffffffu, last);
            while (lm) {
                const int src = __ffs(lm) - 1;
                lm &= lm - 1;
                const int bl = __shfl_sync(0xffffffffu, b, src);
                __threadfence();  // acquire: every other item's partials
                cm_publish_table<R>(a, bl);
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) done[g] = rdy[g];
            CM_PROBE(pb_busy += global_ns() - pb_t; pb_n += total;)
        }
        CM_PROBE(if (lane == 0 && blockIdx.x < 256) { g_cm_probe[blockIdx.x * 8 + 6] = pb_busy; g_cm_probe[blockIdx.x * 8 + 7] = pb_n; })
        return;
    }

    // ===================== compute warps =====================
    const int grp = wid >> 3, lt = tid & 255, lw = wid & 7, bar_id = 1 + grp;
    float *gtabs = tabs + grp * 2 * TABF;
    // the softmax table of the group's NEXT item travels through a register (one 16 B chunk per thread)
    float4 tnext = make_float4(0.f, 0.f, 0.f, 0.f);
    const int me = blockIdx.x + grp * gridDim.x;
    for (int r = 0; !cm_done(a, r); ++r) {  // table of the group's first C item, if that is its first item
        const CmItem d0 = cm_decode(a, me, r);
        if (!d0.valid) continue;
        if (d0.copy && lt < TABF / 4) {
            tnext = __ldcg(reinterpret_cast<const float4 *>(a.table + (int64_t)d0.b * TABF) + lt);
            reinterpret_cast<float4 *>(gtabs)[lt] = tnext;
        }
        break;
    }
    int n_sim = 0, i = 0, n = 0;  // S items / items of the group, valid items of the CTA (stage ring position)
    for (int r = 0; !cm_done(a, r); ++r)
    for (int g = 0; g < NG; ++g) {
        const CmItem d = cm_decode(a, blockIdx.x + g * gridDim.x, r);
        if (!d.valid) continue;
        const int s = n % NST;
        const uint32_t full_ph = (uint32_t)(n / NST) & 1u;
        ++n;
        if (g != grp) continue;
        const CmItem dn = cm_decode(a, me, r + 1);
        const bool tn = dn.valid && dn.copy && lt < TABF / 4;
        if (tn) tnext = __ldcg(reinterpret_cast<const float4 *>(a.table + (int64_t)dn.b * TABF) + lt);
        CM_PROBE(unsigned long long pr_a = global_ns();)
        mbar_wait(smem_u32(full + s), full_ph);
        CM_PROBE(unsigned long long pr_b = global_ns(); pr_wait += pr_b - pr_a; ++pr_n;)
        const int p0 = (d.chunk * 256 + lt) * 4, c0 = d.slab * CH;
        const bool live = p0 < a.P;
        const uint8_t *stb = smem_raw + s * kStageBytes;
        const float4 *st = reinterpret_cast<const float4 *>(stb);
        const uint32_t mw = live ? reinterpret_cast<const uint32_t *>(stb + CH * (R + 1) * 4096)[lt] : 0u;
        if (!d.copy) {
            // ---------------- S item: partial similarity ----------------
            float acc[G2];
#pragma unroll
            for (int r = 0; r < G2; ++r) acc[r] = 0.0f;
            if (live) {
                float4 vm[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {  // vt' * vr'                  :220
                    const uint32_t m = mw & (mw >> (r + 1)) & 0x01010101u;
                    vm[r] = make_float4((float)(m & 1u), (float)((m >> 8) & 1u), (float)((m >> 16) & 1u),
                                        (float)((m >> 24) & 1u));
                    if (d.slab == 0) acc[R + r] = (vm[r].x + vm[r].y) + (vm[r].z + vm[r].w);  // :221
                }
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    if (c0 + k < a.C) {
                        const float4 ct = st[(k * (R + 1)) * 256 + lt];
#pragma unroll
                        for (int r = 0; r < R; ++r) {  // vmap * c_t * c_r            :226
                            const float4 cr = st[(k * (R + 1) + r + 1) * 256 + lt];
                            acc[r] += vm[r].x * ct.x * cr.x;
                            acc[r] += vm[r].y * ct.y * cr.y;
                            acc[r] += vm[r].z * ct.z * cr.z;
                            acc[r] += vm[r].w * ct.w * cr.w;
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(empty + s));  // the stage may be refilled
            // CTA-group sum of the 2R values, fixed order; double-buffered scratch: warp 0 of the group
            // reads while the others move on to their next item
            float *rd = red[grp][n_sim & 1];
            const int slot = n_sim % kMailbox;
            if constexpr (G2 == 8) {
                const float y = warp_sum8(acc);
                if ((lane & 3) == 0) rd[(lane >> 2) * 8 + lw] = y;
                bar_group(bar_id);
                if (lw == 0) {
                    float t = rd[(lane >> 2) * 8 + 2 * (lane & 3)] + rd[(lane >> 2) * 8 + 2 * (lane & 3) + 1];
                    t += __shfl_xor_sync(0xffffffffu, t, 1);
                    t += __shfl_xor_sync(0xffffffffu, t, 2);
                    if (lane == 0) while (n_sim - mb_done[grp] >= kMailbox) __nanosleep(100);
                    __syncwarp();
                    if ((lane & 3) == 0) mbox[grp][slot][lane >> 2] = t;
                }
            } else {
#pragma unroll
                for (int k = 0; k < G2; ++k) {
                    acc[k] = warp_sum(acc[k]);
                    if (lane == 0) rd[k * 8 + lw] = acc[k];
                }
                bar_group(bar_id);
                if (lw == 0) {
                    if (lane == 0) while (n_sim - mb_done[grp] >= kMailbox) __nanosleep(100);
                    __syncwarp();
                    if (lane < G2) {
                        float t = 0.0f;
#pragma unroll
                        for (int w8 = 0; w8 < 8; ++w8) t += rd[lane * 8 + w8];
                        mbox[grp][slot][lane] = t;
                    }
                }
            }
            if (lw == 0) {
                __syncwarp();
                if (lane == 0) {
                    mbox_b[grp][slot] = d.b;
                    mbox_idx[grp][slot] = d.idx;
                    __threadfence_block();
                    mb_ready[grp] = n_sim + 1;
                }
            }
            ++n_sim;
        } else {
            // ---------------- C item: cat[c_t, sum_r c_r * w_r] ----------------
            float *tb = gtabs + (i & 1) * TABF;
            bool bad = false;
            if (lt < TABF / 4) {  // the chunk this thread fetched one item ago
                const float4 t4 = reinterpret_cast<const float4 *>(tb)[lt];
                bad = isnan(t4.x) || isnan(t4.y) || isnan(t4.z) || isnan(t4.w);
            }
            if (bar_group_or(bar_id, bad)) {  // also: the table chunks of the other threads are visible
                CM_PROBE(++pr_slow; unsigned long long sl_t = global_ns();
                         if (lt == 0 && d.b < 64) atomicMin(&g_cm_probe2[1088 + d.b], sl_t);)
                if (lt == 0) {
                    CM_PROBE(if (ld_acquire(a.flag + d.b) != 0u && blockIdx.x < 256) g_cm_probe2[blockIdx.x * 8 + 0] += 1;)
                    if (ld_acquire(a.flag + d.b) == 0u) {
                        const unsigned long long t0 = global_ns();
                        while (ld_acquire(a.flag + d.b) == 0u) {
                            __nanosleep(64);
                            if (global_ns() - t0 > 2000000000ull) __trap();
                        }
                    }
                }
                bar_group(bar_id);
                for (int q = lt; q < TABF; q += 256) tb[q] = __ldcg(a.table + (int64_t)d.b * TABF + q);
                bar_group(bar_id);
                CM_PROBE(if (tid == 0 && blockIdx.x < 256) g_cm_probe2[blockIdx.x * 8 + 1] += global_ns() - sl_t;)
            }
            if (live) {
                int pat[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) pat[j] = (int)((mw >> (8 * j + 1)) & ((1u << R) - 1u)) * (R + 1);
                float *ob = a.out + (int64_t)d.b * (2 * a.C + 1) * a.P + p0;
                if (d.slab == 0) {
                    const float4 c4 = make_float4(tb[pat[0] + R], tb[pat[1] + R], tb[pat[2] + R], tb[pat[3] + R]);
                    st_stream4(ob + (int64_t)(2 * a.C) * a.P, c4);
                    st_stream4(a.c_mask + (int64_t)d.b * a.P + p0, c4);
                }
                float4 wg[R];
#pragma unroll
                for (int r = 0; r < R; ++r) wg[r] = make_float4(tb[pat[0] + r], tb[pat[1] + r], tb[pat[2] + r], tb[pat[3] + r]);
#pragma unroll
                for (int k = 0; k < CH; ++k) {
                    const int c = c0 + k;
                    if (c >= a.C) break;
                    float4 o = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
                    for (int r = 0; r < R; ++r) {  // sum_r c_r * w_r, sequential over r    :238
                        const float4 cr = st[(k * (R + 1) + r + 1) * 256 + lt];
                        o.x = __fadd_rn(o.x, __fmul_rn(cr.x, wg[r].x));
                        o.y = __fadd_rn(o.y, __fmul_rn(cr.y, wg[r].y));
                        o.z = __fadd_rn(o.z, __fmul_rn(cr.z, wg[r].z));
                        o.w = __fadd_rn(o.w, __fmul_rn(cr.w, wg[r].w));
                    }
                    st_stream4(ob + (int64_t)c * a.P, st[(k * (R + 1)) * 256 + lt]);  // cat[c_t, ...]  :243
                    st_stream4(ob + (int64_t)(a.C + c) * a.P, o);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(empty + s));  // the stage may be refilled
        }
        // table of the group's next item: slot (i + 1) & 1 was last read by item i - 1 of the group, and
        // every thread of the group has passed the barrier of item i
        if (tn) reinterpret_cast<float4 *>(gtabs + ((i + 1) & 1) * TABF)[lt] = tnext;
        CM_PROBE(if (d.copy) pr_c += global_ns() - pr_b; else pr_s += global_ns() - pr_b;)
        ++i;
    }
    CM_PROBE(if (lt == 0 && grp == 0 && blockIdx.x < 256) {
        unsigned long long *o = g_cm_probe + blockIdx.x * 8;
        o[0] = global_ns() - pr_t0; o[1] = pr_wait; o[2] = pr_s; o[3] = pr_c; o[4] = pr_slow; o[5] = pr_n;
    })
    if (lt == 0) { __threadfence_block(); mb_fin[grp] = 1; }
}

constexpr int kCmGroups = 2;
template <int R, int CH>
constexpr int cm_pipe_smem(int nst) {
    return nst * cm_stage_bytes<R, CH>() + kCmGroups * 2 * cm_table_floats<R>() * 4;
}

template <int R, int CH, int NST>
int launch_cm_pipe_n(CmArgs a, cudaStream_t st) {
    constexpr int smem = cm_pipe_smem<R, CH>(NST);
    auto kern = cm_pipe_kernel<R, CH, NST, kCmGroups>;
    static bool ready = false;
    if (!ready) {
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
            cudaGetLastError();
            return -1;
        }
        ready = true;
    }
    a.n_items = a.chunks * ((a.C + CH - 1) / CH);
    a.nparts = a.n_items;
    const int64_t total = (int64_t)a.B * a.n_items;  // per stream
    if (total > (1ll << 29)) return -1;
    int64_t grid = sm_count();  // one resident wave (1 CTA/SM): the CTAs wait on each other
    if (grid * kCmGroups > total) grid = (total + kCmGroups - 1) / kCmGroups;
    a.workers = (int)grid * kCmGroups;
    a.rounds = (int)((total + a.workers - 1) / a.workers);
    // every S item of a sample before any C item of it in the common round sequence (cm_decode), plus
    // MT_CM_LAG rounds so that the table of a sample is normally published before its first C item
    int need = 1;
    for (int b = 0; b < a.B; ++b) {
        const int i_max = (int)((((int64_t)b + 1) * a.n_items - 1) / a.workers);
        const int j_min = (int)(((int64_t)b * a.n_items) / a.workers);
        if (i_max - j_min + 1 > need) need = i_max - j_min + 1;
    }
    a.lag = tuning("MT_CM_LAG", 2);
    if (a.lag < 0) a.lag = 0;
    if (a.lag > 4) a.lag = 4;
    a.head = need + a.lag;
    if (a.head > a.rounds) a.head = a.rounds;
    a.dv_items = make_fastdiv((uint32_t)a.n_items);
    a.dv_chunks = make_fastdiv((uint32_t)a.chunks);
    EncodeTiledFn enc = encode_fn();
    if (!enc || (a.P & 255) != 0 || (int64_t)a.B * a.C * a.f > (1ll << 31) - 1) return -1;
    CUtensorMap map_c;
    {
        cuuint64_t dims[3] = {256, (cuuint64_t)(a.P / 256), (cuuint64_t)a.B * a.C * a.f};
        cuuint64_t strides[2] = {1024, (cuuint64_t)a.P * 4};
        cuuint32_t box[3] = {256, 4, (cuuint32_t)(CH * (R + 1))};
        cuuint32_t estr[3] = {1, 1, 1};
        if (enc(&map_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float *>(a.c_feats), dims, strides, box, estr,
                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
    }
    launch(cm_masks_kernel, dim3((a.P + 255) / 256, a.B), 256, 0, st, a);
    launch(kern, dim3((unsigned)grid), kCmGroups * 256 + 64, (size_t)smem, st, map_c, a);
    return launch_status("mt_cm_match_fwd");
}

// deepest ring that fits 227 KB of shared memory (static smem of the kernel: < 3 KB)
template <int R, int CH>
int launch_cm_pipe(CmArgs a, cudaStream_t st) {
    constexpr int kBudget = 224 * 1024;
    const int want = tuning("MT_CM_STAGES", 5);
    if ((a.P & 15) != 0) return -1;  // bulk copies: 16 B aligned rows of the byte masks
    if constexpr (cm_pipe_smem<R, CH>(5) <= kBudget) { if (want >= 5) return launch_cm_pipe_n<R, CH, 5>(a, st); }
    if constexpr (cm_pipe_smem<R, CH>(4) <= kBudget) { if (want >= 4) return launch_cm_pipe_n<R, CH, 4>(a, st); }
    if constexpr (cm_pipe_smem<R, CH>(3) <= kBudget) { if (want >= 3) return launch_cm_pipe_n<R, CH, 3>(a, st); }
    if constexpr (cm_pipe_smem<R, CH>(2) <= kBudget) return launch_cm_pipe_n<R, CH, 2>(a, st);
    return -1;
}

template <int R>
int launch_cm(CmArgs a, cudaStream_t st) {
    // one persistent launch (pass 2 from L2); MT_CM_FUSED=0 keeps the three-launch path
    // MT_CM_FUSED=1 (experimental, off): the persistent pipelined kernel K3p.  Parity-green, but on B200
    // it is 1.3-1.4x slower than the three launches (profiles/r1_experiments.md).
    if (R <= 7 && tuning("MT_CM_FUSED", 0)) {  // the mask byte holds the target and up to 7 references
        const int rc = launch_cm_pipe<R, 2>(a, st);
        if (rc >= 0) return rc;
    }
    a.b_off = 0;
    launch(cm_masks_kernel, dim3((a.P + 255) / 256, a.B), 256, 0, st, a);
    // MT_CM_CHUNK > 0 (off): samples in groups, sim(g0) | weights(g0) | copy(g0) + sim(g1) | weights(g1) |
    // copy(g1) + sim(g2) | ...  - pass 2 of a group shares a launch with pass 1 of the next one and
    // re-reads its c_feats from L2 (a group of 4 samples is 42 MB of the 126 MB).  Measured on B200:
    // 52.0 / 57.3 / 63.5 us for groups of 4 / 3 / 2 against 46.9 us for one group at B=8, 166.6 against
    // 140.3 us at B=32: every additional launch costs 3-5 us of ramp and tail, more than the L2 hits
    // return.  (Separate launches per pass and group were worse still: 111/72/55 us for 1/2/4.)
    int chunk = tuning("MT_CM_CHUNK", 0);
    if (chunk < 1 || chunk > a.B) chunk = a.B;
    const int cc = tuning("MT_CM_COPY_CH", kCopyChannels);
    const bool merged = chunk < a.B;
    if (merged) a.sim_ch = tuning("MT_CM_MERGE_CH", 4) == 2 ? 2 : 4;  // one slab width for both passes
    a.nparts = a.chunks * ((a.C + a.sim_ch - 1) / a.sim_ch);
    const int slabs = (a.C + a.sim_ch - 1) / a.sim_ch;
    for (int b0 = 0; b0 < a.B; b0 += chunk) {
        const int nb = a.B - b0 < chunk ? a.B - b0 : chunk;
        a.b_off = b0;
        if (!merged || b0 == 0) {
            dim3 g1(a.chunks, slabs, nb);
            if (a.sim_ch == 2) launch(cm_sim_kernel<R, 2>, g1, 256, 0, st, a);
            else launch(cm_sim_kernel<R, 4>, g1, 256, 0, st, a);
        }
        if (!merged && R <= 7 && tuning("MT_CM_TABLE", 1)) {  // pass 1b folded into pass 2 (the mask byte holds <= 7 references)
            if (cc == 2) launch(cm_copy2_kernel<R, 2>, dim3(a.chunks, (a.C + 1) / 2, nb), 256, 0, st, a);
            else launch(cm_copy2_kernel<R, 4>, dim3(a.chunks, (a.C + 3) / 4, nb), 256, 0, st, a);
            continue;
        }
        dim3 gw(a.chunks, nb);
        launch(cm_weights_kernel<R>, gw, 256, 0, st, a);
        if (merged && b0 + nb < a.B) {
            a.n_copy = nb;
            a.b_sim = b0 + nb;
            a.n_sim = a.B - a.b_sim < chunk ? a.B - a.b_sim : chunk;
            dim3 g2(a.chunks, 2 * slabs, a.n_copy > a.n_sim ? a.n_copy : a.n_sim);
            if (a.sim_ch == 2) launch(cm_copy_sim_kernel<R, 2>, g2, 256, 0, st, a);
            else launch(cm_copy_sim_kernel<R, 4>, g2, 256, 0, st, a);
        } else if (merged) {
            a.copy_reverse = 0;
            dim3 g2(a.chunks, slabs, nb);
            if (a.sim_ch == 2) launch(cm_copy_kernel<R, 2>, g2, 256, 0, st, a);
            else launch(cm_copy_kernel<R, 4>, g2, 256, 0, st, a);
        } else if (cc == 2) {
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
           align256(B * R * P * 4) + 2 * align256((int64_t)B * 4) + align256(B * P) +
           align256(B * (int64_t)(1 << R) * (R + 1) * 4);
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
    ws += align256((int64_t)B * a.R * a.P * 4);
    a.count = reinterpret_cast<unsigned int *>(ws);
    ws += align256((int64_t)B * 4);
    a.flag = reinterpret_cast<unsigned int *>(ws);
    ws += align256((int64_t)B * 4);
    a.pmask = reinterpret_cast<unsigned char *>(ws);
    ws += align256((int64_t)B * a.P);
    a.table = reinterpret_cast<float *>(ws);
    a.n_items = 0; a.lag = 1;
    a.copy_reverse = tuning("MT_CM_COPY_REVERSE", 1);
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

#ifdef MT_DEV_PROBES
extern "C" __attribute__((visibility("default"))) int mt_debug_cm_reset(void) {
    static unsigned long long init[2048];
    for (int i = 0; i < 2048; ++i) init[i] = (i >= 1088 && i < 1152) ? ~0ull : 0ull;
    return cudaMemcpyToSymbol(mt::g_cm_probe2, init, sizeof(init)) == cudaSuccess ? 0 : -2;
}
extern "C" __attribute__((visibility("default"))) int mt_debug_cm_probe(unsigned long long *dst, int n) {
    cudaDeviceSynchronize();
    if (n < 0) return cudaMemcpyFromSymbol(dst, mt::g_cm_probe2, sizeof(unsigned long long) * 2048) == cudaSuccess ? 0 : -2;
    return cudaMemcpyFromSymbol(dst, mt::g_cm_probe, sizeof(unsigned long long) * (n < 2048 ? n : 2048)) == cudaSuccess
               ? 0 : -2;
}
#endif
