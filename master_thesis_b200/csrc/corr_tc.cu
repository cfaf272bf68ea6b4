// corr_tc.cu - K2 on the 5th-generation tensor cores: tcgen05.mma (kind::tf32),
// operands staged by TMA, accumulator in TMEM, normalisation + masks fused as a
// rank-1 scaling in the tcgen05.ld epilogue.
//
// Replaces CorrelationVGG.correlation_masked_4d, master_thesis/model_dfpn.py:534-565 (a7):
//   out[b,f,m,n] = sum_k ft[b,k,m] fr[b,k,f,n] * sa[b,m] * sb[b,f,n]
//   sa = v_t / (||ft * v_t||_2 + 1e-9)     (:551-560)     sb likewise for the reference (:561-562)
// i.e. the reference's "mask, normalise, matmul" with the two normalisations factored out of
// the contraction.  Masked rows / columns come out exactly 0 (scale 0 times a finite sum).
//
// Data layout.  The features arrive fp32 and MN-major: ft (B,C,P) and fr (B,C,F,P) have the
// pixel index contiguous, the contraction index (channel) strided.  They are consumed AS IS:
// kind::tf32 reads fp32 words from shared memory (10-bit mantissa used), and both operands use
// MN-major shared-memory descriptors.  For MN-major tf32 operands the ONLY legal shared-memory
// layout is "128 B swizzle with a 32 B base" (UMMA layout type 1, Swizzle<2,5,2>: the four 32 B
// chunks of a 128 B row are permuted by the row index mod 4; a plain SWIZZLE_128B descriptor
// is silently ignored - the first version of this kernel produced zeros).  TMA boxes of
// {32 pixels = 128 B, BK channels, 4 pixel groups} with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B land
// exactly in that canonical layout:
//   4 channel rows x 128 B = one 512 B swizzle atom; SBO = 512 B between 4-channel groups;
//   LBO = BK * 128 B between 32-pixel groups; one MMA (K = 8) spans two atoms.
// One CTA = one 128 x 128 output tile of one (b, f) frame, K = C in BK = 32 slices, 3 stages
// (100 KB of shared memory, so two CTAs share an SM and one's epilogue overlaps the other's
// main loop; with 4 stages and one CTA per SM the tensor pipe was 21 % active, profiles/).
// Warp roles: 0 = TMA producer, 1 = TMEM owner + MMA issuer, 2..5 = epilogue (one TMEM lane
// quarter each): tcgen05.ld 32 lanes x 32 columns -> scale -> 128 B per-row stores.
#include <cuda.h>

#include "mt_common.cuh"

namespace mt {
namespace {

constexpr int kTileM = 128, kTileN = 128, kBK = 32, kStages = 3;  // 3 x 32 KB: two CTAs per SM
constexpr int kUmmaK = 8;  // tf32: 32 B of K per instruction
constexpr int kStageBytesA = kTileM * kBK * 4, kStageBytesB = kTileN * kBK * 4;
constexpr int kSmemBytes = kStages * (kStageBytesA + kStageBytesB) + 1024 /*align*/ + 1024 /*barriers, scales*/;
constexpr int kThreadsTc = 6 * 32;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (!ok && ++spins > (1u << 28)) __trap();  // a wedged pipeline must fail, not hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *map, uint32_t bar, int c0, int c1,
                                            int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}

// MN-major, SWIZZLE_128B_BASE32B shared-memory matrix descriptor (sm_100 "version 1").
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);        // start address  [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;  // leading byte offset [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;  // stride byte offset  [32,46)
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                           // layout type: SWIZZLE_128B_BASE32B
    return d;
}
// kind::tf32, fp32 accumulate, A and B MN-major, M x N
constexpr uint32_t instr_desc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) |
           ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct CorrTcArgs {
    const float *sa;  // (B, P)      v_t / (||ft v_t|| + 1e-9)
    const float *sb;  // (B, F, P)
    float *out;       // (B, F, P, P)
    int C, F, P;
};

__global__ void __launch_bounds__(kThreadsTc, 2)
corr_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
               const CorrTcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem;
    uint8_t *smem_b = smem + kStages * kStageBytesA;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + kStages * (kStageBytesA + kStageBytesB));
    uint64_t *full = bars, *empty = bars + kStages, *tmem_full = bars + 2 * kStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 1);
    float *s_sb = reinterpret_cast<float *>(bars + 2 * kStages + 2);  // kTileN floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tile = blockIdx.x, n_tile = blockIdx.y, frame = blockIdx.z;
    const int b = frame / a.F, f = frame - b * a.F;
    const int num_k = a.C / kBK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(full + s), 1);
            mbar_init(smem_u32(empty + s), 1);
        }
        mbar_init(smem_u32(tmem_full), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: kTileN fp32 accumulator columns x 128 lanes
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_slot)), "n"(kTileN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // everything above is on-chip setup and overlaps the tail of the previous kernel (PDL);
    // from here on the scales written by corr_scales_kernel are read
    pdl_sync();
    if (warp >= 2) {  // column scales of this tile
        const int t = threadIdx.x - 64;
        if (t < kTileN) s_sb[t] = __ldg(a.sb + ((int64_t)frame * a.P + n_tile * kTileN + t));
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(smem_u32(empty + s), ph ^ 1);
                mbar_expect_tx(smem_u32(full + s), kStageBytesA + kStageBytesB);
                // A: (pixel-in-group 32, channel C, pixel group P/32, batch B)
                tma_load_4d(smem_u32(smem_a + s * kStageBytesA), &map_a, smem_u32(full + s), 0, kb * kBK,
                            m_tile * (kTileM / 32), b);
                // B: (pixel-in-group 32, channel C, pixel group P/32, frame F, batch B)
                tma_load_5d(smem_u32(smem_b + s * kStageBytesB), &map_b, smem_u32(full + s), 0, kb * kBK,
                            n_tile * (kTileN / 32), f, b);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected lane) =====
        if (lane == 0) {
            constexpr uint32_t idesc = instr_desc(kTileM, kTileN);
            for (int kb = 0; kb < num_k; ++kb) {
                const int s = kb % kStages;
                const uint32_t ph = (kb / kStages) & 1;
                mbar_wait(smem_u32(full + s), ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_u32(smem_a + s * kStageBytesA), b0 = smem_u32(smem_b + s * kStageBytesB);
#pragma unroll
                for (int j = 0; j < kBK / kUmmaK; ++j) {
                    // K advance inside the stage: next 8 channels = two 512 B atoms = +1024 B
                    const uint64_t ad = umma_desc(a0 + j * 1024, kBK * 128, 512);
                    const uint64_t bd = umma_desc(b0 + j * 1024, kBK * 128, 512);
                    umma_tf32(tmem_base, ad, bd, idesc, (kb | j) != 0 ? 1u : 0u);
                }
                umma_commit(smem_u32(empty + s));  // frees the smem slot when these MMAs retire
            }
            umma_commit(smem_u32(tmem_full));      // accumulator complete
        }
    } else {
        // ===== epilogue: TMEM -> registers -> scale -> global =====
        const int q = warp & 3;  // TMEM lane quarter this warp may access
        const int m = m_tile * kTileM + q * 32 + lane;
        const float sa = __ldg(a.sa + ((int64_t)b * a.P + m));
        float *orow = a.out + (((int64_t)frame * a.P + m) * a.P + n_tile * kTileN);
        mbar_wait(smem_u32(tmem_full), 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
        for (int c0 = 0; c0 < kTileN; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
                float4 o;
                o.x = v[i] * sa * s_sb[c0 + i];
                o.y = v[i + 1] * sa * s_sb[c0 + i + 1];
                o.z = v[i + 2] * sa * s_sb[c0 + i + 2];
                o.w = v[i + 3] * sa * s_sb[c0 + i + 3];
                __stcs(reinterpret_cast<float4 *>(orow + c0 + i), o);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTileN) : "memory");
    }
}

// scale[n, p] = v / (||feat[:, p] * v||_2 + 1e-9) for frame n = b*F + f (F = 1 for the target).
// CTA = 32 pixels x 8 channel groups; coalesced 128 B rows; fixed-order reduction.
__global__ void __launch_bounds__(256) corr_scales_kernel(const float *__restrict__ src,
                                                          const float *__restrict__ vis,
                                                          float *__restrict__ scale, int C, int F, int P) {
    pdl_sync();
    __shared__ float part[8][33];
    const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const int64_t n = blockIdx.y;
    const int64_t b = n / F, f = n - b * F;
    const int64_t sc = (int64_t)F * P;
    float ss = 0.0f;
    if (p < P) {
        const float *s = src + b * C * sc + f * P + p;
        int k = grp;
        for (; k + 56 < C; k += 64) {  // 8 independent loads in flight per thread
            float v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = __ldg(s + (int64_t)(k + 8 * u) * sc);
#pragma unroll
            for (int u = 0; u < 8; ++u) ss = __fmaf_rn(v[u], v[u], ss);
        }
        for (; k < C; k += 8) {
            const float v = __ldg(s + k * sc);
            ss = __fmaf_rn(v, v, ss);
        }
    }
    part[grp][lane] = ss;
    __syncthreads();
    if (grp == 0 && p < P) {
        float tot = 0.0f;
#pragma unroll
        for (int g = 0; g < 8; ++g) tot += part[g][lane];
        const float v = vis ? __ldg(vis + n * P + p) : 1.0f;
        // ||f * v|| = |v| * ||f||;  scale = v / (||f v|| + 1e-9)
        scale[n * P + p] = __fdiv_rn(v, __fadd_rn(__fmul_rn(fabsf(v), sqrtf(tot)), 1e-9f));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int64_t align256(int64_t v) { return (v + 255) & ~int64_t(255); }

}  // namespace

int corr4d_tc_supported(int C, int P) {
    if (tuning("MT_CORR_SIMT", 0)) return 0;
    return (P % kTileM == 0 && P % kTileN == 0 && C % kBK == 0 && C >= kBK && P <= 65536) ? 1 : 0;
}

int64_t corr4d_tc_workspace_bytes(int B, int C, int F, int P) {
    (void)C;
    return align256((int64_t)B * P * 4) + align256((int64_t)B * F * P * 4);
}

int corr4d_tc_launch(const float *ft, const float *vt, const float *fr, const float *vr, float *out, void *ws,
                     int64_t ws_bytes, int B, int C, int F, int P, cudaStream_t st) {
    (void)ws_bytes;
    MT_REQUIRE(aligned16(ft) && aligned16(fr) && aligned16(out) && aligned16(ws),
               "mt_corr4d_fwd: pointers must be 16 B aligned");
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        set_error("mt_corr4d_fwd: cuTensorMapEncodeTiled is not available from the driver");
        return MT_ERR_NO_DEVICE;
    }
    float *sa = reinterpret_cast<float *>(ws);
    float *sb = reinterpret_cast<float *>(reinterpret_cast<char *>(ws) + align256((int64_t)B * P * 4));
    dim3 gs_t((P + 31) / 32, B), gs_r((P + 31) / 32, B * F);
    launch(corr_scales_kernel, gs_t, 256, 0, st, ft, vt, sa, C, 1, P);
    launch(corr_scales_kernel, gs_r, 256, 0, st, fr, vr, sb, C, F, P);

    CUtensorMap map_a, map_b;
    {
        // ft (B, C, P) viewed as (32, C, P/32, B): strides in bytes for dims 1..3
        cuuint64_t dims[4] = {32, (cuuint64_t)C, (cuuint64_t)(P / 32), (cuuint64_t)B};
        cuuint64_t strides[3] = {(cuuint64_t)P * 4, 128, (cuuint64_t)C * P * 4};
        cuuint32_t box[4] = {32, (cuuint32_t)kBK, (cuuint32_t)(kTileM / 32), 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float *>(ft), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("mt_corr4d_fwd: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
            return MT_ERR_CUDA;
        }
    }
    {
        // fr (B, C, F, P) viewed as (32, C, P/32, F, B)
        cuuint64_t dims[5] = {32, (cuuint64_t)C, (cuuint64_t)(P / 32), (cuuint64_t)F, (cuuint64_t)B};
        cuuint64_t strides[4] = {(cuuint64_t)F * P * 4, 128, (cuuint64_t)P * 4, (cuuint64_t)C * F * P * 4};
        cuuint32_t box[5] = {32, (cuuint32_t)kBK, (cuuint32_t)(kTileN / 32), 1, 1};
        cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float *>(fr), dims, strides, box,
                         estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            set_error("mt_corr4d_fwd: cuTensorMapEncodeTiled(B) failed with %d", (int)r);
            return MT_ERR_CUDA;
        }
    }
    {
        cudaError_t e = cudaFuncSetAttribute(corr_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) {
            set_error("mt_corr4d_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return MT_ERR_CUDA;
        }
    }
    CorrTcArgs a{sa, sb, out, C, F, P};
    dim3 grid(P / kTileM, P / kTileN, B * F);
    launch(corr_tc_kernel, grid, kThreadsTc, kSmemBytes, st, map_a, map_b, a);
    return launch_status("mt_corr4d_fwd");
}

}  // namespace mt
