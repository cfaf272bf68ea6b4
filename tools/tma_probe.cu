// Developer probe: which (box, coordinate, expected-bytes) combinations of a non-swizzled 4-D TMA
// box load complete on this GPU.  nvcc -arch=sm_100a -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int c3, uint32_t bytes,
                      int nfloats, float *out, int *status) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + ((nfloats * 4 + 127) & ~127));
    uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(bar), dst = (uint32_t)__cvta_generic_to_shared(smem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) reinterpret_cast<float *>(smem)[i] = -7.0f;
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
            ::"r"(dst), "l"(&map), "r"(bar_a), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
        uint32_t ok = 0, spins = 0;
        while (!ok && spins < (1u << 20)) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar_a) : "memory");
            ++spins;
        }
        status[0] = ok; status[1] = spins;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nfloats; i += blockDim.x) out[i] = reinterpret_cast<float *>(smem)[i];
}

int main() {
    const int W = 204, H = 200, F = 4, B = 3;
    std::vector<float> h((size_t)W * H * F * B);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 100003);
    float *d, *out; int *status;
    cudaMalloc(&d, h.size() * 4); cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&out, 1 << 20); cudaMalloc(&status, 8);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int boxes[][2] = {{32, 32}, {48, 48}, {40, 40}, {64, 32}, {48, 8}, {32, 48}};
    const int coords[][2] = {{0, 0}, {16, 8}, {-4, -5}, {-44, -45}, {180, 190}, {300, 300}, {-400, 7}};
    for (auto &bx : boxes) for (auto &co : coords) {
        CUtensorMap map;
        cuuint64_t dims[4] = {W, H, F, B};
        cuuint64_t strides[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * F * 4};
        cuuint32_t box[4] = {(cuuint32_t)bx[0], (cuuint32_t)bx[1], 1, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = cuTensorMapEncodeTiled(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es,
                                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("box %dx%d encode failed %d\n", bx[0], bx[1], (int)r); continue; }
        const int nf = bx[0] * bx[1];
        cudaMemset(status, 0, 8);
        probe<<<1, 128, ((nf * 4 + 127) & ~127) + 64>>>(map, co[0], co[1], 1, 2, (uint32_t)nf * 4, nf, out, status);
        cudaError_t e = cudaDeviceSynchronize();
        int st[2] = {0, 0};
        std::vector<float> o(nf);
        cudaMemcpy(st, status, 8, cudaMemcpyDeviceToHost); cudaMemcpy(o.data(), out, nf * 4, cudaMemcpyDeviceToHost);
        // check a few elements
        int bad = 0, untouched = 0;
        for (int y = 0; y < bx[1]; ++y) for (int x = 0; x < bx[0]; ++x) {
            const int gx = co[0] + x, gy = co[1] + y;
            float exp = 0.0f;
            if (gx >= 0 && gx < W && gy >= 0 && gy < H) exp = h[(((size_t)2 * F + 1) * H + gy) * W + gx];
            const float got = o[y * bx[0] + x];
            if (got == -7.0f) ++untouched; else if (got != exp) ++bad;
        }
        printf("box %2dx%2d at (%4d,%4d): %s done=%d spins=%d bad=%d untouched=%d\n", bx[0], bx[1], co[0], co[1],
               cudaGetErrorString(e), st[0], st[1], bad, untouched);
        if (e != cudaSuccess) return 1;
    }
    return 0;
}
