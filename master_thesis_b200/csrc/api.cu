// api.cu - library plumbing of libmt_b200.so: error reporting, device info and
// the host-buffer (end-to-end) entry points.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "mt_common.cuh"

namespace mt {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int launch_status(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) return MT_OK;
    set_error("%s: %s", what, cudaGetErrorString(e));
    return (e == cudaErrorNoKernelImageForDevice || e == cudaErrorInvalidDeviceFunction ||
            e == cudaErrorNoDevice)
               ? MT_ERR_NO_DEVICE
               : MT_ERR_CUDA;
}

int tuning(const char *name, int dflt);

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

bool pdl_enabled() { return tuning("MT_PDL", 1) != 0; }

namespace {
struct TuningEntry { char name[32]; int value; };
TuningEntry g_tuning[32];
int g_tuning_used = 0;
std::mutex g_tuning_mu;  // launches may come from several host threads (one per GPU): the table is shared
}  // namespace

int tuning(const char *name, int dflt) {
    std::lock_guard<std::mutex> lock(g_tuning_mu);
    for (int i = 0; i < g_tuning_used; ++i)
        if (strcmp(g_tuning[i].name, name) == 0) return g_tuning[i].value;
    const char *e = getenv(name);
    const int v = e ? atoi(e) : dflt;
    if (g_tuning_used < 32 && strlen(name) < sizeof(g_tuning[0].name)) {
        strcpy(g_tuning[g_tuning_used].name, name);
        g_tuning[g_tuning_used++].value = v;
    }
    return v;
}

int set_tuning(const char *name, int value) {
    std::lock_guard<std::mutex> lock(g_tuning_mu);
    for (int i = 0; i < g_tuning_used; ++i)
        if (strcmp(g_tuning[i].name, name) == 0) { g_tuning[i].value = value; return MT_OK; }
    if (g_tuning_used >= 32 || strlen(name) >= sizeof(g_tuning[0].name)) return MT_ERR_INVALID;
    strcpy(g_tuning[g_tuning_used].name, name);
    g_tuning[g_tuning_used++].value = value;
    return MT_OK;
}

}  // namespace mt

using namespace mt;

extern "C" int mt_version(void) { return 100; }
extern "C" int mt_set_tuning(const char *name, int value) {
    MT_REQUIRE(name && name[0], "mt_set_tuning: empty name");
    int rc = mt::set_tuning(name, value);
    if (rc) set_error("mt_set_tuning: table full or name too long (%s)", name);
    return rc;
}
extern "C" const char *mt_last_error(void) { return g_err; }
extern "C" int64_t mt_workspace_bytes(void) { return kWorkspaceBytes; }

extern "C" int mt_device_info(int *sms, int *cc_major, int *cc_minor) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("mt_device_info: %s", cudaGetErrorString(e));
        return MT_ERR_NO_DEVICE;
    }
    int a = 0, b = 0, c = 0;
    cudaDeviceGetAttribute(&a, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&b, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&c, cudaDevAttrComputeCapabilityMinor, dev);
    if (sms) *sms = a;
    if (cc_major) *cc_major = b;
    if (cc_minor) *cc_minor = c;
    return MT_OK;
}

// ---- host-buffer entry points ------------------------------------------------
namespace {

#define MT_CUDA(call)                                                    \
    do {                                                                 \
        cudaError_t e_ = (call);                                         \
        if (e_ != cudaSuccess) {                                         \
            set_error("%s: %s", #call, cudaGetErrorString(e_));          \
            rc = MT_ERR_CUDA;                                            \
            goto done;                                                   \
        }                                                                \
    } while (0)

// Per-thread staging: device buffers + two streams, grown on demand and reused.
struct Stage {
    float *d = nullptr;
    size_t cap = 0;
    cudaStream_t s = nullptr;
};
thread_local Stage g_stage;

int stage_reserve(size_t bytes) {
    if (!g_stage.s && cudaStreamCreateWithFlags(&g_stage.s, cudaStreamNonBlocking) != cudaSuccess) {
        set_error("cudaStreamCreate failed");
        return MT_ERR_CUDA;
    }
    if (bytes <= g_stage.cap) return MT_OK;
    if (g_stage.d) cudaFree(g_stage.d);
    g_stage.d = nullptr;
    g_stage.cap = 0;
    cudaError_t e = cudaMalloc(&g_stage.d, bytes);
    if (e != cudaSuccess) {
        set_error("cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
        return MT_ERR_CUDA;
    }
    g_stage.cap = bytes;
    return MT_OK;
}

int align_host(const float *x_refs, const float *m_refs, const float *m_target, const float *grid,
               size_t grid_elems, float *x_aligned, float *v_aligned, float *v_maps, int B, int F,
               int H, int W, int flags) {
    int rc = MT_OK;
    MT_REQUIRE(x_refs && m_refs && m_target && grid && x_aligned && v_aligned && v_maps,
               "mt_*_align_host: NULL argument");
    MT_REQUIRE(B > 0 && F > 0 && H > 0 && W > 0, "mt_*_align_host: empty shape");
    const size_t P = (size_t)H * W, n = (size_t)B * F;
    auto up4 = [](size_t v) { return (v + 3) & ~size_t(3); };
    const size_t o_x = 0, o_m = o_x + up4(n * 3 * P), o_mt = o_m + up4(n * P),
                 o_g = o_mt + up4((size_t)B * P), o_xa = o_g + up4(grid_elems),
                 o_va = o_xa + up4(n * 3 * P), o_vm = o_va + up4(n * P), total = o_vm + up4(n * P);
    rc = stage_reserve(total * sizeof(float));
    if (rc) return rc;
    {
        float *d = g_stage.d;
        cudaStream_t s = g_stage.s;
        MT_CUDA(cudaMemcpyAsync(d + o_x, x_refs, n * 3 * P * 4, cudaMemcpyHostToDevice, s));
        MT_CUDA(cudaMemcpyAsync(d + o_m, m_refs, n * P * 4, cudaMemcpyHostToDevice, s));
        MT_CUDA(cudaMemcpyAsync(d + o_mt, m_target, (size_t)B * P * 4, cudaMemcpyHostToDevice, s));
        MT_CUDA(cudaMemcpyAsync(d + o_g, grid, grid_elems * 4, cudaMemcpyHostToDevice, s));
        // inputs and outputs in the reference's logical (B, C, F, H, W) layout
        rc = mt_warp_fwd(d + o_x, (int64_t)3 * F * P, (int64_t)F * P, (int64_t)P,
                         d + o_m, (int64_t)F * P, (int64_t)P, d + o_g, d + o_mt, (int64_t)P,
                         d + o_xa, (int64_t)3 * F * P, (int64_t)F * P, (int64_t)P,
                         d + o_va, d + o_vm, B, 3, F, H, W, flags, s);
        if (rc) goto done;
        MT_CUDA(cudaMemcpyAsync(x_aligned, d + o_xa, n * 3 * P * 4, cudaMemcpyDeviceToHost, s));
        MT_CUDA(cudaMemcpyAsync(v_aligned, d + o_va, n * P * 4, cudaMemcpyDeviceToHost, s));
        MT_CUDA(cudaMemcpyAsync(v_maps, d + o_vm, n * P * 4, cudaMemcpyDeviceToHost, s));
        MT_CUDA(cudaStreamSynchronize(s));
    }
done:
    return rc;
}

}  // namespace

extern "C" int mt_cpn_align_host(const float *x_refs, const float *m_refs, const float *m_target,
                                 const float *theta, float *x_aligned, float *v_aligned,
                                 float *v_maps, int B, int F, int H, int W) {
    return align_host(x_refs, m_refs, m_target, theta, (size_t)B * F * 6, x_aligned, v_aligned, v_maps,
                      B, F, H, W, MT_GRID_AFFINE | MT_VIS_BILINEAR | MT_VIS_FROM_MASK);
}

extern "C" int mt_dfpn_align_host(const float *x_refs, const float *m_refs, const float *m_target,
                                  const float *flow, float *x_aligned, float *v_aligned,
                                  float *v_maps, int B, int F, int H, int W) {
    return align_host(x_refs, m_refs, m_target, flow, (size_t)B * F * H * W * 2, x_aligned, v_aligned,
                      v_maps, B, F, H, W, MT_ALIGN_CORNERS | MT_VIS_FROM_MASK);
}
