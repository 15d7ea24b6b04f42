// Training-step tail on one flat fp32 parameter buffer (VRM:78-108): GradScaler's unscale + non-finite
// check and the AdamW update, each as one HBM-bound pass over the 59 M parameters instead of ~633 small per-tensor
// launches.  Both kernels are grid-stride over float4 with a grid of a multiple of the SM count.
#include <cmath>
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ bool finite4(float4 a) {
    // |x| <= FLT_MAX is false for inf and NaN
    return (fabsf(a.x) <= 3.402823466e38f) & (fabsf(a.y) <= 3.402823466e38f) & (fabsf(a.z) <= 3.402823466e38f) &
           (fabsf(a.w) <= 3.402823466e38f);
}

// found[0] = 1.0f if any gradient element is inf/NaN (GradScaler.unscale_'s found_inf, VRM:101); the caller zeroes it
__global__ void __launch_bounds__(kThreads) grad_check_kernel(const float *__restrict__ g, int64_t n, float *found) {
    const int64_t n4 = n >> 2;
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    bool ok = true;
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads)
        ok &= finite4(__ldg(g4 + i));
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        ok &= fabsf(g[i]) <= 3.402823466e38f;
    if (__syncthreads_or(!ok) && threadIdx.x == 0) *found = 1.0f;
}

struct AdamConst {
    float lr, beta1, beta2, eps, decay, inv_bc1, inv_sqrt_bc2, grad_scale;
};

__device__ __forceinline__ void adam1(float &p, float g, float &m, float &v, const AdamConst &c) {
    g *= c.grad_scale;                       // 1/loss_scale and 1/world_size folded into one factor
    p *= c.decay;                            // decoupled weight decay: p *= 1 - lr*wd  (torch.optim.AdamW)
    m = c.beta1 * m + (1.0f - c.beta1) * g;
    v = c.beta2 * v + (1.0f - c.beta2) * g * g;
    const float denom = sqrtf(v) * c.inv_sqrt_bc2 + c.eps;
    p -= c.lr * c.inv_bc1 * (m / denom);
}

__global__ void __launch_bounds__(kThreads)
adamw_flat_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                  int64_t n, AdamConst c, const float *__restrict__ found) {
    if (found != nullptr && *found != 0.0f) return;          // GradScaler.step skips the update on overflow
    const int64_t n4 = n >> 2;
    float4 *p4 = reinterpret_cast<float4 *>(p), *m4 = reinterpret_cast<float4 *>(m), *v4 = reinterpret_cast<float4 *>(v);
    const float4 *g4 = reinterpret_cast<const float4 *>(g);
    for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n4; i += (int64_t)gridDim.x * kThreads) {
        float4 P = p4[i], M = m4[i], V = v4[i];
        const float4 G = __ldg(g4 + i);
        adam1(P.x, G.x, M.x, V.x, c);
        adam1(P.y, G.y, M.y, V.y, c);
        adam1(P.z, G.z, M.z, V.z, c);
        adam1(P.w, G.w, M.w, V.w, c);
        p4[i] = P, m4[i] = M, v4[i] = V;
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
        adam1(p[i], g[i], m[i], v[i], c);
}

int grid_for(int64_t n) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t want = cdiv64(cdiv64(n, 4), kThreads);
    const int64_t cap = (int64_t)sms * 8;
    return (int)(want < cap ? (want > 0 ? want : 1) : cap);
}

bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

extern "C" int turtle_grad_check_finite(const float *g, int64_t n, float *found, void *stream) {
    if (!g || !found || n <= 0 || !aligned16(g)) return TURTLE_EINVAL;
    grad_check_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(g, n, found);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}

extern "C" int turtle_adamw_flat(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1,
                                 float beta2, float eps, float weight_decay, int step, float grad_scale,
                                 const float *found, void *stream) {
    if (!p || !g || !m || !v || n <= 0 || step < 1) return TURTLE_EINVAL;
    if (!aligned16(p) || !aligned16(g) || !aligned16(m) || !aligned16(v)) return TURTLE_EINVAL;
    if (!(beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f)) return TURTLE_EINVAL;
    AdamConst c;
    c.lr = lr, c.beta1 = beta1, c.beta2 = beta2, c.eps = eps, c.decay = 1.0f - lr * weight_decay;
    // bias corrections in double on the host, as torch does for a Python-scalar step
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    c.inv_bc1 = (float)(1.0 / bc1);
    c.inv_sqrt_bc2 = (float)(1.0 / sqrt(bc2));
    c.grad_scale = grad_scale;
    adamw_flat_kernel<<<grid_for(n), kThreads, 0, as_stream(stream)>>>(p, g, m, v, n, c, found);
    TURTLE_CHECK_LAUNCH();
    return TURTLE_OK;
}
