// Probe: can a K-major SWIZZLE_128B A operand of tcgen05.mma start at ANY 128-byte row of a TMA-written tile (not just
// at multiples of the 8-row swizzle repeat)?  A depthwise 3x3 on the tensor cores wants the nine taps as nine shifted
// views of ONE halo tile (shift = (dy * row_pitch + dx) pixels = rows of the operand).
//   D[m][n] = sum_k A[m + shift][kofs*16 + k] * I[n][k]   (M = 128, N = 16, K = 16, fp16, B = identity, no swizzle)
// variants: base_offset field = 0, or = (start >> 7) & 7.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o umma_shift_probe umma_shift_probe.cu -lcuda
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int ROWS = 256;   // tile rows (pixels), 64 halves each

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128) probe(const __grid_constant__ CUtensorMap map, float *out, int shift, int kofs, int use_bo, int split) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar, mbar;
    __shared__ uint32_t tmem_base_sh;
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sB = smem0 + ROWS * 128;            // identity, 4 core matrices of 128 B
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // B = I[16][16] as 2x2 core matrices (8 rows x 16 B each): core (nb, kb) at (nb*2 + kb) * 128
    for (int i = threadIdx.x; i < 256; i += 128) {
        const int n = i / 16, k = i % 16;
        const int off = ((n / 8) * 2 + (k / 8)) * 128 + (n % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<__half *>(smem_raw + (sB - smem_u32(smem_raw)) + off) = __float2half(n == k ? 1.f : 0.f);
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base_sh)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sh;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(ROWS * (split ? 64 : 128)) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem0),
                     "l"(&map), "r"(smem_u32(&bar)), "r"(0), "r"(0)
                     : "memory");
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        } while (!ok);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem0 + shift * (split ? 64 : 128) + kofs * 32;
        uint64_t da = 0;
        da |= (uint64_t)((sa & 0x3FFFF) >> 4);
        da |= (uint64_t)1 << 16;
        da |= (uint64_t)((split ? 512 : 1024) >> 4) << 32;
        da |= (uint64_t)1 << 46;
        if (use_bo) da |= (uint64_t)((sa >> 7) & 7) << 49;
        da |= (uint64_t)(split ? 4 : 2) << 61;          // SWIZZLE_64B : SWIZZLE_128B
        uint64_t db = 0;                       // no swizzle, K-major: LBO = 128 (next core matrix in K), SBO = 256 (next 8 rows)
        db |= (uint64_t)((sB & 0x3FFFF) >> 4);
        db |= (uint64_t)(128 >> 4) << 16;
        db |= (uint64_t)(256 >> 4) << 32;
        db |= (uint64_t)1 << 46;
        const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | ((16u >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem),
                     "l"(da), "l"(db), "r"(idesc), "r"(0)
                     : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        } while (!ok);
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem) : "memory");
}

int main() {
    std::vector<__half> hx(ROWS * 64);
    __half *dx;
    float *dout;
    cudaMalloc(&dx, hx.size() * 2);
    cudaMalloc(&dout, 128 * 16 * 4);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(fn);
    CUtensorMap map;
    cuuint64_t dims[2] = {64, ROWS}, str[1] = {128};
    cuuint32_t box[2] = {64, ROWS}, es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dx, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    // 64-byte rows (the first 32 halves of every row of x), SWIZZLE_64B
    CUtensorMap map3;
    cuuint64_t dims3[2] = {32, ROWS}, str3[1] = {128};
    cuuint32_t box3[2] = {32, ROWS}, es3[2] = {1, 1};
    r = enc(&map3, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dx, dims3, str3, box3, es3, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("SW64 map encode: %d\n", (int)r);
    const size_t smem = ROWS * 128 + 512 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<float> ho(128 * 16);
    for (int split = 0; split < 2; ++split)
    for (int pattern = 0; pattern < 2; ++pattern) {          // 0: value = row, 1: value = column
        for (int i = 0; i < ROWS * 64; ++i) hx[i] = __float2half(pattern == 0 ? (float)(i / 64) : (float)(i % 64));
        cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice);
        for (int use_bo = 0; use_bo < 2; ++use_bo) {
            if (use_bo) continue;
            printf("%s map, pattern %s, base_offset %s: ", split ? "SW64 (64 B rows)" : "SW128", pattern == 0 ? "row" : "col", use_bo ? "set" : "0");
            for (int shift = 0; shift < 20; ++shift) {
                int bad = 0;
                for (int kofs = 0; kofs < (split ? 2 : 4); ++kofs) {
                    probe<<<1, 128, smem>>>(split ? map3 : map, dout, shift, kofs, use_bo, split);
                    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                    cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < 16; ++n) {
                            const float want = pattern == 0 ? (float)(m + shift) : (float)(kofs * 16 + n);
                            if (ho[m * 16 + n] != want) ++bad;
                        }
                }
                printf("%d:%s ", shift, bad ? "BAD" : "ok");
            }
            printf("\n");
        }
    }
    return 0;
}
