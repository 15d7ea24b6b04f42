// Depthwise 3x3 on fp16 channels-last maps (tensor-core mode intermediates), sm_100a.
//
// Same item / TMA halo staging as dwconv_tma.cu -- work item = (image, 8x16 pixel tile, 32-channel block),
// one cp.async.bulk.tensor.4d halo box per item (two for the gated variant), hardware zero padding,
// persistent CTAs with a 4-deep mbarrier ring -- but the kernel is built around its real bound, the SM issue
// rate (measured: the fp32-register version spent 34 instructions per input element, 67 % issue-active):
//   * products run as FHFMA (fma.rn.f32.f16: fp16 x fp16 + fp32 -> fp32, near FFMA rate on B200), so the halo
//     needs no half->float converts and the 3x3 register window and the taps are half as many registers;
//   * GELU is Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7) on packed fp32 ops: 8 issue slots per element, two of them MUFU;
//   * the elected thread decodes the item once and publishes (cb, tx, ty, nb) through shared memory; the
//     mbarrier that guards the halo also orders that write.
// Taps arrive as fp16 [9, C] (same 10-bit significand as the TF32 weights of the GEMMs), bias as fp32.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

bool turtle_get_tmap2(CUtensorMap *out, const void *ptr, int rank, const uint64_t *dims, const uint64_t *strides,
                      const uint32_t *box, int swizzle128, int dtype);

namespace {

constexpr int TH = 8, TW = 16, CK = 32;
constexpr int HR = TH + 2, HC = TW + 2;
constexpr int BOXB = HR * HC * CK * 2;        // 11520 B
// halo boxes in flight per CTA = stages-1 (HBM latency x bandwidth needs > 40 KB in flight per SM)
// PAIR (FUSE 0 / 1 only): one work item covers TWO adjacent 32-channel blocks of the same pixel tile, staged as two halo
// boxes like the gated variant's (u1, u2) pair.  The item bookkeeping (barrier wait, item decode, pointer and bounds
// arithmetic, the block barrier) is ~250 of the ~450 warp instructions of a plain single-block item but only ~130 of the
// gated variant's 711; pairing the blocks amortises it over twice the arithmetic.
template <int FUSE, bool PAIR> struct Cfg {
    static constexpr int NS = (FUSE == 2 || PAIR) ? 2 : 1;
    static constexpr int STAGES = NS == 2 ? 3 : 4;
    static constexpr int CTAS = (FUSE == 0 && !PAIR) ? 4 : 3;          // resident CTAs per SM (registers / shared memory)
};

struct alignas(64) Dw16Params {
    CUtensorMap map;
    const __half *w9;
    const float *bias;
    __half *y;
    int ldy, H, W, C, Cout;
    int tiles_x, tiles_y, cblocks;
    unsigned nitems;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
    float r;
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(r) : "h"(a), "h"(b), "f"(c));
    return r;
}
__device__ __forceinline__ unsigned short lo16(uint32_t v) { return (unsigned short)(v & 0xffffu); }
__device__ __forceinline__ unsigned short hi16(uint32_t v) { return (unsigned short)(v >> 16); }

// acc[4 channels] += x[4 halves] * w[4 halves]
__device__ __forceinline__ void fma4h(float4 &a, const uint2 &x, const uint2 &w) {
    a.x = fhfma(lo16(x.x), lo16(w.x), a.x);
    a.y = fhfma(hi16(x.x), hi16(w.x), a.y);
    a.z = fhfma(lo16(x.y), lo16(w.y), a.z);
    a.w = fhfma(hi16(x.y), hi16(w.y), a.w);
}

template <int FUSE, bool PAIR>
__global__ void __launch_bounds__(256, Cfg<FUSE, PAIR>::CTAS) dwconv16_kernel(const __grid_constant__ Dw16Params p) {
    constexpr int NS = Cfg<FUSE, PAIR>::NS, STAGES = Cfg<FUSE, PAIR>::STAGES;
    // channel offset of stream s of channel-block index cb: the gated variant pairs block cb of u1 with block cb of u2
    // (Cout channels further); PAIR takes the adjacent blocks 2cb, 2cb+1
    auto chan_of = [&](int cb, int s) { return FUSE == 2 ? cb * CK + s * p.Cout : (PAIR ? (2 * cb + s) * CK : cb * CK); };
    // the host sizes the grid as a multiple of cblocks, so a CTA's channel block never changes: its taps and
    // bias are staged in shared memory once
    __shared__ __align__(16) __half wsm[NS * 9 * CK];
    __shared__ __align__(16) float bsm[NS * CK];
    pdl_trigger();
    if (threadIdx.x == 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map) : "memory");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ int4 coord[STAGES];
    const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
    const uint8_t *sgen = smem_raw + (smem0 - smem_u32(smem_raw));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full_bar[s])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const int cb0 = (int)(blockIdx.x % (unsigned)p.cblocks);
        for (int i = tid; i < NS * 9 * CK; i += 256) {
            const int s = i / (9 * CK), t = (i / CK) % 9, ch = i % CK;
            wsm[i] = p.w9[t * p.C + chan_of(cb0, s) + ch];
        }
        if (tid < NS * CK) bsm[tid] = p.bias ? p.bias[chan_of(cb0, tid / CK) + (tid % CK)] : 0.f;
    }
    __syncthreads();
    pdl_wait();          // taps and bias are constants; the map is the first thing an earlier kernel may still write

    auto issue = [&](unsigned it, int buf) {     // elected thread only
        const int cb = (int)(it % (unsigned)p.cblocks);
        unsigned t = it / (unsigned)p.cblocks;
        const int tx = (int)(t % (unsigned)p.tiles_x);
        t /= (unsigned)p.tiles_x;
        const int ty = (int)(t % (unsigned)p.tiles_y);
        const int nb = (int)(t / (unsigned)p.tiles_y);
        coord[buf] = make_int4(cb, tx, ty, nb);
        const uint32_t bar = smem_u32(&full_bar[buf]);
        const uint32_t dst = smem0 + buf * (NS * BOXB);
        // the arrive has release semantics: coord[buf] is visible to every thread that observes the phase flip
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(NS * BOXB) : "memory");
#pragma unroll
        for (int s = 0; s < NS; ++s)
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                    "r"(dst + s * BOXB),
                "l"(&p.map), "r"(bar), "r"(chan_of(cb, s)), "r"(tx * TW - 1), "r"(ty * TH - 1), "r"(nb)
                : "memory");
    };

    // thread role inside a tile: 4-channel group, column, and which 4-row half
    const int c4 = tid & 7, col = (tid >> 3) & 15, half = tid >> 7;
    const uint32_t toff = (uint32_t)(((half * 4) * HC + col) * CK + c4 * 4) * 2;     // byte offset of window origin

    unsigned it = blockIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES - 1; ++s)
            if (it + s * gridDim.x < p.nitems) issue(it + s * gridDim.x, s);
    }
    for (int k = 0; it < p.nitems; it += gridDim.x, ++k) {
        const int buf = k % STAGES;
        const unsigned nxt = it + (STAGES - 1) * gridDim.x;
        if (nxt < p.nitems && tid == 0) issue(nxt, (k + STAGES - 1) % STAGES);
        {
            const uint32_t bar = smem_u32(&full_bar[buf]);
            const uint32_t parity = (k / STAGES) & 1;
            uint32_t ok;
            do {
                asm volatile(
                    "{\n\t.reg .pred q;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n\t"
                    "selp.u32 %0, 1, 0, q;\n\t}"
                    : "=r"(ok)
                    : "r"(bar), "r"(parity)
                    : "memory");
            } while (!ok);
        }
        const int4 cd = coord[buf];
        const int px = cd.y * TW + col;
        const int py0 = cd.z * TH + half * 4;
        const uint8_t *sbase = sgen + buf * (NS * BOXB) + toff;

        float4 out[4];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            uint2 wv[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) wv[t] = *reinterpret_cast<const uint2 *>(wsm + (s * 9 + t) * CK + c4 * 4);
            const float4 bv = *reinterpret_cast<const float4 *>(bsm + s * CK + c4 * 4);
            const uint8_t *sb = sbase + s * BOXB;
            uint2 r[3][3];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                r[0][dx] = *reinterpret_cast<const uint2 *>(sb + (0 * HC + dx) * CK * 2);
                r[1][dx] = *reinterpret_cast<const uint2 *>(sb + (1 * HC + dx) * CK * 2);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx)
                    r[2][dx] = *reinterpret_cast<const uint2 *>(sb + ((i + 2) * HC + dx) * CK * 2);
                float4 a = bv;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) fma4h(a, r[ky][kx], wv[ky * 3 + kx]);
                if (FUSE == 0) {
                    out[i] = a;
                } else if (FUSE == 1) {
                    const float2 g0 = gelu_fast2(make_float2(a.x, a.y)), g1 = gelu_fast2(make_float2(a.z, a.w));
                    out[i] = make_float4(g0.x, g0.y, g1.x, g1.y);
                } else if (s == 0) {
                    const float2 g0 = gelu_fast2(make_float2(a.x, a.y)), g1 = gelu_fast2(make_float2(a.z, a.w));
                    out[i] = make_float4(g0.x, g0.y, g1.x, g1.y);
                } else {
                    const float2 m0 = f2_mul(make_float2(out[i].x, out[i].y), make_float2(a.x, a.y));
                    const float2 m1 = f2_mul(make_float2(out[i].z, out[i].w), make_float2(a.z, a.w));
                    out[i] = make_float4(m0.x, m0.y, m1.x, m1.y);
                }
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    r[0][dx] = r[1][dx];
                    r[1][dx] = r[2][dx];
                }
            }
            // plain / GELU streams are independent outputs: stored as soon as they are computed; the gated pair after both
            if ((FUSE != 2 || s == NS - 1) && px < p.W) {
                const int c0 = (FUSE == 2 ? cd.x * CK : chan_of(cd.x, s)) + c4 * 4;
                __half *yp = p.y + (((long long)cd.w * p.H + py0) * p.W + px) * p.ldy + c0;
                const long long rstride = (long long)p.W * p.ldy;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (py0 + i < p.H) {
                        const __half2 h0 = __floats2half2_rn(out[i].x, out[i].y), h1 = __floats2half2_rn(out[i].z, out[i].w);
                        *reinterpret_cast<uint2 *>(yp + i * rstride) =
                            make_uint2(*reinterpret_cast<const uint32_t *>(&h0), *reinterpret_cast<const uint32_t *>(&h1));
                    }
                }
            }
        }
        __syncthreads();   // everyone is done with `buf` (and coord[buf]) before the next prefetch refills it
    }
}

template <int FUSE, bool PAIR>
int launch16(const Dw16Params &p, cudaStream_t s) {
    constexpr int NS = Cfg<FUSE, PAIR>::NS;
    const size_t smem = Cfg<FUSE, PAIR>::STAGES * NS * BOXB + 128;
    static bool configured_[TURTLE_MAX_DEVICES] = {};      // cudaFuncSetAttribute is a per-device property
    const int dev_ = turtle_device();
    bool &configured = configured_[dev_];
    static int nsm_[TURTLE_MAX_DEVICES];
    int &nsm = nsm_[dev_];
    if (!configured) {
        if (cudaFuncSetAttribute(dwconv16_kernel<FUSE, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return TURTLE_ELAUNCH;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    unsigned grid = (unsigned)nsm * Cfg<FUSE, PAIR>::CTAS;
    if (grid > p.nitems) grid = p.nitems;
    grid -= grid % (unsigned)p.cblocks;          // every CTA keeps one channel block (nitems is a multiple of cblocks)
    launch_pdl(dwconv16_kernel<FUSE, PAIR>, dim3(grid), dim3(256), smem, s, p);
    return cudaGetLastError() == cudaSuccess ? TURTLE_OK : TURTLE_ELAUNCH;
}

}  // namespace

int turtle_dwconv3x3_h16_tc(const void *x, int ldx, const void *w9, const float *bias, void *y, int ldy, int NB, int H, int W,
                            int C, int fuse, void *stream);       // dwconv16_tc.cu: the same conv as tcgen05 MMAs

// x, w9, y are fp16 (ldx / ldy in halves); bias fp32.  TURTLE_ENOTSUP when the shape does not fit.
int turtle_dwconv3x3_h16(const void *x, int ldx, const void *w9, const float *bias, void *y, int ldy, int NB, int H,
                         int W, int C, int fuse, void *stream) {
    const int Cout = fuse == 2 ? C / 2 : C;
    if (!x || !w9 || !y || NB < 1 || fuse < 0 || fuse > 2) return TURTLE_EINVAL;
    // tensor-core form (dwconv16_tc.cu) only on request: TURTLE_DW_TC=N runs it for the fuse variants in bit mask N.
    // Measured 2x SLOWER than the CUDA-core kernel on a 720p frame (14.1 vs 7.2 ms of depthwise time,
    // profiles/r02s_dwconv_tc_ab.txt), so it is off by default and kept for the measurement only.
    static const int tc_mask = getenv("TURTLE_DW_TC") ? atoi(getenv("TURTLE_DW_TC")) : 0;
    if (tc_mask & (1 << fuse)) {
        const int r = turtle_dwconv3x3_h16_tc(x, ldx, w9, bias, y, ldy, NB, H, W, C, fuse, stream);
        if (r != TURTLE_ENOTSUP) return r;
    }
    if (Cout % CK || (ldx & 7) || (ldy & 3) || (((uintptr_t)x) & 15) || (((uintptr_t)y | (uintptr_t)w9) & 7))
        return TURTLE_ENOTSUP;
    Dw16Params p{};
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)ldx * 2, (uint64_t)ldx * 2 * W, (uint64_t)ldx * 2 * W * H};
    uint32_t box[4] = {CK, HC, HR, 1};
    if (!turtle_get_tmap2(&p.map, x, 4, dims, str, box, 0, 1)) return TURTLE_ENOTSUP;
    p.w9 = reinterpret_cast<const __half *>(w9);
    p.bias = bias;
    p.y = reinterpret_cast<__half *>(y);
    p.ldy = ldy; p.H = H; p.W = W; p.C = C; p.Cout = Cout;
    p.tiles_x = (W + TW - 1) / TW;
    p.tiles_y = (H + TH - 1) / TH;
    // plain / GELU maps whose channel count is a multiple of 64 run two channel blocks per item (TURTLE_DW_PAIR=0: off)
    static const bool no_pair = getenv("TURTLE_DW_PAIR") && atoi(getenv("TURTLE_DW_PAIR")) == 0;
    const bool pair = fuse != 2 && !no_pair && Cout % (2 * CK) == 0;
    p.cblocks = Cout / CK / (pair ? 2 : 1);
    const long long nitems = (long long)p.cblocks * p.tiles_x * p.tiles_y * NB;
    if (nitems >= (1LL << 31)) return TURTLE_ENOTSUP;
    p.nitems = (unsigned)nitems;
    cudaStream_t s = as_stream(stream);
    if (fuse == 2) return launch16<2, false>(p, s);
    if (pair) return fuse == 0 ? launch16<0, true>(p, s) : launch16<1, true>(p, s);
    return fuse == 0 ? launch16<0, false>(p, s) : launch16<1, false>(p, s);
}
