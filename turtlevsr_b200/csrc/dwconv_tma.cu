// Depthwise 3x3 with TMA halo staging (sm_100a).
//
// Work item = (image, 8x16 pixel tile, 32-channel block).  The (8+2)x(16+2)x32 halo box of the
// channels-last input is fetched by ONE cp.async.bulk.tensor.4d (two for the gated variant, which
// needs channel blocks c and c+C/2); out-of-bounds rows/columns are zero-filled by the TMA unit,
// which is the conv's zero padding.  CTAs are persistent over items with a 2-deep mbarrier ring, so
// the next halo streams in while the current one is consumed from shared memory.  Each thread
// produces a 4-row column of one float4 channel group with a register sliding window (18 LDS.128
// per 4 outputs) and stores full 128-byte channel runs.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

bool turtle_get_tmap2(CUtensorMap *out, const void *ptr, int rank, const uint64_t *dims, const uint64_t *strides,
                      const uint32_t *box, int swizzle128, int dtype);

namespace {

constexpr int TH = 8, TW = 16, CK = 32;
constexpr int HR = TH + 2, HC = TW + 2;
constexpr int BOX_BYTES = HR * HC * CK * 4;   // 23040 (fp32); the fp16 variant uses half of it

struct alignas(64) DwParams {
    CUtensorMap map;
    const float *w9, *bias;
    float *y;
    __half *y16;               // layout 1 only, nullable: fp16 copy of the patch rows, element for element
    int ldy, NB, H, W, C, Cout, fuse, layout, ws, rnd;
    int tiles_x, tiles_y, cblocks;
    long long nitems;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void fma4(float4 &a, const float4 &x, const float4 &w) {
    a.x = fmaf(x.x, w.x, a.x);
    a.y = fmaf(x.y, w.y, a.y);
    a.z = fmaf(x.z, w.z, a.z);
    a.w = fmaf(x.w, w.w, a.w);
}

template <int FUSE, bool FAST, bool IO16>
__global__ void __launch_bounds__(256, 2) dwconv_tma_kernel(const __grid_constant__ DwParams p) {
    constexpr int NS = FUSE == 2 ? 2 : 1;
    constexpr int BOXB = IO16 ? BOX_BYTES / 2 : BOX_BYTES;
    pdl_trigger();
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[2];
    const uint32_t smem0 = (smem_u32(smem_raw) + 127u) & ~127u;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < 2; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full_bar[s])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();

    auto decode = [&](unsigned it, int &cb, int &tx, int &ty, int &nb) {     // 32-bit: 64-bit div is ~10x dearer
        cb = (int)(it % (unsigned)p.cblocks);
        unsigned t = it / (unsigned)p.cblocks;
        tx = (int)(t % (unsigned)p.tiles_x);
        t /= (unsigned)p.tiles_x;
        ty = (int)(t % (unsigned)p.tiles_y);
        nb = (int)(t / (unsigned)p.tiles_y);
    };
    auto issue = [&](unsigned it, int buf) {
        int cb, tx, ty, nb;
        decode(it, cb, tx, ty, nb);
        const uint32_t bar = smem_u32(&full_bar[buf]);
        const uint32_t dst = smem0 + buf * (NS * BOXB);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(NS * BOXB) : "memory");
#pragma unroll
        for (int s = 0; s < NS; ++s)
            asm volatile(
                "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::
                    "r"(dst + s * BOXB),
                "l"(&p.map), "r"(bar), "r"(cb * CK + s * p.Cout), "r"(tx * TW - 1), "r"(ty * TH - 1), "r"(nb)
                : "memory");
    };

    // thread role inside a tile: float4 channel group, column, and which 4-row half
    const int c4 = tid & 7, col = (tid >> 3) & 15, half = tid >> 7;
    const int Hg = p.layout ? p.H / p.ws : 1, Wg = p.layout ? p.W / p.ws : 1;

    const unsigned nitems = (unsigned)p.nitems;
    unsigned it = blockIdx.x;
    if (it < nitems && tid == 0) issue(it, 0);
    for (int k = 0; it < nitems; it += gridDim.x, ++k) {
        const int buf = k & 1;
        const unsigned nxt = it + gridDim.x;
        if (nxt < nitems && tid == 0) issue(nxt, buf ^ 1);
        int cb, tx, ty, nb;
        decode(it, cb, tx, ty, nb);
        const int c0 = cb * CK + c4 * 4;
        // wait for the halo box
        {
            const uint32_t bar = smem_u32(&full_bar[buf]);
            const uint32_t parity = (k >> 1) & 1;
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
        const int px = tx * TW + col;
        const int py0 = ty * TH + half * 4;
        const uint8_t *sbase = smem_raw + (smem0 - smem_u32(smem_raw)) + buf * (NS * BOXB);
        // smem box layout [HR][HC][CK]; output (r, col) reads rows r..r+2, cols col..col+2.
        // One channel set at a time (the gated variant has two) keeps the register window small.
        float4 acc[NS][4];
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            float4 wv[9];
#pragma unroll
            for (int t = 0; t < 9; ++t)
                wv[t] = __ldg(reinterpret_cast<const float4 *>(p.w9 + t * p.C + c0 + s * p.Cout));
            const float4 bv =
                p.bias ? __ldg(reinterpret_cast<const float4 *>(p.bias + c0 + s * p.Cout)) : make_float4(0, 0, 0, 0);
            const uint8_t *sb = sbase + s * BOXB;
            float4 r[3][3];
            auto ld_row = [&](int slot, int rr) {
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int e = (rr * HC + col + dx) * CK + c4 * 4;       // element index inside the box
                    if (IO16) {
                        const uint2 u = *reinterpret_cast<const uint2 *>(sb + e * 2);
                        const float2 lo = __half22float2(*reinterpret_cast<const __half2 *>(&u.x));
                        const float2 hi = __half22float2(*reinterpret_cast<const __half2 *>(&u.y));
                        r[slot][dx] = make_float4(lo.x, lo.y, hi.x, hi.y);
                    } else {
                        r[slot][dx] = *reinterpret_cast<const float4 *>(sb + e * 4);
                    }
                }
            };
            ld_row(0, half * 4);
            ld_row(1, half * 4 + 1);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ld_row(2, half * 4 + i + 2);
                float4 a = bv;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) fma4(a, r[ky][kx], wv[ky * 3 + kx]);
                acc[s][i] = a;
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    r[0][dx] = r[1][dx];
                    r[1][dx] = r[2][dx];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float4 o = acc[0][i];
            if (FUSE == 1) {
                o.x = gelu_sel<FAST>(o.x); o.y = gelu_sel<FAST>(o.y); o.z = gelu_sel<FAST>(o.z); o.w = gelu_sel<FAST>(o.w);
            } else if (FUSE == 2) {
                const float4 g = acc[NS - 1][i];
                o.x = gelu_sel<FAST>(o.x) * g.x; o.y = gelu_sel<FAST>(o.y) * g.y;
                o.z = gelu_sel<FAST>(o.z) * g.z; o.w = gelu_sel<FAST>(o.w) * g.w;
            }
            if (!IO16 && p.rnd) o = rna_tf32(o);
            const int py = py0 + i;
            if (py < p.H && px < p.W) {
                if (IO16) {
                    __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
                    uint2 pk = make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
                    *reinterpret_cast<uint2 *>(reinterpret_cast<__half *>(p.y) +
                                               (((long long)nb * p.H + py) * p.W + px) * p.ldy + c0) = pk;
                } else if (p.layout == 0) {
                    stg_stream(p.y + (((long long)nb * p.H + py) * p.W + px) * p.ldy + c0, o);
                } else {
                    long long n = (long long)(py % Hg) * Wg + (px % Wg);
                    long long e = ((long long)(py / Hg) * p.ws + (px / Wg)) * p.Cout + c0;
                    const long long oo = (((long long)nb * Hg * Wg + n) * p.ws * p.ws) * p.Cout + e;
                    stg_stream(p.y + oo, o);
                    if (p.y16) {
                        __half2 h0 = __floats2half2_rn(o.x, o.y), h1 = __floats2half2_rn(o.z, o.w);
                        *reinterpret_cast<uint2 *>(p.y16 + oo) =
                            make_uint2(*reinterpret_cast<uint32_t *>(&h0), *reinterpret_cast<uint32_t *>(&h1));
                    }
                }
            }
        }
        __syncthreads();   // everyone is done with `buf` before the next iteration's prefetch refills it
    }
}

template <int FUSE, bool FAST, bool IO16>
int launch(const DwParams &p, cudaStream_t s) {
    constexpr int NS = FUSE == 2 ? 2 : 1;
    const size_t smem = 2 * NS * (IO16 ? BOX_BYTES / 2 : BOX_BYTES) + 128;
    static bool configured_[TURTLE_MAX_DEVICES] = {};      // cudaFuncSetAttribute is a per-device property
    const int dev_ = turtle_device();
    bool &configured = configured_[dev_];
    static int nsm_[TURTLE_MAX_DEVICES];
    int &nsm = nsm_[dev_];
    if (!configured) {
        if (cudaFuncSetAttribute(dwconv_tma_kernel<FUSE, FAST, IO16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess)
            return TURTLE_ELAUNCH;
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    const int per_sm = NS == 2 ? 2 : 4;
    long long grid = (long long)nsm * per_sm;
    if (grid > p.nitems) grid = p.nitems;
    launch_pdl(dwconv_tma_kernel<FUSE, FAST, IO16>, dim3((unsigned)grid), dim3(256), smem, s, p);
    return cudaGetLastError() == cudaSuccess ? TURTLE_OK : TURTLE_ELAUNCH;
}

}  // namespace

// returns TURTLE_ENOTSUP when the shape does not fit the tiled kernel (caller falls back)
int turtle_dwconv3x3_tma_ex(const float *x, int ldx, const float *w, const float *bias, float *y, void *y16, int ldy, int NB,
                            int H, int W, int C, int fuse, int layout, int ws, int rnd, void *stream);
int turtle_dwconv3x3_tma(const float *x, int ldx, const float *w, const float *bias, float *y, int ldy, int NB, int H,
                         int W, int C, int fuse, int layout, int ws, int rnd, void *stream) {
    return turtle_dwconv3x3_tma_ex(x, ldx, w, bias, y, nullptr, ldy, NB, H, W, C, fuse, layout, ws, rnd, stream);
}

// y16 (nullable, layout 1 only): the patch rows are written a second time as fp16
int turtle_dwconv3x3_tma_ex(const float *x, int ldx, const float *w, const float *bias, float *y, void *y16, int ldy, int NB,
                            int H, int W, int C, int fuse, int layout, int ws, int rnd, void *stream) {
    const int Cout = fuse == 2 ? C / 2 : C;
    if (Cout % CK) return TURTLE_ENOTSUP;
    const bool io16 = rnd == 2;
    if (io16 && layout != 0) return TURTLE_ENOTSUP;
    if (y16 && (layout != 1 || io16 || ((uintptr_t)y16 & 7))) return TURTLE_ENOTSUP;
    const uint64_t es = io16 ? 2 : 4;
    DwParams p{};
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)NB};
    uint64_t str[3] = {(uint64_t)ldx * es, (uint64_t)ldx * es * W, (uint64_t)ldx * es * W * H};
    uint32_t box[4] = {CK, HC, HR, 1};
    if (!turtle_get_tmap2(&p.map, x, 4, dims, str, box, 0, io16 ? 1 : 0)) return TURTLE_ENOTSUP;
    p.w9 = w; p.bias = bias; p.y = y; p.ldy = ldy;
    p.y16 = reinterpret_cast<__half *>(y16);
    p.NB = NB; p.H = H; p.W = W; p.C = C; p.Cout = Cout; p.fuse = fuse; p.layout = layout; p.ws = ws; p.rnd = rnd;
    p.tiles_x = (W + TW - 1) / TW;
    p.tiles_y = (H + TH - 1) / TH;
    p.cblocks = Cout / CK;
    p.nitems = (long long)p.cblocks * p.tiles_x * p.tiles_y * NB;
    cudaStream_t s = as_stream(stream);
    if (p.nitems >= (1LL << 31)) return TURTLE_ENOTSUP;
    // rnd marks the tensor-core (tf32) mode: its 1e-3-class numerics admit the 1.5e-7-accurate fast erf
    if (io16) return fuse == 0 ? launch<0, false, true>(p, s) : fuse == 1 ? launch<1, true, true>(p, s) : launch<2, true, true>(p, s);
    if (fuse == 0) return launch<0, false, false>(p, s);
    if (fuse == 1) return rnd ? launch<1, true, false>(p, s) : launch<1, false, false>(p, s);
    return rnd ? launch<2, true, false>(p, s) : launch<2, false, false>(p, s);
}
