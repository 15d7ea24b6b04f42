/*
 * turtle_b200.h -- C ABI of libturtle_b200.so: the sm_100a kernels behind Turtle's inference
 * hot path (truncated Causal History Model + per-frame U-Net blocks).
 *
 * The reference (sflindrs/TurtleVSR) is pure PyTorch: every op below replaces a group of ATen
 * calls issued from basicsr/models/archs/turtle_t1_arch.py ("T1"), turtle_arch.py ("T0") or
 * turtlesuper_t1_arch.py ("TS").  The citation on each entry point is the reference code it
 * replaces.  A maintainer binds these with ctypes; nothing here takes or
 * returns a torch type.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless stated; activations are channels-last
 *     ("NHWC"): element (b,y,x,c) of a [B,H,W,C] map lives at ((b*H+y)*W+x)*ld + c, where the
 *     row pitch ld (in floats) lets a call address a channel slice of a wider tensor;
 *   - every function enqueues work on `stream` (a cudaStream_t passed as void*) and returns
 *     immediately: 0 on success, a negative TURTLE_E* code on bad arguments or launch failure.
 *     Nothing allocates, synchronises or throws; the caller owns all buffers;
 *   - `round_tf32` (producer kernels) / `round_out` (gemm): in TURTLE_TF32 mode the host asks producers whose
 *     output feeds a tensor-core contraction to round it to nearest TF32 (cvt.rna), so that the tensor core's
 *     operand truncation is exact and unbiased; pass 0 in TURTLE_FP32 mode.  The same flag (and TURTLE_TF32 in
 *     turtle_gemm) selects a 1.5e-7-accurate polynomial erf for the fused GELUs instead of erff;
 *   - `mode`: TURTLE_FP32 = CUDA-core fp32 FMA (exact mode: bit-exact top-k contract),
 *             TURTLE_TF32 = tcgen05 tensor cores, TF32 operands, fp32 accumulate in TMEM.
 */
#ifndef TURTLE_B200_H
#define TURTLE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TURTLE_OK 0
#define TURTLE_EINVAL (-1)   /* bad shape / alignment / unsupported combination */
#define TURTLE_ELAUNCH (-2)  /* cudaGetLastError() != cudaSuccess after the launch */
#define TURTLE_ENOTSUP (-3)  /* shape not supported by the requested mode (caller picks TURTLE_FP32) */

#define TURTLE_FP32 0
#define TURTLE_TF32 1

#define TURTLE_ACT_NONE 0
#define TURTLE_ACT_GELU 1    /* exact erf GELU (F.gelu default) */

#define TURTLE_STORE_PLAIN 0
#define TURTLE_STORE_UNSHUFFLE2 1  /* nn.PixelUnshuffle(2) folded into the store (T1:140-141) */
#define TURTLE_STORE_SHUFFLE2 2    /* nn.PixelShuffle(2) folded into the store   (T1:150-151) */

#define TURTLE_MAX_SEG 48
#define TURTLE_SAB_SLOTS 48  /* 5 top-k + 41 local-window entries, padded to 48 */

int turtle_abi_version(void);          /* bumps when a signature changes */
int turtle_sizeof_gemm_args(void);       /* sizeof(TurtleGemmArgs) as compiled: bindings check their mirror of the struct */

/* fp32 -> fp16 copy of a dense buffer (n % 8 == 0): the residual stream as the fp16 A operand of the dense 3x3
 * resampling convs (Downsample / Upsample, T1:136-154) in tensor-core mode. */
int turtle_cast_f16(const float *x, void *y, int64_t n, void *stream);
const char *turtle_build_info(void);   /* "sm_100a nvcc <ver> ..." */

/* ---------------------------------------------------------------------------------------
 * Frame entry / exit
 * ------------------------------------------------------------------------------------- */

/* check_image_size zero pad (T1:1134-1140) + NCHW->NHWC of the selected frame inp[:,1]
 * (T1:1059-1061); with upscale==4 also nn.Upsample(x4, bilinear, align_corners=False)
 * (TS:975-978, 1063-1070).  src: [B,C,Hs,Ws] with batch stride src_bstride (floats);
 * dst: [B,Hp,Wp,C] dense, Hp>=Hs*upscale, Wp>=Ws*upscale, zero outside. */
int turtle_pack_frame(const float *src, int64_t src_bstride, float *dst, int B, int C, int Hs, int Ws,
                      int Hp, int Wp, int upscale, void *stream);

/* input_projection: 3x3 conv, tiny Cin (3 or 6), zero pad 1 (T1:975-978, 1063).
 * x [B,H,W,Cin] dense; w [Cout,Cin,3,3] (reference layout); bias nullable; y [B,H,W,Cout]. */
int turtle_conv3x3_first(const float *x, const float *w, const float *bias, float *y, int B, int H, int W,
                         int Cin, int Cout, void *stream);

/* ending: 3x3 conv Cin->Cout(<=4) + bias + residual `current` + crop, written NCHW (T1:1128-1132).
 * x [B,H,W,Cin]; w [Cout,Cin,3,3]; cur [B,H,W,cur_ld] (first Cout channels used);
 * out [B,Cout,Hc,Wc] dense NCHW with Hc<=H, Wc<=W. */
int turtle_conv3x3_last(const float *x, const float *w, const float *bias, const float *cur, int cur_ld,
                        int cur_coff, float *out, int B, int H, int W, int Cin, int Cout, int Hc, int Wc,
                        void *stream);

/* ---------------------------------------------------------------------------------------
 * Per-pixel channel LayerNorm (WithBias_LayerNorm / BiasFree_LayerNorm, T1:67-112)
 *   y[p,:] = (x[p,:]-mu)/sqrt(var_biased+1e-5)*w + b      (b==NULL: BiasFree, no mean subtraction
 *   in the numerator, exactly as T1:79-81)
 * ------------------------------------------------------------------------------------- */
int turtle_layernorm(const float *x, int ldx, const float *w, const float *b, float *y, int ldy, int64_t P,
                     int C, int round_tf32, void *stream);

/* ---------------------------------------------------------------------------------------
 * Contractions over channels: 1x1 convs, the skip-cat + reduce_chan, the folded
 * channel-attention apply, and (im2col==1) the dense 3x3 convs of Down/Upsample.
 *
 *   acc[p,o] = sum_k A(p,k) * W[o,k]                        k in [0, nseg*segw)
 *   v        = acc + bias[o];  v = act(v);  v *= scale[o];  v += res[p*ldres+o]
 *   out      = store(v)
 *
 * A(p,k): im2col==0 -> segment s=k/segw: A[s][p*lda[s] + k%segw]   (K-concatenated sources:
 *                      torch.cat([up, enc],1) T1:1098; history rows T1:272-273)
 *         im2col==1 -> one segment, k=(tap*Cin+c), zero-padded 3x3 neighbourhood of pixel p in
 *                      an [B,H,W,Cin] map (T1:140,150); W is [Cout, 9*Cin] tap-major.
 * Replaces: nn.Conv2d 1x1 everywhere (T1:166,171,188,193,236,238,298,301,305,307,310,623,674,
 * 676,708,724,1008,1015,1023), the bias/GELU/beta/gamma/residual elementwise ops around them
 * (T1:176,206,210,739,742,808-810) and attn@v + project_out (T1:697-701, 280-284).
 * ------------------------------------------------------------------------------------- */
typedef struct TurtleGemmArgs {
    int32_t mode;        /* TURTLE_FP32 | TURTLE_TF32 */
    int32_t im2col;      /* 0 | 1 */
    int64_t P;           /* rows (pixels, batch folded in) */
    int32_t B, H, W;     /* geometry, used by im2col and by the (un)shuffle stores */
    int32_t Cout;
    int32_t nseg, segw;  /* K = nseg*segw (im2col: nseg=1, segw=Cin, K=9*Cin) */
    const float *A[TURTLE_MAX_SEG];
    int32_t lda[TURTLE_MAX_SEG];
    const float *Wt;     /* [Cout, K] row-major */
    const float *bias;   /* [Cout] or NULL */
    const float *scale;  /* [Cout] or NULL */
    int32_t act;         /* TURTLE_ACT_* */
    const float *res;    /* or NULL */
    int32_t ldres;
    float *out;
    int32_t ldo;
    int32_t store;       /* TURTLE_STORE_* */
    int32_t round_out;   /* !=0: round the stored values to nearest TF32 (they feed another tensor-core op) */
    int32_t a_dtype;     /* 0: A segments and Wt are fp32 (TF32 MMA); 1: they are fp16 (kind::f16 MMA). TURTLE_TF32 only;
                            lda counts elements of that type, the pointers are carried in the float* slots */
    int32_t out_dtype;   /* 0: out is fp32; 1: out is fp16 (ldo in halves; no residual, plain store) */
    /* Fused WithBias LayerNorm of the rows just produced (T1:96-112: the norm1/norm2 that reads this output next):
       ln_out[p,:] = fp16( (out[p,:]-mu)/sqrt(var_biased+1e-5)*ln_w + ln_b ),  ld_ln in halves.
       TURTLE_TF32 only; needs Cout in {64,128,256}, a residual, a plain fp32 store; else TURTLE_ENOTSUP. */
    void *ln_out;        /* or NULL */
    int32_t ld_ln;
    const float *ln_w, *ln_b;
    /* Per-batch weights (the folded channel-attention matrices of B images in one launch): rows
       [b*rows_per_batch, (b+1)*rows_per_batch) are multiplied by the matrix at Wt + b*w_bstride (elements of Wt's type).
       w_batches <= 1: one matrix for all rows.  TURTLE_TF32 only, no im2col, rows_per_batch % 128 == 0,
       P == w_batches*rows_per_batch; TURTLE_ENOTSUP otherwise (the caller then launches once per image). */
    int32_t w_batches;
    int32_t reserved_;
    int64_t w_bstride;
    int64_t rows_per_batch;
} TurtleGemmArgs;

int turtle_gemm(const TurtleGemmArgs *args, void *stream);

/* ---------------------------------------------------------------------------------------
 * Depthwise 3x3, stride 1, zero pad 1 (T1:168-170, 237, 299, 302, 624, 675, 716-722)
 *   fuse: 0 plain(+bias) | 1 GELU(dw+bias)  (ReducedAttn T1:738-739)
 *         2 gate: y[:,j] = GELU(dw[:,j]) * dw[:,j+C/2], y has C/2 channels (GFFW T1:175-176)
 *   layout: 0 NHWC | 1 SAB "dilated patch" rows 'b d (p1 h)(p2 w) -> b (h w) (p1 p2 d)' (T1:573),
 *           y is then [NB, (H/ws)*(W/ws), ws*ws*C] dense.
 * x [NB,H,W,C] pitch ldx; w tap-major [9,C] (= weight.view(C,9).t(), packed once by the host);
 * bias nullable.
 * round_tf32: 0 exact | 1 round the stored values to nearest TF32 | 2 fp16 maps: x, y AND the taps w are fp16
 *   (ldx/ldy in halves, pointers carried in the float* slots; layout 0 and 32-aligned channel counts only),
 *   products accumulate in fp32.
 * ------------------------------------------------------------------------------------- */
int turtle_dwconv3x3(const float *x, int ldx, const float *w, const float *bias, float *y, int ldy, int NB,
                     int H, int W, int C, int fuse, int layout, int ws, int round_tf32, void *stream);
/* layout 1 (dilated patch rows, the StateAlignBlock value rows T1:571-574) with a second, fp16 copy of every row
 * written by the same kernel: y fp32 [NB, N, ws*ws*C], y16 fp16 likewise.  The tensor-core aggregation reads the copy
 * (turtle_sab_aggregate_tc, v_dtype 1).  C % 32 == 0; TURTLE_ENOTSUP otherwise (caller: turtle_dwconv3x3 + turtle_cast_f16). */
int turtle_dwconv3x3_patch_rows(const float *x, int ldx, const float *w, const float *bias, float *y, void *y16, int NB,
                                int H, int W, int C, int ws, void *stream);

/* ---------------------------------------------------------------------------------------
 * Transposed (channel) attention: ChannelAttention T1:680-702, FrameHistoryRouter T1:243-286,
 * and the router inside CausalHistoryModel T1:649-660.
 *
 * Step 1 (per key segment): per-head Gram over pixels plus squared column norms
 *     G[h,i,j] = sum_p q[p, h*q_hs+i] * k[p, h*k_hs+j]     i,j in [0,ch)
 *   written as `nsplit` partial sums (deterministic two-stage reduction):
 *     gpart [nsplit, heads, ch, ch], sqq [nsplit, heads*ch], sqk [nsplit, heads*ch]
 * Step 2: P = softmax_j over all segments( G / (max(|q_i|,1e-12) * max(|k_j|,1e-12)) * temp[h] )
 *   (F.normalize folded in as diagonal scaling; seg_prenorm[s]!=0 means the key rows of segment s
 *   were cached already normalised, T1:272 -> divisor 1).
 * Step 3: M[o, s*C + h*ch + j] = sum_i Wo[o, h*ch+i] * P[h,i,s*ch+j]   (project_out folded in)
 *   so that out = res + M @ [v_seg0; v_seg1; ...] is one turtle_gemm.
 * ------------------------------------------------------------------------------------- */
int turtle_chan_gram(const float *q, int ldq, int q_hs, const float *k, int ldk, int k_hs, int64_t P, int heads,
                     int ch, int nsplit, float *gpart, float *sqq, float *sqk, int mode, void *stream);

int turtle_chan_softmax(const float *gpart, const float *sqq, const float *sqk, const int32_t *seg_prenorm,
                        const float *temperature, int nseg, int nsplit, int heads, int ch, float *Pout,
                        float *inv_knorm /* [nseg, heads*ch], 1/max(|k|,1e-12) (1 for prenormalised) */,
                        void *stream);

int turtle_chan_fold(const float *Pm, const float *Wo, int nseg, int heads, int ch, float *M, int round_tf32,
                     void *stream);

/* Batched forms of the three steps above: B batch elements per launch (the tiles of tiled inference, training
 * batches), element b of every operand at a fixed stride from element 0, so the launch count of a channel-attention
 * block no longer scales with B.
 *   gram_b:    q_bs / k_bs = element strides of the q / k maps between batch elements; gpart [B][nsplit,heads,ch,ch] at
 *              g_bs floats, sqq / sqk [B][nsplit,C] at s_bs floats (one call per key segment, as turtle_chan_gram);
 *   softmax_b: gpart [B][nseg,nsplit,heads,ch,ch] at g_bs, sqq / sqk [B][nseg,nsplit,C] at s_bs; Pout [B,heads,ch,nseg*ch]
 *              and inv_knorm [B,nseg,C] dense;
 *   fold_b:    Pm [B,heads,ch,nseg*ch] -> M [B,C,nseg*C] dense. */
int turtle_chan_gram_b(const float *q, int ldq, int q_hs, int64_t q_bs, const float *k, int ldk, int k_hs, int64_t k_bs,
                       int64_t P, int heads, int ch, int nsplit, float *gpart, float *sqq, float *sqk, int64_t g_bs,
                       int64_t s_bs, int B, int mode, void *stream);
int turtle_chan_softmax_b(const float *gpart, const float *sqq, const float *sqk, const int32_t *seg_prenorm,
                          const float *temperature, int nseg, int nsplit, int heads, int ch, float *Pout, float *inv_knorm,
                          int64_t g_bs, int64_t s_bs, int B, void *stream);
int turtle_chan_fold_b(const float *Pm, const float *Wo, int nseg, int heads, int ch, float *M, int round_tf32, int B,
                       void *stream);

/* y[p, h*y_hs + j] = x[p, h*x_hs + j] * s[h*ch + j]  (s==NULL: copy).  Used to push the normalised
 * key rows / raw value rows of a frame into the history ring (T1:272-273, 286). */
int turtle_scale_cols(const float *x, int ldx, int x_hs, const float *s, float *y, int ldy, int y_hs, int64_t P,
                      int heads, int ch, void *stream);

/* ---------------------------------------------------------------------------------------
 * StateAlignBlock (T1:548-610; T0:459-533 with halve=1)
 * ------------------------------------------------------------------------------------- */

/* k2_dwconv / q2_dwconv: depthwise ws x ws, stride ws, pad 1, then 'b d h w -> b (h w) d' and
 * F.normalize over d (T1:306-308, 559-560, 569-572, 577-578).
 * t [B,H,W,D] pitch ldt (the 1x1 k2/q2 output); w tap-major [ws*ws,D] (= weight.view(D,-1).t()); bias [D] or NULL
 * (the arch's `bias` option, T1:306-308), added before the normalisation; out [B, N=(H/ws)*(W/ws), D] dense,
 * batch stride out_bstride floats. */
int turtle_sab_window_reduce(const float *t, int ldt, const float *w, const float *bias, float *out, int64_t out_bstride,
                             int B, int H, int W, int D, int ws, void *stream);
/* the same on an fp16 map (ldt in halves; tensor-core mode intermediates); taps, bias, sum and output stay fp32 */
int turtle_sab_window_reduce_h16(const void *t, int ldt, const float *w, const float *bias, float *out,
                                 int64_t out_bstride, int B, int H, int W, int D, int ws, void *stream);

/* 'b d (p1 h)(p2 w) -> b (h w) (p1 p2 d)' + F.normalize, for the T0 q/k path (T0:487-498). */
int turtle_sab_patch_normalize(float *rows, int64_t n_rows, int D, void *stream);

/* Correlation + top-5 + local L1 window + clipped softmax, never materialising [F,N,N]
 * (T1:585-596, 394-416, 448-464, 115-132).
 *   qn [N,D] normalised queries; kn: F key frames, frame f at kn + f*k_fstride, each [N,D];
 *   S[f,i,j] = temp * <qn_i, kn_fj>.
 * Outputs, per (f,i): idx[f,i,0:5] = top-5 keys in descending score order, idx[f,i,5:46] = the
 * local-window keys (|dy|+|dx|<=4) not already in the top-5, -1 = empty; wgt = clipped-softmax
 * weight of each slot (entries in both sets carry the doubled logit, T1:595). */
int turtle_sab_select(const float *qn, const float *kn, int64_t k_fstride, int F, int Hg, int Wg, int D,
                      const float *temperature, int halve, int32_t *idx, float *wgt, int mode, void *stream);

/* Same contract on the tensor cores (tcgen05, 3xTF32 split product => fp32-accurate scores, top-5 kept
 * per TMEM lane in registers).  Needs D % 32 == 0, dense key frames (k_fstride == N*D) and a device
 * workspace of turtle_sab_select_tc_workspace(F,N,D) bytes; returns TURTLE_ENOTSUP otherwise. */
long long turtle_sab_select_tc_workspace(int F, int N, int D);
int turtle_sab_select_tc(const float *qn, const float *kn, int64_t k_fstride, int F, int Hg, int Wg, int D,
                         const float *temperature, int halve, int32_t *idx, float *wgt, void *workspace,
                         void *stream);

/* attn @ v with the sparse weights, un-patched straight to NHWC (T1:599-604):
 *   y[f, p1*Hg+iy, p2*Wg+ix, d] = sum_t wgt[f,i,t] * V[f][idx[f,i,t]][(p1*ws+p2)*c + d]
 * v: frame f at v + f*v_fstride, [N, ws*ws*c];  y [F,H,W,c] dense.
 * passthrough!=0 reproduces T0:523 (`out = v`): y is the un-patched V itself. */
int turtle_sab_aggregate(const int32_t *idx, const float *wgt, const float *v, int64_t v_fstride, float *y, int F,
                         int Hg, int Wg, int ws, int c, int passthrough, int round_tf32, void *stream);

/* The same aggregation on the tensor cores (tensor-core mode; csrc/sab_agg_tc.cu).  For a tile of 8 x 16 neighbouring
 * queries every local-window key lies in the 16 x 24 box of keys around the tile, so the window part is one dense
 * tcgen05 contraction per tile (TF32 operands, fp32 accumulation; the weights are TF32-rounded here, V should hold
 * TF32-rounded values -- turtle_dwconv3x3(round_tf32 = 1) writes them so); the top-k keys outside the box are added
 * by a gather pass.  round_mode: 0 fp32 y, 1 fp32 y TF32-rounded, 2 fp16 y (feeds a kind::f16 GEMM).
 * v_dtype 1: v points at an fp16 copy of the value rows (v_fstride in halves): kind::f16 MMAs and half the bytes in the
 * contraction and in the gather (the engine keeps such a copy next to the fp32 history ring in tensor-core mode).
 * Needs c % 32 == 0, (ws*ws*c) % 256 == 0, dense frames (v_fstride >= N*ws*ws*c) and a device workspace of
 * turtle_sab_aggregate_tc_workspace(F, Hg, Wg) bytes; TURTLE_ENOTSUP otherwise (the caller runs turtle_sab_aggregate). */
long long turtle_sab_aggregate_tc_workspace(int F, int Hg, int Wg);
int turtle_sab_aggregate_tc(const int32_t *idx, const float *wgt, const void *v, int v_dtype, int64_t v_fstride, void *y,
                            int F, int Hg, int Wg, int ws, int c, int round_mode, void *workspace, void *stream);

/* T0 only: x + positionalencoding2d(c,h,w) (T0:412-439, 475-476), evaluated analytically. */
int turtle_add_posenc(const float *x, float *y, int B, int H, int W, int C, void *stream);

/* ---------------------------------------------------------------------------------------
 * Training-step tail (cfg 5, VRM:78-108): one flat fp32 buffer holds all 59,079,548 parameters
 * (and one each their gradients, exp_avg, exp_avg_sq), so the optimizer is two HBM-bound passes.
 * ------------------------------------------------------------------------------------- */

/* GradScaler.unscale_'s non-finite check (VRM:101): *found (device, zeroed by the caller) becomes 1.0f if any
 * of the n gradients is inf/NaN.  g must be 16-byte aligned. */
int turtle_grad_check_finite(const float *g, int64_t n, float *found, void *stream);

/* torch.optim.AdamW single step on flat buffers (VRM:68-69, 104): g is multiplied by grad_scale first
 * (1/loss_scale x 1/world_size), p *= 1-lr*weight_decay, m/v are the exp_avg / exp_avg_sq states, `step` >= 1 is
 * the 1-based step count for the bias corrections.  found (nullable, device): the whole update is skipped when
 * *found != 0 (GradScaler.step).  All four buffers 16-byte aligned. */
int turtle_adamw_flat(float *p, const float *g, float *m, float *v, int64_t n, float lr, float beta1, float beta2,
                      float eps, float weight_decay, int step, float grad_scale, const float *found, void *stream);

/* Channel LayerNorm of the TRAINING graph (WithBias_LayerNorm, T1:83-112, as autograd sees it) on NCHW maps
 * x [B,C,H,W] dense, x_dtype 0 = fp32, 1 = fp16, 2 = bf16 (the autocast residual stream); HW = H*W.
 * fwd: y fp32 [B,C,H,W] = (x - mean) * rstd * w + b, mean / rstd [B*HW] saved for the backward (biased variance,
 *      eps 1e-5).  Replaces the reference's to_3d / mean / var / sqrt / div / mul / add / to_4d chain (T1:96-112).
 * bwd: dx (x's dtype), dw[C], db[C] from dy fp32; workspace of turtle_ln2d_bwd_workspace(C, B*HW) bytes holds the
 *      per-block partial sums, reduced in a fixed order (deterministic dw / db). */
long long turtle_ln2d_bwd_workspace(int C, long long n_pixels);
int turtle_ln2d_fwd(const void *x, int x_dtype, const float *w, const float *b, float *y, float *mean, float *rstd,
                    int B, int C, long long HW, void *stream);
int turtle_ln2d_bwd(const float *dy, const void *x, int x_dtype, const float *w, const float *mean, const float *rstd,
                    void *dx, float *dw, float *db, void *workspace, int B, int C, long long HW, void *stream);

/* The same with y (forward) / dy (backward) in a chosen dtype (0 fp32, 1 fp16, 2 bf16): under autocast the conv that
 * consumes the LayerNorm casts its output to the 16-bit type and sends the gradient back in it; writing / reading that
 * type directly is the same single rounding and saves two ATen cast launches per LayerNorm and direction. */
int turtle_ln2d_fwd_cast(const void *x, int x_dtype, const float *w, const float *b, void *y, int y_dtype, float *mean,
                         float *rstd, int B, int C, long long HW, void *stream);
int turtle_ln2d_bwd_cast(const void *dy, int dy_dtype, const void *x, int x_dtype, const float *w, const float *mean,
                         const float *rstd, void *dx, float *dw, float *db, void *workspace, int B, int C, long long HW,
                         void *stream);

/* Depthwise 3x3 (stride 1, zero pad 1, groups = C) of the TRAINING graph on NCHW maps x [B,C,H,W] dense; dtype as in
 * turtle_ln2d_fwd; w9 [C,9] fp32 taps (row-major 3x3), bias [C] fp32 or NULL.  Replaces nn.Conv2d(groups=C).forward
 * (e.g. GatedFeedForward.dwconv T1:163, qkv_dwconv T1:674) and the three kernels autograd runs for it:
 *   turtle_dwconv3x3_nchw(flip=0)          forward, y in x's dtype
 *   turtle_dwconv3x3_nchw(dy, flip=1)      input gradient (taps flipped, bias NULL)
 *   turtle_dwconv3x3_nchw_wgrad            dw9 [C,9] and db [C] (nullable); workspace of ..._wgrad_workspace bytes;
 *                                          tile partials are reduced in a fixed order (deterministic). */
int turtle_dwconv3x3_nchw(const void *x, int dtype, const float *w9, const float *bias, void *y, int B, int C, int H,
                          int W, int flip, void *stream);
long long turtle_dwconv3x3_nchw_wgrad_workspace(int B, int C, int H, int W);
int turtle_dwconv3x3_nchw_wgrad(const void *x, const void *dy, int dtype, float *dw9, float *db, void *workspace, int B,
                                int C, int H, int W, void *stream);

/* GELU gate of the TRAINING graph on NCHW maps (GatedFeedForward T1:175-176: `x1, x2 = dwconv(..).chunk(2, dim=1);
 * F.gelu(x1) * x2` and what autograd runs for its backward): u [B, 2*Ch, HW] dense, y / dy [B, Ch, HW], du like u; dtype
 * as in turtle_ln2d_fwd; intermediates are rounded to the map's dtype where the ATen chain rounds them under autocast.
 * HW % 8 == 0 and 16-byte (fp32: 32-byte) aligned pointers; TURTLE_ENOTSUP otherwise (caller: the torch ops). */
int turtle_gelu_gate_nchw(const void *u, int dtype, void *y, int B, int Ch, long long HW, void *stream);
int turtle_gelu_gate_nchw_bwd(const void *u, const void *dy, int dtype, void *du, int B, int Ch, long long HW, void *stream);

/* F.normalize(x, dim=-1) of the channel attention's q / k rows (T1:686-687) in the TRAINING graph.  Row (b, ch) of x is
 * the H*W = len contiguous elements at x + b * bstride + ch * len (a channel chunk of the NCHW qkv map); y [B*CH, len] fp32
 * (autocast runs normalize in fp32), denom [B*CH] = max(||row||, 1e-12).  Backward: dx = (dy - y <dy, y>) / denom in x's
 * dtype, [rows, len] dense.  len % 4 == 0. */
int turtle_rownorm_fwd(const void *x, int dtype, long long bstride, int B, int CH, int len, float *y, float *denom,
                       void *stream);
int turtle_rownorm_bwd(const float *dy, const float *y, const float *denom, int dtype, long long rows, int len, void *dx,
                       void *stream);

/* GatedFeedForward (T1:159-178) as one kernel, tensor-core mode:  x += W_out . ( gelu(u1) * u2 ),
 * [u1 | u2] = dw3x3( W_in . xn ), with the 5c-wide hidden map kept on the SM (project_in recomputed on the 1-pixel halo
 * of each 8x16 tile, depthwise + gate feeding the second tcgen05 contraction through shared memory).
 *   xn16    fp16 [B,H,W,C] dense: LayerNorm(x) (norm2 of the block)         w_in16  fp16 [2*hid, C]  (project_in.weight)
 *   taps16  fp16 [hid/32][2][9][32]: dwconv.weight regrouped per chunk of 32 gated channels (u1 block, u2 block; tap-major)
 *   w_out16 fp16 [C, hid] (project_out.weight)                                x       fp32 [B,H,W,C] dense, updated in place
 *   ln_out16 (nullable) fp16 [B,H,W,C]: LayerNorm(x_new) with ln_w / ln_b [C] for the norm that reads x next (T1:96-112)
 * C in {64, 128, 256}, hid % 32 == 0, no biases (the shipped ymls have bias = False); TURTLE_ENOTSUP otherwise -- the
 * caller then runs turtle_gemm -> turtle_dwconv3x3(fuse = 2) -> turtle_gemm. */
int turtle_gffw_fused(const void *xn16, const void *w_in16, const void *taps16, const void *w_out16, float *x,
                      void *ln_out16, const float *ln_w, const float *ln_b, int B, int H, int W, int C, int hid,
                      void *stream);
/* The second half of GatedFeedForward on its own:  x += W_out . ( gelu(u1) * u2 ),  [u1 | u2] = dw3x3(t16), where
 * t16 fp16 [B,H,W,2*hid] dense is the hidden map the project_in GEMM wrote (turtle_gemm, fp16 output).  Depthwise + gate
 * feed the project_out tcgen05 contraction through shared memory (halo boxes by TMA, two CTAs per SM), so the gated map
 * never reaches HBM and project_out is not a separate pass.  Other arguments and limits as turtle_gffw_fused. */
int turtle_gffw_tail(const void *t16, const void *taps16, const void *w_out16, float *x, void *ln_out16, const float *ln_w,
                     const float *ln_b, int B, int H, int W, int C, int hid, void *stream);

/* ---------------------------------------------------------------------------------------
 * Frame side of the per-clip loop: decode/normalise, quantise/encode, metrics, tiling.
 * Frames are planar fp32 [C,H,W] (what the arch's forward takes and returns); 8-bit images are interleaved
 * [H,W,C] with a row pitch in bytes (cv2 / PNG layout).
 * ------------------------------------------------------------------------------------- */

/* uint8 HWC -> fp32 CHW / 255 (swap_rb: BGR source, RGB planes).  Replaces the dataset path of
 * inference_no_ground_truth.py (cv2.imread -> cvtColor -> /255 -> permute(2,0,1), INFN:88-120) so that only
 * 1 byte per sample crosses PCIe. */
int turtle_u8_to_frame(const void *src_u8, long long src_pitch, float *dst, int H, int W, int C, int swap_rb, void *stream);

/* fp32 CHW -> clamp(0,1) -> x255 -> uint8 HWC.  round_mode 1 = round half to even (tensor2img, utils/img_util.py:73,99);
 * 0 = truncate ((x*255).astype(np.uint8), INFN:268-269).  swap_rb writes BGR (cv2.cvtColor(RGB2BGR), INFN:272). */
int turtle_frame_to_u8(const float *src, void *dst_u8, long long dst_pitch, int H, int W, int C, int swap_rb, int round_mode,
                       void *stream);

#define TURTLE_METRICS_INFERENCE 0 /* inference.py:33-61,313-327: uint8 frames, calc_PSNR, scipy-gaussian SSIM */
#define TURTLE_METRICS_BASICSR 1   /* VRM:171-200 -> metrics/psnr_ssim.py:13-68,136-180,229 on uint8 frames    */
#define TURTLE_METRICS_FLOAT 2     /* the same formulas on un-quantised [0,1] data, peak 1                        */
/* PSNR and SSIM of restored vs gt, both fp32 [C,H,W] dense on the device (C <= 4).  result (device, 4 doubles) =
 * {PSNR dB (+inf if identical), SSIM, MSE, element count}; workspace of turtle_frame_metrics_workspace(H, W) bytes.
 * Two launches, deterministic (fixed-order double reductions), no host round trip: the reference converts both
 * frames to numpy uint8 on the CPU every frame (INF:313-327). */
long long turtle_frame_metrics_workspace(int H, int W);
int turtle_frame_metrics(const float *restored, const float *gt, int C, int H, int W, int flavour, double *result,
                         void *workspace, void *stream);

/* Tiled inference (inference.py:172-246).  y0[ny] / x0[nx] are HOST arrays: the tile origins per axis in the
 * reflect-padded frame (h_idx_list / w_idx_list, INF:196-197), ny, nx <= 64; tiles are numbered row-major.
 * gather: out [ny*nx, 2, C, tile, tile] = (previous, current) frame pairs cut from prev / cur [C,H,W]; coordinates
 *         beyond H, W are reflect-padded on the fly (F.pad(..., 'reflect'), INF:185-187), so the padded frame never
 *         exists in memory.
 * blend:  out [C,Ho,Wo] = clamp(mean of the tiles [ny*nx, C, tile, tile] covering each pixel, 0, 1): E.div_(W) and
 *         torch.clamp of INF:239-245 as one gather pass (no accumulators, no atomics); clamp01 = 0 skips the clamp. */
int turtle_tile_gather(const float *prev, const float *cur, float *out, int C, int H, int W, int tile, const int *y0,
                       int ny, const int *x0, int nx, void *stream);
int turtle_tile_blend(const float *tiles, float *out, int C, int Ho, int Wo, int tile, const int *y0, int ny,
                      const int *x0, int nx, int clamp01, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TURTLE_B200_H */
