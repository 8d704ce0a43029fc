/* tcct_b200.h -- C ABI of libtcct_b200.so, the sm_100a kernel library behind the `stc_tt` training / inference
 * hot path of tyb311/TCCT.
 *
 * The reference has no native layer: its "FFI" for this path is the PyTorch operator surface its Python code calls
 * (nn.Conv2d, nn.BatchNorm2d, F.gelu, nn.AvgPool2d ... dispatched to ATen/cuDNN).  Every entry point below replaces
 * one such call site (cited as task1/<file>:<lines> of the reference) with a hand-written kernel.  The host side
 * (tcct_b200/ops.py, ctypes) binds exactly these symbols; INTEGRATION.md shows the binding a maintainer of the
 * reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current CUDA device unless named *_host
 *   - `stream` is a cudaStream_t passed as void*; kernels are launched on it and the call returns immediately
 *     (no synchronisation, no allocation: safe under CUDA-graph capture)
 *   - activations are NHWC fp32 `[B,H,W,C]` (tokens `[B,N,C]` share the layout); logits / labels at the module
 *     boundary are NCHW fp32 `[B,C,H,W]` / uint8 `[B,H,W]` class-index maps
 *   - functions returning int return TCCT_OK (0) or an error code; tcct_last_error() gives the message
 *     (thread-local).  Invalid shapes fail loudly; there is no CPU or library fallback.
 *   - `stats` arguments are zero-initialised double[2*C] buffers that receive per-channel [sum | sum of squares]
 *     (train-mode BatchNorm statistics fused into the producing kernel's epilogue)
 *   - weight gradients are ACCUMULATED (+=) into the buffers given (the flat gradient buffer of the model)
 */
#ifndef TCCT_B200_H
#define TCCT_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define TCCT_OK 0
#define TCCT_ERR_ARG 1
#define TCCT_ERR_CUDA 2

/* activation codes (pre/post/stats_act arguments): LeakyReLU(0.01) tcct.py:811, Hardswish tcct.py:69, exact-erf GELU
 * tcct.py:35,826 */
#define TCCT_ACT_NONE 0
#define TCCT_ACT_LRELU 1
#define TCCT_ACT_HSWISH 2
#define TCCT_ACT_GELU 3

/* ------------------------------------------------------------------------------------------------ runtime */
const char* tcct_last_error(void);          /* message of the last failed call on this thread */
long long tcct_launch_count(void);          /* kernels launched (or recorded into a graph) by this library so far */
/* launches so far per tensor-core route: 0 conv2d_tma, 1 wgrad_tma, 2 gemm_tma, 3 wgrad_gemm_tma (tcgen05 + TMA kernels);
 * -1 for an unknown id.  Lets a caller assert that a shape was served by the tcgen05 path and not by the mma.sync one. */
long long tcct_route_count(int id);
int tcct_abi_version(void);
int tcct_device_arch(void);                 /* compute capability major*10+minor of the current device, -1 if none */

/* ------------------------------------------------------------------------------------------------ weights
 * One launch re-packs every dense weight of the model into tensor-core operand order (hi plane | lo plane of the
 * 3xTF32 split).  `table_dev` is a device array of n_entries records of tcct_pack_entry_size() bytes
 * (layout: tcct_b200/nets/flat.py PackPlan.DTYPE).  Replaces cuDNN's internal filter transforms. */
int tcct_pack_weights(const void* table_dev, int n_entries, int total_elems, void* stream);
int tcct_pack_entry_size(void);

/* ------------------------------------------------------------------------------------------------ dense contractions
 * nn.Conv2d (stride 1, 'same' zero padding) of CrossCNNBlock tcct.py:803-823, MPUpBlock 887-900, head 975:
 * y = conv(x, W) + bias [; y = res + res_scale[b] * y]; optional per-channel statistics of stats_act(y).
 * Cin in {32,64}, Cout % 32 == 0, odd KH x KW <= 25 taps.  TF32 tensor cores (mma.sync), fp32 accumulate;
 * lo_off != 0 selects the error-compensated 3xTF32 mode (element offset from wpk to the residual plane).
 * The same entry computes the data gradient when given the transposed/flipped pack. */
int tcct_conv2d_nhwc(const float* x, const float* wpk, long long lo_off, const float* bias, float* y, int B, int H, int W,
                     int Cin, int Cout, int KH, int KW, const float* res, const float* res_scale, double* stats,
                     int stats_act, void* stream);
/* One 32- or 64-channel slice of the reduction of a wider conv -- the (32,64,96,128,256)-channel CrossResNet and decoder of
 * stc_tb / gtc_tb (tcct.py:861-864, 887-900, 975): x points at the slice's first channel inside a [B,H,W,x_ch] tensor, wpk packs that
 * slice of the weight; slices chain through res = y (in place), bias on the first, stats on the last. */
int tcct_conv2d_nhwc_slice(const float* x, int x_ch, const float* wpk, long long lo_off, const float* bias, float* y, int B, int H,
                           int W, int Cin, int Cout, int KH, int KW, const float* res, double* stats, int stats_act, void* stream);
/* The 32->32 spatial convs (3x3, 1xk, kx1; k <= 13) as a TMA-fed tcgen05 pipeline (cp.async.bulk.tensor loads with
 * 128-byte swizzle -> tcgen05.mma kind::tf32 -> TMEM -> TMA stores): same contract as above without res.  The line
 * length (W, or H for kx1) must be a multiple of 128 -- query with tcct_conv_tma_supported (1 = supported).
 * wu is the fmt-2 pack of tcct_pack_weights. */
int tcct_conv_tma_supported(int H, int W, int Cin, int Cout, int KH, int KW);
int tcct_conv2d_tma(const float* x, const float* wu, const float* bias, float* y, int B, int H, int W, int KH, int KW,
                    double* stats, int stats_act, void* stream);
/* One 32 -> 32 channel slice of a wider convolution (the 32 -> 64 stem conv of MPViT, tcct.py:682-689, and its data gradient): reads
 * channels [x_c0, x_c0+32) of x [B,H,W,x_ch], writes (accumulate = 0) or adds into (1, a TMA reduce-add store) channels
 * [y_c0, y_c0+32) of y [B,H,W,y_ch].  wu: fmt-2 pack of the (output tile, input tile) weight block; stats: the full [2*y_ch]
 * statistics buffer of y or null. */
int tcct_conv2d_tma_slice(const float* x, int x_ch, int x_c0, const float* wu, const float* bias, float* y, int y_ch, int y_c0,
                          int accumulate, int B, int H, int W, int KH, int KW, double* stats, int stats_act, void* stream);
/* 1x1 convs / nn.Linear over pixels (Conv2d_BN tcct.py:55-97, Mlp 29-53, tran_vit/tran_cnn 966-973, t32x 988-991):
 * y[M][N] = x[M][K] . W^T (+bias) [; y = res + res_scale[m / px_per_sample] * y].  K, N % 32 == 0. */
int tcct_gemm_px(const float* x, const float* wpk, long long lo_off, const float* bias, float* y, long long M, int K, int N,
                 const float* res, const float* res_scale, int px_per_sample, double* stats, int stats_act, void* stream);
/* The same GEMM on large maps (M % 128 == 0, M >= 8192, weights <= ~100 KB) as a TMA-fed tcgen05 pipeline: x tiles by
 * cp.async.bulk.tensor with 128-byte swizzle, the weight matrix resident in shared memory, accumulators in TMEM,
 * TMA stores.  wu is the fmt-3 pack of tcct_pack_weights.  Query with tcct_gemm_tma_supported (1 = supported). */
int tcct_gemm_tma_supported(long long M, int K, int N);
int tcct_gemm_tma(const float* x, const float* wu, const float* bias, float* y, long long M, int K, int N, const float* res,
                  const float* res_scale, int px_per_sample, double* stats, int stats_act, void* stream);
/* The MLP of an MHCA block (Mlp tcct.py:29-53, MHCABlock.forward 467-468; SURVEY 8(b) `ln_mlp`: LayerNorm2 is the second output of
 * tcct_ln_metapool_fwd) without separate activation passes: fc1 writes the pre-activation y (kept for the backward) and
 * y_act = GELU(y) from the same epilogue; the data gradient through fc2 multiplies by GELU'(h) in its epilogue
 * (dh = (dy W2) * GELU'(h); K = channels of dy, N = channels of h).  Same shape support as tcct_gemm_tma. */
int tcct_gemm_tma_gelu(const float* x, const float* wu, const float* bias, float* y, float* y_act, long long M, int K, int N, void* stream);
int tcct_gemm_tma_dgelu(const float* dy, const float* wu_t, const float* h, float* dh, long long M, int K, int N, void* stream);
/* Weight / bias gradient of those GEMMs on large maps (N <= 128, K <= 256), contraction over pixels with both operands
 * MN-major from 32B-atom-swizzled TMA tiles; accumulator resident in TMEM, partials -> workspace -> grid barrier ->
 * sliced reduction.  Row n of the gradient is accumulated at dw + n*ld (dw already offset to the first input column of
 * a concat slice); ws: tcct_wgrad_gemm_tma_ws_floats floats; counter: unused, pass null (ABI slot of a removed software grid barrier). */
int tcct_wgrad_gemm_tma_supported(long long M, int K, int N);
long long tcct_wgrad_gemm_tma_ws_floats(long long M, int K, int N);
int tcct_wgrad_gemm_tma(const float* x, const float* dy, float* dw, float* dbias, long long M, int K, int N, int ld, float* ws,
                        unsigned int* counter, void* stream);
/* Weight / bias gradients of both (autograd's convolution_backward weight path):
 * dw[co*sco + ci*sci + tap*stp] += sum_px dy[px][co] * x[px + tap][ci];  dbias[co] += sum_px dy[px][co].
 * KH*KW == 1 selects the linear mode (x rows of Cin channels, B*H*W pixels). x3 = 1: 3xTF32. */
int tcct_wgrad(const float* x, const float* dy, float* dw, float* dbias, int B, int H, int W, int Cin, int Cout, int KH,
               int KW, int sco, int sci, int stp, int x3, void* stream);
/* Spatial weight gradient of one 32-channel input slice of a wider conv (same layers): x points at the slice inside a
 * [B,H,W,x_ch] tensor, dw at dW[0][slice][0]; dbias (or null) on one slice only. */
int tcct_wgrad_slice(const float* x, int x_ch, const float* dy, float* dw, float* dbias, int B, int H, int W, int Cout, int KH,
                     int KW, int sco, int sci, int stp, int x3, void* stream);
/* The weight gradient of the 32->Cout spatial convs (3x3, 1x13, 13x1, 1x11, 11x1; line length % 128 == 0; Cout = 32,
 * 64, 96, 128) as a TMA-fed tcgen05 pipeline with the contraction over pixels (both operands MN-major from the swizzled
 * NHWC line buffers), one launch per 32 output channels followed by a partial-sum reduce kernel.
 * dw: PyTorch [Cout][32][KH][KW], dbias [Cout] or null (both accumulated); ws: tcct_wgrad_tma_ws_floats floats of scratch;
 * counter: unused (earlier versions ran a grid-wide barrier on it). */
int tcct_wgrad_tma_supported(int H, int W, int Cin, int Cout, int KH, int KW);
long long tcct_wgrad_tma_ws_floats(int B, int H, int W, int KH, int KW);
int tcct_wgrad_tma(const float* x, const float* dy, float* dw, float* dbias, int B, int H, int W, int KH, int KW, int Cout,
                   float* ws, unsigned int* counter, void* stream);
/* Deferred second phase: the tcgen05 weight-gradient kernels leave per-CTA partial sums in a workspace and a small launch folds them
 * into dW.  Per layer that launch is pure latency (7-12 us); a backward pass has 48 of them on the streams that finish the step.
 * The *_partial entries run the first phase only (dbias is complete, dW untouched; ws: Cout/32 regions for the conv form) and
 * report the slab geometry (conv: parts[4] = {slabs per region, S, KA, KL}; 1x1: parts[1] = {slabs}); tcct_wgrad_reduce_batch then
 * folds any number of such workspaces in ONE launch (jobs: host array, copied into the launch parameters). */
typedef struct {
  const float* ws; float* dw;
  int kind;                 /* 0: conv (p0 = S, p1 = KA, p2 = KL, dw = dW of the 32-output-channel slice); 1: 1x1 (p0 = N, p1 = K, p2 = row stride of dW) */
  int nparts;
  int p0, p1, p2, reserved;
} tcct_reduce_job;
int tcct_reduce_job_size(void);
int tcct_wgrad_tma_partial(const float* x, const float* dy, float* dbias, int B, int H, int W, int KH, int KW, int Cout, float* ws, int* parts,
                           void* stream);
int tcct_wgrad_gemm_tma_partial(const float* x, const float* dy, float* dbias, long long M, int K, int N, float* ws, int* parts, void* stream);
int tcct_wgrad_reduce_batch(const void* jobs_host, int n, void* stream);

/* ------------------------------------------------------------------------------------------------ normalisation family
 * nn.BatchNorm2d train/eval (eps 1e-5, momentum 0.1, unbiased running variance; tcct.py:63,811,817,823 ...).
 * stats_nhwc: per-channel sums of an NHWC tensor.  bn_finalize: coef = [scale | shift | mean | invstd] (4*C floats)
 * from the batch sums (stats != null; running statistics updated when update_running) or from the running ones. */
int tcct_stats_nhwc(const float* x, long long npix, int C, double* stats, void* stream);
int tcct_bn_finalize(const double* stats, double count, const float* gamma, const float* beta, float eps, float momentum,
                     float* running_mean, float* running_var, long long* num_batches, int update_running, float* coef,
                     int C, void* stream);
/* out = post( opA(a) + opB(b) ), op(v) = scale*pre(v) + shift (coef null: identity; b null: one operand).
 * Covers BN+activation, GELU(BN(a)+BN(b)) of CrossCNNBlock.forward tcct.py:825-828, x + BN(conv) of ResBlock
 * 562-571 and the plain skip adds of FTC.forward 1026-1040.  The backward returns da, db and accumulates
 * dgamma/dbeta; `sums` is a zeroed double[8*3*C + 1] workspace
 * (8 replicas of the batch sums; null: eval-mode statistics). */
int tcct_bn_act2_fwd(const float* a, const float* coefA, int preA, const float* b, const float* coefB, int preB, int post,
                     float* out, long long npix, int C, void* stream);
/* The same with tcct_bn_finalize fused into the kernel's prologue (one launch per BatchNorm+activation instead of two
 * or three): bnA / bnB are HOST pointers to tcct_bn_src records read at call time (null: operand not normalised);
 * the kernel writes each record's coef array (for the backward) and updates its running statistics. */
typedef struct tcct_bn_src {
  const double* stats;        /* device [2C] sum | sum of squares (null: eval mode, use the running statistics) */
  double count;               /* elements per channel behind stats */
  const float* gamma; const float* beta;
  float eps, momentum;
  float* running_mean; float* running_var; long long* num_batches;
  int update_running;
  float* coef;                /* device out [4C]: scale | shift | mean | invstd */
} tcct_bn_src;
int tcct_bn_act2_fwd_bn(const float* a, const tcct_bn_src* bnA, int preA, const float* b, const tcct_bn_src* bnB, int preB,
                        int post, float* out, long long npix, int C, void* stream);
int tcct_bn_act2_bwd(const float* a, const float* coefA, int preA, const float* gammaA, const float* b,
                     const float* coefB, int preB, const float* gammaB, int post, const float* dout, double* sums,
                     float* da, float* db, float* dgammaA, float* dbetaA, float* dgammaB, float* dbetaB, long long npix,
                     int C, void* stream);
/* nn.LayerNorm over C (eps 1e-6, tcct.py:427,454-455); mean_rstd: [ntok][2] saved for the backward */
int tcct_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean_rstd, long long ntok,
                       int C, float eps, void* stream);
int tcct_layernorm_bwd(const float* x, const float* gamma, const float* mean_rstd, const float* dy, float* dx,
                       float* dgamma, float* dbeta, long long ntok, int C, void* stream);
/* alpha * F.normalize(x, dim=C) over 32 channels (norm_add tcct.py:937-942); the backward scales dy by alpha */
int tcct_l2norm32_fwd(const float* x, float* y, long long npix, float alpha, void* stream);
int tcct_l2norm32_bwd(const float* x, const float* dy, float* dx, long long npix, float alpha, void* stream);
/* norm_add (tcct.py:937-942) in one pass: out = alpha * (x0/|x0| + up(n1) + up(n2)); n1 [B,h1,w1,32], n2 [B,h2,w2,32]
 * (null: absent) are already normalised lower-resolution maps, bilinear align_corners=False */
int tcct_norm_add3_fwd(const float* x0, const float* n1, const float* n2, float* out, int B, int H, int W, int h1, int w1,
                       int h2, int w2, float alpha, void* stream);

/* ------------------------------------------------------------------------------------------------ pooling / depthwise / token mixing */
/* GateFusion of the gtc_* models (tcct.py:916-932): out = x1 * a + x2 * (1 - a) with a = clamp(bicubic_up(alpha), 0, 1) evaluated in
 * registers (alpha: the small [B,C,hs,ws] field of torch.rand draws, NCHW; null = eval mode, a = 0.5); bwd: d1 = dy * a, d2 = dy * (1 - a). */
int tcct_gate_fuse_fwd(const float* x1, const float* x2, const float* alpha, float* out, int B, int H, int W, int C, int hs, int ws,
                       void* stream);
int tcct_gate_fuse_bwd(const float* dy, const float* alpha, float* d1, float* d2, int B, int H, int W, int C, int hs, int ws,
                       void* stream);

/* nn.MaxPool2d(2) tcct.py:867,883 (backward routes to the first maximum in window scan order, like ATen) */
int tcct_maxpool2_fwd(const float* x, float* y, int B, int H, int W, int C, void* stream);
int tcct_maxpool2_bwd(const float* x, const float* dy, float* dx, int B, int H, int W, int C, void* stream);
/* depthwise 3x3, pad 1, stride 1|2 (DWConv2d_BN tcct.py:99-147, ResBlock 535-543); add_input: y = dw(x)+bias+x
 * (ConvPosEnc tcct.py:197-217).  w is PyTorch's [C,1,3,3]. */
int tcct_dwconv3_fwd(const float* x, const float* w, const float* bias, float* y, int B, int H, int W, int C, int stride,
                     int add_input, double* stats, void* stream);
int tcct_dwconv3_bwd(const float* x, const float* w, const float* dy, float* dx, float* dw, float* dbias, int B, int H,
                     int W, int C, int stride, int add_input, void* stream);
/* MetaPool token mixer with residual and DropPath scale (MHCABlock.forward tcct.py:457-469, MetaPool 405-415):
 * out = t + scale[b] * (avgpool3x3 over the (token, channel) plane, count_include_pad=False, of cur  -  cur) */
/* MHCABlock token mixer in one pass (nets/tcct.py:457-469, MetaPool 405-415; SURVEY 8(b) `ln_metapool_{fwd,bwd}`):
 *   cur = LayerNorm1(t); t2 = t + scale[b] * (avgpool3x3_{(token,channel) plane, valid count}(cur) - cur); cur2 = LayerNorm2(t2).
 * t [B,N,C]; scale [B] (DropPath mask / keep) or null; out t2, cur2 [B,N,C], stats [B*N*4] (mean1, rstd1, mean2, rstd2).
 * Backward: dt2 / dcur2 = gradients of the two outputs (either may be null); dt written; dg1, db1, dg2, db2 [C] accumulated. */
int tcct_ln_metapool_fwd(const float* t, const float* g1, const float* b1, const float* g2, const float* b2, const float* scale, float* t2,
                         float* cur2, float* stats, int B, int N, int C, float eps, void* stream);
int tcct_ln_metapool_bwd(const float* t, const float* t2, const float* stats, const float* g1, const float* g2, const float* scale,
                         const float* dt2, const float* dcur2, float* dt, float* dg1, float* db1, float* dg2, float* db2, int B, int N,
                         int C, void* stream);
int tcct_metapool_fwd(const float* t, const float* cur, const float* scale, float* out, int B, int N, int C, void* stream);
int tcct_metapool_bwd(const float* dy, const float* scale, float* dcur, int B, int N, int C, void* stream);
/* y = x * scale[sample] (DropPath tcct.py:452,465,468) */
int tcct_scale_per_sample(const float* x, const float* scale, float* y, long long n, int per_sample, void* stream);

/* ------------------------------------------------------------------------------------------------ resampling
 * bilinear, ATen index rules: align=1 nn.Upsample(align_corners=True) of MPUpBlock tcct.py:890;
 * align=0 F.interpolate of norm_add / the auxiliary heads tcct.py:941,1042-1044.
 * nhwc: out = alpha * up(x) (+ add) (+ out when accumulate). */
int tcct_resize_nhwc_fwd(const float* x, const float* add, float* out, int B, int h, int w, int H, int W, int C, int align,
                         float alpha, int accumulate, void* stream);
int tcct_resize_nhwc_bwd(const float* dout, float* dx, int B, int h, int w, int H, int W, int C, int align, float alpha,
                         void* stream);
int tcct_resize_nchw_fwd(const float* x, float* out, int planes, int h, int w, int H, int W, void* stream);
int tcct_resize_nchw_bwd(const float* dout, float* dx, int planes, int h, int w, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------ stems and heads */
/* 3x3 conv 3->32 on the NCHW image, NHWC out: CrossResNet.cnn tcct.py:873 (stride 1, bias) and MPViT stem 673
 * (stride 2, no bias).  The image needs no gradient. */
int tcct_stem_conv_fwd(const float* img, const float* w, const float* bias, float* y, int B, int H, int W, int stride,
                       double* stats, void* stream);
int tcct_stem_conv_wgrad(const float* img, const float* dy, float* dw, float* dbias, int B, int H, int W, int stride,
                         void* stream);
/* aux0/1/2/4: 1x1 conv 32 -> n_class, NHWC features in, NCHW logits out (tcct.py:993-996,1041-1044) */
int tcct_head_fwd(const float* x, const float* w, const float* bias, float* out, int B, int HW, int Cc, void* stream);
int tcct_head_bwd(const float* x, const float* w, const float* dl, float* dx, float* dw, float* db, int B, int HW, int Cc,
                  void* stream);

/* ------------------------------------------------------------------------------------------------ labels, Dice, metrics */
/* int64 one-hot [B,C,HW] (kite/loop_seg.py:119) or int64 index map -> uint8 index map */
int tcct_onehot_to_index(const long long* onehot, unsigned char* lab, int B, int C, int HW, void* stream);
int tcct_index64_to_u8(const long long* idx, unsigned char* lab, long long n, void* stream);
/* MultiLoss(DiceLoss) mode 0 / MultiLoss(MSELoss) mode 1 (kite/losses/loss.py:9-37,70-110): softmax over C, whole-
 * batch sums, smooth 1.  sums: zeroed double[3*C+1]; coef: float[2*C+1] saved for the backward.
 * bwd: dlogits (=|+=) weight * gscale[0] * dloss/dlogits. */
/* Deep-supervision Dice of KiteBack.grad_calc (kite/loopback.py:62-73) over the four heads of FTC.forward (nets/tcct.py:1041-1044) in
 * one launch: z0 [B,C,H,W] full-resolution logits, z1..z3 [B,C,hs[k],ws[k]] the auxiliary logits at their NATIVE resolution (the
 * F.interpolate(bilinear, align_corners=False) of tcct.py:1042-1044 is evaluated in registers; the up-sampled maps are never
 * materialised: SURVEY 8(b) `dice_multi_{fwd,bwd}`).  hs / ws / weights are HOST arrays (3, 3, 4 entries; weights = 1, coff_ds x 3).
 * sums: zeroed double[tcct_dice_multi_sums_doubles(C)]; loss: float[5] = the four MultiLoss(DiceLoss) values and their weighted
 * sum; coef: float[4*2*C] for the backward.  Backward: d0 [B,C,H,W] is written, d1..d3 (ZEROED, low resolution) are accumulated
 * through the adjoint of the interpolation; gscale = dL/d(total) (device scalar). */
/* Validation scores of kite/losses/miou.py:28-44,69-91 (MIouLoss.score / scorem, MDiceLoss.score / scores / scorem) on arbitrary soft or
 * hard maps: out[b][c] = {sum(pr*gt), sum(pr), sum(gt)} per image and class plane.  pr float [B,C,H,W]; gt float (gt_is_i64 = 0) or
 * int64 one-hot (1); out: ZEROED double[B*C*3]. */
int tcct_score_sums(const float* pr, const void* gt, int gt_is_i64, int B, int C, int HW, double* out, void* stream);
long long tcct_dice_multi_sums_doubles(int C);
int tcct_dice_multi_fwd(const float* z0, const float* z1, const float* z2, const float* z3, const int* hs, const int* ws,
                        const unsigned char* lab, int B, int C, int H, int W, const float* weights, double* sums, float* loss,
                        float* coef, void* stream);
int tcct_dice_multi_bwd(const float* z0, const float* z1, const float* z2, const float* z3, const int* hs, const int* ws,
                        const unsigned char* lab, int B, int C, int H, int W, const float* weights, const float* coef,
                        const float* gscale, float* d0, float* d1, float* d2, float* d3, void* stream);
int tcct_dice_fwd(const float* logits, const unsigned char* lab, int B, int C, int HW, int mode, double* sums, float* loss,
                  float* coef, void* stream);
int tcct_dice_bwd(const float* logits, const unsigned char* lab, int B, int C, int HW, const float* coef,
                  const float* gscale, float weight, float* dlogits, int accumulate, void* stream);
/* KiteSeg.predict argmax (kite/loop_seg.py:21-33; first maximum wins like torch.argmax) and the per-image
 * [C][3] = (intersection, predicted, true) pixel counts behind MDiceLoss/MIouLoss (kite/losses/miou.py:28-44,69-91) */
int tcct_argmax_nchw(const float* logits, unsigned char* lab, int B, int C, int HW, void* stream);
/* soft_argmax (nets/reg.py:27-35): out [B,1,H,W] = sum_c c * softmax_C(beta * logits).  boundary_positions: the soft-argmax
 * boundary extraction of the inference contract (defined here, the reference has none):
 * pos [B,C-1,W] = sum_h h * softmax_H(beta * |p_c[h] - p_c[h-1]|), p = softmax_C(logits), classes 1..C-1. */
int tcct_soft_argmax(const float* logits, float* out, int B, int C, int HW, float beta, void* stream);
int tcct_boundary_positions(const float* logits, float* pos, int B, int C, int H, int W, float beta, void* stream);
int tcct_label_counts(const unsigned char* pred, const unsigned char* truth, int B, int C, int HW, int* counts,
                      void* stream);

/* ------------------------------------------------------------------------------------------------ optimizer
 * clip_grad_norm_(12) + AdamW (kite/loop_seg.py:128-130, kite/loopback.py:126-128) over the flat buffers:
 * sqnorm accumulates sum g^2 (double[2], zeroed by the caller); adamw_step applies grad_scale, the clip factor
 * derived from sqnorm, bias-corrected AdamW with decoupled weight decay.  state = [step, lr, last grad norm, -]. */
int tcct_sqnorm(const float* g, long long n, double* out, void* stream);
int tcct_adamw_step(float* p, const float* g, float* m, float* v, long long n, const double* sqnorm, float* state,
                    float max_norm, float beta1, float beta2, float eps, float wd, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------ boundary regression
 * RegNet.regular_reg (nets/reg.py:109-156 with the modules of 64-77): both branches (logits[:,1:] and the one-hot
 * labels) in the same launches.  eps: [2][B][C-1][H][W] uniform(0,1) Gumbel noise (pred, true); jit: [2][H]
 * uniform(0,1) row jitter (pred, true).  ws: float[tcct_breg_ws_floats] (no initialisation needed);
 * dws: zeroed double[10]; both are kept for the backward.  bws: float[tcct_breg_bwd_ws_floats], zeroed from
 * float 6*B*H*W on; dlogits zero-initialised [B,C,H,W]; gout: device scalar upstream gradient. */
long long tcct_breg_ws_floats(int B, int C, int H, int W);
long long tcct_breg_bwd_ws_floats(int B, int C, int H, int W);
int tcct_breg_forward(const float* logits, const unsigned char* lab, const float* eps, const float* jit, const float* w0,
                      const float* b0, const float* w1, const float* b1, const float* wm0, const float* bm0,
                      const float* gamma, const float* beta, const float* wm2, const float* bm2, float* rmean, float* rvar,
                      long long* nbt, int training, int B, int C, int H, int W, float* ws, double* dws, float* loss,
                      void* stream);
int tcct_breg_backward(const float* logits, const unsigned char* lab, const float* eps, const float* jit, const float* w0,
                       const float* b0, const float* w1, const float* b1, const float* wm0, const float* gamma,
                       const float* wm2, int training, int B, int C, int H, int W, float* ws, double* dws, float* bws,
                       const float* gout, float* dlogits, float* dw0, float* db0, float* dw1, float* db1, float* dwm0,
                       float* dbm0, float* dgamma, float* dbeta, float* dwm2, float* dbm2, void* stream);

/* ------------------------------------------------------------------------------------------------ feature polarisation
 * RegNet.regular_udh (nets/reg.py:86-105) = FeatConSuper.select1 / points_selection_bins (nets/fcs.py:25-50,82-96)
 * + cosinesim / foreach_loss (fcs.py:63-80) + FeatConPolar.choice (nets/fcp.py:72-75).  feat: NHWC [B,H,W,32];
 * proto: [C][32] (fcp.buf_grad).  iws: unsigned int[tcct_fpolar_ws_words(B*H*W)], first 16 words zeroed, kept for the
 * backward (sorted pixel order); fws: zeroed tcct_fpolar_fws_bytes() bytes; pro_last: float[1024]. */
long long tcct_fpolar_ws_words(long long n);
long long tcct_fpolar_fws_bytes(void);
int tcct_fpolar_forward(const float* feat, const float* logits, const unsigned char* lab, const float* proto, int B, int C,
                        int H, int W, unsigned int* iws, void* fws, float* loss, float* pro_last, void* stream);
int tcct_fpolar_backward(const unsigned char* lab, const float* proto, const float* pro_last, int B, int C, int H, int W,
                         const unsigned int* iws, const float* gout, float* dfeat, void* stream);

/* ------------------------------------------------------------------------------------------------ input / deploy formats
 * The deterministic part of the reference's data path on decoded uint8 frames.
 * prep_pair: EyeSetResource.readPair (data/octnpy.py:117-129) + the tensor conversion of EyeSetGenerator.__getitem__
 * (data/octgen.py:124-126): rows [row0, row0+rows) of img [B][Hs][Ws][3] / lab [B][Hs][Ws] (gray levels), label // divide,
 * cv2 INTER_NEAREST resize to H x W, image -> [B][3][H][W] float in [0,1], label -> [B][H][W] uint8 class indices
 * (either pair of pointers may be null).  post_labels: EyeSetResource.postprocess (octnpy.py:95-112): index map * divide,
 * INTER_NEAREST resize to Ho x Wo, pasted into rows [row0, row0+Ho) of a zero [B][Hfull][Wo] frame. */
int tcct_prep_pair(const unsigned char* img, const unsigned char* lab, int B, int Hs, int Ws, int row0, int rows, int H, int W,
                   int divide, float* out_img, unsigned char* out_lab, void* stream);
int tcct_post_labels(const unsigned char* lab, int B, int H, int W, int Ho, int Wo, int row0, int Hfull, int divide,
                     unsigned char* out, void* stream);
/* readPair + make_tran (task1/data/octgen.py:9-19: PadIfNeeded, CropNonEmptyMaskIfExists, Horizontal/VerticalFlip, RGBShift,
 * HueSaturationValue, RandomContrast, RandomBrightness on the uint8 pair) + the tensor conversion of octgen.py:124-126 in one launch.
 * The random draws arrive as one record per sample (params_dev: B records of tcct_aug_params_size() bytes; layout AugParams in
 * csrc/prep.cu, mirrored by tcct_b200/data/octgen.py): crop origin, flips, the shifts / factors of the colour tables.
 * rows [row0, row0+rows) of the decoded frame are nearest-resized to Hp x Wp (readPair), padded to >= H x W and cropped. */
int tcct_aug_params_size(void);
int tcct_prep_augment(const unsigned char* img, const unsigned char* lab, const void* params_dev, int B, int Hs, int Ws, int row0, int rows,
                      int Hp, int Wp, int H, int W, int divide, float* out_img, unsigned char* out_lab, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TCCT_B200_H */
