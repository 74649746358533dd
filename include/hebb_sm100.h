/*
 * hebb_sm100.h — C ABI of libhebb_sm100.so: the B200 (sm_100a) implementation of the
 * reference's Hebbian-convolution hot path (forward + soft-WTA plasticity update).
 *
 * The reference exposes NO FFI: its seam is the Python class (hebb/hebb.py, hebb/hebb3d.py).
 * Each entry point below names the reference lines it replaces.  The Python drop-in
 * (hebbian-bootstraping-semi-supervised-medical-imaging_b200/hebb/) binds these with ctypes;
 * INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is DEVICE memory unless it says "host";
 *  - tensors are fp32, contiguous NCHW / NCDHW unless a stride argument says otherwise;
 *  - the library never allocates, frees or keeps device memory: scratch comes in through
 *    (ws, ws_bytes) sized by hebb_workspace_bytes(); delta_w is read-modify-write (+=);
 *  - every call only ENQUEUES work on `stream` (a cudaStream_t passed as void*), never syncs;
 *  - return value: 0 on success, a negative hebb_status otherwise; no C++ exception crosses;
 *  - there is no CPU path: without an sm_100 device every compute call returns HEBB_EARCH.
 */
#ifndef HEBB_SM100_H_
#define HEBB_SM100_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum hebb_status {
  HEBB_OK = 0,
  HEBB_EARCH = -1,   /* no CUDA device, or device is not sm_100 */
  HEBB_ESHAPE = -2,  /* unsupported / inconsistent shape */
  HEBB_EALIGN = -3,  /* pointer not aligned as required (16 B) */
  HEBB_EWS = -4,     /* workspace missing or too small */
  HEBB_ECUDA = -5,   /* a CUDA runtime call failed; see hebb_last_cuda_error() */
  HEBB_EARG = -6,    /* null pointer / bad enum */
  HEBB_EKERNEL = -7  /* a kernel reported an internal error (watchdog) */
} hebb_status;

/* Operand precision of the two contractions.  State it per run (BASELINE.json north_star).
 *  FP32   : CUDA-core fp32 FMA implicit GEMM (exact-order fp32; parity anchor, any shape)
 *  BF16X3 : tcgen05 kind::f16, every operand split hi+lo bf16, 3 MMAs per product
 *           (hi*hi + hi*lo + lo*hi), fp32 accumulate in TMEM  -> meets the 1e-4 bar
 *  BF16   : tcgen05 kind::f16, single-pass bf16 operands on the dW contraction, split
 *           (3-pass) forward so winner indices stay exact            -> meets the 1e-2 bar */
typedef enum hebb_prec { HEBB_PREC_FP32 = 0, HEBB_PREC_BF16X3 = 1, HEBB_PREC_BF16 = 2 } hebb_prec;

/* Geometry of one Hebbian (transposed) convolution.  nd==2 uses the last two entries of
 * every 3-array (index 0 is the depth axis and must be size 1 / kernel 1 / stride 1 / pad 0).
 * pad_lo/pad_hi are the zero halo the reference's pad() really applies per axis
 * (hebb/hebb.py:83-85, hebb/hebb3d.py:82-84 — note its tuple form pads W with padding[0]). */
typedef struct HebbDesc {
  int32_t nd;          /* 2 or 3 */
  int32_t B, Cin, Cout;
  int32_t in[3];       /* unpadded input extent  (D, H, W) */
  int32_t k[3];        /* kernel extent          (kd, kh, kw) */
  int32_t stride[3];
  int32_t pad_lo[3];
  int32_t pad_hi[3];
  int32_t transposed;  /* 0: y = conv(x, W[Cout,Cin,k]);  1: y = convT(x, W[Cin,Cout,k]) */
} HebbDesc;

/* flags for hebb_conv_swta_step / hebb_convT_swta_step */
#define HEBB_F_UPDATE 1u   /* also accumulate the plasticity update into delta_w */
#define HEBB_F_WNRM   2u   /* divide each filter by its L2 norm (w_nrm=True) */
#define HEBB_F_RULE_HPCA 4u /* plasticity rule = HPCA / Sanger (hebb/hebb.py:122-135): delta_w += y X - tril(y y^T) W,
                              * patchwise; runs on the fp32 CUDA-core kernels whatever `prec` says */
/* Profiling aids (tensor-core path only): re-run ONE stage of the step on the scratch a
 * preceding full call with the same arguments left in `ws`; outputs are rewritten. */
#define HEBB_F_ONLY_PACK 0x100u   /* filter norms + bf16 packing of x and W */
#define HEBB_F_ONLY_FWD  0x200u   /* forward shift-GEMM + soft-WTA epilogue */
#define HEBB_F_ONLY_DW   0x400u   /* dW shift-GEMM + decay/accumulate (needs HEBB_F_UPDATE) */

/* Device / build query.  Pointers may be NULL.  Returns HEBB_EARCH when no sm_100 device. */
int hebb_query(int* sm_major, int* sm_minor, int* num_sms);
const char* hebb_status_str(int status);
/* cudaError_t of the last failing runtime call seen by this thread (0 if none). */
int hebb_last_cuda_error(void);
/* Code left by a kernel whose bounded mbarrier wait timed out (0 = none).  Such a kernel traps, the CUDA context is
 * lost and every later call returns HEBB_ECUDA; the word is kept in pinned host memory so it stays readable.  The Python
 * binding reports it as HEBB_EKERNEL.  Codes 1-6: forward kernel, 11-13: update kernel, 20-28: fused kernel. */
int hebb_watchdog_code(void);
const char* hebb_version(void);

/* Output extent (D,H,W) of the layer described by d. */
int hebb_out_shape(const HebbDesc* d, int32_t out[3]);

/* Scratch bytes needed by the *_step calls for this geometry and precision. */
int hebb_workspace_bytes(const HebbDesc* d, int prec, size_t* bytes);

/* a1  normalize()  hebb/hebb.py:10-13, hebb/hebb3d.py:9-12.
 * Treats W as rows x (mid x inner): element (r,m,i) lives at r*row_stride + m*mid_stride + i.
 * Wn[...] = W[...] / ||W_r||_2 (norm 0 -> divide by 1), same addressing.  inv_norm (nullable)
 * receives 1/||W_r|| (1 for zero rows).  Conv weight [Cout,Cin,k..]: rows=Cout,
 * row_stride=Cin*taps, mid=1, inner=Cin*taps.  Transposed view [Cin,Cout,k..] of a
 * [Cout,Cin,k..] buffer (hebb.py:222-224): rows=Cin, row_stride=taps, mid=Cout,
 * mid_stride=Cin*taps, inner=taps. */
int hebb_wnorm(const float* W, float* Wn, float* inv_norm, int64_t rows, int64_t row_stride,
               int64_t mid, int64_t mid_stride, int64_t inner, void* stream);

/* a2+a3+a5 (+a4+a6+a7 with HEBB_F_UPDATE): one HebbianConv{2,3}d.forward in SWTA/patchwise
 * mode — hebb/hebb.py:87-115, hebb/hebb3d.py:86-125.
 *   y[B,Cout,out..]      = conv(zero_pad(x), W/||W||, bias, stride)          (overwritten)
 *   winner[B,out..]      = argmax_c y (lowest index wins ties); nullable      (overwritten)
 *   delta_w[Cout,Cin,k..] += sum_p r[c,p] X[p,:] - (sum_p r[c,p]) W[c,:],  r = softmax_c(kinv*y)
 * W is the UN-normalised weight; bias is nullable. */
int hebb_conv_swta_step(const HebbDesc* d, const float* x, const float* W, const float* bias,
                        float kinv, float* y, int32_t* winner, float* delta_w,
                        void* ws, size_t ws_bytes, unsigned flags, int prec, void* stream);

/* hebb_conv_swta_step that also hands back the BatchNorm statistics of its output: y_stats[c][0] = sum over
 * batch and pixels of y[:,c], y_stats[c][1] = sum of squares (doubles, overwritten), accumulated in the forward
 * kernel's epilogue so that the BatchNorm that follows the layer in the networks (models/networks_2d/unet.py:53-61,
 * networks_3d/unet3d.py:97-125) needs no statistics pass of its own (hebb_bn_act_from_stats).  *y_stats_written
 * is 1 if the statistics were produced (tensor-core path, at most 512 output channels), else 0 — the step
 * itself is carried out either way and the caller falls back to hebb_bn_act_train.  SURVEY §8f row 2. */
int hebb_conv_swta_step_stats(const HebbDesc* d, const float* x, const float* W, const float* bias, float kinv,
                              float* y, int32_t* winner, float* delta_w, void* ws, size_t ws_bytes,
                              unsigned flags, int prec, double* y_stats, int* y_stats_written, void* stream);

/* Weight gradient of a stride-1 convolution on the same tcgen05 contraction kernel as the plasticity update
 * ("the wgrad kernel *is* a6 with dL/dy in place of r"):
 *     grad_w[co][ci][tap] += sum_p grad_y[b][co][p] * xpad[b][ci][p + tap]
 * Replaces the weight-gradient half of torch's convolution backward that the reference reaches through
 * loss.backward() for layers with alpha < 1 (hebb/hebb.py:185-191, train_sup_2d.py:150-168).
 * x: [B][Cin][in...], grad_y: [B][gy_channels][out...], grad_w: [Cout][Cin][taps], all fp32.
 * gy_channels: channels actually present in grad_y (0 = d->Cout); rows gy_channels..Cout-1 of grad_w receive 0 —
 *   lets a layer with few output channels (a 2-class head) use a descriptor padded to a multiple of 16.
 * channels_last: 0 = x and grad_y are contiguous NCHW / NCDHW, 1 = contiguous NHWC / NDHWC (torch channels_last).
 * prec: HEBB_PREC_BF16X3 (fp32-equivalent) or HEBB_PREC_BF16.  HEBB_ESHAPE for layers the tcgen05 planner
 * does not take (the caller then uses its own fallback; nothing is computed). */
int hebb_conv_wgrad(const HebbDesc* d, const float* x, const float* grad_y, float* grad_w, int gy_channels,
                    int channels_last, void* ws, size_t ws_bytes, int prec, void* stream);

/* a9: one HebbianConvTranspose{2,3}d.forward in swta_t/patchwise mode —
 * hebb/hebb.py:226-264, hebb/hebb3d.py:250-289.  W and delta_w are the CONTIGUOUS
 * [Cout,Cin,k..] buffers underneath the reference's transposed (Cin,Cout,k..) view; the
 * normalisation is per input channel (the view's leading dim).  Padding must be 0. */
int hebb_convT_swta_step(const HebbDesc* d, const float* x, const float* W, const float* bias,
                         float kinv, float* y, int32_t* winner, float* delta_w,
                         void* ws, size_t ws_bytes, unsigned flags, int prec, void* stream);

/* a8  local_update() for n layers in one launch — hebb/hebb.py:174-192.
 * For i<n:  grad[i][j] = has_grad[i] ? (1-alpha[i])*grad[i][j] - alpha[i]*dw[i][j]
 *                                    : -alpha[i]*dw[i][j];   dw[i][j] = 0.
 * The four arrays are HOST arrays of length n; grad[i]/dw[i] are device pointers with the
 * same (dense) memory layout, numel[i] elements each. */
int hebb_local_update_multi(int n, float* const* grad, float* const* dw, const int64_t* numel,
                            const float* alpha, const int32_t* has_grad, void* stream);

/* ---- the bandwidth-bound ops between two Hebbian convs (opt-in, hebb/fused.py) ---- */

/* nn.BatchNorm{2,3}d in TRAINING mode fused with (Leaky)ReLU — models/networks_2d/unet.py:53-61, unet3d.py:97-125:
 *   out = act((y - mean_c) / sqrt(var_c + eps) * gamma_c + beta_c),  act(v) = v >= 0 ? v : slope*v
 * (slope 0 = ReLU, 0.01 = LeakyReLU default, 1 = no activation); batch statistics over (B, S) per channel,
 * biased variance for the normalisation, running_mean/var (nullable) updated with `momentum` and the unbiased
 * variance exactly like torch.  y,out: [B][C][S] fp32 (out may alias y).  ws: >= C*24 bytes, 16-byte aligned. */
int hebb_bn_act_train(const float* y, float* out, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, int64_t B, int64_t C, int64_t S, float eps, float momentum, float slope,
                      void* ws, size_t ws_bytes, void* stream);

/* BatchNorm(train) + activation from statistics that hebb_conv_swta_step_stats already produced: only the
 * scale/shift + running-statistics update and the normalise+activate pass run (one read, one write of y).
 * ws: >= 2*C floats. */
int hebb_bn_act_from_stats(const float* y, float* out, const double* y_stats, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, int64_t B, int64_t C, int64_t S, float eps,
                           float momentum, float slope, void* ws, size_t ws_bytes, void* stream);

/* The two calls above with the nn.Dropout(p) that follows BatchNorm + activation inside the networks' blocks
 * (models/networks_2d/unet.py:53-61: Conv -> BatchNorm -> LeakyReLU -> Dropout -> Conv ...) folded into the same pass:
 * out = dropout_p(act(bn(y))), kept values scaled by 1/(1-p); no mask is returned (the pass is for tensors nothing
 * back-propagates through).  state: the device-resident Philox state {seed, launches so far} of
 * hebb_bias_relu_dropout_state (incremented on `stream`).  0 <= p < 1; p = 0 is the plain call. */
int hebb_bn_act_train_dropout(const float* y, float* out, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, int64_t B, int64_t C, int64_t S, float eps, float momentum, float slope,
                              float p, uint64_t* state, void* ws, size_t ws_bytes, void* stream);
int hebb_bn_act_from_stats_dropout(const float* y, float* out, const double* y_stats, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, int64_t B, int64_t C, int64_t S, float eps,
                                   float momentum, float slope, float p, uint64_t* state, void* ws, size_t ws_bytes,
                                   void* stream);

/* nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) — models/networks_2d/unet.py:171-172.
 * in: [N][H][W], out: [N][2H][2W] fp32, N = batch*channels. */
int hebb_upsample2x_bilinear(const float* in, float* out, int64_t N, int64_t H, int64_t W, void* stream);

/* 2x max pooling (kernel = stride = 2, no padding, floor mode, NaNs propagate) of N planes [D][H][W]:
 * nn.MaxPool2d(2) (pool_depth = 0, D = 1) and nn.MaxPool3d(kernel_size=2, stride=2) (pool_depth = 1) between
 * the Hebbian blocks (reference models/networks_2d/unet.py:70-80, networks_3d/unet3d.py:31-43).  SURVEY §8f row 2. */
int hebb_maxpool2x(const float* in, float* out, int64_t N, int64_t D, int64_t H, int64_t W, int pool_depth,
                   void* stream);

/* out = dropout_p(relu(z + bias[c])) in one pass, plus the 1-byte mask (kept AND positive) for the backward:
 * the Conv -> ReLU -> Dropout runs of the back-prop head (reference models/networks_2d/unet.py:449-457) with the
 * bias add taken out of the convolution.  z, out: n floats; element i belongs to channel (i / inner) % C
 * (inner = 1 for channels_last storage, = spatial size for NCHW).  Kept values are scaled by 1/(1-p); the random
 * stream is Philox4x32-10 keyed by (seed, thread) -- statistically the nn.Dropout mask, not torch's bit stream.
 * hebb_mask_scale is its backward: gz = gout * mask * scale.  Opt-in (hebb.fused.fuse_norm_act). */
int hebb_bias_relu_dropout(const float* z, const float* bias, float* out, uint8_t* mask, int64_t n, int64_t C,
                           int64_t inner, float p, uint64_t seed, void* stream);
/* Same, with the stream state on the device: state[0] = seed, state[1] = launches so far (incremented by the call,
 * on `stream`).  No host value enters the launch, so the call can be recorded into a CUDA graph and every replay
 * draws a new mask. */
int hebb_bias_relu_dropout_state(const float* z, const float* bias, float* out, uint8_t* mask, int64_t n, int64_t C,
                                 int64_t inner, float p, uint64_t* state, void* stream);
int hebb_mask_scale(const float* gout, const uint8_t* mask, float* gz, int64_t n, float scale, void* stream);
/* hebb_mask_scale for channels_last storage (element i belongs to channel i % C; C a power of two, 4..1024) that also
 * returns gb[c] = sum over the pixels of gz[., c] -- the bias gradient of the convolution in front of the fused
 * activation (reference: autograd of `Conv -> ReLU -> Dropout`, models/networks_2d/unet.py:449-457), saving one more
 * pass over gz.  partial: caller-owned scratch of partial_rows x C floats (the launch uses at most partial_rows blocks);
 * sums are folded in a fixed order (deterministic). */
int hebb_mask_scale_gb(const float* gout, const uint8_t* mask, float* gz, float* gb, int64_t n, int64_t C, float scale,
                       float* partial, int64_t partial_rows, void* stream);

/* ---- exported for tests and profiling ---- */

/* Kernels launched by this library in this process so far (bench.py's gpu_launches). */
unsigned long long hebb_debug_launch_count(void);

/* 1 if (d, prec) runs on the tcgen05 kernels, 0 if on the CUDA-core kernels. */
int hebb_uses_tensor_cores(const HebbDesc* d, int prec);

/* Which kernels a step of (d, prec, flags) runs on: 0 = fp32 CUDA-core kernels, 1 = tcgen05 pack / forward / update
 * kernels, 2 = the fused small-channel kernel (2-D, Cin and Cout in {16, 32}, stride 1, kernel <= 3x3: forward, soft-WTA
 * and update in ONE launch fed by TMA tensor-map loads of the fp32 NCHW input -- SURVEY 8b `hebb_fwd_dw_fused`; reached
 * through hebb_conv_swta_step / hebb_conv_swta_step_stats, hebb/hebb.py:87-115); -1 for an invalid descriptor. */
int hebb_layer_path(const HebbDesc* d, int prec, unsigned flags);

/* Which kernels hebb_conv_wgrad(d, ...) runs on: 0 = not taken (HEBB_ESHAPE), 1 = the tcgen05 pack + update kernels,
 * 2 = the fused kernel in weight-gradient mode (2-D, stride 1, kernel <= 3x3, Cin and Cout multiples of 16 that split
 * into at most two (16|32) x (16|32) channel passes: x through TMA tensor maps -- NCHW or channels_last --, dL/dy read
 * once per pass in place of the responses); -1 for an invalid descriptor. */
int hebb_wgrad_path(const HebbDesc* d, int prec);

/* Tile plan of the fused kernel (0 if the layer does not take it): {TH, TW, tile row pitch, tiles, 128-position blocks
 * per tile, x rows per tile, shared memory, TMEM columns, grid}; returns the number of fields. */
int hebb_debug_fused_plan(const HebbDesc* d, int* out, int n);

/* Profiling aid: with HEBB_FUSED_PROF=1 in the environment the fused kernel accumulates, per CTA and warp, the cycles
 * spent in each of its bounded waits (10 counters, codes 20..29) and the warp's total cycles; this copies the table of
 * the last launch ([grid][15 warps][11]) to the host buffer `out` (n entries) and returns the count.  Synchronises. */
int hebb_debug_fused_prof(long long* out, int n);

/* Tile plan the tensor-core path would use (0 if it would not run there): fills out[0..n) with
 * {MB, fwd SEGLEN, XST, WST, NACC, fwd TMEM cols, fwd tiles, fwd smem, dW by_kh, CM, CN, BLK, ST, dW SEGLEN,
 *  tap groups, cin tiles, cout tiles, position splits, position blocks, dW TMEM cols, dW smem, dW HL, ws MiB,
 *  stackM, stackN, CT, channel tiles, nrep, WG, dW collector re-use, dW tap halo, swizzled-response variant available,
 *  and its BLK, ST, smem, TMEM cols, position splits, stackM};
 * returns the number of fields. */
int hebb_debug_plan(const HebbDesc* d, int prec, int* out, int n);

/* One CTA: copy two raw operand images to shared memory, issue `ksteps` tcgen05.mma
 * (kind::f16, bf16 in, fp32 out) with descriptors {hi | (start + i*step) >> 4}, and dump
 * TMEM lanes 0..127 x n columns to d_out[128][n].  Pins the SWIZZLE_NONE descriptor
 * semantics the shift-GEMM kernels rely on (tests/test_umma_probe.py). */
int hebb_debug_umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                          uint64_t a_desc_hi, uint32_t a_start, uint32_t a_step,
                          uint64_t b_desc_hi, uint32_t b_start, uint32_t b_step,
                          uint32_t idesc, int ksteps, int m, int n, float* d_out, void* stream);

/* Issue `iters` rounds of `per_round` back-to-back tcgen05.mma per CTA from resident shared memory
 * (two zeroed regions of region_bytes each; descriptor start advances by a_step/b_step per MMA)
 * and report the elapsed SM cycles per CTA in cycles[ctas].  Used by scripts/umma_rate.py. */
int hebb_debug_umma_rate(uint64_t a_desc_hi, uint32_t a_step, uint64_t b_desc_hi, uint32_t b_step,
                         uint32_t region_bytes, uint32_t idesc, int per_round, int iters, int n, int ctas,
                         long long* cycles, void* stream);

/* Same, with groups of `group` consecutive MMAs sharing one A tile (A advances per group, B per MMA) under the
 * A-operand collector hints (hint = 0: none, 1: fill/use/lastuse, 2: fill/use/use); MMA j of a group accumulates
 * into TMEM column j*d_step.  Used by scripts/umma_rate3.py. */
int hebb_debug_umma_rate_shared_a(uint64_t a_desc_hi, uint32_t a_step, uint64_t b_desc_hi, uint32_t b_step,
                                  uint32_t region_bytes, uint32_t idesc, int per_round, int group, int hint,
                                  int d_step, int iters, int n, int ctas, long long* cycles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HEBB_SM100_H_ */
