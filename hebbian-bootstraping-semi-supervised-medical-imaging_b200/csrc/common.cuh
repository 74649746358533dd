// Shared declarations of libhebb_sm100: status codes, geometry, launch helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <atomic>
#include "hebb_sm100.h"

namespace hebb {

// Thread-local record of the last failing CUDA runtime call (hebb_last_cuda_error()).
extern thread_local int g_last_cuda_error;
// Number of kernels this library has launched in this process (hebb_debug_launch_count()).
extern std::atomic<unsigned long long> g_launches;
#define HEBB_LAUNCHED() (::hebb::g_launches.fetch_add(1, std::memory_order_relaxed))

#define HEBB_CUDA_TRY(expr)                                   \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) {                                  \
      ::hebb::g_last_cuda_error = (int)_e;                    \
      return HEBB_ECUDA;                                      \
    }                                                         \
  } while (0)

#define HEBB_TRY(expr)               \
  do {                               \
    int _s = (expr);                 \
    if (_s != HEBB_OK) return _s;    \
  } while (0)

// internal flag of tc_conv_step (never accepted through the C ABI): weight-gradient mode, see hebb_conv_wgrad
#define HEBB_F_WGRAD_INTERNAL 0x10000u

// Geometry of one layer, resolved from HebbDesc (device-friendly POD).
struct Geo {
  int nd, B, Cin, Cout;
  int iD, iH, iW;     // unpadded input extent
  int kD, kH, kW;
  int sD, sH, sW;
  int pD, pH, pW;     // pad_lo
  int qD, qH, qW;     // pad_hi
  int oD, oH, oW;     // output extent
  int taps;           // kD*kH*kW
  int K;              // Cin*taps (conv) — length of one filter
  long long inS, outS;  // spatial sizes
  int transposed;
};

int resolve_geo(const HebbDesc* d, Geo* g);
int device_ok();          // HEBB_OK iff current device is sm_100
int num_sms();
int* watchdog_word();   // device pointer of the pinned watchdog word (nullptr if it could not be allocated)

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

// ---- elementwise stage (elementwise.cu) ----
int launch_wnorm(const float* W, float* Wn, float* inv, long long rows, long long row_stride,
                 long long mid, long long mid_stride, long long inner, cudaStream_t st);
int launch_local_update_multi(int n, float* const* grad, float* const* dw, const int64_t* numel,
                              const float* alpha, const int32_t* has_grad, cudaStream_t st);
// delta_w[c,j] += H[c*(K+1)+j] - H[c*(K+1)+K] * W[c,j]
int launch_finalize_conv(const float* H, const float* W, float* delta_w, int Cout, int K, cudaStream_t st);
// transposed: see simt_path.cu
int launch_finalize_convT(const float* H, const float* W, float* delta_w, int Cin, int Cout, int taps,
                          cudaStream_t st);

// ---- fp32 CUDA-core path (simt_path.cu) ----
size_t simt_workspace_bytes(const Geo& g);
int simt_conv_step(const Geo& g, const float* x, const float* W, const float* bias, float kinv, float* y,
                   int32_t* winner, float* delta_w, void* ws, size_t ws_bytes, unsigned flags,
                   cudaStream_t st);
int simt_convT_step(const Geo& g, const float* x, const float* W, const float* bias, float kinv, float* y,
                    int32_t* winner, float* delta_w, void* ws, size_t ws_bytes, unsigned flags,
                    cudaStream_t st);

// delta_w[c][j] -= sum_{c' <= c} G[c][c'] W[c'][j]   (HPCA decay on the CUDA cores; G, W fp32)
int launch_hpca_decay(const float* G, const float* W, float* delta_w, int Cout, int K, cudaStream_t st);

// ---- exact winners for near-tie pixels (fixup.cu) ----
// Re-evaluates the pixels listed in list[0 .. min(*count, cap)) (indices into `winner`) from the fp32 inputs with
// fp64 accumulation and rewrites winner[]: argmax_c of  (sum_k x_k W[c][k]) * inv[c] + bias[c]  (plain conv,
// inv per output channel) or of the transposed convolution with inv per INPUT channel.  W: [Cout][Cin][taps].
int launch_winner_fixup(const Geo& g, const float* x, const float* W, const float* inv, const float* bias,
                        int32_t* winner, const int* list, const int* count, int cap, cudaStream_t st);

// ---- tcgen05 path (tc_path.cu) ----
bool tc_supported(const Geo& g, int prec);
size_t tc_workspace_bytes(const Geo& g, int prec);
int tc_describe_plan(const Geo& g, int prec, int* out, int n);
int tc_launch_pack_w(const float* W, void* wp, int Cin, int Cout, int taps, int NSLAB, int CT, cudaStream_t st);
int tc_launch_finalize(const float* hpart, const float* rsum, const float* W, float* dw, int n_part, int taps, int Cin,
                       int CinP, int Cout, cudaStream_t st);
int tc_conv_step(const Geo& g, const float* x, const float* W, const float* bias, float kinv, float* y,
                 int32_t* winner, float* delta_w, void* ws, size_t ws_bytes, unsigned flags, int prec,
                 cudaStream_t st, int aux = 0, double* ystats = nullptr, int* ystats_written = nullptr);

// ---- fused forward + update kernel of the small-channel 2-D layers (fused_path.cu) ----
// one launch: 1/|W_c| per filter, packed bf16 hi/lo weights ([slab][tap][k-chunk][hi|lo][Cout]), zero_bytes of zeroed
// accumulators from zero_from, zeroed BatchNorm sums (nullable)
int launch_layer_prep(const float* W, void* wp, float* inv, void* zero_from, size_t zero_bytes, double* ystats, int Cin, int Cout,
                      int taps, int wnrm, cudaStream_t st);
bool fused_supported(const Geo& g, int prec, unsigned flags);
size_t fused_workspace_bytes(const Geo& g);
int fused_describe_plan(const Geo& g, int* out, int n);
int fused_conv_step(const Geo& g, const float* x, const float* W, const float* bias, float kinv, float* y, int32_t* winner,
                    float* delta_w, void* ws, size_t ws_bytes, unsigned flags, cudaStream_t st, double* ystats = nullptr,
                    int* ystats_written = nullptr);
// hebb_conv_wgrad on the fused kernel (dL/dy in place of the responses): channel counts that split into at most two
// passes of (16|32) x (16|32) channels; x and gy dense NCHW or dense channels_last; gw[Cout][Cin][taps] +=
bool fused_wgrad_supported(const Geo& g);
size_t fused_wgrad_workspace_bytes(const Geo& g);
int fused_conv_wgrad(const Geo& g, const float* x, const float* gy, float* gw, int gy_channels, int channels_last, void* ws,
                     size_t ws_bytes, cudaStream_t st);

}  // namespace hebb
