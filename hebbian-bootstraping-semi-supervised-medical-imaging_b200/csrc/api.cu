// C-ABI entry points of libhebb_sm100.so (see include/hebb_sm100.h).
#include "common.cuh"
#include <atomic>
#include <mutex>

namespace hebb {

thread_local int g_last_cuda_error = 0;
std::atomic<unsigned long long> g_launches{0};

// Per-device cache of the properties the launches need (a process may drive several B200s, one current at a time).
constexpr int kMaxDev = 64;
static int g_dev_state[kMaxDev];   // 0 unknown, 1 sm_100, -1 anything else
static int g_dev_sms[kMaxDev], g_dev_major[kMaxDev], g_dev_minor[kMaxDev];
static std::mutex g_mu;

// Returns the current device's slot (after probing it once), or -1 without a usable CUDA device.
static int current_slot() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) { (void)cudaGetLastError(); return -1; }
  if (g_dev_state[dev]) return dev;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_dev_state[dev]) return dev;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
  g_dev_major[dev] = p.major; g_dev_minor[dev] = p.minor; g_dev_sms[dev] = p.multiProcessorCount;
  g_dev_state[dev] = (p.major == 10) ? 1 : -1;
  return dev;
}

// Watchdog word: every tcgen05 kernel writes the code of a timed-out mbarrier wait here before it traps (umma.cuh:
// mbar_wait).  The word lives in pinned, mapped host memory so that the host can still read it after the trap has
// taken the CUDA context down; hebb_watchdog_code() returns it and the Python binding reports HEBB_EKERNEL.
static int* g_wd_host = nullptr;
static int* g_wd_dev = nullptr;
int* watchdog_word() {
  if (g_wd_dev) return g_wd_dev;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_wd_dev) return g_wd_dev;
  int* h = nullptr;
  if (cudaHostAlloc(reinterpret_cast<void**>(&h), 64, cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  h[0] = 0;
  int* d = nullptr;
  if (cudaHostGetDevicePointer(reinterpret_cast<void**>(&d), h, 0) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
  g_wd_host = h; g_wd_dev = d;
  return d;
}

int device_ok() { const int s = current_slot(); return (s >= 0 && g_dev_state[s] == 1) ? HEBB_OK : HEBB_EARCH; }
int num_sms() { const int s = current_slot(); return s >= 0 ? g_dev_sms[s] : 148; }

int resolve_geo(const HebbDesc* d, Geo* g) {
  if (!d || !g) return HEBB_EARG;
  if (d->nd != 2 && d->nd != 3) return HEBB_ESHAPE;
  if (d->B <= 0 || d->Cin <= 0 || d->Cout <= 0) return HEBB_ESHAPE;
  for (int i = 0; i < 3; ++i) {
    if (d->in[i] <= 0 || d->k[i] <= 0 || d->stride[i] <= 0 || d->pad_lo[i] < 0 || d->pad_hi[i] < 0) return HEBB_ESHAPE;
  }
  if (d->nd == 2 && (d->in[0] != 1 || d->k[0] != 1 || d->stride[0] != 1 || d->pad_lo[0] || d->pad_hi[0])) return HEBB_ESHAPE;
  g->nd = d->nd; g->B = d->B; g->Cin = d->Cin; g->Cout = d->Cout;
  g->iD = d->in[0]; g->iH = d->in[1]; g->iW = d->in[2];
  g->kD = d->k[0]; g->kH = d->k[1]; g->kW = d->k[2];
  g->sD = d->stride[0]; g->sH = d->stride[1]; g->sW = d->stride[2];
  g->pD = d->pad_lo[0]; g->pH = d->pad_lo[1]; g->pW = d->pad_lo[2];
  g->qD = d->pad_hi[0]; g->qH = d->pad_hi[1]; g->qW = d->pad_hi[2];
  g->transposed = d->transposed ? 1 : 0;
  const int xD = g->iD + g->pD + g->qD, xH = g->iH + g->pH + g->qH, xW = g->iW + g->pW + g->qW;
  if (!g->transposed) {
    if (xD < g->kD || xH < g->kH || xW < g->kW) return HEBB_ESHAPE;
    g->oD = (xD - g->kD) / g->sD + 1; g->oH = (xH - g->kH) / g->sH + 1; g->oW = (xW - g->kW) / g->sW + 1;
  } else {
    g->oD = (xD - 1) * g->sD + g->kD; g->oH = (xH - 1) * g->sH + g->kH; g->oW = (xW - 1) * g->sW + g->kW;
  }
  g->taps = g->kD * g->kH * g->kW;
  g->K = g->Cin * g->taps;
  g->inS = (long long)g->iD * g->iH * g->iW;
  g->outS = (long long)g->oD * g->oH * g->oW;
  // index arithmetic in the kernels is 32-bit within one image and 64-bit across the batch
  if (g->inS * g->Cin >= (1LL << 31) || g->outS * g->Cout >= (1LL << 31)) return HEBB_ESHAPE;
  if ((long long)xD * xH * xW >= (1LL << 30)) return HEBB_ESHAPE;
  return HEBB_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace hebb

using namespace hebb;

extern "C" {

int hebb_query(int* sm_major, int* sm_minor, int* n_sms) {
  const int ok = device_ok();
  const int sl = current_slot();
  if (sm_major) *sm_major = sl >= 0 ? g_dev_major[sl] : 0;
  if (sm_minor) *sm_minor = sl >= 0 ? g_dev_minor[sl] : 0;
  if (n_sms) *n_sms = sl >= 0 ? g_dev_sms[sl] : 0;
  return ok;
}

const char* hebb_status_str(int s) {
  switch (s) {
    case HEBB_OK: return "ok";
    case HEBB_EARCH: return "no sm_100 (B200) CUDA device: this library has no other backend";
    case HEBB_ESHAPE: return "unsupported or inconsistent layer geometry";
    case HEBB_EALIGN: return "pointer must be 16-byte aligned";
    case HEBB_EWS: return "workspace missing or smaller than hebb_workspace_bytes()";
    case HEBB_ECUDA: return "CUDA runtime error (see hebb_last_cuda_error)";
    case HEBB_EARG: return "null pointer or invalid enum argument";
    case HEBB_EKERNEL: return "kernel watchdog reported an internal error";
    default: return "unknown hebb status";
  }
}

int hebb_last_cuda_error(void) { return g_last_cuda_error; }

int hebb_watchdog_code(void) { return g_wd_host ? *reinterpret_cast<volatile int*>(g_wd_host) : 0; }

unsigned long long hebb_debug_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* hebb_version(void) { return "hebb_sm100 0.1 (sm_100a; fp32 CUDA-core + tcgen05 bf16/bf16x3)"; }

int hebb_out_shape(const HebbDesc* d, int32_t out[3]) {
  Geo g;
  HEBB_TRY(resolve_geo(d, &g));
  if (!out) return HEBB_EARG;
  out[0] = g.oD; out[1] = g.oH; out[2] = g.oW;
  return HEBB_OK;
}

static bool use_tc(const Geo& g, int prec) {
  return prec != HEBB_PREC_FP32 && tc_supported(g, prec);
}

int hebb_workspace_bytes(const HebbDesc* d, int prec, size_t* bytes) {
  Geo g;
  HEBB_TRY(resolve_geo(d, &g));
  if (!bytes) return HEBB_EARG;
  if (prec < HEBB_PREC_FP32 || prec > HEBB_PREC_BF16) return HEBB_EARG;
  // the CUDA-core scratch is also what the HPCA rule uses, whatever the precision mode
  const size_t a = use_tc(g, prec) ? tc_workspace_bytes(g, prec) : 0, b = simt_workspace_bytes(g);
  const size_t c = (prec != HEBB_PREC_FP32 && fused_supported(g, prec, 0)) ? fused_workspace_bytes(g) : 0;
  const size_t e = (prec != HEBB_PREC_FP32 && !g.transposed) ? fused_wgrad_workspace_bytes(g) : 0;      // hebb_conv_wgrad
  size_t m = a > b ? a : b;
  if (c > m) m = c;
  if (e > m) m = e;
  *bytes = m;
  return HEBB_OK;
}

int hebb_uses_tensor_cores(const HebbDesc* d, int prec) {
  Geo g;
  if (resolve_geo(d, &g) != HEBB_OK) return 0;
  if (prec != HEBB_PREC_FP32 && !g.transposed && fused_supported(g, prec, 0)) return 1;
  return use_tc(g, prec) ? 1 : 0;
}

int hebb_layer_path(const HebbDesc* d, int prec, unsigned flags) {
  Geo g;
  if (resolve_geo(d, &g) != HEBB_OK) return -1;
  if (prec != HEBB_PREC_FP32 && !g.transposed && fused_supported(g, prec, flags & 0xFFFFu)) return 2;
  return use_tc(g, prec) ? 1 : 0;
}

int hebb_wgrad_path(const HebbDesc* d, int prec) {
  Geo g;
  if (resolve_geo(d, &g) != HEBB_OK) return -1;
  if (g.transposed || prec == HEBB_PREC_FP32) return 0;
  if (fused_wgrad_supported(g)) return 2;
  return use_tc(g, prec) ? 1 : 0;
}

int hebb_debug_fused_plan(const HebbDesc* d, int* out, int n) {
  Geo g;
  if (resolve_geo(d, &g) != HEBB_OK || !out) return 0;
  return fused_describe_plan(g, out, n);
}

int hebb_debug_plan(const HebbDesc* d, int prec, int* out, int n) {
  Geo g;
  if (resolve_geo(d, &g) != HEBB_OK || !out) return 0;
  if (!use_tc(g, prec)) return 0;
  return tc_describe_plan(g, prec, out, n);
}

int hebb_wnorm(const float* W, float* Wn, float* inv_norm, int64_t rows, int64_t row_stride, int64_t mid,
               int64_t mid_stride, int64_t inner, void* stream) {
  HEBB_TRY(device_ok());
  if (!W || (!Wn && !inv_norm)) return HEBB_EARG;
  if (rows < 0 || mid <= 0 || inner <= 0) return HEBB_ESHAPE;
  return launch_wnorm(W, Wn, inv_norm, rows, row_stride, mid, mid_stride, inner, (cudaStream_t)stream);
}

int hebb_conv_swta_step(const HebbDesc* d, const float* x, const float* W, const float* bias, float kinv,
                        float* y, int32_t* winner, float* delta_w, void* ws, size_t ws_bytes,
                        unsigned flags, int prec, void* stream) {
  HEBB_TRY(device_ok());
  Geo g;
  HEBB_TRY(resolve_geo(d, &g));
  if (g.transposed) return HEBB_EARG;
  if (!x || !W || !y) return HEBB_EARG;
  if ((flags & HEBB_F_UPDATE) && !delta_w) return HEBB_EARG;
  if (prec < HEBB_PREC_FP32 || prec > HEBB_PREC_BF16) return HEBB_EARG;
  if (!aligned16(ws)) return HEBB_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  flags &= 0xFFFFu;                       // bits above are library-internal
  if (prec != HEBB_PREC_FP32 && fused_supported(g, prec, flags))     // small-channel 2-D layers: one fused kernel
    return fused_conv_step(g, x, W, bias, kinv, y, winner, delta_w, ws, ws_bytes, flags, st);
  if (use_tc(g, prec)) {      // HPCA layers the planner cannot give a Gram plan fall through to the fp32 kernels
    const int s = tc_conv_step(g, x, W, bias, kinv, y, winner, delta_w, ws, ws_bytes, flags, prec, st);
    if (!(s == HEBB_ESHAPE && (flags & HEBB_F_RULE_HPCA))) return s;
  }
  return simt_conv_step(g, x, W, bias, kinv, y, winner, delta_w, ws, ws_bytes, flags, st);
}

int hebb_conv_swta_step_stats(const HebbDesc* d, const float* x, const float* W, const float* bias, float kinv,
                              float* y, int32_t* winner, float* delta_w, void* ws, size_t ws_bytes,
                              unsigned flags, int prec, double* y_stats, int* y_stats_written, void* stream) {
  if (y_stats_written) *y_stats_written = 0;
  HEBB_TRY(device_ok());
  Geo g;
  HEBB_TRY(resolve_geo(d, &g));
  flags &= 0xFFFFu;
  if (!y_stats || g.transposed || !use_tc(g, prec) || (flags & HEBB_F_RULE_HPCA))      // nothing to fuse: plain step
    return hebb_conv_swta_step(d, x, W, bias, kinv, y, winner, delta_w, ws, ws_bytes, flags, prec, stream);
  if (!x || !W || !y) return HEBB_EARG;
  if ((flags & HEBB_F_UPDATE) && !delta_w) return HEBB_EARG;
  if (prec < HEBB_PREC_FP32 || prec > HEBB_PREC_BF16) return HEBB_EARG;
  if (!aligned16(ws)) return HEBB_EALIGN;
  if (fused_supported(g, prec, flags))
    return fused_conv_step(g, x, W, bias, kinv, y, winner, delta_w, ws, ws_bytes, flags, (cudaStream_t)stream, y_stats,
                           y_stats_written);
  return tc_conv_step(g, x, W, bias, kinv, y, winner, delta_w, ws, ws_bytes, flags, prec, (cudaStream_t)stream, 0,
                      y_stats, y_stats_written);
}

int hebb_conv_wgrad(const HebbDesc* d, const float* x, const float* grad_y, float* grad_w, int gy_channels,
                    int channels_last, void* ws, size_t ws_bytes, int prec, void* stream) {
  HEBB_TRY(device_ok());
  Geo g;
  HEBB_TRY(resolve_geo(d, &g));
  if (g.transposed) return HEBB_ESHAPE;
  if (!x || !grad_y || !grad_w) return HEBB_EARG;
  if (prec != HEBB_PREC_BF16X3 && prec != HEBB_PREC_BF16) return HEBB_EARG;
  if (gy_channels < 0 || gy_channels > g.Cout) return HEBB_EARG;
  if (!aligned16(ws)) return HEBB_EALIGN;
  // few-channel layers reduced over many pixels (the back-prop head): the fused kernel with dL/dy as the responses
  if (fused_wgrad_supported(g))
    return fused_conv_wgrad(g, x, grad_y, grad_w, gy_channels, channels_last, ws, ws_bytes, (cudaStream_t)stream);
  if (!use_tc(g, prec)) return HEBB_ESHAPE;          // shapes outside the tcgen05 planner: caller's choice what to do
  if (channels_last && g.Cin <= 4 && g.taps > 1) return HEBB_ESHAPE;   // the patch-gathering pack reads NCHW only
  const int aux = (gy_channels & 0xFFFF) | ((channels_last ? 1 : 0) << 16);
  return tc_conv_step(g, x, grad_w, nullptr, 1.f, const_cast<float*>(grad_y), nullptr, grad_w, ws, ws_bytes,
                      HEBB_F_WGRAD_INTERNAL, prec, (cudaStream_t)stream, aux);
}

int hebb_convT_swta_step(const HebbDesc* d, const float* x, const float* W, const float* bias, float kinv,
                         float* y, int32_t* winner, float* delta_w, void* ws, size_t ws_bytes,
                         unsigned flags, int prec, void* stream) {
  HEBB_TRY(device_ok());
  Geo g;
  HEBB_TRY(resolve_geo(d, &g));
  if (!g.transposed) return HEBB_EARG;
  if (!x || !W || !y) return HEBB_EARG;
  if ((flags & HEBB_F_UPDATE) && !delta_w) return HEBB_EARG;
  if (prec < HEBB_PREC_FP32 || prec > HEBB_PREC_BF16) return HEBB_EARG;
  if (!aligned16(ws)) return HEBB_EALIGN;
  if (use_tc(g, prec) && !(flags & HEBB_F_RULE_HPCA))
    return tc_conv_step(g, x, W, bias, kinv, y, winner, delta_w, ws, ws_bytes, flags, prec, (cudaStream_t)stream);
  return simt_convT_step(g, x, W, bias, kinv, y, winner, delta_w, ws, ws_bytes, flags, (cudaStream_t)stream);
}

int hebb_local_update_multi(int n, float* const* grad, float* const* dw, const int64_t* numel,
                            const float* alpha, const int32_t* has_grad, void* stream) {
  HEBB_TRY(device_ok());
  if (n < 0) return HEBB_EARG;
  if (n == 0) return HEBB_OK;
  if (!grad || !dw || !numel || !alpha || !has_grad) return HEBB_EARG;
  return launch_local_update_multi(n, grad, dw, numel, alpha, has_grad, (cudaStream_t)stream);
}

}  // extern "C"
