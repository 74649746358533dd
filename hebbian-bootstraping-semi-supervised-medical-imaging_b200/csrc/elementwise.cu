// Elementwise / normalisation stage of the Hebbian path (HBM-bound kernels).
//   wnorm_kernel            <- normalize()      hebb/hebb.py:10-13
//   local_update_kernel     <- local_update()   hebb/hebb.py:174-192 (all layers, one launch)
//   finalize_conv_kernel    <- the decay term   hebb/hebb.py:114-115
//   finalize_convT_kernel   <- the decay term   hebb/hebb.py:262-264
#include "common.cuh"

namespace hebb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per row.  Pass 1: sum of squares (fp32, tree order).  Pass 2 (optional): scale.
__global__ void __launch_bounds__(256)
wnorm_kernel(const float* __restrict__ W, float* __restrict__ Wn, float* __restrict__ inv_out,
             long long row_stride, long long mid, long long mid_stride, long long inner) {
  const long long r = blockIdx.x;
  const float* src = W + r * row_stride;
  const long long n = mid * inner;
  float acc = 0.f;
  const bool vec = (mid == 1) && ((inner & 3) == 0) && ((row_stride & 3) == 0) &&
                   ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
  if (vec) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (long long i = threadIdx.x; i < (inner >> 2); i += blockDim.x) {
      float4 v = __ldg(s4 + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      long long m = i / inner, j = i - m * inner;
      float v = __ldg(src + m * mid_stride + j);
      acc += v * v;
    }
  }
  __shared__ float part[8];
  __shared__ float s_inv, s_nrm;
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part[i];
    float nrm = sqrtf(t);
    if (nrm == 0.f) nrm = 1.f;            // hebb.py:12
    s_nrm = nrm;
    s_inv = 1.f / nrm;
    if (inv_out) inv_out[r] = s_inv;
  }
  __syncthreads();
  if (!Wn) return;
  // The reference divides (x / nrm); keep a true division so Wn matches it bit for bit
  // up to the norm's own rounding.
  const float nrm = s_nrm;
  float* dst = Wn + r * row_stride;
  if (vec && ((reinterpret_cast<uintptr_t>(Wn) & 15) == 0)) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (long long i = threadIdx.x; i < (inner >> 2); i += blockDim.x) {
      float4 v = __ldg(s4 + i);
      v.x /= nrm; v.y /= nrm; v.z /= nrm; v.w /= nrm;
      d4[i] = v;
    }
  } else {
    for (long long i = threadIdx.x; i < n; i += blockDim.x) {
      long long m = i / inner, j = i - m * inner;
      dst[m * mid_stride + j] = __ldg(src + m * mid_stride + j) / nrm;
    }
  }
}

int launch_wnorm(const float* W, float* Wn, float* inv, long long rows, long long row_stride,
                 long long mid, long long mid_stride, long long inner, cudaStream_t st) {
  if (rows <= 0) return HEBB_OK;
  wnorm_kernel<<<(unsigned)rows, 256, 0, st>>>(W, Wn, inv, row_stride, mid, mid_stride, inner);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

// ---------------------------------------------------------------------------------------
constexpr int kMaxUpd = 32;
struct UpdBatch {
  float* grad[kMaxUpd];
  float* dw[kMaxUpd];
  long long numel[kMaxUpd];
  float alpha[kMaxUpd];
  int has_grad[kMaxUpd];
};

__global__ void __launch_bounds__(256)
local_update_kernel(const __grid_constant__ UpdBatch p) {
  const int t = blockIdx.y;
  float* __restrict__ g = p.grad[t];
  float* __restrict__ d = p.dw[t];
  const long long n = p.numel[t];
  const float a = p.alpha[t];
  const float oma = 1.f - a;
  const bool hg = p.has_grad[t] != 0;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  const bool vec = (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(d)) & 15) == 0);
  long long done = 0;
  if (vec) {
    const long long n4 = n >> 2;
    float4* g4 = reinterpret_cast<float4*>(g);
    float4* d4 = reinterpret_cast<float4*>(d);
    for (long long i = tid; i < n4; i += nth) {
      float4 dv = d4[i];
      float4 gv;
      if (hg) {
        gv = g4[i];
        gv.x = oma * gv.x - a * dv.x; gv.y = oma * gv.y - a * dv.y;
        gv.z = oma * gv.z - a * dv.z; gv.w = oma * gv.w - a * dv.w;
      } else {
        gv.x = -a * dv.x; gv.y = -a * dv.y; gv.z = -a * dv.z; gv.w = -a * dv.w;
      }
      g4[i] = gv;
      d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    done = n4 << 2;
  }
  for (long long i = done + tid; i < n; i += nth) {
    float dv = d[i];
    g[i] = hg ? (oma * g[i] - a * dv) : (-a * dv);
    d[i] = 0.f;
  }
}

int launch_local_update_multi(int n, float* const* grad, float* const* dw, const int64_t* numel,
                              const float* alpha, const int32_t* has_grad, cudaStream_t st) {
  for (int base = 0; base < n; base += kMaxUpd) {
    UpdBatch b;
    const int cnt = (n - base < kMaxUpd) ? (n - base) : kMaxUpd;
    long long mx = 0;
    for (int i = 0; i < kMaxUpd; ++i) {
      const int j = base + (i < cnt ? i : 0);
      b.grad[i] = grad[j]; b.dw[i] = dw[j];
      b.numel[i] = (i < cnt) ? numel[j] : 0;
      b.alpha[i] = alpha[j]; b.has_grad[i] = has_grad[j];
      if (b.numel[i] > mx) mx = b.numel[i];
      if (i < cnt && (!grad[j] || !dw[j])) return HEBB_EARG;
    }
    long long gx = cdiv(mx, 256LL * 4 * 4);
    const long long cap = (long long)num_sms() * 8;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, (unsigned)cnt);
    local_update_kernel<<<grid, 256, 0, st>>>(b);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }
  return HEBB_OK;
}

// ---------------------------------------------------------------------------------------
// H is [Cout][K+1]; column K holds sum_p r[c,p].
__global__ void __launch_bounds__(256)
finalize_conv_kernel(const float* __restrict__ H, const float* __restrict__ W, float* __restrict__ dw,
                     int Cout, int K) {
  const long long n = (long long)Cout * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i / K);
    const int j = (int)(i - (long long)c * K);
    const float rs = H[(long long)c * (K + 1) + K];
    dw[i] += H[(long long)c * (K + 1) + j] - rs * W[i];
  }
}

int launch_finalize_conv(const float* H, const float* W, float* delta_w, int Cout, int K, cudaStream_t st) {
  const long long n = (long long)Cout * K;
  long long gx = cdiv(n, 256);
  if (gx > (long long)num_sms() * 8) gx = (long long)num_sms() * 8;
  finalize_conv_kernel<<<(unsigned)gx, 256, 0, st>>>(H, W, delta_w, Cout, K);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

// H is [(Cin+1)][Cout*taps]; row Cin holds sum_p r[co,off,p].  W/dw are the contiguous
// [Cout][Cin][taps] buffers under the reference's (Cin,Cout,k..) view.
// dec[ci,co] = sum_off rsum[co,off] * W[ci,co,off], broadcast to every off (hebb.py:262-263).
__global__ void __launch_bounds__(256)
finalize_convT_kernel(const float* __restrict__ H, const float* __restrict__ W, float* __restrict__ dw,
                      int Cin, int Cout, int taps) {
  const long long n = (long long)Cin * Cout;
  const long long N = (long long)Cout * taps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / Cin);
    const int ci = (int)(i - (long long)co * Cin);
    const float* w = W + ((long long)co * Cin + ci) * taps;
    float* d = dw + ((long long)co * Cin + ci) * taps;
    const float* rs = H + (long long)Cin * N + (long long)co * taps;
    const float* h = H + (long long)ci * N + (long long)co * taps;
    float dec = 0.f;
    for (int t = 0; t < taps; ++t) dec += rs[t] * w[t];
    for (int t = 0; t < taps; ++t) d[t] += h[t] - dec;
  }
}

int launch_finalize_convT(const float* H, const float* W, float* delta_w, int Cin, int Cout, int taps,
                          cudaStream_t st) {
  const long long n = (long long)Cin * Cout;
  long long gx = cdiv(n, 256);
  if (gx > (long long)num_sms() * 8) gx = (long long)num_sms() * 8;
  finalize_convT_kernel<<<(unsigned)gx, 256, 0, st>>>(H, W, delta_w, Cin, Cout, taps);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

}  // namespace hebb
