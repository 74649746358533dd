// fp32 CUDA-core path: exact-order fp32 implicit GEMMs.  This is the parity anchor
// (HEBB_PREC_FP32) and the path for geometries the tcgen05 kernels do not take
// (strided convs, transposed convs, Cout not a multiple of 8, ...).  No im2col is ever
// materialised: the patch matrix of hebb/hebb.py:105-106 is gathered on the fly.
//
// One generic 64x64x16 register-tiled kernel is instantiated with four "problems":
//   ConvFwd   y[P,Cout]      = X[P,K] * W^T                        hebb.py:75-80
//   ConvDw    H[Cout,K+1]    = r[Cout,P] * [X | 1][P,K+1]          hebb.py:114-115
//   ConvTFwd  y[Pout,Cout]   = Xup[Pout,Cin*taps] * W              hebb.py:226-232
//   ConvTDw   H[Cin+1,Cout*taps] = [x;1][Cin+1,Pin] * R[Pin,Cout*taps]   hebb.py:252-264
// The appended ones column/row makes sum_p r a by-product of the same contraction.
#include "common.cuh"

namespace hebb {

struct DevGeo {
  int B, Cin, Cout;
  int iD, iH, iW, kD, kH, kW, sD, sH, sW, pD, pH, pW, oD, oH, oW;
  int taps, K;
  int inS, outS, oHW, iHW, kHW;
  int xD, xH, xW, xS, xHW;   // zero-padded input extent (used by the transposed problems)
};

static DevGeo to_dev(const Geo& g) {
  DevGeo d;
  d.B = g.B; d.Cin = g.Cin; d.Cout = g.Cout;
  d.iD = g.iD; d.iH = g.iH; d.iW = g.iW; d.kD = g.kD; d.kH = g.kH; d.kW = g.kW;
  d.sD = g.sD; d.sH = g.sH; d.sW = g.sW; d.pD = g.pD; d.pH = g.pH; d.pW = g.pW;
  d.oD = g.oD; d.oH = g.oH; d.oW = g.oW; d.taps = g.taps; d.K = g.K;
  d.inS = (int)g.inS; d.outS = (int)g.outS; d.oHW = g.oH * g.oW; d.iHW = g.iH * g.iW; d.kHW = g.kH * g.kW;
  d.xD = g.iD + g.pD + g.qD; d.xH = g.iH + g.pH + g.qH; d.xW = g.iW + g.pW + g.qW;
  d.xHW = d.xH * d.xW; d.xS = d.xD * d.xHW;
  return d;
}

// value of the zero-padded input at patch p (output pixel index), column kk = (ci, kd, kh, kw)
__device__ __forceinline__ float patch_value(const DevGeo& g, const float* __restrict__ x, long long p, int kk) {
  const int b = (int)(p / g.outS);
  int s = (int)(p - (long long)b * g.outS);
  const int od = s / g.oHW; s -= od * g.oHW;
  const int oh = s / g.oW;
  const int ow = s - oh * g.oW;
  const int ci = kk / g.taps;
  int t = kk - ci * g.taps;
  const int kd = t / g.kHW; t -= kd * g.kHW;
  const int kh = t / g.kW;
  const int kw = t - kh * g.kW;
  const int id = od * g.sD + kd - g.pD;
  const int ih = oh * g.sH + kh - g.pH;
  const int iw = ow * g.sW + kw - g.pW;
  if ((unsigned)id >= (unsigned)g.iD || (unsigned)ih >= (unsigned)g.iH || (unsigned)iw >= (unsigned)g.iW) return 0.f;
  return __ldg(x + ((long long)(b * g.Cin + ci) * g.iD + id) * g.iHW + (long long)ih * g.iW + iw);
}

struct ConvFwd {
  DevGeo g; const float* x; const float* W; const float* inv; const float* bias; float* y;
  static constexpr bool kAMajorM = true;    // A contiguous along M (pixels)
  static constexpr bool kBMajorK = true;    // B contiguous along K (filter taps)
  static constexpr bool kAtomic = false;
  __device__ long long M() const { return (long long)g.B * g.outS; }
  __device__ int N() const { return g.Cout; }
  __device__ long long Kd() const { return g.K; }
  __device__ float a(long long m, long long k) const { return patch_value(g, x, m, (int)k); }
  __device__ float b(long long k, int n) const { return __ldg(W + (long long)n * g.K + k); }
  __device__ void store(long long m, int n, float v) const {
    const long long bb = m / g.outS;
    const long long s = m - bb * g.outS;
    if (inv) v *= inv[n];
    if (bias) v += bias[n];
    y[(bb * g.Cout + n) * g.outS + s] = v;
  }
};

struct ConvDw {
  DevGeo g; const float* x; const float* r; float* H;
  static constexpr bool kAMajorM = false;   // A = r[c][p]: contiguous along K (pixels)
  static constexpr bool kBMajorK = true;    // B = X[p][kk]: contiguous along K (pixels) for fixed kk
  static constexpr bool kAtomic = true;
  __device__ long long M() const { return g.Cout; }
  __device__ int N() const { return g.K + 1; }
  __device__ long long Kd() const { return (long long)g.B * g.outS; }
  __device__ float a(long long m, long long p) const {
    const long long bb = p / g.outS;
    return __ldg(r + (bb * g.Cout + m) * g.outS + (p - bb * g.outS));
  }
  __device__ float b(long long p, int n) const { return n == g.K ? 1.f : patch_value(g, x, p, n); }
  __device__ void store(long long m, int n, float v) const { atomicAdd(H + m * (g.K + 1) + n, v); }
};

// Transposed conv, forward, as a gather over (ci, tap): y[b,co,o] = sum x[b,ci,(o-k)/s] W[ci,co,k]
struct ConvTFwd {
  DevGeo g; const float* x; const float* W; const float* inv; const float* bias; float* y;
  static constexpr bool kAMajorM = true;
  static constexpr bool kBMajorK = true;
  static constexpr bool kAtomic = false;
  __device__ long long M() const { return (long long)g.B * g.outS; }
  __device__ int N() const { return g.Cout; }
  __device__ long long Kd() const { return g.K; }
  __device__ float a(long long m, long long kk) const {
    const int b = (int)(m / g.outS);
    int s = (int)(m - (long long)b * g.outS);
    const int od = s / g.oHW; s -= od * g.oHW;
    const int oh = s / g.oW;
    const int ow = s - oh * g.oW;
    const int ci = (int)kk / g.taps;
    int t = (int)kk - ci * g.taps;
    const int kd = t / g.kHW; t -= kd * g.kHW;
    const int kh = t / g.kW;
    const int kw = t - kh * g.kW;
    const int nd = od - kd, nh = oh - kh, nw = ow - kw;
    if (nd < 0 || nh < 0 || nw < 0) return 0.f;
    if (nd % g.sD || nh % g.sH || nw % g.sW) return 0.f;
    // coordinates in the zero-padded input the reference feeds to conv_transpose (hebb.py:88,232)
    const int id = nd / g.sD - g.pD, ih = nh / g.sH - g.pH, iw = nw / g.sW - g.pW;
    if ((unsigned)id >= (unsigned)g.iD || (unsigned)ih >= (unsigned)g.iH || (unsigned)iw >= (unsigned)g.iW) return 0.f;
    return __ldg(x + ((long long)(b * g.Cin + ci) * g.iD + id) * g.iHW + (long long)ih * g.iW + iw);
  }
  // memory is [Cout][Cin][taps]; normalisation is per input channel
  __device__ float b(long long kk, int n) const {
    const int ci = (int)kk / g.taps;
    const int t = (int)kk - ci * g.taps;
    float w = __ldg(W + ((long long)n * g.Cin + ci) * g.taps + t);
    return inv ? w * inv[ci] : w;
  }
  __device__ void store(long long m, int n, float v) const {
    const long long bb = m / g.outS;
    const long long s = m - bb * g.outS;
    if (bias) v += bias[n];
    y[(bb * g.Cout + n) * g.outS + s] = v;
  }
};

struct ConvTDw {
  DevGeo g; const float* x; const float* r; float* H;
  static constexpr bool kAMajorM = false;   // A = x[ci][p]: contiguous along K (input pixels)
  static constexpr bool kBMajorK = true;
  static constexpr bool kAtomic = true;
  __device__ long long M() const { return g.Cin + 1; }
  __device__ int N() const { return g.Cout * g.taps; }
  // the contraction runs over the pixels of the zero-PADDED input: halo pixels carry x = 0
  // but their r still counts in sum_p r (the ones row)
  __device__ long long Kd() const { return (long long)g.B * g.xS; }
  __device__ float a(long long m, long long p) const {
    if (m == g.Cin) return 1.f;
    const int bb = (int)(p / g.xS);
    int s = (int)(p - (long long)bb * g.xS);
    const int id = s / g.xHW - g.pD; s %= g.xHW;
    const int ih = s / g.xW - g.pH;
    const int iw = s % g.xW - g.pW;
    if ((unsigned)id >= (unsigned)g.iD || (unsigned)ih >= (unsigned)g.iH || (unsigned)iw >= (unsigned)g.iW) return 0.f;
    return __ldg(x + ((long long)(bb * g.Cin + (int)m) * g.iD + id) * g.iHW + (long long)ih * g.iW + iw);
  }
  __device__ float b(long long p, int n) const {
    const int b = (int)(p / g.xS);
    int s = (int)(p - (long long)b * g.xS);
    const int id = s / g.xHW; s -= id * g.xHW;
    const int ih = s / g.xW;
    const int iw = s - ih * g.xW;
    const int co = n / g.taps;
    int t = n - co * g.taps;
    const int kd = t / g.kHW; t -= kd * g.kHW;
    const int kh = t / g.kW;
    const int kw = t - kh * g.kW;
    const int od = id * g.sD + kd, oh = ih * g.sH + kh, ow = iw * g.sW + kw;
    return __ldg(r + ((long long)(b * g.Cout + co) * g.oD + od) * g.oHW + (long long)oh * g.oW + ow);
  }
  __device__ void store(long long m, int n, float v) const {
    atomicAdd(H + m * ((long long)g.Cout * g.taps) + n, v);
  }
};

// ---- HPCA (Sanger / generalised Hebbian rule), hebb/hebb.py:122-135, hebb/hebb3d.py:139-153 ----
//   delta_w += y X - (tril(y y^T)) W        (patchwise; y = the layer output incl. bias)
// G = y y^T  [Cout x Cout], contraction over pixels
struct YYt {
  DevGeo g; const float* y; float* G;
  static constexpr bool kAMajorM = false;
  static constexpr bool kBMajorK = true;
  static constexpr bool kAtomic = true;
  __device__ long long M() const { return g.Cout; }
  __device__ int N() const { return g.Cout; }
  __device__ long long Kd() const { return (long long)g.B * g.outS; }
  __device__ float a(long long m, long long p) const {
    const long long bb = p / g.outS;
    return __ldg(y + (bb * g.Cout + m) * g.outS + (p - bb * g.outS));
  }
  __device__ float b(long long p, int n) const {
    const long long bb = p / g.outS;
    return __ldg(y + (bb * g.Cout + n) * g.outS + (p - bb * g.outS));
  }
  __device__ void store(long long m, int n, float v) const { atomicAdd(G + m * g.Cout + n, v); }
};

// delta_w[c][j] += H[c][j] - sum_{c' <= c} G[c][c'] W[c'][j]      (H is [Cout][K+1])
struct HpcaDecay {
  DevGeo g; const float* G; const float* W; const float* H; float* dw;
  static constexpr bool kAMajorM = false;
  static constexpr bool kBMajorK = false;
  static constexpr bool kAtomic = false;
  __device__ long long M() const { return g.Cout; }
  __device__ int N() const { return g.K; }
  __device__ long long Kd() const { return g.Cout; }
  __device__ float a(long long m, long long c) const { return c <= m ? __ldg(G + m * g.Cout + c) : 0.f; }
  __device__ float b(long long c, int n) const { return __ldg(W + c * g.K + n); }
  __device__ void store(long long m, int n, float v) const {
    dw[m * g.K + n] += __ldg(H + m * (g.K + 1) + n) - v;
  }
};

// delta_w[c][j] -= sum_{c' <= c} G[c][c'] W[c'][j]   (decay alone: the tensor-core path adds y X itself)
struct HpcaDecayOnly {
  const float* G; const float* W; float* dw; int Cout, K;
  static constexpr bool kAMajorM = false;
  static constexpr bool kBMajorK = false;
  static constexpr bool kAtomic = false;
  __device__ long long M() const { return Cout; }
  __device__ int N() const { return K; }
  __device__ long long Kd() const { return Cout; }
  __device__ float a(long long m, long long c) const { return c <= m ? __ldg(G + m * Cout + c) : 0.f; }
  __device__ float b(long long c, int n) const { return __ldg(W + c * (long long)K + n); }
  __device__ void store(long long m, int n, float v) const { dw[m * K + n] -= v; }
};

// Transposed layers in mode 'hpca' use the conv rule with x and y exchanged (hebb.py:243-246): the layer
// INPUT is the response, the unfolded OUTPUT is the presynaptic patch.
// G = x x^T  [Cin x Cin], contraction over input pixels
struct XXt {
  DevGeo g; const float* x; float* G;
  static constexpr bool kAMajorM = false;
  static constexpr bool kBMajorK = true;
  static constexpr bool kAtomic = true;
  __device__ long long M() const { return g.Cin; }
  __device__ int N() const { return g.Cin; }
  __device__ long long Kd() const { return (long long)g.B * g.inS; }
  __device__ float a(long long m, long long p) const {
    const long long bb = p / g.inS;
    return __ldg(x + (bb * g.Cin + m) * g.inS + (p - bb * g.inS));
  }
  __device__ float b(long long p, int n) const {
    const long long bb = p / g.inS;
    return __ldg(x + (bb * g.Cin + n) * g.inS + (p - bb * g.inS));
  }
  __device__ void store(long long m, int n, float v) const { atomicAdd(G + m * g.Cin + n, v); }
};

// delta_w[co][ci][off] += H[ci][(co,off)] - sum_{ci' <= ci} G[ci][ci'] W[co][ci'][off]   (H is [Cin+1][Cout*taps])
struct HpcaDecayT {
  DevGeo g; const float* G; const float* W; const float* H; float* dw;
  static constexpr bool kAMajorM = false;
  static constexpr bool kBMajorK = true;
  static constexpr bool kAtomic = false;
  __device__ long long M() const { return g.Cin; }
  __device__ int N() const { return g.Cout * g.taps; }
  __device__ long long Kd() const { return g.Cin; }
  __device__ float a(long long m, long long c) const { return c <= m ? __ldg(G + m * g.Cin + c) : 0.f; }
  __device__ float b(long long c, int n) const {
    const int co = n / g.taps, off = n - co * g.taps;
    return __ldg(W + ((long long)co * g.Cin + c) * g.taps + off);
  }
  __device__ void store(long long m, int n, float v) const {
    const int co = n / g.taps, off = n - co * g.taps;
    dw[((long long)co * g.Cin + m) * g.taps + off] += __ldg(H + m * ((long long)g.Cout * g.taps) + n) - v;
  }
};

constexpr int BM = 64, BN = 64, BK = 16;

template <class Prob>
__global__ void __launch_bounds__(256)
simt_gemm_kernel(const __grid_constant__ Prob pr, long long k_per_split) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const long long M = pr.M();
  const int N = pr.N();
  const long long Kd = pr.Kd();
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const long long kbeg = (long long)blockIdx.z * k_per_split;
  long long kend = kbeg + k_per_split;
  if (kend > Kd) kend = Kd;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // 16 x 16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (long long k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- stage A tile (BM x BK) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mm, kk;
      if (Prob::kAMajorM) { mm = tid & 63; kk = (tid >> 6) + 4 * i; }
      else                { kk = tid & 15; mm = (tid >> 4) + 16 * i; }
      const long long m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < kend) ? pr.a(m, k) : 0.f;
    }
    // ---- stage B tile (BK x BN) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int nn, kk;
      if (Prob::kBMajorK) { kk = tid & 15; nn = (tid >> 4) + 16 * i; }
      else                { nn = tid & 63; kk = (tid >> 6) + 4 * i; }
      const int n = n0 + nn;
      const long long k = k0 + kk;
      Bs[kk][nn] = (n < N && k < kend) ? pr.b(k, n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) pr.store(m, n, acc[i][j]);
    }
  }
}

// ConvFwd/ConvTFwd write y with m varying fastest across ty: make the pixel index the
// fast thread index for coalesced stores by transposing the roles (pixels on tx).
template <class Prob>
__global__ void __launch_bounds__(256)
simt_gemm_pixfast_kernel(const __grid_constant__ Prob pr) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const long long M = pr.M();
  const int N = pr.N();
  const long long Kd = pr.Kd();
  const long long m0 = (long long)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // tx -> pixels (m), ty -> channels (n)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long long k0 = 0; k0 < Kd; k0 += BK) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int mm = tid & 63, kk = (tid >> 6) + 4 * i;
      const long long m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < Kd) ? pr.a(m, k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = tid & 15, nn = (tid >> 4) + 16 * i;
      const int n = n0 + nn;
      const long long k = k0 + kk;
      Bs[kk][nn] = (n < N && k < Kd) ? pr.b(k, n) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][tx + 16 * i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = Bs[kk][ty * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + ty * 4 + j;
    if (n >= N) continue;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long m = m0 + tx + 16 * i;
      if (m < M) pr.store(m, n, acc[i][j]);
    }
  }
}

// Soft-WTA over channels, one thread per pixel (hebb.py:107).  y is [B][C][S]; r gets the
// same layout; winner = argmax_c y with the lowest index winning ties (torch.argmax).
__global__ void __launch_bounds__(256)
swta_softmax_kernel(const float* __restrict__ y, float* __restrict__ r, int32_t* __restrict__ winner,
                    long long P, int C, long long S, float kinv) {
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P;
       p += (long long)gridDim.x * blockDim.x) {
    const long long b = p / S;
    const long long s = p - b * S;
    const float* yp = y + b * C * S + s;
    float mx = -INFINITY, best = -INFINITY;
    int bi = 0;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(yp + (long long)c * S);
      const float kv = v * kinv;
      mx = fmaxf(mx, kv);
      if (v > best) { best = v; bi = c; }
    }
    if (winner) winner[p] = bi;
    if (r) {
      float sum = 0.f;
      for (int c = 0; c < C; ++c) sum += expf(__ldg(yp + (long long)c * S) * kinv - mx);
      const float inv = 1.f / sum;
      float* rp = r + b * C * S + s;
      for (int c = 0; c < C; ++c) rp[(long long)c * S] = expf(__ldg(yp + (long long)c * S) * kinv - mx) * inv;
    }
  }
}

static long long pick_splits(long long tiles, long long Kd, int sms) {
  long long want = (long long)sms * 4 / (tiles > 0 ? tiles : 1);
  if (want < 1) want = 1;
  long long maxs = cdiv(Kd, 4 * BK);
  if (want > maxs) want = maxs;
  if (want > 65535) want = 65535;
  return want < 1 ? 1 : want;
}

// workspace: [inv_norm (max(Cout,Cin))] [r: B*Cout*outS] [H: (Cout*(K+1)) or ((Cin+1)*Cout*taps)]
struct SimtWs { float* inv; float* r; float* H; size_t h_bytes; float* G; };

static size_t h_elems(const Geo& g) {
  return g.transposed ? (size_t)(g.Cin + 1) * g.Cout * g.taps : (size_t)g.Cout * (g.K + 1);
}

size_t simt_workspace_bytes(const Geo& g) {
  size_t inv = align_up(sizeof(float) * (size_t)(g.Cout > g.Cin ? g.Cout : g.Cin), 256);
  size_t r = align_up(sizeof(float) * (size_t)g.B * g.Cout * g.outS, 256);
  size_t H = align_up(sizeof(float) * h_elems(g), 256);
  const size_t gd = g.transposed ? g.Cin : g.Cout;
  size_t G = align_up(sizeof(float) * gd * gd, 256);                      // HPCA only
  return inv + r + H + G;
}

static int carve(const Geo& g, void* ws, size_t ws_bytes, SimtWs* o) {
  if (!ws || ws_bytes < simt_workspace_bytes(g)) return HEBB_EWS;
  char* p = static_cast<char*>(ws);
  o->inv = reinterpret_cast<float*>(p);
  p += align_up(sizeof(float) * (size_t)(g.Cout > g.Cin ? g.Cout : g.Cin), 256);
  o->r = reinterpret_cast<float*>(p);
  p += align_up(sizeof(float) * (size_t)g.B * g.Cout * g.outS, 256);
  o->H = reinterpret_cast<float*>(p);
  o->h_bytes = sizeof(float) * h_elems(g);
  p += align_up(o->h_bytes, 256);
  o->G = reinterpret_cast<float*>(p);
  return HEBB_OK;
}

static unsigned softmax_grid(long long P) {
  long long gx = cdiv(P, 256);
  const long long cap = (long long)num_sms() * 16;
  return (unsigned)(gx > cap ? cap : (gx < 1 ? 1 : gx));
}

int simt_conv_step(const Geo& g, const float* x, const float* W, const float* bias, float kinv, float* y,
                   int32_t* winner, float* delta_w, void* ws, size_t ws_bytes, unsigned flags,
                   cudaStream_t st) {
  SimtWs w;
  HEBB_TRY(carve(g, ws, ws_bytes, &w));
  const DevGeo dg = to_dev(g);
  const long long P = (long long)g.B * g.outS;
  const bool upd = (flags & HEBB_F_UPDATE) != 0;
  if (flags & HEBB_F_WNRM)
    HEBB_TRY(launch_wnorm(W, nullptr, w.inv, g.Cout, g.K, 1, 0, g.K, st));
  ConvFwd f{dg, x, W, (flags & HEBB_F_WNRM) ? w.inv : nullptr, bias, y};
  dim3 grid((unsigned)cdiv(P, BM), (unsigned)cdiv(g.Cout, BN));
  simt_gemm_pixfast_kernel<ConvFwd><<<grid, 256, 0, st>>>(f);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  const bool hpca = (flags & HEBB_F_RULE_HPCA) != 0;
  if ((upd && !hpca) || winner) {
    swta_softmax_kernel<<<softmax_grid(P), 256, 0, st>>>(y, (upd && !hpca) ? w.r : nullptr, winner, P, g.Cout, g.outS, kinv);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }
  if (upd) {
    HEBB_CUDA_TRY(cudaMemsetAsync(w.H, 0, w.h_bytes, st));
    ConvDw d{dg, x, hpca ? y : w.r, w.H};       // HPCA: the response is y itself (hebb.py:127)
    const long long tiles = cdiv(g.Cout, BM) * cdiv(g.K + 1, BN);
    const long long splits = pick_splits(tiles, P, num_sms());
    long long kps = cdiv(cdiv(P, splits), BK) * BK;
    dim3 g2((unsigned)cdiv(g.Cout, BM), (unsigned)cdiv(g.K + 1, BN), (unsigned)cdiv(P, kps));
    simt_gemm_kernel<ConvDw><<<g2, 256, 0, st>>>(d, kps);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
    if (!hpca) {
      HEBB_TRY(launch_finalize_conv(w.H, W, delta_w, g.Cout, g.K, st));
    } else {
      HEBB_CUDA_TRY(cudaMemsetAsync(w.G, 0, sizeof(float) * (size_t)g.Cout * g.Cout, st));
      YYt yy{dg, y, w.G};
      const long long t2 = cdiv(g.Cout, BM) * cdiv(g.Cout, BN);
      const long long sp2 = pick_splits(t2, P, num_sms());
      long long kps2 = cdiv(cdiv(P, sp2), BK) * BK;
      dim3 g3((unsigned)cdiv(g.Cout, BM), (unsigned)cdiv(g.Cout, BN), (unsigned)cdiv(P, kps2));
      simt_gemm_kernel<YYt><<<g3, 256, 0, st>>>(yy, kps2);
      HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
      HpcaDecay hd{dg, w.G, W, w.H, delta_w};
      dim3 g4((unsigned)cdiv(g.Cout, BM), (unsigned)cdiv(g.K, BN), 1);
      simt_gemm_kernel<HpcaDecay><<<g4, 256, 0, st>>>(hd, (long long)cdiv(g.Cout, BK) * BK);
      HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
    }
  }
  return HEBB_OK;
}

int launch_hpca_decay(const float* G, const float* W, float* delta_w, int Cout, int K, cudaStream_t st) {
  HpcaDecayOnly hd{G, W, delta_w, Cout, K};
  dim3 grid((unsigned)cdiv(Cout, BM), (unsigned)cdiv(K, BN), 1);
  simt_gemm_kernel<HpcaDecayOnly><<<grid, 256, 0, st>>>(hd, (long long)cdiv(Cout, BK) * BK);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

int simt_convT_step(const Geo& g, const float* x, const float* W, const float* bias, float kinv, float* y,
                    int32_t* winner, float* delta_w, void* ws, size_t ws_bytes, unsigned flags,
                    cudaStream_t st) {
  SimtWs w;
  HEBB_TRY(carve(g, ws, ws_bytes, &w));
  const DevGeo dg = to_dev(g);
  const long long Pout = (long long)g.B * g.outS;
  const long long Pin = (long long)g.B * dg.xS;
  const bool upd = (flags & HEBB_F_UPDATE) != 0;
  // per-INPUT-channel norm over (Cout, taps) of the [Cout][Cin][taps] buffer (hebb3d.py:78 on the view)
  if (flags & HEBB_F_WNRM)
    HEBB_TRY(launch_wnorm(W, nullptr, w.inv, g.Cin, g.taps, g.Cout, (long long)g.Cin * g.taps, g.taps, st));
  ConvTFwd f{dg, x, W, (flags & HEBB_F_WNRM) ? w.inv : nullptr, bias, y};
  dim3 grid((unsigned)cdiv(Pout, BM), (unsigned)cdiv(g.Cout, BN));
  simt_gemm_pixfast_kernel<ConvTFwd><<<grid, 256, 0, st>>>(f);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  const bool hpca = (flags & HEBB_F_RULE_HPCA) != 0;
  if ((upd && !hpca) || winner) {
    swta_softmax_kernel<<<softmax_grid(Pout), 256, 0, st>>>(y, (upd && !hpca) ? w.r : nullptr, winner, Pout, g.Cout, g.outS, kinv);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }
  if (upd) {
    HEBB_CUDA_TRY(cudaMemsetAsync(w.H, 0, w.h_bytes, st));
    ConvTDw d{dg, x, hpca ? y : w.r, w.H};
    const int N = g.Cout * g.taps;
    const long long tiles = cdiv(g.Cin + 1, BM) * cdiv(N, BN);
    const long long splits = pick_splits(tiles, Pin, num_sms());
    long long kps = cdiv(cdiv(Pin, splits), BK) * BK;
    dim3 g2((unsigned)cdiv(g.Cin + 1, BM), (unsigned)cdiv(N, BN), (unsigned)cdiv(Pin, kps));
    simt_gemm_kernel<ConvTDw><<<g2, 256, 0, st>>>(d, kps);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
    if (!hpca) {
      HEBB_TRY(launch_finalize_convT(w.H, W, delta_w, g.Cin, g.Cout, g.taps, st));
    } else {
      const long long Px = (long long)g.B * g.inS;
      HEBB_CUDA_TRY(cudaMemsetAsync(w.G, 0, sizeof(float) * (size_t)g.Cin * g.Cin, st));
      XXt xx{dg, x, w.G};
      const long long t2 = cdiv(g.Cin, BM) * cdiv(g.Cin, BN);
      const long long sp2 = pick_splits(t2, Px, num_sms());
      long long kps2 = cdiv(cdiv(Px, sp2), BK) * BK;
      dim3 g3((unsigned)cdiv(g.Cin, BM), (unsigned)cdiv(g.Cin, BN), (unsigned)cdiv(Px, kps2));
      simt_gemm_kernel<XXt><<<g3, 256, 0, st>>>(xx, kps2);
      HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
      HpcaDecayT hd{dg, w.G, W, w.H, delta_w};
      dim3 g4((unsigned)cdiv(g.Cin, BM), (unsigned)cdiv(N, BN), 1);
      simt_gemm_kernel<HpcaDecayT><<<g4, 256, 0, st>>>(hd, (long long)cdiv(g.Cin, BK) * BK);
      HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
    }
  }
  return HEBB_OK;
}

}  // namespace hebb
