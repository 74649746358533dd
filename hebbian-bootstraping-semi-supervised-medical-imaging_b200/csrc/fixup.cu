// Exact resolution of near-tie winners (reference: y.argmax over channels, hebb/hebb.py:107 -- first maximum).
//
// The tensor-core forward computes y with a relative error of a few 1e-6 (bf16x3 split).  Where the two largest
// channel responses of a pixel are closer than that, its argmax may pick the runner-up.  The forward epilogues
// therefore append every pixel whose top-2 margin is below `tie_rel * max_c |y_c|` to a worklist, and this
// kernel re-evaluates those pixels from the fp32 inputs with fp64 accumulation in a fixed order and rewrites
// their winner entries.  Only a fraction ~1e-3 of the pixels is ever listed, so the pass costs microseconds.
#include "common.cuh"

namespace hebb {

struct FixGeo {
  int B, Cin, Cout;
  int iD, iH, iW, kD, kH, kW, sD, sH, sW, pD, pH, pW, oD, oH, oW;
  int taps, K, transposed;
  long long inS, outS;
};

struct FixParams {
  FixGeo g;
  const float* x; const float* W; const float* inv; const float* bias;
  int32_t* winner; const int* list; const int* count; int cap;
};

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One CTA per listed pixel; warp w evaluates channels w, w+8, ...; lanes stride over the filter taps k = (ci, tap).
__global__ void __launch_bounds__(256)
winner_fixup_kernel(const __grid_constant__ FixParams p) {
  extern __shared__ float s_x[];                 // the pixel's patch: K values (plain conv) / Cin values (transposed)
  __shared__ double s_best[8];
  __shared__ int s_bi[8];
  const FixGeo& g = p.g;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int n = *p.count;
  if (n > p.cap) n = p.cap;
  const int oHW = g.oH * g.oW;
  const long long iHW = (long long)g.iH * g.iW;
  for (int it = blockIdx.x; it < n; it += gridDim.x) {
    const long long pid = p.list[it];
    const int b = (int)(pid / g.outS);
    int s = (int)(pid - (long long)b * g.outS);
    const int od = s / oHW; s -= od * oHW;
    const int oh = s / g.oW;
    const int ow = s - oh * g.oW;
    const int kHW = g.kH * g.kW;
    const float* xb = p.x + (long long)b * g.Cin * g.inS;
    __syncthreads();                             // previous pixel fully consumed
    if (!g.transposed) {
      for (int k = threadIdx.x; k < g.K; k += blockDim.x) {
        const int ci = k / g.taps;
        int t = k - ci * g.taps;
        const int kd = t / kHW; t -= kd * kHW;
        const int kh = t / g.kW;
        const int kw = t - kh * g.kW;
        const int id = od * g.sD + kd - g.pD, ih = oh * g.sH + kh - g.pH, iw = ow * g.sW + kw - g.pW;
        float v = 0.f;
        if ((unsigned)id < (unsigned)g.iD && (unsigned)ih < (unsigned)g.iH && (unsigned)iw < (unsigned)g.iW)
          v = __ldg(xb + (long long)ci * g.inS + (long long)id * iHW + (long long)ih * g.iW + iw);
        s_x[k] = v;
      }
    } else {
      // y[b,co,o] = sum_{ci,tap} x[b,ci,(o - tap)/stride] W[ci,co,tap]: patch entry k = (ci, tap), 0 where the tap
      // does not hit an input voxel
      for (int k = threadIdx.x; k < g.K; k += blockDim.x) {
        const int ci = k / g.taps;
        int t = k - ci * g.taps;
        const int kd = t / kHW; t -= kd * kHW;
        const int kh = t / g.kW;
        const int kw = t - kh * g.kW;
        const int nd = od - kd, nh = oh - kh, nw = ow - kw;
        float v = 0.f;
        if (nd >= 0 && nh >= 0 && nw >= 0 && nd % g.sD == 0 && nh % g.sH == 0 && nw % g.sW == 0) {
          const int id = nd / g.sD - g.pD, ih = nh / g.sH - g.pH, iw = nw / g.sW - g.pW;
          if ((unsigned)id < (unsigned)g.iD && (unsigned)ih < (unsigned)g.iH && (unsigned)iw < (unsigned)g.iW)
            v = __ldg(xb + (long long)ci * g.inS + (long long)id * iHW + (long long)ih * g.iW + iw);
        }
        // the transposed layer normalises per INPUT channel (hebb.py:222-232 on the (Cin,Cout,k) view)
        s_x[k] = p.inv ? v * p.inv[ci] : v;
      }
    }
    __syncthreads();
    double best = -1e300;
    int bi = 0x7fffffff;
    for (int c = warp; c < g.Cout; c += 8) {
      const float* w = p.W + (long long)c * g.K;               // [Cout][Cin][taps] storage in both cases
      double acc = 0.0;
      for (int k = lane; k < g.K; k += 32) acc += (double)s_x[k] * (double)__ldg(w + k);
      acc = warp_sum_d(acc);
      if (!g.transposed && p.inv) acc *= (double)p.inv[c];
      if (p.bias) acc += (double)p.bias[c];
      const float yf = (float)acc;                              // the reference's y is an fp32 value
      if ((double)yf > best) { best = (double)yf; bi = c; }     // channels ascend within a warp: first maximum kept
    }
    if (lane == 0) { s_best[warp] = best; s_bi[warp] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double bb = s_best[0]; int bc = s_bi[0];
      for (int w = 1; w < 8; ++w)
        if (s_best[w] > bb || (s_best[w] == bb && s_bi[w] < bc)) { bb = s_best[w]; bc = s_bi[w]; }
      p.winner[pid] = bc;
    }
  }
}

int launch_winner_fixup(const Geo& g, const float* x, const float* W, const float* inv, const float* bias,
                        int32_t* winner, const int* list, const int* count, int cap, cudaStream_t st) {
  FixParams p;
  p.g.B = g.B; p.g.Cin = g.Cin; p.g.Cout = g.Cout;
  p.g.iD = g.iD; p.g.iH = g.iH; p.g.iW = g.iW; p.g.kD = g.kD; p.g.kH = g.kH; p.g.kW = g.kW;
  p.g.sD = g.sD; p.g.sH = g.sH; p.g.sW = g.sW; p.g.pD = g.pD; p.g.pH = g.pH; p.g.pW = g.pW;
  p.g.oD = g.oD; p.g.oH = g.oH; p.g.oW = g.oW; p.g.taps = g.taps; p.g.K = g.K; p.g.transposed = g.transposed;
  p.g.inS = g.inS; p.g.outS = g.outS;
  p.x = x; p.W = W; p.inv = inv; p.bias = bias; p.winner = winner; p.list = list; p.count = count; p.cap = cap;
  const size_t smem = (size_t)g.K * sizeof(float);
  if (smem > 200 * 1024) return HEBB_ESHAPE;
  HEBB_CUDA_TRY(cudaFuncSetAttribute(winner_fixup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  winner_fixup_kernel<<<2 * num_sms(), 256, smem, st>>>(p);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

}  // namespace hebb
