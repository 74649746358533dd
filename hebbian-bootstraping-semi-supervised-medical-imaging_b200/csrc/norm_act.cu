// SURVEY.md §8(f) row 2 — the bandwidth-bound ops that sit between two Hebbian convolutions in the
// reference's UNet blocks (models/networks_2d/unet.py:53-61,170-183; models/networks_3d/unet3d.py:97-125):
//   BatchNorm(train) -> (Leaky)ReLU            : hebb_bn_stats + hebb_bn_act_apply   (2 reads + 1 write of y)
//   Upsample(scale 2, bilinear, align_corners) : hebb_upsample2x_bilinear
// They are opt-in (hebb/fused.py); the reference API surface is untouched.
#include "common.cuh"
#include <curand_kernel.h>

namespace hebb {

// Per-channel sum and sum of squares of y[B][C][S].  grid = (splits, C); each block reduces a
// contiguous range of the B*S elements of its channel in fp32 and adds its partial in fp64.
__global__ void __launch_bounds__(256)
bn_stats_kernel(const float* __restrict__ y, double* __restrict__ sums, int C, long long S, long long BS,
                long long per_block) {
  const int c = blockIdx.y;
  long long beg = (long long)blockIdx.x * per_block;
  long long end = beg + per_block;
  if (end > BS) end = BS;
  float s1 = 0.f, s2 = 0.f;
  const bool vec = ((S & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((per_block & 3) == 0);
  if (vec) {
    for (long long i = beg + 4LL * threadIdx.x; i < end; i += 4LL * blockDim.x) {
      const long long b = i / S, s = i - b * S;
      const float4 v = *reinterpret_cast<const float4*>(y + (b * C + c) * S + s);
      s1 += (v.x + v.y) + (v.z + v.w);
      s2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
    }
  } else {
    for (long long i = beg + threadIdx.x; i < end; i += blockDim.x) {
      const long long b = i / S, s = i - b * S;
      const float v = y[(b * C + c) * S + s];
      s1 += v; s2 += v * v;
    }
  }
  __shared__ float p1[8], p2[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  if ((threadIdx.x & 31) == 0) { p1[threadIdx.x >> 5] = s1; p2[threadIdx.x >> 5] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int i = 0; i < 8; ++i) { a += p1[i]; b += p2[i]; }
    atomicAdd(sums + 2 * c, a);
    atomicAdd(sums + 2 * c + 1, b);
  }
}

// scale/shift per channel from the sums; also the running-stat update of nn.BatchNorm (momentum form)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ scale_shift,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   int C, double count, float eps, float momentum) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = sums[2 * c] / count;
  double var = sums[2 * c + 1] / count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  scale_shift[2 * c] = g * invstd;
  scale_shift[2 * c + 1] = b - (float)mean * g * invstd;
  if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
  if (running_var) {
    const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

// out = dropout_p(act(y * scale[c] + shift[c])): the normalise + activate pass with the nn.Dropout that follows it in the
// networks' blocks (models/networks_2d/unet.py:53-61) folded in.  No mask is kept: the pass is only taken where nothing
// back-propagates through it.  Philox stream as in bias_relu_dropout_fwd_kernel (device-resident {seed, launches} state).
__global__ void __launch_bounds__(256)
bn_act_dropout_apply_kernel(const float* __restrict__ y, float* __restrict__ out, const float* __restrict__ scale_shift,
                            int C, long long S, long long total, float slope, float p, float keep_scale,
                            const unsigned long long* __restrict__ state) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  curandStatePhilox4_32_10_t rng;
  curand_init(state[0], (unsigned long long)tid, state[1] << 24, &rng);
  const bool vec = ((S & 3) == 0) && (((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  if (vec) {
    for (long long i = tid * 4; i < total; i += nth * 4) {
      const int c = (int)((i / S) % C);
      const float sc = __ldg(scale_shift + 2 * c), sh = __ldg(scale_shift + 2 * c + 1);
      const float4 u = curand_uniform4(&rng);
      float4 v = *reinterpret_cast<const float4*>(y + i);
      v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
      v.x = v.x >= 0.f ? v.x : v.x * slope; v.y = v.y >= 0.f ? v.y : v.y * slope;
      v.z = v.z >= 0.f ? v.z : v.z * slope; v.w = v.w >= 0.f ? v.w : v.w * slope;
      v.x = u.x > p ? v.x * keep_scale : 0.f; v.y = u.y > p ? v.y * keep_scale : 0.f;      // curand_uniform is in (0, 1]
      v.z = u.z > p ? v.z * keep_scale : 0.f; v.w = u.w > p ? v.w * keep_scale : 0.f;
      *reinterpret_cast<float4*>(out + i) = v;
    }
  } else {
    for (long long i = tid; i < total; i += nth) {
      const int c = (int)((i / S) % C);
      float v = fmaf(y[i], scale_shift[2 * c], scale_shift[2 * c + 1]);
      v = v >= 0.f ? v : v * slope;
      out[i] = curand_uniform(&rng) > p ? v * keep_scale : 0.f;
    }
  }
}

// out = act(y * scale[c] + shift[c]), act(v) = v >= 0 ? v : slope * v   (slope 0 -> ReLU, 1 -> identity)
__global__ void __launch_bounds__(256)
bn_act_apply_kernel(const float* __restrict__ y, float* __restrict__ out, const float* __restrict__ scale_shift,
                    int C, long long S, long long total, float slope) {
  const bool vec = ((S & 3) == 0) && (((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(out)) & 15) == 0);
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  if (vec) {
    for (long long i = tid * 4; i < total; i += nth * 4) {
      const int c = (int)((i / S) % C);
      const float sc = __ldg(scale_shift + 2 * c), sh = __ldg(scale_shift + 2 * c + 1);
      float4 v = *reinterpret_cast<const float4*>(y + i);
      v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
      v.x = v.x >= 0.f ? v.x : v.x * slope; v.y = v.y >= 0.f ? v.y : v.y * slope;
      v.z = v.z >= 0.f ? v.z : v.z * slope; v.w = v.w >= 0.f ? v.w : v.w * slope;
      *reinterpret_cast<float4*>(out + i) = v;
    }
  } else {
    for (long long i = tid; i < total; i += nth) {
      const int c = (int)((i / S) % C);
      float v = fmaf(y[i], scale_shift[2 * c], scale_shift[2 * c + 1]);
      out[i] = v >= 0.f ? v : v * slope;
    }
  }
}

// nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True) on [N][H][W] planes (N = B*C).
// Source coordinate = o * (in-1)/(out-1), like ATen's area_pixel_compute_source_index(align_corners=true).
// A block works on whole output rows (256 / W of them at a time; a thread owns the output columns 2w, 2w+1 of its row and
// keeps their horizontal taps in registers), so the only integer division left is one 32-bit one per row: the first
// version decoded a flat 64-bit index per output pair and ran at a quarter of the HBM rate.
__global__ void __launch_bounds__(256)
upsample2x_bilinear_kernel(const float* __restrict__ in, float* __restrict__ out, unsigned rows, int H, int W) {
  const int OH = 2 * H, OW = 2 * W;
  const float ry = (OH > 1) ? (float)(H - 1) / (float)(OH - 1) : 0.f;
  const float rx = (OW > 1) ? (float)(W - 1) / (float)(OW - 1) : 0.f;
  const int rpb = W < 256 ? 256 / W : 1;                 // output rows per block and round
  const int lr = W < 256 ? (int)threadIdx.x / W : 0;
  if (lr >= rpb) return;
  const int w0 = W < 256 ? (int)threadIdx.x - lr * W : (int)threadIdx.x;
  for (int w = w0; w < W; w += 256) {
    int x0[2], x1[2]; float lx[2];
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float fx = rx * (float)(2 * w + k);
      x0[k] = (int)fx;
      x1[k] = x0[k] + (x0[k] < W - 1 ? 1 : 0);
      lx[k] = fx - (float)x0[k];
    }
    for (unsigned row = blockIdx.x * (unsigned)rpb + (unsigned)lr; row < rows; row += gridDim.x * (unsigned)rpb) {
      const unsigned n = row / (unsigned)OH;
      const int oy = (int)(row - n * (unsigned)OH);
      const float fy = ry * (float)oy;
      const int y0 = (int)fy;
      const int y1 = y0 + (y0 < H - 1 ? 1 : 0);
      const float ly = fy - (float)y0, hy = 1.f - ly;
      const float* r0 = in + ((long long)n * H + y0) * W;
      const float* r1 = in + ((long long)n * H + y1) * W;
      float o2[2];
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float hx = 1.f - lx[k];
        o2[k] = hy * (hx * __ldg(r0 + x0[k]) + lx[k] * __ldg(r0 + x1[k])) + ly * (hx * __ldg(r1 + x0[k]) + lx[k] * __ldg(r1 + x1[k]));
      }
      *reinterpret_cast<float2*>(out + (long long)row * OW + 2 * w) = make_float2(o2[0], o2[1]);
    }
  }
}

// 2x max pooling (kernel = stride = 2, no padding, floor mode) of [N][D][H][W] planes; POOL_D = false keeps D
// (2-D pooling).  NaNs propagate like torch's max_pool.  One thread per output element, 8-byte loads when W is even.
__device__ __forceinline__ float nanmax(float a, float b) { return (a > b || a != a) ? a : b; }

template <bool POOL_D, bool VEC>
__global__ void __launch_bounds__(256)
maxpool2x_kernel(const float* __restrict__ in, float* __restrict__ out, long long N, int D, int H, int W) {
  const int OD = POOL_D ? D / 2 : D, OH = H / 2, OW = W / 2;
  const long long total = N * OD * OH * OW;
  const long long HW = (long long)H * W;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int ow = (int)(idx % OW);
    long long t = idx / OW;
    const int oh = (int)(t % OH); t /= OH;
    const int od = (int)(t % OD);
    const long long n = t / OD;
    const float* p = in + (n * D + (POOL_D ? 2 * od : od)) * HW + (long long)(2 * oh) * W + 2 * ow;
    float m;
    if (VEC) {
      const float2 a = __ldg(reinterpret_cast<const float2*>(p)), b = __ldg(reinterpret_cast<const float2*>(p + W));
      m = nanmax(nanmax(a.x, a.y), nanmax(b.x, b.y));
      if (POOL_D) {
        const float2 c = __ldg(reinterpret_cast<const float2*>(p + HW)), d = __ldg(reinterpret_cast<const float2*>(p + HW + W));
        m = nanmax(m, nanmax(nanmax(c.x, c.y), nanmax(d.x, d.y)));
      }
    } else {
      m = nanmax(nanmax(__ldg(p), __ldg(p + 1)), nanmax(__ldg(p + W), __ldg(p + W + 1)));
      if (POOL_D) m = nanmax(m, nanmax(nanmax(__ldg(p + HW), __ldg(p + HW + 1)), nanmax(__ldg(p + HW + W), __ldg(p + HW + W + 1))));
    }
    out[idx] = m;
  }
}

// out = dropout(relu(z + bias[c])) in ONE pass (the Conv -> ReLU -> Dropout runs of the back-prop head:
// models/networks_2d/unet.py:449-457), with the 1-byte mask (kept AND positive) the backward needs.
// Element i belongs to channel (i / inner) % C: inner = 1 for channels_last storage, = spatial size for NCHW.
// Philox4x32-10 counter-based stream: (seed, thread index) -> independent of the launch geometry's order.
__global__ void __launch_bounds__(256)
bias_relu_dropout_fwd_kernel(const float* __restrict__ z, const float* __restrict__ bias, float* __restrict__ out,
                             uint8_t* __restrict__ mask, long long n, int C, long long inner, float p, float scale,
                             unsigned long long seed, const unsigned long long* __restrict__ state) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  curandStatePhilox4_32_10_t rng;
  // device-resident stream state {seed, launches so far}: each launch skips 2^24 draws per thread ahead of the last one
  // (a thread draws n / threads values), so a replayed CUDA graph still gets a fresh mask every time
  unsigned long long off = 0;
  if (state) { seed = state[0]; off = state[1] << 24; }
  curand_init(seed, (unsigned long long)tid, off, &rng);
  const bool vec = ((n & 3) == 0) && ((inner & 3) == 0 || (inner == 1 && (C & 3) == 0)) &&
                   (((reinterpret_cast<uintptr_t>(z) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(mask) & 3) == 0);
  if (vec) {
    for (long long i = tid * 4; i < n; i += nth * 4) {
      const float4 v = *reinterpret_cast<const float4*>(z + i);
      const float4 r = curand_uniform4(&rng);
      const int c0 = (int)((i / inner) % C);
      float b[4];
      if (inner == 1) { b[0] = __ldg(bias + c0); b[1] = __ldg(bias + c0 + 1); b[2] = __ldg(bias + c0 + 2); b[3] = __ldg(bias + c0 + 3); }
      else { b[0] = b[1] = b[2] = b[3] = __ldg(bias + c0); }
      const float x[4] = {v.x + b[0], v.y + b[1], v.z + b[2], v.w + b[3]};
      const float u[4] = {r.x, r.y, r.z, r.w};
      float o[4]; uint32_t m = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool keep = (u[k] > p) && (x[k] > 0.f);          // curand_uniform is in (0, 1]: P(keep) = 1 - p
        o[k] = keep ? x[k] * scale : 0.f;
        m |= (keep ? 1u : 0u) << (8 * k);
      }
      *reinterpret_cast<float4*>(out + i) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<uint32_t*>(mask + i) = m;
    }
  } else {
    for (long long i = tid; i < n; i += nth) {
      const float x = z[i] + __ldg(bias + (int)((i / inner) % C));
      const bool keep = (curand_uniform(&rng) > p) && (x > 0.f);
      out[i] = keep ? x * scale : 0.f;
      mask[i] = keep ? 1 : 0;
    }
  }
}

__global__ void dropout_state_bump_kernel(unsigned long long* state) { state[1] += 1; }

// gz = gout * mask * scale
__global__ void __launch_bounds__(256)
mask_scale_bwd_kernel(const float* __restrict__ gout, const uint8_t* __restrict__ mask, float* __restrict__ gz, long long n,
                      float scale) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  const bool vec = ((n & 3) == 0) && (((reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(gz)) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(mask) & 3) == 0);
  if (vec) {
    for (long long i = tid * 4; i < n; i += nth * 4) {
      const float4 g = *reinterpret_cast<const float4*>(gout + i);
      const uint32_t m = *reinterpret_cast<const uint32_t*>(mask + i);
      *reinterpret_cast<float4*>(gz + i) = make_float4((m & 0xffu) ? g.x * scale : 0.f, (m & 0xff00u) ? g.y * scale : 0.f,
                                                       (m & 0xff0000u) ? g.z * scale : 0.f, (m & 0xff000000u) ? g.w * scale : 0.f);
    }
  } else {
    for (long long i = tid; i < n; i += nth) gz[i] = mask[i] ? gout[i] * scale : 0.f;
  }
}

// gz = gout * mask * scale for channels_last storage (element i belongs to channel i % C), plus the per-channel sums of gz
// (the bias gradient of the convolution in front: it would otherwise be one more full pass over gz).  The grid stride is
// a multiple of C, so a thread meets the same four channels in every round: register sums, one fixed-order fold per
// block into partial[block][C], one fixed-order fold over the blocks by bias_partial_fold_kernel -- deterministic.
__global__ void __launch_bounds__(256)
mask_scale_gb_cl_kernel(const float* __restrict__ gout, const uint8_t* __restrict__ mask, float* __restrict__ gz,
                        float* __restrict__ partial, long long n, int C, float scale) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (long long i = tid * 4; i < n; i += nth * 4) {
    const float4 g = *reinterpret_cast<const float4*>(gout + i);
    const uint32_t m = *reinterpret_cast<const uint32_t*>(mask + i);
    const float4 o = make_float4((m & 0xffu) ? g.x * scale : 0.f, (m & 0xff00u) ? g.y * scale : 0.f,
                                 (m & 0xff0000u) ? g.z * scale : 0.f, (m & 0xff000000u) ? g.w * scale : 0.f);
    *reinterpret_cast<float4*>(gz + i) = o;
    a0 += o.x; a1 += o.y; a2 += o.z; a3 += o.w;
  }
  __shared__ float4 s[256];
  s[threadIdx.x] = make_float4(a0, a1, a2, a3);
  __syncthreads();
  const int q = C >> 2;                       // channel quads; 256 % q == 0: thread t holds quad t % q
  if ((int)threadIdx.x < q) {
    float4 acc = s[threadIdx.x];
    for (int t = threadIdx.x + q; t < 256; t += q) { const float4 v = s[t]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    *reinterpret_cast<float4*>(partial + (long long)blockIdx.x * C + 4 * threadIdx.x) = acc;
  }
}

// gb[c] = sum over the blocks' partial rows, in a fixed order: 1024 / C row groups each sum every (1024 / C)-th row, then
// one thread per channel folds the groups (a single thread per channel walking all ~1200 rows took 0.1 ms)
__global__ void __launch_bounds__(1024)
bias_partial_fold_kernel(const float* __restrict__ partial, float* __restrict__ gb, int rows, int C) {
  __shared__ float s[1024];
  const int G = 1024 / C;                       // C is a power of two <= 1024
  const int c = threadIdx.x % C, g = threadIdx.x / C;
  float acc = 0.f;
  for (int r = g; r < rows; r += G) acc += __ldg(partial + (long long)r * C + c);
  s[threadIdx.x] = acc;
  __syncthreads();
  if (g == 0) {
    for (int k = 1; k < G; ++k) acc += s[k * C + c];
    gb[c] = acc;
  }
}

}  // namespace hebb

using namespace hebb;

// the normalise + activate (+ dropout) pass shared by the two BatchNorm entry points
static int launch_bn_apply(const float* y, float* out, const float* ss, int64_t C, int64_t S, long long total, float slope, float p,
                           uint64_t* state, cudaStream_t st) {
  long long gx = cdiv(total, 256 * 4 * 4);
  const long long cap = (long long)num_sms() * 16;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  if (p > 0.f) {
    if (!state) return HEBB_EARG;
    if (!(p < 1.f)) return HEBB_ESHAPE;
    bn_act_dropout_apply_kernel<<<(unsigned)gx, 256, 0, st>>>(y, out, ss, (int)C, S, total, slope, p, 1.f / (1.f - p),
                                                             reinterpret_cast<const unsigned long long*>(state));
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
    dropout_state_bump_kernel<<<1, 1, 0, st>>>(reinterpret_cast<unsigned long long*>(state));
  } else {
    bn_act_apply_kernel<<<(unsigned)gx, 256, 0, st>>>(y, out, ss, (int)C, S, total, slope);
  }
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

static int bn_act_train_impl(const float* y, float* out, const float* gamma, const float* beta, float* running_mean,
                             float* running_var, int64_t B, int64_t C, int64_t S, float eps, float momentum, float slope,
                             float p, uint64_t* state, void* ws, size_t ws_bytes, void* stream);
static int bn_act_from_stats_impl(const float* y, float* out, const double* y_stats, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, int64_t B, int64_t C, int64_t S, float eps,
                                  float momentum, float slope, float p, uint64_t* state, void* ws, size_t ws_bytes, void* stream);

extern "C" {

int hebb_bn_act_train(const float* y, float* out, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, int64_t B, int64_t C, int64_t S, float eps, float momentum, float slope,
                      void* ws, size_t ws_bytes, void* stream) {
  return bn_act_train_impl(y, out, gamma, beta, running_mean, running_var, B, C, S, eps, momentum, slope, 0.f, nullptr, ws, ws_bytes, stream);
}

int hebb_bn_act_train_dropout(const float* y, float* out, const float* gamma, const float* beta, float* running_mean,
                              float* running_var, int64_t B, int64_t C, int64_t S, float eps, float momentum, float slope,
                              float p, uint64_t* state, void* ws, size_t ws_bytes, void* stream) {
  return bn_act_train_impl(y, out, gamma, beta, running_mean, running_var, B, C, S, eps, momentum, slope, p, state, ws, ws_bytes, stream);
}

int hebb_bn_act_from_stats(const float* y, float* out, const double* y_stats, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, int64_t B, int64_t C, int64_t S, float eps,
                           float momentum, float slope, void* ws, size_t ws_bytes, void* stream) {
  return bn_act_from_stats_impl(y, out, y_stats, gamma, beta, running_mean, running_var, B, C, S, eps, momentum, slope, 0.f, nullptr, ws,
                                ws_bytes, stream);
}

int hebb_bn_act_from_stats_dropout(const float* y, float* out, const double* y_stats, const float* gamma, const float* beta,
                                   float* running_mean, float* running_var, int64_t B, int64_t C, int64_t S, float eps,
                                   float momentum, float slope, float p, uint64_t* state, void* ws, size_t ws_bytes, void* stream) {
  return bn_act_from_stats_impl(y, out, y_stats, gamma, beta, running_mean, running_var, B, C, S, eps, momentum, slope, p, state, ws,
                                ws_bytes, stream);
}

}  // extern "C"

static int bn_act_train_impl(const float* y, float* out, const float* gamma, const float* beta, float* running_mean,
                             float* running_var, int64_t B, int64_t C, int64_t S, float eps, float momentum, float slope,
                             float p, uint64_t* state, void* ws, size_t ws_bytes, void* stream) {
  HEBB_TRY(device_ok());
  if (!y || !out || !ws) return HEBB_EARG;
  if (B <= 0 || C <= 0 || S <= 0 || C > 65535) return HEBB_ESHAPE;
  const size_t need = (size_t)C * (2 * sizeof(double) + 2 * sizeof(float));
  if (ws_bytes < need) return HEBB_EWS;
  if (reinterpret_cast<uintptr_t>(ws) & 15) return HEBB_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  double* sums = static_cast<double*>(ws);
  float* ss = reinterpret_cast<float*>(sums + 2 * C);
  HEBB_CUDA_TRY(cudaMemsetAsync(sums, 0, (size_t)C * 2 * sizeof(double), st));
  const long long BS = B * S;
  long long splits = ((long long)num_sms() * 8 + C - 1) / C;
  const long long max_splits = cdiv(BS, 4096);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  long long per_block = cdiv(cdiv(BS, splits), 4) * 4;
  splits = cdiv(BS, per_block);
  dim3 grid((unsigned)splits, (unsigned)C);
  bn_stats_kernel<<<grid, 256, 0, st>>>(y, sums, (int)C, S, BS, per_block);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  bn_finalize_kernel<<<(unsigned)cdiv(C, 128), 128, 0, st>>>(sums, gamma, beta, ss, running_mean, running_var, (int)C,
                                                             (double)BS, eps, momentum);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return launch_bn_apply(y, out, ss, C, S, B * C * S, slope, p, state, st);
}

static int bn_act_from_stats_impl(const float* y, float* out, const double* y_stats, const float* gamma, const float* beta,
                                  float* running_mean, float* running_var, int64_t B, int64_t C, int64_t S, float eps,
                                  float momentum, float slope, float p, uint64_t* state, void* ws, size_t ws_bytes, void* stream) {
  HEBB_TRY(device_ok());
  if (!y || !out || !ws || !y_stats) return HEBB_EARG;
  if (B <= 0 || C <= 0 || S <= 0 || C > 65535) return HEBB_ESHAPE;
  if (ws_bytes < (size_t)C * 2 * sizeof(float)) return HEBB_EWS;
  if (reinterpret_cast<uintptr_t>(ws) & 15) return HEBB_EALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  float* ss = static_cast<float*>(ws);
  bn_finalize_kernel<<<(unsigned)cdiv(C, 128), 128, 0, st>>>(y_stats, gamma, beta, ss, running_mean, running_var, (int)C,
                                                             (double)(B * S), eps, momentum);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return launch_bn_apply(y, out, ss, C, S, B * C * S, slope, p, state, st);
}

static int bias_relu_dropout_impl(const float* z, const float* bias, float* out, uint8_t* mask, int64_t n, int64_t C,
                                  int64_t inner, float p, uint64_t seed, uint64_t* state, void* stream) {
  HEBB_TRY(device_ok());
  if (!z || !bias || !out || !mask) return HEBB_EARG;
  if (n <= 0 || C <= 0 || inner <= 0 || !(p >= 0.f && p < 1.f)) return HEBB_ESHAPE;
  long long gx = cdiv(n, 256 * 4 * 2);
  const long long cap = (long long)num_sms() * 16;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  bias_relu_dropout_fwd_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(
      z, bias, out, mask, n, (int)C, inner, p, 1.f / (1.f - p), seed, reinterpret_cast<const unsigned long long*>(state));
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  if (state) {
    dropout_state_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(reinterpret_cast<unsigned long long*>(state));
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }
  return HEBB_OK;
}

extern "C" {

int hebb_bias_relu_dropout(const float* z, const float* bias, float* out, uint8_t* mask, int64_t n, int64_t C,
                           int64_t inner, float p, uint64_t seed, void* stream) {
  return bias_relu_dropout_impl(z, bias, out, mask, n, C, inner, p, seed, nullptr, stream);
}

int hebb_bias_relu_dropout_state(const float* z, const float* bias, float* out, uint8_t* mask, int64_t n, int64_t C,
                                 int64_t inner, float p, uint64_t* state, void* stream) {
  if (!state) return HEBB_EARG;
  return bias_relu_dropout_impl(z, bias, out, mask, n, C, inner, p, 0, state, stream);
}

int hebb_mask_scale(const float* gout, const uint8_t* mask, float* gz, int64_t n, float scale, void* stream) {
  HEBB_TRY(device_ok());
  if (!gout || !mask || !gz) return HEBB_EARG;
  if (n <= 0) return HEBB_ESHAPE;
  long long gx = cdiv(n, 256 * 4 * 2);
  const long long cap = (long long)num_sms() * 16;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  mask_scale_bwd_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(gout, mask, gz, n, scale);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

int hebb_mask_scale_gb(const float* gout, const uint8_t* mask, float* gz, float* gb, int64_t n, int64_t C, float scale,
                       float* partial, int64_t partial_rows, void* stream) {
  HEBB_TRY(device_ok());
  if (!gout || !mask || !gz || !gb || !partial) return HEBB_EARG;
  if (n <= 0 || C < 4 || C > 1024 || (C & (C - 1)) != 0 || n % C != 0 || partial_rows < 1) return HEBB_ESHAPE;
  if (((reinterpret_cast<uintptr_t>(gout) | reinterpret_cast<uintptr_t>(gz) | reinterpret_cast<uintptr_t>(partial)) & 15) ||
      (reinterpret_cast<uintptr_t>(mask) & 3))
    return HEBB_EALIGN;
  long long gx = cdiv(n, 256 * 4 * 2);
  const long long cap = (long long)num_sms() * 8;
  if (gx > cap) gx = cap;
  if (gx > partial_rows) gx = partial_rows;
  if (gx < 1) gx = 1;
  // (256 threads x 4 channels per round and block: the grid stride gx * 1024 is a multiple of every power-of-two C <= 1024)
  mask_scale_gb_cl_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(gout, mask, gz, partial, n, (int)C, scale);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  bias_partial_fold_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(partial, gb, (int)gx, (int)C);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

int hebb_upsample2x_bilinear(const float* in, float* out, int64_t N, int64_t H, int64_t W, void* stream) {
  HEBB_TRY(device_ok());
  if (!in || !out) return HEBB_EARG;
  if (N <= 0 || H <= 0 || W <= 0 || H > (1 << 20) || W > (1 << 20)) return HEBB_ESHAPE;
  const long long rows = N * 2 * H;                       // output rows
  if (rows >= (1LL << 31)) return HEBB_ESHAPE;
  const long long rpb = W < 256 ? 256 / W : 1;
  long long gx = cdiv(rows, rpb * 4);                     // ~4 rounds per block
  const long long cap = (long long)num_sms() * 16;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  upsample2x_bilinear_kernel<<<(unsigned)gx, 256, 0, (cudaStream_t)stream>>>(in, out, (unsigned)rows, (int)H, (int)W);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

int hebb_maxpool2x(const float* in, float* out, int64_t N, int64_t D, int64_t H, int64_t W, int pool_depth, void* stream) {
  HEBB_TRY(device_ok());
  if (!in || !out) return HEBB_EARG;
  if (N <= 0 || D <= 0 || H < 2 || W < 2 || (pool_depth && D < 2) || D > (1 << 20) || H > (1 << 20) || W > (1 << 20))
    return HEBB_ESHAPE;
  const long long total = N * (pool_depth ? D / 2 : D) * (H / 2) * (W / 2);
  long long gx = cdiv(total, 256);
  const long long cap = (long long)num_sms() * 32;
  if (gx > cap) gx = cap;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (W % 2 == 0) && ((reinterpret_cast<uintptr_t>(in) & 7) == 0);
  if (pool_depth) {
    if (vec) maxpool2x_kernel<true, true><<<(unsigned)gx, 256, 0, st>>>(in, out, N, (int)D, (int)H, (int)W);
    else maxpool2x_kernel<true, false><<<(unsigned)gx, 256, 0, st>>>(in, out, N, (int)D, (int)H, (int)W);
  } else {
    if (vec) maxpool2x_kernel<false, true><<<(unsigned)gx, 256, 0, st>>>(in, out, N, (int)D, (int)H, (int)W);
    else maxpool2x_kernel<false, false><<<(unsigned)gx, 256, 0, st>>>(in, out, N, (int)D, (int)H, (int)W);
  }
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

}  // extern "C"
