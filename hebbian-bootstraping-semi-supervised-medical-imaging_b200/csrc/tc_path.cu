// tcgen05 / TMEM path of the Hebbian convolution (stride-1 convs, Cout % 16 == 0, Cout <= 256).
//
// Both contractions are "shift GEMMs" over ONE packed copy of the activations, so no patch
// matrix (hebb/hebb.py:105-106, hebb/hebb3d.py:92-101) is ever built:
//
//   packed activations  Xp[hl][c8][p][8]  bf16   p = b*Qimg + (d*HP + h)*WP + w  over the ZERO-PADDED
//                                                 image (halo materialised once, 2 B/elem), channels in
//                                                 chunks of 8 (one 16-byte vector per position)
//   packed responses    Rp[hl][c8][p][8]  bf16   same position index, r = softmax_c(k*y), 0 where p is
//                                                 not a real output pixel
//   output pixel p, tap (kd,kh,kw)  ->  input position p + (kd*HP + kh)*WP + kw          (stride 1)
//
//   forward  D[p, co]        = sum_tap sum_ci Xp[ci][p + shift(tap)] * W[co][ci][tap]
//            A = 128 consecutive positions x 16 channels, K-major, SWIZZLE_NONE: a tap is just a
//            16-byte-granular start-address offset into the staged segment; B = packed weights.
//   dW       H[tap][ci, co]  = sum_p Xp[ci][p + shift(tap)] * Rp[co][p]
//            A = channels x 16 positions (MN-major), B = channels x 16 positions (MN-major).
//
// hl = 0/1 are the hi / lo halves of the bf16x3 split (v = hi + lo up to 2^-17); a product is
// hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM.
//
// Every global->shared transfer is a 1-D bulk async copy (TMA engine) of a contiguous byte
// range completing on an mbarrier; one thread issues tcgen05.mma; 4 warps run the epilogue
// out of TMEM (bias, y store, winner, softmax, bf16 split of r, column sums of r).
#include "common.cuh"
#include "umma.cuh"
#include <cstdlib>
#include <vector>

// the single-pass instantiations of the forward epilogue leave the generic multi-pass code unreachable
#pragma nv_diag_suppress 128

namespace hebb {

using namespace ptx;

constexpr int kMaxTaps = 27;
constexpr int kSmemLimit = 227 * 1024;

// Index (in 16-byte vectors) of the 8-channel group g8 of packed position pp in the response planes (the hi and
// the lo plane set use the same index).  rsw = 0: planes of 8 channels, one vector per position.  rsw = 1: planes of
// 64 channels, 128 bytes per position,
// the 16-byte chunks XOR-ed with (position % 8) -- the MN-major SWIZZLE_128B image that the dW kernel's B operand
// reads with a ONE-position atom stride, so that an N = 192 instruction covers the three taps of a kernel row
// without staging shifted copies (tests/test_umma_probe.py pins that descriptor behaviour).
__device__ __forceinline__ long long r_index(int rsw, long long PRS, int g8, long long pp) {
  return rsw ? (((long long)(g8 >> 3) * PRS + pp) * 8 + ((g8 & 7) ^ (int)(pp & 7))) : ((long long)g8 * PRS + pp);
}

// Appends winner-array index `pid` to the near-tie worklist (see fixup.cu).
__device__ __forceinline__ void flag_tie(int* list, int* count, int cap, long long pid) {
  const int i = atomicAdd(count, 1);
  if (i < cap) list[i] = (int)pid;
}

// -------------------------------------------------------------------------------------
// Packing kernels
// -------------------------------------------------------------------------------------
struct PackGeo {
  int B, Cin, iD, iH, iW, pD, pH, pW, HP, WP, plane, Qimg, CC;
  long long PA;   // positions per chunk plane (incl. zero slack)
  long long PTOT; // B*Qimg
  long long sB, sC, sD, sH, sW;   // element strides of x (NCHW: sC = inS, sW = 1; channels_last: sC = 1, sW = C)
  // zero pads of the packed responses (written here so the step needs no extra launch): in each of the C8 planes
  // of Rp, r_lead positions before and r_tail positions after the PR the forward pass writes; r_lead + r_tail = 0: none
  // r_ushift = 0: C8 planes of one vector per position; 3: C8/8 planes of 8 vectors per position (see r_index)
  uint4* rhi; uint4* rlo; int C8, r_lead, r_tail, r_ushift; long long PR, PRS;
};

__global__ void __launch_bounds__(256, 8)
pack_x_kernel(const float* __restrict__ x, uint4* __restrict__ xhi, uint4* __restrict__ xlo, const __grid_constant__ PackGeo g) {
  const long long total = (long long)g.CC * g.PA;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx / g.PA);
    const long long p = idx - (long long)c8 * g.PA;
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
    if (p < g.PTOT) {
      const int b = (int)(p / g.Qimg);
      int q = (int)(p - (long long)b * g.Qimg);
      const int d = q / g.plane; q -= d * g.plane;
      const int hh = q / g.WP;
      const int w = q - hh * g.WP;
      const int id = d - g.pD, ih = hh - g.pH, iw = w - g.pW;
      if ((unsigned)id < (unsigned)g.iD && (unsigned)ih < (unsigned)g.iH && (unsigned)iw < (unsigned)g.iW) {
        const float* src = x + (long long)b * g.sB + (long long)(c8 * 8) * g.sC + (long long)id * g.sD + (long long)ih * g.sH +
                           (long long)iw * g.sW;
        __nv_bfloat16 vh[8], vl[8];
        if (g.sC == 1 && c8 * 8 + 8 <= g.Cin && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {   // channels_last: 2 x 16 B
          const float4 a = __ldg(reinterpret_cast<const float4*>(src)), c = __ldg(reinterpret_cast<const float4*>(src) + 1);
          const float v8[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) split_bf16(v8[i], vh[i], vl[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float v = (c8 * 8 + i < g.Cin) ? __ldg(src + (long long)i * g.sC) : 0.f;
            split_bf16(v, vh[i], vl[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { h[i] = pack_bf16x2(vh[2 * i], vh[2 * i + 1]); l[i] = pack_bf16x2(vl[2 * i], vl[2 * i + 1]); }
      }
    }
    xhi[idx] = make_uint4(h[0], h[1], h[2], h[3]);
    if (xlo) xlo[idx] = make_uint4(l[0], l[1], l[2], l[3]);
  }
  // zero pads of Rp: one block per plane, no divisions (keeps the kernel at 32 registers per thread)
  const int per = (g.r_lead + g.r_tail) << g.r_ushift;
  const int planes = g.C8 >> g.r_ushift;
  for (int pl = blockIdx.x; pl < planes && per > 0; pl += gridDim.x) {
    uint4* ph = g.rhi + (((long long)pl * g.PRS) << g.r_ushift);
    uint4* plo = g.rlo ? g.rlo + (((long long)pl * g.PRS) << g.r_ushift) : nullptr;
    const int nlead = g.r_lead << g.r_ushift;
    for (int i = threadIdx.x; i < per; i += blockDim.x) {
      const long long off = i < nlead ? (long long)(i - nlead) : (g.PR << g.r_ushift) + (i - nlead);
      ph[off] = make_uint4(0, 0, 0, 0);
      if (plo) plo[off] = make_uint4(0, 0, 0, 0);
    }
  }
}

// Patch-gathering pack for layers with very few input channels (see equivalent_1x1): position p runs over
// the OUTPUT grid, pseudo-channel k' = ci*taps + tap holds x[ci] at the tap's input position (0 in the halo).
struct GatherGeo {
  int B, Cin, iD, iH, iW, pD, pH, pW, kH, kW, taps, oD, oH, oW, Kp, CC;
  long long PA, PTOT;
};

// KD/KH/KW > 0: compile-time kernel extent (3x3 and 3x3x3 get fully unrolled loops, so all loads of a
// position are in flight together); 0: run-time extent from g.
template <int KD, int KH, int KW>
__global__ void __launch_bounds__(256)
pack_x_gather_kernel(const float* __restrict__ x, uint4* __restrict__ xhi, uint4* __restrict__ xlo, const __grid_constant__ GatherGeo g) {
  // one thread per output position: walks (ci, kd, kh, kw) with counters (no divisions per element) and
  // shifts each bf16 into a 128-bit register window that is stored once per 8 pseudo-channels
  const long long iHW = (long long)g.iH * g.iW;
  const long long inS = (long long)g.iD * iHW;
  const int oHW = g.oH * g.oW;
  const long long oS = (long long)g.oD * oHW;
  const int kD = KD > 0 ? KD : g.taps / (g.kH * g.kW);
  const int kH = KH > 0 ? KH : g.kH, kW = KW > 0 ? KW : g.kW;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < g.PA;
       p += (long long)gridDim.x * blockDim.x) {
    const bool real = p < g.PTOT;
    int b = 0, od = 0, oh = 0, ow = 0;
    if (real) {
      b = (int)(p / oS);
      int q = (int)(p - (long long)b * oS);
      od = q / oHW; q -= od * oHW;
      oh = q / g.oW;
      ow = q - oh * g.oW;
    }
    unsigned long long h0 = 0, h1 = 0, l0 = 0, l1 = 0;     // 8 x bf16 windows (hi and lo parts), oldest in the low bits
    int filled = 0, c8 = 0;
    const float* xb = x + (long long)b * g.Cin * inS;
    for (int ci = 0; ci < g.Cin; ++ci) {
#pragma unroll
      for (int kd = 0; kd < kD; ++kd) {
        const int id = od + kd - g.pD;
#pragma unroll
        for (int kh = 0; kh < kH; ++kh) {
          const int ih = oh + kh - g.pH;
          const bool row_ok = real && (unsigned)id < (unsigned)g.iD && (unsigned)ih < (unsigned)g.iH;
          const float* row = xb + (long long)ci * inS + (long long)id * iHW + (long long)ih * g.iW;
#pragma unroll
          for (int kw = 0; kw < kW; ++kw) {
            const int iw = ow + kw - g.pW;
            float v = 0.f;
            if (row_ok && (unsigned)iw < (unsigned)g.iW) v = __ldg(row + iw);
            __nv_bfloat16 vh, vl;
            split_bf16(v, vh, vl);
            h0 = (h0 >> 16) | (h1 << 48); h1 = (h1 >> 16) | ((unsigned long long)__bfloat16_as_ushort(vh) << 48);
            l0 = (l0 >> 16) | (l1 << 48); l1 = (l1 >> 16) | ((unsigned long long)__bfloat16_as_ushort(vl) << 48);
            if (++filled == 8) {
              const long long idx = (long long)c8 * g.PA + p;
              xhi[idx] = make_uint4((uint32_t)h0, (uint32_t)(h0 >> 32), (uint32_t)h1, (uint32_t)(h1 >> 32));
              if (xlo) xlo[idx] = make_uint4((uint32_t)l0, (uint32_t)(l0 >> 32), (uint32_t)l1, (uint32_t)(l1 >> 32));
              filled = 0; ++c8;
            }
          }
        }
      }
    }
    for (; c8 < g.CC; ++c8) {          // zero-pad the last partly filled chunk and any wholly empty ones
      for (; filled < 8; ++filled) {
        h0 = (h0 >> 16) | (h1 << 48); h1 >>= 16;
        l0 = (l0 >> 16) | (l1 << 48); l1 >>= 16;
      }
      const long long idx = (long long)c8 * g.PA + p;
      xhi[idx] = make_uint4((uint32_t)h0, (uint32_t)(h0 >> 32), (uint32_t)h1, (uint32_t)(h1 >> 32));
      if (xlo) xlo[idx] = make_uint4((uint32_t)l0, (uint32_t)(l0 >> 32), (uint32_t)l1, (uint32_t)(l1 >> 32));
      filled = 0; h0 = h1 = l0 = l1 = 0;
    }
  }
}

// Packs an OUTPUT-shaped fp32 tensor (the back-propagated dL/dy of hebb_conv_wgrad) into the response layout
// Rp[hl][c8][p][8] over the padded position index space: 0 at positions that are not output pixels and in
// channels >= C (the packed channel count is a multiple of 16).
struct PackRGeo {
  int B, C, C8, oD, oH, oW, WP, plane, Qimg;     // C: channels actually present in gy (the rest of C8*8 is zero)
  long long outS, PR, PTOT;
  long long PRS;                                 // positions between the 8-channel planes of Rp (PR + zero pads)
  long long sB, sC, sD, sH, sW;                  // element strides of gy
};

__global__ void __launch_bounds__(256)
pack_r_kernel(const float* __restrict__ gy, uint4* __restrict__ rhi, uint4* __restrict__ rlo, const __grid_constant__ PackRGeo g) {
  const long long total = (long long)g.C8 * g.PR;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(idx / g.PR);
    const long long p = idx - (long long)c8 * g.PR;
    uint32_t h[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0};
    if (p < g.PTOT) {
      const int b = (int)(p / g.Qimg);
      int q = (int)(p - (long long)b * g.Qimg);
      const int od = q / g.plane; q -= od * g.plane;
      const int oh = q / g.WP;
      const int ow = q - oh * g.WP;
      if (od < g.oD && oh < g.oH && ow < g.oW) {
        const float* src = gy + (long long)b * g.sB + (long long)(c8 * 8) * g.sC + (long long)od * g.sD + (long long)oh * g.sH +
                           (long long)ow * g.sW;
        __nv_bfloat16 vh[8], vl[8];
        if (g.sC == 1 && c8 * 8 + 8 <= g.C && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(src)), c = __ldg(reinterpret_cast<const float4*>(src) + 1);
          const float v8[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) split_bf16(v8[i], vh[i], vl[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float v = (c8 * 8 + i < g.C) ? __ldg(src + (long long)i * g.sC) : 0.f;
            split_bf16(v, vh[i], vl[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { h[i] = pack_bf16x2(vh[2 * i], vh[2 * i + 1]); l[i] = pack_bf16x2(vl[2 * i], vl[2 * i + 1]); }
      }
    }
    const long long ridx = (long long)c8 * g.PRS + p;
    rhi[ridx] = make_uint4(h[0], h[1], h[2], h[3]);
    if (rlo) rlo[ridx] = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// Wp[slab][ct][tap][c2][hl][CT] (uint4 = 8 input channels) : the forward B operand, K-major.  For one
// (slab, channel tile) any run of consecutive taps is ONE contiguous block already in the shared-memory stage
// layout ([tap][k-chunk][hi|lo][CT rows]), so the producer moves a whole tap group with a single bulk copy.
__global__ void __launch_bounds__(256)
pack_w_kernel(const float* __restrict__ W, uint4* __restrict__ wp, int Cin, int Cout, int taps, int NSLAB, int HL,
              int CT, int tr_taps, int tr_q, const float* __restrict__ inv_ci) {
  const long long total = (long long)NSLAB * taps * HL * 2 * Cout;
  const int n_ct = Cout / CT;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long t = idx;
    const int cl = (int)(t % CT); t /= CT;
    const int hl = (int)(t % HL); t /= HL;
    const int c2 = (int)(t % 2); t /= 2;
    const int tap = (int)(t % taps); t /= taps;
    const int ct = (int)(t % n_ct); t /= n_ct;
    const int slab = (int)t;
    const int co = ct * CT + cl;
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat16 e[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ci = (slab * 2 + c2) * 8 + 2 * i + j;
        float v = 0.f;
        if (ci < Cin) {
          if (tr_taps) {   // transposed conv as a 1x1 conv onto (co, offset) channels; W is [Cout][Cin][taps]
            int cr, off;
            if (tr_q) { const int bk = co / tr_q, rem = co - bk * tr_q; cr = rem >> 1; off = 2 * bk + (rem & 1); }
            else { cr = co / tr_taps; off = co - cr * tr_taps; }
            v = __ldg(W + ((long long)cr * Cin + ci) * tr_taps + off) * (inv_ci ? inv_ci[ci] : 1.f);
          } else {
            v = __ldg(W + ((long long)co * Cin + ci) * taps + tap);
          }
        }
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        e[j] = hl ? lo : hi;
      }
      o[i] = pack_bf16x2(e[0], e[1]);
    }
    wp[idx] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Same packed layout for plain convolutions with LARGE weights, read coalesced: a block stages the 16 input
// channels x all taps of 16 filters (each a contiguous run of 16*taps floats of [Cout][Cin][taps]) in shared
// memory and emits 16 consecutive filters per 256-byte store run.  (The element-wise kernel above reads with a
// stride of Cin*taps floats between threads, which costs ~8x the bytes on the 57-113 MB bottleneck weights.)
__global__ void __launch_bounds__(256)
pack_w_tiled_kernel(const float* __restrict__ W, uint4* __restrict__ wp, int Cin, int Cout, int taps, int HL, int CT) {
  extern __shared__ float tile[];               // [16 filters][16*taps + 1]
  const int slab = blockIdx.x, co0 = blockIdx.y * 16;
  const int row = 16 * taps, pitch = row + 1;
  const int ci0 = slab * 16;
  const int nci = (Cin - ci0 < 16) ? (Cin - ci0) : 16;
  for (int i = threadIdx.x; i < 16 * row; i += blockDim.x) {
    const int f = i / row, j = i - f * row;
    tile[f * pitch + j] = (j < nci * taps) ? __ldg(W + ((long long)(co0 + f) * Cin + ci0) * taps + j) : 0.f;
  }
  __syncthreads();
  const int n_ct = Cout / CT;
  const int ct = co0 / CT, cl0 = co0 - ct * CT;
  const int items = taps * 2 * HL * 16;         // (tap, c2, hl, filter) with the filter fastest
  for (int i = threadIdx.x; i < items; i += blockDim.x) {
    const int f = i & 15;
    int t = i >> 4;
    const int hl = t % HL; t /= HL;
    const int c2 = t & 1;
    const int tap = t >> 1;
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat16 e[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float v = tile[f * pitch + (c2 * 8 + 2 * k + j) * taps + tap];
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        e[j] = hl ? lo : hi;
      }
      o[k] = pack_bf16x2(e[0], e[1]);
    }
    wp[((((long long)(slab * n_ct + ct) * taps + tap) * 2 + c2) * HL + hl) * CT + cl0 + f] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// -------------------------------------------------------------------------------------
// Forward shift-GEMM + fused soft-WTA epilogue
// -------------------------------------------------------------------------------------
struct FwdParams {
  const uint4* xp[2];
  const uint4* wp;
  uint4* rp[2];
  float* y; int32_t* winner; const float* inv; const float* bias; float* rsum; int* err;
  double* ystats;              // optional [Cout][2]: per-channel sum and sum of squares of y (BatchNorm statistics)
  int Cout, CC, NSLAB, taps, nseg, HL;
  int stackF;                  // bf16x3 forward: B = [w_hi | w_lo] stacked along N -> 2 MMAs instead of 3
  int CT, n_ct, fuse;          // output-channel tile handled by one CTA (<= 512 TMEM columns); fuse: softmax in the epilogue
  int tr, tD, tH, tW, CoutR;   // transposed conv (k == stride == 2): y scatter to the (tD,tH,tW) grid, CoutR real channels
                               // tr == 1: columns co*8 + off; tr == 2: columns (off>>1)*trQ + 2*co + (off&1), trQ = 2*CoutR
  int trQ;
  int RHL;                     // 1: r is consumed as single bf16 (hi only), 2: hi + lo
  long long PA, PR, PTOT;      // positions per chunk plane in Xp; packed output positions; real positions B*Qimg
  long long PRS;               // positions between the 8-channel planes of Rp (PR + the zero pads the dW kernel reads)
  int rsw;                     // responses in the 64-channel swizzled layout (see r_index)
  int MB, TILE_M, ntiles, SEGLEN;
  int XST, WST, NACC, WG;      // WG: taps per weight stage (one bulk copy)
  int WP, plane, Qimg, oD, oH, oW;
  float kinv; int write_r;
  int dbg;                     // HEBB_FWD_DBG bit mask (profiling only): 1 skip epilogue, 2 skip MMAs, 4 no y stores, 8 no r stores
  // near-tie worklist (fixup.cu): pixels whose top-2 margin is below tie_rel * max_c |y_c| are re-evaluated exactly
  int* fix_list; int* fix_count; int fix_cap; float tie_rel;
  int seg_base[4]; int seg_tap_begin[5]; int tap_off[kMaxTaps];
  uint32_t x_stage_bytes, w_stage_bytes, off_w, off_misc;
  uint32_t tmem_cols;
};

// Loads CH accumulator columns; with the stacked forward the x*w_lo partial lives `second` columns further on.
template <int CH>
__device__ __forceinline__ void ld_acc(uint32_t a, int second, uint32_t (&v)[CH]) {
  TmemLd<CH>::ld(a, v);
  if (second) {
    uint32_t w[CH];
    TmemLd<CH>::ld(a + second, w);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < CH; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
  } else {
    tmem_ld_wait();
  }
}

// One tap of one 16-channel slab for all M-blocks of the tile (called by the elected lane of the MMA warp).
// a_tap / wa are shared-memory byte addresses of the tap's first position and of the tap's weight tile.
// Operand-A collector hints: an instruction that shares its A tile (128 positions x 16 channels, 4 KB) with
// the next one keeps it and the next re-uses it instead of fetching it from shared memory again -- small-N
// instructions are bound by that fetch (profiles/README.md).
template <int ACC>
__device__ __forceinline__ void fwd_issue(const FwdParams& p, int mode, uint32_t a_tap, uint32_t wa, uint32_t d0, int cw,
                                          int NP, uint32_t a_lbo, uint32_t a_hi32, uint32_t b_lbo, uint32_t b_hi32,
                                          uint32_t idesc, uint32_t idesc2) {
  uint32_t ah = a_lbo | (a_tap >> 4);                              // x_hi tile of M-block 0
  uint32_t al = a_lbo | ((a_tap + 2u * p.SEGLEN * 16u) >> 4);      // x_lo tile
  const uint32_t bh = b_lbo | (wa >> 4);                           // w_hi rows
  const uint32_t bl = b_lbo | ((wa + (uint32_t)p.CT * 16u) >> 4);  // w_lo rows
  uint32_t d = d0;
  if (mode == 1) {                 // bf16x3, one N block: x_hi*w_lo + x_hi*w_hi + x_lo*w_hi
    for (int j = 0; j < p.MB; ++j, ah += 128u, al += 128u, d += cw) {
      umma_lo<1, ACC>(d, ah, a_hi32, bl, b_hi32, idesc);
      umma_lo<3, 1>(d, ah, a_hi32, bh, b_hi32, idesc);
      umma_lo<0, 1>(d, al, a_hi32, bh, b_hi32, idesc);
    }
  } else if (mode == 2) {          // two N halves (CT > 256): x_hi serves four instructions, x_lo two
    const uint32_t bh1 = bh + (uint32_t)NP, bl1 = bl + (uint32_t)NP;
    for (int j = 0; j < p.MB; ++j, ah += 128u, al += 128u, d += cw) {
      umma_lo<1, ACC>(d, ah, a_hi32, bl, b_hi32, idesc);
      umma_lo<2, 1>(d, ah, a_hi32, bh, b_hi32, idesc);
      umma_lo<2, ACC>(d + NP, ah, a_hi32, bl1, b_hi32, idesc);
      umma_lo<3, 1>(d + NP, ah, a_hi32, bh1, b_hi32, idesc);
      umma_lo<1, 1>(d, al, a_hi32, bh, b_hi32, idesc);
      umma_lo<3, 1>(d + NP, al, a_hi32, bh1, b_hi32, idesc);
    }
  } else if (mode == 3) {          // stacked [w_hi | w_lo]: x_hi*[w_hi|w_lo] -> columns [0,2CT) ; x_lo*w_hi -> [0,CT)
    for (int j = 0; j < p.MB; ++j, ah += 128u, al += 128u, d += cw) {
      umma_lo<0, ACC>(d, ah, a_hi32, bh, b_hi32, idesc2);
      umma_lo<0, 1>(d, al, a_hi32, bh, b_hi32, idesc);
    }
  } else {                         // single bf16 pass
    const int NH = p.CT / NP;
    for (int j = 0; j < p.MB; ++j, ah += 128u, d += cw)
      for (int h = 0; h < NH; ++h)
        umma_lo<0, ACC>(d + h * NP, ah, a_hi32, bh + (uint32_t)(h * NP), b_hi32, idesc);
  }
}

// SP > 0: the whole channel row (SP * CH <= 64 values) is held in registers -- a single pass out of TMEM with
// one exp per value and per-thread running column sums; SP == 0: generic multi-pass epilogue.
template <int CH, int SP>
__global__ void __launch_bounds__(320, 1)
fwd_swta_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  // misc region: [inv CT][bias CT][rs 8*CT][ysum CT][ysq CT] floats, then barriers, then tmem ptr
  float* s_inv = reinterpret_cast<float*>(smem + p.off_misc);
  float* s_bias = s_inv + p.CT;
  float* s_rs = s_bias + p.CT;
  float* s_ys = s_rs + 8 * p.CT;             // per-CTA sum of y per channel (shared by the 8 epilogue warps)
  float* s_yq = s_ys + p.CT;                 //         sum of y^2
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_yq + p.CT);
  const int total_work = p.ntiles * p.n_ct;      // work item = (position tile, output-channel tile)
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t x_full = bar0, x_empty = x_full + 8 * p.XST;
  const uint32_t w_full = x_empty + 8 * p.XST, w_empty = w_full + 8 * p.WST;
  const uint32_t t_full = w_empty + 8 * p.WST, t_empty = t_full + 8 * p.NACC;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * p.XST + 2 * p.WST + 2 * p.NACC);

  for (int i = threadIdx.x; i < 10 * p.CT; i += blockDim.x) s_rs[i] = 0.f;      // rs, ysum, ysq
  if (threadIdx.x == 0) {
    for (int i = 0; i < p.XST; ++i) { mbar_init(x_full + 8 * i, 1); mbar_init(x_empty + 8 * i, 1); }
    for (int i = 0; i < p.WST; ++i) { mbar_init(w_full + 8 * i, 1); mbar_init(w_empty + 8 * i, 1); }
    for (int i = 0; i < p.NACC; ++i) { mbar_init(t_full + 8 * i, 1); mbar_init(t_empty + 8 * i, 4); }
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(s_tmem), p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;

  if (warp == 0) {
    // ===================== producer: bulk copies =====================
    if (elect_one()) {
      int xs = 0, ws = 0; uint32_t xph = 0, wph = 0;
      const uint32_t w_tap_bytes = (uint32_t)p.HL * 2 * p.CT * 16;   // one tap: [k-chunk][hi|lo][CT rows]
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const int ct = work / p.ntiles, tile = work - ct * p.ntiles;   // channel tile slowest: it changes at most n_ct times per CTA
        const long long p0 = (long long)tile * p.TILE_M;
        const int slab0 = (p.dbg & 64) ? (int)(blockIdx.x % (unsigned)p.NSLAB) : 0;
        for (int s_i = 0; s_i < p.NSLAB; ++s_i) {
          const int slab = (s_i + slab0) % p.NSLAB;
          for (int seg = 0; seg < p.nseg; ++seg) {
            mbar_wait(x_empty + 8 * xs, xph ^ 1, p.err, 1);
            mbar_expect_tx(x_full + 8 * xs, (p.dbg & 16) ? 0u : p.x_stage_bytes);
            const uint32_t dst = sbase + xs * p.x_stage_bytes;
            for (int hl = 0; hl < ((p.dbg & 16) ? 0 : p.HL); ++hl)
              for (int c = 0; c < 2; ++c)
                bulk_g2s(dst + (hl * 2 + c) * p.SEGLEN * 16,
                         p.xp[hl] + (long long)(slab * 2 + c) * p.PA + p0 + p.seg_base[seg],
                         p.SEGLEN * 16, x_full + 8 * xs);
            if (++xs == p.XST) { xs = 0; xph ^= 1; }
            const int t_end = p.seg_tap_begin[seg + 1];
            for (int t0 = p.seg_tap_begin[seg]; t0 < t_end; t0 += p.WG) {
              const int nt = (t_end - t0 < p.WG) ? (t_end - t0) : p.WG;
              const uint32_t bytes = (uint32_t)nt * w_tap_bytes;
              mbar_wait(w_empty + 8 * ws, wph ^ 1, p.err, 2);
              mbar_expect_tx(w_full + 8 * ws, (p.dbg & 32) ? 0u : bytes);
              if (!(p.dbg & 32))
                bulk_g2s(sbase + p.off_w + ws * p.w_stage_bytes,
                         p.wp + ((long long)(slab * p.n_ct + ct) * p.taps + t0) * (w_tap_bytes >> 4), bytes,
                         w_full + 8 * ws);
              if (++ws == p.WST) { ws = 0; wph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp walks the (warp-uniform) loop nest; only the elected lane issues tcgen05.mma /
    // commit.  Keeping control flow uniform lets ptxas keep descriptors in uniform registers.
    {
      const int NH = (p.CT > 256) ? 2 : 1;          // one tcgen05.mma covers at most 256 output channels
      const int NP = p.CT / NH;
      const uint32_t idesc = idesc_bf16(128, NP, 0, 0);
      const uint32_t idesc2 = idesc_bf16(128, 2 * NP, 0, 0);      // stacked [w_hi | w_lo]
      const int cw = (p.stackF ? 2 : 1) * p.CT;                   // TMEM columns per M-block
      const uint64_t a_hi64 = smem_desc_hi(p.SEGLEN * 16, 128);   // LBO: chunk stride, SBO: 8 positions
      const uint64_t b_hi64 = smem_desc_hi(p.HL * p.CT * 16, 128);
      const uint32_t a_lbo = (uint32_t)a_hi64, a_hi32 = (uint32_t)(a_hi64 >> 32);
      const uint32_t b_lbo = (uint32_t)b_hi64, b_hi32 = (uint32_t)(b_hi64 >> 32);
      const int mode = p.stackF ? 3 : (p.HL == 2 ? (NH == 1 ? 1 : 2) : 0);
      const uint32_t w_tap_bytes = (uint32_t)p.HL * 2 * p.CT * 16;
      int xs = 0, ws = 0, acc = 0; uint32_t xph = 0, wph = 0, aph = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        mbar_wait(t_empty + 8 * acc, aph ^ 1, p.err, 3);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * p.MB * cw;
        uint32_t accum = 0u;
        for (int slab = 0; slab < p.NSLAB; ++slab) {
          for (int seg = 0; seg < p.nseg; ++seg) {
            mbar_wait(x_full + 8 * xs, xph, p.err, 4);
            const uint32_t xa = sbase + xs * p.x_stage_bytes;
            const int t_end = p.seg_tap_begin[seg + 1];
            for (int t0 = p.seg_tap_begin[seg]; t0 < t_end; t0 += p.WG) {
              const int nt = (t_end - t0 < p.WG) ? (t_end - t0) : p.WG;
              mbar_wait(w_full + 8 * ws, wph, p.err, 5);
              tc_fence_after();
              if (elect_one()) {
                const uint32_t wst = sbase + p.off_w + ws * p.w_stage_bytes;
                const int ntd = (p.dbg & 2) ? 0 : nt;
                if (accum == 0u && ntd > 0) {          // first tap of a tile overwrites the accumulators
                  fwd_issue<0>(p, mode, xa + p.tap_off[t0] * 16, wst, d0, cw, NP, a_lbo, a_hi32, b_lbo, b_hi32, idesc, idesc2);
                  for (int ti = 1; ti < ntd; ++ti)
                    fwd_issue<1>(p, mode, xa + p.tap_off[t0 + ti] * 16, wst + ti * w_tap_bytes, d0, cw, NP, a_lbo, a_hi32,
                                 b_lbo, b_hi32, idesc, idesc2);
                } else {
                  for (int ti = 0; ti < ntd; ++ti)
                    fwd_issue<1>(p, mode, xa + p.tap_off[t0 + ti] * 16, wst + ti * w_tap_bytes, d0, cw, NP, a_lbo, a_hi32,
                                 b_lbo, b_hi32, idesc, idesc2);
                }
                umma_commit(w_empty + 8 * ws);
              }
              __syncwarp();
              accum = 1u;
              if (++ws == p.WST) { ws = 0; wph ^= 1; }
            }
            if (elect_one()) umma_commit(x_empty + 8 * xs);
            __syncwarp();
            if (++xs == p.XST) { xs = 0; xph ^= 1; }
          }
        }
        if (elect_one()) umma_commit(t_full + 8 * acc);
        __syncwarp();
        if (++acc == p.NACC) { acc = 0; aph ^= 1; }
      }
    }
  } else {
    // ===================== epilogue warps (2..9): two sets of four =====================
    // A warp may only read its own TMEM lane quadrant (warp % 4), so the second set of four warps does not
    // split rows: with two accumulator buffers set h drains buffer h, i.e. the sets alternate work items and
    // each SM sub-partition has two epilogue warps to hide the tcgen05.ld / MUFU / store latencies.
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int ew = warp - 2;
    const int eset = ew >> 2;
    const int row = quad * 32 + lane;
    const long long outS = (long long)p.oD * p.oH * p.oW;
    const int oHW = p.oH * p.oW;
    float* my_rs = s_rs + ew * p.CT;
    uint32_t aph = 0;
    int cur_ct = -1;
    int it = 0;
    constexpr int NV = SP > 0 ? SP * CH : 1;
    float racc[NV];                           // SP > 0: this thread's running sum of r per channel
#pragma unroll
    for (int i = 0; i < NV; ++i) racc[i] = 0.f;
    // BatchNorm statistics of y (p.ystats): rows of up to 32 channels keep per-thread running sums like racc;
    // wider rows are folded per M-block by the butterfly (those layers are tensor-bound, the epilogue has slack)
    constexpr bool YACC = SP > 0 && NV <= 32;
    constexpr int NY = YACC ? NV : 1;
    float ysacc[NY], yqacc[NY];
#pragma unroll
    for (int i = 0; i < NY; ++i) { ysacc[i] = 0.f; yqacc[i] = 0.f; }
    const bool want_ys = p.ystats != nullptr;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x, ++it) {
      const int acc = (p.NACC == 2) ? (it & 1) : 0;
      if (acc != eset) continue;            // the other set's buffer (NACC == 1: set 1 has nothing to do)
      const int ct = work / p.ntiles, tile = work - ct * p.ntiles;   // channel tile slowest: it changes at most n_ct times per CTA
      const int cbase = ct * p.CT;
      if (ct != cur_ct) {      // (re)load this channel tile's 1/|W| and bias
        if (cur_ct >= 0 && p.fuse && p.write_r) {      // flush the finished tile's column sums
          __syncwarp();
          for (int c = lane; c < p.CT; c += 32) { atomicAdd(p.rsum + cur_ct * p.CT + c, my_rs[c]); my_rs[c] = 0.f; }
          __syncwarp();
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eset));
        for (int i = (int)threadIdx.x - 64 - eset * 128; i < p.CT; i += 128) {
          s_inv[i] = p.inv ? p.inv[cbase + i] : 1.f;
          s_bias[i] = p.bias ? p.bias[p.tr == 2 ? (((cbase + i) % p.trQ) >> 1) : (p.tr ? ((cbase + i) >> 3) : (cbase + i))] : 0.f;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + eset));
        cur_ct = ct;
      }
      mbar_wait(t_full + 8 * acc, aph, p.err, 6);
      tc_fence_after();
      for (int j = 0; j < ((p.dbg & 1) ? 0 : p.MB); ++j) {
        const long long pp = (long long)tile * p.TILE_M + j * 128 + row;
        const unsigned pp32 = (unsigned)pp;                 // the planner guarantees PR < 2^31: 32-bit divisions
        const int b = (int)(pp32 / (unsigned)p.Qimg);
        int q = (int)(pp32 - (unsigned)b * (unsigned)p.Qimg);
        const int od = (int)((unsigned)q / (unsigned)p.plane); q -= od * p.plane;
        const int oh = (int)((unsigned)q / (unsigned)p.WP);
        const int ow = q - oh * p.WP;
        const bool valid = (od < p.oD) && (oh < p.oH) && (ow < p.oW) && (pp < p.PTOT) && !(p.dbg & 4);
        const long long s = (long long)od * oHW + (long long)oh * p.oW + ow;
        float* yb = p.y + ((long long)b * p.Cout + cbase) * outS + s;
        const long long tS = (long long)p.tD * p.tH * p.tW;
        float* ytb = p.y + (long long)b * p.CoutR * tS + ((long long)(2 * od) * p.tH + 2 * oh) * p.tW + 2 * ow;
        const int cw = (p.stackF ? 2 : 1) * p.CT;
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * p.MB * cw + j * cw;
        float mx = -INFINITY, best = -INFINITY;
        float second = -INFINITY, amax = 0.f;       // runner-up and largest |y| of the pixel: near-tie test (fixup.cu)
        int bi = 0;
        float mx8[8], best8[8], second8[8], amax8[8];
        int bi8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { mx8[i] = -INFINITY; best8[i] = -INFINITY; bi8[i] = 0; second8[i] = -INFINITY; amax8[i] = 0.f; }
        uint32_t v[CH];
        if constexpr (SP > 0) {
          // ---- single pass: y = acc * 1/|w| + b, winner, r = softmax(k y), bf16 split, running column sums ----
          float f[NV];
#pragma unroll
          for (int ck = 0; ck < SP; ++ck) {
            ld_acc<CH>(ta + ck * CH, p.stackF ? p.CT : 0, v);
#pragma unroll
            for (int i4 = 0; i4 < CH; i4 += 4) {
              const float4 sc = *reinterpret_cast<const float4*>(s_inv + ck * CH + i4);
              const float4 bs = *reinterpret_cast<const float4*>(s_bias + ck * CH + i4);
              f[ck * CH + i4 + 0] = fmaf(__uint_as_float(v[i4 + 0]), sc.x, bs.x);
              f[ck * CH + i4 + 1] = fmaf(__uint_as_float(v[i4 + 1]), sc.y, bs.y);
              f[ck * CH + i4 + 2] = fmaf(__uint_as_float(v[i4 + 2]), sc.z, bs.z);
              f[ck * CH + i4 + 3] = fmaf(__uint_as_float(v[i4 + 3]), sc.w, bs.w);
            }
          }
          if (want_ys) {
            if constexpr (YACC) {
#pragma unroll
              for (int i = 0; i < NV; ++i) { const float t = valid ? f[i] : 0.f; ysacc[i] += t; yqacc[i] = fmaf(t, t, yqacc[i]); }
            } else {
#pragma unroll
              for (int ck = 0; ck < SP; ++ck) {
                float t1[CH], t2[CH];
#pragma unroll
                for (int i = 0; i < CH; ++i) { t1[i] = valid ? f[ck * CH + i] : 0.f; t2[i] = t1[i] * t1[i]; }
                const float a1 = lane_col_sum<CH>(t1, lane), a2 = lane_col_sum<CH>(t2, lane);
                if (lane < CH) { atomicAdd(s_ys + ck * CH + lane, a1); atomicAdd(s_yq + ck * CH + lane, a2); }
              }
            }
          }
          const float k2 = p.kinv * 1.4426950408889634f;      // exp(k y) = 2^(k2 y)
          float mx2 = -INFINITY;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            if (valid) yb[(long long)i * outS] = f[i];
            if (f[i] > best) { best = f[i]; bi = i; }          // strict: the lowest index wins ties
          }
          if (p.winner && valid) {
            p.winner[(long long)b * outS + s] = bi;
            float second = -INFINITY, amax = 0.f;              // runner-up and scale of this pixel
#pragma unroll
            for (int i = 0; i < NV; ++i) { second = fmaxf(second, i == bi ? -INFINITY : f[i]); amax = fmaxf(amax, fabsf(f[i])); }
            if (best - second <= p.tie_rel * amax) flag_tie(p.fix_list, p.fix_count, p.fix_cap, (long long)b * outS + s);
          }
#pragma unroll
          for (int i = 0; i < NV; ++i) { f[i] *= k2; mx2 = fmaxf(mx2, f[i]); }
          if (p.write_r) {
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
              float e;
              asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(f[i] - mx2));
              f[i] = e; sum += e;
            }
            const float rinv = valid ? (1.f / sum) : 0.f;
#pragma unroll
            for (int g8 = 0; g8 < NV / 8; ++g8) {
              uint32_t oh4[4], ol4[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int c = g8 * 8 + i * 2;
                const float r0 = f[c] * rinv, r1 = f[c + 1] * rinv;
                uint32_t hp, lp;
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(r1), "f"(r0));
                const float h0 = __uint_as_float(hp << 16), h1 = __uint_as_float(hp & 0xffff0000u);
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp) : "f"(r1 - h1), "f"(r0 - h0));
                oh4[i] = hp; ol4[i] = lp;
                if (p.RHL == 2) {
                  racc[c] += h0 + __uint_as_float(lp << 16);
                  racc[c + 1] += h1 + __uint_as_float(lp & 0xffff0000u);
                } else {
                  racc[c] += h0; racc[c + 1] += h1;
                }
              }
              if (p.dbg & 8) continue;
              const long long ridx = r_index(p.rsw, p.PRS, g8, pp);
              p.rp[0][ridx] = make_uint4(oh4[0], oh4[1], oh4[2], oh4[3]);
              if (p.RHL == 2) p.rp[1][ridx] = make_uint4(ol4[0], ol4[1], ol4[2], ol4[3]);
            }
          }
          continue;
        }
        if (p.tr == 2) {
          // ---- transposed layer, columns ordered (od2, oh2) block > channel > ow2: every block of trQ = 2*CoutR
          // columns holds two complete soft-WTA groups (one per x-parity), so any whole number of blocks is a
          // self-contained channel tile and wide layers still get the fused epilogue ----
          const int nblk = p.CT / p.trQ;
          for (int bk = 0; bk < nblk; ++bk) {
            const int off_hi = cbase / p.trQ + bk;            // (od2, oh2)
            const long long sub = ((long long)(off_hi >> 1) * p.tH + (off_hi & 1)) * p.tW;
            float* yo = ytb + sub;
            const int cb = bk * p.trQ;
            float mxa = -INFINITY, mxb = -INFINITY, besta = -INFINITY, bestb = -INFINITY;
            float seca = -INFINITY, secb = -INFINITY, amaxa = 0.f, amaxb = 0.f;
            int bia = 0, bib = 0;
            for (int c0 = cb; c0 < cb + p.trQ; c0 += CH) {
              ld_acc<CH>(ta + c0, 0, v);
#pragma unroll
              for (int i = 0; i < CH; i += 2) {
                const int co = (c0 - cb + i) >> 1;
                float2 o;
                o.x = __uint_as_float(v[i]) + s_bias[c0 + i];
                o.y = __uint_as_float(v[i + 1]) + s_bias[c0 + i + 1];
                if (valid) *reinterpret_cast<float2*>(yo + (long long)co * tS) = o;
                mxa = fmaxf(mxa, o.x * p.kinv); mxb = fmaxf(mxb, o.y * p.kinv);
                if (o.x > besta) { seca = besta; besta = o.x; bia = co; } else seca = fmaxf(seca, o.x);
                if (o.y > bestb) { secb = bestb; bestb = o.y; bib = co; } else secb = fmaxf(secb, o.y);
                amaxa = fmaxf(amaxa, fabsf(o.x)); amaxb = fmaxf(amaxb, fabsf(o.y));
              }
            }
            if (p.winner && valid) {
              int32_t* wb = p.winner + (long long)b * tS + ((long long)(2 * od) * p.tH + 2 * oh) * p.tW + 2 * ow + sub;
              *reinterpret_cast<int2*>(wb) = make_int2(bia, bib);
              if (besta - seca <= p.tie_rel * amaxa) flag_tie(p.fix_list, p.fix_count, p.fix_cap, wb - p.winner);
              if (bestb - secb <= p.tie_rel * amaxb) flag_tie(p.fix_list, p.fix_count, p.fix_cap, wb - p.winner + 1);
            }
            if (!p.write_r) continue;
            float suma = 0.f, sumb = 0.f;
            for (int c0 = cb; c0 < cb + p.trQ; c0 += CH) {
              ld_acc<CH>(ta + c0, 0, v);
#pragma unroll
              for (int i = 0; i < CH; i += 2) {
                suma += __expf(fmaf(__uint_as_float(v[i]) + s_bias[c0 + i], p.kinv, -mxa));
                sumb += __expf(fmaf(__uint_as_float(v[i + 1]) + s_bias[c0 + i + 1], p.kinv, -mxb));
              }
            }
            const float ria = valid ? (1.f / suma) : 0.f, rib = valid ? (1.f / sumb) : 0.f;
            for (int c0 = cb; c0 < cb + p.trQ; c0 += CH) {
              ld_acc<CH>(ta + c0, 0, v);
              float rr[CH];
#pragma unroll
              for (int g8 = 0; g8 < CH / 8; ++g8) {
                uint32_t oh4[4], ol4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const int c = g8 * 8 + i * 2;
                  const float r0 = __expf(fmaf(__uint_as_float(v[c]) + s_bias[c0 + c], p.kinv, -mxa)) * ria;
                  const float r1 = __expf(fmaf(__uint_as_float(v[c + 1]) + s_bias[c0 + c + 1], p.kinv, -mxb)) * rib;
                  __nv_bfloat16 h0, l0, h1, l1;
                  split_bf16(r0, h0, l0);
                  split_bf16(r1, h1, l1);
                  rr[c] = __bfloat162float(h0) + (p.RHL == 2 ? __bfloat162float(l0) : 0.f);
                  rr[c + 1] = __bfloat162float(h1) + (p.RHL == 2 ? __bfloat162float(l1) : 0.f);
                  oh4[i] = pack_bf16x2(h0, h1);
                  ol4[i] = pack_bf16x2(l0, l1);
                }
                if (p.dbg & 8) continue;
                const long long ridx = r_index(p.rsw, p.PRS, (cbase + c0) / 8 + g8, pp);
                p.rp[0][ridx] = make_uint4(oh4[0], oh4[1], oh4[2], oh4[3]);
                if (p.RHL == 2) p.rp[1][ridx] = make_uint4(ol4[0], ol4[1], ol4[2], ol4[3]);
              }
              const float cs = lane_col_sum<CH>(rr, lane);
              if (lane < CH) my_rs[c0 + lane] += cs;
            }
          }
          continue;
        }
        for (int c0 = 0; c0 < p.CT; c0 += CH) {
          ld_acc<CH>(ta + c0, p.stackF ? p.CT : 0, v);
          if (p.tr) {
            // (co, offset) columns: offsets 2j, 2j+1 are x-neighbours in the up-sampled grid -> 8-byte stores
#pragma unroll
            for (int i = 0; i < CH; i += 2) {
              const int cc = cbase + c0 + i, off = cc & 7;
              float2 o;
              o.x = __uint_as_float(v[i]) + s_bias[c0 + i];
              o.y = __uint_as_float(v[i + 1]) + s_bias[c0 + i + 1];
              if (valid)
                *reinterpret_cast<float2*>(ytb + (long long)(cc >> 3) * tS + ((long long)(off >> 2) * p.tH + ((off >> 1) & 1)) * p.tW) = o;
              // per-offset statistics: one softmax over the real channels for each of the 8 output voxels
              mx8[i & 7] = fmaxf(mx8[i & 7], o.x * p.kinv);
              mx8[(i + 1) & 7] = fmaxf(mx8[(i + 1) & 7], o.y * p.kinv);
              if (o.x > best8[i & 7]) { second8[i & 7] = best8[i & 7]; best8[i & 7] = o.x; bi8[i & 7] = cc >> 3; }
              else second8[i & 7] = fmaxf(second8[i & 7], o.x);
              if (o.y > best8[(i + 1) & 7]) { second8[(i + 1) & 7] = best8[(i + 1) & 7]; best8[(i + 1) & 7] = o.y; bi8[(i + 1) & 7] = cc >> 3; }
              else second8[(i + 1) & 7] = fmaxf(second8[(i + 1) & 7], o.y);
              amax8[i & 7] = fmaxf(amax8[i & 7], fabsf(o.x)); amax8[(i + 1) & 7] = fmaxf(amax8[(i + 1) & 7], fabsf(o.y));
            }
          } else {
            float t1[CH], t2[CH];
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const float f = fmaf(__uint_as_float(v[i]), s_inv[c0 + i], s_bias[c0 + i]);
              if (valid) yb[(long long)(c0 + i) * outS] = f;
              mx = fmaxf(mx, f * p.kinv);
              if (f > best) { second = best; best = f; bi = c0 + i; } else second = fmaxf(second, f);
              amax = fmaxf(amax, fabsf(f));
              t1[i] = valid ? f : 0.f; t2[i] = t1[i] * t1[i];
            }
            if (want_ys) {
              const float a1 = lane_col_sum<CH>(t1, lane), a2 = lane_col_sum<CH>(t2, lane);
              if (lane < CH) { atomicAdd(s_ys + c0 + lane, a1); atomicAdd(s_yq + c0 + lane, a2); }
            }
          }
        }
        if (p.fuse && p.tr) {
          // ---- transposed layer whose Cout*8 columns fit TMEM: grouped soft-WTA, group = column % 8 ----
          if (p.winner && valid) {
            int32_t* wb = p.winner + (long long)b * tS + ((long long)(2 * od) * p.tH + 2 * oh) * p.tW + 2 * ow;
#pragma unroll
            for (int off = 0; off < 8; ++off) {
              const long long wo = ((long long)(off >> 2) * p.tH + ((off >> 1) & 1)) * p.tW + (off & 1);
              wb[wo] = bi8[off];
              if (best8[off] - second8[off] <= p.tie_rel * amax8[off]) flag_tie(p.fix_list, p.fix_count, p.fix_cap, wb + wo - p.winner);
            }
          }
          if (p.write_r) {
            float sum8[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) sum8[i] = 0.f;
            for (int c0 = 0; c0 < p.CT; c0 += CH) {
              ld_acc<CH>(ta + c0, 0, v);
#pragma unroll
              for (int i = 0; i < CH; ++i)
                sum8[i & 7] += __expf(fmaf(__uint_as_float(v[i]) + s_bias[c0 + i], p.kinv, -mx8[i & 7]));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) sum8[i] = valid ? (1.f / sum8[i]) : 0.f;
            for (int c0 = 0; c0 < p.CT; c0 += CH) {
              ld_acc<CH>(ta + c0, 0, v);
              float rr[CH];
#pragma unroll
              for (int g8 = 0; g8 < CH / 8; ++g8) {
                uint32_t oh4[4], ol4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  __nv_bfloat16 h2[2], l2[2];
#pragma unroll
                  for (int k2 = 0; k2 < 2; ++k2) {
                    const int c = g8 * 8 + i * 2 + k2, off = i * 2 + k2;
                    const float r = __expf(fmaf(__uint_as_float(v[c]) + s_bias[c0 + c], p.kinv, -mx8[off])) * sum8[off];
                    split_bf16(r, h2[k2], l2[k2]);
                    rr[c] = __bfloat162float(h2[k2]) + (p.RHL == 2 ? __bfloat162float(l2[k2]) : 0.f);
                  }
                  oh4[i] = pack_bf16x2(h2[0], h2[1]);
                  ol4[i] = pack_bf16x2(l2[0], l2[1]);
                }
                const long long ridx = r_index(p.rsw, p.PRS, c0 / 8 + g8, pp);
                if (p.dbg & 8) continue;
                p.rp[0][ridx] = make_uint4(oh4[0], oh4[1], oh4[2], oh4[3]);
                if (p.RHL == 2) p.rp[1][ridx] = make_uint4(ol4[0], ol4[1], ol4[2], ol4[3]);
              }
              const float cs = lane_col_sum<CH>(rr, lane);
              if (lane < CH) my_rs[c0 + lane] += cs;
            }
          }
          continue;
        }
        if (p.fuse && p.winner && valid) {
          p.winner[(long long)b * outS + s] = bi;
          if (best - second <= p.tie_rel * amax) flag_tie(p.fix_list, p.fix_count, p.fix_cap, (long long)b * outS + s);
        }
        if (p.fuse && p.write_r) {
          float sum = 0.f;
          for (int c0 = 0; c0 < p.CT; c0 += CH) {
            ld_acc<CH>(ta + c0, p.stackF ? p.CT : 0, v);
#pragma unroll
            for (int i = 0; i < CH; ++i) {
              const float f = fmaf(__uint_as_float(v[i]), s_inv[c0 + i], s_bias[c0 + i]);
              sum += __expf(fmaf(f, p.kinv, -mx));
            }
          }
          const float rinv = valid ? (1.f / sum) : 0.f;
          for (int c0 = 0; c0 < p.CT; c0 += CH) {
            ld_acc<CH>(ta + c0, p.stackF ? p.CT : 0, v);
            float rr[CH];
#pragma unroll
            for (int g8 = 0; g8 < CH / 8; ++g8) {
              uint32_t oh4[4], ol4[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                __nv_bfloat16 h2[2], l2[2];
#pragma unroll
                for (int k2 = 0; k2 < 2; ++k2) {
                  const int c = g8 * 8 + i * 2 + k2;
                  const float f = fmaf(__uint_as_float(v[c]), s_inv[c0 + c], s_bias[c0 + c]);
                  const float r = __expf(fmaf(f, p.kinv, -mx)) * rinv;
                  split_bf16(r, h2[k2], l2[k2]);
                  rr[c] = __bfloat162float(h2[k2]) + (p.RHL == 2 ? __bfloat162float(l2[k2]) : 0.f);
                }
                oh4[i] = pack_bf16x2(h2[0], h2[1]);
                ol4[i] = pack_bf16x2(l2[0], l2[1]);
              }
              const long long ridx = r_index(p.rsw, p.PRS, c0 / 8 + g8, pp);
              if (pp < p.PR && !(p.dbg & 8)) {
                p.rp[0][ridx] = make_uint4(oh4[0], oh4[1], oh4[2], oh4[3]);
                if (p.RHL == 2) p.rp[1][ridx] = make_uint4(ol4[0], ol4[1], ol4[2], ol4[3]);
              }
            }
            const float cs = lane_col_sum<CH>(rr, lane);
            if (lane < CH) my_rs[c0 + lane] += cs;
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(t_empty + 8 * acc);
      aph ^= 1;                             // this set's buffer is reused every NACC-th item
    }
    __syncwarp();
    if constexpr (SP > 0) {                 // fold the per-thread column sums: one butterfly per CH columns
#pragma unroll
      for (int ck = 0; ck < SP; ++ck) {
        float t[CH];
#pragma unroll
        for (int i = 0; i < CH; ++i) t[i] = racc[ck * CH + i];
        const float cs = lane_col_sum<CH>(t, lane);
        if (lane < CH) my_rs[ck * CH + lane] += cs;
      }
      __syncwarp();
    }
    if (p.fuse && p.write_r && cur_ct >= 0)
      for (int c = lane; c < p.CT; c += 32) atomicAdd(p.rsum + cur_ct * p.CT + c, my_rs[c]);
    if (want_ys) {                          // host side guarantees n_ct == 1: columns are channels
      if constexpr (YACC) {
#pragma unroll
        for (int ck = 0; ck < SP; ++ck) {
          float t1[CH], t2[CH];
#pragma unroll
          for (int i = 0; i < CH; ++i) { t1[i] = ysacc[ck * CH + i]; t2[i] = yqacc[ck * CH + i]; }
          const float a1 = lane_col_sum<CH>(t1, lane), a2 = lane_col_sum<CH>(t2, lane);
          if (lane < CH) { atomicAdd(s_ys + ck * CH + lane, a1); atomicAdd(s_yq + ck * CH + lane, a2); }
        }
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");        // the 8 epilogue warps
      for (int c = (int)threadIdx.x - 64; c < p.CT; c += 256) {
        atomicAdd(p.ystats + 2 * c, (double)s_ys[c]);
        atomicAdd(p.ystats + 2 * c + 1, (double)s_yq[c]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// Unfused soft-WTA for layers whose channel count exceeds one CTA's TMEM (Cout > 512): reads the y rows back
// from global memory (L2-resident: a few thousand positions) and writes the packed responses.  Only the
// bottleneck layers of the 3-D network take this path.
struct SmxParams {
  const float* y; uint4* rp[2]; int32_t* winner; float* rsum;
  int Cout, RHL, WP, plane, Qimg, oD, oH, oW;
  long long PR, PTOT, PRS;
  float kinv;
  int* fix_list; int* fix_count; int fix_cap; float tie_rel;      // near-tie worklist (fixup.cu)
};

// A block takes 32 consecutive packed positions (the lanes: for one channel their y values are contiguous, so
// loads and the packed stores are coalesced) and its 8 warps share the channels in 8-channel chunks; maximum, winner
// and the exponential sum are combined across the warps through shared memory.  (The first version ran one thread
// per position over all channels: 28 blocks of 128 threads and 380 us for the 1024-channel layers of the 3-D net.)
__global__ void __launch_bounds__(256)
swta_softmax_pack_kernel(const __grid_constant__ SmxParams p) {
  __shared__ float s_mx[8][32], s_best[8][32], s_sum[8][32], s_sec[8][32], s_amax[8][32];
  __shared__ int s_bi[8][32];
  const int lane = threadIdx.x & 31, wg = threadIdx.x >> 5;
  const long long pp = (long long)blockIdx.x * 32 + lane;
  const long long outS = (long long)p.oD * p.oH * p.oW;
  const int oHW = p.oH * p.oW;
  const int b = (int)(pp / p.Qimg);
  int q = (int)(pp - (long long)b * p.Qimg);
  const int od = q / p.plane; q -= od * p.plane;
  const int oh = q / p.WP;
  const int ow = q - oh * p.WP;
  const bool valid = (pp < p.PTOT) && (od < p.oD) && (oh < p.oH) && (ow < p.oW);
  const long long s = (long long)od * oHW + (long long)oh * p.oW + ow;
  const float* yb = p.y + (long long)b * p.Cout * outS + s;
  const int C8 = p.Cout / 8;
  // pass 1: maximum of k*y and the winner (largest y, lowest index on ties) over this warp's chunks
  float mx = -INFINITY, best = -INFINITY, second = -INFINITY, amax = 0.f;
  int bi = 0x7fffffff;
  if (valid)
    for (int c8 = wg; c8 < C8; c8 += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int c = c8 * 8 + i;
        const float f = __ldg(yb + (long long)c * outS);
        mx = fmaxf(mx, f * p.kinv);
        if (f > best) { second = best; best = f; bi = c; } else second = fmaxf(second, f);
        amax = fmaxf(amax, fabsf(f));
      }
    }
  s_mx[wg][lane] = mx; s_best[wg][lane] = best; s_bi[wg][lane] = bi; s_sec[wg][lane] = second; s_amax[wg][lane] = amax;
  __syncthreads();
  best = -INFINITY; second = -INFINITY; bi = 0x7fffffff;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    mx = fmaxf(mx, s_mx[w][lane]);
    amax = fmaxf(amax, s_amax[w][lane]);
    const float ob = s_best[w][lane];
    const int oi = s_bi[w][lane];
    second = fmaxf(second, s_sec[w][lane]);
    if (ob > best || (ob == best && oi < bi)) { second = fmaxf(second, best); best = ob; bi = oi; }
    else second = fmaxf(second, ob);
  }
  if (wg == 0 && valid && p.winner) {
    p.winner[(long long)b * outS + s] = bi;
    if (best - second <= p.tie_rel * amax) flag_tie(p.fix_list, p.fix_count, p.fix_cap, (long long)b * outS + s);
  }
  // pass 2: sum of exponentials
  float sum = 0.f;
  if (valid)
    for (int c8 = wg; c8 < C8; c8 += 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += __expf(fmaf(__ldg(yb + (long long)(c8 * 8 + i) * outS), p.kinv, -mx));
    }
  s_sum[wg][lane] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) sum += s_sum[w][lane];
  const float rinv = valid ? 1.f / sum : 0.f;
  // pass 3: responses, packed stores (0 where the position is not an output pixel), per-channel sums
  for (int c8 = wg; c8 < C8; c8 += 8) {
    uint32_t oh4[4], ol4[4];
    float rr[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat16 h2[2], l2[2];
#pragma unroll
      for (int k2 = 0; k2 < 2; ++k2) {
        const int c = c8 * 8 + i * 2 + k2;
        const float r = valid ? __expf(fmaf(__ldg(yb + (long long)c * outS), p.kinv, -mx)) * rinv : 0.f;
        split_bf16(r, h2[k2], l2[k2]);
        rr[i * 2 + k2] = __bfloat162float(h2[k2]) + (p.RHL == 2 ? __bfloat162float(l2[k2]) : 0.f);
      }
      oh4[i] = pack_bf16x2(h2[0], h2[1]);
      ol4[i] = pack_bf16x2(l2[0], l2[1]);
    }
    if (pp < p.PR) {
      p.rp[0][(long long)c8 * p.PRS + pp] = make_uint4(oh4[0], oh4[1], oh4[2], oh4[3]);
      if (p.RHL == 2) p.rp[1][(long long)c8 * p.PRS + pp] = make_uint4(ol4[0], ol4[1], ol4[2], ol4[3]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = rr[i];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && v != 0.f) atomicAdd(p.rsum + c8 * 8 + i, v);
    }
  }
}

// Transposed conv (k == stride == 2, 3-D) with more than 512 packed channels: one WARP per INPUT voxel,
// lanes stride over the real output channels; each of its 8 output voxels (one per kernel offset) gets
// a softmax over the channels (hebb3d.py:276-289).  Packed channel index = co*8 + off, so one 16-byte
// vector holds the 8 offsets of one output channel.  These layers have few voxels and many channels.
__global__ void __launch_bounds__(256)
swta_softmax_pack_T_kernel(const __grid_constant__ SmxParams p, int tD, int tH, int tW, int CoutR) {
  const int lane = threadIdx.x & 31;
  const long long pp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pp >= p.PR) return;
  const long long tS = (long long)tD * tH * tW;
  const int b = (int)(pp / p.Qimg);
  int q = (int)(pp - (long long)b * p.Qimg);
  const int id = q / p.plane; q -= id * p.plane;
  const int ih = q / p.WP;
  const int iw = q - ih * p.WP;
  const bool valid = pp < p.PTOT;
  const float* yb = p.y + (long long)b * CoutR * tS + ((long long)(2 * id) * tH + 2 * ih) * tW + 2 * iw;
  long long o8[8];
#pragma unroll
  for (int off = 0; off < 8; ++off) o8[off] = ((long long)(off >> 2) * tH + ((off >> 1) & 1)) * tW + (off & 1);
  float mx[8], sum[8], best[8], sec[8], amx[8];
  int bi[8];
#pragma unroll
  for (int off = 0; off < 8; ++off) { mx[off] = -INFINITY; sum[off] = 0.f; best[off] = -INFINITY; bi[off] = 0x7fffffff; sec[off] = -INFINITY; amx[off] = 0.f; }
  if (valid) {
    for (int c = lane; c < CoutR; c += 32)
#pragma unroll
      for (int off = 0; off < 8; off += 2) {
        const float2 f2 = __ldg(reinterpret_cast<const float2*>(yb + (long long)c * tS + o8[off]));
        mx[off] = fmaxf(mx[off], f2.x * p.kinv); mx[off + 1] = fmaxf(mx[off + 1], f2.y * p.kinv);
        if (f2.x > best[off]) { sec[off] = best[off]; best[off] = f2.x; bi[off] = c; } else sec[off] = fmaxf(sec[off], f2.x);
        if (f2.y > best[off + 1]) { sec[off + 1] = best[off + 1]; best[off + 1] = f2.y; bi[off + 1] = c; } else sec[off + 1] = fmaxf(sec[off + 1], f2.y);
        amx[off] = fmaxf(amx[off], fabsf(f2.x)); amx[off + 1] = fmaxf(amx[off + 1], fabsf(f2.y));
      }
#pragma unroll
    for (int off = 0; off < 8; ++off) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        mx[off] = fmaxf(mx[off], __shfl_xor_sync(0xffffffffu, mx[off], o));
        amx[off] = fmaxf(amx[off], __shfl_xor_sync(0xffffffffu, amx[off], o));
        const float ob = __shfl_xor_sync(0xffffffffu, best[off], o);
        const float os = __shfl_xor_sync(0xffffffffu, sec[off], o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi[off], o);
        sec[off] = fmaxf(fmaxf(sec[off], os), fminf(best[off], ob));      // runner-up of the merged sets
        if (ob > best[off] || (ob == best[off] && oi < bi[off])) { best[off] = ob; bi[off] = oi; }   // lowest index wins ties
      }
    }
    for (int c = lane; c < CoutR; c += 32)
#pragma unroll
      for (int off = 0; off < 8; off += 2) {
        const float2 f2 = __ldg(reinterpret_cast<const float2*>(yb + (long long)c * tS + o8[off]));
        sum[off] += __expf(fmaf(f2.x, p.kinv, -mx[off]));
        sum[off + 1] += __expf(fmaf(f2.y, p.kinv, -mx[off + 1]));
      }
#pragma unroll
    for (int off = 0; off < 8; ++off) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum[off] += __shfl_xor_sync(0xffffffffu, sum[off], o);
      sum[off] = 1.f / sum[off];
    }
    if (p.winner && lane < 8) {
      int32_t* wb = p.winner + (long long)b * tS + ((long long)(2 * id) * tH + 2 * ih) * tW + 2 * iw;
      int sel = bi[0];
#pragma unroll
      for (int off = 1; off < 8; ++off) sel = (lane == off) ? bi[off] : sel;
      const long long wo = o8[0] + (((long long)(lane >> 2) * tH + ((lane >> 1) & 1)) * tW + (lane & 1));
      wb[wo] = sel;
      float bsel = best[0], ssel = sec[0], asel = amx[0];
#pragma unroll
      for (int off = 1; off < 8; ++off) { bsel = (lane == off) ? best[off] : bsel; ssel = (lane == off) ? sec[off] : ssel; asel = (lane == off) ? amx[off] : asel; }
      if (bsel - ssel <= p.tie_rel * asel) flag_tie(p.fix_list, p.fix_count, p.fix_cap, wb + wo - p.winner);
    }
  }
  for (int c = lane; c < CoutR; c += 32) {
    uint32_t oh4[4], ol4[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f2 = make_float2(0.f, 0.f);
      if (valid) f2 = __ldg(reinterpret_cast<const float2*>(yb + (long long)c * tS + o8[2 * i]));
      __nv_bfloat16 h2[2], l2[2];
      const float r0 = valid ? __expf(fmaf(f2.x, p.kinv, -mx[2 * i])) * sum[2 * i] : 0.f;
      const float r1 = valid ? __expf(fmaf(f2.y, p.kinv, -mx[2 * i + 1])) * sum[2 * i + 1] : 0.f;
      split_bf16(r0, h2[0], l2[0]);
      split_bf16(r1, h2[1], l2[1]);
      oh4[i] = pack_bf16x2(h2[0], h2[1]);
      ol4[i] = pack_bf16x2(l2[0], l2[1]);
    }
    p.rp[0][(long long)c * p.PRS + pp] = make_uint4(oh4[0], oh4[1], oh4[2], oh4[3]);
    if (p.RHL == 2) p.rp[1][(long long)c * p.PRS + pp] = make_uint4(ol4[0], ol4[1], ol4[2], ol4[3]);
  }
}

// rsum[c8*8 + i] = sum_p (hi + lo)(Rp[c8][p][i]) : one block per 8-channel chunk, fixed summation order.
__global__ void __launch_bounds__(256)
rsum_from_packed_kernel(const uint4* __restrict__ rhi, const uint4* __restrict__ rlo, float* __restrict__ rsum, long long PR,
                        long long PRS) {
  const int c8 = blockIdx.x;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
  for (long long pp = threadIdx.x; pp < PR; pp += blockDim.x) {
    const uint4 h = __ldg(rhi + (long long)c8 * PRS + pp);
    const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      acc[2 * i] += __uint_as_float(hw[i] << 16);
      acc[2 * i + 1] += __uint_as_float(hw[i] & 0xffff0000u);
    }
    if (rlo) {
      const uint4 l = __ldg(rlo + (long long)c8 * PRS + pp);
      const uint32_t lw[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[2 * i] += __uint_as_float(lw[i] << 16);
        acc[2 * i + 1] += __uint_as_float(lw[i] & 0xffff0000u);
      }
    }
  }
  __shared__ float part[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][i] = acc[i];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += part[w][threadIdx.x];
    rsum[c8 * 8 + threadIdx.x] = t;
  }
}

// delta_w[co][ci][off] += sum_s Hpart[s][0][ci][co*8+off] - sum_off' rsum[co*8+off'] * W[co][ci][off']
__global__ void __launch_bounds__(256)
tc_finalize_T_kernel(const float* __restrict__ hpart, const float* __restrict__ rsum, const float* __restrict__ W,
                     float* __restrict__ dw, int PS, int Cin, int CinP, int CoutR, int tr_q) {
  // 256 threads = 8 split lanes x 32 consecutive (ci, co) outputs: the position splits of one output are summed
  // by 8 threads in a fixed order (deterministic), so layers with few weights and many splits stay parallel
  __shared__ float red[8][32][9];
  const long long n = (long long)Cin * CoutR;
  const long long Cp = (long long)CoutR * 8;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (long long base = (long long)blockIdx.x * 32; base < n; base += (long long)gridDim.x * 32) {
    const long long idx = base + tx;
    const bool ok = idx < n;
    const int co = ok ? (int)(idx % CoutR) : 0;
    const int ci = ok ? (int)(idx / CoutR) : 0;
    int col[8];      // packed column of (co, off): co*8 + off, or (off>>1)*tr_q + 2*co + (off&1)
#pragma unroll
    for (int off = 0; off < 8; ++off) col[off] = tr_q ? ((off >> 1) * tr_q + 2 * co + (off & 1)) : (co * 8 + off);
    float h[8];
#pragma unroll
    for (int off = 0; off < 8; ++off) h[off] = 0.f;
    if (ok)
      for (int s = ty; s < PS; s += 8) {
        const float* hp = hpart + ((long long)s * CinP + ci) * Cp;
#pragma unroll
        for (int off = 0; off < 8; ++off) h[off] += hp[col[off]];
      }
#pragma unroll
    for (int off = 0; off < 8; ++off) red[ty][tx][off] = h[off];
    __syncthreads();
    if (ty == 0 && ok) {
      const float* w = W + ((long long)co * Cin + ci) * 8;
      float* d = dw + ((long long)co * Cin + ci) * 8;
      float dec = 0.f;
#pragma unroll
      for (int off = 0; off < 8; ++off) dec += rsum[col[off]] * w[off];
#pragma unroll
      for (int off = 0; off < 8; ++off) {
        float tot = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot += red[k][tx][off];
        d[off] += tot - dec;
      }
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------
// dW shift-GEMM:  Hpart[split][tap][ci][co] = sum_q Xp[ci][q] * Rp[co][q - shift(tap)]
//
// The A operand (x, the M rows) is the SAME tile for every tap of a position block; the tap is an address offset
// of the B operand (the responses), whose packed planes carry zero pads of `maxshift` positions on both sides.
// Shared-memory bandwidth bounds this kernel for Cout <= 64 (an M=128 A tile is 4 KB per 16 positions against
// 32 cycles of math at N=64: profiles/README.md), so keeping A fixed lets a run of instructions fetch it once
// through the A-operand collector (fill / use / lastuse) -- measured 37-42 cycles per MMA in runs of >= 6
// against 59 with one A fetch per instruction (profiles/umma_rate3_r1.txt).
// -------------------------------------------------------------------------------------
struct DwParams {
  const uint4* xp[2];
  const uint4* rp[2];
  float* hpart; int* err;
  int Cin, Cout, CC, C8, taps, HL;
  long long PA, PRS;            // positions between the 8-channel planes of the x / r operands
  int BLK, SEGLEN, total_blocks, blocks_per_split, PS;     // SEGLEN: staged r positions = BLK + tap halo
  int ngrp; int grp_base[9];
  // "super-taps": with nrep > 1 the staged x tile is replicated nrep times, copy r shifted by r positions, so the
  // M rows of ONE tcgen05.mma cover nrep horizontally adjacent taps (kw0 .. kw0+nrep-1) of a kernel row.
  // st_boff: where the B operand of a super-tap starts inside the staged r tile (positions).
  int nrep; int grp_st_begin[10]; int st_boff[kMaxTaps]; int st_first[kMaxTaps]; int st_n[kMaxTaps];
  int rhalo;                    // the staged r tile starts rhalo + grp_base positions before the x tile
  // replicas of the x tile are rep_stride positions apart and tap_rep taps apart (1 / 1: kw neighbours; rsw: WP / kW)
  int rep_stride, tap_rep; int st_aoff[kMaxTaps];           // st_aoff: A start of a super-tap inside the x region (16 B units)
  // rsw: responses staged as the 64-channel swizzled image; ONE instruction reads ncopy position-shifted copies of
  // the tile (N = 64 * ncopy: the kW taps of a kernel row); b_rows staged rows, fetched from an 8-aligned start
  int rsw, ncopy, b_rows, r_lead;
  // tap groups may carry different numbers of taps: group g is split grp_w[g] * (grid / virtual tiles) ways over the
  // positions, so that every CTA issues the same number of instructions.  Virtual tile v = (vt_grp[v], vt_w[v]).
  int n_vt, vt_grp[2 * 9], vt_w[2 * 9], grp_w[9], blocks_per_split2;       // blocks_per_split2: for groups of weight 2
  int reuse;                    // 1: k-step outer / tap inner with A collector re-use; 0: tap outer, one A fetch per MMA
  int CM, n_cin_tiles, CN, n_cout_tiles, ST, CinP;
  int stackM, stackN, cpt;      // bf16x3 "precision stacking": [x_hi; x_lo] along M and/or [r_hi | r_lo] along N, so one
                                // tcgen05.mma yields several of the hi/lo partial products; cpt = 8-channel chunks per cin tile
  uint32_t stage_bytes, x_bytes, off_bar, tmem_cols;
};

constexpr int kMaxRun = 9;       // super-taps per tap group the collector loop order handles (3 x 3 taps of a plane)

// One pass of a k-step: the n super-taps of the CTA's tap group against ONE A tile, NB B tiles per super-tap
// (bks + bo[j] + bx0, then + bx1).  With n > 1 the run shares A through the collector (fill, then use).  The B
// offsets live in registers and the loop is fully unrolled: a load from the parameter block between two
// `asm volatile` issues would sit on the issuing thread's critical path.  ACC = 0: the first instruction into each
// accumulator overwrites it.
template <int NB, int ACC>
__device__ __forceinline__ void dw_pass(int n, uint32_t d0, int colw, uint32_t ah, uint32_t a_hi32, uint32_t bks,
                                        uint32_t b_hi32, uint32_t idesc, uint32_t bx0, uint32_t bx1,
                                        const uint32_t (&bo)[kMaxRun]) {
  if (n == 1) {
    const uint32_t bh = bks + bo[0];
    if (NB == 1) umma_lo<0, ACC>(d0, ah, a_hi32, bh + bx0, b_hi32, idesc);
    else { umma_lo<1, ACC>(d0, ah, a_hi32, bh + bx0, b_hi32, idesc); umma_lo<3, 1>(d0, ah, a_hi32, bh + bx1, b_hi32, idesc); }
    return;
  }
#pragma unroll
  for (int j = 0; j < kMaxRun; ++j) {
    if (j < n) {
      const uint32_t bh = bks + bo[j];
      const uint32_t d = d0 + j * colw;
      if (j == 0) umma_lo<1, ACC>(d, ah, a_hi32, bh + bx0, b_hi32, idesc);
      else umma_lo<2, ACC>(d, ah, a_hi32, bh + bx0, b_hi32, idesc);
      if (NB == 2) umma_lo<2, 1>(d, ah, a_hi32, bh + bx1, b_hi32, idesc);
    }
  }
}

// All MMAs of one k-step (16 positions).  MODE: 1 one instruction per super-tap (bf16, or both operands
// precision-stacked), 2 rows stacked ([x_hi; x_lo] against r_lo, r_hi), 3 columns stacked (x_lo, x_hi against
// [r_hi | r_lo]), 4 classic 3-product split.
template <int MODE, int ACC>
__device__ __forceinline__ void dw_kstep(int n, uint32_t d, int colw, uint32_t ah, uint32_t a_hi32, uint32_t bks,
                                         uint32_t b_hi32, uint32_t idesc, uint32_t xhl16, uint32_t rhl16,
                                         const uint32_t (&bo)[kMaxRun]) {
  if (MODE == 2) {
    dw_pass<2, ACC>(n, d, colw, ah, a_hi32, bks, b_hi32, idesc, rhl16, 0u, bo);
  } else if (MODE == 3) {
    dw_pass<1, ACC>(n, d, colw, ah + xhl16, a_hi32, bks, b_hi32, idesc, 0u, 0u, bo);
    dw_pass<1, 1>(n, d, colw, ah, a_hi32, bks, b_hi32, idesc, 0u, 0u, bo);
  } else if (MODE == 4) {
    dw_pass<2, ACC>(n, d, colw, ah, a_hi32, bks, b_hi32, idesc, rhl16, 0u, bo);
    dw_pass<1, 1>(n, d, colw, ah + xhl16, a_hi32, bks, b_hi32, idesc, 0u, 0u, bo);
  } else {
    dw_pass<1, ACC>(n, d, colw, ah, a_hi32, bks, b_hi32, idesc, 0u, 0u, bo);
  }
}

// All MMAs of one staged position block (called by the elected lane of the MMA warp).  ACC0 = 0 on the first
// block of the split (overwrite the accumulators).  REUSE = 1: k-step outer, the taps of the group share the A
// tile of a k-step through the collector; REUSE = 0 (small tap groups, where the collector hand-over costs more
// than it saves): tap outer, BLK/16 k-steps inner, one A fetch per instruction.
template <int MODE, int ACC0, int REUSE>
__device__ __forceinline__ void dw_issue(const DwParams& p, uint32_t xa, uint32_t ra, int st_b, int st_e, uint32_t tmem_base,
                                         int colw, int ksteps, uint32_t a_lbo, uint32_t a_hi32, uint32_t b_lbo,
                                         uint32_t b_hi32, uint32_t idesc, uint32_t xhl16, uint32_t rhl16, uint32_t bstep) {
  uint32_t ah = a_lbo | (xa >> 4);
  uint32_t bks = b_lbo | (ra >> 4);
  uint32_t bo[kMaxRun];
  if (REUSE) {
    const int n = st_e - st_b;
#pragma unroll
    for (int j = 0; j < kMaxRun; ++j) bo[j] = (j < n) ? (uint32_t)p.st_boff[st_b + j] : 0u;
    dw_kstep<MODE, ACC0>(n, tmem_base, colw, ah, a_hi32, bks, b_hi32, idesc, xhl16, rhl16, bo);
    for (int ks = 1; ks < ksteps; ++ks) {
      ah += 16u; bks += bstep;                   // next 16 positions: 256 bytes (2048 in the swizzled response image)
      dw_kstep<MODE, 1>(n, tmem_base, colw, ah, a_hi32, bks, b_hi32, idesc, xhl16, rhl16, bo);
    }
  } else {
#pragma unroll
    for (int j = 0; j < kMaxRun; ++j) bo[j] = 0u;
    uint32_t d = tmem_base;
    for (int stp = st_b; stp < st_e; ++stp, d += colw) {
      uint32_t a = ah + (uint32_t)p.st_aoff[stp], b = bks + (uint32_t)p.st_boff[stp];
      dw_kstep<MODE, ACC0>(1, d, colw, a, a_hi32, b, b_hi32, idesc, xhl16, rhl16, bo);
      for (int ks = 1; ks < ksteps; ++ks) {
        a += 16u; b += bstep;
        dw_kstep<MODE, 1>(1, d, colw, a, a_hi32, b, b_hi32, idesc, xhl16, rhl16, bo);
      }
    }
  }
}

__global__ void __launch_bounds__(192, 1)
dw_swta_kernel(const __grid_constant__ DwParams p) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  // stages start on 1024-byte boundaries of the shared-memory window: the swizzled response image repeats every 1024 B
  uint8_t* const smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform
  const int lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.off_bar);
  const uint32_t full = smem_u32(bars), empty = full + 8 * p.ST, done = empty + 8 * p.ST;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * p.ST + 1);

  // output tile fastest: CTAs that stream the same positions run side by side (L2 reuse)
  int task = blockIdx.x;
  const int out_tiles = p.n_vt * p.n_cin_tiles * p.n_cout_tiles;
  const int usplit = task / out_tiles; task -= usplit * out_tiles;
  const int cout_tile = task % p.n_cout_tiles; task /= p.n_cout_tiles;
  const int cin_tile = task % p.n_cin_tiles; task /= p.n_cin_tiles;
  const int grp = p.vt_grp[task];
  const int split = usplit * p.grp_w[grp] + p.vt_w[task];
  const int bps = p.grp_w[grp] == 2 ? p.blocks_per_split2 : p.blocks_per_split;
  const int cm_chunks = min(p.cpt, p.CC - cin_tile * p.cpt);
  const int N = min(p.CN, p.Cout - cout_tile * p.CN);
  const int rn_chunks = N / 8;
  const int Neff = p.rsw ? p.ncopy * N : (p.stackN ? 2 * N : N);
  const int colw = p.rsw ? p.ncopy * p.CN : (p.stackN ? 2 : 1) * p.CN;        // TMEM columns per super-tap
  const int st_b = p.grp_st_begin[grp], st_e = p.grp_st_begin[grp + 1];
  const int blk_b = min(split * bps, p.total_blocks);
  const int blk_e = min(blk_b + bps, p.total_blocks);
  // A weight-1 group next to weight-2 groups covers the position range of TWO of their splits: it walks the two
  // halves interleaved, so that at any time all CTAs of a unit split stream the same two windows (L2 re-use).
  const int nblk = blk_e - blk_b;
  const int ihalf = (p.blocks_per_split2 > 0 && p.grp_w[grp] == 1) ? p.blocks_per_split2 : nblk;
  const int ih2 = nblk > ihalf ? nblk - ihalf : 0;       // blocks in the second half
#define HEBB_DW_BLK(i) (blk_b + ((i) < 2 * ih2 ? ((i) >> 1) + ((i) & 1) * ihalf : (i) - ih2))

  if (threadIdx.x == 0) {
    for (int i = 0; i < p.ST; ++i) { mbar_init(full + 8 * i, 1); mbar_init(empty + 8 * i, 1); }
    mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(s_tmem), p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t r_off = p.x_bytes;                 // R region follows X region inside a stage
  const uint32_t x_hl_stride = cm_chunks * p.BLK * 16;      // lo chunks directly follow the hi chunks
  const uint32_t x_rep_stride = p.HL * x_hl_stride;         // replica r+1 directly follows replica r
  const uint32_t r_hl_stride = p.rsw ? (uint32_t)(p.b_rows * 128) : rn_chunks * p.SEGLEN * 16;

  if (warp == 0) {
    if (elect_one()) {
      int st = 0; uint32_t ph = 0;
      const uint32_t bytes = p.rsw ? (uint32_t)(p.HL * (p.nrep * cm_chunks * p.BLK * 16 + p.b_rows * 128))
                                   : p.HL * (p.nrep * cm_chunks * p.BLK + rn_chunks * p.SEGLEN) * 16;
      const long long r_back = (long long)p.grp_base[grp] + p.rhalo;      // the r tile starts this far before the x tile
      for (int i = 0; i < nblk; ++i) {
        const long long q0 = (long long)HEBB_DW_BLK(i) * p.BLK;
        mbar_wait(empty + 8 * st, ph ^ 1, p.err, 11);
        mbar_expect_tx(full + 8 * st, bytes);
        const uint32_t dst = sbase + st * p.stage_bytes;
        for (int hl = 0; hl < p.HL; ++hl) {
          for (int rep = 0; rep < p.nrep; ++rep)
            for (int c = 0; c < cm_chunks; ++c)
              bulk_g2s(dst + rep * x_rep_stride + hl * x_hl_stride + c * p.BLK * 16,
                       p.xp[hl] + (long long)(cin_tile * p.cpt + c) * p.PA + q0 + (long long)rep * p.rep_stride,
                       p.BLK * 16, full + 8 * st);
          if (p.rsw) {
            // one copy: b_rows whole positions (128 B each) of this 64-channel plane, from an 8-aligned row so the
            // staged image keeps the swizzle phase of the global one (r_lead and q0 are multiples of 8)
            const long long r0a = ((q0 - r_back + p.r_lead) & ~7LL) - p.r_lead;
            bulk_g2s(dst + r_off + hl * (uint32_t)(p.b_rows * 128), p.rp[hl] + ((long long)cout_tile * p.PRS + r0a) * 8,
                     p.b_rows * 128, full + 8 * st);
          } else {
            for (int c = 0; c < rn_chunks; ++c)
              bulk_g2s(dst + r_off + hl * r_hl_stride + c * p.SEGLEN * 16,
                       p.rp[hl] + (long long)(cout_tile * (p.CN / 8) + c) * p.PRS + q0 - r_back, p.SEGLEN * 16,
                       full + 8 * st);
          }
        }
        if (++st == p.ST) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {
      const uint32_t idesc = idesc_bf16(p.CM, Neff, 1, 1);
      const uint64_t a_hi64 = smem_desc_hi(128, p.BLK * 16);      // MN-major: LBO = next 8 positions, SBO = next chunk
      // rsw: MN-major SWIZZLE_128B (layout type 2), 8-position groups 1024 B apart, 64-channel atoms ONE position apart
      const uint64_t b_hi64 = p.rsw ? (smem_desc_hi(128, 1024) | (2ull << 61)) : smem_desc_hi(128, p.SEGLEN * 16);
      const uint32_t bstep = p.rsw ? 128u : 16u;
      // first staged row = (q0 - r_back) rounded down to a multiple of 8: the B operand starts r_rem rows further
      const uint32_t r_rem = p.rsw ? (uint32_t)(((long long)blk_b * p.BLK - p.grp_base[grp] - p.rhalo + p.r_lead) & 7LL) : 0u;
      const uint32_t a_lbo = (uint32_t)a_hi64, a_hi32 = (uint32_t)(a_hi64 >> 32);
      const uint32_t b_lbo = (uint32_t)b_hi64, b_hi32 = (uint32_t)(b_hi64 >> 32);
      const uint32_t xhl16 = x_hl_stride >> 4, rhl16 = r_hl_stride >> 4;
      int st = 0; uint32_t ph = 0;
      const int ksteps = p.BLK / 16;
      const int mode = (p.HL == 2) ? (p.stackM ? (p.stackN ? 1 : 2) : (p.stackN ? 3 : 4)) : 1;
      const int sel = (mode - 1) * 2 + (p.reuse ? 1 : 0);
      for (int i = 0; i < nblk; ++i) {
        mbar_wait(full + 8 * st, ph, p.err, 12);
        tc_fence_after();
        const uint32_t xa = sbase + st * p.stage_bytes;
        const uint32_t ra = xa + r_off + r_rem * 128u;
        if (elect_one()) {
          const bool first = (i == 0);
#define HEBB_DW_ISSUE(M, R)                                                                                               \
          (first ? dw_issue<M, 0, R>(p, xa, ra, st_b, st_e, tmem_base, colw, ksteps, a_lbo, a_hi32, b_lbo, b_hi32, idesc,  \
                                     xhl16, rhl16, bstep)                                                                 \
                 : dw_issue<M, 1, R>(p, xa, ra, st_b, st_e, tmem_base, colw, ksteps, a_lbo, a_hi32, b_lbo, b_hi32, idesc,  \
                                     xhl16, rhl16, bstep))
          switch (sel) {
            case 0: HEBB_DW_ISSUE(1, 0); break;
            case 1: HEBB_DW_ISSUE(1, 1); break;
            case 2: HEBB_DW_ISSUE(2, 0); break;
            case 3: HEBB_DW_ISSUE(2, 1); break;
            case 4: HEBB_DW_ISSUE(3, 0); break;
            case 5: HEBB_DW_ISSUE(3, 1); break;
            case 6: HEBB_DW_ISSUE(4, 0); break;
            default: HEBB_DW_ISSUE(4, 1); break;
          }
#undef HEBB_DW_ISSUE
          umma_commit(empty + 8 * st);
        }
        __syncwarp();
        if (++st == p.ST) { st = 0; ph ^= 1; }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    mbar_wait(done, 0, p.err, 13);
    tc_fence_after();
    int row; bool row_ok;
    if (p.CM == 128) { row = quad * 32 + lane; row_ok = true; }
    else { row = quad * 16 + (lane & 15); row_ok = lane < 16; }    // M=64: D row r lives in lane (r%16)+32*(r/16)
    // rows of one replica: [0, R) pair with x_hi and, when stacked, [R, 2R) with x_lo (R = real channels of
    // this cin tile); replica `rep` (rows rep*RR ...) belongs to tap st_first + rep of the super-tap
    const int R = cm_chunks * 8;
    const int RR = (p.stackM ? 2 : 1) * R;
    const int rep = row / RR;
    const int rrow = row - rep * RR;
    const int row_lo = (p.stackM && rrow >= R) ? 1 : 0;
    const int rloc = rrow - row_lo * R;
    const int ci = cin_tile * p.cpt * 8 + rloc;
    const bool st_ok = row_ok && rep < p.nrep && ci < p.Cin;
    const bool have = blk_e > blk_b;
    const int Q = (p.stackM ? 2 : 1) * (p.stackN ? 2 : 1);
    for (int stp = st_b; stp < st_e; ++stp) {
      const uint32_t ta = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + (stp - st_b) * colw;
      const int tap0 = p.st_first[stp] + rep * p.tap_rep;
      const bool tap_ok = st_ok && rep < p.st_n[stp];
      for (int c0 = 0; c0 < Neff; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(ta + c0, v);
        tmem_ld_wait();
        // rsw: column block j holds the tile shifted by +j positions = tap kw = ncopy-1-j of the kernel row
        const int jcopy = p.rsw ? c0 / N : 0;
        const int tap = tap0 + (p.rsw ? p.ncopy - 1 - jcopy : 0);
        const int col_lo = (!p.rsw && p.stackN && c0 >= N) ? 1 : 0;
        const int quadrant = row_lo * (p.stackN ? 2 : 1) + col_lo;
        float* dst = p.hpart + ((((long long)split * Q + quadrant) * p.taps + tap) * p.CinP + ci) * p.Cout +
                     cout_tile * p.CN + (c0 - (p.rsw ? jcopy : col_lo) * N);
        if (tap_ok) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 o;
            o.x = have ? __uint_as_float(v[4 * i + 0]) : 0.f; o.y = have ? __uint_as_float(v[4 * i + 1]) : 0.f;
            o.z = have ? __uint_as_float(v[4 * i + 2]) : 0.f; o.w = have ? __uint_as_float(v[4 * i + 3]) : 0.f;
            *reinterpret_cast<float4*>(dst + 4 * i) = o;
          }
        }
      }
    }
  }
#undef HEBB_DW_BLK
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}

// delta_w[co][ci][t] += sum_s Hpart[s][t][ci][co] - rsum[co] * W[co][ci][t]
// 256 threads = 8 split lanes x 32 consecutive outputs (co fastest -> coalesced partial reads); the
// partials of one output are summed in a fixed order, so the result is deterministic.
__global__ void __launch_bounds__(256)
tc_finalize_kernel(const float* __restrict__ hpart, const float* __restrict__ rsum, const float* __restrict__ W,
                   float* __restrict__ dw, int PS, int taps, int Cin, int CinP, int Cout) {
  __shared__ float red[8][33];
  const long long n = (long long)taps * Cin * Cout;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long plane = (long long)taps * CinP * Cout;
  for (long long base = (long long)blockIdx.x * 32; base < n; base += (long long)gridDim.x * 32) {
    const long long idx = base + tx;
    float acc = 0.f;
    int co = 0, ci = 0, t = 0;
    if (idx < n) {
      co = (int)(idx % Cout);
      const long long t2 = idx / Cout;
      ci = (int)(t2 % Cin);
      t = (int)(t2 / Cin);
      const float* hp = hpart + ((long long)t * CinP + ci) * Cout + co;
      for (int s = ty; s < PS; s += 8) acc += hp[(long long)s * plane];
    }
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && idx < n) {
      float tot = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += red[k][tx];
      const long long wi = ((long long)co * Cin + ci) * taps + t;
      dw[wi] += rsum ? (tot - rsum[co] * W[wi]) : tot;
    }
    __syncthreads();
  }
}

// Same result for LARGE weight tensors (few position splits, millions of outputs).  It is a transposition:
// the partials are contiguous along co, delta_w / W along j = ci*taps + t (a row of the [Cout][Cin*taps] weight).
// A T(j) x T(co) tile goes through shared memory; for the largest tensors T = 128, so that BOTH sides move in
// 512-byte runs -- with 32 x 32 tiles every 128-byte piece fell into a different DRAM page and the pass ran at a
// quarter of the HBM rate (1024 -> 1024 3x3x3: 283 us for 452 MB).
template <int kFinT>
__global__ void __launch_bounds__(256)
tc_finalize_tiled_kernel(const float* __restrict__ hpart, const float* __restrict__ rsum, const float* __restrict__ W,
                         float* __restrict__ dw, int PS, int taps, int Cin, int CinP, int Cout) {
  extern __shared__ float fin_tile[];                 // [kFinT j][kFinT + 1 co]
  constexpr int V = kFinT / 32;                       // consecutive co per lane: one 4/8/16-byte load
  constexpr int RU = 4;                               // rows in flight per warp (memory-level parallelism)
  const int K = Cin * taps;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long plane = (long long)taps * CinP * Cout;
  const int j0 = blockIdx.x * kFinT, co0 = blockIdx.y * kFinT;
  const bool co_ok = co0 + V * lane < Cout;           // Cout is a multiple of 16 >= V: a lane is all in or all out
  for (int r0 = warp * RU; r0 < kFinT; r0 += 8 * RU) {
    float acc[RU][V];
    const float* hp[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
#pragma unroll
      for (int k = 0; k < V; ++k) acc[u][k] = 0.f;
      const int j = j0 + r0 + u;
      hp[u] = nullptr;
      if (j < K && co_ok) {
        const int ci = j / taps, t = j - ci * taps;
        hp[u] = hpart + ((long long)t * CinP + ci) * Cout + co0 + V * lane;
      }
    }
    for (int sp = 0; sp < PS; ++sp) {
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        if (hp[u]) {
          if (V == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(hp[u])); acc[u][0] += v.x; acc[u][1 % V] += v.y; acc[u][2 % V] += v.z; acc[u][3 % V] += v.w; }
          else if (V == 2) { const float2 v = __ldg(reinterpret_cast<const float2*>(hp[u])); acc[u][0] += v.x; acc[u][1 % V] += v.y; }
          else acc[u][0] += __ldg(hp[u]);
          hp[u] += plane;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u)
#pragma unroll
      for (int k = 0; k < V; ++k) fin_tile[(r0 + u) * (kFinT + 1) + V * lane + k] = acc[u][k];
  }
  __syncthreads();
#pragma unroll 2
  for (int c = warp; c < kFinT; c += 8) {
    const int co = co0 + c;
    if (co >= Cout) break;
    const float rs = rsum ? rsum[co] : 0.f;
    const long long row = (long long)co * K + j0;
    float wv[V], dv[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int jj = lane + 32 * i;
      const bool ok = j0 + jj < K;
      dv[i] = ok ? dw[row + jj] : 0.f;
      wv[i] = (ok && rsum) ? __ldg(W + row + jj) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int jj = lane + 32 * i;
      if (j0 + jj < K) dw[row + jj] = dv[i] + (fin_tile[jj * (kFinT + 1) + c] - rs * wv[i]);
    }
  }
}

// -------------------------------------------------------------------------------------
// Host-side planning
// -------------------------------------------------------------------------------------
static int round_up_i(long long v, int a) { return (int)((v + a - 1) / a * a); }
__host__ __device__ static inline uint32_t pow2_cols(int c) { uint32_t r = 32; while ((int)r < c) r <<= 1; return r; }

struct Plan {
  // geometry of the packed position space
  int HP, WP, plane, Qimg, CC, NSLAB, C8;
  long long PTOT, PR, PA;
  int maxshift;
  // forward
  int f_HL, MB, TILE_M, f_SEGLEN, XST, WST, NACC, WG, f_nseg, f_ntiles, CT, n_ct, stackF;
  uint32_t f_x_stage, f_w_stage, f_off_w, f_off_misc, f_smem, f_tmem;
  // dW
  int d_HL, BLK, d_SEGLEN, d_by_kh, ngrp, CM, n_cin_tiles, CN, n_cout_tiles, PS, total_blocks, blocks_per_split, ST, CinP;
  int stackM, stackN, cpt, Q, nrep, rhalo, reuse;
  long long PRS; int r_lead;      // Rp plane: r_lead zero positions, PR packed responses, zero tail (PRS in all)
  // swizzled-response variant of the dW kernel (bf16 operands, 64 response channels, 3-wide kernel rows): its own
  // tiling, chosen per call (the weight-gradient and HPCA calls keep the plain layout); 0 = not available
  int rsw, rs_BLK, rs_ST, rs_nrep, rs_by_kh, rs_cpt, rs_n_cin, rs_CinP, rs_PS, rs_total_blocks, rs_blocks_per_split, rs_b_rows;
  int rs_pair, rs_rhalo, rs_PSu, rs_bps2;     // rs_pair: kernel rows grouped (kh0, kh1) + (kh2), weights 2 : 1
  int rs_stackM;                               // split precision, Cin = 64: M = [x_hi; x_lo] (two partial planes per split)
  uint32_t rs_stage, rs_x_bytes, rs_off_bar, rs_smem, rs_tmem;
  uint32_t d_stage, d_x_bytes, d_off_bar, d_smem, d_tmem;
  // workspace carve (byte offsets)
  size_t o_inv, o_rsum, o_err, o_xp[2], o_rp[2], o_wp, o_hpart, o_gram, o_hpart2, o_fix, total;
  int fix_cap;                    // entries of the near-tie worklist (fixup.cu); its counter sits 16 bytes after the error word
  bool ok, gram_ok;
};

// A transposed conv with kernel == stride == 2 (3-D) and no padding is a 1x1 conv onto Cout*8
// "(co, offset)" channels followed by a pixel shuffle; its plasticity update is the matching 1x1 update.
static bool equivalent_1x1(const Geo& g, Geo* e) {
  if (!g.transposed) {
    *e = g;
    // Few input channels (first layers: 1 or 3): a 16-channel K slab would be mostly zero padding and every
    // tap would cost a full MMA.  Gather the patch at pack time instead — (ci, tap) pairs become Cin*taps
    // pseudo-channels of a 1x1 layer over the OUTPUT grid; [Cout][Cin][taps] is already that layer's weight.
    if (g.Cin <= 4 && g.taps > 1 && g.sD == 1 && g.sH == 1 && g.sW == 1 && g.Cin * g.taps <= 128) {
      e->Cin = g.Cin * g.taps;
      e->kD = e->kH = e->kW = 1;
      e->pD = e->pH = e->pW = e->qD = e->qH = e->qW = 0;
      e->iD = g.oD; e->iH = g.oH; e->iW = g.oW;
      e->taps = 1; e->K = e->Cin;
      e->inS = g.outS;
    }
    return true;
  }
  if (g.nd != 3 || g.kD != 2 || g.kH != 2 || g.kW != 2 || g.sD != 2 || g.sH != 2 || g.sW != 2) return false;
  if (g.pD || g.pH || g.pW || g.qD || g.qH || g.qW) return false;
  if ((long long)g.Cout * 8 > 8192) return false;
  *e = g;
  e->transposed = 0;
  e->Cout = g.Cout * 8;
  e->kD = e->kH = e->kW = 1; e->sD = e->sH = e->sW = 1;
  e->oD = g.iD; e->oH = g.iH; e->oW = g.iW;
  e->taps = 1; e->K = g.Cin;
  e->outS = g.inS;
  return true;
}

// Transposed layers whose 2*Cout columns (one (od2, oh2) block: all channels x both x-parities) fit a channel
// tile use the block-major column order, which keeps the soft-WTA fused however many tiles the layer needs.
static int tr_quantum(const Geo& g0) {
  if (!g0.transposed) return 0;
  const int q = 2 * g0.Cout;
  return (q <= 512 && q % 32 == 0) ? q : 0;
}

// The Gram matrix G = y y^T of the HPCA rule as a layer: 1x1, Cout -> Cout, over a 1-D image of PR positions.
static Geo gram_geo(const Geo& g, long long PR) {
  Geo e = g;
  e.nd = 3; e.B = 1; e.Cin = g.Cout; e.Cout = g.Cout;
  e.iD = 1; e.iH = 1; e.iW = (int)PR;
  e.kD = e.kH = e.kW = 1; e.sD = e.sH = e.sW = 1;
  e.pD = e.pH = e.pW = e.qD = e.qH = e.qW = 0;
  e.oD = 1; e.oH = 1; e.oW = (int)PR;
  e.taps = 1; e.K = g.Cout; e.inS = PR; e.outS = PR; e.transposed = 0;
  return e;
}

static bool plan_layer(const Geo& g, int prec, Plan* P, int trq = 0, bool gram = false);

static bool plan_layer_search(const Geo& g, int prec, Plan* P, int trq, bool gram) {
  Plan& q = *P;
  q.ok = false;
  if (g.transposed || g.sD != 1 || g.sH != 1 || g.sW != 1) return false;
  if (g.Cout % 16 || g.taps > kMaxTaps) return false;
  q.CT = g.Cout <= 512 ? g.Cout : 512;           // channel tile of the forward kernel (TMEM: 512 columns)
  if (trq) {     // whole blocks, and at most 256 columns when possible so that two accumulator buffers fit TMEM
    q.CT = trq <= 256 ? trq * (256 / trq) : trq;
    if (q.CT > g.Cout) q.CT = g.Cout;
  }
  if (g.Cout % q.CT) return false;
  q.n_ct = g.Cout / q.CT;
  if (g.kD > 3 || g.kH > 9 || g.kW > 9) return false;
  q.HP = g.iH + g.pH + g.qH; q.WP = g.iW + g.pW + g.qW;
  const int xD = g.iD + g.pD + g.qD;
  q.plane = q.HP * q.WP;
  const long long qimg = (long long)xD * q.plane;
  if (qimg * g.B >= (1LL << 31)) return false;
  q.Qimg = (int)qimg;
  q.PTOT = qimg * g.B;
  q.NSLAB = (g.Cin + 15) / 16; q.CC = q.NSLAB * 2; q.C8 = g.Cout / 8;
  q.maxshift = ((g.kD - 1) * q.HP + (g.kH - 1)) * q.WP + (g.kW - 1);
  const int sms = num_sms();

  // ---------------- forward ----------------
  q.f_HL = 2;                                   // forward is always split (exact winners)
  const int halo = (g.kH - 1) * q.WP + (g.kW - 1);
  q.f_nseg = g.kD;
  // Stacking [w_hi | w_lo] along N (2 MMAs instead of 3) measured SLOWER on B200 for every channel count
  // (the doubled TMEM footprint halves the M-blocks per tile and doubles the epilogue's TMEM reads:
  // profiles/README.md, "forward stacking"), so the planner keeps the 3-MMA form.  HEBB_STACKF=1 re-enables it.
  static const bool want_stackf = [] { const char* e = getenv("HEBB_STACKF"); return e && e[0] == '1'; }();
  q.stackF = (want_stackf && q.f_HL == 2 && q.CT <= 64 && !g.transposed) ? 1 : 0;
  const int fcw = (q.stackF ? 2 : 1) * q.CT;
  const uint32_t w_tap = (uint32_t)q.f_HL * 2 * q.CT * 16;     // one tap of one slab: [k-chunk][hi|lo][CT rows]
  const uint32_t misc = (uint32_t)(12 * q.CT * 4 + 8 * 64 + 64);
  // Weight stages hold a GROUP of taps moved by one bulk copy: every stage costs the producer and the MMA
  // warp a full mbarrier round trip (~600 cycles measured with per-tap stages, which starved the tensor
  // pipe), so prefer the largest group (a whole kd plane, else a kernel row, else one tap) that fits.
  const int seg_taps = g.kH * g.kW;
  const int wg_opts[3] = {seg_taps, g.kW, 1};
  bool found = false;
  for (int mb = 4; mb >= 1 && !found; mb >>= 1) {
    if (mb > 1 && mb * fcw > 256) continue;
    // do not make tiles so large that the grid cannot fill the machine
    static const double mb_waves = [] { const char* e = getenv("HEBB_MB_WAVES"); return e ? atof(e) : 2.0; }();
    if (mb > 1 && (double)cdiv(q.PTOT, 128LL * mb) < mb_waves * sms) continue;
    const int seglen = round_up_i(128 * mb + halo, 8);
    const uint32_t xst = (uint32_t)q.f_HL * 2 * seglen * 16;
    for (int nx = 3; nx >= 2 && !found; --nx) {
      for (int wi = 0; wi < 3 && !found; ++wi) {
        const int wg = wg_opts[wi];
        if (wi > 0 && wg == wg_opts[wi - 1]) continue;
        for (int nw = (wg == 1 ? 8 : 4); nw >= 2 && !found; --nw) {
          const uint32_t tot = nx * xst + nw * wg * w_tap + misc + 256;
          if (tot <= (uint32_t)kSmemLimit - 1024) {
            q.MB = mb; q.TILE_M = 128 * mb; q.f_SEGLEN = seglen; q.XST = nx; q.WST = nw; q.WG = wg;
            q.f_w_stage = (uint32_t)wg * w_tap;
            q.f_x_stage = xst; found = true;
          }
        }
      }
    }
  }
  if (!found) return false;
  q.NACC = (2 * q.MB * fcw <= 512) ? 2 : 1;
  q.f_tmem = pow2_cols(q.NACC * q.MB * fcw);
  q.f_off_w = q.XST * q.f_x_stage;
  q.f_off_misc = q.f_off_w + q.WST * q.f_w_stage;
  q.f_smem = q.f_off_misc + misc;

  // ---------------- dW ----------------
  // Search (tap grouping, cin tile, cout tile, positions per stage, stages) for the cheapest plan that
  // fits shared memory and TMEM.  Cost model (cycles per SM): a SWIZZLE_NONE MN-major tcgen05.mma costs
  // ~135 cycles however small it is (measured, profiles/umma_rate_r1.txt) or N/2 when math-bound; staged
  // bytes arrive at ~40 B/cycle/SM from L2.
  q.d_HL = (prec == HEBB_PREC_BF16) ? 1 : 2;
  const int hl3 = (q.d_HL == 2) ? 3 : 1;
  double best_cost = 1e300;
  found = false;
  // HEBB_DW_REUSE=0/1 forces the loop order of the contraction kernel (profiling aid)
  static const int want_reuse = [] { const char* e = getenv("HEBB_DW_REUSE"); return (e && (e[0] == '0' || e[0] == '1')) ? e[0] - '0' : -1; }();
  for (int by_kh = 0; by_kh <= 1; ++by_kh) {
    const int ngrp = by_kh ? g.kD * g.kH : g.kD;
    if (by_kh && g.kH == 1) continue;
    for (int cm = 128; cm >= 64; cm -= 64) {
     for (int sm = 0; sm <= (q.d_HL == 2 ? 1 : 0); ++sm) {
      for (int sn = 0; sn <= (q.d_HL == 2 ? 1 : 0); ++sn) {
      const int cpt = sm ? cm / 16 : cm / 8;             // 8-channel chunks of x per cin tile
      const int cm_chunks = (q.CC < cpt) ? q.CC : cpt;
      const int n_cin = (int)cdiv(q.CC, cpt);
      for (int nrep = 1; nrep <= 4; ++nrep) {
      if (nrep == 1 && cm == 128 && q.CC <= cpt / 2) continue;   // a 64-row instruction already holds everything
      // tap replication: nrep shifted copies of the x tile share one instruction's M rows; needs every channel in
      // one tile, room for nrep replicas in the M rows, and (split mode) hi/lo stacked inside each replica
      if (nrep > 1 && (n_cin != 1 || nrep > g.kW || (q.d_HL == 2 && !sm) || nrep * (sm ? 2 : 1) * cm_chunks > cm / 8)) continue;
      const int st_row = (int)cdiv(g.kW, nrep);          // super-taps per kernel row
      const int gst = by_kh ? st_row : g.kH * st_row;    // super-taps (= accumulator column groups) per tap group
      const int rhalo = (st_row - 1) * nrep + (by_kh ? 0 : (g.kH - 1) * q.WP);
      const int cn_opts[7] = {256, 128, 64, 48, 32, 16, g.Cout};
      for (int ci = 0; ci < 7; ++ci) {
        const int cn = cn_opts[ci];
        const int neff = sn ? 2 * cn : cn;
        if (cn > g.Cout || neff > 256 || gst * neff > 512 || cn % 16) continue;
        const int n_cout = (int)cdiv(g.Cout, cn);
        for (int blk = 1024; blk >= 64; blk >>= 1) {
          if (gram && q.PTOT % blk) continue;          // must tile the host plan's packed positions exactly
          const int seglen = round_up_i(blk + rhalo, 8);     // staged r positions: the block plus the tap halo
          const uint32_t xb = (uint32_t)nrep * q.d_HL * cm_chunks * blk * 16;
          const uint32_t rb = (uint32_t)q.d_HL * (cn / 8) * seglen * 16;
          for (int st = 3; st >= 2; --st) {
            // the A descriptor always spans cm/8 chunks: rows past the real ones read whatever follows in
            // shared memory (discarded rows) but must stay inside the allocation
            const uint64_t ring = (uint64_t)st * (xb + rb);
            const uint64_t last_read = (uint64_t)(st - 1) * (xb + rb) + (uint64_t)(q.d_HL - 1) * cm_chunks * blk * 16 +
                                       (uint64_t)(cm / 8) * blk * 16 + 256;
            uint64_t tot = ring > last_read ? ring : last_read;
            tot = (tot + 127) / 128 * 128 + 8 * 16 + 64;
            if (tot > (uint64_t)kSmemLimit - 1024) continue;
            // measured cycles per SWIZZLE_NONE MN-major tcgen05.mma (profiles/umma_rate_r1.txt); a run of >= 5
            // instructions that share the A tile through the collector costs ~38 (profiles/umma_rate3_r1.txt)
            (void)hl3;
            const int n_mma = (q.d_HL == 2) ? (sm && sn ? 1 : ((sm || sn) ? 2 : 3)) : 1;
            const int run = gst * ((q.d_HL == 2 && !sn) ? 2 : 1);      // longest run of instructions with one A tile
            const int reuse = (want_reuse >= 0) ? (want_reuse && gst <= 9) : ((run >= 5 && gst <= 9 && neff <= 96 && cm == 128) ? 1 : 0);
            const double floor_c = reuse ? 38.0 : ((cm == 64) ? 45.0 : 60.0);
            const double mma = (neff * 0.5625 > floor_c) ? neff * 0.5625 : floor_c;
            const double t_mma = (double)gst * (blk / 16) * n_mma * mma;
            const double t_ld = (double)q.d_HL * 16.0 * ((double)nrep * cm_chunks * blk + (cn / 8.0) * seglen) / 40.0;
            const double per_blk = (t_mma > t_ld ? t_mma : t_ld) + 800.0 + (st == 2 ? 0.15 * t_ld : 0.0);
            const double out_tiles = (double)ngrp * n_cin * n_cout;
            const double blocks = (double)cdiv(q.PTOT + q.maxshift, blk);
            double waves = out_tiles / sms;               // how unevenly the tasks fill the SMs
            waves = waves < 1.0 ? 1.0 : waves;
            double ctas = out_tiles < sms ? out_tiles * (double)((int)(sms / out_tiles)) : out_tiles;   // CTAs of the single wave
            if (ctas > sms) ctas = sms;
            const double cost = blocks * per_blk * out_tiles / ctas + 30000.0 * waves + 4.0 * gst * neff * waves;
            if (cost < best_cost) {
              best_cost = cost; found = true;
              q.d_by_kh = by_kh; q.BLK = blk; q.d_SEGLEN = seglen; q.ST = st; q.CN = cn; q.CM = cm;
              q.stackM = sm; q.stackN = sn; q.cpt = cpt; q.nrep = nrep; q.rhalo = rhalo; q.reuse = reuse;
              q.d_x_bytes = xb; q.d_stage = xb + rb;
              q.d_off_bar = (uint32_t)(tot - (8 * 16 + 64));
              q.d_smem = (uint32_t)tot;
            }
          }
        }
      }
      }
      }
     }
    }
  }
  if (!found) return false;
  q.n_cin_tiles = (int)cdiv(q.CC, q.cpt);
  q.CinP = q.n_cin_tiles * q.cpt * 8;
  q.Q = (q.stackM ? 2 : 1) * (q.stackN ? 2 : 1);
  q.ngrp = q.d_by_kh ? g.kD * g.kH : g.kD;
  q.n_cout_tiles = (int)cdiv(g.Cout, q.CN);
  q.d_tmem = pow2_cols((q.d_by_kh ? 1 : g.kH) * (int)cdiv(g.kW, q.nrep) * q.CN * (q.stackN ? 2 : 1));

  // ---------------- dW, swizzled-response variant ----------------
  // Layers with 64 response channels are bound by shared-memory bandwidth (one 4 KB A fetch per 32 cycles of math).
  // With the responses stored as [position][64 channels] (SWIZZLE_128B image) the B descriptor's atom stride can be
  // ONE position, so a single N = 192 instruction computes the three taps of a kernel row from one staged tile.
  // Cin = 64: three kh-shifted replicas of the x tile (rows of M), tap groups = kd planes; Cin = k*128: tap groups =
  // kernel rows, 128-channel tiles.
  static const int want_rsw = [] { const char* e = getenv("HEBB_DW_RSW"); return (e && e[0] == '0') ? 0 : 1; }();
  q.rsw = 0; q.rs_BLK = 0;
  // (Cout = 128 only with Cin = 64: two 64-channel response planes; wider layers are math-bound as they are)
  // Split precision (bf16x3): the hi and lo response images are staged side by side; Cin = k*128 issues the classic
  // three products per kernel row (x_hi r_lo, x_hi r_hi, x_lo r_hi: 3 x N = 192 instead of 9 x N = 64), Cin = 64 stacks
  // [x_hi; x_lo] along M (two instructions per kernel row, tap groups = kernel rows).
  if (want_rsw && (prec == HEBB_PREC_BF16 || prec == HEBB_PREC_BF16X3) && !gram && !trq &&
      (g.Cout == 64 || (g.Cout == 128 && g.Cin == 64)) && q.n_ct == 1 &&
      g.kW == 3 && g.kH == 3 && (g.Cin == 64 || g.Cin % 128 == 0)) {
    const int HL = q.d_HL;
    const bool rep3 = g.Cin == 64 && HL == 1;
    q.rs_stackM = (g.Cin == 64 && HL == 2) ? 1 : 0;
    q.rs_nrep = rep3 ? 3 : 1; q.rs_by_kh = rep3 ? 0 : 1; q.rs_cpt = g.Cin == 64 ? 8 : 16;
    q.rs_n_cin = g.Cin == 64 ? 1 : g.Cin / 128; q.rs_CinP = g.Cin;
    // Cin = k*128, opt-in (HEBB_DW_RSW_PAIR=1): one CTA handles the kernel rows (kh 0, kh 1) of a plane -- two
    // instructions per k-step share the staged x tile, the response tile carries a halo of one image row -- and
    // another one the row kh 2 with half as many position splits.  Measured no faster (128 -> 64 @96x96x80: 2.19 vs
    // 2.09 ms): with one N = 192 instruction per k-step the layer already runs at 77 % of the measured tensor peak.
    static const int want_pair = [] { const char* e = getenv("HEBB_DW_RSW_PAIR"); return (e && e[0] == '1') ? 1 : 0; }();
    q.rs_pair = (!rep3 && HL == 1 && want_pair && g.Cout == 64 && q.rs_n_cin * g.kD * 3 <= sms) ? 1 : 0;
    q.rs_rhalo = (g.kW - 1) + (q.rs_pair ? q.WP : 0);
    const int blk_opts[2] = {(rep3 || HL == 2) ? 128 : 256, 128}, st_opts[2] = {(rep3 || q.rs_stackM) ? 3 : 2, 2};
    for (int o = 0; o < 2 && !q.rsw; ++o) {
      q.rs_BLK = blk_opts[o]; q.rs_ST = st_opts[o];
      q.rs_b_rows = round_up_i(q.rs_BLK + q.rs_rhalo + 7, 8);
      q.rs_x_bytes = (uint32_t)q.rs_nrep * HL * q.rs_cpt * q.rs_BLK * 16;
      q.rs_stage = q.rs_x_bytes + (uint32_t)HL * q.rs_b_rows * 128;
      // the last A descriptor (rep3: start at replica 2, spanning 16 chunks) must stay inside the allocation
      const uint64_t ring = (uint64_t)q.rs_ST * q.rs_stage;
      const uint64_t last_read = (uint64_t)(q.rs_ST - 1) * q.rs_stage + (rep3 ? 2ull * 8 * q.rs_BLK * 16 : 0ull) +
                                 16ull * q.rs_BLK * 16 + 256;
      uint64_t tot = ring > last_read ? ring : last_read;
      tot = (tot + 1023) / 1024 * 1024 + 8 * 16 + 64;
      q.rs_off_bar = (uint32_t)(tot - (8 * 16 + 64)); q.rs_smem = (uint32_t)tot;
      q.rs_tmem = pow2_cols(((rep3 || q.rs_pair) ? 2 : 1) * 64 * g.kW);
      if (q.rs_stage % 1024 == 0 && q.rs_x_bytes % 1024 == 0 && tot <= (uint64_t)kSmemLimit - 1024 && q.rs_tmem <= 512) q.rsw = 1;
    }
  }

  // packed position space: multiples of both tile sizes
  // (a Gram plan runs over another plan's PR positions: its stage size divides them, see the search above)
  int big = gram ? q.BLK : (q.TILE_M > q.BLK ? q.TILE_M : q.BLK);
  if (q.rsw && q.rs_BLK > big) big = q.rs_BLK;
  q.PR = (q.PTOT + big - 1) / big * big;
  // the contraction pairs x[q] with r[q - shift]: its position blocks run over q in [0, PTOT + maxshift), so the
  // x planes extend that far (zeros) and the r planes carry zero pads of maxshift positions on both sides
  q.total_blocks = (int)cdiv(q.PTOT + q.maxshift, q.BLK);
  const int blk_max = (q.rsw && q.rs_BLK > q.BLK) ? q.rs_BLK : q.BLK;
  // rsw: the x replicas reach (nrep-1)*WP positions past a block
  q.PA = (q.PR + q.maxshift + blk_max + (q.rsw ? (q.rs_nrep - 1) * q.WP : 0) + 16 + 7) / 8 * 8;
  q.r_lead = (q.maxshift + ((q.rsw && q.rs_pair) ? q.WP : 0) + 7) / 8 * 8;      // pair mode stages one more image row ahead
  q.PRS = q.r_lead + ((q.PR + q.maxshift + blk_max + 16 + 7) / 8 * 8);
  q.f_ntiles = (int)(q.PR / q.TILE_M);
  const int out_tiles = q.ngrp * q.n_cin_tiles * q.n_cout_tiles;
  // one wave: never more CTAs than SMs (a 149th CTA would double the kernel's duration)
  int ps = sms / out_tiles;
  if (ps > q.total_blocks) ps = q.total_blocks;
  if (ps < 1) ps = 1;
  q.blocks_per_split = (int)cdiv(q.total_blocks, ps);
  q.PS = (int)cdiv(q.total_blocks, q.blocks_per_split);
  if (q.rsw) {
    q.rs_total_blocks = (int)cdiv(q.PTOT + q.maxshift, q.rs_BLK);
    // virtual tiles: pair mode counts a (kh0, kh1) group twice (it gets twice the position splits)
    const int rs_tiles = (q.rs_pair ? g.kD * 3 : (q.rs_by_kh ? g.kD * g.kH : g.kD)) * q.rs_n_cin * (g.Cout / 64);
    int rps = sms / rs_tiles;
    if (rps > q.rs_total_blocks) rps = q.rs_total_blocks;
    if (rps < 1) rps = 1;
    q.rs_PSu = rps;
    q.rs_blocks_per_split = (int)cdiv(q.rs_total_blocks, rps);
    q.rs_bps2 = (int)cdiv(q.rs_total_blocks, 2 * rps);
    q.rs_PS = q.rs_pair ? 2 * rps : (int)cdiv(q.rs_total_blocks, q.rs_blocks_per_split);     // partial planes summed by finalize
  }

  // ---------------- workspace ----------------
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  q.o_inv = take(sizeof(float) * (g.Cout > g.Cin ? g.Cout : g.Cin));
  q.o_rsum = take(sizeof(float) * g.Cout);
  q.o_err = take(256);
  q.o_xp[0] = take((size_t)q.CC * q.PA * 16);
  q.o_xp[1] = take((size_t)q.CC * q.PA * 16);
  q.o_rp[0] = take((size_t)q.C8 * q.PRS * 16);
  q.o_rp[1] = take((size_t)q.C8 * q.PRS * 16);
  q.o_wp = take((size_t)q.NSLAB * g.taps * q.f_HL * 2 * g.Cout * 16);
  q.fix_cap = (int)((q.PTOT * 8) < (1LL << 18) ? (q.PTOT * 8) : (1LL << 18));     // (a transposed layer has 8 outputs per position)
  q.o_fix = take(sizeof(int) * (size_t)q.fix_cap);
  {
    size_t hp = (size_t)q.PS * q.Q * g.taps * q.CinP * g.Cout * sizeof(float);
    const size_t hp2 = q.rsw ? (size_t)q.rs_PS * (q.rs_stackM ? 2 : 1) * g.taps * q.rs_CinP * g.Cout * sizeof(float) : 0;
    q.o_hpart = take(hp > hp2 ? hp : hp2);
  }
  q.gram_ok = false; q.o_gram = q.o_hpart2 = 0;
  if (!gram && !trq) {
    // HPCA (hebb.py:122-135) also needs G = y y^T: the same contraction kernel over the packed responses,
    // planned as a 1x1 layer Cout -> Cout over the PR packed positions (a 1-D "image")
    Geo g2 = gram_geo(g, q.PR);
    Plan P2;
    if (plan_layer(g2, prec, &P2, 0, true) && P2.PR == q.PR) {
      q.gram_ok = true;
      q.o_gram = take(sizeof(float) * (size_t)g.Cout * g.Cout);
      q.o_hpart2 = take((size_t)P2.PS * P2.Q * P2.CinP * g.Cout * sizeof(float));
    }
  }
  q.total = off;
  q.ok = true;
  return true;
}

// The search above costs tens of microseconds and runs for every launch and workspace query: keep the last
// few plans per host thread (a network has a few dozen distinct layer shapes).
static bool geo_eq(const Geo& a, const Geo& b) {
  return a.nd == b.nd && a.B == b.B && a.Cin == b.Cin && a.Cout == b.Cout && a.iD == b.iD && a.iH == b.iH &&
         a.iW == b.iW && a.kD == b.kD && a.kH == b.kH && a.kW == b.kW && a.sD == b.sD && a.sH == b.sH && a.sW == b.sW &&
         a.pD == b.pD && a.pH == b.pH && a.pW == b.pW && a.qD == b.qD && a.qH == b.qH && a.qW == b.qW &&
         a.oD == b.oD && a.oH == b.oH && a.oW == b.oW && a.taps == b.taps && a.K == b.K && a.inS == b.inS &&
         a.outS == b.outS && a.transposed == b.transposed;
}

static bool plan_layer(const Geo& g, int prec, Plan* P, int trq, bool gram) {
  struct Entry { Geo g; int prec, trq, gram, sms; Plan plan; bool ok; };
  thread_local std::vector<Entry> cache;
  const int sms = num_sms();
  for (const Entry& e : cache)
    if (e.prec == prec && e.trq == trq && e.gram == (int)gram && e.sms == sms && geo_eq(e.g, g)) {
      *P = e.plan;
      return e.ok;
    }
  Entry e;
  e.g = g; e.prec = prec; e.trq = trq; e.gram = gram; e.sms = sms;
  e.ok = plan_layer_search(g, prec, &e.plan, trq, gram);
  if (cache.size() >= 128) cache.erase(cache.begin());
  cache.push_back(e);
  *P = e.plan;
  return e.ok;
}

bool tc_supported(const Geo& g, int prec) {
  Plan P;
  Geo e;
  return equivalent_1x1(g, &e) && plan_layer(e, prec, &P, tr_quantum(g));
}

size_t tc_workspace_bytes(const Geo& g, int prec) {
  Plan P;
  Geo e;
  if (!equivalent_1x1(g, &e) || !plan_layer(e, prec, &P, tr_quantum(g))) return 0;
  return P.total;
}

int tc_describe_plan(const Geo& g0, int prec, int* o, int n) {
  Plan P;
  Geo g;
  if (!equivalent_1x1(g0, &g) || !plan_layer(g, prec, &P, tr_quantum(g0))) return 0;
  const int v[] = {P.MB, P.f_SEGLEN, P.XST, P.WST, P.NACC, (int)P.f_tmem, P.f_ntiles, (int)P.f_smem,
                   P.d_by_kh, P.CM, P.CN, P.BLK, P.ST, P.d_SEGLEN, P.ngrp, P.n_cin_tiles, P.n_cout_tiles, P.PS,
                   P.total_blocks, (int)P.d_tmem, (int)P.d_smem, P.d_HL, (int)(P.total >> 20), P.stackM, P.stackN, P.CT, P.n_ct, P.nrep, P.WG,
                   P.reuse, P.rhalo, P.rsw, P.rsw ? P.rs_BLK : 0, P.rsw ? P.rs_ST : 0, P.rsw ? (int)P.rs_smem : 0,
                   P.rsw ? (int)P.rs_tmem : 0, P.rsw ? P.rs_PS : 0, P.rsw ? P.rs_stackM : 0};
  const int m = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < m; ++i) o[i] = v[i];
  return m;
}

static unsigned ew_grid(long long n) {
  long long gx = cdiv(n, 256);
  const long long cap = (long long)num_sms() * 16;
  return (unsigned)(gx > cap ? cap : (gx < 1 ? 1 : gx));
}

// Fills the parameter block of the contraction kernel for plan P / geometry g and launches it.  `PA` is the
// position stride between 8-channel planes of the x operand (P.PA for packed activations).
static int launch_dw(const Plan& P, const Geo& g, const uint4* xp0, const uint4* xp1, const uint4* rp0, const uint4* rp1,
                     float* hpart, int* err, long long PA, long long PRS, cudaStream_t st, bool rsw = false) {
  DwParams d;
  d.xp[0] = xp0; d.xp[1] = xp1; d.rp[0] = rp0; d.rp[1] = rp1; d.hpart = hpart; d.err = err;
  d.Cin = g.Cin; d.Cout = g.Cout; d.CC = P.CC; d.C8 = P.C8; d.taps = g.taps; d.HL = P.d_HL;
  d.PA = PA; d.PRS = PRS;
  for (int i = 0; i < 9; ++i) d.grp_base[i] = 0;
  for (int t = 0; t < kMaxTaps; ++t) { d.st_boff[t] = 0; d.st_aoff[t] = 0; d.st_first[t] = 0; d.st_n[t] = 0; }
  d.r_lead = P.r_lead;
  int dgrid;
  for (int i = 0; i < 9; ++i) d.grp_w[i] = 1;
  for (int i = 0; i < 18; ++i) { d.vt_grp[i] = i < 9 ? i : 0; d.vt_w[i] = 0; }
  d.blocks_per_split2 = 0;
  if (rsw) {
    // swizzled responses: one N = 64*kW instruction per kernel row (see plan_layer_search)
    d.rsw = 1; d.ncopy = g.kW; d.b_rows = P.rs_b_rows; d.reuse = 0;
    d.BLK = P.rs_BLK; d.SEGLEN = P.rs_b_rows; d.total_blocks = P.rs_total_blocks;
    d.blocks_per_split = P.rs_blocks_per_split; d.blocks_per_split2 = P.rs_bps2; d.PS = P.rs_PS;
    d.rhalo = P.rs_rhalo;
    d.nrep = P.rs_nrep; d.rep_stride = P.WP; d.tap_rep = g.kW;
    int nst = 0, gi = 0, nvt = 0;
    for (int kd = 0; kd < g.kD; ++kd) {
      if (!P.rs_by_kh) {
        // tap group = kd plane; super-tap 0: x replicas 0,1 = kernel rows kh 0,1 (M = 128), super-tap 1: replica 2 = kh 2
        d.grp_base[gi] = kd * P.plane; d.grp_st_begin[gi] = nst;
        for (int kh0 = 0; kh0 < g.kH; kh0 += 2, ++nst) {
          d.st_aoff[nst] = kh0 * (P.rs_cpt * P.rs_BLK);              // replica kh0 starts kh0 * cpt * BLK vectors into the x region
          d.st_first[nst] = (kd * g.kH + kh0) * g.kW;
          d.st_n[nst] = (g.kH - kh0 < 2) ? (g.kH - kh0) : 2;
        }
        d.vt_grp[nvt] = gi; d.vt_w[nvt++] = 0;
        ++gi;
      } else if (P.rs_pair) {
        // group (kh 0, kh 1): two super-taps whose B operands start one image row apart in the staged response tile
        // (tap offset = grp_base + rhalo - row offset - column copy); weight 2.  Group (kh 2): weight 1.
        d.grp_base[gi] = kd * P.plane; d.grp_st_begin[gi] = nst; d.grp_w[gi] = 2;
        d.st_boff[nst] = P.WP * 8; d.st_first[nst] = (kd * g.kH + 0) * g.kW; d.st_n[nst] = 1; ++nst;
        d.st_boff[nst] = 0;        d.st_first[nst] = (kd * g.kH + 1) * g.kW; d.st_n[nst] = 1; ++nst;
        d.vt_grp[nvt] = gi; d.vt_w[nvt++] = 0; d.vt_grp[nvt] = gi; d.vt_w[nvt++] = 1;
        ++gi;
        d.grp_base[gi] = kd * P.plane + 2 * P.WP; d.grp_st_begin[gi] = nst; d.grp_w[gi] = 1;
        d.st_boff[nst] = P.WP * 8; d.st_first[nst] = (kd * g.kH + 2) * g.kW; d.st_n[nst] = 1; ++nst;
        d.vt_grp[nvt] = gi; d.vt_w[nvt++] = 0;
        ++gi;
      } else {
        for (int kh = 0; kh < g.kH; ++kh, ++gi, ++nst) {
          d.grp_base[gi] = kd * P.plane + kh * P.WP; d.grp_st_begin[gi] = nst;
          d.st_first[nst] = (kd * g.kH + kh) * g.kW; d.st_n[nst] = 1;
          d.vt_grp[nvt] = gi; d.vt_w[nvt++] = 0;
        }
      }
    }
    d.ngrp = gi; d.n_vt = nvt;
    for (int i = gi; i < 10; ++i) d.grp_st_begin[i] = nst;
    d.stackM = P.rs_stackM; d.stackN = 0; d.cpt = P.rs_cpt;
    d.CM = 128; d.n_cin_tiles = P.rs_n_cin; d.CN = 64; d.n_cout_tiles = g.Cout / 64; d.ST = P.rs_ST; d.CinP = P.rs_CinP;
    d.stage_bytes = P.rs_stage; d.x_bytes = P.rs_x_bytes; d.off_bar = P.rs_off_bar; d.tmem_cols = P.rs_tmem;
    dgrid = nvt * d.n_cin_tiles * d.n_cout_tiles * P.rs_PSu;
    if (P.rs_pair)     // the (kh 2) groups fill only half of the partial planes the finalize pass sums
      HEBB_CUDA_TRY(cudaMemsetAsync(hpart, 0, (size_t)P.rs_PS * (P.rs_stackM ? 2 : 1) * g.taps * P.rs_CinP * g.Cout * sizeof(float), st));
  } else {
    d.rsw = 0; d.ncopy = 1; d.b_rows = 0; d.rep_stride = 1; d.tap_rep = 1; d.n_vt = P.ngrp;
    d.BLK = P.BLK; d.SEGLEN = P.d_SEGLEN; d.total_blocks = P.total_blocks;
    d.rhalo = P.rhalo; d.reuse = P.reuse;
    d.blocks_per_split = P.blocks_per_split; d.PS = P.PS; d.ngrp = P.ngrp;
    d.nrep = P.nrep;
    // tap groups (one CTA column set each) -> super-taps (one accumulator column group each) -> taps
    int nst = 0, gi = 0;
    for (int kd = 0; kd < g.kD; ++kd) {
      if (!P.d_by_kh) { d.grp_base[gi] = kd * P.plane; d.grp_st_begin[gi] = nst; }
      for (int kh = 0; kh < g.kH; ++kh) {
        if (P.d_by_kh) { d.grp_base[gi] = kd * P.plane + kh * P.WP; d.grp_st_begin[gi] = nst; }
        for (int kw0 = 0; kw0 < g.kW; kw0 += P.nrep, ++nst) {
          d.st_boff[nst] = P.rhalo - ((P.d_by_kh ? 0 : kh * P.WP) + kw0);    // tap offset = grp_base + rhalo - st_boff (+ replica)
          d.st_first[nst] = (kd * g.kH + kh) * g.kW + kw0;
          d.st_n[nst] = (g.kW - kw0 < P.nrep) ? (g.kW - kw0) : P.nrep;
        }
        if (P.d_by_kh) ++gi;
      }
      if (!P.d_by_kh) ++gi;
    }
    for (int i = gi; i < 10; ++i) d.grp_st_begin[i] = nst;
    d.stackM = P.stackM; d.stackN = P.stackN; d.cpt = P.cpt;
    d.CM = P.CM; d.n_cin_tiles = P.n_cin_tiles; d.CN = P.CN; d.n_cout_tiles = P.n_cout_tiles; d.ST = P.ST; d.CinP = P.CinP;
    d.stage_bytes = P.d_stage; d.x_bytes = P.d_x_bytes; d.off_bar = P.d_off_bar; d.tmem_cols = P.d_tmem;
    dgrid = P.ngrp * P.n_cin_tiles * P.n_cout_tiles * P.PS;
  }
  HEBB_CUDA_TRY(cudaFuncSetAttribute(dw_swta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  dw_swta_kernel<<<dgrid, 192, kSmemLimit, st>>>(d);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

// Launch wrappers for the fused small-channel path (fused_path.cu), which shares the weight packing and the
// deterministic finalize pass with the kernels above.
int tc_launch_pack_w(const float* W, void* wp, int Cin, int Cout, int taps, int NSLAB, int CT, cudaStream_t st) {
  const long long n = (long long)NSLAB * taps * 2 * 2 * Cout;
  pack_w_kernel<<<ew_grid(n), 256, 0, st>>>(W, reinterpret_cast<uint4*>(wp), Cin, Cout, taps, NSLAB, 2, CT, 0, 0, nullptr);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

int tc_launch_finalize(const float* hpart, const float* rsum, const float* W, float* dw, int n_part, int taps, int Cin,
                       int CinP, int Cout, cudaStream_t st) {
  long long gx = cdiv((long long)taps * Cin * Cout, 32);
  const long long cap = (long long)num_sms() * 32;
  if (gx > cap) gx = cap;
  tc_finalize_kernel<<<(unsigned)gx, 256, 0, st>>>(hpart, rsum, W, dw, n_part, taps, Cin, CinP, Cout);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

int tc_conv_step(const Geo& g0, const float* x, const float* W, const float* bias, float kinv, float* y,
                 int32_t* winner, float* delta_w, void* ws, size_t ws_bytes, unsigned flags, int prec,
                 cudaStream_t st, int aux, double* ystats, int* ystats_written) {
  if (ystats_written) *ystats_written = 0;
  Plan P;
  Geo g;
  const int trq = tr_quantum(g0);
  if (!equivalent_1x1(g0, &g) || !plan_layer(g, prec, &P, trq)) return HEBB_ESHAPE;
  const bool tr = g0.transposed != 0;
  if (!ws || ws_bytes < P.total) return HEBB_EWS;
  char* base = static_cast<char*>(ws);
  float* inv = reinterpret_cast<float*>(base + P.o_inv);
  float* rsum = reinterpret_cast<float*>(base + P.o_rsum);
  int* err = reinterpret_cast<int*>(base + P.o_err);       // the near-tie counter sits 16 bytes after this word
  int* wd = watchdog_word() ? watchdog_word() : err;       // where a timed-out wait leaves its code (pinned host memory)
  uint4* xp0 = reinterpret_cast<uint4*>(base + P.o_xp[0]);
  uint4* xp1 = reinterpret_cast<uint4*>(base + P.o_xp[1]);
  uint4* rp0 = reinterpret_cast<uint4*>(base + P.o_rp[0]);
  uint4* rp1 = reinterpret_cast<uint4*>(base + P.o_rp[1]);
  uint4* wp = reinterpret_cast<uint4*>(base + P.o_wp);
  float* hpart = reinterpret_cast<float*>(base + P.o_hpart);
  // weight-gradient mode (hebb_conv_wgrad): `y` holds dL/dy; only x is packed, dL/dy takes the place of the
  // responses, and the finalize pass adds the plain contraction (no decay term) to delta_w
  const bool wgrad = (flags & HEBB_F_WGRAD_INTERNAL) != 0;
  if (wgrad && tr) return HEBB_ESHAPE;
  // HPCA (hebb.py:122-135): the response is y itself, the decay is tril(y y^T) W
  const bool hpca = (flags & HEBB_F_RULE_HPCA) != 0 && (flags & HEBB_F_UPDATE) != 0;
  if (hpca && (tr || !P.gram_ok)) return HEBB_ESHAPE;
  const bool upd = (flags & HEBB_F_UPDATE) != 0 || wgrad;
  // swizzled-response variant of the update (forward epilogue writes the layout, the dW kernel reads it)
  const bool rsw = P.rsw && (flags & HEBB_F_UPDATE) != 0 && !wgrad && !hpca && !tr && g.taps == g0.taps;
  rp0 += (long long)P.r_lead * (rsw ? 8 : 1);      // position 0 of the first plane (rsw: 8 vectors per position)
  rp1 += (long long)P.r_lead * (rsw ? 8 : 1);

  // profiling aid: HEBB_F_ONLY_* re-run one stage on the scratch left by a preceding full call
  const unsigned only = flags & (HEBB_F_ONLY_PACK | HEBB_F_ONLY_FWD | HEBB_F_ONLY_DW);
  const bool do_pack = !only || (only & HEBB_F_ONLY_PACK);
  const bool do_fwd = !wgrad && (!only || (only & HEBB_F_ONLY_FWD));
  const bool do_dw = upd && (!only || (only & HEBB_F_ONLY_DW));
  // Small weight tensors of plain layers: filter norms, weight packing and the zeroing of the per-call accumulators are
  // ONE launch (fused_path.cu: fused_prep_kernel) instead of memset + wnorm + pack_w -- these layers are launch-bound
  const long long n_wp = (long long)P.NSLAB * g.taps * P.f_HL * 2 * g.Cout;
  const bool merged_prep = !only && !wgrad && !tr && P.n_ct == 1 && P.f_HL == 2 && n_wp < (1LL << 16);
  const bool want_stats = ystats && do_fwd && !tr && P.n_ct == 1;
  if (merged_prep)
    HEBB_TRY(launch_layer_prep(W, wp, inv, base + P.o_rsum, P.o_xp[0] - P.o_rsum, want_stats ? ystats : nullptr, g.Cin, g.Cout, g.taps,
                               (flags & HEBB_F_WNRM) ? 1 : 0, st));
  if (do_fwd && !merged_prep) HEBB_CUDA_TRY(cudaMemsetAsync(base + P.o_rsum, 0, (P.o_xp[0] - P.o_rsum), st));   // rsum + err word
  if ((flags & HEBB_F_WNRM) && do_pack && !wgrad && !merged_prep) {
    if (tr)   // per INPUT channel over (Cout, taps) of the [Cout][Cin][taps] buffer (hebb3d.py:78 on the view)
      HEBB_TRY(launch_wnorm(W, nullptr, inv, g0.Cin, g0.taps, g0.Cout, (long long)g0.Cin * g0.taps, g0.taps, st));
    else
      HEBB_TRY(launch_wnorm(W, nullptr, inv, g.Cout, g.K, 1, 0, g.K, st));
  }

  PackGeo pg;
  pg.B = g.B; pg.Cin = g.Cin; pg.iD = g.iD; pg.iH = g.iH; pg.iW = g.iW; pg.pD = g.pD; pg.pH = g.pH; pg.pW = g.pW;
  pg.HP = P.HP; pg.WP = P.WP; pg.plane = P.plane; pg.Qimg = P.Qimg; pg.CC = P.CC; pg.PA = P.PA; pg.PTOT = P.PTOT;
  // the contraction kernel reads r[q - shift]: zero pads around the packed responses, written by the pack pass
  const bool zpads = upd && P.maxshift > 0;
  pg.rhi = rp0; pg.rlo = P.d_HL == 2 ? rp1 : nullptr; pg.C8 = P.C8; pg.PR = P.PR; pg.PRS = P.PRS;
  pg.r_lead = zpads ? P.r_lead : 0; pg.r_tail = zpads ? (int)(P.PRS - P.r_lead - P.PR) : 0; pg.r_ushift = rsw ? 3 : 0;
  const bool nhwc = wgrad && ((aux >> 16) & 1);         // hebb_conv_wgrad on channels_last tensors
  if (nhwc) { pg.sC = 1; pg.sW = g.Cin; pg.sH = (long long)g.iW * g.Cin; pg.sD = (long long)g.iH * g.iW * g.Cin; pg.sB = g.inS * g.Cin; }
  else { pg.sW = 1; pg.sH = g.iW; pg.sD = (long long)g.iH * g.iW; pg.sC = g.inS; pg.sB = g.inS * g.Cin; }
  const bool gathered = !tr && g.taps != g0.taps;      // few-input-channel layer re-stated as a 1x1 layer
  if (do_pack && gathered) {
    GatherGeo gg;
    gg.B = g0.B; gg.Cin = g0.Cin; gg.iD = g0.iD; gg.iH = g0.iH; gg.iW = g0.iW; gg.pD = g0.pD; gg.pH = g0.pH; gg.pW = g0.pW;
    gg.kH = g0.kH; gg.kW = g0.kW; gg.taps = g0.taps; gg.oD = g0.oD; gg.oH = g0.oH; gg.oW = g0.oW; gg.Kp = g.Cin;
    gg.CC = P.CC; gg.PA = P.PA; gg.PTOT = P.PTOT;
    if (g0.kD == 3 && g0.kH == 3 && g0.kW == 3) pack_x_gather_kernel<3, 3, 3><<<ew_grid(P.PA), 256, 0, st>>>(x, xp0, xp1, gg);
    else if (g0.kD == 1 && g0.kH == 3 && g0.kW == 3) pack_x_gather_kernel<1, 3, 3><<<ew_grid(P.PA), 256, 0, st>>>(x, xp0, xp1, gg);
    else pack_x_gather_kernel<0, 0, 0><<<ew_grid(P.PA), 256, 0, st>>>(x, xp0, xp1, gg);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  } else if (do_pack) {
    pack_x_kernel<<<ew_grid((long long)P.CC * P.PA), 256, 0, st>>>(x, xp0, xp1, pg);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }
  if (wgrad) {
    PackRGeo rg;
    const int gyC = (aux & 0xFFFF) ? (aux & 0xFFFF) : g.Cout;      // channels really present in dL/dy
    rg.B = g.B; rg.C = gyC; rg.C8 = P.C8;
    if (nhwc) { rg.sC = 1; rg.sW = gyC; rg.sH = (long long)g.oW * gyC; rg.sD = (long long)g.oH * g.oW * gyC; rg.sB = g.outS * gyC; }
    else { rg.sW = 1; rg.sH = g.oW; rg.sD = (long long)g.oH * g.oW; rg.sC = g.outS; rg.sB = g.outS * gyC; } rg.oD = g.oD; rg.oH = g.oH; rg.oW = g.oW; rg.WP = P.WP; rg.plane = P.plane;
    rg.Qimg = P.Qimg; rg.outS = g.outS; rg.PR = P.PR; rg.PTOT = P.PTOT; rg.PRS = P.PRS;
    pack_r_kernel<<<ew_grid((long long)P.C8 * P.PR), 256, 0, st>>>(y, rp0, P.d_HL == 2 ? rp1 : nullptr, rg);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }
  if (do_pack && !wgrad && !merged_prep) {
    const long long n = (long long)P.NSLAB * g.taps * P.f_HL * 2 * g.Cout;
    const size_t tile_bytes = (size_t)16 * (16 * g.taps + 1) * sizeof(float);
    if (!tr && n >= (1LL << 16) && tile_bytes <= 48 * 1024) {
      dim3 wg((unsigned)P.NSLAB, (unsigned)(g.Cout / 16));
      pack_w_tiled_kernel<<<wg, 256, tile_bytes, st>>>(W, wp, g.Cin, g.Cout, g.taps, P.f_HL, P.CT);
    } else {
      pack_w_kernel<<<ew_grid(n), 256, 0, st>>>(W, wp, g.Cin, g.Cout, g.taps, P.NSLAB, P.f_HL, P.CT, tr ? g0.taps : 0, trq,
                                                (tr && (flags & HEBB_F_WNRM)) ? inv : nullptr);
    }
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }

  // ---- forward ----
  FwdParams f;
  f.xp[0] = xp0; f.xp[1] = xp1; f.wp = wp; f.rp[0] = rp0; f.rp[1] = rp1;
  f.y = y; f.winner = winner; f.inv = ((flags & HEBB_F_WNRM) && !tr) ? inv : nullptr; f.bias = bias; f.rsum = rsum; f.err = wd;
  // BatchNorm statistics of y ride along when the layer is one channel tile of a plain convolution
  f.ystats = (ystats && do_fwd && !tr && P.n_ct == 1) ? ystats : nullptr;
  if (f.ystats) {
    if (!merged_prep) HEBB_CUDA_TRY(cudaMemsetAsync(ystats, 0, sizeof(double) * 2 * (size_t)g.Cout, st));
    if (ystats_written) *ystats_written = 1;
  }
  f.tr = tr ? (trq ? 2 : 1) : 0; f.trQ = trq; f.tD = g0.oD; f.tH = g0.oH; f.tW = g0.oW; f.CoutR = g0.Cout;
  f.Cout = g.Cout; f.CC = P.CC; f.NSLAB = P.NSLAB; f.taps = g.taps; f.nseg = P.f_nseg; f.HL = P.f_HL; f.RHL = P.d_HL;
  static const int fwd_dbg = [] { const char* e = getenv("HEBB_FWD_DBG"); return e ? atoi(e) : 0; }();
  f.dbg = fwd_dbg;
  // near-tie worklist: the forward error is a few 1e-6 of the pixel's response scale; 2.5e-4 leaves a wide margin and
  // still lists only ~1e-3 of the pixels (HEBB_TIE_REL overrides; 0 disables the exact pass)
  static const float tie_rel = [] { const char* e = getenv("HEBB_TIE_REL"); return e ? (float)atof(e) : 2.5e-4f; }();
  int* fix_count = err + 4;
  int* fix_list = reinterpret_cast<int*>(base + P.o_fix);
  f.fix_list = fix_list; f.fix_count = fix_count; f.fix_cap = P.fix_cap; f.tie_rel = tie_rel;
  f.stackF = P.stackF; f.CT = P.CT; f.n_ct = P.n_ct; f.fuse = (P.n_ct == 1 || trq) ? 1 : 0;   // transposed: grouped softmax when Cout*8 <= 512
  f.PA = P.PA; f.PR = P.PR; f.PRS = P.PRS; f.rsw = rsw ? 1 : 0; f.PTOT = P.PTOT; f.MB = P.MB; f.TILE_M = P.TILE_M; f.ntiles = P.f_ntiles; f.SEGLEN = P.f_SEGLEN;
  f.XST = P.XST; f.WST = P.WST; f.NACC = P.NACC; f.WG = P.WG;
  f.WP = P.WP; f.plane = P.plane; f.Qimg = P.Qimg; f.oD = g.oD; f.oH = g.oH; f.oW = g.oW;
  f.kinv = kinv; f.write_r = (upd && !hpca) ? 1 : 0;
  for (int s = 0; s < 4; ++s) f.seg_base[s] = s * P.plane;
  for (int s = 0; s <= 4; ++s) f.seg_tap_begin[s] = (s <= g.kD ? s : g.kD) * g.kH * g.kW;
  for (int t = 0; t < kMaxTaps; ++t) f.tap_off[t] = 0;
  for (int kd = 0, t = 0; kd < g.kD; ++kd)
    for (int kh = 0; kh < g.kH; ++kh)
      for (int kw = 0; kw < g.kW; ++kw, ++t) f.tap_off[t] = kh * P.WP + kw;
  f.x_stage_bytes = P.f_x_stage; f.w_stage_bytes = P.f_w_stage; f.off_w = P.f_off_w; f.off_misc = P.f_off_misc;
  f.tmem_cols = P.f_tmem;
  const long long fwork = (long long)P.f_ntiles * P.n_ct;
  const int fgrid = fwork < num_sms() ? (int)fwork : num_sms();
  if (do_fwd) {
    const int ch = (g.Cout % 32 == 0) ? 32 : 16;
    // single-pass epilogue: whole channel row in registers (fused soft-WTA, plain conv, <= 64 channels)
    const int sp = (f.fuse && !tr && P.CT <= 64 && P.CT % ch == 0 && !P.stackF) ? P.CT / ch : 0;
#define HEBB_FWD_LAUNCH(CHV, SPV)                                                                                   \
    do {                                                                                                            \
      HEBB_CUDA_TRY(cudaFuncSetAttribute(fwd_swta_kernel<CHV, SPV>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                         kSmemLimit));                                                               \
      fwd_swta_kernel<CHV, SPV><<<fgrid, 320, kSmemLimit, st>>>(f);                                                  \
    } while (0)
    if (ch == 32) {
      if (sp == 1) HEBB_FWD_LAUNCH(32, 1); else if (sp == 2) HEBB_FWD_LAUNCH(32, 2); else HEBB_FWD_LAUNCH(32, 0);
    } else {
      if (sp == 1) HEBB_FWD_LAUNCH(16, 1); else if (sp == 3) HEBB_FWD_LAUNCH(16, 3); else HEBB_FWD_LAUNCH(16, 0);
    }
#undef HEBB_FWD_LAUNCH
  }
  if (do_fwd) { HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED(); }
  if (do_fwd && P.n_ct > 1 && !trq && ((upd && !hpca) || winner)) {
    SmxParams sp;
    sp.y = y; sp.rp[0] = rp0; sp.rp[1] = rp1; sp.winner = winner; sp.rsum = rsum;
    sp.Cout = g.Cout; sp.RHL = P.d_HL; sp.WP = P.WP; sp.plane = P.plane; sp.Qimg = P.Qimg;
    sp.oD = g.oD; sp.oH = g.oH; sp.oW = g.oW; sp.PR = P.PR; sp.PRS = P.PRS; sp.PTOT = P.PTOT; sp.kinv = kinv;
    sp.fix_list = fix_list; sp.fix_count = fix_count; sp.fix_cap = P.fix_cap; sp.tie_rel = tie_rel;
    if (tr) {
      sp.Cout = g.Cout;
      swta_softmax_pack_T_kernel<<<(unsigned)cdiv(P.PR, 8), 256, 0, st>>>(sp, g0.oD, g0.oH, g0.oW, g0.Cout);
      if (upd) {
        HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
        rsum_from_packed_kernel<<<(unsigned)P.C8, 256, 0, st>>>(rp0, P.d_HL == 2 ? rp1 : nullptr, rsum, P.PR, P.PRS);
      }
    } else {
      swta_softmax_pack_kernel<<<(unsigned)(P.PR / 32), 256, 0, st>>>(sp);
    }
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }
  // exact re-evaluation of the listed near-tie pixels (winner indices bit-exact with the reference's argmax)
  if (do_fwd && winner && tie_rel > 0.f)
    HEBB_TRY(launch_winner_fixup(g0, x, W, (flags & HEBB_F_WNRM) ? inv : nullptr, bias, winner, fix_list, fix_count, P.fix_cap, st));
  if (!do_dw) return HEBB_OK;

  if (hpca) {
    PackRGeo rg;
    rg.sW = 1; rg.sH = g.oW; rg.sD = (long long)g.oH * g.oW; rg.sC = g.outS; rg.sB = g.outS * g.Cout;
    rg.B = g.B; rg.C = g.Cout; rg.C8 = P.C8; rg.oD = g.oD; rg.oH = g.oH; rg.oW = g.oW; rg.WP = P.WP; rg.plane = P.plane;
    rg.Qimg = P.Qimg; rg.outS = g.outS; rg.PR = P.PR; rg.PTOT = P.PTOT; rg.PRS = P.PRS;
    pack_r_kernel<<<ew_grid((long long)P.C8 * P.PR), 256, 0, st>>>(y, rp0, P.d_HL == 2 ? rp1 : nullptr, rg);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }
  // ---- dW ----
  HEBB_TRY(launch_dw(P, g, xp0, xp1, rp0, rp1, hpart, wd, P.PA, P.PRS, st, rsw));
  const int n_part = rsw ? P.rs_PS * (P.rs_stackM ? 2 : 1) : P.PS * P.Q;           // partial planes the finalize pass sums
  const int cin_p = rsw ? P.rs_CinP : P.CinP;
  {
    const long long n = (long long)g.taps * g.Cin * g.Cout;
    if (tr)
    {
      long long gx = cdiv((long long)g0.Cin * g0.Cout, 32);
      const long long cap = (long long)num_sms() * 32;
      tc_finalize_T_kernel<<<(unsigned)(gx > cap ? cap : gx), 256, 0, st>>>(hpart, rsum, W, delta_w, P.PS * P.Q, g0.Cin, P.CinP, g0.Cout, trq);
    }
    else
    if (n >= (1LL << 18) && n_part <= 32) {
      const float* rs = (wgrad || hpca) ? nullptr : rsum;
#define HEBB_FIN_LAUNCH(T)                                                                                              \
      do {                                                                                                              \
        dim3 fg((unsigned)cdiv((long long)g.Cin * g.taps, T), (unsigned)cdiv(g.Cout, T));                               \
        const int fin_smem = T * (T + 1) * (int)sizeof(float);                                                          \
        HEBB_CUDA_TRY(cudaFuncSetAttribute(tc_finalize_tiled_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                           fin_smem));                                                                  \
        tc_finalize_tiled_kernel<T><<<fg, 256, fin_smem, st>>>(hpart, rs, W, delta_w, n_part, g.taps, g.Cin,             \
                                                               cin_p, g.Cout);                                          \
      } while (0)
      if (n >= (1LL << 22)) HEBB_FIN_LAUNCH(128); else if (n >= (1LL << 20)) HEBB_FIN_LAUNCH(64); else HEBB_FIN_LAUNCH(32);
#undef HEBB_FIN_LAUNCH
    } else {
      long long gx = cdiv(n, 32);
      const long long cap = (long long)num_sms() * 32;
      if (gx > cap) gx = cap;
      tc_finalize_kernel<<<(unsigned)gx, 256, 0, st>>>(hpart, (wgrad || hpca) ? nullptr : rsum, W, delta_w, n_part, g.taps, g.Cin, cin_p, g.Cout);
    }
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  }
  if (hpca) {
    // G = y y^T on the same contraction kernel (both operands are the packed responses), then the
    // triangular decay  delta_w -= tril(G) W  as a small fp32 GEMM
    Geo g2 = gram_geo(g, P.PR);
    Plan P2;
    if (!plan_layer(g2, prec, &P2, 0, true) || P2.PR != P.PR) return HEBB_ESHAPE;
    float* G = reinterpret_cast<float*>(base + P.o_gram);
    float* hpart2 = reinterpret_cast<float*>(base + P.o_hpart2);
    HEBB_CUDA_TRY(cudaMemsetAsync(G, 0, sizeof(float) * (size_t)g.Cout * g.Cout, st));
    HEBB_TRY(launch_dw(P2, g2, rp0, rp1, rp0, rp1, hpart2, wd, P.PRS, P.PRS, st));
    long long gx = cdiv((long long)g.Cout * g.Cout, 32);
    const long long cap = (long long)num_sms() * 32;
    if (gx > cap) gx = cap;
    tc_finalize_kernel<<<(unsigned)gx, 256, 0, st>>>(hpart2, nullptr, hpart2, G, P2.PS * P2.Q, 1, g.Cout, P2.CinP, g.Cout);
    HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
    HEBB_TRY(launch_hpca_decay(G, W, delta_w, g.Cout, g.K, st));
  }
  return HEBB_OK;
}

// -------------------------------------------------------------------------------------
// Descriptor probe (tests/test_umma_probe.py): one CTA, raw images, raw descriptors.
// -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint8_t* __restrict__ a_img, int a_bytes, const uint8_t* __restrict__ b_img, int b_bytes,
                  uint64_t a_hi, uint32_t a_start, uint32_t a_step, uint64_t b_hi, uint32_t b_start,
                  uint32_t b_step, uint32_t idesc, int ksteps, int m, int n, float* __restrict__ d_out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // both images start on 1024-byte boundaries of the shared-memory window (the swizzled layouts repeat every 1024 B)
  uint8_t* const base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
  const uint32_t sa = smem_u32(base);
  const uint32_t b_off = (uint32_t)((a_bytes + 1023) / 1024 * 1024);
  for (int i = threadIdx.x; i < a_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(base)[i] = reinterpret_cast<const uint4*>(a_img)[i];
  for (int i = threadIdx.x; i < b_bytes / 16; i += blockDim.x)
    reinterpret_cast<uint4*>(base + b_off)[i] = reinterpret_cast<const uint4*>(b_img)[i];
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  const uint32_t cols = pow2_cols(n);
  if (warp == 0) { tmem_alloc(smem_u32(&s_tmem), cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = s_tmem;
  if (threadIdx.x == 0) {
    for (int k = 0; k < ksteps; ++k)
      umma_bf16(tb, smem_desc(a_hi, sa + a_start + k * a_step), smem_desc(b_hi, sa + b_off + b_start + k * b_step),
                idesc, k ? 1u : 0u);
    umma_commit(smem_u32(&bar));
  }
  mbar_wait(smem_u32(&bar), 0, nullptr, 0);
  tc_fence_after();
  // dump all 128 lanes x n columns; the host maps lanes to rows (M=64 uses lanes (r%16)+32*(r/16))
  const uint32_t ta = tb + (static_cast<uint32_t>(warp * 32) << 16);
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(ta + c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (c0 + i < n) d_out[(long long)(warp * 32 + lane) * n + c0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, cols);
  (void)m;
}


// MMA issue-rate microbenchmark: every CTA issues `iters` rounds of `per_round` tcgen05.mma from operands
// already resident in (uninitialised) shared memory and reports elapsed SM cycles of the whole sequence.
__global__ void __launch_bounds__(128, 1)
umma_rate_kernel(uint64_t a_hi, uint32_t a_step, uint64_t b_hi, uint32_t b_step, uint32_t b_off, uint32_t idesc,
                 int per_round, int iters, int n, long long* __restrict__ cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  const uint32_t sa = smem_u32(smem);
  // finite contents (zeros) so no NaN handling paths are exercised
  for (int i = threadIdx.x; i < (int)(b_off * 2 / 16); i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  const uint32_t cols = pow2_cols(n);
  if (warp == 0) { tmem_alloc(smem_u32(&s_tmem), cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = s_tmem;
  if (__shfl_sync(0xffffffffu, warp, 0) == 0) {
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        for (int k = 0; k < per_round; ++k)
          umma_bf16(tb, smem_desc(a_hi, sa + k * a_step), smem_desc(b_hi, sa + b_off + k * b_step), idesc, 1u);
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), ph, nullptr, 0);
      ph ^= 1;
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, cols);
}

// Same measurement with groups of `group` consecutive MMAs that share their A tile (the A descriptor advances
// once per group, the B descriptor per MMA) under the A-operand collector hints fill / use / lastuse
// (hint = 0 issues the same sequence without hints).
__global__ void __launch_bounds__(128, 1)
umma_rate_shared_a_kernel(uint64_t a_hi, uint32_t a_step, uint64_t b_hi, uint32_t b_step, uint32_t b_off, uint32_t idesc,
                          int per_round, int group, int hint, int d_step, int iters, int n,
                          long long* __restrict__ cycles) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t s_tmem;
  const int warp = threadIdx.x >> 5;
  const uint32_t sa = smem_u32(smem);
  for (int i = threadIdx.x; i < (int)(b_off * 2 / 16); i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
  const uint32_t cols = pow2_cols(d_step ? d_step * group : n);
  if (warp == 0) { tmem_alloc(smem_u32(&s_tmem), cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = s_tmem;
  if (__shfl_sync(0xffffffffu, warp, 0) == 0) {
    const uint32_t a_lbo = (uint32_t)a_hi, a_hi32 = (uint32_t)(a_hi >> 32);
    const uint32_t b_lbo = (uint32_t)b_hi, b_hi32 = (uint32_t)(b_hi >> 32);
    const long long t0 = clock64();
    uint32_t ph = 0;
    for (int it = 0; it < iters; ++it) {
      if (elect_one()) {
        uint32_t al = a_lbo | (sa >> 4), bl = b_lbo | ((sa + b_off) >> 4);
        for (int k = 0; k < per_round; k += group, al += a_step >> 4) {
          uint32_t d = tb;
          if (!hint) {
            for (int j = 0; j < group; ++j, bl += b_step >> 4, d += d_step) umma_lo<0, 1>(d, al, a_hi32, bl, b_hi32, idesc);
          } else {
            umma_lo<1, 1>(d, al, a_hi32, bl, b_hi32, idesc); bl += b_step >> 4; d += d_step;
            for (int j = 1; j < group - 1; ++j, bl += b_step >> 4, d += d_step) umma_lo<2, 1>(d, al, a_hi32, bl, b_hi32, idesc);
            if (hint == 1) umma_lo<3, 1>(d, al, a_hi32, bl, b_hi32, idesc);
            else umma_lo<2, 1>(d, al, a_hi32, bl, b_hi32, idesc);
            bl += b_step >> 4;
          }
        }
        umma_commit(smem_u32(&bar));
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), ph, nullptr, 0);
      ph ^= 1;
    }
    if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, cols);
}

}  // namespace hebb

extern "C" int hebb_debug_umma_probe(const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                                     uint64_t a_desc_hi, uint32_t a_start, uint32_t a_step,
                                     uint64_t b_desc_hi, uint32_t b_start, uint32_t b_step, uint32_t idesc,
                                     int ksteps, int m, int n, float* d_out, void* stream) {
  using namespace hebb;
  HEBB_TRY(device_ok());
  if (!a_img || !b_img || !d_out) return HEBB_EARG;
  if (a_bytes % 16 || b_bytes % 16 || n % 8 || n < 8 || n > 256 || (m != 64 && m != 128)) return HEBB_ESHAPE;
  const size_t smem = (size_t)((a_bytes + 1023) / 1024 * 1024) + b_bytes + 2048;
  if (smem > (size_t)kSmemLimit) return HEBB_ESHAPE;
  HEBB_CUDA_TRY(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(
      static_cast<const uint8_t*>(a_img), a_bytes, static_cast<const uint8_t*>(b_img), b_bytes, a_desc_hi, a_start,
      a_step, b_desc_hi, b_start, b_step, idesc, ksteps, m, n, d_out);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

extern "C" int hebb_debug_umma_rate(uint64_t a_desc_hi, uint32_t a_step, uint64_t b_desc_hi, uint32_t b_step,
                                    uint32_t region_bytes, uint32_t idesc, int per_round, int iters, int n, int ctas,
                                    long long* cycles, void* stream) {
  using namespace hebb;
  HEBB_TRY(device_ok());
  if (!cycles || region_bytes % 1024 || region_bytes * 2 + 1024 > (uint32_t)kSmemLimit) return HEBB_EARG;
  const size_t smem = (size_t)region_bytes * 2 + 1024;
  HEBB_CUDA_TRY(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_rate_kernel<<<ctas, 128, smem, (cudaStream_t)stream>>>(a_desc_hi, a_step, b_desc_hi, b_step, region_bytes, idesc,
                                                             per_round, iters, n, cycles);
  HEBB_CUDA_TRY(cudaGetLastError());
  return HEBB_OK;
}

extern "C" int hebb_debug_umma_rate_shared_a(uint64_t a_desc_hi, uint32_t a_step, uint64_t b_desc_hi, uint32_t b_step,
                                             uint32_t region_bytes, uint32_t idesc, int per_round, int group, int hint,
                                             int d_step, int iters, int n, int ctas, long long* cycles, void* stream) {
  using namespace hebb;
  HEBB_TRY(device_ok());
  if (!cycles || region_bytes % 1024 || region_bytes * 2 + 1024 > (uint32_t)kSmemLimit || group < 2 || per_round % group ||
      d_step < 0 || d_step * group > 512)
    return HEBB_EARG;
  const size_t smem = (size_t)region_bytes * 2 + 1024;
  HEBB_CUDA_TRY(cudaFuncSetAttribute(umma_rate_shared_a_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_rate_shared_a_kernel<<<ctas, 128, smem, (cudaStream_t)stream>>>(a_desc_hi, a_step, b_desc_hi, b_step, region_bytes,
                                                                      idesc, per_round, group, hint, d_step, iters, n, cycles);
  HEBB_CUDA_TRY(cudaGetLastError());
  return HEBB_OK;
}
