// Fused forward + plasticity-update kernel for the small-channel 2-D layers (Cin, Cout in {16, 32}; 1x1 or 3x3,
// stride 1) -- hebb/hebb.py:87-115 as ONE kernel (SURVEY.md 8b `hebb_fwd_dw_fused`).
//
// These layers are HBM-bound (48 % of the 2-D network's flops, 82 % of its bytes).  The
// two-kernel path moves every activation five times (pack: read fp32 + write bf16 hi/lo; forward: read packed x,
// write y and the packed responses; update: read packed x and r).  Here x is read ONCE as fp32 NCHW by TMA tensor-map
// loads (out-of-bounds zero fill materialises the zero halo of hebb.py:83-85), converted to bf16 hi/lo in shared
// memory, y is written once, and the responses r = softmax_c(k y) never leave the SM.
//
//   persistent CTA (one per SM), work item = (image b, TH x TW output tile); per CTA (NCW converter warps):
//   warp 0        TMA producer: per padded tile row and 16-channel group one 3-D box  [16 ch][1 row][pitch] fp32
//                 (weight-gradient mode on channels_last tensors: 4-D boxes [pitch pixels][16 ch])
//   warp 1        tcgen05.mma issuer of the forward, per 128-position block:
//                   D[pos, co] += x[pos + tap][ci] W[tap][ci][co]: A = the x image read K-major (a tap is a start-address
//                   offset of whole rows, hi/lo halves are 32-byte offsets), B = packed weights [w_hi | w_lo] stacked along
//                   N for x_hi, w_hi for x_lo: 2 instructions per tap and 16-channel slab into one fp32 accumulator
//   warp 2        tcgen05.mma issuer of the update, lagging until a block's responses are through the epilogue:
//                   H[(kh, hl, ci), (kw, hl', co)] += x[q + kh*pitch] r[q - kw]: the SAME x image read MN-major with an
//                   atom stride of one tile row (M = 128 rows = kh copies x [hi; lo] x Cin), the response image read with
//                   an atom stride of ONE position (N = kW copies x [hi | lo] x Cout): ONE instruction per 16 positions
//                   covers all taps and all four hi/lo products (tests/test_umma_probe.py pins these descriptor forms)
//   warps 3..2+NCW  converter: fp32 staging -> x image [pos][hi Cin | lo Cin] bf16, 16-byte chunks XOR-swizzled on absolute
//                 address bits (SWIZZLE_64B for Cin = 16, SWIZZLE_128B for Cin = 32); a warp owns whole stages in turn
//                 (NCW = 2; 4 for the patch gather of the 3-channel first layer; 6 in weight-gradient mode)
//   last 8 warps  epilogue (two sets alternate blocks): TMEM -> y = acc/|w| + b -> global; winner (+ near-tie list);
//                 BatchNorm sums; r = softmax -> bf16 hi/lo -> response image in shared memory; running sums of r.
//                 Weight-gradient mode (template flag WG, hebb_conv_wgrad): no forward, no softmax -- these warps read
//                 dL/dy[b, co, pixel] from global memory and split it into the response ring instead.
//   The update accumulators stay in TMEM for the whole kernel; each CTA writes ONE partial [taps][Cin][Cout] (x2 for the
//   hi/lo rows of x), summed in a fixed order by tc_finalize_kernel together with the decay term -(sum_p r) W
//   (fused_wgrad_finalize_kernel in weight-gradient mode: no decay, += into grad_w).
#include "common.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cstdlib>

namespace hebb {

using namespace ptx;

namespace {

constexpr int kSmemLimitF = 227 * 1024;
constexpr int kRLead = 8;                 // zero / carried-over positions in front of every response slot
constexpr int kRSlotPos = kRLead + 128;
constexpr int kMaxRows = 18;              // x tile rows (TH + kH - 1) a CTA keeps barriers for
constexpr int kMaxStages = 10;            // fp32 staging ring (TMA boxes in flight)
constexpr int kDwLag = 2;                 // the update MMAs of a block are issued this many blocks after its forward MMAs
constexpr int kRSlots = kDwLag + 2;       // response ring: at most this many slots of kRSlotPos positions (p.nrs are in use)
constexpr int kNumBars = 2 * kMaxStages + 2 * kMaxRows + 2 + 2 + 2 * kRSlots + 2;
// warps: 0 TMA producer, 1 forward issuer, 2 update issuer, 3 .. 2+NCW converter, then the two epilogue sets of four.
// NCW = 2 converter warps for the plain layers (13 warps: the issuing warps then share their scheduler with epilogue
// warps only -- with a converter warp on every scheduler the same kernel ran 25 % slower), 4 for the patch gather, 6 in
// weight-gradient mode (no forward issuer to disturb; the converter is the critical path there).

struct FusedParams {
  float* y; int32_t* winner; const float* inv; const float* bias; float* rsum; double* ystats; float* hpart; int* err;
  int* fix_list; int* fix_count; int fix_cap; float tie_rel;
  const uint4* wp;
  int B, oH, oW, kH, kW, pH, pW, taps;
  int TH, TW, pitch, nTH, nTW, ntiles, XROWS, NBLK, XPOS, NST;
  unsigned pitch_magic;        // ceil(2^32 / pitch): q / pitch == umulhi(q, pitch_magic) for the q < 2^16 of a tile
  int nrs;                     // response ring slots in use (kRSlots, or one fewer where shared memory is short)
  int gather, gcin, gk, srows; // few-input-channel layers: the converter gathers the gk x gk patch of the gcin real channels into
                               // gcin*gk*gk pseudo-channels of a 1x1 layer (srows = TH + gk - 1 staged input rows per tile)
  int BW, padl;                // TMA box width (floats) and left pad of the box start: staging column = tile column + padl - pW
  float kinv; int update;
  // weight-gradient mode (hebb_conv_wgrad on this kernel): dL/dy takes the place of the responses -- no forward MMAs, no
  // softmax; the "epilogue" warps read gy[b][co][pixel] (element strides gy_sb / gy_sc / gy_sp) and feed the response ring
  const float* gy; long long gy_sb, gy_sc, gy_sp;
  int gy_nc;                   // channels of this pass present in gy (< COUT: a 2-class layer padded to 16 filters; the rest read 0)
  int xcl;                     // x is channels_last: boxes [BW pixels][16 channels] through a 4-D tensor map
  int cin_tot, ci_off;         // channels of the tensor behind the map, first channel of this launch (channel passes)
  long long* prof;             // HEBB_FUSED_PROF=1: per CTA [32] cycles spent in each bounded wait (index = code - 16) + totals
  int dbg;                     // HEBB_FUSED_DBG (profiling only): 1 one forward MMA per block, 2 no update MMAs, 4 no epilogue math,
                               // 8 no conversion, 16 no y stores
  uint32_t off_r, off_stage, off_w, off_misc, w_bytes, stage_bytes, tmem_cols;
};

__device__ __forceinline__ void flag_tie_f(int* list, int* count, int cap, long long pid) {
  const int i = atomicAdd(count, 1);       // near-tie worklist, see fixup.cu
  if (i < cap) list[i] = (int)pid;
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

__device__ __forceinline__ void st_shared_v2(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// tcgen05.mma with a run-time accumulate flag (the update accumulators are overwritten exactly once per kernel)
__device__ __forceinline__ void umma_rt(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                        uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n"
      ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}

// byte address of 16-byte chunk `c` of the position whose (unswizzled) row starts at `row_addr`; NCH chunks per row.
// The hardware XORs address bits [4, 4+log2 NCH) with the bits from 7 up (Swizzle<b,4,3> on absolute addresses).
template <int NCH>
__device__ __forceinline__ uint32_t swz(uint32_t row_addr, int c) {
  return row_addr + ((((uint32_t)c) ^ ((row_addr >> 7) & (NCH - 1))) << 4);
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// bounded wait; with PROF the cycles spent in it are accumulated per wait code (20..28)
#define FWAIT(bar, par, code)                                                        \
  do {                                                                               \
    if (PROF) {                                                                      \
      const long long _t = clock64();                                                \
      mbar_wait(bar, par, p.err, code);                                              \
      prof_acc[(code) - 20] += clock64() - _t;                                       \
    } else {                                                                         \
      mbar_wait(bar, par, p.err, code);                                              \
    }                                                                                \
  } while (0)

template <int CIN, int COUT, int KS, bool PROF, int NCW, bool WG = false>
__global__ void __launch_bounds__(32 * (3 + NCW + 8), 1)
fused_small_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ FusedParams p) {
  constexpr int XB = CIN * 4;              // bytes per position of the x image: [hi Cin | lo Cin] bf16
  constexpr int RB = COUT * 4;             // ... of the response image
  constexpr int XCH = XB / 16, RCH = RB / 16;
  constexpr uint32_t XLAY = (XB == 128) ? 2u : 4u, RLAY = (RB == 128) ? 2u : 4u;     // SWIZZLE_128B : SWIZZLE_64B
  constexpr int NSL = CIN / 16;            // 16-channel slabs (K steps of the forward)
  constexpr int COPIES = 128 / (2 * CIN);  // kh copies of the x tile in the M = 128 rows of an update instruction
  constexpr int FCOLS = 2 * COUT;          // forward accumulator: [x w_hi (+ x_lo w_hi) | x_hi w_lo]

  extern __shared__ __align__(128) uint8_t smem_raw[];
  uint8_t* const smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  long long prof_acc[10];
#pragma unroll
  for (int i = 0; i < 10; ++i) prof_acc[i] = 0;
  const long long prof_t0 = PROF ? clock64() : 0;

  float* s_inv = reinterpret_cast<float*>(smem + p.off_misc);
  float* s_bias = s_inv + COUT;
  float* s_rs = s_bias + COUT;               // [8 warps][COUT]
  float* s_ys = s_rs + 8 * COUT;
  float* s_yq = s_ys + COUT;
  float* s_tab = s_yq + COUT;                 // 2 x 32 ints: per-block row tables of the issuing warp
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_tab + 64);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t st_full = bar0, st_empty = st_full + 8 * kMaxStages;
  const uint32_t xr_full = st_empty + 8 * kMaxStages, xr_empty = xr_full + 8 * kMaxRows;
  const uint32_t tf_full = xr_empty + 8 * kMaxRows, tf_empty = tf_full + 16;
  const uint32_t r_full = tf_empty + 16, r_empty = r_full + 8 * kRSlots, w_full = r_empty + 8 * kRSlots, done = w_full + 8;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kNumBars);

  const int NACC = (p.kH + COPIES - 1) / COPIES;          // update accumulators actually used (at most 2)
  // accumulator 1 re-reads the last COPIES kernel rows (kh = kH-COPIES ..), so that no copy reaches past kh = kH-1
  const int kh_base1 = p.kH - COPIES;
  const int NW = p.kW * 2 * COUT;                          // columns of one update accumulator
  const uint32_t xb = sbase;
  const uint32_t rb = sbase + p.off_r;

  // ---- one-time set-up: constants, zero fill of what the tensor pipe may read before anyone wrote it ----
  for (int i = threadIdx.x; i < COUT; i += blockDim.x) {
    s_inv[i] = p.inv ? p.inv[i] : 1.f;
    s_bias[i] = p.bias ? p.bias[i] : 0.f;
  }
  for (int i = threadIdx.x; i < 10 * COUT; i += blockDim.x) s_rs[i] = 0.f;     // rs, ys, yq
  if ((int)threadIdx.x < p.NBLK && threadIdx.x < 32) {
    const int kb = threadIdx.x;
    int need = ((kb + 1) * 128 - 1 + (p.kH - 1) * p.pitch + (p.kW - 1)) / p.pitch + 1;
    if (need > p.XROWS) need = p.XROWS;
    int free_to = (kb == p.NBLK - 1) ? p.XROWS : ((kb + 1) * 128) / p.pitch;
    if (free_to > p.XROWS) free_to = p.XROWS;
    reinterpret_cast<int*>(s_tab)[kb] = need;
    reinterpret_cast<int*>(s_tab)[32 + kb] = free_to;
  }
  {
    // x image beyond the converted rows (read by the wasted kh copies / the last block) and the response ring
    const uint32_t x_tail = (uint32_t)p.XROWS * p.pitch * XB, x_end = (uint32_t)p.XPOS * XB;
    for (uint32_t a = x_tail + threadIdx.x * 16; a < x_end; a += blockDim.x * 16) st_shared_v4(xb + a, 0, 0, 0, 0);
    const uint32_t r_end = (uint32_t)p.nrs * kRSlotPos * RB;
    for (uint32_t a = threadIdx.x * 16; a < r_end; a += blockDim.x * 16) st_shared_v4(rb + a, 0, 0, 0, 0);
  }
  fence_proxy_async();
  if (threadIdx.x == 0) {
    // (one converter warp per stage, see the converter; the patch gather releases a staged input row three times)
    for (int i = 0; i < p.NST; ++i) { mbar_init(st_full + 8 * i, 1); mbar_init(st_empty + 8 * i, p.gather ? 3 : 1); }
    for (int i = 0; i < kMaxRows; ++i) { mbar_init(xr_full + 8 * i, NSL); mbar_init(xr_empty + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(tf_full + 8 * i, 1); mbar_init(tf_empty + 8 * i, 4); }
    for (int i = 0; i < kRSlots; ++i) { mbar_init(r_full + 8 * i, 4); mbar_init(r_empty + 8 * i, 1); }
    mbar_init(w_full, 1); mbar_init(done, 1);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(smem_u32(s_tmem), p.tmem_cols); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_f = tmem_base;                      // 2 forward accumulators of FCOLS columns
  const uint32_t tmem_d = tmem_base + 2 * FCOLS;          // NACC update accumulators of NW columns
  const int per_img = p.nTH * p.nTW;
  const int my_tiles = ((int)blockIdx.x < p.ntiles) ? (p.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      if (!WG) {
        mbar_expect_tx(w_full, p.w_bytes);
        bulk_g2s(sbase + p.off_w, p.wp, p.w_bytes, w_full);
      }
      int s = 0; uint32_t ph = 0;
      const uint32_t row_bytes = (uint32_t)p.BW * 16 * 4;
      for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        const int b = tile / per_img, rem = tile - b * per_img;
        const int th = rem / p.nTW, tw = rem - th * p.nTW;
        const int h0 = th * p.TH - p.pH, w0 = tw * p.TW - p.padl;      // box start: a multiple of 4 floats (16 bytes)
        if (p.gather) {
          for (int r = 0; r < p.srows; ++r) {          // one box per input row: [gcin channels][1 row][BW]
            FWAIT(st_empty + 8 * s, ph ^ 1, 21);
            mbar_expect_tx(st_full + 8 * s, (uint32_t)p.BW * p.gcin * 4);
            tma_load_3d(sbase + p.off_stage + s * p.stage_bytes, &tmap, w0, h0 + r, b * p.gcin, st_full + 8 * s);
            if (++s == p.NST) { s = 0; ph ^= 1; }
          }
          continue;
        }
        for (int r = 0; r < p.XROWS; ++r)
          for (int cg = 0; cg < NSL; ++cg) {
            FWAIT(st_empty + 8 * s, ph ^ 1, 21);
            mbar_expect_tx(st_full + 8 * s, row_bytes);
            if (WG && p.xcl) tma_load_4d(sbase + p.off_stage + s * p.stage_bytes, &tmap, p.ci_off + cg * 16, w0, h0 + r, b, st_full + 8 * s);
            else if (WG) tma_load_3d(sbase + p.off_stage + s * p.stage_bytes, &tmap, w0, h0 + r, b * p.cin_tot + p.ci_off + cg * 16, st_full + 8 * s);
            else tma_load_3d(sbase + p.off_stage + s * p.stage_bytes, &tmap, w0, h0 + r, b * CIN + cg * 16, st_full + 8 * s);
            if (++s == p.NST) { s = 0; ph ^= 1; }
          }
      }
    }
  } else if (warp == 1 || warp == 2) {
    // ===================== MMA issuers: warp 1 forward, warp 2 update =====================
    // The issuing thread is the critical resource of this kernel (26-50 instructions per 128 positions, each worth
    // ~45 cycles of tensor-pipe time): everything it needs per instruction is a compile-time constant or sits in a
    // register before the loop -- tap offsets, weight descriptors, per-block row counts -- so that an issue is one or
    // two integer adds plus the tcgen05.mma itself, and the loop nest is warp-uniform (descriptors in uniform registers).
    constexpr int TAPS = KS * KS;
    const uint32_t idesc_f1 = idesc_bf16(128, 2 * COUT, 0, 0);     // x_hi * [w_hi | w_lo]
    const uint32_t idesc_f2 = idesc_bf16(128, COUT, 0, 0);         // x_lo * w_hi
    const uint32_t idesc_d = idesc_bf16(128, KS * 2 * COUT, 1, 1);
    // forward A: K-major swizzled rows of XB bytes, 8-row groups 8*XB apart (LBO unused: K = 32 bytes < row)
    const uint32_t fa_hi = ((8u * XB) >> 4) | (1u << 14) | (XLAY << 29);
    // forward B: SWIZZLE_NONE K-major packed weights [k-chunk][hi|lo][COUT rows][16 B]: for one k-chunk the w_hi rows are
    // directly followed by the w_lo rows, so N = 2 COUT rows from the w_hi start are [w_hi | w_lo]
    const uint32_t fb_hi = (128u >> 4) | (1u << 14);
    const uint32_t fb_base = (((2u * COUT * 16u) >> 4) << 16) | (((sbase + p.off_w) >> 4) & 0x3FFFu);
    // update A: MN-major swizzled, atoms (kh copies) one tile row apart, 8-position groups 8*XB apart
    const uint32_t da_hi = ((8u * XB) >> 4) | (1u << 14) | (XLAY << 29);
    const uint32_t da_base = ((((uint32_t)p.pitch * XB) >> 4) << 16) | ((xb >> 4) & 0x3FFFu);
    // update B: MN-major swizzled response image, atoms (kw copies) ONE position apart
    const uint32_t db_hi = ((8u * RB) >> 4) | (1u << 14) | (RLAY << 29);
    const uint32_t db_base = (((uint32_t)RB >> 4) << 16) | (((rb + (uint32_t)(kRLead - (KS - 1)) * RB) >> 4) & 0x3FFFu);
    const uint32_t fa_base = (xb >> 4) & 0x3FFFu;
    uint32_t toff[TAPS];                      // tap -> start offset of the A operand, in 16-byte units
#pragma unroll
    for (int t = 0; t < TAPS; ++t) toff[t] = (uint32_t)(((t / KS) * p.pitch + (t % KS)) * XB) >> 4;
    const uint32_t a2 = (uint32_t)(kh_base1 * p.pitch * XB) >> 4;    // second update accumulator: kh copies kH-COPIES ..
    const int* s_need = reinterpret_cast<const int*>(s_tab);          // rows the forward of block k needs converted
    const int* s_free = s_need + 32;                                  // rows free once the update of block k is done
    const int total = my_tiles * p.NBLK;
    if (warp == 1) {
      // ---------- forward issuer: block g as soon as its x rows are converted and its TMEM buffer has been drained ----------
      if (!WG) {
      FWAIT(w_full, 0, 20);
      int rows_ready = 0, rows_freed = 0;
      int k = 0, it = 0;                        // block within the tile, tile count
      for (int g = 0; g < total; ++g) {
        if (k == 0) rows_ready = 0;
        const int need = s_need[k];
        for (; rows_ready < need; ++rows_ready) FWAIT(xr_full + 8 * rows_ready, it & 1, 22);
        const uint32_t acc = (uint32_t)g & 1u;
        FWAIT(tf_empty + 8 * acc, (((uint32_t)g >> 1) & 1u) ^ 1u, 23);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem_f + acc * FCOLS;
          const uint32_t ablk = fa_base + (uint32_t)k * (128u * XB >> 4);
          const int ntap = (p.dbg & 1) ? 1 : TAPS;
#pragma unroll
          for (int t = 0; t < TAPS; ++t) {
            if (t < ntap) {
#pragma unroll
              for (int s = 0; s < NSL; ++s) {
                const uint32_t ah = ablk + toff[t] + 2u * s, al = ah + (2u * CIN >> 4);
                const uint32_t bh = fb_base + (uint32_t)((s * TAPS + t) * 4 * COUT);
                if (t == 0 && s == 0) umma_lo<0, 0>(d, ah, fa_hi, bh, fb_hi, idesc_f1);
                else umma_lo<0, 1>(d, ah, fa_hi, bh, fb_hi, idesc_f1);
                umma_lo<0, 1>(d, al, fa_hi, bh, fb_hi, idesc_f2);
              }
            }
          }
          umma_commit(tf_full + 8 * acc);
        }
        __syncwarp();
        if (!p.update) {
          // forward only: rows are free once the blocks that read them have been issued
          if (k == 0) rows_freed = 0;
          const int free_to = s_free[k];
          if (elect_one())
            for (int r = rows_freed; r < free_to; ++r) umma_commit(xr_empty + 8 * r);
          __syncwarp();
          rows_freed = free_to > rows_freed ? free_to : rows_freed;
        }
        if (++k == p.NBLK) { k = 0; ++it; }
      }
      }
    } else if (p.update) {
      // ---------- update issuer: block g once its responses are in the ring (i.e. after its forward has completed and
      // been through the epilogue).  It runs on its own warp so that its waits overlap the forward issue; the tensor
      // pipe takes both instruction streams.  A tcgen05.commit only tracks the MMAs of its own thread: the x rows it
      // releases are last read by forward blocks <= g, which have completed by the time r_full(g) is set. ----------
      int rows_freed = 0;
      int jd = 0; uint32_t slot_d = 0, ph_d = 0;   // block within its tile, response slot, its phase
      int rows_ready = 0, itd = 0;                  // weight-gradient mode: nobody in front of this warp waited for the x rows
      bool d_first = true;
      for (int g = 0; g < total; ++g) {
        if (jd == 0) { rows_freed = 0; rows_ready = 0; }
        if (WG) {
          const int need = s_need[jd];
          for (; rows_ready < need; ++rows_ready) FWAIT(xr_full + 8 * rows_ready, itd & 1, 22);
        }
        FWAIT(r_full + 8 * slot_d, ph_d, 24);
        tc_fence_after();
        const int free_to = s_free[jd];
        if (elect_one()) {
          uint32_t a = da_base + (uint32_t)jd * (128u * XB >> 4);
          uint32_t b = db_base + slot_d * ((uint32_t)(kRSlotPos * RB) >> 4);
          const uint32_t nks = (p.dbg & 2) ? 0u : 8u;
          if (nks) {
            if (d_first) {
              umma_lo<0, 0>(tmem_d, a, da_hi, b, db_hi, idesc_d);
              if (NACC > 1) umma_lo<0, 0>(tmem_d + NW, a + a2, da_hi, b, db_hi, idesc_d);
            } else {
              umma_lo<0, 1>(tmem_d, a, da_hi, b, db_hi, idesc_d);
              if (NACC > 1) umma_lo<0, 1>(tmem_d + NW, a + a2, da_hi, b, db_hi, idesc_d);
            }
#pragma unroll
            for (int ks = 1; ks < 8; ++ks) {                  // 16 positions per instruction
              a += XB; b += RB;
              umma_lo<0, 1>(tmem_d, a, da_hi, b, db_hi, idesc_d);
              if (NACC > 1) umma_lo<0, 1>(tmem_d + NW, a + a2, da_hi, b, db_hi, idesc_d);
            }
          }
          umma_commit(r_empty + 8 * slot_d);             // the response slot (and the lead it shares) may be rewritten
          for (int r = rows_freed; r < free_to; ++r) umma_commit(xr_empty + 8 * r);
        }
        __syncwarp();
        d_first = false;
        rows_freed = free_to > rows_freed ? free_to : rows_freed;
        if (++jd == p.NBLK) { jd = 0; ++itd; }
        if (++slot_d == (uint32_t)p.nrs) { slot_d = 0; ph_d ^= 1u; }
      }
      if (elect_one()) umma_commit(done);
      __syncwarp();
    }
  } else if (warp < 3 + NCW) {
    // ===================== converter (warps 3 .. 2+NCW) =====================
    const int conv_ntiles = p.ntiles;
    int s = 0; uint32_t ph = 0;
    int it = 0;
    constexpr int CSTEP = 32;
    const int t0 = lane;
    int own = 0;
    if constexpr (CIN == 32 && KS == 1 && NCW == 4) {
      if (p.gather) {
        // ---- patch gather (first layers: 1-4 real channels, gk x gk kernel): image row r of the 1x1 layer holds, per
        // output pixel, the gcin*gk*gk patch values (zero-padded to 32 pseudo-channels) as [hi 32 | lo 32] ----
        // A converter warp owns whole image rows in turn (global row n belongs to warp n mod NCW): the per-row fixed cost
        // (barrier polls, proxy fence, arrives) is paid by the four warps in parallel instead of by each of them for a
        // quarter of the pixels.  Staged input row j serves image rows j-2 .. j: its st_empty barrier counts 3 arrivals,
        // and the rows at a tile's edge make up for the neighbours they do not have.
        const int cw = warp - 3;
        const int off = p.padl - p.pW;
        int grow = 0;
        int cons = 0, cs = 0;                // staged input rows observed so far, and the ring slot of the next one
        uint32_t cph = 0;
        int s0 = 0;                          // ring slot of the first input row under the current image row
        for (int tile = blockIdx.x; tile < conv_ntiles; tile += gridDim.x, ++it) {
          const int base = it * p.srows;
          for (int r = 0; r < p.XROWS; ++r, ++grow) {
            // every warp observes every barrier phase in order (see the plain converter below); the owner converts
            for (; cons <= base + r + 2; ++cons) {
              FWAIT(st_full + 8 * cs, cph, 25);
              if (++cs == p.NST) { cs = 0; cph ^= 1u; }
            }
            FWAIT(xr_empty + 8 * r, (it & 1) ^ 1, 26);
            int sl[3];
            sl[0] = s0; sl[1] = (s0 + 1 == p.NST) ? 0 : s0 + 1; sl[2] = (sl[1] + 1 == p.NST) ? 0 : sl[1] + 1;
            s0 = sl[1];
            if ((grow & (NCW - 1)) != cw) continue;
            const float* row0 = reinterpret_cast<const float*>(smem + p.off_stage + sl[0] * p.stage_bytes) + off;
            const float* row1 = reinterpret_cast<const float*>(smem + p.off_stage + sl[1] * p.stage_bytes) + off;
            const float* row2 = reinterpret_cast<const float*>(smem + p.off_stage + sl[2] * p.stage_bytes) + off;
            for (int c = lane; c < ((p.dbg & 8) ? 0 : p.pitch); c += 32) {
              float vv[32];                      // pseudo-channel ci*9 + kh*3 + kw; zero beyond the real channels
#pragma unroll
              for (int i = 27; i < 32; ++i) vv[i] = 0.f;
#pragma unroll
              for (int ci = 0; ci < 3; ++ci) {
                if (ci < p.gcin) {
                  const float* q0 = row0 + ci * p.BW + c;
                  const float* q1 = row1 + ci * p.BW + c;
                  const float* q2 = row2 + ci * p.BW + c;
#pragma unroll
                  for (int kw = 0; kw < 3; ++kw) { vv[ci * 9 + kw] = q0[kw]; vv[ci * 9 + 3 + kw] = q1[kw]; vv[ci * 9 + 6 + kw] = q2[kw]; }
                } else {
#pragma unroll
                  for (int i = 0; i < 9; ++i) vv[ci * 9 + i] = 0.f;
                }
              }
              uint32_t hp[16], lp[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp[i]) : "f"(vv[2 * i + 1]), "f"(vv[2 * i]));
                const float h0 = __uint_as_float(hp[i] << 16), h1 = __uint_as_float(hp[i] & 0xffff0000u);
                asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp[i]) : "f"(vv[2 * i + 1] - h1), "f"(vv[2 * i] - h0));
              }
              const uint32_t row = xb + (uint32_t)(r * p.pitch + c) * XB;
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                st_shared_v4(swz<XCH>(row, ch), hp[4 * ch], hp[4 * ch + 1], hp[4 * ch + 2], hp[4 * ch + 3]);
                st_shared_v4(swz<XCH>(row, 4 + ch), lp[4 * ch], lp[4 * ch + 1], lp[4 * ch + 2], lp[4 * ch + 3]);
              }
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              for (int a = 0; a < NSL; ++a) mbar_arrive(xr_full + 8 * r);
              const int top = r == 0 ? 1 : 0, bot = r == p.XROWS - 1 ? 1 : 0;
              for (int a = 0; a < 1 + 2 * top; ++a) mbar_arrive(st_empty + 8 * sl[0]);
              for (int a = 0; a < 1 + top + bot; ++a) mbar_arrive(st_empty + 8 * sl[1]);
              for (int a = 0; a < 1 + 2 * bot; ++a) mbar_arrive(st_empty + 8 * sl[2]);
            }
          }
          for (int j = 0; j < p.srows - p.XROWS; ++j) s0 = (s0 + 1 == p.NST) ? 0 : s0 + 1;      // the trailing gk-1 input rows
        }
        goto converter_done;
      }
    }
    // A converter warp owns WHOLE stages in turn (stage n belongs to warp n mod NCW) instead of a share of every stage's
    // pixels: the per-stage fixed cost (two barrier polls, the proxy fence, two arrives: ~300 of 450 cycles for a 68-pixel
    // row) is then paid by the warps in parallel.  In weight-gradient mode, where nothing but the update runs behind it,
    // the converter is the critical path (six warps: 2.31 -> 1.50 ms for the 2-D head's three layers); the forward +
    // update kernels gained 2-10 % (32->16 @256^2 0.415 -> 0.374 ms, 32->32 @128^2 0.194 -> 0.181, 16->32 0.105 -> 0.101)
    for (int tile = blockIdx.x; tile < conv_ntiles; tile += gridDim.x, ++it) {
      for (int r = 0; r < p.XROWS; ++r)
        for (int cg = 0; cg < NSL; ++cg) {
          // EVERY warp observes every barrier phase in order (a parity wait cannot tell a phase from the one two
          // completions earlier: a warp that skipped ahead to its own stage could pass on a stale phase); only the
          // conversion, the proxy fence and the arrives belong to the owner
          FWAIT(st_full + 8 * s, ph, 25);
          if (cg == 0) FWAIT(xr_empty + 8 * r, (it & 1) ^ 1, 26);      // the previous tile no longer reads this row
          {
            const bool mine = own == warp - 3;
            own = (own + 1 == NCW) ? 0 : own + 1;
            if (!mine) {
              if (++s == p.NST) { s = 0; ph ^= 1; }
              continue;
            }
          }
          const long long tc0 = PROF ? clock64() : 0;
          if (WG && p.xcl) {
            // channels_last box: [pixel][16 channels] fp32; four lanes share a pixel (one float4 = 4 channels each)
            const float4* st4 = reinterpret_cast<const float4*>(smem + p.off_stage + s * p.stage_bytes) + (p.padl - p.pW) * 4;
            // four pixels-quarters per lane and round: the loads first (the shared-memory stores are ordered asm
            // statements, so the compiler cannot overlap one item's load with the previous item's stores by itself)
            const int n_items = p.pitch * 4;
            for (int base = t0; base < n_items; base += 4 * CSTEP) {
              float4 v4[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int idx = base + u * CSTEP;
                v4[u] = idx < n_items ? st4[idx] : make_float4(0.f, 0.f, 0.f, 0.f);
              }
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const int idx = base + u * CSTEP;
                if (idx < n_items) {
                  const int c = idx >> 2, j = idx & 3;
                  const float4 v = v4[u];
                  uint32_t h0, h1, l0, l1;
                  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h0) : "f"(v.y), "f"(v.x));
                  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(v.w), "f"(v.z));
                  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l0) : "f"(v.y - __uint_as_float(h0 & 0xffff0000u)), "f"(v.x - __uint_as_float(h0 << 16)));
                  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l1) : "f"(v.w - __uint_as_float(h1 & 0xffff0000u)), "f"(v.z - __uint_as_float(h1 << 16)));
                  const uint32_t row = xb + (uint32_t)(r * p.pitch + c) * XB;
                  st_shared_v2(swz<XCH>(row, 2 * cg + (j >> 1)) + (uint32_t)(j & 1) * 8u, h0, h1);
                  st_shared_v2(swz<XCH>(row, XCH / 2 + 2 * cg + (j >> 1)) + (uint32_t)(j & 1) * 8u, l0, l1);
                }
              }
            }
          }
          const float* st = reinterpret_cast<const float*>(smem + p.off_stage + s * p.stage_bytes) + (p.padl - p.pW);
          if constexpr (WG) {
            if (!p.xcl) {
              // NCHW box, two pixels per lane and round (loads of both first, see above)
              for (int c = t0; c < p.pitch; c += 2 * CSTEP) {
                float v[2][16];
                const bool two = c + CSTEP < p.pitch;
#pragma unroll
                for (int i = 0; i < 16; ++i) { v[0][i] = st[i * p.BW + c]; v[1][i] = two ? st[i * p.BW + c + CSTEP] : 0.f; }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                  if (u == 1 && !two) break;
                  uint32_t hp[8], lp[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp[i]) : "f"(v[u][2 * i + 1]), "f"(v[u][2 * i]));
                    const float h0 = __uint_as_float(hp[i] << 16), h1 = __uint_as_float(hp[i] & 0xffff0000u);
                    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp[i]) : "f"(v[u][2 * i + 1] - h1), "f"(v[u][2 * i] - h0));
                  }
                  const uint32_t row = xb + (uint32_t)(r * p.pitch + c + u * CSTEP) * XB;
                  st_shared_v4(swz<XCH>(row, 2 * cg), hp[0], hp[1], hp[2], hp[3]);
                  st_shared_v4(swz<XCH>(row, 2 * cg + 1), hp[4], hp[5], hp[6], hp[7]);
                  st_shared_v4(swz<XCH>(row, XCH / 2 + 2 * cg), lp[0], lp[1], lp[2], lp[3]);
                  st_shared_v4(swz<XCH>(row, XCH / 2 + 2 * cg + 1), lp[4], lp[5], lp[6], lp[7]);
                }
              }
            }
          }
          for (int c = t0; c < (((p.dbg & 8) || WG) ? 0 : p.pitch); c += CSTEP) {
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = st[i * p.BW + c];
            uint32_t hp[8], lp[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
              const float h0 = __uint_as_float(hp[i] << 16), h1 = __uint_as_float(hp[i] & 0xffff0000u);
              asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp[i]) : "f"(v[2 * i + 1] - h1), "f"(v[2 * i] - h0));
            }
            const uint32_t row = xb + (uint32_t)(r * p.pitch + c) * XB;
            st_shared_v4(swz<XCH>(row, 2 * cg), hp[0], hp[1], hp[2], hp[3]);
            st_shared_v4(swz<XCH>(row, 2 * cg + 1), hp[4], hp[5], hp[6], hp[7]);
            st_shared_v4(swz<XCH>(row, XCH / 2 + 2 * cg), lp[0], lp[1], lp[2], lp[3]);
            st_shared_v4(swz<XCH>(row, XCH / 2 + 2 * cg + 1), lp[4], lp[5], lp[6], lp[7]);
          }
          const long long tc1 = PROF ? clock64() : 0;
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) { mbar_arrive(xr_full + 8 * r); mbar_arrive(st_empty + 8 * s); }
          if (PROF) { prof_acc[0] += tc1 - tc0; prof_acc[1] += clock64() - tc1; }      // [0] conversion, [1] fence + arrives
          if (++s == p.NST) { s = 0; ph ^= 1; }
        }
    }
  converter_done:;
  } else {
    // ===================== epilogue warps 3+NCW .. 10+NCW: two sets of four alternate blocks =====================
    const int quad = warp & 3;
    const int ew = warp - (3 + NCW);
    const int eset = ew >> 2;
    float* my_rs = s_rs + ew * COUT;
    const long long outS = (long long)p.oH * p.oW;
    // BatchNorm sums of y: 16-channel rows keep per-thread running sums, 32-channel rows fold every block with a
    // butterfly (96 more live registers would not fit next to the row itself)
    constexpr bool YACC = COUT <= 16;
    constexpr int NY = YACC ? COUT : 1;
    float racc[COUT], ysacc[NY], yqacc[NY];
#pragma unroll
    for (int i = 0; i < COUT; ++i) racc[i] = 0.f;
#pragma unroll
    for (int i = 0; i < NY; ++i) { ysacc[i] = 0.f; yqacc[i] = 0.f; }
    const bool want_ys = p.ystats != nullptr;
    const float k2 = p.kinv * 1.4426950408889634f;      // exp(k y) = 2^(k2 y)
    uint32_t kk = 0;
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      const int b = tile / per_img, rem = tile - b * per_img;
      const int th = rem / p.nTW, tw = rem - th * p.nTW;
      const int h0 = th * p.TH, w0 = tw * p.TW;
      const int THv = min(p.TH, p.oH - h0), TWv = min(p.TW, p.oW - w0);
      for (int k = 0; k < p.NBLK; ++k, ++kk) {
        if ((int)(kk & 1u) != eset) continue;
        const int loc = quad * 32 + lane;
        float f[COUT];
        float rinv;
        long long tq0 = PROF ? clock64() : 0;
        if constexpr (WG) {
          // weight-gradient mode: this pixel's dL/dy row takes the place of the responses
          const int q = k * 128 + loc;
          const int r_ = (int)__umulhi((unsigned)q, p.pitch_magic), c_ = q - r_ * p.pitch;      // q / pitch (q < 2^16)
          const bool valid = r_ < THv && c_ < TWv;
          const long long pix = (long long)(h0 + r_) * p.oW + (w0 + c_);
          if (valid) {
            const float* gp = p.gy + (long long)b * p.gy_sb + pix * p.gy_sp;
            if (p.gy_sc == 1 && p.gy_nc == COUT) {
#pragma unroll
              for (int i4 = 0; i4 < COUT; i4 += 4) {
                const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp + i4));
                f[i4] = g4.x; f[i4 + 1] = g4.y; f[i4 + 2] = g4.z; f[i4 + 3] = g4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < COUT; ++i) { f[i] = i < p.gy_nc ? __ldg(gp) : 0.f; gp += p.gy_sc; }
            }
          } else {
#pragma unroll
            for (int i = 0; i < COUT; ++i) f[i] = 0.f;
          }
          rinv = 1.f;
        } else {
        const uint32_t acc = kk & 1u;
        FWAIT(tf_full + 8 * acc, (kk >> 1) & 1u, 27);
        tc_fence_after();
        tq0 = PROF ? clock64() : 0;
        uint32_t v[COUT], v2[COUT];
        const uint32_t ta = tmem_f + (static_cast<uint32_t>(quad * 32) << 16) + acc * FCOLS;
        TmemLd<COUT>::ld(ta, v);
        TmemLd<COUT>::ld(ta + COUT, v2);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tf_empty + 8 * acc);        // the accumulator is in registers: hand the buffer back
        if (PROF) { const long long tq = clock64(); prof_acc[0] += tq - tq0; tq0 = tq; }      // [0] TMEM load
        if (p.dbg & 4) {
          if (p.update) { fence_proxy_async(); __syncwarp(); if (lane == 0) mbar_arrive(r_full + 8 * (kk % (uint32_t)p.nrs)); }
          continue;
        }
        const int q = k * 128 + loc;
        const int r_ = (int)__umulhi((unsigned)q, p.pitch_magic), c_ = q - r_ * p.pitch;      // q / pitch (q < 2^16)
        const bool valid = r_ < THv && c_ < TWv;
#pragma unroll
        for (int i4 = 0; i4 < COUT; i4 += 4) {
          const float4 sc = *reinterpret_cast<const float4*>(s_inv + i4);
          const float4 bs = *reinterpret_cast<const float4*>(s_bias + i4);
          f[i4 + 0] = fmaf(__uint_as_float(v[i4 + 0]) + __uint_as_float(v2[i4 + 0]), sc.x, bs.x);
          f[i4 + 1] = fmaf(__uint_as_float(v[i4 + 1]) + __uint_as_float(v2[i4 + 1]), sc.y, bs.y);
          f[i4 + 2] = fmaf(__uint_as_float(v[i4 + 2]) + __uint_as_float(v2[i4 + 2]), sc.z, bs.z);
          f[i4 + 3] = fmaf(__uint_as_float(v[i4 + 3]) + __uint_as_float(v2[i4 + 3]), sc.w, bs.w);
        }
        const long long pix = (long long)(h0 + r_) * p.oW + (w0 + c_);
        if (valid && !(p.dbg & 16)) {          // ONE branch around all the stores (a condition per store costs a
          float* yp = p.y + (long long)b * COUT * outS + pix;      // reconvergence region each: measured 11 % of the kernel)
#pragma unroll
          for (int i = 0; i < COUT; ++i) { *yp = f[i]; yp += outS; }
        }
        float best = -INFINITY;
        int bi = 0;
        if (p.winner) {
#pragma unroll
          for (int i = 0; i < COUT; ++i)
            if (f[i] > best) { best = f[i]; bi = i; }            // strict: the lowest index wins ties
        }
        if (want_ys) {
          if constexpr (YACC) {
#pragma unroll
            for (int i = 0; i < COUT; ++i) { const float tt = valid ? f[i] : 0.f; ysacc[i] += tt; yqacc[i] = fmaf(tt, tt, yqacc[i]); }
          } else {
            float t1[COUT], t2[COUT];
#pragma unroll
            for (int i = 0; i < COUT; ++i) { t1[i] = valid ? f[i] : 0.f; t2[i] = t1[i] * t1[i]; }
            const float a1 = lane_col_sum<COUT>(t1, lane), a2 = lane_col_sum<COUT>(t2, lane);
            atomicAdd(s_ys + lane, a1); atomicAdd(s_yq + lane, a2);
          }
        }
        if (p.winner && valid) {
          p.winner[(long long)b * outS + pix] = bi;
          float second = -INFINITY, amax = 0.f;
#pragma unroll
          for (int i = 0; i < COUT; ++i) { second = fmaxf(second, i == bi ? -INFINITY : f[i]); amax = fmaxf(amax, fabsf(f[i])); }
          if (best - second <= p.tie_rel * amax) flag_tie_f(p.fix_list, p.fix_count, p.fix_cap, (long long)b * outS + pix);
        }
        if (PROF) { const long long tq = clock64(); prof_acc[1] += tq - tq0; tq0 = tq; }      // [1] y, winner, sums
        if (!p.update) continue;
        // ---- responses: r = softmax_c(k y), bf16 hi/lo, into this block's slot of the response ring ----
        float mx2 = -INFINITY;
#pragma unroll
        for (int i = 0; i < COUT; ++i) { f[i] *= k2; mx2 = fmaxf(mx2, f[i]); }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < COUT; ++i) {
          float e;
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(f[i] - mx2));
          f[i] = e; sum += e;
        }
        rinv = valid ? (1.f / sum) : 0.f;
        }
        const uint32_t nrs = (uint32_t)p.nrs;
        const uint32_t slot = kk % nrs, nslot = (kk + 1u) % nrs;
        // this block writes its own slot (last read by the update of block kk-nrs) and the lead of the next slot
        // (block kk-nrs+1): the update issuer works in order, so one wait on the younger of the two covers both
        if (kk >= nrs - 1) FWAIT(r_empty + 8 * nslot, ((kk - (nrs - 1)) / nrs) & 1u, 29);
        const uint32_t row = rb + slot * (kRSlotPos * RB) + (uint32_t)(kRLead + loc) * RB;
        // the last positions of a block are also the lead of the next block's slot (zeros at a tile's end)
        const bool carry = loc >= 128 - kRLead;
        const uint32_t lrow = rb + nslot * (kRSlotPos * RB) + (uint32_t)(loc - (128 - kRLead)) * RB;
        const bool tile_end = (k == p.NBLK - 1);
#pragma unroll
        for (int g8 = 0; g8 < COUT / 8; ++g8) {
          uint32_t oh4[4], ol4[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int c = g8 * 8 + i * 2;
            const float r0 = f[c] * rinv, r1 = f[c + 1] * rinv;
            uint32_t hp, lp;
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(r1), "f"(r0));
            const float hh0 = __uint_as_float(hp << 16), hh1 = __uint_as_float(hp & 0xffff0000u);
            asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp) : "f"(r1 - hh1), "f"(r0 - hh0));
            oh4[i] = hp; ol4[i] = lp;
            racc[c] += hh0 + __uint_as_float(lp << 16);
            racc[c + 1] += hh1 + __uint_as_float(lp & 0xffff0000u);
          }
          st_shared_v4(swz<RCH>(row, g8), oh4[0], oh4[1], oh4[2], oh4[3]);
          st_shared_v4(swz<RCH>(row, RCH / 2 + g8), ol4[0], ol4[1], ol4[2], ol4[3]);
          if (carry) {
            if (tile_end) { oh4[0] = oh4[1] = oh4[2] = oh4[3] = 0u; ol4[0] = ol4[1] = ol4[2] = ol4[3] = 0u; }
            st_shared_v4(swz<RCH>(lrow, g8), oh4[0], oh4[1], oh4[2], oh4[3]);
            st_shared_v4(swz<RCH>(lrow, RCH / 2 + g8), ol4[0], ol4[1], ol4[2], ol4[3]);
          }
        }
        if (PROF) { const long long tq = clock64(); prof_acc[2] += tq - tq0; tq0 = tq; }      // [2] softmax, split, response stores
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(r_full + 8 * slot);
        if (PROF) { const long long tq = clock64(); prof_acc[3] += tq - tq0; tq0 = tq; }      // [3] proxy fence + arrive
      }
    }
    // ---- fold the per-thread running sums ----
    __syncwarp();
    if (p.update) {
      constexpr int CHF = COUT >= 32 ? 32 : 16;
#pragma unroll
      for (int ck = 0; ck < COUT / CHF; ++ck) {
        float tt[CHF];
#pragma unroll
        for (int i = 0; i < CHF; ++i) tt[i] = racc[ck * CHF + i];
        const float cs = lane_col_sum<CHF>(tt, lane);
        if (lane < CHF) my_rs[ck * CHF + lane] = cs;
      }
      __syncwarp();
      for (int c = lane; c < COUT; c += 32) atomicAdd(p.rsum + c, my_rs[c]);
    }
    if (want_ys) {
      if constexpr (YACC) {
        float t1[COUT], t2[COUT];
#pragma unroll
        for (int i = 0; i < COUT; ++i) { t1[i] = ysacc[i]; t2[i] = yqacc[i]; }
        const float a1 = lane_col_sum<COUT>(t1, lane), a2 = lane_col_sum<COUT>(t2, lane);
        if (lane < COUT) { atomicAdd(s_ys + lane, a1); atomicAdd(s_yq + lane, a2); }
      }
      asm volatile("bar.sync 3, 256;" ::: "memory");        // the 8 epilogue warps
      for (int c = (int)threadIdx.x - 32 * (3 + NCW); c < COUT; c += 256) {
        atomicAdd(p.ystats + 2 * c, (double)s_ys[c]);
        atomicAdd(p.ystats + 2 * c + 1, (double)s_yq[c]);
      }
    }
    // ---- the update accumulators: one partial per CTA and hi/lo half of x; set s drains accumulator s ----
    if (p.update) {
      FWAIT(done, 0, 28);
      tc_fence_after();
      const bool have = my_tiles > 0;
      if (eset < NACC) {
        const int rowi = quad * 32 + lane;
        const int jj = rowi / (2 * CIN), rr = rowi - jj * (2 * CIN);
        const int hl = rr / CIN, ci = rr - hl * CIN;
        const int kh = eset == 0 ? jj : kh_base1 + jj;
        const bool mine = kh < p.kH && (eset == 0 || kh >= COPIES);      // accumulator 1 repeats rows accumulator 0 owns
        const uint32_t ta = tmem_d + (static_cast<uint32_t>(quad * 32) << 16) + eset * NW;
        for (int i = 0; i < p.kW; ++i) {
          const int tap = kh * p.kW + (p.kW - 1 - i);          // column copy i holds r shifted by +i positions: kw = kW-1-i
          float* dst = p.hpart + ((((long long)blockIdx.x * 2 + hl) * p.taps + tap) * CIN + ci) * COUT;
#pragma unroll
          for (int c0 = 0; c0 < COUT; c0 += 16) {
            uint32_t vh[16], vl[16];
            tmem_ld16(ta + i * 2 * COUT + c0, vh);
            tmem_ld16(ta + i * 2 * COUT + COUT + c0, vl);
            tmem_ld_wait();
            if (mine) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                float4 o;
                o.x = have ? __uint_as_float(vh[4 * u + 0]) + __uint_as_float(vl[4 * u + 0]) : 0.f;
                o.y = have ? __uint_as_float(vh[4 * u + 1]) + __uint_as_float(vl[4 * u + 1]) : 0.f;
                o.z = have ? __uint_as_float(vh[4 * u + 2]) + __uint_as_float(vl[4 * u + 2]) : 0.f;
                o.w = have ? __uint_as_float(vh[4 * u + 3]) + __uint_as_float(vl[4 * u + 3]) : 0.f;
                *reinterpret_cast<float4*>(dst + c0 + 4 * u) = o;
              }
            }
          }
        }
      }
    }
  }
  if (PROF && p.prof && lane == 0) {
    long long* dst = p.prof + ((long long)blockIdx.x * 15 + warp) * 11;      // (15 = the widest layout)
#pragma unroll
    for (int i = 0; i < 10; ++i) dst[i] = prof_acc[i];
    dst[10] = clock64() - prof_t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, p.tmem_cols);
}
#undef FWAIT

// One launch that prepares a layer call: 1/|W_c| per filter (hebb.py:10-13), the packed bf16 hi/lo weight image of the
// forward B operand ([slab][tap][k-chunk][hi|lo][Cout] x 16 bytes, as pack_w_kernel writes it), and the zeroing of the
// per-call accumulators (sum_p r, error word, near-tie counter, BatchNorm sums).  One block per filter.
__global__ void __launch_bounds__(128)
fused_prep_kernel(const float* __restrict__ W, uint4* __restrict__ wp, float* __restrict__ inv, uint32_t* __restrict__ zero32,
                  int nzero32, double* __restrict__ ystats, int Cin, int NSLAB, int Cout, int taps, int wnrm) {
  const int c = blockIdx.x, K = Cin * taps;
  const float* w = W + (long long)c * K;
  if (wnrm) {
    float acc = 0.f;
    for (int i = threadIdx.x; i < K; i += blockDim.x) { const float v = __ldg(w + i); acc = fmaf(v, v, acc); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ float part[4];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float nrm = sqrtf(part[0] + part[1] + part[2] + part[3]);
      inv[c] = nrm == 0.f ? 1.f : 1.f / nrm;                 // hebb.py:12: zero norms divide by 1
    }
  }
  const int items = NSLAB * taps * 4;                         // (slab, tap, k-chunk, hi|lo); channels >= Cin are zero
  for (int it = threadIdx.x; it < items; it += blockDim.x) {
    const int hl = it & 1, c2 = (it >> 1) & 1, st = it >> 2;
    const int tap = st % taps, slab = st / taps;
    uint32_t o[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat16 e[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int ci = (slab * 2 + c2) * 8 + 2 * i + j;
        const float v = ci < Cin ? __ldg(w + ci * taps + tap) : 0.f;
        __nv_bfloat16 hi, lo;
        split_bf16(v, hi, lo);
        e[j] = hl ? lo : hi;
      }
      o[i] = pack_bf16x2(e[0], e[1]);
    }
    wp[((((long long)slab * taps + tap) * 2 + c2) * 2 + hl) * Cout + c] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  if (c == 0) {
    for (int i = threadIdx.x; i < nzero32; i += blockDim.x) zero32[i] = 0u;
    if (ystats)
      for (int i = threadIdx.x; i < 2 * Cout; i += blockDim.x) ystats[i] = 0.0;
  }
}

// Weight-gradient mode: gw[co_off + co][ci_off + ci][tap] += the per-CTA partials, summed in a fixed order.
__global__ void __launch_bounds__(256)
fused_wgrad_finalize_kernel(const float* __restrict__ hpart, float* __restrict__ gw, int n_part, int taps, int CI, int CO,
                            int cin_tot, int ci_off, int co_off) {
  const int n = taps * CI * CO;
  const int idx = blockIdx.x * (blockDim.x / 8) + (threadIdx.x >> 3);      // 8 threads per output element
  const int sub = threadIdx.x & 7;
  float acc = 0.f;
  if (idx < n)
    for (int part = sub; part < n_part; part += 8) acc += __ldg(hpart + (long long)part * n + idx);
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (idx < n && sub == 0) {
    const int co = idx % CO, ci = (idx / CO) % CI, tap = idx / (CO * CI);
    gw[((long long)(co_off + co) * cin_tot + ci_off + ci) * taps + tap] += acc;
  }
}

// ---------------------------------------------------------------------------------------------------------------
struct FPlan {
  int TH, TW, pitch, nTH, nTW, ntiles, XROWS, NBLK, XPOS, NST, grid, BW, padl;
  int nrs;
  int gather, eCin, ek, etaps, srows;    // gather: few-channel 3x3 layer run as a 1x1 layer on eCin = 32 pseudo-channels
  uint32_t off_r, off_stage, off_w, off_misc, w_bytes, stage_bytes, smem, tmem_cols;
  size_t o_inv, o_rsum, o_err, o_wp, o_hpart, o_fix, o_prof, total;
  int fix_cap;
};

bool fused_plan(const Geo& g, FPlan* P) {
  if (g.nd != 2 || g.transposed || g.kD != 1 || g.sH != 1 || g.sW != 1) return false;
  if (!(g.Cout == 16 || g.Cout == 32) || g.iW % 4 != 0) return false;
  // Few input channels (the first layer: 3 -> 16): the converter gathers the 3x3 patch of the real channels into
  // Cin*9 <= 32 pseudo-channels, and the layer runs as a 1x1 layer with 32 input channels over the output grid
  // ([Cout][Cin][3][3] already is that layer's weight) -- 4 forward instructions per block instead of 18
  const bool gather = g.Cin <= 3 && g.kH == 3 && g.kW == 3;
  if (!gather && !((g.Cin == 16 || g.Cin == 32) && ((g.kH == 1 && g.kW == 1) || (g.kH == 3 && g.kW == 3)))) return false;
  static const int want = [] { const char* e = getenv("HEBB_FUSED"); return (e && e[0] == '0') ? 0 : 1; }();
  if (!want) return false;
  FPlan& q = *P;
  const int sms = num_sms();
  const int eCin = gather ? 32 : g.Cin, ekH = gather ? 1 : g.kH, ekW = gather ? 1 : g.kW;      // what the MMA side sees
  q.gather = gather ? 1 : 0; q.eCin = eCin; q.ek = ekH; q.etaps = ekH * ekW;
  const int XB = eCin * 4, RB = g.Cout * 4;
  // The TMA box starts padl = round_up(pW, 4) columns left of the tile, so that its first byte is 16-byte aligned in
  // global memory: a box whose first element is not 16-byte aligned raises an illegal-instruction fault on sm_100
  // (HEBB_FUSED_ALIGN=0 reproduces it)
  static const int want_align = [] { const char* e = getenv("HEBB_FUSED_ALIGN"); return (e && e[0] == '0') ? 0 : 1; }();
  q.padl = want_align ? (g.pW + 3) / 4 * 4 : g.pW;
  q.w_bytes = (uint32_t)((eCin / 16) * q.etaps * 4 * g.Cout * 16);
  const uint32_t misc = (uint32_t)(12 * g.Cout * 4 + 64 * 4 + 8 * kNumBars + 64);
  bool found = false;
  double best = 1e300, best_waste = 1e300;
  const int copies = 128 / (2 * eCin);
  const int reach = (copies - 1) > (ekH - 1) ? (copies - 1) : (ekH - 1);
  // Search tile width (whole rows up to 128 pixels, or 64-pixel columns where the 128-byte-per-position images would
  // otherwise leave no room), response-ring depth and tile height.  The kernel is bound by the tensor pipe's operand
  // fetches, so the positions the 128-wide blocks compute beyond the real pixels (row pitch > TW, the block that
  // overhangs the tile, the last tile row of the image) cost time one to one; the re-converted halo rows cost a
  // little; and the fp32 staging ring must keep ~48 KB of TMA boxes in flight per SM to cover the DRAM latency
  // (measured: with 4 x 8 KB in flight the bare pipeline took 0.1 ms per 2304 tiles).
  for (int twc = 0; twc < 2; ++twc) {
    const int tw_max = twc == 0 ? 128 : 64;
    if (twc == 1 && (g.oW <= 64 || gather)) break;      // (the gather converts one pixel per thread: 128-pixel rows)
    int nTW = (int)cdiv(g.oW, tw_max);
    const int TW = (int)(cdiv(cdiv(g.oW, nTW), 4) * 4);          // tile columns start on multiples of 4 floats
    nTW = (int)cdiv(g.oW, TW);
    const int pitch = (int)((TW + ekW - 1 + 3) / 4 * 4);
    const int BW = (pitch + (gather ? g.kW - 1 : 0) + (q.padl - g.pW) + 3) / 4 * 4;
    if (BW > 256) continue;
    const uint32_t stage = (uint32_t)align_up((size_t)BW * (gather ? g.Cin : 16) * 4, 128);
    const long long tiles_w = (long long)g.B * nTW;
    for (int nrs = kRSlots; nrs >= kRSlots - 1; --nrs) {
      const uint32_t r_bytes = (uint32_t)nrs * kRSlotPos * RB;
      const uint32_t fixed = (uint32_t)align_up(r_bytes, 1024) + (uint32_t)align_up(q.w_bytes, 128) + misc + 1024;
      for (int th = 16; th >= 1; --th) {
        if (th + ekH - 1 > kMaxRows) continue;
        if (th > g.oH && th > 1) continue;
        // the update of a block is issued after its responses went through the epilogue: the rows the next tile's first
        // block waits for (kH) must have been released by the updates that can have been issued by then: TH - 1 >= kH
        if (th < ekH + 1) continue;
        const long long tiles = tiles_w * cdiv(g.oH, th);
        if (th > ekH + 1 && tiles < 4LL * sms && tiles_w * g.oH >= 4LL * sms) continue;      // enough tiles to balance the CTAs
        const int nblk = (int)cdiv((long long)th * pitch, 128);
        const int xpos = nblk * 128 + reach * pitch + 8;
        const uint32_t x_bytes = (uint32_t)align_up((size_t)xpos * XB, 1024);
        const int min_st = gather ? 4 : 3;                  // the gather keeps 3 staged rows open at a time
        if (x_bytes + fixed + min_st * stage > (uint32_t)kSmemLimitF) continue;
        int nst = (int)(((uint32_t)kSmemLimitF - x_bytes - fixed) / stage);
        if (nst > kMaxStages) nst = kMaxStages;
        const double waste = (double)nblk * 128.0 / ((double)th * TW) * ((double)cdiv(g.oH, th) * th / g.oH) *
                             ((double)nTW * TW / g.oW);
        const double halo = (double)(th + g.kH - 1) / th * (double)(TW + g.kW - 1) / TW;
        double fl = gather ? 1.0 : (double)nst * stage / 49152.0; if (fl > 1.0) fl = 1.0;
        const double cost = waste * (0.8 + 0.2 * halo) / (0.5 + 0.5 * fl) * (nrs < kRSlots ? 1.03 : 1.0);
        if (cost < best) {
          best = cost; best_waste = waste * (0.8 + 0.2 * halo); found = true;
          q.TW = TW; q.nTW = nTW; q.pitch = pitch; q.BW = BW; q.stage_bytes = stage; q.nrs = nrs;
          q.TH = th; q.NBLK = nblk; q.XPOS = xpos; q.XROWS = th + ekH - 1; q.NST = nst; q.srows = th + g.kH - 1;
          q.off_r = x_bytes;
          q.off_stage = q.off_r + (uint32_t)align_up(r_bytes, 1024);
          q.off_w = q.off_stage + nst * stage;
          q.off_misc = q.off_w + (uint32_t)align_up(q.w_bytes, 128);
          q.smem = x_bytes + fixed + nst * stage;
        }
      }
    }
  }
  // a tile that wastes this much (tiny TH for the 128-byte-per-position images) is slower than the two-kernel path
  if (!found || best_waste > 1.6) return false;
  q.nTH = (int)cdiv(g.oH, q.TH);
  q.ntiles = g.B * q.nTH * q.nTW;
  q.grid = q.ntiles < sms ? q.ntiles : sms;
  const int nacc = (ekH + copies - 1) / copies;
  const int cols = 4 * g.Cout + nacc * ekW * 2 * g.Cout;      // 2 forward accumulators of 2 Cout columns + the update's
  uint32_t tc = 32; while ((int)tc < cols) tc <<= 1;
  if (tc > 512) return false;
  q.tmem_cols = tc;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
  q.o_inv = take(sizeof(float) * g.Cout);
  q.o_rsum = take(sizeof(float) * g.Cout);
  q.o_err = take(256);
  q.o_wp = take(q.w_bytes);
  q.o_hpart = take((size_t)q.grid * 2 * q.etaps * eCin * g.Cout * sizeof(float));
  const long long px = (long long)g.B * g.outS;
  q.fix_cap = (int)(px < (1LL << 18) ? px : (1LL << 18));
  q.o_fix = take(sizeof(int) * (size_t)q.fix_cap);
  q.o_prof = take(sizeof(long long) * (size_t)sms * 15 * 11);
  q.total = off;
  return true;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
      (void)cudaGetLastError();
      p = nullptr;
    }
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

}  // namespace

static long long* g_last_prof = nullptr;
static int g_last_prof_n = 0;

// Launch wrapper of fused_prep_kernel for the two-kernel path (tc_path.cu): small weight tensors, one channel tile.
int launch_layer_prep(const float* W, void* wp, float* inv, void* zero_from, size_t zero_bytes, double* ystats, int Cin, int Cout,
                      int taps, int wnrm, cudaStream_t st) {
  fused_prep_kernel<<<Cout, 128, 0, st>>>(W, reinterpret_cast<uint4*>(wp), inv, reinterpret_cast<uint32_t*>(zero_from), (int)(zero_bytes / 4),
                                          ystats, Cin, (Cin + 15) / 16, Cout, taps, wnrm);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  return HEBB_OK;
}

bool fused_supported(const Geo& g, int prec, unsigned flags) {
  if (prec != HEBB_PREC_BF16X3 && prec != HEBB_PREC_BF16) return false;
  if (flags & (HEBB_F_RULE_HPCA | HEBB_F_WGRAD_INTERNAL | HEBB_F_ONLY_PACK | HEBB_F_ONLY_FWD | HEBB_F_ONLY_DW)) return false;
  FPlan P;
  return fused_plan(g, &P) && encode_fn() != nullptr;
}

size_t fused_workspace_bytes(const Geo& g) {
  FPlan P;
  return fused_plan(g, &P) ? P.total : 0;
}

int fused_describe_plan(const Geo& g, int* out, int n) {
  FPlan P;
  if (!fused_plan(g, &P)) return 0;
  const int v[] = {P.TH, P.TW, P.pitch, P.ntiles, P.NBLK, P.XROWS, (int)P.smem, (int)P.tmem_cols, P.grid, P.NST, P.nrs};
  const int m = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < m; ++i) out[i] = v[i];
  return m;
}

int fused_conv_step(const Geo& g, const float* x, const float* W, const float* bias, float kinv, float* y, int32_t* winner,
                    float* delta_w, void* ws, size_t ws_bytes, unsigned flags, cudaStream_t st, double* ystats,
                    int* ystats_written) {
  if (ystats_written) *ystats_written = 0;
  FPlan P;
  if (!fused_plan(g, &P)) return HEBB_ESHAPE;
  if (!ws || ws_bytes < P.total) return HEBB_EWS;
  if (reinterpret_cast<uintptr_t>(x) & 15) return HEBB_EALIGN;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return HEBB_ECUDA;
  char* base = static_cast<char*>(ws);
  float* inv = reinterpret_cast<float*>(base + P.o_inv);
  float* rsum = reinterpret_cast<float*>(base + P.o_rsum);
  int* err = reinterpret_cast<int*>(base + P.o_err);
  const bool upd = (flags & HEBB_F_UPDATE) != 0;
  // one launch: filter norms, packed weights, zeroed per-call accumulators (sum_p r, error word, near-tie counter, BatchNorm sums)
  const int pCin = P.gather ? g.Cin * g.taps : g.Cin;          // the gathered layer's weight is [Cout][Cin*9] of a 1x1 layer
  fused_prep_kernel<<<g.Cout, 128, 0, st>>>(W, reinterpret_cast<uint4*>(base + P.o_wp), inv, reinterpret_cast<uint32_t*>(base + P.o_rsum),
                                            (int)((P.o_wp - P.o_rsum) / 4), ystats, pCin, P.eCin / 16, g.Cout, P.etaps, (flags & HEBB_F_WNRM) ? 1 : 0);
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();

  CUtensorMap tm;
  const cuuint64_t dims[3] = {(cuuint64_t)g.iW, (cuuint64_t)g.iH, (cuuint64_t)g.B * g.Cin};      // (the REAL channels)
  const cuuint64_t strides[2] = {(cuuint64_t)g.iW * 4, (cuuint64_t)g.iW * g.iH * 4};
  const cuuint32_t box[3] = {(cuuint32_t)P.BW, 1u, (cuuint32_t)(P.gather ? g.Cin : 16)};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return HEBB_ESHAPE;

  FusedParams f;
  f.y = y; f.winner = winner; f.inv = (flags & HEBB_F_WNRM) ? inv : nullptr; f.bias = bias; f.rsum = rsum;
  f.ystats = ystats; f.hpart = reinterpret_cast<float*>(base + P.o_hpart);
  f.err = watchdog_word() ? watchdog_word() : err;
  static const float tie_rel = [] { const char* e = getenv("HEBB_TIE_REL"); return e ? (float)atof(e) : 2.5e-4f; }();
  f.fix_list = reinterpret_cast<int*>(base + P.o_fix); f.fix_count = err + 4; f.fix_cap = P.fix_cap; f.tie_rel = tie_rel;
  f.wp = reinterpret_cast<const uint4*>(base + P.o_wp);
  f.B = g.B; f.oH = g.oH; f.oW = g.oW; f.kH = P.ek; f.kW = P.ek; f.pH = g.pH; f.pW = g.pW; f.taps = P.etaps;
  f.gather = P.gather; f.gcin = g.Cin; f.gk = g.kH; f.srows = P.srows; f.nrs = P.nrs;
  f.TH = P.TH; f.TW = P.TW; f.pitch = P.pitch; f.nTH = P.nTH; f.nTW = P.nTW; f.ntiles = P.ntiles; f.XROWS = P.XROWS;
  f.NBLK = P.NBLK; f.XPOS = P.XPOS; f.NST = P.NST; f.BW = P.BW; f.padl = P.padl;
  f.pitch_magic = (unsigned)((0x100000000ULL + (unsigned)P.pitch - 1) / (unsigned)P.pitch);

  f.kinv = kinv; f.update = upd ? 1 : 0;
  f.gy = nullptr; f.gy_sb = f.gy_sc = f.gy_sp = 0; f.xcl = 0; f.cin_tot = g.Cin; f.ci_off = 0;
  static const int fdbg = [] { const char* e = getenv("HEBB_FUSED_DBG"); return e ? atoi(e) : 0; }();
  f.dbg = fdbg;
  static const int fprof = [] { const char* e = getenv("HEBB_FUSED_PROF"); return (e && e[0] == '1') ? 1 : 0; }();
  f.prof = fprof ? reinterpret_cast<long long*>(base + P.o_prof) : nullptr;
  g_last_prof = f.prof; g_last_prof_n = P.grid * 15 * 11;
  f.off_r = P.off_r; f.off_stage = P.off_stage; f.off_w = P.off_w; f.off_misc = P.off_misc; f.w_bytes = P.w_bytes;
  f.stage_bytes = P.stage_bytes; f.tmem_cols = P.tmem_cols;
  if (ystats && ystats_written) *ystats_written = 1;           // (zeroed by the prep kernel)
#define HEBB_FUSED_LAUNCH2(CI, CO, K, PR, NC)                                                                             \
  do {                                                                                                                    \
    HEBB_CUDA_TRY(cudaFuncSetAttribute(fused_small_kernel<CI, CO, K, PR, NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimitF)); \
    fused_small_kernel<CI, CO, K, PR, NC><<<P.grid, 32 * (3 + NC + 8), kSmemLimitF, st>>>(tm, f);                         \
  } while (0)
#define HEBB_FUSED_LAUNCH(CI, CO)                                                                                         \
  do {                                                                                                                    \
    if (P.ek == 3) { if (f.prof) HEBB_FUSED_LAUNCH2(CI, CO, 3, true, 2); else HEBB_FUSED_LAUNCH2(CI, CO, 3, false, 2); }  \
    else { f.prof = nullptr; HEBB_FUSED_LAUNCH2(CI, CO, 1, false, 2); }                                                   \
  } while (0)
  if (P.gather) {
    if (g.Cout == 16) { if (f.prof) HEBB_FUSED_LAUNCH2(32, 16, 1, true, 4); else HEBB_FUSED_LAUNCH2(32, 16, 1, false, 4); }
    else { f.prof = nullptr; HEBB_FUSED_LAUNCH2(32, 32, 1, false, 4); }
  }
  else if (P.eCin == 16 && g.Cout == 16) HEBB_FUSED_LAUNCH(16, 16);
  else if (P.eCin == 16 && g.Cout == 32) HEBB_FUSED_LAUNCH(16, 32);
  else if (P.eCin == 32 && g.Cout == 16) HEBB_FUSED_LAUNCH(32, 16);
  else HEBB_FUSED_LAUNCH(32, 32);
#undef HEBB_FUSED_LAUNCH2
#undef HEBB_FUSED_LAUNCH
  HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
  if (winner && tie_rel > 0.f)
    HEBB_TRY(launch_winner_fixup(g, x, W, (flags & HEBB_F_WNRM) ? inv : nullptr, bias, winner, f.fix_list, f.fix_count, P.fix_cap, st));
  if (upd)
    HEBB_TRY(tc_launch_finalize(f.hpart, rsum, W, delta_w, P.grid * 2, P.etaps, pCin, P.eCin, g.Cout, st));
  return HEBB_OK;
}

// ---- hebb_conv_wgrad on the fused kernel ("the wgrad kernel is a6 with dL/dy in place of r") ----
// Channel counts that are multiples of 16 run as passes of (16|32) x (16|32) channels: every pass re-reads one of the
// two tensors, so only layers that need at most kMaxWgradPasses passes are taken (the back-prop head of the 2-D
// network: 16 -> 64 and 64 -> 32; its tall-skinny reduction over 4 M pixels is the shape this kernel was built for).
static bool fused_wgrad_split(const Geo& g, Geo* sub, int* cip, int* cop) {
  if (g.nd != 2 || g.transposed) return false;
  if (g.Cin % 16 != 0 || g.Cout % 16 != 0) return false;
  static const int want = [] { const char* e = getenv("HEBB_FUSED_WGRAD"); return (e && e[0] == '0') ? 0 : 1; }();
  static const int max_passes = [] { const char* e = getenv("HEBB_FUSED_WGRAD_PASSES"); return e ? atoi(e) : 2; }();
  if (!want) return false;
  const int ci = (g.Cin % 32 == 0) ? 32 : 16, co = (g.Cout % 32 == 0) ? 32 : 16;
  if ((g.Cin / ci) * (g.Cout / co) > max_passes) return false;
  *sub = g;
  sub->Cin = ci; sub->Cout = co;
  *cip = ci; *cop = co;
  return true;
}

bool fused_wgrad_supported(const Geo& g) {
  Geo sub; int ci, co; FPlan P;
  return fused_wgrad_split(g, &sub, &ci, &co) && fused_plan(sub, &P) && !P.gather && encode_fn() != nullptr;
}

size_t fused_wgrad_workspace_bytes(const Geo& g) {
  Geo sub; int ci, co; FPlan P;
  return (fused_wgrad_split(g, &sub, &ci, &co) && fused_plan(sub, &P)) ? P.total : 0;
}

int fused_conv_wgrad(const Geo& g, const float* x, const float* gy, float* gw, int gy_channels, int channels_last, void* ws,
                     size_t ws_bytes, cudaStream_t st) {
  Geo sub; int CI, CO; FPlan P;
  if (!fused_wgrad_split(g, &sub, &CI, &CO) || !fused_plan(sub, &P) || P.gather) return HEBB_ESHAPE;
  const int gyC = gy_channels > 0 ? gy_channels : g.Cout;      // channels stored in gy; filters gyC.. get a zero gradient
  if (!ws || ws_bytes < P.total) return HEBB_EWS;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(gy) & 15)) return HEBB_EALIGN;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return HEBB_ECUDA;
  char* base = static_cast<char*>(ws);
  float* rsum = reinterpret_cast<float*>(base + P.o_rsum);
  int* err = reinterpret_cast<int*>(base + P.o_err);

  CUtensorMap tm;
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult rc;
  if (channels_last) {
    const cuuint64_t dims[4] = {(cuuint64_t)g.Cin, (cuuint64_t)g.iW, (cuuint64_t)g.iH, (cuuint64_t)g.B};
    const cuuint64_t strides[3] = {(cuuint64_t)g.Cin * 4, (cuuint64_t)g.iW * g.Cin * 4, (cuuint64_t)g.iH * g.iW * g.Cin * 4};
    const cuuint32_t box[4] = {16u, (cuuint32_t)P.BW, 1u, 1u};
    rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[3] = {(cuuint64_t)g.iW, (cuuint64_t)g.iH, (cuuint64_t)g.B * g.Cin};
    const cuuint64_t strides[2] = {(cuuint64_t)g.iW * 4, (cuuint64_t)g.iW * g.iH * 4};
    const cuuint32_t box[3] = {(cuuint32_t)P.BW, 1u, 16u};
    rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (rc != CUDA_SUCCESS) return HEBB_ESHAPE;

  FusedParams f;
  memset(&f, 0, sizeof(f));
  f.rsum = rsum; f.hpart = reinterpret_cast<float*>(base + P.o_hpart);
  f.err = watchdog_word() ? watchdog_word() : err;
  f.fix_list = reinterpret_cast<int*>(base + P.o_fix); f.fix_count = err + 4; f.fix_cap = P.fix_cap;
  f.B = g.B; f.oH = g.oH; f.oW = g.oW; f.kH = P.ek; f.kW = P.ek; f.pH = g.pH; f.pW = g.pW; f.taps = P.etaps;
  f.gcin = CI; f.gk = g.kH; f.srows = P.srows; f.nrs = P.nrs;
  f.TH = P.TH; f.TW = P.TW; f.pitch = P.pitch; f.nTH = P.nTH; f.nTW = P.nTW; f.ntiles = P.ntiles; f.XROWS = P.XROWS;
  f.NBLK = P.NBLK; f.XPOS = P.XPOS; f.NST = P.NST; f.BW = P.BW; f.padl = P.padl;
  f.pitch_magic = (unsigned)((0x100000000ULL + (unsigned)P.pitch - 1) / (unsigned)P.pitch);
  f.kinv = 1.f; f.update = 1;
  f.off_r = P.off_r; f.off_stage = P.off_stage; f.off_w = P.off_w; f.off_misc = P.off_misc; f.w_bytes = 0;
  f.stage_bytes = P.stage_bytes; f.tmem_cols = P.tmem_cols;
  f.xcl = channels_last ? 1 : 0; f.cin_tot = g.Cin;
  if (channels_last) { f.gy_sb = g.outS * gyC; f.gy_sc = 1; f.gy_sp = gyC; }
  else { f.gy_sb = g.outS * gyC; f.gy_sc = g.outS; f.gy_sp = 1; }
  const int n_out = P.etaps * CI * CO;
  // converter warps of the weight-gradient mode: 6 (measured on the head's layers, channels_last: 2 -> 4 -> 6 warps with whole-
  // stage ownership: 2.31 -> 1.58 -> 1.50 ms for the three of them); HEBB_FUSED_WG_NCW=4 selects the profiled 4-warp build
  static const int wg_ncw = [] { const char* e = getenv("HEBB_FUSED_WG_NCW"); return e ? atoi(e) : 6; }();
  static const int wprof = [] { const char* e = getenv("HEBB_FUSED_PROF"); return (e && e[0] == '1') ? 1 : 0; }();
  f.prof = (wprof && CI == 32 && P.ek == 3) ? reinterpret_cast<long long*>(base + P.o_prof) : nullptr;
  if (f.prof) { g_last_prof = f.prof; g_last_prof_n = P.grid * 15 * 11; }
  for (int co0 = 0; co0 < g.Cout; co0 += CO)
    for (int ci0 = 0; ci0 < g.Cin; ci0 += CI) {
      HEBB_CUDA_TRY(cudaMemsetAsync(base + P.o_rsum, 0, P.o_wp - P.o_rsum, st));      // sum of gy, error word
      f.gy = gy + (long long)co0 * f.gy_sc;
      f.gy_nc = gyC - co0 < CO ? (gyC - co0 > 0 ? gyC - co0 : 0) : CO;
      f.ci_off = ci0;
#define HEBB_FUSED_WG2(CIv, COv, Kv, NCv)                                                                                \
  do {                                                                                                                    \
    HEBB_CUDA_TRY(cudaFuncSetAttribute(fused_small_kernel<CIv, COv, Kv, false, NCv, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimitF)); \
    fused_small_kernel<CIv, COv, Kv, false, NCv, true><<<P.grid, 32 * (3 + NCv + 8), kSmemLimitF, st>>>(tm, f);           \
  } while (0)
#define HEBB_FUSED_WG(CIv, COv)                                                                                          \
  do {                                                                                                                    \
    if (P.ek == 3) {                                                                                                      \
      if (wg_ncw == 4) HEBB_FUSED_WG2(CIv, COv, 3, 4);                                                                    \
      else HEBB_FUSED_WG2(CIv, COv, 3, 6);                                                                                \
    } else {                                                                                                              \
      HEBB_FUSED_WG2(CIv, COv, 1, 2);                                                                                     \
    }                                                                                                                     \
  } while (0)
      if (f.prof && CO == 16) {          // (the wait-cycle table exists for the 4-converter-warp layout)
        HEBB_CUDA_TRY(cudaFuncSetAttribute(fused_small_kernel<32, 16, 3, true, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimitF));
        fused_small_kernel<32, 16, 3, true, 4, true><<<P.grid, 32 * 15, kSmemLimitF, st>>>(tm, f);
      } else if (f.prof) {
        HEBB_CUDA_TRY(cudaFuncSetAttribute(fused_small_kernel<32, 32, 3, true, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimitF));
        fused_small_kernel<32, 32, 3, true, 4, true><<<P.grid, 32 * 15, kSmemLimitF, st>>>(tm, f);
      }
      else if (CI == 16 && CO == 16) HEBB_FUSED_WG(16, 16);
      else if (CI == 16 && CO == 32) HEBB_FUSED_WG(16, 32);
      else if (CI == 32 && CO == 16) HEBB_FUSED_WG(32, 16);
      else HEBB_FUSED_WG(32, 32);
#undef HEBB_FUSED_WG
#undef HEBB_FUSED_WG2
      HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
      fused_wgrad_finalize_kernel<<<(unsigned)cdiv(n_out, 32), 256, 0, st>>>(f.hpart, gw, P.grid * 2, P.etaps, CI, CO, g.Cin, ci0, co0);
      HEBB_CUDA_TRY(cudaGetLastError()); HEBB_LAUNCHED();
    }
  return HEBB_OK;
}

}  // namespace hebb

// Profiling aid (HEBB_FUSED_PROF=1): copies the wait-cycle table of the last fused launch to the host: per CTA and warp
// 9 wait accumulators (codes 20..28) + the warp's total cycles.  Synchronises the device.
extern "C" int hebb_debug_fused_prof(long long* out, int n) {
  if (!hebb::g_last_prof || !out) return 0;
  const int m = n < hebb::g_last_prof_n ? n : hebb::g_last_prof_n;
  if (cudaMemcpy(out, hebb::g_last_prof, sizeof(long long) * (size_t)m, cudaMemcpyDeviceToHost) != cudaSuccess) return 0;
  return m;
}
