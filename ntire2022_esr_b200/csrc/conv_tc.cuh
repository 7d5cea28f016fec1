// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), fp16 in / fp32 accumulate in TMEM.
//
// GEMM view of one output tile: M = 128 consecutive pixels of one image row, N = output channels,
// K = taps x input channels.  Activations are NHWC fp16 with 64-channel (128-byte) pixels, so an
// image-row strip [130 px x 64 ch] lands in shared memory through ONE TMA box load in exactly the
// K-major SWIZZLE_128B layout tcgen05.mma wants (one pixel = one 128-byte row).  The 3x3 taps are
// not materialised: tap (dy,dx) is the same strip ring with the descriptor start address moved by
// dy strips and dx rows (the swizzle is a function of the absolute shared-memory address, so a
// row-shifted start needs no re-layout), so every input row is read from L2 once per strip
// (im2col-free, and the zero padding is TMA out-of-bounds fill).  A persistent CTA walks down a
// column strip keeping a ring of row strips in shared memory; weights are staged once per CTA as
// pre-swizzled blocks.
//
// Warp roles (352 threads, one CTA per SM): warp 0 = TMA producer, warps 1 and 10 = MMA issuers for even / odd
// tiles (one elected lane each; warp 1 also owns the TMEM allocation), warps 2..9 = epilogue (TMEM ->
// registers -> bias/residual/activation -> fp16 -> swizzled staging -> TMA store, or the fused pixel-shuffle
// store of the network tail).  Two or four TMEM accumulator slots let the epilogue of tile t overlap the MMAs
// of the following tiles.
//
// What bounds it (profiles/README.md): both MMA operands come from shared memory, 268 KB per tile of the
// 3x3 50->50(+25) layer = 2.06 k cycles at 128 B/clk against 1.35 k cycles of tensor math; with the epilogue's
// staging traffic the kernel sits at ~90 % of its shared-memory roofline (3.1 k cycles per tile at batch 16).
//
// Measured on B200 (tools/micro/mma_bench.cu): an M=128, K=16 SS-mode MMA costs 32 + N/4 cycles for
// N <= 128 (A and B are both fetched from shared memory at 128 B/clk), i.e. 48 cycles at N = 64.
#pragma once
#include <type_traits>

#include "kernels_generic.cuh"
#include "tc_common.cuh"

namespace esr {

constexpr int TC_MAX_ENTRIES = 16;
constexpr int TC_TILE_PX = 128;
constexpr int TC_MAX_SLOTS = 8;
constexpr int TC_MAX_GROUPS = 3;
constexpr int TC_MAX_CHUNKS = 10;    // 16-column epilogue units of a layer (host-side table only)
constexpr int TC_PREFETCH_ROWS = 6;  // L2 prefetch distance of the producer, in image rows
constexpr int TC_EPI_WARPS = 8;
constexpr int TC_THREADS = 96 + 32 * TC_EPI_WARPS;  // warp 0 TMA, warp 1 + warp 10 MMA issuers, warps 2..9 epilogue
constexpr int TC_DONE_BARS = 8;

// One MMA group = one (tap, K-chunk, column segment), pre-digested on the host so that the issuing thread
// spends a handful of instructions per tcgen05.mma (the issue loop, not the tensor pipe, was the limiter).
struct __align__(16) TcEntry {
  uint32_t a_row;   // bits 0-27: (chunk * chunk_bytes + px_off * 128) >> 4, bits 28-31: strip of the row window
  uint32_t b_off16; // (byte offset of the [n x 64] pre-swizzled weight block inside the blob) >> 4
  uint32_t idesc;   // instruction descriptor (M = 128, N = n)
  uint32_t misc;    // bits 0-15: first accumulator column, 16-19: K=16 steps, bit 31: first MMA overwrites
};

struct TcOutGroup {
  int32_t col0;        // first accumulator column of the group
  int32_t ncols;       // columns stored (multiple of 16, <= 64)
  int32_t act;         // Act
  float slope;
  int32_t res_after;   // residual added after (1) or before (0) the activation
  int32_t res_stride;  // residual: elements per pixel
  int32_t res_coff;
  int32_t mode;        // 0: TMA store NHWC fp16   1: pixel-shuffle x4 store into NCHW output
  int32_t swizzle;     // staging layout of the store tensor map: 0 linear, 1 SWIZZLE_128B (64 ch), 2 SWIZZLE_64B (32 ch), 3 SWIZZLE_32B (16 ch)
  int32_t stage_off;   // smem offset of the two staging buffers
  int32_t stage_bytes; // bytes of one staging buffer
  const float* bias;   // [ncols]
  const __half* res;   // nullptr = none
  // First group only, nullptr = none: bias per border class [9][64], class = 3 * (y == 0 ? 0 : y == H-1 ? 2 : 1) +
  // (x == 0 ? 0 : x == W-1 ? 2 : 1).  A BSConvU (pointwise Linear, THEN zero-padded depthwise 3x3) run as one dense
  // 3x3 convolution needs it: the Linear's bias reaches an output pixel only through the taps that lie inside the
  // image (models/team18_bsrn.py:82-88).  Used instead of `bias` for that group.
  const float* bias9;
};

// one unit of epilogue work: 16 accumulator columns of one output group
struct __align__(8) TcChunk {
  uint16_t tcol;    // accumulator column
  uint8_t group, c0, width, pad_[3];
};

struct TcParams {
  int32_t B, H, W;
  int32_t halo;          // 0 (1x1) or 1 (3x3)
  int32_t nchunks;       // 64-channel chunks per strip
  int32_t strip_px;      // 128 + 2*halo
  int32_t chunk_bytes;   // smem bytes reserved per chunk (multiple of 1024)
  int32_t strip_bytes;   // nchunks * chunk_bytes
  int32_t nslots;        // strip ring depth
  int32_t rows_per_item;
  int32_t strips_x, segs_y, n_items;
  int32_t n_entries, ngroups, n_epi_chunks;
  int32_t tmem_cols;     // TMEM allocation (power of two >= 2*acc_cols)
  int32_t acc_cols;      // columns of one accumulator slot
  int32_t acc_slots;     // accumulator slots in TMEM (2 or 4): how far the MMA warp may run ahead of the epilogue
  int32_t w_off, w_bytes, ring_off;
  int32_t ps_fp32;
  int32_t dbg_flags;     // experiments: 1 = issue no MMA, 2 = epilogue does no work
  int32_t pdl;           // launched with programmatic stream serialization
  int32_t chunk_c0[4];   // channel coordinate of each chunk in the A tensor
  const uint8_t* wblob;
  void* ps_out;
  long long* dbg;        // optional timeline buffer (block 0 only): [role][event] clock64 stamps
  TcEntry e[TC_MAX_ENTRIES];
  TcOutGroup g[TC_MAX_GROUPS];
  TcChunk ck[TC_MAX_CHUNKS];
};

#define TC_STAMP(role, idx)                                                                        \
  do {                                                                                             \
    if (kDbg && dbg != nullptr && blockIdx.x == 0 && (idx) < 32) dbg[(role) * 32 + (idx)] = clock64();      \
  } while (0)

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_nc(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
// MMA without a memory clobber: ordering against the barrier waits / commits comes from `volatile`
__device__ __forceinline__ void umma_f16_ss_nc(uint32_t tmem_d, uint32_t adesc_lo, uint32_t bdesc_lo, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(adesc_lo), "r"(bdesc_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u));
}

// GELU of 16 accumulator columns (gelu_fast of kernels_generic.cuh), written stage by stage over 8 values so that
// eight independent chains (two MUFU each) are in flight.  As an out-of-line call every element paid the call
// ABI's register save / restore around a 168-register caller (measured: 13 k cycles per tile for 96 GELU columns).
__device__ __forceinline__ void tc_gelu16(float (&f)[16]) {
#pragma unroll
  for (int h = 0; h < 16; h += 8) {
    float x[8], t[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fabsf(f[h + j]) * 0.70710678118654752440f;
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = __fdividef(1.f, fmaf(0.3275911f, x[j], 1.f));
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __expf(-x[j] * x[j]);
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = fmaf(t[j], 1.061405429f, -1.453152027f);
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = fmaf(q[j], t[j], 1.421413741f);
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = fmaf(q[j], t[j], -0.284496736f);
#pragma unroll
    for (int j = 0; j < 8; ++j) q[j] = fmaf(q[j], t[j], 0.254829592f);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float e = 1.f - q[j] * t[j] * x[j];   // erf(|v| / sqrt 2)
      f[h + j] = 0.5f * f[h + j] * (1.f + copysignf(e, f[h + j]));
    }
  }
}

// bias + residual + activation on 16 accumulator columns of one pixel.  NONE / RELU / LRELU are all
// max(v, v * slope) with slope 1 / 0 / s (set by the host), so the common path has no activation branch.
__device__ __forceinline__ void tc_epi_math16(const uint32_t (&v)[16], const float* __restrict__ bias_s, bool gelu, float slope,
                                              bool has_res, const uint4& u0, const uint4& u1, int res_after, float (&f)[16]) {
#pragma unroll
  for (int j = 0; j < 16; j += 4) {
    const float4 b4 = *reinterpret_cast<const float4*>(bias_s + j);
    f[j] = __uint_as_float(v[j]) + b4.x;
    f[j + 1] = __uint_as_float(v[j + 1]) + b4.y;
    f[j + 2] = __uint_as_float(v[j + 2]) + b4.z;
    f[j + 3] = __uint_as_float(v[j + 3]) + b4.w;
  }
  if (!has_res && !gelu) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], f[j] * slope);
    return;
  }
  float rv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) rv[j] = 0.f;
  if (has_res) {
    const __half2* h0 = reinterpret_cast<const __half2*>(&u0);
    const __half2* h1 = reinterpret_cast<const __half2*>(&u1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 a = __half22float2(h0[j]), c = __half22float2(h1[j]);
      rv[2 * j] = a.x; rv[2 * j + 1] = a.y; rv[8 + 2 * j] = c.x; rv[8 + 2 * j + 1] = c.y;
    }
  }
  if (res_after == 0) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] += rv[j];
  }
  if (gelu) {
    tc_gelu16(f);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], f[j] * slope);
  }
  if (res_after == 2) {          // gate: sigmoid(v) * res (FMEN's high-frequency attention, team03_fmen.py:72-74)
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = rv[j] * __fdividef(1.f, 1.f + __expf(-f[j]));
  } else if (res_after) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] += rv[j];
  }
}

// fp16 pack + store of 16 columns: swizzled staging row (mode 0) or fused PixelShuffle(4) (mode 1)
__device__ __forceinline__ void tc_epi_store16(const float (&f)[16], int mode, uint8_t* stage_row, int c0, int swz_row,
                                               bool valid, void* ps_out, int ps_fp32, int b, int y, int x, int H, int W) {
  if (mode == 0) {
#pragma unroll
    for (int hseg = 0; hseg < 2; ++hseg) {
      uint4 u;
      __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
      for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(f[hseg * 8 + 2 * j], f[hseg * 8 + 2 * j + 1]);
      const int chunk = (c0 >> 3) + hseg;
      *reinterpret_cast<uint4*>(stage_row + ((chunk ^ swz_row) << 4)) = u;
    }
  } else if (valid) {
    // column 16*c + 4*i + j of pixel (y,x) -> out[b, c, 4y+i, 4x+j]
    const int Ho = 4 * H, Wo = 4 * W;
    const int ch = c0 >> 4;
    const long long o0 = (((long long)b * 3 + ch) * Ho + 4 * y) * Wo + 4 * x;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long o = o0 + (long long)i * Wo;
      const float* ff = f + 4 * i;
      if (ps_fp32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(ps_out) + o) = make_float4(ff[0], ff[1], ff[2], ff[3]);
      } else {
        uint2 u;
        __half2* h = reinterpret_cast<__half2*>(&u);
        h[0] = __floats2half2_rn(ff[0], ff[1]);
        h[1] = __floats2half2_rn(ff[2], ff[3]);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ps_out) + o) = u;
      }
    }
  }
}

// One kernel for every network.  (A second instantiation without the GELU / border-class-bias code measured 1-2 %
// faster on RFDN; it also made a latent barrier bug - see the strip waits of the MMA issuers below - fail often
// enough to be found.  The split itself was not worth keeping.)
// kDbg: compiled with the timeline stamps and the timing-experiment switches; kBsrn: with the GELU and the border-class
// bias table (BSRN).  The production instantiations carry only what their network needs (the single-thread roles pay for
// every instruction-cache line; the same measure gave the fused chain kernel 3-6 %).
// kTail: with the residual / gate operand of group 0 and the pixel-shuffle store (LR_conv, upsampler, FMEN's gates).
template <bool kDbg, bool kBsrn, bool kTail>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmO0,
               const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmO2,
               const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[TC_MAX_SLOTS], tdone_bar[TC_DONE_BARS], tfull_bar[4], tempty_bar[4], w_bar;
  __shared__ int need_a_s[TC_MAX_SLOTS], need_b_s[TC_MAX_SLOTS];   // producer-private: last reader tiles of the strip in each slot
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[TC_MAX_GROUPS][64];
  __shared__ __align__(16) float bias9_s[9][64];
  __shared__ __align__(16) TcEntry ent_s[TC_MAX_ENTRIES];
  __shared__ TcOutGroup grp_s[TC_MAX_GROUPS];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* const dbg = kDbg ? p.dbg : nullptr;
  if (threadIdx.x == 0) TC_STAMP(0, 0);
  // 1024-byte aligned view of dynamic shared memory (SWIZZLE_128B atoms)
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
  uint8_t* const smem = smem_raw + pad;
  const uint32_t smem_base = raw_u32 + pad;

  // every parameter the role loops touch lives in a register from here on (the param bank would be
  // re-read after each asm volatile with a memory clobber otherwise)
  const int S = p.nslots, halo = p.halo, nchunks = p.nchunks, strip_bytes = p.strip_bytes, chunk_bytes = p.chunk_bytes;
  const int n_items = p.n_items, strips_x = p.strips_x, segs_y = p.segs_y, rows_per_item = p.rows_per_item;
  const int H = p.H, W = p.W, ring_off = p.ring_off, acc_cols = p.acc_cols, n_entries = p.n_entries, ngroups = p.ngroups;
  const int dbg_flags = kDbg ? p.dbg_flags : 0;
  const uint32_t NS = (uint32_t)p.acc_slots, ns_shift = NS == 4 ? 2u : 1u;

  auto decode = [&](int item, int& b, int& y0, int& y1, int& x0) {
    const int per_img = strips_x * segs_y;
    b = item / per_img;
    const int rem = item - b * per_img;
    const int seg = rem / strips_x;
    const int sx = rem - seg * strips_x;
    y0 = seg * rows_per_item;
    y1 = min(y0 + rows_per_item, H);
    x0 = sx * TC_TILE_PX;
  };

  // ---- producer state (warp 0, elected lane): strips are issued in two phases so that the first ring
  // fill overlaps the rest of the CTA set-up (TMEM allocation, bias / entry tables)
  const uint32_t strip_tx = (uint32_t)(nchunks * p.strip_px * 128);
  const int cc0 = p.chunk_c0[0], cc1 = p.chunk_c0[1], cc2 = p.chunk_c0[2], cc3 = p.chunk_c0[3];
  uint32_t pr_slot = 0, pr_seq = 0;
  int pr_item = blockIdx.x, pr_row = 0, pr_b = 0, pr_y0 = 0, pr_y1 = 0, pr_x0 = 0, pr_tile_base = 0;
  bool pr_open = false;
  auto produce = [&](uint32_t limit) {   // issue strips until `limit` have been issued in total
    while (pr_seq < limit) {
      if (!pr_open) {
        if (pr_item >= n_items) return;
        decode(pr_item, pr_b, pr_y0, pr_y1, pr_x0);
        pr_row = pr_y0 - halo;
        pr_open = true;
      }
      if (pr_seq >= (uint32_t)S) {
        // the slot still holds an older strip: every tile that reads it must have completed.  Its readers are
        // up to 1 + 2*halo consecutive tiles, issued alternately by the two MMA warps, so the last reader of
        // each warp is waited for (tile-done barriers, committed by the issuers after each tile).
        const int ua = need_a_s[pr_slot], ub = need_b_s[pr_slot];
        mbar_wait(&tdone_bar[ua & (TC_DONE_BARS - 1)], (uint32_t)(ua >> 3) & 1u);
        if (ub >= 0) mbar_wait(&tdone_bar[ub & (TC_DONE_BARS - 1)], (uint32_t)(ub >> 3) & 1u);
      }
      {
        const int nrows = pr_y1 - pr_y0, j = pr_row - (pr_y0 - halo);
        const int last = min(j, nrows - 1), first = max(0, j - 2 * halo);
        need_a_s[pr_slot] = pr_tile_base + last;
        need_b_s[pr_slot] = (last - 1 >= first) ? pr_tile_base + last - 1 : -1;
        // A strip with ONE reader tile (every strip of a 1x1 layer): the reader is one issuing thread, and the other issuing thread only watches its barrier to
        // stay in phase (see the MMA issuers).  Nothing kept the producer from recycling the slot twice before that
        // thread had looked at it - under multi-stream load (four requests in flight) it was lapped about once per
        // 10^5 forwards and then waited for a phase that had already gone by (mbarrier time-out).  The slot is therefore
        // only reused once the NEXT tile (the watcher's own, which it starts after having seen this strip) is done.
        // (The same holds for the first and the last strip of an item in the 3x3 layers: one reader tile.)
        if (need_b_s[pr_slot] < 0 && (last + 1 < nrows || pr_item + (int)gridDim.x < n_items)) need_b_s[pr_slot] = pr_tile_base + last + 1;
      }
      mbar_arrive_expect_tx(&full_bar[pr_slot], strip_tx);
      uint8_t* dst = smem + ring_off + pr_slot * strip_bytes;
      tma_load_4d(&tmA, &full_bar[pr_slot], dst, cc0, pr_x0 - halo, pr_row, pr_b);
      if (nchunks > 1) tma_load_4d(&tmA, &full_bar[pr_slot], dst + chunk_bytes, cc1, pr_x0 - halo, pr_row, pr_b);
      if (nchunks > 2) tma_load_4d(&tmA, &full_bar[pr_slot], dst + 2 * chunk_bytes, cc2, pr_x0 - halo, pr_row, pr_b);
      if (nchunks > 3) tma_load_4d(&tmA, &full_bar[pr_slot], dst + 3 * chunk_bytes, cc3, pr_x0 - halo, pr_row, pr_b);
      if (pr_row + TC_PREFETCH_ROWS < pr_y1 + halo) {   // warm L2 for the strip the ring cannot hold yet
        tma_prefetch_4d(&tmA, cc0, pr_x0 - halo, pr_row + TC_PREFETCH_ROWS, pr_b);
        if (nchunks > 1) tma_prefetch_4d(&tmA, cc1, pr_x0 - halo, pr_row + TC_PREFETCH_ROWS, pr_b);
        if (nchunks > 2) tma_prefetch_4d(&tmA, cc2, pr_x0 - halo, pr_row + TC_PREFETCH_ROWS, pr_b);
        if (nchunks > 3) tma_prefetch_4d(&tmA, cc3, pr_x0 - halo, pr_row + TC_PREFETCH_ROWS, pr_b);
      }
      TC_STAMP(1, pr_seq);
      ++pr_seq;
      if (++pr_slot == (uint32_t)S) pr_slot = 0;
      if (++pr_row >= pr_y1 + halo) { pr_open = false; pr_tile_base += pr_y1 - pr_y0; pr_item += gridDim.x; }
    }
  };

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmA);
      for (int i = 0; i < S; ++i) mbar_init(&full_bar[i], 1);
      for (int i = 0; i < TC_DONE_BARS; ++i) mbar_init(&tdone_bar[i], 1);
      for (int i = 0; i < 4; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], TC_EPI_WARPS); }
      mbar_init(&w_bar, 1);
      fence_mbar_init();
      // With programmatic dependent launch the weights (which never depend on the previous kernel) go before
      // the grid dependency is resolved; otherwise the strips the first tile needs are requested first (the TMA
      // queue is served in order and the 77 KB weight copy would delay them).
      if (p.pdl) {
        mbar_arrive_expect_tx(&w_bar, (uint32_t)p.w_bytes);
        bulk_load_1d(smem + p.w_off, p.wblob, (uint32_t)p.w_bytes, &w_bar);
        griddep_wait();
        produce((uint32_t)S);
      } else {
        produce((uint32_t)(1 + 2 * halo));
        mbar_arrive_expect_tx(&w_bar, (uint32_t)p.w_bytes);
        bulk_load_1d(smem + p.w_off, p.wblob, (uint32_t)p.w_bytes, &w_bar);
        produce((uint32_t)S);
      }
      tma_prefetch_desc(&tmO0);
      tma_prefetch_desc(&tmO1);
      tma_prefetch_desc(&tmO2);
    }
    __syncwarp();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  if (warp >= 2) {
    const int tid = threadIdx.x - 64;
    for (int i = tid; i < TC_MAX_GROUPS * 64; i += 32 * TC_EPI_WARPS) {
      const int g = i >> 6, c = i & 63;
      bias_s[g][c] = (g < ngroups && c < p.g[g].ncols) ? p.g[g].bias[c] : 0.f;
    }
    if (kBsrn && p.g[0].bias9 != nullptr)
      for (int i = tid; i < 9 * 64; i += 32 * TC_EPI_WARPS) bias9_s[i >> 6][i & 63] = p.g[0].bias9[i];
    if (tid < TC_MAX_ENTRIES) ent_s[tid] = p.e[tid];
    if (tid >= 32 && tid < 32 + TC_MAX_GROUPS) grp_s[tid - 32] = p.g[tid - 32];
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) {
    TC_STAMP(0, 1);
    griddep_launch_dependents();   // this grid is fully resident (<= 1 CTA per SM): let the next kernel set up
  }

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      produce(0xffffffffu);
      TC_STAMP(0, 2);
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2 + TC_EPI_WARPS) {
    // ================================ MMA issuers =================================
    // Two issuing threads (warp 1: even tiles, warp 10: odd tiles).  One thread cannot keep the tensor pipe
    // busy: between two tiles it spends ~0.8k cycles on barrier traffic (commits, strip / accumulator waits)
    // while the short MMA queue drains; with two threads that bookkeeping hides behind the other one's MMAs.
    if (elect_one()) {
      const uint32_t me = warp == 1 ? 0u : 1u;
      mbar_wait(&w_bar, 0);
      if (me == 0) TC_STAMP(0, 3);
      const uint32_t ring_base = smem_base + ring_off;
      const uint32_t w_lo = 0x10000u | ((smem_base + p.w_off) >> 4);   // descriptor low word: LBO = 1, start >> 4
      // strips are numbered q = 0, 1, ... in the order the producer issues them: slot = q % S, phase = (q / S) & 1.
      // `tslot/tpar` follow the top strip of the current tile; `seen` = strips this thread has already waited for
      uint32_t tslot = 0, tpar = 0, q_top = 0, seen = 0, t = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        int b, y0, y1, x0;
        decode(item, b, y0, y1, x0);
        const int nrows = y1 - y0;
        for (int r = 0; r < nrows; ++r, ++t) {
          {
            // Wait for the strips of this tile's row window that this thread has not seen yet - for EVERY tile,
            // also the ones the other issuing thread owns.  A parity wait is only valid one phase deep: with
            // halo = 0 and an odd ring (the 1x1 c5 layers: S = 3) a thread that only waited for its own tiles'
            // strips skipped every other phase of each slot, and a wait for strip q + 6 returned at once while
            // strip q + 3 was still in flight (TMA completions are not ordered).  The thread then ran one phase
            // ahead for the rest of the kernel and the CTA could exit with loads outstanding: the intermittent
            // "unspecified launch failure" at 16 x 270x480.
            uint32_t sl = tslot, pa = tpar;
            for (uint32_t q = q_top; q <= q_top + 2 * halo; ++q) {
              if (q >= seen) mbar_wait(&full_bar[sl], pa);
              if (++sl == (uint32_t)S) { sl = 0; pa ^= 1; }
            }
            seen = q_top + 2 * halo + 1;
          }
          if ((t & 1u) == me) {
            const uint32_t aslot = t & (NS - 1);
            if (t < 16) TC_STAMP(2, 2 * t);
            mbar_wait(&tempty_bar[aslot], ((t >> ns_shift) & 1) ^ 1);
            tc_fence_after_sync();
            const uint32_t d_base = tmem_base + aslot * acc_cols;
            const int ne = (dbg_flags & 1) ? 0 : n_entries;
            // descriptor low words of the (up to three) strips of this tile's row window
            uint32_t s1 = tslot + 1, s2 = tslot + 2;
            if (s1 >= (uint32_t)S) s1 -= S;
            if (s2 >= (uint32_t)S) s2 -= S;
            const uint32_t rb0 = 0x10000u | ((ring_base + tslot * strip_bytes) >> 4);
            const uint32_t rb1 = 0x10000u | ((ring_base + s1 * strip_bytes) >> 4);
            const uint32_t rb2 = 0x10000u | ((ring_base + s2 * strip_bytes) >> 4);
            uint4 en = *reinterpret_cast<const uint4*>(&ent_s[0]);
            for (int ei = 0; ei < ne; ++ei) {
              const uint4 e = en;
              if (ei + 1 < ne) en = *reinterpret_cast<const uint4*>(&ent_s[ei + 1]);   // next entry in flight
              const uint32_t row = e.x >> 28;
              const uint32_t a_lo = (row == 0 ? rb0 : (row == 1 ? rb1 : rb2)) + (e.x & 0x0fffffffu);
              const uint32_t b_lo = w_lo + e.y;
              const uint32_t steps = (e.w >> 16) & 15u;
              const uint32_t d = d_base + (e.w & 0xffffu);
              umma_f16_ss_nc(d, a_lo, b_lo, e.z, (e.w >> 31) ? 0u : 1u);
              if (steps > 1) umma_f16_ss_nc(d, a_lo + 2, b_lo + 2, e.z, 1u);
              if (steps > 2) umma_f16_ss_nc(d, a_lo + 4, b_lo + 4, e.z, 1u);
              if (steps > 3) umma_f16_ss_nc(d, a_lo + 6, b_lo + 6, e.z, 1u);
            }
            umma_commit(&tfull_bar[aslot]);
            umma_commit(&tdone_bar[t & (TC_DONE_BARS - 1)]);
            if (t < 16) TC_STAMP(2, 2 * t + 1);
          }
          ++q_top;
          if (++tslot == (uint32_t)S) { tslot = 0; tpar ^= 1; }
        }
        // the next item starts 2*halo strips further down the ring
        for (int k = 0; k < 2 * halo; ++k) {
          ++q_top;
          if (++tslot == (uint32_t)S) { tslot = 0; tpar ^= 1; }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue ====================================
    // 8 warps: warp w reads TMEM lanes 32*(w%4).. (hardware rule); the two warps of a lane quadrant split
    // the epilogue chunks (32 or 16 accumulator columns each) between them by chunk parity
    const int q = warp & 3;           // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2; // 0: warps 2..5, 1: warps 6..9
    const int m = q * 32 + lane;      // pixel of the tile == TMEM lane
    const bool store_thread = (warp == 2) && (lane == 0);
    const int ps_fp32 = p.ps_fp32;
    void* const ps_out = p.ps_out;
    const int ng = (dbg_flags & 2) ? 0 : ngroups;
    const int nck = ng;
    griddep_wait();   // residual reads, staging stores and the fused pixel-shuffle store touch global memory
    // per-group constants in registers (the group index is a compile-time constant in the unrolled loops)
    int gcol0[TC_MAX_GROUPS], gncols[TC_MAX_GROUPS], gmode[TC_MAX_GROUPS], gswz[TC_MAX_GROUPS];
    int gstage[TC_MAX_GROUPS], gstage_bytes[TC_MAX_GROUPS];
    bool ggelu[TC_MAX_GROUPS];
    float gslope[TC_MAX_GROUPS];
#pragma unroll
    for (int gi = 0; gi < TC_MAX_GROUPS; ++gi) {
      const TcOutGroup& g = grp_s[gi < ngroups ? gi : 0];
      gcol0[gi] = g.col0; gncols[gi] = gi < ngroups ? g.ncols : 0; gmode[gi] = g.mode; gswz[gi] = g.swizzle;
      gstage[gi] = g.stage_off; gstage_bytes[gi] = g.stage_bytes;
      ggelu[gi] = kBsrn && g.act == ACT_GELU; gslope[gi] = g.slope;
    }
    const bool g0_bias9 = kBsrn && grp_s[0].bias9 != nullptr;
    const bool g0_has_res = kTail && ng > 0 && grp_s[0].res != nullptr;
    const __half* const g0_res = g0_has_res ? grp_s[0].res + grp_s[0].res_coff : nullptr;
    const int g0_res_stride = grp_s[0].res_stride;
    uint32_t t = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      int b, y0, y1, x0;
      decode(item, b, y0, y1, x0);
      for (int y = y0; y < y1; ++y, ++t) {
        const int x = x0 + m;
        const bool valid = x < W;
        const long long pix = ((long long)b * H + y) * W + x;
        const uint32_t aslot = t & (NS - 1), sbuf = t & 1;
        // residual rows (group 0 only) are fetched before the accumulator is ready: their latency hides
        // behind the MMAs
        uint4 rpre[2][2];
        if (g0_has_res) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            rpre[k][0] = make_uint4(0, 0, 0, 0);
            rpre[k][1] = make_uint4(0, 0, 0, 0);
            const int c0 = (half + 2 * k) * 16;
            if (valid && c0 < gncols[0]) {
              const uint4* rp = reinterpret_cast<const uint4*>(g0_res + pix * g0_res_stride + c0);
              rpre[k][0] = rp[0];
              rpre[k][1] = rp[1];
            }
          }
        }
        // staging buffer `sbuf` was last read by the TMA store of tile t-2
        if (store_thread) tma_store_wait_read<1>();
        __syncwarp();   // bar.sync is a warp-aligned instruction: the store thread's lane must have rejoined its warp
        named_bar_sync(1, 32 * TC_EPI_WARPS);
        if (threadIdx.x == 64) TC_STAMP(3, 3 * t);
        mbar_wait(&tfull_bar[aslot], (t >> ns_shift) & 1);
        tc_fence_after_sync();
        if (threadIdx.x == 64) TC_STAMP(3, 3 * t + 1);
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + aslot * acc_cols;
        // tcgen05.ld is queued behind the MMAs already issued for the next tiles (in-order tensor pipe), so
        // this warp issues ALL of its TMEM loads for the tile back to back and waits once.  Within every
        // output group the two warps of a lane quadrant take alternate 16-column units.
        // (group, k) slots are compile-time; slot s+1's TMEM load is in flight while slot s is converted and
        // staged (two register buffers)
        uint32_t va[16], vb[16];
        auto slot_valid = [&](int s) { return (s >> 1) < ng && (half + 2 * (s & 1)) * 16 < gncols[s >> 1]; };
        auto slot_load = [&](int s, uint32_t (&v)[16]) { tmem_ld16_nc(taddr + gcol0[s >> 1] + (half + 2 * (s & 1)) * 16, v); };
        auto slot_run = [&](int s, const uint32_t (&v)[16]) {
          const int gi = s >> 1, c0 = (half + 2 * (s & 1)) * 16;
          const TcOutGroup& g = grp_s[gi];
          const bool has_res = gi == 0 && g0_has_res;
          uint8_t* stage_row = smem + gstage[gi] + sbuf * gstage_bytes[gi] + m * (gncols[gi] * 2);
          const int swz = gswz[gi] == 1 ? (m & 7) : (gswz[gi] == 2 ? ((m >> 1) & 3) : (gswz[gi] == 3 ? ((m >> 2) & 1) : 0));
          float f[16];
          const float* bias_row = &bias_s[gi][c0];
          if (gi == 0 && g0_bias9)   // border class of this thread's pixel
            bias_row = &bias9_s[3 * (y == 0 ? 0 : (y == H - 1 ? 2 : 1)) + (x == 0 ? 0 : (x == W - 1 ? 2 : 1))][c0];
          tc_epi_math16(v, bias_row, ggelu[gi], gslope[gi], has_res, rpre[s & 1][0], rpre[s & 1][1], g.res_after, f);
          tc_epi_store16(f, kTail ? gmode[gi] : 0, stage_row, c0, swz, valid, ps_out, ps_fp32, b, y, x, H, W);
        };
        // With both MMA warps keeping the tensor pipe busy a tcgen05.ld takes ~0.7k cycles to come back, so
        // the loads of three slots are issued back to back and waited for once (two rounds cover all six)
        uint32_t vc[16];
#pragma unroll
        for (int s0 = 0; s0 < 2 * TC_MAX_GROUPS; s0 += 3) {
          if (!(slot_valid(s0) || slot_valid(s0 + 1) || slot_valid(s0 + 2))) continue;
          if (slot_valid(s0)) slot_load(s0, va);
          if (slot_valid(s0 + 1)) slot_load(s0 + 1, vb);
          if (slot_valid(s0 + 2)) slot_load(s0 + 2, vc);
          tmem_ld_wait();
          if (slot_valid(s0)) slot_run(s0, va);
          if (slot_valid(s0 + 1)) slot_run(s0 + 1, vb);
          if (slot_valid(s0 + 2)) slot_run(s0 + 2, vc);
        }
        if (threadIdx.x == 64 && t < 8) TC_STAMP(1, 16 + 2 * t);
        if (threadIdx.x == 64 && t < 8) TC_STAMP(1, 17 + 2 * t);
        // accumulator slot drained: hand it back to the MMA warp
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[aslot]);
        fence_proxy_async_smem();
        __syncwarp();
        named_bar_sync(2, 32 * TC_EPI_WARPS);
        if (store_thread) {
          if (nck > 0) {
            if (grp_s[0].mode == 0) tma_store_4d(&tmO0, smem + grp_s[0].stage_off + sbuf * grp_s[0].stage_bytes, 0, x0, y, b);
            if (ngroups > 1 && grp_s[1].mode == 0)
              tma_store_4d(&tmO1, smem + grp_s[1].stage_off + sbuf * grp_s[1].stage_bytes, 0, x0, y, b);
            if (ngroups > 2 && grp_s[2].mode == 0)
              tma_store_4d(&tmO2, smem + grp_s[2].stage_off + sbuf * grp_s[2].stage_bytes, 0, x0, y, b);
          }
          tma_store_commit();
          TC_STAMP(3, 3 * t + 2);
        }
      }
    }
    if (store_thread) {
      tma_store_wait_all<0>();
      TC_STAMP(0, 4);
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (threadIdx.x == 0) TC_STAMP(0, 5);
}

}  // namespace esr
