// Implicit-GEMM convolution on tcgen05 tensor cores (sm_100a), fp16 in / fp32 accumulate in TMEM.
//
// GEMM view of one output tile: M = 128 consecutive pixels of one image row, N = output channels,
// K = taps x input channels.  Activations are NHWC fp16 with 64-channel (128-byte) pixels, so an
// image-row strip [130 px x 64 ch] lands in shared memory through ONE TMA box load in exactly the
// K-major SWIZZLE_128B layout tcgen05.mma wants (one pixel = one 128-byte row).  The 3x3 taps are
// not materialised: tap (dy,dx) is the same strip ring with the descriptor start address moved by
// dy strips and dx rows, so every input row is read from L2 once per strip (im2col-free, and the
// zero padding is TMA out-of-bounds fill).  A persistent CTA walks down a column strip keeping a
// ring of row strips in shared memory; weights are staged once per CTA as pre-swizzled blocks.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2..5 = epilogue (TMEM -> registers -> bias/residual/activation -> fp16 -> swizzled staging
// -> TMA store, or the fused pixel-shuffle store of the network tail).  Two TMEM accumulator slots
// let the epilogue of tile t overlap the MMAs of tile t+1.
#pragma once
#include "tc_common.cuh"

namespace esr {

constexpr int TC_MAX_ENTRIES = 16;
constexpr int TC_TILE_PX = 128;
constexpr int TC_MAX_SLOTS = 8;
constexpr int TC_THREADS = 192;

struct TcEntry {
  int16_t row;     // strip of the tile's row window (0 .. 2*halo)
  int16_t px_off;  // first pixel of the strip used by this tap (dx + halo)
  int16_t chunk;   // 64-channel chunk of the strip
  int16_t nsteps;  // K=16 steps issued (ceil(real channels / 16))
  int32_t b_off;   // byte offset of the [n x 64] pre-swizzled weight block
  int16_t n;       // MMA N (multiple of 16)
  int16_t dcol;    // first accumulator column
  int32_t first;   // 1: first MMA of the entry overwrites the accumulator columns
};

struct TcOutGroup {
  int32_t col0;        // first accumulator column of the group
  int32_t ncols;       // columns stored (multiple of 8, <= 64)
  int32_t act;         // Act
  float slope;
  int32_t res_after;   // residual added after (1) or before (0) the activation
  int32_t res_stride;  // residual: elements per pixel
  int32_t res_coff;
  int32_t mode;        // 0: TMA store NHWC fp16   1: pixel-shuffle x4 store into NCHW output
  int32_t swizzle;     // staging layout of the store tensor map (1 = SWIZZLE_128B, 0 = linear)
  int32_t stage_off;   // smem offset of the two staging buffers
  int32_t stage_bytes; // bytes of one staging buffer
  int32_t pad_;
  const float* bias;   // [ncols]
  const __half* res;   // nullptr = none
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
  int32_t n_entries, ngroups;
  int32_t tmem_cols;     // TMEM allocation (power of two >= 2*acc_cols)
  int32_t acc_cols;      // columns of one accumulator slot
  int32_t w_off, w_bytes, ring_off;
  int32_t shift_mode;    // 0: base_offset 0 for shifted starts, 1: base_offset = row phase
  int32_t ps_fp32;
  int32_t chunk_c0[4];   // channel coordinate of each chunk in the A tensor
  const uint8_t* wblob;
  void* ps_out;
  TcEntry e[TC_MAX_ENTRIES];
  TcOutGroup g[2];
};

__device__ __forceinline__ void tc_decode_item(const TcParams& p, int item, int& b, int& y0, int& y1, int& x0) {
  const int per_img = p.strips_x * p.segs_y;
  b = item / per_img;
  const int rem = item - b * per_img;
  const int seg = rem / p.strips_x;
  const int sx = rem - seg * p.strips_x;
  y0 = seg * p.rows_per_item;
  y1 = min(y0 + p.rows_per_item, p.H);
  x0 = sx * TC_TILE_PX;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmO0,
               const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t full_bar[TC_MAX_SLOTS], empty_bar[TC_MAX_SLOTS], tfull_bar[2], tempty_bar[2], w_bar;
  __shared__ uint32_t tmem_base_s;
  __shared__ float bias_s[2][64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // 1024-byte aligned view of dynamic shared memory (SWIZZLE_128B atoms)
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
  uint8_t* smem = smem_raw + pad;
  const uint32_t smem_base = raw_u32 + pad;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmO0);
    tma_prefetch_desc(&tmO1);
    for (int i = 0; i < p.nslots; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    mbar_init(&w_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  for (int i = threadIdx.x; i < 128; i += blockDim.x) {
    const int g = i >> 6, c = i & 63;
    bias_s[g][c] = (g < p.ngroups && c < p.g[g].ncols) ? p.g[g].bias[c] : 0.f;
  }
  tc_fence_before_sync();
  __syncthreads();
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  const int S = p.nslots;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      mbar_arrive_expect_tx(&w_bar, (uint32_t)p.w_bytes);
      bulk_load_1d(smem + p.w_off, p.wblob, (uint32_t)p.w_bytes, &w_bar);
      const uint32_t strip_tx = (uint32_t)(p.nchunks * p.strip_px * 128);
      uint32_t seq = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int b, y0, y1, x0;
        tc_decode_item(p, item, b, y0, y1, x0);
        for (int row = y0 - p.halo; row < y1 + p.halo; ++row, ++seq) {
          const uint32_t slot = seq % S, par = (seq / S) & 1;
          mbar_wait(&empty_bar[slot], par ^ 1);
          mbar_arrive_expect_tx(&full_bar[slot], strip_tx);
          for (int c = 0; c < p.nchunks; ++c)
            tma_load_4d(&tmA, &full_bar[slot], smem + p.ring_off + slot * p.strip_bytes + c * p.chunk_bytes,
                        p.chunk_c0[c], x0 - p.halo, row, b);
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      mbar_wait(&w_bar, 0);
      const uint32_t ring_base = smem_base + p.ring_off;
      const uint32_t w_base = smem_base + p.w_off;
      uint32_t waited = 0, released = 0, seq_base = 0, t = 0;
      for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
        int b, y0, y1, x0;
        tc_decode_item(p, item, b, y0, y1, x0);
        const int nrows = y1 - y0;
        for (int r = 0; r < nrows; ++r, ++t) {
          const uint32_t need_hi = seq_base + r + 2 * p.halo;
          while (waited <= need_hi) {
            mbar_wait(&full_bar[waited % S], (waited / S) & 1);
            ++waited;
          }
          const uint32_t aslot = t & 1;
          mbar_wait(&tempty_bar[aslot], ((t >> 1) & 1) ^ 1);
          tc_fence_after_sync();
          const uint32_t d_base = tmem_base + aslot * p.acc_cols;
          for (int ei = 0; ei < p.n_entries; ++ei) {
            const TcEntry e = p.e[ei];
            const uint32_t sq = seq_base + r + e.row;
            const uint32_t a_addr =
                ring_base + (sq % S) * p.strip_bytes + e.chunk * p.chunk_bytes + e.px_off * 128;
            const uint32_t b_addr = w_base + e.b_off;
            const uint32_t idesc = umma_idesc_f16((uint32_t)e.n);
            const uint32_t bo = p.shift_mode ? (uint32_t)(e.px_off & 7) : 0u;
            for (int ks = 0; ks < e.nsteps; ++ks) {
              umma_f16_ss(d_base + e.dcol, umma_desc_sw128(a_addr + ks * 32, bo), umma_desc_sw128(b_addr + ks * 32),
                          idesc, (e.first && ks == 0) ? 0u : 1u);
            }
          }
          umma_commit(&tfull_bar[aslot]);
          const uint32_t limit = (r == nrows - 1) ? need_hi + 1 : seq_base + r + 1;
          while (released < limit) {
            umma_commit(&empty_bar[released % S]);
            ++released;
          }
        }
        seq_base += nrows + 2 * p.halo;
      }
    }
    __syncwarp();
  } else {
    // ================================ epilogue ====================================
    const int q = warp & 3;           // TMEM lane quadrant this warp may read
    const int m = q * 32 + lane;      // pixel of the tile == TMEM lane
    const bool issuer = (threadIdx.x == 64);
    uint32_t t = 0;
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      int b, y0, y1, x0;
      tc_decode_item(p, item, b, y0, y1, x0);
      for (int y = y0; y < y1; ++y, ++t) {
        const int x = x0 + m;
        const bool valid = x < p.W;
        const long long pix = ((long long)b * p.H + y) * p.W + x;
        const uint32_t aslot = t & 1, sbuf = t & 1;
        // staging buffer `sbuf` was last read by the TMA store of tile t-2
        if (issuer) tma_store_wait_read<1>();
        named_bar_sync(1, 128);
        mbar_wait(&tfull_bar[aslot], (t >> 1) & 1);
        tc_fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + aslot * p.acc_cols;
        for (int gi = 0; gi < p.ngroups; ++gi) {
          const TcOutGroup& g = p.g[gi];
          uint8_t* stage = smem + g.stage_off + sbuf * g.stage_bytes;
          const int row_bytes = g.ncols * 2;
          for (int c0 = 0; c0 < g.ncols; c0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + g.col0 + c0, v);
            tmem_ld_wait();
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + bias_s[gi][c0 + j];
            if (g.res != nullptr) {
              float rv[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) rv[j] = 0.f;
              if (valid) {
                const __half* rp = g.res + pix * g.res_stride + g.res_coff + c0;
                const uint4 u0 = *reinterpret_cast<const uint4*>(rp);
                const uint4 u1 = *reinterpret_cast<const uint4*>(rp + 8);
                const __half2* h0 = reinterpret_cast<const __half2*>(&u0);
                const __half2* h1 = reinterpret_cast<const __half2*>(&u1);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const float2 a = __half22float2(h0[j]);
                  const float2 c = __half22float2(h1[j]);
                  rv[2 * j] = a.x; rv[2 * j + 1] = a.y; rv[8 + 2 * j] = c.x; rv[8 + 2 * j + 1] = c.y;
                }
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                float z = f[j];
                if (!g.res_after) z += rv[j];
                z = (g.act == 1) ? (z >= 0.f ? z : z * g.slope) : z;
                if (g.res_after) z += rv[j];
                f[j] = z;
              }
            } else if (g.act == 1) {
#pragma unroll
              for (int j = 0; j < 16; ++j) f[j] = f[j] >= 0.f ? f[j] : f[j] * g.slope;
            }
            if (g.mode == 0) {
              // two 16-byte chunks of this pixel's row in the staging tile
#pragma unroll
              for (int hseg = 0; hseg < 2; ++hseg) {
                if (c0 + hseg * 8 >= g.ncols) break;
                uint4 u;
                __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
                for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(f[hseg * 8 + 2 * j], f[hseg * 8 + 2 * j + 1]);
                const int chunk = (c0 >> 3) + hseg;
                const int pos = g.swizzle ? (chunk ^ (m & 7)) : chunk;
                *reinterpret_cast<uint4*>(stage + m * row_bytes + pos * 16) = u;
              }
            } else if (valid) {
              // fused PixelShuffle(4): column 16*c + 4*i + j of pixel (y,x) -> out[b, c, 4y+i, 4x+j]
              const int ch = c0 >> 4;
              const int Ho = 4 * p.H, Wo = 4 * p.W;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const long long o = (((long long)b * 3 + ch) * Ho + 4 * y + i) * Wo + 4 * x;
                if (p.ps_fp32) {
                  *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.ps_out) + o) =
                      make_float4(f[4 * i], f[4 * i + 1], f[4 * i + 2], f[4 * i + 3]);
                } else {
                  uint2 u;
                  __half2* h = reinterpret_cast<__half2*>(&u);
                  h[0] = __floats2half2_rn(f[4 * i], f[4 * i + 1]);
                  h[1] = __floats2half2_rn(f[4 * i + 2], f[4 * i + 3]);
                  *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.ps_out) + o) = u;
                }
              }
            }
          }
        }
        // accumulator slot drained: hand it back to the MMA warp
        tc_fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[aslot]);
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (issuer) {
          for (int gi = 0; gi < p.ngroups; ++gi) {
            const TcOutGroup& g = p.g[gi];
            if (g.mode != 0) continue;
            tma_store_4d(gi == 0 ? &tmO0 : &tmO1, smem + g.stage_off + sbuf * g.stage_bytes, 0, x0, y, b);
          }
          tma_store_commit();
        }
      }
    }
    if (issuer) tma_store_wait_all<0>();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
}

}  // namespace esr
