// Fused chain of 3x3 convolutions on tcgen05 (sm_100a): the stacked 3x3 layers of one distillation block
// (RFDB c1_r+d .. c4, IMDB conv1..4, RLFB c1_r..c3_r) or of the network tail (LR_conv + upsampler) as ONE
// persistent launch.  Activations stay in shared memory between the layers of a chain; only what a
// neighbouring CTA or a later kernel needs goes to global memory.
//
// Reference: models/rfdn_baseline/block.py:148-166 (RFDB.forward), models/basicblock.py:259-265 (IMDBlock),
// models/team04_rlfn.py:109-122 (RLFB), models/rfdn_baseline/RFDN.py:37-39 (LR_conv, upsampler).
//
// Work decomposition.  The image is cut into bands of R = 4 rows and column strips of 128 pixels; one CTA owns
// one (band, strip) patch for ALL layers of the chain, the strips of a band form one thread-block cluster.  A
// 3x3 layer needs one more row / column of its input on every side, so the patch of layer l is shifted UP by
// one row per layer (layer l of the band at y0 produces rows [y0 - l, y0 - l + R)): every vertical dependency
// then points to the band above (two halo rows per layer, fetched from global memory behind a flag), never
// below, and bands can be processed top-down by a persistent grid of any size without deadlock.  The left /
// right halo pixel columns are exchanged inside the cluster through distributed shared memory.
//
// Shared-memory row ring: R + 2 slots of [144 px x 64 ch] fp16 (SWIZZLE_128B, one pixel = one 128-byte row; tile
// pixel 0 sits at position 8 so that every TMA box starts on a 1024-byte swizzle atom).  Slot i holds input row
// i of the current layer (rows 0, 1 = halo from the band above, 2.. = this CTA's own rows).  Output row j of a
// layer reads slots j, j+1, j+2 and is written IN PLACE into slot j+2 by the epilogue (its old content is dead by
// then), so it becomes input row j+2 of the next layer without ever leaving the SM.
//
// MMA formulation ("row stationary"): input row i contributes to output rows i-2, i-1, i through the tap rows
// dy = +1, 0, -1.  Per (dx, 16-channel K step) ONE tcgen05.mma with A = slot i (shifted by dx pixels) and
// B = [W(+1,dx) | W(0,dx) | W(-1,dx)] (N = 3 x Np columns) feeds the accumulators of three output rows at once:
// N = 192 instead of 64 per instruction, which makes the SS-mode MMA math bound (32 + N/4 cycles of shared-memory
// operand fetch against N/2 cycles of tensor math) and halves the number of instructions one thread must issue.
// The weights of a layer live in three parts ordered by dy (atoms interleaved over dx: SBO = 3 KB) so that the
// stacked B operand is one strided descriptor; the parts are replaced one by one while the last steps of the
// previous layer still run (part +1 is dead after step 2, part 0 after step 1, part -1 after step 0).
//
// Warp roles (384 threads, 1 CTA / SM): warp 0 TMA producer (input rows, halo rows, weights, flag polling),
// warp 1 MMA issuer + TMEM owner, warps 2..9 epilogue (TMEM -> bias / residual / activation -> fp16 -> ring slot,
// staging or pixel-shuffle store; the two edge pixels also go to the neighbour CTAs' rings), warp 10 store warp
// (TMA stores of halo rows / staged groups, flag release).
#pragma once
#include "conv_tc.cuh"

namespace esr {

constexpr int CH_MAX_LAYERS = 6;
constexpr int CH_R = 4;                            // output rows per band
constexpr int CH_SLOTS = CH_R + 2;
constexpr int CH_SLOT_PX = 144;
constexpr int CH_SLOT_BYTES = CH_SLOT_PX * 128;    // 18432
constexpr int CH_PX0 = 8;                          // ring position of tile pixel 0
constexpr int CH_THREADS = 384;
constexpr int CH_MAX_MAPS = 20;
constexpr int CH_W_BYTES = 3 * 64 * 384;           // three dy parts of a 64-column layer
constexpr int CH_CTR_BYTES = 32 * 128;

struct ChLayer {
  int32_t np;           // accumulator columns per output row (multiple of 16, <= 64)
  int32_t ksteps;       // K / 16 of the convolution MMAs
  int32_t ctr_n;        // columns of the centre-tap-only block (distillation 1x1), 0 = none
  int32_t part_bytes;   // np * 384
  int32_t w_goff;       // byte offset of the layer's weights inside the chain blob: 3 parts, then the centre block
  int32_t ring_out;     // group 0 is written in place into the row ring (every layer but the last)
  int32_t res_smem;     // add the layer's own input (centre pixel) before the activation (RFDB `+ input`)
  int32_t n0;           // group 0: accumulator columns [0, n0)
  float slope0;
  int32_t mode0;        // last layer: 0 = staging + TMA store, 1 = fused PixelShuffle(4) store
  int32_t swz0;         // staging swizzle of group 0 (last layer, mode 0)
  int32_t res_stride, res_coff, res_after;
  int32_t n1;           // group 1 (0 = none): n1 columns, always staged
  int32_t g1_ctr;       // 1: group 1 = the centre block's accumulator, 0: accumulator columns [col1, col1 + n1)
  int32_t col1;
  float slope1;
  int32_t swz1;
  int32_t map_in;       // tensor map (box 128 px) of the layer's input buffer; map_in + 1 = the same with box 8 px
  int32_t map_out;      // tensor map (box 128 px x 64 ch) of the global copy of group 0 (ring layers) / staged group 0
  int32_t map_g1;       // tensor map of staged group 1
  int32_t acc_col;      // first TMEM column of the layer's accumulators (output row j at acc_col + j * np)
  const __half* res;    // residual from global memory (LR_conv + fea), nullptr = none
};

struct ChainParams {
  int32_t B, H, W, n_layers;
  int32_t strips, nbands, n_items;     // item = one band of one image (all strips = one cluster)
  int32_t ring_off, w_off, ctr_off, stage_off, stage_bytes;
  int32_t tmem_cols, ctr_acc_col;      // consecutive layers of different width use disjoint accumulator regions: the
                                       // per-row "accumulator drained" hand-over only holds between equal layouts
  int32_t store_all;                   // debug: every ring row also goes to global memory
  int32_t ps_fp32;
  void* ps_out;
  int32_t* flags;                      // [n_items][strips][n_layers], zeroed before the launch
  const uint8_t* wblob;
  long long* dbg;
  ChLayer L[CH_MAX_LAYERS];
};
struct ChainMaps { CUtensorMap m[CH_MAX_MAPS]; };

// ---- cluster / distributed shared memory primitives -------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_nid_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cl(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait_cl(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cl(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("esr chain: cluster mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, smem_u32(bar), parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// MMA with explicit descriptor high words (the stacked B operand uses SBO = 3 KB)
__device__ __forceinline__ void umma_f16_ss_hi(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(a_hi), "r"(b_hi));
}

#define CH_STAMP(role, idx)                                                                        \
  do {                                                                                             \
    if (dbg != nullptr && blockIdx.x == 0 && (idx) < 64) dbg[(role) * 64 + (idx)] = clock64();      \
  } while (0)

__global__ void __launch_bounds__(CH_THREADS, 1)
conv_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t tma_full[CH_SLOTS], sd[CH_SLOTS], ready[CH_R], wrote[CH_R], nfree[CH_R], wfull[4], sfree[2], ring_read;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[CH_MAX_LAYERS][128];   // [0,64): group 0, [64,128): group 1

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* const dbg = p.dbg;
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
  uint8_t* const smem = smem_raw + pad;
  const uint32_t smem_base = raw_u32 + pad;

  const int R = CH_R;
  const int H = p.H, W = p.W, nL = p.n_layers, strips = p.strips, nbands = p.nbands, n_items = p.n_items;
  const uint32_t rank = cluster_ctarank(), cid = cluster_id_x(), ncl = cluster_nid_x();
  const int strip = (int)rank, x0 = strip * TC_TILE_PX;
  const bool hasL = strip > 0, hasR = strip + 1 < strips;
  const int n_side = (hasL ? 1 : 0) + (hasR ? 1 : 0);
  const uint32_t ring_base = smem_base + p.ring_off;

  if (warp == 0 && elect_one()) {
    for (int i = 0; i < CH_SLOTS; ++i) { mbar_init(&tma_full[i], 1); mbar_init(&sd[i], 1); }
    for (int j = 0; j < CH_R; ++j) {
      mbar_init(&ready[j], TC_EPI_WARPS + 2 * n_side);
      mbar_init(&wrote[j], TC_EPI_WARPS);
      mbar_init(&nfree[j], n_side > 0 ? n_side : 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&wfull[i], 1);
    mbar_init(&sfree[0], 1); mbar_init(&sfree[1], 1);
    mbar_init(&ring_read, 1);
    fence_mbar_init();
    for (int i = 0; i < CH_MAX_MAPS; ++i) tma_prefetch_desc(&maps.m[i]);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  if (warp >= 2 && warp < 2 + TC_EPI_WARPS) {
    const int tid = threadIdx.x - 64;
    for (int i = tid; i < nL * 128; i += 32 * TC_EPI_WARPS) {
      const int l = i >> 7, c = i & 127;
      const ChLayer& Lr = p.L[l];
      float v = 0.f;
      if (c < 64) { if (c < Lr.n0) v = reinterpret_cast<const float*>(p.wblob + Lr.w_goff + 3 * Lr.part_bytes + CH_CTR_BYTES)[c]; }
      else if (c - 64 < Lr.n1) v = reinterpret_cast<const float*>(p.wblob + Lr.w_goff + 3 * Lr.part_bytes + CH_CTR_BYTES)[c];
      bias_s[l][c] = v;
    }
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();      // the peers' barriers must be initialised before any remote arrive
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) CH_STAMP(0, 0);

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      griddep_wait();
      uint32_t g = 0, ctr_cnt = 0, ring_cnt = 0;
      for (int item = (int)cid; item < n_items; item += (int)ncl) {
        const int img = item / nbands, band = item - img * nbands, y0 = band * R;
        for (int l = 0; l < nL; ++l, ++g) {
          const ChLayer& Lr = p.L[l];
          const int row0 = y0 - l - 1;             // image row of input row 0 of this layer
          const uint8_t* wsrc = p.wblob + Lr.w_goff;
          bool flags_ok = (l == 0) || band == 0;
          // Weight part p of this layer overwrites [p, p+1) * part_bytes of the weight area; the previous layer's part q
          // (dead once its step 2 - q has completed) occupies [q, q+1) * its own part_bytes.  Equal widths: part p
          // replaces part p.  A wider layer after a narrower one (the first layer of the next band after the last
          // layer of this one) must wait for every old part it covers.
          int need_step[3];
          {
            const int ob = g > 0 ? p.L[l == 0 ? nL - 1 : l - 1].part_bytes : 1;
            for (int pp = 0; pp < 3; ++pp) need_step[pp] = g > 0 ? 2 - min(2, ((pp + 1) * Lr.part_bytes - 1) / ob) : R + 1;
          }
          bool part_loaded[3] = {false, false, false};
          if (l == 0 && g > 0) mbar_wait(&ring_read, (ring_cnt - 1) & 1u);   // TMA stores out of the ring slots have read them
          for (int i = R + 1; i >= 0; --i) {
            // slot i, and the weight part whose last reader was step i, are free once step i of the previous
            // layer has completed
            if (g > 0) mbar_wait(&sd[i], (g - 1) & 1u);
            // ... and, where the previous layer's epilogue reads its input row back (block residual from the ring:
            // output row i-1 reads slot i), once that epilogue is through
            if (g > 0 && i >= 1 && i <= R && (l == 0 || i < 2) && p.L[l == 0 ? nL - 1 : l - 1].res_smem)
              mbar_wait(&wrote[i - 1], (g - 1) & 1u);
            for (int part = 0; part < 3; ++part) {   // part 0: dy = +1 (needed first), 1: dy = 0, 2: dy = -1
              if (part_loaded[part] || need_step[part] < i) continue;
              part_loaded[part] = true;
              mbar_arrive_expect_tx(&wfull[part], (uint32_t)Lr.part_bytes);
              bulk_load_1d(smem + p.w_off + part * Lr.part_bytes, wsrc + part * Lr.part_bytes, (uint32_t)Lr.part_bytes, &wfull[part]);
            }
            if (i == (g > 0 ? 1 : R + 1) && Lr.ctr_n > 0) {   // the centre block's last reader is step 1
              mbar_arrive_expect_tx(&wfull[3], (uint32_t)(Lr.ctr_n * 128));
              bulk_load_1d(smem + p.ctr_off, wsrc + 3 * Lr.part_bytes, (uint32_t)(Lr.ctr_n * 128), &wfull[3]);
              ++ctr_cnt;
            }
            if (l == 0 || i < 2) {
              if (!flags_ok) {
                // halo rows of the band above: its stores of layer l-1 (strips s-1, s, s+1: the halo row includes the
                // corner pixels) must have reached global memory
                const int* f = p.flags + ((size_t)(item - 1) * strips) * nL + (l - 1);
                const long long t0 = clock64();
                for (int s = max(strip - 1, 0); s <= min(strip + 1, strips - 1); ++s) {
                  while (ld_acquire_gpu(f + (size_t)s * nL) == 0) {
                    if (clock64() - t0 > 4000000000LL) {
                      printf("esr chain: flag timeout block %d item %d layer %d strip %d\n", blockIdx.x, item, l, s);
                      __trap();
                    }
                  }
                }
                fence_proxy_async_all();
                flags_ok = true;
              }
              const CUtensorMap* m128 = &maps.m[Lr.map_in];
              const CUtensorMap* m8 = &maps.m[Lr.map_in + 1];
              uint8_t* dst = smem + p.ring_off + i * CH_SLOT_BYTES;
              mbar_arrive_expect_tx(&tma_full[i], (uint32_t)CH_SLOT_BYTES);
              tma_load_4d(m128, &tma_full[i], dst + CH_PX0 * 128, 0, x0, row0 + i, img);
              tma_load_4d(m8, &tma_full[i], dst, 0, x0 - 8, row0 + i, img);
              tma_load_4d(m8, &tma_full[i], dst + (CH_PX0 + TC_TILE_PX) * 128, 0, x0 + TC_TILE_PX, row0 + i, img);
            }
          }
          if (Lr.ring_out) ++ring_cnt;
        }
      }
      (void)ctr_cnt;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      uint32_t g = 0, item_cnt = 0, ctr_cnt = 0;
      const uint32_t w_base = smem_base + p.w_off, ctr_base = smem_base + p.ctr_off;
      const uint32_t HI_A = 0x40004040u, HI_B3 = 0x400040C0u;   // SBO 1024 / 3072 bytes, version 1, SWIZZLE_128B
      for (int item = (int)cid; item < n_items; item += (int)ncl, ++item_cnt) {
        for (int l = 0; l < nL; ++l, ++g) {
          const ChLayer& Lr = p.L[l];
          const int np = Lr.np, ks = Lr.ksteps, ctr_n = Lr.ctr_n;
          const uint32_t part_bytes = (uint32_t)Lr.part_bytes;
          for (int i = R + 1; i >= 0; --i) {
            if (l == 0 || i < 2) mbar_wait(&tma_full[i], (i < 2 ? g : item_cnt) & 1u);
            if (g > 0 && i >= 2) mbar_wait_cl(&ready[i - 2], (g - 1) & 1u);   // input row written (own + side pixels), accumulator drained
            if (i == R + 1) mbar_wait(&wfull[0], g & 1u);
            if (i == R) { mbar_wait(&wfull[1], g & 1u); if (ctr_n > 0) mbar_wait(&wfull[3], ctr_cnt & 1u); }
            if (i == R - 1) mbar_wait(&wfull[2], g & 1u);
            tc_fence_after_sync();
            const int jlo = max(i - 2, 0), jhi = min(i, R - 1), nj = jhi - jlo + 1;
            const int part0 = i >= 2 ? 0 : (i == 1 ? 1 : 2);
            const uint32_t a_base = ring_base + (uint32_t)i * CH_SLOT_BYTES + (CH_PX0 - 1) * 128;
            const uint32_t b_base = w_base + (uint32_t)part0 * part_bytes;
            const uint32_t d0 = tmem_base + (uint32_t)(Lr.acc_col + jlo * np);
            const uint32_t id_all = umma_idesc_f16((uint32_t)(np * nj)), id_one = umma_idesc_f16((uint32_t)np);
            const uint32_t id_rest = umma_idesc_f16((uint32_t)(np * (nj > 1 ? nj - 1 : 1)));
            const bool first = i >= 2;     // output row i-2 receives its first contribution in this step
#pragma unroll 1
            for (int dxi = 0; dxi < 3; ++dxi) {
              const uint32_t a_lo0 = 0x10000u | (((a_base + (uint32_t)dxi * 128u) & 0x3FFFFu) >> 4);
              const uint32_t b_lo0 = 0x10000u | (((b_base + (uint32_t)dxi * 1024u) & 0x3FFFFu) >> 4);
              for (int k = 0; k < ks; ++k) {
                const uint32_t a_lo = a_lo0 + 2u * k, b_lo = b_lo0 + 2u * k;
                if (first && dxi == 0 && k == 0) {
                  umma_f16_ss_hi(d0, a_lo, HI_A, b_lo, HI_B3, id_one, 0u);
                  if (nj > 1) umma_f16_ss_hi(d0 + (uint32_t)np, a_lo, HI_A, b_lo + (part_bytes >> 4), HI_B3, id_rest, 1u);
                } else {
                  umma_f16_ss_hi(d0, a_lo, HI_A, b_lo, HI_B3, id_all, 1u);
                }
              }
            }
            if (ctr_n > 0 && i >= 1 && i <= R) {   // centre-tap-only block (distillation 1x1) of output row i-1
              const uint32_t dc = tmem_base + (uint32_t)(p.ctr_acc_col + (i - 1) * 32);
              const uint32_t a_lo0 = 0x10000u | (((a_base + 128u) & 0x3FFFFu) >> 4);
              const uint32_t b_lo0 = 0x10000u | ((ctr_base & 0x3FFFFu) >> 4);
              const uint32_t idc = umma_idesc_f16((uint32_t)ctr_n);
              for (int k = 0; k < ks; ++k) umma_f16_ss_hi(dc, a_lo0 + 2u * k, HI_A, b_lo0 + 2u * k, HI_A, idc, k > 0 ? 1u : 0u);
            }
            umma_commit(&sd[i]);
            if (g < 8) CH_STAMP(1, g * 6 + i);
          }
          if (ctr_n > 0) ++ctr_cnt;
        }
      }
    }
    __syncwarp();
  } else if (warp < 2 + TC_EPI_WARPS) {
    // ================================ epilogue ====================================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int pos = CH_PX0 + m;                 // ring position of this thread's pixel
    const bool edgeL = hasL && m == 0, edgeR = hasR && m == TC_TILE_PX - 1;
    const int x = x0 + m;
    void* const ps_out = p.ps_out;
    const int ps_fp32 = p.ps_fp32;
    griddep_wait();
    uint32_t g = 0, stage_cnt = 0, ring_cnt = 0;
    for (int item = (int)cid; item < n_items; item += (int)ncl) {
      const int img = item / nbands, band = item - img * nbands, y0 = band * R;
      for (int l = 0; l < nL; ++l, ++g) {
        const ChLayer& Lr = p.L[l];
        const int np = Lr.np, n0 = Lr.n0, n1 = Lr.n1, ring_out = Lr.ring_out;
        const float slope0 = Lr.slope0, slope1 = Lr.slope1;
        const bool has_gres = Lr.res != nullptr;
        const bool staged = (n1 > 0) || (!ring_out && Lr.mode0 == 0);
        // the (up to) three 16-column units of this thread: accumulator units `half` and `half + 2`, centre unit `half`
        // kind: 0 none, 1 group 0, 2 group 1, 3 zero fill of the ring lanes beyond n0
        int ukind[3], uc0[3], ucol[3];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int c = (half + 2 * s) * 16;
          ucol[s] = c;
          if (c < n0) { ukind[s] = 1; uc0[s] = c; }
          else if (!Lr.g1_ctr && n1 > 0 && c >= Lr.col1 && c < Lr.col1 + n1) { ukind[s] = 2; uc0[s] = c - Lr.col1; }
          else if (ring_out) { ukind[s] = 3; uc0[s] = c; }
          else { ukind[s] = 0; uc0[s] = 0; }
        }
        ukind[2] = (Lr.g1_ctr && half * 16 < n1) ? 2 : 0;
        uc0[2] = half * 16;
        ucol[2] = half * 16;
        for (int j = R - 1; j >= 0; --j) {
          const int y = y0 - l + j;
          const bool valid = y >= 0 && y < H && x < W;
          const long long pix = ((long long)img * H + y) * W + x;
          uint4 rg[2][2];
          if (has_gres) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              rg[s][0] = make_uint4(0, 0, 0, 0); rg[s][1] = make_uint4(0, 0, 0, 0);
              if (valid && ukind[s] == 1) {
                const uint4* rp = reinterpret_cast<const uint4*>(Lr.res + pix * Lr.res_stride + Lr.res_coff + uc0[s]);
                rg[s][0] = rp[0]; rg[s][1] = rp[1];
              }
            }
          }
          mbar_wait(&sd[j], g & 1u);
          tc_fence_after_sync();
          if (warp == 2 && lane == 0 && n_side > 0) {   // my slot j+2 has been read: the neighbours may drop their edge pixels into it
            if (hasL) mbar_arrive_remote(mapa_u32(smem_u32(&nfree[j]), rank - 1));
            if (hasR) mbar_arrive_remote(mapa_u32(smem_u32(&nfree[j]), rank + 1));
          }
          uint32_t v[3][16];
          const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll
          for (int s = 0; s < 2; ++s)
            if (ukind[s] == 1 || ukind[s] == 2) tmem_ld16_nc(trow + (uint32_t)(Lr.acc_col + j * np + ucol[s]), v[s]);
          if (ukind[2]) tmem_ld16_nc(trow + (uint32_t)(p.ctr_acc_col + j * 32 + ucol[2]), v[2]);
          tmem_ld_wait();
          if (j == R - 1 && l > 0) mbar_wait(&ring_read, (ring_cnt - 1) & 1u);   // TMA stores out of the ring slots of the previous layer have read them
          const uint32_t buf = stage_cnt & 1u;
          if (staged) mbar_wait(&sfree[buf], ((stage_cnt >> 1) & 1u) ^ 1u);
          if ((edgeL || edgeR) && n_side > 0) mbar_wait_cl(&nfree[j], g & 1u);
          uint8_t* const slot_out = smem + p.ring_off + (j + 2) * CH_SLOT_BYTES;
          const uint8_t* const slot_in = smem + p.ring_off + (j + 1) * CH_SLOT_BYTES;
          uint8_t* const stage = smem + p.stage_off + buf * p.stage_bytes;
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int kind = ukind[s];
            if (kind == 0) continue;
            float f[16];
            const int c0 = uc0[s];
            if (kind == 3) {
#pragma unroll
              for (int e = 0; e < 16; ++e) f[e] = 0.f;
            } else {
              const bool g0 = kind == 1;
              uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0;
              bool has_res = false;
              if (g0 && Lr.res_smem) {
                const int ch = c0 >> 3;
                u0 = *reinterpret_cast<const uint4*>(slot_in + pos * 128 + (((ch) ^ (pos & 7)) << 4));
                u1 = *reinterpret_cast<const uint4*>(slot_in + pos * 128 + (((ch + 1) ^ (pos & 7)) << 4));
                has_res = true;
              } else if (g0 && has_gres) {
                u0 = rg[s][0]; u1 = rg[s][1];
                has_res = true;
              }
              tc_epi_math16(v[s], &bias_s[l][(g0 ? 0 : 64) + c0], false, g0 ? slope0 : slope1, has_res, u0, u1,
                            g0 ? Lr.res_after : 0, f);
            }
            if (kind == 3 || (kind == 1 && ring_out)) {
              uint4 o[2];
#pragma unroll
              for (int hs = 0; hs < 2; ++hs) {
                __half2* h2 = reinterpret_cast<__half2*>(&o[hs]);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  h2[e] = valid ? __floats2half2_rn(f[hs * 8 + 2 * e], f[hs * 8 + 2 * e + 1]) : __floats2half2_rn(0.f, 0.f);
                const int ch = (c0 >> 3) + hs;
                *reinterpret_cast<uint4*>(slot_out + pos * 128 + ((ch ^ (pos & 7)) << 4)) = o[hs];
              }
              if (edgeL || edgeR) {
                const int rpos = edgeL ? (CH_PX0 + TC_TILE_PX) : (CH_PX0 - 1);
                const uint32_t rbase = mapa_u32(smem_u32(slot_out) + (uint32_t)rpos * 128u, edgeL ? rank - 1 : rank + 1);
#pragma unroll
                for (int hs = 0; hs < 2; ++hs) st_cluster_v4(rbase + (uint32_t)((((c0 >> 3) + hs) ^ (rpos & 7)) << 4), o[hs]);
              }
            } else {
              const bool g0 = kind == 1;
              const int ncols = g0 ? n0 : n1;
              const int swzm = g0 ? Lr.swz0 : Lr.swz1;
              const int mode = g0 ? Lr.mode0 : 0;
              const int swz = swzm == 1 ? (m & 7) : (swzm == 2 ? ((m >> 1) & 3) : (swzm == 3 ? ((m >> 2) & 1) : 0));
              tc_epi_store16(f, mode, stage + m * (ncols * 2), c0, swz, valid, ps_out, ps_fp32, img, y, x, H, W);
            }
          }
          tc_fence_before_sync();
          if (edgeL || edgeR) {
            fence_proxy_async_all();
            mbar_arrive_remote(mapa_u32(smem_u32(&ready[j]), edgeL ? rank - 1 : rank + 1));
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { mbar_arrive(&ready[j]); mbar_arrive(&wrote[j]); }
          if (staged) ++stage_cnt;
          if (threadIdx.x == 64 && g < 8) CH_STAMP(2, g * 4 + j);
        }
        if (ring_out) ++ring_cnt;
      }
    }
  } else if (warp == 2 + TC_EPI_WARPS) {
    // ================================ store warp ==================================
    if (elect_one()) {
      uint32_t g = 0, stage_cnt = 0;
      for (int item = (int)cid; item < n_items; item += (int)ncl) {
        const int img = item / nbands, band = item - img * nbands, y0 = band * R;
        for (int l = 0; l < nL; ++l, ++g) {
          const ChLayer& Lr = p.L[l];
          const bool staged = (Lr.n1 > 0) || (!Lr.ring_out && Lr.mode0 == 0);
          for (int j = R - 1; j >= 0; --j) {
            const int y = y0 - l + j;
            const bool row_valid = y >= 0 && y < H;
            mbar_wait(&wrote[j], g & 1u);
            const uint32_t buf = stage_cnt & 1u;
            if (row_valid) {
              if (Lr.ring_out && (p.store_all || j >= R - 2))
                tma_store_4d(&maps.m[Lr.map_out], smem + p.ring_off + (j + 2) * CH_SLOT_BYTES + CH_PX0 * 128, 0, x0, y, img);
              if (!Lr.ring_out && Lr.mode0 == 0) tma_store_4d(&maps.m[Lr.map_out], smem + p.stage_off + buf * p.stage_bytes, 0, x0, y, img);
              if (Lr.n1 > 0) tma_store_4d(&maps.m[Lr.map_g1], smem + p.stage_off + buf * p.stage_bytes, 0, x0, y, img);
            }
            tma_store_commit();
            if (staged) {
              tma_store_wait_read<0>();
              mbar_arrive(&sfree[buf]);
              ++stage_cnt;
            }
            if (Lr.ring_out && j == R - 2) {
              // rows R-1 and R-2 (the halo of the band below) are on their way: publish them
              tma_store_wait_all<0>();
              fence_proxy_async_all();
              __threadfence();
              st_release_gpu(p.flags + ((size_t)item * strips + strip) * nL + l, 1);
              if (!p.store_all) mbar_arrive(&ring_read);   // ... and the ring slots they came from may be overwritten
            }
          }
          if (Lr.ring_out && p.store_all) {
            tma_store_wait_read<0>();
            mbar_arrive(&ring_read);
          }
        }
      }
      tma_store_wait_all<0>();
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();     // no CTA may leave while a peer can still write into its shared memory
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (threadIdx.x == 0) CH_STAMP(0, 1);
}

}  // namespace esr
