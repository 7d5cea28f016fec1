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
// Warp roles (640 threads, 1 CTA / SM): warp 0 TMA producer (input rows, halo rows, weights, flag polling),
// warp 1 MMA issuer + TMEM owner, warps 2..17 epilogue (TMEM -> bias / residual / activation -> fp16 -> ring slot,
// staging or pixel-shuffle store; the two edge pixels also go to the neighbour CTAs' rings with st.async), warp 18
// store warp (TMA stores of halo rows / staged groups, flag release), warp 19 relay ("slot read" to the neighbours).
#pragma once
#include "conv_tc.cuh"

namespace esr {

constexpr int CH_MAX_LAYERS = 6;
constexpr int CH_R = 4;                            // output rows per band
constexpr int CH_SLOTS = CH_R + 2;
constexpr int CH_SLOT_PX = 144;
constexpr int CH_SLOT_BYTES = CH_SLOT_PX * 128;    // 18432
constexpr int CH_PX0 = 8;                          // ring position of tile pixel 0
constexpr int CH_EPI_WARPS = 16;
constexpr int CH_THREADS = 32 * (CH_EPI_WARPS + 4);   // warp 0 producer, 1 MMA issuer, 2..17 epilogue, 18 store, 19 relay
constexpr int CH_MAX_MAPS = 20;
constexpr int CH_W_BYTES = 3 * 64 * 384;           // three dy parts of a 64-column layer
constexpr int CH_CTR_BYTES = 32 * 128;
constexpr int CH_IDENT_BYTES = 64 * 128;           // [64 x 64] fp16 identity, K-major SWIZZLE_128B (block residual as an MMA)

struct ChLayer {
  int32_t np;           // accumulator columns per output row (multiple of 16, <= 64)
  int32_t ksteps;       // K / 16 of the convolution MMAs
  int32_t ctr_n;        // columns of the centre-tap-only block (distillation 1x1), 0 = none
  int32_t part_bytes;   // np * 384
  int32_t w_goff;       // byte offset of the layer's weights inside the chain blob: 3 parts, then the centre block
  int32_t ring_out;     // group 0 is written in place into the row ring (every layer but the last)
  int32_t res_smem;     // block residual `+ input` (RFDB): issued as an exact identity MMA on the centre tap
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

// Optional last stage of a chain: a pointwise (1x1) convolution over the rows the last 3x3 layer has just produced
// (RFDB: c5 with the ESA entry columns over [d1|d2|d3|r4]; IMDB: conv1x1 over the distilled features).  Its input comes
// back from global memory (the chain's own staged stores + those of the band above, behind the same flags), its
// 64-channel output groups are staged in ring slots that are free by then.
struct ChPw {
  int32_t enabled;
  int32_t nchunks;        // 64-channel K chunks of the input buffer (1 or 2)
  int32_t ksteps;         // K steps per chunk
  int32_t n;              // accumulator columns (multiple of 16, <= 144)
  int32_t acc_col;        // rows alternate between TMEM columns acc_col and acc_col + n
  int32_t w_soff;         // byte offset of the weights inside the shared-memory weight area
  int32_t w_bytes;        // nchunks * n * 128
  int32_t w_early;        // 1: that range is dead during the last 3x3 layer (the weights arrive one layer ahead)
  int32_t map_in;         // tensor map (box 64 ch x 128 px) over the input buffer
  int32_t chunk_c0[2];    // channel coordinate of each chunk
  int32_t ngroups;
  int32_t g_col0[3], g_ncols[3], g_map[3];
  int32_t g_stage[3];     // 0 / 1: first / second 64-channel group, 2: the 16-channel group (ordinary staging buffer).
  int32_t nbuf;           // 64-channel groups are staged in ring slots 1 and 0.  nbuf = 2: rows alternate between two
  int32_t n64;            // buffers per group - one group: slot 1 / slot 0; two groups: slot 1 / weight area [0, 16K) and
                          // slot 0 / weight area [16K, 32K) (free while the pointwise weights sit behind them)
  float g_slope[3];
  int32_t res_stride, res_coff, res_after;
  int32_t from_smem;      // 1: the last 3x3 layer's group 0 is a channel range of this stage's input: its epilogue writes it
  int32_t fs_chunk;       //    straight into the A slot (chunk fs_chunk, channels fs_lane0 ..) instead of going through memory
  int32_t fs_lane0;
  int32_t bias_goff;      // byte offset of 160 bias floats (per accumulator column) inside the chain blob
  const __half* res;      // residual of group 0 from global memory, nullptr = none
  const uint8_t* w;       // device pointer: chunk c at w + c * n * 128 (K-major SWIZZLE_128B)
};

struct ChainParams {
  int32_t B, H, W, n_layers;
  int32_t strips, nbands, n_items;     // item = one band of one image (all strips = one cluster)
  int32_t ring_off, w_off, ctr_off, ident_off, stage_off, stage_bytes, ident_bytes;
  int32_t tmem_cols, ctr_acc_col;      // consecutive layers of different width use disjoint accumulator regions: the
                                       // per-row "accumulator drained" hand-over only holds between equal layouts
  int32_t store_all;                   // debug: every ring row also goes to global memory
  int32_t dbg_flags;                   // timing experiments (results are wrong): 1 no MMAs, 2 epilogue only synchronises, 4 no TMA stores
  int32_t ps_fp32;
  int32_t ps_u8;                       // the pixel-shuffle layer writes uint8 HWC (tensor2uint of the reference, utils_image.py:204-208)
  float ps_dr;                         // ... with this data range
  void* ps_out;
  int32_t* flags;                      // [n_items][strips][n_layers], zeroed before the launch
  int32_t* item_counter;               // bands beyond the first wave are handed out in order through this counter (zeroed with the flags)
  const uint8_t* wblob;
  const uint8_t* ident;                // device copy of the identity block
  long long* dbg;
  long long* dbg_blocks;               // optional: [gridDim.x][4] globaltimer stamps (kernel entry, set-up done, roles done, exit)
  ChLayer L[CH_MAX_LAYERS];
  ChPw pw;
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
// 16-byte store into a peer CTA's shared memory that completes 16 transaction bytes on a peer mbarrier (no fence
// and no separate arrive in the sending warp)
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, const uint4& v, uint32_t cluster_mbar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(cluster_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_mbar) : "memory");
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
// The same on a precomputed 32-bit shared address.  In a cluster launch every use of a __shared__ symbol's address
// costs an S2R of the CTA's cluster rank plus address arithmetic; the single-thread roles keep ONE laundered base
// address in a register and add compile-time offsets instead.
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t addr, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}\n"
      : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow_a(uint32_t addr, uint32_t parity) {
  const long long t0 = clock64();
  while (!mbar_try_wait_a(addr, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("esr chain: mbarrier timeout block %d thread %d bar %u parity %u\n", blockIdx.x, threadIdx.x, addr, parity);
      __trap();
    }
  }
}
__device__ __forceinline__ void mbar_wait_a(uint32_t addr, uint32_t parity) {
  if (mbar_try_wait_a(addr, parity)) return;
  mbar_wait_slow_a(addr, parity);
}
__device__ __forceinline__ void umma_commit_a(uint32_t addr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(addr) : "memory");
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

// n (0..4) consecutive K steps of one tap in ONE asm statement: the descriptors advance by 32 bytes (+2) per K step
// inside the statement, so ptxas keeps them in uniform registers (UIADD3.64) instead of moving four operands from
// vector registers (R2UR) and two constants (UMOV) for every MMA - ~4 instead of ~8 instructions per MMA for the
// single issuing thread, whose instruction rate is what bounds a step.  The first MMA uses `acc0`, the others accumulate.
__device__ __forceinline__ void umma_f16_ss_run(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                                uint32_t idesc, uint32_t acc0, int n) {
  asm volatile(
      "{\n\t.reg .pred p, q0, q1, q2, q3;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "setp.gt.s32 q0, %7, 0;\n\t"
      "setp.gt.s32 q1, %7, 1;\n\t"
      "setp.gt.s32 q2, %7, 2;\n\t"
      "setp.gt.s32 q3, %7, 3;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %6};\n\t"
      "@q0 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n\t"
      "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
      "@q1 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t"
      "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
      "@q2 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t"
      "add.s64 da, da, 2;\n\tadd.s64 db, db, 2;\n\t"
      "@q3 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n\t"
      "}\n"
      ::"r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(acc0), "r"(a_hi), "r"(b_hi), "r"(n));
}

__device__ __forceinline__ long long global_timer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define CH_STAMP(role, idx)                                                                        \
  do {                                                                                             \
    if (kDbg && dbg != nullptr && (int)blockIdx.x == (dbg_flags >> 8) && (idx) < 64) dbg[(role) * 64 + (idx)] = clock64();  /* dbg_flags >> 8: the block whose roles are stamped */      \
  } while (0)

// kPw: compiled with / without the optional pointwise last stage (its code costs the common instantiation registers)
// kU8: compiled with the uint8 pixel-shuffle epilogue (esr_forward_u8); measured: carrying that code in the common
// instantiation costs every chain 5 % (37.2 vs 35.5 us per RFDB chain)
// kDbg: compiled with the clock64 / globaltimer stamps and the timing-experiment switches (dbg_flags); the production
// instantiations carry none of that code
// kTail: compiled with the global residual / gate operand and the pixel-shuffle store (the tail chain and FMEN's chains);
// the block chains (RFDB, IMDB, RLFB) use the instantiation without them
// kCtr: compiled with the centre block (the distillation 1x1 of an RFDB stage: its weights, MMAs, accumulator, epilogue unit)
// kDyn: compiled with the dynamic band scheduling (more bands than clusters in flight); the single-pass instantiations
// (every cluster processes exactly one band: batch 1 at 256 x 256) drop the scheduling code from all five roles
template <bool kPw, bool kU8 = false, bool kDbg = false, bool kTail = true, bool kCtr = true, bool kDyn = true>
__global__ void __launch_bounds__(CH_THREADS, 1)
conv_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  // every mbarrier of the CTA in one array (the MMA issuer addresses them as base + compile-time offset)
  enum { B_TMA_FULL = 0, B_SD = B_TMA_FULL + CH_SLOTS, B_READY = B_SD + CH_SLOTS, B_WROTE = B_READY + CH_R, B_NFREE = B_WROTE + CH_R,
         B_WFULL = B_NFREE + CH_R, B_SFREE = B_WFULL + 4, B_RING_READ = B_SFREE + 2, B_IDENT = B_RING_READ + 1, B_PW_AFULL = B_IDENT + 1,
         B_PW_WFULL = B_PW_AFULL + 2, B_PW_DROW = B_PW_WFULL + 1, B_PW_SFREE = B_PW_DROW + CH_R, B_PW_DONE = B_PW_SFREE + 2,
         B_PW_DPRE = B_PW_DONE + 1, B_COUNT = B_PW_DPRE + 1 };
  __shared__ uint64_t bars[B_COUNT];
  uint64_t* const tma_full = bars + B_TMA_FULL; uint64_t* const sd = bars + B_SD; uint64_t* const ready = bars + B_READY;
  uint64_t* const wrote = bars + B_WROTE; uint64_t* const nfree = bars + B_NFREE; uint64_t* const wfull = bars + B_WFULL;
  uint64_t* const sfree = bars + B_SFREE; uint64_t& ring_read = bars[B_RING_READ]; uint64_t& ident_bar = bars[B_IDENT];
  uint64_t* const pw_afull = bars + B_PW_AFULL; uint64_t& pw_wfull = bars[B_PW_WFULL]; uint64_t* const pw_drow = bars + B_PW_DROW;
  uint64_t* const pw_sfree = bars + B_PW_SFREE; uint64_t& pw_done = bars[B_PW_DONE]; uint64_t& pw_dpre = bars[B_PW_DPRE];
  __shared__ uint64_t sched_bar[2];
  __shared__ int sched_item[2];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float bias_s[CH_MAX_LAYERS][128];   // [0,64): group 0, [64,128): group 1
  __shared__ __align__(16) float pw_bias_s[160];               // pointwise stage: per accumulator column

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long* const dbg = kDbg ? p.dbg : nullptr;
  long long* const dbg_blocks = kDbg ? p.dbg_blocks : nullptr;
  const int dbg_flags = kDbg ? p.dbg_flags : 0;
  if (kDbg && dbg_blocks != nullptr && threadIdx.x == 0) dbg_blocks[blockIdx.x * 4 + 0] = global_timer_ns();
  const uint32_t raw_u32 = smem_u32(smem_raw);
  const uint32_t pad = (1024u - (raw_u32 & 1023u)) & 1023u;
  uint8_t* const smem = smem_raw + pad;
  const uint32_t smem_base = raw_u32 + pad;

  const int R = CH_R;
  const int H = p.H, W = p.W, nL = p.n_layers, strips = p.strips, nbands = p.nbands, n_items = p.n_items;
  const bool has_pw = kPw && p.pw.enabled != 0;
  const int nG = nL + (has_pw ? 1 : 0);      // stages per band: the 3x3 layers (+ the pointwise stage)
  const uint32_t rank = cluster_ctarank(), cid = cluster_id_x(), ncl = cluster_nid_x();
  const int strip = (int)rank, x0 = strip * TC_TILE_PX;
  const bool hasL = strip > 0, hasR = strip + 1 < strips;
  const int n_side = (hasL ? 1 : 0) + (hasR ? 1 : 0);
  const uint32_t ring_base = smem_base + p.ring_off;
  // Band scheduling.  The first band of cluster c is band c; when there are more bands than clusters the rest are
  // handed out in increasing order through a global counter (the leader CTA's producer fetches the k+1-th band while
  // the k-th is loading and posts it into every CTA of the cluster).  A band only depends on lower-numbered bands, and
  // a lower-numbered band is always held by a running (or finished) cluster, so the kernel makes progress with any
  // number of co-resident clusters - also next to kernels of other streams.
  const bool dyn_sched = kDyn && n_items > (int)ncl;
  auto next_item = [&](uint32_t k) -> int {   // k-th band of this cluster, k >= 1
    if (!dyn_sched) return n_items;
    mbar_wait_cl(&sched_bar[k & 1u], ((k - 1) >> 1) & 1u);
    return *reinterpret_cast<volatile int*>(&sched_item[k & 1u]);
  };

  if (warp == 0 && elect_one()) {
    for (int i = 0; i < CH_SLOTS; ++i) { mbar_init(&tma_full[i], 1); mbar_init(&sd[i], 1); }
    for (int j = 0; j < CH_R; ++j) {
      mbar_init(&ready[j], CH_EPI_WARPS);
      mbar_init(&wrote[j], CH_EPI_WARPS);
      mbar_init(&nfree[j], n_side > 0 ? n_side : 1);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&wfull[i], 1);
    mbar_init(&sfree[0], 1); mbar_init(&sfree[1], 1);
    mbar_init(&ring_read, 1);
    mbar_init(&ident_bar, 1);
    mbar_init(&sched_bar[0], 1); mbar_init(&sched_bar[1], 1);
    mbar_init(&pw_afull[0], 1); mbar_init(&pw_afull[1], 1);
    mbar_init(&pw_wfull, 1);
    for (int j = 0; j < CH_R; ++j) mbar_init(&pw_drow[j], 1);
    mbar_init(&pw_sfree[0], 1); mbar_init(&pw_sfree[1], 1);
    mbar_init(&pw_done, 1);
    mbar_init(&pw_dpre, 1);
    fence_mbar_init();
    for (int i = 0; i < CH_MAX_MAPS; ++i) tma_prefetch_desc(&maps.m[i]);
  }
  if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
  if (warp >= 2 && warp < 2 + CH_EPI_WARPS) {
    const int tid = threadIdx.x - 64;
    for (int i = tid; i < nL * 128; i += 32 * CH_EPI_WARPS) {
      const int l = i >> 7, c = i & 127;
      const ChLayer& Lr = p.L[l];
      float v = 0.f;
      if (c < 64) { if (c < Lr.n0) v = reinterpret_cast<const float*>(p.wblob + Lr.w_goff + 3 * Lr.part_bytes + CH_CTR_BYTES)[c]; }
      else if (c - 64 < Lr.n1) v = reinterpret_cast<const float*>(p.wblob + Lr.w_goff + 3 * Lr.part_bytes + CH_CTR_BYTES)[c];
      bias_s[l][c] = v;
    }
    if (p.pw.enabled)
      for (int i = tid; i < 160; i += 32 * CH_EPI_WARPS) pw_bias_s[i] = reinterpret_cast<const float*>(p.wblob + p.pw.bias_goff)[i];
  }
  tc_fence_before_sync();
  __syncthreads();
  cluster_sync_all();      // the peers' barriers must be initialised before any remote arrive
  tc_fence_after_sync();
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) CH_STAMP(0, 0);
  if (kDbg && dbg_blocks != nullptr && threadIdx.x == 0) dbg_blocks[blockIdx.x * 4 + 1] = global_timer_ns();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      // the identity block (constant: does not depend on the previous kernel) - only chains with a block residual have it
      mbar_arrive_expect_tx(&ident_bar, kCtr ? (uint32_t)p.ident_bytes : 0u);
      if (kCtr && p.ident_bytes > 0) bulk_load_1d(smem + p.ident_off, p.ident, (uint32_t)p.ident_bytes, &ident_bar);
      griddep_wait();
      uint32_t g = 0, ctr_cnt = 0, ring_cnt = 0, items = 0;
      for (int item = (int)cid; item < n_items; ++items, item = next_item(items)) {
        const int img = item / nbands, band = item - img * nbands, y0 = band * R;
        if (has_pw && g > 0) {
          // the pointwise stage of the previous band: its MMAs have read their ring slots, its stores have read the two
          // staging slots, and its weights / accumulators are done with
          mbar_wait(&sd[0], (g - 1) & 1u);
          mbar_wait(&pw_done, (items - 1) & 1u);
        }
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
            for (int pp = 0; pp < 3; ++pp) need_step[pp] = g > 0 ? 2 - min(2, ((pp + 1) * Lr.part_bytes - 1) / ob) : R + 1 - pp;
            if (l == 0 && g > 0 && has_pw)   // the previous stage was the pointwise one (all of it has completed, see above)
              for (int pp = 0; pp < 3; ++pp) need_step[pp] = R + 1 - pp;
          }
          bool part_loaded[3] = {false, false, false};
          if (l == 0 && g > 0) mbar_wait(&ring_read, (ring_cnt - 1) & 1u);   // TMA stores out of the ring slots have read them
          auto load_row = [&](int i) {
            // the slot is also read back by the previous layer's epilogue where that layer adds a residual from the
            // ring (none does at present: the block residual is an identity MMA)
            if (!flags_ok) {
              // halo rows of the band above: its stores of layer l-1 (strips s-1, s, s+1: the halo row includes the
              // corner pixels) must have reached global memory
              const int* f = p.flags + ((size_t)(item - 1) * strips) * nL + (l - 1);
              const long long t0 = clock64();
              for (int s2 = max(strip - 1, 0); s2 <= min(strip + 1, strips - 1); ++s2) {
                while (ld_acquire_gpu(f + (size_t)s2 * nL) == 0) {
                  if (clock64() - t0 > 4000000000LL) {
                    printf("esr chain: flag timeout block %d item %d layer %d strip %d\n", blockIdx.x, item, l, s2);
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
          };
          for (int i = R + 1; i >= 0; --i) {
            // slot i, and the weight part whose last reader was step i, are free once step i of the previous
            // layer has completed
            if (g > 0) mbar_wait(&sd[i], (g - 1) & 1u);
            for (int part = 0; part < 3; ++part) {   // part 0: dy = +1 (needed first), 1: dy = 0, 2: dy = -1
              if (part_loaded[part] || need_step[part] < i) continue;
              part_loaded[part] = true;
              mbar_arrive_expect_tx(&wfull[part], (uint32_t)Lr.part_bytes);
              bulk_load_1d(smem + p.w_off + part * Lr.part_bytes, wsrc + part * Lr.part_bytes, (uint32_t)Lr.part_bytes, &wfull[part]);
            }
            if (kCtr && i == (g > 0 ? 1 : R) && Lr.ctr_n > 0) {   // the centre block's last reader is step 1
              mbar_arrive_expect_tx(&wfull[3], (uint32_t)(Lr.ctr_n * 128));
              bulk_load_1d(smem + p.ctr_off, wsrc + 3 * Lr.part_bytes, (uint32_t)(Lr.ctr_n * 128), &wfull[3]);
              ++ctr_cnt;
            }
            if (l == 0) load_row(i);     // the chain's input: every row comes from global memory
          }
          if (l == 0 && dyn_sched && rank == 0) {
            // leader: fetch this cluster's next band and post it to every CTA of the cluster
            const int nxt = (int)ncl + atomicAdd(p.item_counter, 1);
            const uint32_t slot = (items + 1) & 1u;
            for (uint32_t r = 0; r < (uint32_t)strips; ++r) {
              asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(mapa_u32(smem_u32(&sched_item[slot]), r)), "r"(nxt) : "memory");
              mbar_arrive_remote(mapa_u32(smem_u32(&sched_bar[slot]), r));
            }
          }
          // the halo rows of the later layers go last: they are needed last (steps 1 and 0), and the flag they wait
          // for must not hold up the weights
          if (l > 0) { load_row(1); load_row(0); }
          if (Lr.ring_out) ++ring_cnt;
          if (has_pw && l == nL - 1 && p.pw.w_early) {   // the pointwise weights go where the wider layers' parts used to be
            mbar_arrive_expect_tx(&pw_wfull, (uint32_t)p.pw.w_bytes);
            bulk_load_1d(smem + p.w_off + p.pw.w_soff, p.pw.w, (uint32_t)p.pw.w_bytes, &pw_wfull);
          }
        }
        if (has_pw) {
          // ---- pointwise stage: rows j = R-1 .. 0 of the last 3x3 layer's output rows, K chunks into ring slot pairs
          const ChPw& P = p.pw;
          if (!P.w_early) {
            mbar_wait(&sd[0], (g - 1) & 1u);
            mbar_arrive_expect_tx(&pw_wfull, (uint32_t)P.w_bytes);
            bulk_load_1d(smem + p.w_off + P.w_soff, P.w, (uint32_t)P.w_bytes, &pw_wfull);
          }
          const CUtensorMap* const mp = &maps.m[P.map_in];
          if (P.from_smem) {   // everything this CTA stored in the layers before the last one has landed
            mbar_wait(&pw_dpre, items & 1u);
            fence_proxy_async_all();
          }
          for (int j = R - 1; j >= 0; --j) {
            const int pr = (R - 1 - j) & 1;          // slot pair: 0 -> slots (5, 4), 1 -> slots (3, 2)
            if (j >= R - 2) {   // first use: the last 3x3 layer's steps that read those slots
              mbar_wait(&sd[5 - 2 * pr], (g - 1) & 1u);
              mbar_wait(&sd[4 - 2 * pr], (g - 1) & 1u);
            } else {            // second use: the MMAs of pointwise row j+2
              mbar_wait(&sd[j + 2], g & 1u);
            }
            if (!P.from_smem) {
              mbar_wait(&pw_drow[j], items & 1u);    // this CTA's stores up to row j of the last layer have landed
              fence_proxy_async_all();
            }
            const int y = y0 - (nL - 1) + j;
            mbar_arrive_expect_tx(&pw_afull[pr], (uint32_t)(P.nchunks * TC_TILE_PX * 128));
            for (int c = 0; c < P.nchunks; ++c)
              tma_load_4d(mp, &pw_afull[pr], smem + p.ring_off + (5 - 2 * pr - c) * CH_SLOT_BYTES + CH_PX0 * 128, P.chunk_c0[c], x0, y, img);
          }
          ++g;
        }
      }
      (void)ctr_cnt;
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    // One thread issues every MMA.  It measures ~100 cycles per instruction when the loop is branchy, and the tensor
    // pipe's queue is short, so (a) the MMAs of a step are issued from straight-line code and (b) the barrier waits of
    // the NEXT step are taken in the middle of the current step's MMAs: the pipe never drains while this thread polls.
    if (elect_one()) {
      const bool no_mma = (dbg_flags & 1) != 0;
      const uint32_t w_base = smem_base + p.w_off, ctr_base = smem_base + p.ctr_off, ident_base = smem_base + p.ident_off;
      mbar_wait(&ident_bar, 0);
      uint32_t bars_a;                                           // laundered: must not be rematerialised from the symbol
      asm volatile("mov.u32 %0, %1;" : "=r"(bars_a) : "r"(smem_u32(bars)));
      auto BA = [&](int idx) -> uint32_t { return bars_a + 8u * (uint32_t)idx; };
      const uint32_t HI_A = 0x40004040u, HI_B3 = 0x400040C0u;   // SBO 1024 / 3072 bytes, version 1, SWIZZLE_128B
      // position in the (band, layer, step) sequence
      int it_item = (int)cid, it_l = 0, it_i = R + 1;
      uint32_t it_g = 0, it_items = 0, it_ctr = 0;
      bool it_valid = it_item < n_items;
      uint32_t ctr_mask = 0;
      for (int l = 0; l < nL; ++l) if (kCtr && p.L[l].ctr_n > 0) ctr_mask |= 1u << l;
      const bool no_wait = (dbg_flags & 16) != 0, no_short = (dbg_flags & 32) != 0;   // timing experiments (results are wrong)
      auto wait_step = [&](int l, int i, uint32_t g, uint32_t items, uint32_t ctrc) {
        if (no_wait) return;
        if (kPw && l == nL) {   // pointwise stage: steps R+1, R are empty; step j < R = output row j
          if (i >= R) return;
          const int pr = (R - 1 - i) & 1;
          if (i == R - 1) mbar_wait_a(BA(B_PW_WFULL), items & 1u);
          mbar_wait_a(BA(B_PW_AFULL + pr), i < R - 2 ? 1u : 0u);     // each slot pair is filled twice per band
          if (p.pw.from_smem) mbar_wait_a(BA(B_READY + i), (g - 1) & 1u);   // ... and the last 3x3 layer's epilogue has added its channels
          if (i < R - 2) mbar_wait_a(BA(B_READY + i + 2), g & 1u);    // row i+2 used the same accumulator: drained
          return;
        }
        if (l == 0 && g > 0 && has_pw && i == R + 1)          // the pointwise accumulators overlap every row's columns
          for (int j = 0; j < R; ++j) mbar_wait_a(BA(B_READY + j), (g - 1) & 1u);
        const uint32_t g3 = g - items * (uint32_t)(nG - nL);   // 3x3 layers so far (the weight / halo-row barriers skip the pointwise stage)
        if (l == 0 || i < 2) mbar_wait_a(BA(B_TMA_FULL + i), (i < 2 ? g3 : items) & 1u);
        // input row written (own + side pixels), accumulator drained.  The side pixels arrive as st.async transactions
        // on this barrier (async proxy, like a multicast TMA load): a plain CTA-scope wait orders them
        if (g > 0 && i >= 2) mbar_wait_a(BA(B_READY + i - 2), (g - 1) & 1u);
        // weight part q is first needed by step R+1-q; the centre block by step R
        if (i >= R - 1) mbar_wait_a(BA(B_WFULL + R + 1 - i), g3 & 1u);
        if (i == R && ((ctr_mask >> l) & 1u)) mbar_wait_a(BA(B_WFULL + 3), ctrc & 1u);
      };
      if (it_valid) wait_step(0, R + 1, 0, 0, 0);
      // the current layer's constants live in registers and change only at a layer boundary (an indexed read of the
      // parameter bank per step costs the single issuing thread tens of cycles each)
      int np = 0, ks = 0, ctr_n = 0, acc_col = 0, part_bytes = 0, cur_l = -1;
      bool res_ident = false;
      const uint32_t d_ctr = tmem_base + (uint32_t)p.ctr_acc_col;
      const uint32_t A_RING = 0x10000u | (((ring_base + (CH_PX0 - 1) * 128) & 0x3FFFFu) >> 4);   // + SLOT >> 4 per row
      const uint32_t B_W = 0x10000u | ((w_base & 0x3FFFFu) >> 4);
      const uint32_t B_CTR = 0x10000u | ((ctr_base & 0x3FFFFu) >> 4), B_ID = 0x10000u | ((ident_base & 0x3FFFFu) >> 4);
      while (it_valid) {
        const int l = it_l, i = it_i;
        const uint32_t g = it_g;
        if (kPw && l == nL) {
          // ---- pointwise stage
          tc_fence_after_sync();
          if (i < R && !no_mma) {
            const ChPw& P = p.pw;
            const int pr = (R - 1 - i) & 1;
            const uint32_t d = tmem_base + (uint32_t)(P.acc_col + (i & 1) * P.n);
            const uint32_t idp = umma_idesc_f16((uint32_t)P.n);
            for (int c = 0; c < P.nchunks; ++c) {
              const uint32_t Ap = 0x10000u | (((ring_base + (uint32_t)(5 - 2 * pr - c) * CH_SLOT_BYTES + CH_PX0 * 128) & 0x3FFFFu) >> 4);
              const uint32_t Bp = 0x10000u | (((w_base + (uint32_t)(P.w_soff + c * P.n * 128)) & 0x3FFFFu) >> 4);
              for (int k = 0; k < P.ksteps; ++k) umma_f16_ss_hi(d, Ap + 2u * k, HI_A, Bp + 2u * k, HI_A, idp, (c | k) ? 1u : 0u);
            }
          }
          umma_commit_a(BA(B_SD + i));
          if (g < 8) CH_STAMP(1, g * 6 + i);
          if (--it_i < 0) {
            it_i = R + 1;
            ++it_g;
            it_l = 0;
            ++it_items;
            it_item = next_item(it_items);
            it_valid = it_item < n_items;
          }
          if (it_valid) wait_step(it_l, it_i, it_g, it_items, it_ctr);
          continue;
        }
        if (l != cur_l) {
          const ChLayer& Lr = p.L[l];
          np = Lr.np; ks = no_mma ? 0 : Lr.ksteps; ctr_n = kCtr ? Lr.ctr_n : 0; acc_col = Lr.acc_col; part_bytes = Lr.part_bytes;
          res_ident = kCtr && Lr.res_smem != 0;   // (the identity tap only occurs in RFDB stages: same instantiations as the centre block)
          cur_l = l;
        }
        const uint32_t pb16 = (uint32_t)part_bytes >> 4;
        tc_fence_after_sync();
        if (g < 5) CH_STAMP(3, g * 12 + i * 2);
        const int jlo = max(i - 2, 0), jhi = min(i, R - 1), nj = jhi - jlo + 1;
        const int part0 = i >= 2 ? 0 : (i == 1 ? 1 : 2);
        const uint32_t A0 = A_RING + (uint32_t)i * (CH_SLOT_BYTES >> 4);                     // + 8 per dx, + 2 per K step
        const uint32_t B0 = B_W + (uint32_t)part0 * pb16;                                    // + 64 per dx, + 2 per K step
        const uint32_t d0 = tmem_base + (uint32_t)(acc_col + jlo * np);
        const uint32_t id_all = umma_idesc_f16((uint32_t)(np * nj)), id_one = umma_idesc_f16((uint32_t)np);
        // the very first MMA of a step with i >= 2 initialises output row i-2 (and accumulates onto the others)
        if (ks > 0) {
          if (i >= 2) {
            umma_f16_ss_hi(d0, A0, HI_A, B0, HI_B3, id_one, 0u);
            if (nj > 1) umma_f16_ss_hi(d0 + (uint32_t)np, A0, HI_A, B0 + pb16, HI_B3, umma_idesc_f16((uint32_t)(np * (nj - 1))), 1u);
          } else {
            umma_f16_ss_hi(d0, A0, HI_A, B0, HI_B3, id_all, 1u);
          }
        }
        // The short MMAs (the centre block of the distillation 1x1, N <= 32, on output row i-1, and the identity tap of
        // the block residual) go right behind the first tap: the step ends with full-width MMAs, which is what the
        // tensor pipe works on while this thread commits and sets up the next step
        if (i >= 1 && i <= R && !no_short) {
          if (kCtr && ctr_n > 0) umma_f16_ss_run(d_ctr + (uint32_t)((i - 1) * 32), A0 + 8, HI_A, B_CTR, HI_A, umma_idesc_f16((uint32_t)ctr_n), 0u, ks);
          if (res_ident) umma_f16_ss_run(tmem_base + (uint32_t)(acc_col + (i - 1) * np), A0 + 8, HI_A, B_ID, HI_A, id_one, 1u, ks);
        }
        // dx = -1 (remaining K steps) and dx = 0 taps
        umma_f16_ss_run(d0, A0 + 2, HI_A, B0 + 2, HI_B3, id_all, 1u, ks - 1);
        umma_f16_ss_run(d0, A0 + 8, HI_A, B0 + 64, HI_B3, id_all, 1u, ks);
        // advance, and take the next step's waits while the MMAs above execute - unless the next step opens a new
        // stage whose weights may only be requested once THIS step has been committed (a wider layer after a narrower
        // one waits for sd[0]; the pointwise stage): waiting for them here would be a cycle
        bool deferred_wait = false;
        {
          if (--it_i < 0) {
            it_i = R + 1;
            if (ctr_n > 0) ++it_ctr;
            ++it_g;
            if (++it_l == nG) {
              it_l = 0;
              ++it_items;
              it_item = next_item(it_items);
              it_valid = it_item < n_items;
            }
          }
          if (it_valid) {
            // (equal widths: the next stage's first weight part was requested after step 2 - safe to wait for here)
            if (it_i == R + 1 && (it_l == nL || p.L[it_l].part_bytes > 2 * part_bytes)) deferred_wait = true;
            else {
              if (g == 2) CH_STAMP(3, 48 + i * 2);        // (timeline: how long the mid-step wait of layer 2 blocks)
              wait_step(it_l, it_i, it_g, it_items, it_ctr);
              if (g == 2) CH_STAMP(3, 49 + i * 2);
            }
          }
        }
        // dx = +1 taps
        umma_f16_ss_run(d0, A0 + 16, HI_A, B0 + 128, HI_B3, id_all, 1u, ks);
        if (g < 5) CH_STAMP(3, g * 12 + i * 2 + 1);
        umma_commit_a(BA(B_SD + i));
        if (g < 8) CH_STAMP(1, g * 6 + i);
        if (deferred_wait) wait_step(it_l, it_i, it_g, it_items, it_ctr);
      }
    }
    __syncwarp();
  } else if (warp < 2 + CH_EPI_WARPS) {
    // ================================ epilogue ====================================
    // 16 warps: TMEM lane quadrant q = warp % 4 (hardware rule), `sub` = which 16-column unit of the accumulator row
    // this warp converts (plus the same unit of the centre block's accumulator for sub < 2).  Everything a tile
    // needs is hoisted into registers per layer: the parameter bank is re-read after every asm volatile otherwise.
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int m = q * 32 + lane;
    const int pos = CH_PX0 + m;                 // ring position of this thread's pixel
    const bool edgeL = hasL && m == 0, edgeR = hasR && m == TC_TILE_PX - 1, edge = edgeL || edgeR;
    const int x = x0 + m;
    void* const ps_out = p.ps_out;
    const int ps_fp32 = p.ps_fp32, ps_u8 = kU8 ? p.ps_u8 : 0;
    const float ps_dr = p.ps_dr;
    uint32_t u8_cnt = 0;
    const int ring_off = p.ring_off, stage_off = p.stage_off, stage_bytes = p.stage_bytes, ctr_acc_col = p.ctr_acc_col;
    // byte offsets of this thread's two 16-byte pieces inside a ring slot, and of the neighbour's halo position
    const uint32_t roff0 = (uint32_t)(pos * 128 + (((2 * sub) ^ (pos & 7)) << 4)), roff1 = (uint32_t)(pos * 128 + (((2 * sub + 1) ^ (pos & 7)) << 4));
    const int rpos = edgeL ? (CH_PX0 + TC_TILE_PX) : (CH_PX0 - 1);
    const uint32_t xoff0 = (uint32_t)(rpos * 128 + (((2 * sub) ^ (rpos & 7)) << 4)), xoff1 = (uint32_t)(rpos * 128 + (((2 * sub + 1) ^ (rpos & 7)) << 4));
    const uint32_t nb_rank = edgeL ? rank - 1 : rank + 1;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool epi_skip = (dbg_flags & 2) != 0;
    griddep_wait();
    uint32_t g = 0, stage_cnt = 0, ring_cnt = 0, pw_rows = 0, items = 0;
    for (int item = (int)cid; item < n_items; ++items, item = next_item(items)) {
      const int img = item / nbands, band = item - img * nbands, y0 = band * R;
      for (int l = 0; l < nL; ++l, ++g) {
        const ChLayer& Lr = p.L[l];
        const int np = Lr.np, n0 = Lr.n0, n1 = Lr.n1;
        const bool ring_out = Lr.ring_out != 0;
        const bool to_pw = has_pw && l == nL - 1 && p.pw.from_smem != 0;   // group 0 goes straight into the pointwise stage's A slot
        const bool staged = (n1 > 0) || (!ring_out && Lr.mode0 == 0 && !to_pw);
        // uint8 output: the pixel-shuffle layer converts in the epilogue (tensor2uint) and stages one LR row = 4 output rows
        // x 128 x 4 pixels x 3 bytes in the (otherwise unused) staging buffers; the epilogue warps copy the tile out themselves
        const bool u8_layer = kU8 && ps_u8 != 0 && Lr.mode0 == 1 && !ring_out;
        const int pw_slot_c = p.pw.fs_chunk, pw_lane0 = p.pw.fs_lane0;
        const int c = sub * 16;
        // unit A = accumulator columns [c, c+16): kind 0 none, 1 group 0, 2 group 1 (IMDN: columns of the same conv)
        int kindA = 0, c0A = 0;
        if (c < n0) { kindA = 1; c0A = c; }
        else if (!Lr.g1_ctr && n1 > 0 && c >= Lr.col1 && c < Lr.col1 + n1) { kindA = 2; c0A = c - Lr.col1; }
        const bool zfillA = ring_out && kindA != 1;            // ring lanes this layer does not produce are written as zeros
        const bool ringA = ring_out && kindA == 1;
        const bool kindB = kCtr && Lr.g1_ctr && c < n1;        // unit B = the centre block's columns [c, c+16): always group 1
        const uint32_t tcolA = (uint32_t)(Lr.acc_col + c), tcolB = (uint32_t)(ctr_acc_col + c);
        const float* const biasA = &bias_s[l][(kindA == 2 ? 64 : 0) + c0A];
        const float* const biasB = &bias_s[l][64 + c];
        const float slopeA = kindA == 2 ? Lr.slope1 : Lr.slope0, slopeB = Lr.slope1;
        const int resA = (kTail && kindA == 1 && Lr.res != nullptr) ? 2 : 0;   // (the block residual arrives through the accumulator)
        const int res_after = Lr.res_after;
        const __half* const gres = Lr.res;
        const int gres_stride = Lr.res_stride, gres_coff = Lr.res_coff + c0A;
        // staged destinations (unit A when it is not a ring unit, unit B always)
        const int ncolsA = kindA == 1 ? n0 : n1, swzmA = kindA == 1 ? Lr.swz0 : Lr.swz1, modeA = kindA == 1 ? Lr.mode0 : 0;
        const int swzA = swzmA == 1 ? (m & 7) : (swzmA == 2 ? ((m >> 1) & 3) : (swzmA == 3 ? ((m >> 2) & 1) : 0));
        const int swzB = Lr.swz1 == 1 ? (m & 7) : (Lr.swz1 == 2 ? ((m >> 1) & 3) : (Lr.swz1 == 3 ? ((m >> 2) & 1) : 0));
        const uint32_t strowA = (uint32_t)(m * ncolsA * 2), strowB = (uint32_t)(m * n1 * 2);
        for (int j = R - 1; j >= 0; --j) {
          const int y = y0 - l + j;
          const bool valid = y >= 0 && y < H && x < W;
          uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0;
          if (resA == 2 && valid) {     // residual from global memory (LR_conv + fea): fetched before the accumulator is ready
            const uint4* rp = reinterpret_cast<const uint4*>(gres + (((long long)img * H + y) * W + x) * gres_stride + gres_coff);
            u0 = rp[0]; u1 = rp[1];
          }
          mbar_wait(&sd[j], g & 1u);
          tc_fence_after_sync();
          uint32_t va[16], vb[16];
          if (kindA) tmem_ld16_nc(trow + tcolA + (uint32_t)(j * np), va);
          if (kindB) tmem_ld16_nc(trow + tcolB + (uint32_t)(j * 32), vb);
          tmem_ld_wait();
          if (j == R - 1 && l > 0) mbar_wait(&ring_read, (ring_cnt - 1) & 1u);   // TMA stores out of the ring slots of the previous layer have read them
          const uint32_t buf = stage_cnt & 1u;
          if (staged) mbar_wait(&sfree[buf], ((stage_cnt >> 1) & 1u) ^ 1u);
          if (edge && ring_out) mbar_wait(&nfree[j], ring_cnt & 1u);            // the neighbour has read the slot my edge pixel goes into (a permission, no data)
          uint8_t* const slot_out = smem + ring_off + (j + 2) * CH_SLOT_BYTES;
          uint8_t* const stage = smem + stage_off + buf * stage_bytes;
          if ((kindA || zfillA) && !epi_skip) {
            float f[16];
            if (kindA) {
              tc_epi_math16(va, biasA, false, slopeA, resA != 0, u0, u1, res_after, f);
            }
            if (ringA || zfillA) {
              uint4 o0, o1;
              __half2* h0 = reinterpret_cast<__half2*>(&o0);
              __half2* h1 = reinterpret_cast<__half2*>(&o1);
              const bool nz = valid && ringA;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h0[e] = nz ? __floats2half2_rn(f[2 * e], f[2 * e + 1]) : __floats2half2_rn(0.f, 0.f);
                h1[e] = nz ? __floats2half2_rn(f[8 + 2 * e], f[8 + 2 * e + 1]) : __floats2half2_rn(0.f, 0.f);
              }
              *reinterpret_cast<uint4*>(slot_out + roff0) = o0;
              *reinterpret_cast<uint4*>(slot_out + roff1) = o1;
              if (edge) {   // the same 32 bytes into the neighbour's halo position; the store itself signals its ready[j]
                const uint32_t rslot = mapa_u32(smem_u32(slot_out), nb_rank), rbar = mapa_u32(smem_u32(&ready[j]), nb_rank);
                st_async_v4(rslot + xoff0, o0, rbar);
                st_async_v4(rslot + xoff1, o1, rbar);
              }
            }
            if (kindA == 1 && to_pw) {
              const int pr = (R - 1 - j) & 1;
              mbar_wait(&pw_afull[pr], j < R - 2 ? 1u : 0u);     // the rest of that pixel row has arrived from global memory
              uint8_t* const slot = smem + ring_off + (5 - 2 * pr - pw_slot_c) * CH_SLOT_BYTES;
              uint4 o0, o1;
              __half2* h0 = reinterpret_cast<__half2*>(&o0);
              __half2* h1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h0[e] = valid ? __floats2half2_rn(f[2 * e], f[2 * e + 1]) : __floats2half2_rn(0.f, 0.f);
                h1[e] = valid ? __floats2half2_rn(f[8 + 2 * e], f[8 + 2 * e + 1]) : __floats2half2_rn(0.f, 0.f);
              }
              const int ch = (pw_lane0 + c0A) >> 3;
              *reinterpret_cast<uint4*>(slot + pos * 128 + ((ch ^ (pos & 7)) << 4)) = o0;
              *reinterpret_cast<uint4*>(slot + pos * 128 + (((ch + 1) ^ (pos & 7)) << 4)) = o1;
            } else if (kindA && !ringA && u8_layer) {
              if (valid) {   // column 16*ch + 4*i + jj of pixel m -> byte (i, 4m + jj, ch) of the tile
                uint8_t* const tile = smem + stage_off + (u8_cnt & 1u) * stage_bytes + (m * 4) * 3 + (c0A >> 4);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                  for (int jj = 0; jj < 4; ++jj) {
                    const float hv = __half2float(__float2half_rn(f[4 * i + jj]));       // the fp16 engine's output value
                    const float q = __fdiv_rn(__fmul_rn(fminf(fmaxf(hv, 0.f), ps_dr), 255.0f), ps_dr);
                    tile[i * 1536 + jj * 3] = (uint8_t)rintf(q);                          // np.round: half to even
                  }
              }
            } else if (kindA && !ringA)
              tc_epi_store16(f, kTail ? modeA : 0, stage + strowA, c0A, swzA, valid, ps_out, ps_fp32, img, y, x, H, W);
          }
          if (kindB && !epi_skip) {
            float f[16];
            const uint4 z = make_uint4(0, 0, 0, 0);
            tc_epi_math16(vb, biasB, false, slopeB, false, z, z, 0, f);
            tc_epi_store16(f, 0, stage + strowB, c, swzB, valid, ps_out, ps_fp32, img, y, x, H, W);
          }
          if (u8_layer) {
            // every epilogue warp has written its bytes: copy the tile out, 4 output rows x (3 * valid pixels) words.  Two
            // tiles alternate, so the barrier of the next row also says that everybody is done reading this one.
            named_bar_sync(1, 32 * CH_EPI_WARPS);
            if (y >= 0 && y < H && !epi_skip) {
              const uint32_t* const t32 = reinterpret_cast<const uint32_t*>(smem + stage_off + (u8_cnt & 1u) * stage_bytes);
              const int nwords = 3 * min(TC_TILE_PX, W - x0);
              uint8_t* const ob = reinterpret_cast<uint8_t*>(ps_out) + ((((long long)img * 4 * H + 4 * y) * (4 * W)) + 4 * x0) * 3;
#pragma unroll
              for (int t = 0; t < 3; ++t) {
                const int k = (int)threadIdx.x - 64 + 32 * CH_EPI_WARPS * t;
                const int i = k / 384, kk = k - 384 * i;
                if (kk < nwords) reinterpret_cast<uint32_t*>(ob + (long long)i * (12 * W))[kk] = t32[k];
              }
            }
            ++u8_cnt;
          }
          tc_fence_before_sync();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // ready[j]: CH_EPI_WARPS arrivals + (ring layers) 128 bytes of edge pixels from every side neighbour
            if (warp == 2 && ring_out && n_side > 0 && !epi_skip) mbar_arrive_expect_tx(&ready[j], (uint32_t)(128 * n_side));
            else mbar_arrive(&ready[j]);
            mbar_arrive(&wrote[j]);
          }
          if (staged) ++stage_cnt;
          if (threadIdx.x == 64 && g < 8) CH_STAMP(2, g * 4 + j);
        }
        if (ring_out) ++ring_cnt;
      }
      if (has_pw) {
        // ---- pointwise stage: output row j = accumulator slot (j & 1); 16-column units u = sub, sub + 4, sub + 8.
        // The unit descriptors are resolved into registers here (compile-time indexed), not inside the row loop.
        const ChPw& P = p.pw;
        const int n = P.n, nunits = n >> 4, acc_col = P.acc_col, ngr = P.ngroups;
        const bool dbl = P.nbuf == 2, one64 = P.n64 <= 1;
        const int w_off_s = p.w_off;
        const __half* const gres = P.res;
        const int gres_stride = P.res_stride, gres_coff = P.res_coff, res_after = P.res_after;
        bool has16 = false;
        for (int k = 0; k < ngr; ++k) has16 = has16 || P.g_stage[k] == 2;
        bool uval[3], ures[3];
        int ucol[3], uc0[3], ust[3];
        float usl[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
          const int u = sub + 4 * t, c = u * 16;
          int k = 0;
          if (ngr > 1 && c >= P.g_col0[1]) k = 1;
          if (ngr > 2 && c >= P.g_col0[2]) k = 2;
          const int col0 = k == 0 ? P.g_col0[0] : (k == 1 ? P.g_col0[1] : P.g_col0[2]);
          const int nc = k == 0 ? P.g_ncols[0] : (k == 1 ? P.g_ncols[1] : P.g_ncols[2]);
          ucol[t] = c;
          uc0[t] = c - col0;
          uval[t] = u < nunits && uc0[t] < nc;
          ust[t] = k == 0 ? P.g_stage[0] : (k == 1 ? P.g_stage[1] : P.g_stage[2]);
          usl[t] = k == 0 ? P.g_slope[0] : (k == 1 ? P.g_slope[1] : P.g_slope[2]);
          ures[t] = k == 0 && gres != nullptr;
        }
        for (int j = R - 1; j >= 0; --j) {
          const int y = y0 - (nL - 1) + j;
          const bool valid = y >= 0 && y < H && x < W;
          mbar_wait(&sd[j], g & 1u);
          tc_fence_after_sync();
          const uint32_t buf = stage_cnt & 1u;
          const uint32_t pb = dbl ? (pw_rows & 1u) : 0u, pn = dbl ? (pw_rows >> 1) : pw_rows;   // staging buffer of this row, its use count
          if (pn > 0) mbar_wait(&pw_sfree[pb], (pn - 1) & 1u);          // the TMA stores of its previous user have read it
          if (has16) mbar_wait(&sfree[buf], ((stage_cnt >> 1) & 1u) ^ 1u);
          uint8_t* const stage = smem + stage_off + buf * stage_bytes;
          const uint32_t tacc = trow + (uint32_t)(acc_col + (j & 1) * n);
#pragma unroll
          for (int t = 0; t < 3; ++t) {
            if (!uval[t]) continue;
            uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0;
            if (ures[t] && valid) {
              const uint4* rp = reinterpret_cast<const uint4*>(gres + (((long long)img * H + y) * W + x) * gres_stride + gres_coff + uc0[t]);
              u0 = rp[0]; u1 = rp[1];
            }
            uint32_t va[16];
            tmem_ld16_nc(tacc + (uint32_t)ucol[t], va);
            tmem_ld_wait();
            if (epi_skip) continue;
            float f[16];
            tc_epi_math16(va, &pw_bias_s[ucol[t]], false, usl[t], ures[t], u0, u1, res_after, f);
            if (ust[t] == 2) {
              tc_epi_store16(f, 0, stage + m * 32, uc0[t], (m >> 2) & 1, valid, ps_out, ps_fp32, img, y, x, H, W);
            } else {
              // staging row m of this group's buffer for this row (ring slot: behind the 8 pad positions)
              uint8_t* const sb = one64 ? smem + ring_off + (int)(1u - pb) * CH_SLOT_BYTES + CH_PX0 * 128
                                        : (pb == 0 ? smem + ring_off + (1 - ust[t]) * CH_SLOT_BYTES + CH_PX0 * 128
                                                   : smem + w_off_s + ust[t] * 16384);
              uint4 o0, o1;
              __half2* h0 = reinterpret_cast<__half2*>(&o0);
              __half2* h1 = reinterpret_cast<__half2*>(&o1);
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                h0[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
                h1[e] = __floats2half2_rn(f[8 + 2 * e], f[8 + 2 * e + 1]);
              }
              const int ch = uc0[t] >> 3;
              *reinterpret_cast<uint4*>(sb + m * 128 + ((ch ^ (m & 7)) << 4)) = o0;
              *reinterpret_cast<uint4*>(sb + m * 128 + (((ch + 1) ^ (m & 7)) << 4)) = o1;
            }
          }
          tc_fence_before_sync();
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) { mbar_arrive(&ready[j]); mbar_arrive(&wrote[j]); }
          if (has16) ++stage_cnt;
          ++pw_rows;
          if (threadIdx.x == 64 && g < 8) CH_STAMP(2, g * 4 + j);
        }
        ++g;
      }
    }
  } else if (warp == 3 + CH_EPI_WARPS) {
    // ================================ relay warp ==================================
    // tells the side neighbours that this CTA's MMAs have read ring slot j+2 (step j, and with it step j+2, has
    // completed): they may now drop the edge pixels of their output row j into it
    if (n_side > 0 && elect_one()) {
      uint32_t g = 0, items = 0;
      for (int item = (int)cid; item < n_items; ++items, item = next_item(items))
        for (int l = 0; l < nG; ++l, ++g) {
          const bool ring_out = l < nL && p.L[l].ring_out != 0;
          for (int j = R - 1; j >= 0; --j) {
            mbar_wait(&sd[j], g & 1u);     // every phase is waited for (a parity wait is only valid one phase deep)
            if (!ring_out) continue;
            if (hasL) mbar_arrive_remote(mapa_u32(smem_u32(&nfree[j]), rank - 1));
            if (hasR) mbar_arrive_remote(mapa_u32(smem_u32(&nfree[j]), rank + 1));
          }
        }
    }
    __syncwarp();
  } else if (warp == 2 + CH_EPI_WARPS) {
    // ================================ store warp ==================================
    if (elect_one()) {
      uint32_t g = 0, stage_cnt = 0, pw_rows = 0, items = 0;
      for (int item = (int)cid; item < n_items; ++items, item = next_item(items)) {
        const int img = item / nbands, band = item - img * nbands, y0 = band * R;
        for (int l = 0; l < nL; ++l, ++g) {
          const ChLayer& Lr = p.L[l];
          const bool to_pw = has_pw && l == nL - 1 && p.pw.from_smem != 0;
          const bool ring_out = Lr.ring_out != 0, g0_staged = !ring_out && Lr.mode0 == 0 && !to_pw, g1_staged = Lr.n1 > 0;
          const bool staged = g1_staged || g0_staged;
          const CUtensorMap* const map_out = &maps.m[Lr.map_out];
          const CUtensorMap* const map_g1 = &maps.m[Lr.map_g1];
          const int store_all = p.store_all, ring_off = p.ring_off, stage_off = p.stage_off, stage_bytes = p.stage_bytes;
          for (int j = R - 1; j >= 0; --j) {
            const int y = y0 - l + j;
            const bool row_valid = y >= 0 && y < H;
            mbar_wait(&wrote[j], g & 1u);
            const uint32_t buf = stage_cnt & 1u;
            if (row_valid && !(dbg_flags & 4)) {
              if (ring_out && (store_all || j >= R - 2))
                tma_store_4d(map_out, smem + ring_off + (j + 2) * CH_SLOT_BYTES + CH_PX0 * 128, 0, x0, y, img);
              if (g0_staged) tma_store_4d(map_out, smem + stage_off + buf * stage_bytes, 0, x0, y, img);
              if (g1_staged) tma_store_4d(map_g1, smem + stage_off + buf * stage_bytes, 0, x0, y, img);
            }
            tma_store_commit();
            if (staged) {
              tma_store_wait_read<0>();
              mbar_arrive(&sfree[buf]);
              ++stage_cnt;
            }
            if (has_pw && l == nL - 1 && !to_pw) {
              // the pointwise stage reads rows back from global memory: everything this CTA has stored for output row j
              // and before (all layers) must have landed
              tma_store_wait_all<0>();
              fence_proxy_async_all();
              mbar_arrive(&pw_drow[j]);
            }
            if (ring_out && j == R - 2) {
              // rows R-1 and R-2 (the halo of the band below) are on their way: publish them
              tma_store_wait_all<0>();
              fence_proxy_async_all();
              __threadfence();
              st_release_gpu(p.flags + ((size_t)item * strips + strip) * nL + l, 1);
              if (!store_all) mbar_arrive(&ring_read);   // ... and the ring slots they came from may be overwritten
            }
          }
          if (ring_out && store_all) {
            tma_store_wait_read<0>();
            mbar_arrive(&ring_read);
          }
          if (has_pw && p.pw.from_smem && l == nL - 2) {
            // the pointwise stage reads these layers' staged outputs back from global memory
            tma_store_wait_all<0>();
            fence_proxy_async_all();
            mbar_arrive(&pw_dpre);
          }
        }
        if (has_pw) {
          const ChPw& P = p.pw;
          const int ngr = P.ngroups, ring_off = p.ring_off, stage_off = p.stage_off, stage_bytes = p.stage_bytes;
          bool has16 = false;
          for (int k = 0; k < ngr; ++k) has16 = has16 || P.g_stage[k] == 2;
          const bool dbl = P.nbuf == 2, one64 = P.n64 <= 1;
          for (int j = R - 1; j >= 0; --j) {
            const int y = y0 - (nL - 1) + j;
            mbar_wait(&wrote[j], g & 1u);
            const uint32_t buf = stage_cnt & 1u;
            const uint32_t pb = dbl ? (pw_rows & 1u) : 0u;
            if (y >= 0 && y < H && !(dbg_flags & 4))
              for (int k = 0; k < ngr; ++k) {
                const int q = P.g_stage[k];
                const uint8_t* src = q == 2 ? smem + stage_off + buf * stage_bytes
                                   : one64 ? smem + ring_off + (int)(1u - pb) * CH_SLOT_BYTES + CH_PX0 * 128
                                   : (pb == 0 ? smem + ring_off + (1 - q) * CH_SLOT_BYTES + CH_PX0 * 128 : smem + p.w_off + q * 16384);
                tma_store_4d(&maps.m[P.g_map[k]], src, 0, x0, y, img);
              }
            tma_store_commit();
            tma_store_wait_read<0>();
            mbar_arrive(&pw_sfree[pb]);
            if (has16) { mbar_arrive(&sfree[buf]); ++stage_cnt; }
            ++pw_rows;
          }
          mbar_arrive(&pw_done);
          ++g;
        }
      }
      tma_store_wait_all<0>();
    }
    __syncwarp();
  }
  tc_fence_before_sync();
  __syncthreads();
  if (kDbg && dbg_blocks != nullptr && threadIdx.x == 0) dbg_blocks[blockIdx.x * 4 + 2] = global_timer_ns();
  // Every remote access to this CTA's shared memory has been consumed by now (the side pixels through ready[], the
  // "slot read" arrivals through nfree[]); the cluster barrier is only insurance and carries no data, so the relaxed
  // form is enough.  (The release / acquire form measured 4-8 us here: it drains the CTA's outstanding global writes.)
  if (!(dbg_flags & 8)) asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
  if (warp == 1) tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  if (threadIdx.x == 0) CH_STAMP(0, 1);
  if (kDbg && dbg_blocks != nullptr && threadIdx.x == 0) dbg_blocks[blockIdx.x * 4 + 3] = global_timer_ns();
}

}  // namespace esr
