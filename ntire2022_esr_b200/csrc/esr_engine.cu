// esr_b200 engine: C ABI (include/esr_b200.h), weight store, launch plans.
//
// Boundary replaced: `model(img_lq)` in the reference's forward() (test_demo.py:364-367) for the
// modules select_model() builds (test_demo.py:17-30,52-58,150-157).  There is no CPU fallback: a
// handle created without a usable sm_100 device can be loaded / finalized (host-side packing, used by
// the CPU-only tests) but every compute entry point returns ESR_E_NOGPU.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>

#include "chain_host.cuh"
#include "conv_chain.cuh"
#include "conv_tc.cuh"
#include "engine.h"
#include "graph_builder.cuh"
#include "kernels_generic.cuh"

namespace esr {

static const int kNumArch = 6;

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct Launch {
  std::string name;
  std::function<cudaError_t(cudaStream_t)> fn;
  double flops = 0;  // algorithmic FLOPs (2 x unpadded MACs) of this launch
};

struct Plan {
  int B = 0, H = 0, W = 0, dtype = 0, gid = 0;
  const void* in = nullptr;
  void* out = nullptr;
  void* ws = nullptr;
  // uint8 I/O (esr_forward_u8): `in` is HWC uint8, `out` the engine's own NCHW buffer, `u8_out` the caller's HWC uint8
  bool u8 = false;
  bool u8_folded = false;   // the pixel-shuffle layer wrote the uint8 image itself
  float data_range = 1.f;
  void* u8_out = nullptr;
  std::vector<Launch> launches;
  cudaGraphExec_t gexec = nullptr;
  bool graph_failed = false;
  int hits = 0;   // the CUDA graph is captured on the second use of a plan: a one-off shape never pays the ~1 ms of capture + instantiate
};

struct DevGraph {            // a Graph plus its device-side parameters
  Graph g;
  std::vector<Table> tables;
  float* d_params = nullptr;   // all tables (w, b) back to back
  uint8_t* d_blobs = nullptr;  // all tcgen05 weight blobs
  uint8_t* d_ident = nullptr;  // [64 x 64] fp16 identity block (K-major SWIZZLE_128B) for the fused chains
  bool built = false;
};

struct Engine {
  int arch = 0, nf = 0, nblocks = 0, device = -1;
  bool has_gpu = false;
  bool finalized = false;
  int num_sms = 148;
  std::map<std::string, HostTensor> weights;
  DevGraph graphs[2];  // 0: CUDA-core graph (fp32 mode, and fp16 when tc is off), 1: tcgen05 graph
  std::string err;
  // options
  int opt_tc = 1, opt_shift_mode = 0, opt_use_graph = 1, opt_rows_per_item = 0, opt_timeline = 0, opt_dbg_flags = 0, opt_acc_slots = 4, opt_pdl = 0, opt_esa_front_old = 0;
  int opt_chain = 1, opt_chain_store_all = 0, opt_chain_mask = -1;   // mask: bit k enables the k-th chain of the graph (debug)
  int opt_chain_pw = 0;   // 1: the pointwise layer behind a chain (RFDB c5, IMDB conv1x1) runs as the chain's last stage.  Off by
                          // default: measured 385.6 vs 375.2 us per RFDN forward at batch 1 - the 144-column epilogue of c5 is
                          // issue bound either way and inside the chain it queues behind the last 3x3 layer's epilogues
  bool attr_tc = false, attr_chain = false, attr_esa = false;   // per-handle (= per-device) function attribute opt-ins
  int chain_capacity[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};          // co-resident clusters of the fused chain kernel by cluster size (0 = not queried)
  long long* d_timeline = nullptr;  // 128 stamps per tcgen05 launch (debug option tc_timeline)
  std::list<Plan> plans;
  // host-buffer path: kHostSlots requests in flight.  Every slot owns its device buffers, its workspace and its
  // compute stream, so the H2D copy, the forward and the D2H copy of consecutive requests overlap AND the forwards
  // of independent requests run concurrently: at batch 1 a forward is a chain of ~36 short kernels whose launch
  // gaps, epilogue tails and small ESA kernels leave most SMs idle (measured: 2274 img/s on one stream, 2688 on
  // two, 2985 on three, 3036 on four; tools/gpu_two_streams.py).
  static const int kHostSlots = 4;
  struct HostSlot {
    void* d_in = nullptr; void* d_out = nullptr;
    size_t in_sz = 0, out_sz = 0;
    cudaEvent_t ev_in = nullptr, ev_fwd = nullptr, ev_done = nullptr;
    bool busy = false;
    long long ticket = -1;
    void* ws = nullptr; size_t ws_sz = 0;
    cudaStream_t s_cmp = nullptr;
  } hslot[kHostSlots];
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  long long next_ticket = 0;
  struct WsKey { void* p; size_t need; int B, H, W, dtype, gid; };
  std::vector<WsKey> ws_zeroed;   // workspaces whose padded lanes have been cleared, keyed on the full buffer layout
  PFN_encodeTiled encode = nullptr;
};

// Every ABI entry point that touches the device runs on the handle's device and restores the caller's current device
// afterwards (under PyTorch, torch.cuda.current_device() IS cudaGetDevice(): an engine bound to cuda:1 must not leave the
// caller on cuda:1).
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (dev < 0) return;
    if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; cudaGetLastError(); }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) { ok = false; cudaGetLastError(); }
    if (prev == dev) prev = -1;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

static int fail(Engine* e, int code, const std::string& msg) {
  if (e) e->err = msg;
  return code;
}
#define CUDA_TRY(e, expr)                                                                           \
  do {                                                                                              \
    cudaError_t _err = (expr);                                                                      \
    if (_err != cudaSuccess)                                                                        \
      return fail(e, ESR_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_err));              \
  } while (0)

static void esa_dims(int H, int W, int& H2, int& W2, int& H3, int& W3) {
  H2 = (H - 3) / 2 + 1; W2 = (W - 3) / 2 + 1;       // conv2: 3x3 stride 2 pad 0
  H3 = (H2 - 7) / 3 + 1; W3 = (W2 - 7) / 3 + 1;     // max_pool2d(7, 3)
}

static std::string build_dev_graph(Engine& e, int gid) {
  DevGraph& dg = e.graphs[gid];
  GraphBuilder gb;
  for (auto& kv : e.weights) kv.second.used = false;
  gb.wts.store = &e.weights;
  const bool tc = gid == 1;
  switch (e.arch) {
    case ESR_ARCH_RFDN: gb.build_rfdn(e.nf, e.nblocks, tc); break;
    case ESR_ARCH_RFDN_PRUNED: gb.build_rfdn(e.nf, e.nblocks, tc, /*residual=*/false, /*esa_f=*/12); break;
    case ESR_ARCH_RLFN: gb.build_rlfn(e.nf, e.nblocks, tc); break;
    case ESR_ARCH_IMDN: gb.build_imdn(e.nf, e.nblocks, tc); break;
    case ESR_ARCH_BSRN: gb.build_bsrn(e.nf, e.nblocks, tc); break;
    case ESR_ARCH_FMEN: gb.build_fmen(e.nf, e.nblocks, tc); break;
    default: return "unknown architecture";
  }
  if (!gb.wts.err.empty()) return gb.wts.err;
  for (auto& kv : e.weights)
    if (!kv.second.used) return "unexpected key in state_dict: " + kv.first;
  dg.g = std::move(gb.g);
  dg.tables = std::move(gb.tables);
  if (tc) {
    for (int i = 0; i < 3; ++i) {
      dg.g.bufs.push_back(BufDecl{BK_FULL, 64, false});
      dg.g.chain_buf[i] = (int)dg.g.bufs.size() - 1;
    }
    find_chains(dg.g);
  }
  // parameter arena layout
  size_t nf32 = 0;
  for (auto& t : dg.tables) {
    t.off_w = nf32; nf32 += (t.w.size() + 63) / 64 * 64;
    t.off_b = nf32; nf32 += (t.b.size() + 63) / 64 * 64;
  }
  size_t nblob = 0;
  for (auto& c : dg.g.tc) {
    c.off_blob = nblob;
    nblob += (c.blob.size() + 1023) / 1024 * 1024;
    for (auto& gd : c.groups) {
      gd.off_bias = dg.tables[gd.off_bias].off_b;
      if (gd.off_bias9 >= 0) gd.off_bias9 = (long long)dg.tables[gd.off_bias9].off_b;
    }
  }
  for (auto& ch : dg.g.chains) {
    ch.off_blob = nblob;
    nblob += (ch.blob.size() + 1023) / 1024 * 1024;
  }
  if (e.has_gpu) {
    std::vector<float> host(nf32, 0.f);
    for (auto& t : dg.tables) {
      std::copy(t.w.begin(), t.w.end(), host.begin() + t.off_w);
      std::copy(t.b.begin(), t.b.end(), host.begin() + t.off_b);
    }
    if (cudaMalloc(&dg.d_params, std::max<size_t>(nf32, 64) * sizeof(float)) != cudaSuccess) return "cudaMalloc(params) failed";
    if (cudaMemcpy(dg.d_params, host.data(), nf32 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
      return "cudaMemcpy(params) failed";
    if (nblob) {
      std::vector<uint8_t> hb(nblob, 0);
      for (auto& c : dg.g.tc) std::copy(c.blob.begin(), c.blob.end(), hb.begin() + c.off_blob);
      for (auto& ch : dg.g.chains) std::copy(ch.blob.begin(), ch.blob.end(), hb.begin() + ch.off_blob);
      if (cudaMalloc(&dg.d_blobs, nblob) != cudaSuccess) return "cudaMalloc(blobs) failed";
      if (cudaMemcpy(dg.d_blobs, hb.data(), nblob, cudaMemcpyHostToDevice) != cudaSuccess) return "cudaMemcpy(blobs) failed";
    }
    if (tc) {
      std::vector<uint8_t> id(CH_IDENT_BYTES, 0);
      const __half one = __float2half_rn(1.f);
      for (uint32_t n = 0; n < 64; ++n) memcpy(id.data() + sw128_offset(n, n), &one, 2);
      if (cudaMalloc(&dg.d_ident, CH_IDENT_BYTES) != cudaSuccess) return "cudaMalloc(ident) failed";
      if (cudaMemcpy(dg.d_ident, id.data(), CH_IDENT_BYTES, cudaMemcpyHostToDevice) != cudaSuccess) return "cudaMemcpy(ident) failed";
    }
  }
  dg.built = true;
  return "";
}

// ---- workspace layout -----------------------------------------------------------------------------
struct WsLayout {
  std::vector<size_t> off;
  std::vector<int> H, W;
  size_t flags_off = 0, flags_bytes = 0;   // halo flags of the fused chains (conv_chain.cuh), zeroed at the start of a forward
  size_t total = 0;
};
static size_t chain_flag_ints(int B, int H, int W, int n_layers) {   // halo flags + the band counter
  const int nbands = (H + n_layers - 1 + CH_R - 1) / CH_R, strips = (W + TC_TILE_PX - 1) / TC_TILE_PX;
  return (size_t)B * nbands * strips * n_layers + 1;
}
static WsLayout ws_layout(const Graph& g, int B, int H, int W, int dtype) {
  WsLayout L;
  int H2, W2, H3, W3;
  esa_dims(H, W, H2, W2, H3, W3);
  const size_t elt = dtype == ESR_DTYPE_F16 ? 2 : 4;
  size_t off = 0;
  for (auto& b : g.bufs) {
    const int h = b.kind == BK_FULL ? H : (b.kind == BK_S2 ? H2 : H3);
    const int w = b.kind == BK_FULL ? W : (b.kind == BK_S2 ? W2 : W3);
    L.off.push_back(off);
    L.H.push_back(h);
    L.W.push_back(w);
    const size_t bytes = (size_t)B * h * w * b.C * (b.f32 ? 4 : elt);
    off += (bytes + 1023) / 1024 * 1024;
  }
  L.flags_off = off;
  for (auto& ch : g.chains) L.flags_bytes += (chain_flag_ints(B, H, W, ch.n_ops) * sizeof(int) + 1023) / 1024 * 1024;
  off += L.flags_bytes;
  L.total = off + 2048;
  return L;
}

// ---- generic launches -----------------------------------------------------------------------------
// Every kernel goes through this helper.  With g_pdl set the launch carries the programmatic stream
// serialization attribute: the kernel may start while its predecessor drains, and synchronises on it
// with griddepcontrol.wait (all kernels of the engine call it before touching activations).
static thread_local bool g_pdl = false;
template <typename... KArgs, typename... Args>
static cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}
template <typename K, typename P>
static cudaError_t launch1(K kern, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const P& p) {
  return launch_k(kern, grid, block, smem, s, p);
}

static int make_tensor_map(Engine* e, CUtensorMap* m, void* base, int C_stride_elems, int c_extent, int W, int H, int B,
                           int box_c, int box_w, int swizzle /* 0 none, 1 128B, 2 64B, 3 32B */) {
  cuuint64_t dims[4] = {(cuuint64_t)c_extent, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C_stride_elems * 2, (cuuint64_t)W * C_stride_elems * 2,
                           (cuuint64_t)H * W * C_stride_elems * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = e->encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE,
                         swizzle == 1 ? CU_TENSOR_MAP_SWIZZLE_128B
                                      : (swizzle == 2 ? CU_TENSOR_MAP_SWIZZLE_64B
                                                      : (swizzle == 3 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE)),
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(e, ESR_E_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return ESR_OK;
}

static double op_flops(const OpDecl& op, int B, int H, int W) {
  int H2, W2, H3, W3;
  esa_dims(H, W, H2, W2, H3, W3);
  const double px = op.macs_res == BK_FULL ? (double)H * W : (op.macs_res == BK_S2 ? (double)H2 * W2 : (double)H3 * W3);
  return 2.0 * op.macs_pp * px * B;
}

static const size_t kMaxSmem = 232448 - 4096;  // 227 KB minus the kernel's static shared memory (barriers, bias)

static int plan_tc(Engine* e, const DevGraph& dg, const TcConv& c, const std::string& name, const WsLayout& L, Plan& pl) {
  struct Packed {
    CUtensorMap tmA, tmO[TC_MAX_GROUPS];
    TcParams p;
  };
  auto pk = std::make_shared<Packed>();
  memset(pk.get(), 0, sizeof(Packed));
  TcParams& p = pk->p;
  const int B = pl.B, H = pl.H, W = pl.W;
  uint8_t* ws = reinterpret_cast<uint8_t*>(pl.ws);
  p.B = B; p.H = H; p.W = W;
  p.halo = c.halo;
  p.nchunks = c.nchunks;
  p.strip_px = TC_TILE_PX + 2 * c.halo;
  p.chunk_bytes = (p.strip_px * 128 + 1023) / 1024 * 1024;
  p.strip_bytes = p.nchunks * p.chunk_bytes;
  p.n_entries = (int)c.entries.size();
  p.ngroups = (int)c.groups.size();
  if (p.n_entries > TC_MAX_ENTRIES) return fail(e, ESR_E_INVALID, name + ": too many MMA entries");
  p.acc_cols = (c.acc_cols + 15) / 16 * 16;
  p.acc_slots = (e->opt_acc_slots == 4 && 4 * p.acc_cols <= 512) ? 4 : 2;
  int tm = 32;
  while (tm < p.acc_slots * p.acc_cols) tm *= 2;
  if (tm > 512) return fail(e, ESR_E_INVALID, name + ": accumulator does not fit TMEM");
  p.tmem_cols = tm;
  p.ps_fp32 = 0;
  p.ps_out = pl.out;
  for (int i = 0; i < 4; ++i) p.chunk_c0[i] = c.chunk_c0[i];
  p.wblob = dg.d_blobs + c.off_blob;
  p.dbg = nullptr;
  p.dbg_flags = e->opt_dbg_flags;
  p.pdl = e->opt_pdl;
  if (e->opt_timeline) {
    if (!e->d_timeline) {
      CUDA_TRY(e, cudaMalloc(&e->d_timeline, 256 * 128 * sizeof(long long)));
      CUDA_TRY(e, cudaMemset(e->d_timeline, 0, 256 * 128 * sizeof(long long)));
    }
    int idx = 0;
    for (auto& l : pl.launches) idx += l.name.rfind("conv_tc", 0) == 0 ? 1 : 0;
    if (idx < 224) p.dbg = e->d_timeline + (size_t)idx * 128;
  }
  // shared memory carve-up
  size_t off = 0;
  p.w_off = 0;
  p.w_bytes = (int)c.blob.size();
  off += (c.blob.size() + 1023) / 1024 * 1024;
  if (p.ngroups > TC_MAX_GROUPS) return fail(e, ESR_E_INVALID, name + ": too many output groups");
  std::vector<TcChunk> chunks;
  for (int gi = 0; gi < p.ngroups; ++gi) {
    const TcGroupDecl& gd = c.groups[gi];
    TcOutGroup& g = p.g[gi];
    g.col0 = gd.col0; g.ncols = gd.ncols; g.act = gd.act;
    g.slope = gd.act == ACT_RELU ? 0.f : (gd.act == ACT_LRELU ? gd.slope : 1.f);   // NONE/RELU/LRELU = max(v, v*slope)
    g.res_after = gd.res_after;
    g.mode = gd.mode;
    // staging rows are one pixel each (ncols * 2 bytes); the matching TMA swizzle keeps the 16-byte staging
    // stores of a quarter warp on distinct banks (linear 64 B / 32 B rows conflict 4-way / 2-way)
    g.swizzle = gd.ncols == 64 ? 1 : (gd.ncols == 32 ? 2 : (gd.ncols == 16 ? 3 : 0));
    g.bias = dg.d_params + gd.off_bias;
    g.bias9 = gd.off_bias9 >= 0 ? dg.d_params + gd.off_bias9 : nullptr;
    if (g.bias9 && gi != 0) return fail(e, ESR_E_INVALID, name + ": only the first output group may carry a border-class bias");
    g.res = nullptr;
    if (gd.res != BUF_NONE && gi != 0) return fail(e, ESR_E_INVALID, name + ": only the first output group may carry a residual");
    if (gd.res != BUF_NONE) {
      g.res = reinterpret_cast<const __half*>(ws + L.off[gd.res]);
      g.res_stride = dg.g.bufs[gd.res].C;
      g.res_coff = gd.res_coff;
    }
    for (int c0 = 0; c0 < gd.ncols;) {
      const int wdt = 16;
      TcChunk ck;
      memset(&ck, 0, sizeof(ck));
      ck.tcol = (uint16_t)(gd.col0 + c0); ck.group = (uint8_t)gi; ck.c0 = (uint8_t)c0; ck.width = (uint8_t)wdt;
      chunks.push_back(ck);
      c0 += wdt;
    }
    if (gd.mode == 0) {
      g.stage_off = (int)off;
      g.stage_bytes = (TC_TILE_PX * gd.ncols * 2 + 1023) / 1024 * 1024;
      off += 2 * (size_t)g.stage_bytes;
      const int Cs = dg.g.bufs[gd.out].C;
      __half* base = reinterpret_cast<__half*>(ws + L.off[gd.out]) + gd.out_coff;
      int rc = make_tensor_map(e, &pk->tmO[gi], base, Cs, gd.ncols, W, H, B, gd.ncols, TC_TILE_PX, g.swizzle);
      if (rc) return rc;
    }
  }
  // 16-column units alternate between the two epilogue warp sets (unit index parity)
  if ((int)chunks.size() > TC_MAX_CHUNKS) return fail(e, ESR_E_INVALID, name + ": too many epilogue chunks");
  p.n_epi_chunks = (int)chunks.size();
  for (size_t i = 0; i < chunks.size(); ++i) p.ck[i] = chunks[i];
  p.ring_off = (int)off;
  const size_t avail = kMaxSmem - 1024 - off;
  int nslots = (int)std::min<size_t>(TC_MAX_SLOTS, avail / p.strip_bytes);
  if (nslots < 2 * c.halo + 2) return fail(e, ESR_E_INVALID, name + ": shared memory budget exceeded");
  p.nslots = nslots;
  const size_t smem = off + (size_t)nslots * p.strip_bytes + 1024;
  // A operand map
  {
    const int Cs = dg.g.bufs[c.in].C;
    int rc = make_tensor_map(e, &pk->tmA, ws + L.off[c.in], Cs, Cs, W, H, B, 64, p.strip_px, 1);
    if (rc) return rc;
  }
  for (int gi = 0; gi < TC_MAX_GROUPS; ++gi)   // unused descriptor slots must still hold a valid map (they are prefetched)
    if (gi >= p.ngroups || c.groups[gi].mode != 0) pk->tmO[gi] = pk->tmA;
  // work decomposition: items = (image, 128-px column strip, segment of R rows)
  p.strips_x = (W + TC_TILE_PX - 1) / TC_TILE_PX;
  int bestR = 1;
  if (e->opt_rows_per_item > 0) {
    bestR = std::min(e->opt_rows_per_item, H);
  } else {
    double best = 1e30;
    for (int R = 1; R <= std::min(H, 64); ++R) {
      const long long items = (long long)B * p.strips_x * ((H + R - 1) / R);
      const long long waves = (items + e->num_sms - 1) / e->num_sms;
      const double cost = (double)waves * (R + 0.35 * (R + 2 * c.halo) + 0.5);
      if (cost < best - 1e-9) { best = cost; bestR = R; }
    }
  }
  p.rows_per_item = bestR;
  p.segs_y = (H + bestR - 1) / bestR;
  p.n_items = B * p.strips_x * p.segs_y;
  for (int i = 0; i < p.n_entries; ++i) {
    const TcPlaneEntry& s = c.entries[i];
    TcEntry& d = p.e[i];
    d.a_row = (uint32_t)((s.chunk * p.chunk_bytes + (s.dx + c.halo) * 128) >> 4) | ((uint32_t)(s.dy + c.halo) << 28);
    d.b_off16 = (uint32_t)(s.b_off >> 4);
    d.idesc = umma_idesc_f16((uint32_t)s.n);
    d.misc = (uint32_t)s.dcol | ((uint32_t)(s.nsteps & 15) << 16) | (s.first ? 0x80000000u : 0u);
  }
  const int grid = std::min(p.n_items, e->num_sms);
  if (!e->attr_tc) {   // function attributes are per device: tracked per handle (a handle is bound to one device)
    CUDA_TRY(e, cudaFuncSetAttribute(conv_tc_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    CUDA_TRY(e, cudaFuncSetAttribute(conv_tc_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    CUDA_TRY(e, cudaFuncSetAttribute(conv_tc_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    CUDA_TRY(e, cudaFuncSetAttribute(conv_tc_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    CUDA_TRY(e, cudaFuncSetAttribute(conv_tc_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmem));
    e->attr_tc = true;
  }
  // instantiation by feature set: instrumented (timeline / timing experiments), BSRN (GELU, border-class bias), tail
  // (residual / gate operand, pixel-shuffle store), or plain
  bool bsrn = false, tail = false;
  for (auto& gd : c.groups) {
    bsrn = bsrn || gd.act == ACT_GELU || gd.off_bias9 >= 0;
    tail = tail || gd.res != BUF_NONE || gd.mode != 0;
  }
  const int variant = (p.dbg != nullptr || p.dbg_flags != 0) ? 4 : ((bsrn ? 2 : 0) + (tail ? 1 : 0));
  pl.launches.push_back(Launch{"conv_tc:" + name, [pk, grid, smem, variant](cudaStream_t s) {
                                 auto kern = variant == 4 ? conv_tc_kernel<true, true, true>
                                           : variant == 3 ? conv_tc_kernel<false, true, true>
                                           : variant == 2 ? conv_tc_kernel<false, true, false>
                                           : variant == 1 ? conv_tc_kernel<false, false, true> : conv_tc_kernel<false, false, false>;
                                 return launch_k(kern, dim3(grid), dim3(TC_THREADS), smem, s, pk->tmA, pk->tmO[0], pk->tmO[1],
                                                 pk->tmO[2], pk->p);
                               }});
  return ESR_OK;
}


// Is the chain starting at op `oi` executed as one fused launch for this call?
// The fused kernel is the small-image (latency) path: it is used when the bands of ONE image fit the device with one
// cluster each (a property of H and W only, so that image i of a batch is bit-identical to its single-image run whatever
// the batch size).  Larger images go through the per-layer kernel, which amortises a layer's weights over many tiles;
// the fused kernel re-streams them for every 4-row band.  Option chain_enable = 2 lifts the limit (tests).
static const size_t kChainSmemMax = 232448 - 8192;   // dynamic shared memory the chain kernel may ask for (static part: barriers, biases)
// How many clusters of `strips` chain CTAs the device holds at once (host-only handles: SM count / cluster size).
static int chain_clusters(Engine* e, int strips) {
  if (strips < 1 || strips > 8) return 0;
  if (!e->has_gpu) return e->num_sms / strips;
  if (e->chain_capacity[strips] == 0) {
    if (!e->attr_chain) {
      auto optin = [](auto kern) { return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemMax) == cudaSuccess; };
      const bool ok_attr =
          optin(conv_chain_kernel<false, false, false, true, true, true>) && optin(conv_chain_kernel<false, false, false, true, true, false>) &&
          optin(conv_chain_kernel<false, false, false, true, false, true>) && optin(conv_chain_kernel<false, false, false, true, false, false>) &&
          optin(conv_chain_kernel<false, false, false, false, true, true>) && optin(conv_chain_kernel<false, false, false, false, true, false>) &&
          optin(conv_chain_kernel<false, false, false, false, false, true>) && optin(conv_chain_kernel<false, false, false, false, false, false>) &&
          optin(conv_chain_kernel<false, true>) && optin(conv_chain_kernel<false, true, false, true, false, true>) &&
          optin(conv_chain_kernel<false, true, false, true, false, false>) &&
          optin(conv_chain_kernel<false, false, true>) && optin(conv_chain_kernel<true, false, true>);
      if (!ok_attr ||
          cudaFuncSetAttribute(conv_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kChainSmemMax) != cudaSuccess) {
        cudaGetLastError();
        return 0;
      }
      e->attr_chain = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(strips * (e->num_sms / strips))); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = kChainSmemMax;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)strips; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, conv_chain_kernel<true>, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = -1; }
    e->chain_capacity[strips] = n;
  }
  return std::max(e->chain_capacity[strips], 0);
}
static const ChainDecl* chain_at(Engine* e, const Graph& g, int oi, int B, int H, int W) {
  if (!e->opt_chain) return nullptr;
  const int strips = (W + TC_TILE_PX - 1) / TC_TILE_PX;
  if (strips > 8) return nullptr;   // one cluster (<= 8 CTAs) spans the strips of a band
  if (e->opt_chain != 2) {
    int max_layers = 2;
    for (auto& ch : g.chains) max_layers = std::max(max_layers, ch.n_ops);
    const long long bands = (H + max_layers - 1 + CH_R - 1) / CH_R;   // per image: the choice must not depend on the batch
    if (bands > chain_clusters(e, strips)) return nullptr;              // size (image i of a batch == its single-image run)
    (void)B;
  }
  for (size_t k = 0; k < g.chains.size(); ++k)
    if (g.chains[k].first_op == oi) return ((e->opt_chain_mask >> (k < 31 ? k : 31)) & 1) ? &g.chains[k] : nullptr;
  return nullptr;
}

static int plan_chain(Engine* e, const DevGraph& dg, const ChainDecl& ch, const WsLayout& L, size_t flags_off, Plan& pl,
                      std::string* name_out) {
  struct Packed {
    ChainMaps maps;
    ChainParams p;
  };
  auto pk = std::make_shared<Packed>();
  memset(pk.get(), 0, sizeof(Packed));
  ChainParams& p = pk->p;
  const Graph& g = dg.g;
  const int B = pl.B, H = pl.H, W = pl.W, nL = ch.n_ops;
  uint8_t* ws = reinterpret_cast<uint8_t*>(pl.ws);
  p.B = B; p.H = H; p.W = W; p.n_layers = nL;
  p.strips = (W + TC_TILE_PX - 1) / TC_TILE_PX;
  p.nbands = (H + nL - 1 + CH_R - 1) / CH_R;
  p.n_items = B * p.nbands;
  p.store_all = e->opt_chain_store_all;
  p.dbg_flags = e->opt_dbg_flags;
  p.ps_fp32 = 0;
  p.ps_out = pl.out;
  p.ps_u8 = 0; p.ps_dr = 1.f;
  // uint8 output (esr_forward_u8): when this chain ends in the pixel-shuffle layer, tensor2uint happens in its epilogue
  // and the bytes go straight into the caller's HWC image - no NCHW intermediate, no conversion launch
  if (pl.u8 && !(ch.pw_tc >= 0 && e->opt_chain_pw) && g.tc[ch.layers[nL - 1].tc].groups[0].mode == 1 &&
      g.tc[ch.layers[nL - 1].tc].groups.size() == 1 &&
      (reinterpret_cast<uintptr_t>(pl.u8_out) & 3) == 0) {
    p.ps_u8 = 1;
    p.ps_dr = pl.data_range;
    p.ps_out = pl.u8_out;
    pl.u8_folded = true;
  }
  p.flags = reinterpret_cast<int32_t*>(ws + flags_off);
  p.item_counter = p.flags + chain_flag_ints(B, H, W, nL) - 1;
  p.wblob = dg.d_blobs + ch.off_blob;
  p.dbg = nullptr;
  if (e->opt_timeline) {
    if (!e->d_timeline) {
      CUDA_TRY(e, cudaMalloc(&e->d_timeline, 256 * 128 * sizeof(long long)));
      CUDA_TRY(e, cudaMemset(e->d_timeline, 0, 256 * 128 * sizeof(long long)));
    }
    int idx = 0;
    for (auto& l : pl.launches) idx += l.name.rfind("conv_chain", 0) == 0 ? 1 : 0;
    if (idx < 8) p.dbg = e->d_timeline + (size_t)(224 + 2 * idx) * 128;   // chain records: two 128-stamp slots each, from slot 224
    if (idx < 8) p.dbg_blocks = e->d_timeline + (size_t)(100 + 5 * idx) * 128;   // per-block globaltimer stamps: 148 x 4 = 592 longs
  }
  int nmaps = 0;
  auto add_map = [&](void* base, int Cs, int c_extent, int box_c, int box_w, int swz) -> int {
    if (nmaps >= CH_MAX_MAPS) return -1;
    if (make_tensor_map(e, &pk->maps.m[nmaps], base, Cs, c_extent, W, H, B, box_c, box_w, swz)) return -1;
    return nmaps++;
  };
  std::string name;
  int stage_cols = 0;
  for (int l = 0; l < nL; ++l) {
    const ChainLayerDecl& d = ch.layers[l];
    const TcConv& c = g.tc[d.tc];
    ChLayer& Lr = p.L[l];
    const bool last = l == nL - 1;
    Lr.np = d.np; Lr.ksteps = d.ksteps; Lr.ctr_n = d.ctr_n; Lr.part_bytes = d.part_bytes;
    Lr.w_goff = (int)d.w_goff;
    Lr.ring_out = last ? 0 : 1;
    Lr.res_smem = d.res_smem;
    const TcGroupDecl& g0 = c.groups[0];
    auto slope_of = [](const TcGroupDecl& gd) { return gd.act == ACT_RELU ? 0.f : (gd.act == ACT_LRELU ? gd.slope : 1.f); };
    Lr.n0 = d.n0; Lr.slope0 = slope_of(g0); Lr.mode0 = g0.mode;
    Lr.swz0 = d.n0 == 64 ? 1 : (d.n0 == 32 ? 2 : (d.n0 == 16 ? 3 : 0));
    Lr.res = nullptr;
    if (g0.res != BUF_NONE) {
      Lr.res = reinterpret_cast<const __half*>(ws + L.off[g0.res]);
      Lr.res_stride = g.bufs[g0.res].C;
      Lr.res_coff = g0.res_coff;
      Lr.res_after = g0.res_after;
    }
    Lr.n1 = d.n1; Lr.g1_ctr = d.g1_ctr; Lr.col1 = d.col1;
    // input maps (box 128 px and box 8 px): layer 0 reads the chain's input buffer, layer l > 0 the global copy of
    // layer l-1's rows (chain buffer l-1)
    {
      __half* base;
      int Cs;
      if (l == 0) { Cs = g.bufs[c.in].C; base = reinterpret_cast<__half*>(ws + L.off[c.in]) + c.chunk_c0[0]; }
      else { Cs = 64; base = reinterpret_cast<__half*>(ws + L.off[g.chain_buf[l - 1]]); }
      Lr.map_in = add_map(base, Cs, 64, 64, TC_TILE_PX, 1);
      const int m8 = add_map(base, Cs, 64, 64, 8, 1);
      if (Lr.map_in < 0 || m8 != Lr.map_in + 1) return fail(e, ESR_E_INVALID, "chain: too many tensor maps");
    }
    if (!last) {
      if (l >= 3) return fail(e, ESR_E_INVALID, "chain: more than three intermediate layers");
      Lr.map_out = add_map(ws + L.off[g.chain_buf[l]], 64, 64, 64, TC_TILE_PX, 1);
    } else if (g0.mode == 0) {
      const int Cs = g.bufs[g0.out].C;
      __half* base = reinterpret_cast<__half*>(ws + L.off[g0.out]) + g0.out_coff;
      Lr.map_out = add_map(base, Cs, d.n0, d.n0, TC_TILE_PX, Lr.swz0);
      stage_cols = std::max(stage_cols, d.n0);
    }
    if (d.n1 > 0) {
      const TcGroupDecl& g1 = c.groups[1];
      Lr.slope1 = slope_of(g1);
      Lr.swz1 = d.n1 == 64 ? 1 : (d.n1 == 32 ? 2 : (d.n1 == 16 ? 3 : 0));
      const int Cs = g.bufs[g1.out].C;
      __half* base = reinterpret_cast<__half*>(ws + L.off[g1.out]) + g1.out_coff;
      Lr.map_g1 = add_map(base, Cs, d.n1, d.n1, TC_TILE_PX, Lr.swz1);
      stage_cols = std::max(stage_cols, d.n1);
    }
    if (Lr.map_out < 0 || Lr.map_g1 < 0) return fail(e, ESR_E_INVALID, "chain: too many tensor maps");
    name += (l ? " | " : "") + g.ops[ch.first_op + l].name;
  }
  // TMEM: the centre blocks' accumulators sit at columns [256, 384); a layer hands its accumulator of output row j to
  // the next layer row by row (ready[j]), which is only valid when both use the same columns for row j.  A layer whose
  // width differs from its predecessor's therefore gets a region disjoint from it.
  {
    bool any_ctr = false;
    for (int l = 0; l < nL; ++l) any_ctr = any_ctr || ch.layers[l].ctr_n > 0;
    p.ctr_acc_col = 256;
    for (int l = 0; l < nL; ++l) {
      ChLayer& Lr = p.L[l];
      if (l == 0) { Lr.acc_col = 0; continue; }
      const ChLayer& P = p.L[l - 1];
      if (P.np == Lr.np) { Lr.acc_col = P.acc_col; continue; }
      const int need = CH_R * Lr.np;
      int pick = -1;
      for (int cand : {0, 256, 384}) {
        if (cand + need > 512) continue;
        if (any_ctr && cand < 384 && cand + need > 256) continue;                       // the centre blocks' columns
        if (cand < P.acc_col + CH_R * P.np && P.acc_col < cand + need) continue;         // the predecessor's columns
        pick = cand;
        break;
      }
      if (pick < 0) return fail(e, ESR_E_INVALID, "chain: no disjoint TMEM region for layer " + std::to_string(l));
      Lr.acc_col = pick;
    }
  }
  if (ch.pw_tc >= 0 && e->opt_chain_pw) {
    const TcConv& c = g.tc[ch.pw_tc];
    ChPw& P = p.pw;
    const ChLayer& last = p.L[nL - 1];
    P.enabled = 1;
    P.nchunks = c.nchunks;
    P.ksteps = 1;
    for (auto& en : c.entries) P.ksteps = std::max(P.ksteps, en.nsteps);
    P.n = (c.acc_cols + 15) / 16 * 16;
    P.w_bytes = c.nchunks * P.n * 128;
    P.w = dg.d_blobs + c.off_blob;
    P.bias_goff = (int)ch.pw_bias_goff;
    // weights: behind the last 3x3 layer's parts if they fit (then they can arrive one layer ahead)
    if (3 * last.part_bytes + P.w_bytes <= CH_W_BYTES) { P.w_soff = 3 * last.part_bytes; P.w_early = 1; }
    else { P.w_soff = 0; P.w_early = 0; }
    // accumulators: two row slots, disjoint from the last 3x3 layer's columns
    P.acc_col = -1;
    for (int cand : {0, 256}) {
      if (cand + 2 * P.n > 512) continue;
      if (cand < last.acc_col + CH_R * last.np && last.acc_col < cand + 2 * P.n) continue;
      P.acc_col = cand;
      break;
    }
    if (P.acc_col < 0) return fail(e, ESR_E_INVALID, "chain: no TMEM region for the pointwise stage");
    {
      const int Cs = g.bufs[c.in].C;
      P.map_in = add_map(ws + L.off[c.in], Cs, Cs, 64, TC_TILE_PX, 1);
      for (int k = 0; k < c.nchunks; ++k) P.chunk_c0[k] = c.chunk_c0[k];
    }
    P.ngroups = (int)c.groups.size();
    int n64 = 0;
    for (int k = 0; k < P.ngroups; ++k) {
      const TcGroupDecl& gd = c.groups[k];
      P.g_col0[k] = gd.col0; P.g_ncols[k] = gd.ncols;
      P.g_slope[k] = gd.act == ACT_RELU ? 0.f : (gd.act == ACT_LRELU ? gd.slope : 1.f);
      P.g_stage[k] = gd.ncols == 64 ? n64++ : 2;
      const int Cs = g.bufs[gd.out].C;
      __half* base = reinterpret_cast<__half*>(ws + L.off[gd.out]) + gd.out_coff;
      P.g_map[k] = add_map(base, Cs, gd.ncols, gd.ncols, TC_TILE_PX, gd.ncols == 64 ? 1 : 3);
      if (gd.ncols == 16) stage_cols = std::max(stage_cols, 16);
      if (P.g_map[k] < 0) return fail(e, ESR_E_INVALID, "chain: too many tensor maps");
    }
    P.n64 = n64;
    P.nbuf = (n64 <= 1 || P.w_soff >= 32768) ? 2 : 1;
    const TcGroupDecl& g0 = c.groups[0];
    if (g0.res != BUF_NONE) {
      P.res = reinterpret_cast<const __half*>(ws + L.off[g0.res]);
      P.res_stride = g.bufs[g0.res].C;
      P.res_coff = g0.res_coff;
      P.res_after = g0.res_after;
    }
    if (P.map_in < 0) return fail(e, ESR_E_INVALID, "chain: too many tensor maps");
    {   // the last 3x3 layer's output is a channel range of this stage's input (r4 inside [d1|d2|d3|r4]): hand it over in
        // shared memory instead of through a store + load
      const TcGroupDecl& lg = g.tc[ch.layers[nL - 1].tc].groups[0];
      const int rel = lg.out_coff - c.chunk_c0[0];
      if (lg.out == c.in && lg.mode == 0 && lg.res == BUF_NONE && rel >= 0 && rel / 64 < c.nchunks && rel % 16 == 0 &&
          rel % 64 + last.n0 <= 64 && last.n0 % 16 == 0 && last.n1 == 0) {
        P.from_smem = 1;
        P.fs_chunk = rel / 64;
        P.fs_lane0 = rel % 64;
      }
    }
    name += " | " + g.ops[ch.first_op + nL].name;
  }
  for (int i = nmaps; i < CH_MAX_MAPS; ++i) pk->maps.m[i] = pk->maps.m[0];   // every slot is prefetched: keep them valid
  // shared memory carve-up
  p.ring_off = 0;
  p.w_off = CH_SLOTS * CH_SLOT_BYTES;
  p.ctr_off = p.w_off + CH_W_BYTES;
  // the centre block and the identity block only exist in chains that use them (an IMDB / FMEN chain whose last layer
  // stages 64 columns needs the room)
  bool any_ctr_blk = false, any_ident = false;
  for (int l = 0; l < nL; ++l) { any_ctr_blk = any_ctr_blk || ch.layers[l].ctr_n > 0; any_ident = any_ident || ch.layers[l].res_smem != 0; }
  p.ident_off = p.ctr_off + (any_ctr_blk ? CH_CTR_BYTES : 0);
  p.ident_bytes = any_ident ? CH_IDENT_BYTES : 0;
  p.stage_off = p.ident_off + p.ident_bytes;
  p.ident = dg.d_ident;
  if (p.ps_u8) stage_cols = std::max(stage_cols, 24);   // one LR row of uint8 HWC output: 4 x 1536 bytes
  p.stage_bytes = (TC_TILE_PX * std::max(stage_cols, 16) * 2 + 1023) / 1024 * 1024;
  const size_t smem = (size_t)p.stage_off + 2 * (size_t)p.stage_bytes + 1024;
  if (smem > kChainSmemMax) return fail(e, ESR_E_INVALID, "chain: shared memory budget exceeded");
  p.tmem_cols = 512;
  // persistent grid: as many clusters (one per band in flight) as fit the device
  const int strips = p.strips;
  const int max_clusters = std::max(1, chain_clusters(e, strips));
  const int nclusters = std::max(1, std::min(p.n_items, max_clusters));
  const int grid = nclusters * strips;
  pl.launches.push_back(Launch{"conv_chain:" + name, [pk, grid, strips, smem](cudaStream_t s) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)strips; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = g_pdl ? 2 : 1;
    // Instantiation by feature set (the single-thread roles pay for every instruction-cache line: each of these cuts
    // measured 3-6 % off a chain): `tail` = any layer with a global residual / gate operand or the pixel-shuffle store,
    // `ctr` = any layer with a centre block, `dyn` = more bands than clusters in flight (dynamic band scheduling)
    bool tail = false, ctr = false;
    for (int l = 0; l < pk->p.n_layers; ++l) {
      tail = tail || pk->p.L[l].res != nullptr || pk->p.L[l].mode0 != 0;
      ctr = ctr || pk->p.L[l].ctr_n > 0 || pk->p.L[l].res_smem != 0;   // (the identity tap lives in the same instantiations)
    }
    const bool dyn = pk->p.n_items * strips > grid;
    if (pk->p.ps_u8) {
      if (ctr) return cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, true>, pk->maps, pk->p);
      return dyn ? cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, true, false, true, false, true>, pk->maps, pk->p)
                 : cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, true, false, true, false, false>, pk->maps, pk->p);
    }
    if (pk->p.dbg != nullptr || pk->p.dbg_flags != 0)     // timeline stamps / timing experiments: the instrumented instantiations
      return pk->p.pw.enabled ? cudaLaunchKernelEx(&cfg, conv_chain_kernel<true, false, true>, pk->maps, pk->p)
                              : cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, false, true>, pk->maps, pk->p);
    if (pk->p.pw.enabled) return cudaLaunchKernelEx(&cfg, conv_chain_kernel<true>, pk->maps, pk->p);
    const int v = (tail ? 4 : 0) + (ctr ? 2 : 0) + (dyn ? 1 : 0);
    switch (v) {
      case 7: return cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, false, false, true, true, true>, pk->maps, pk->p);
      case 6: return cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, false, false, true, true, false>, pk->maps, pk->p);
      case 5: return cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, false, false, true, false, true>, pk->maps, pk->p);
      case 4: return cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, false, false, true, false, false>, pk->maps, pk->p);
      case 3: return cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, false, false, false, true, true>, pk->maps, pk->p);
      case 2: return cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, false, false, false, true, false>, pk->maps, pk->p);
      case 1: return cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, false, false, false, false, true>, pk->maps, pk->p);
      default: return cudaLaunchKernelEx(&cfg, conv_chain_kernel<false, false, false, false, false, false>, pk->maps, pk->p);
    }
  }});
  if (name_out) *name_out = name;
  return ESR_OK;
}

template <typename TAcc>
static cudaError_t launch_conv_generic(bool in_f32, bool out_f32, dim3 grid, size_t smem, cudaStream_t s,
                                       const ConvGenericParams& p) {
  if (in_f32 && out_f32) return launch1(k_conv_generic<float, float, TAcc>, grid, dim3(128), smem, s, p);
  if (!in_f32 && !out_f32) return launch1(k_conv_generic<__half, __half, TAcc>, grid, dim3(128), smem, s, p);
  if (!in_f32 && out_f32) return launch1(k_conv_generic<__half, float, TAcc>, grid, dim3(128), smem, s, p);
  return cudaErrorInvalidValue;
}
template <typename TAcc>
static cudaError_t launch_conv16(bool in_f32, bool out_f32, dim3 grid, cudaStream_t s, const ConvGenericParams& p) {
  if (in_f32 && out_f32) return launch1(k_conv16<float, float, TAcc>, grid, dim3(128), 0, s, p);
  if (!in_f32 && !out_f32) return launch1(k_conv16<__half, __half, TAcc>, grid, dim3(128), 0, s, p);
  if (!in_f32 && out_f32) return launch1(k_conv16<__half, float, TAcc>, grid, dim3(128), 0, s, p);
  return cudaErrorInvalidValue;
}
template <typename TAcc>
static cudaError_t launch_dw(bool in_f32, bool out_f32, dim3 grid, cudaStream_t s, const DwParams& p) {
  if (in_f32 && out_f32) return launch1(k_dwconv3x3<float, float, TAcc>, grid, dim3(128), 0, s, p);
  if (!in_f32 && !out_f32) return launch1(k_dwconv3x3<__half, __half, TAcc>, grid, dim3(128), 0, s, p);
  return cudaErrorInvalidValue;
}

static int build_plan(Engine* e, Plan& pl) {
  const DevGraph& dg = e->graphs[pl.gid];
  const Graph& g = dg.g;
  const int B = pl.B, H = pl.H, W = pl.W;
  const bool f16 = pl.dtype == ESR_DTYPE_F16;
  const WsLayout L = ws_layout(g, B, H, W, pl.dtype);
  uint8_t* ws = reinterpret_cast<uint8_t*>(pl.ws);
  const size_t elt = f16 ? 2 : 4;
  auto is_f32 = [&](int b) { return b < 0 ? !f16 : (g.bufs[b].f32 || !f16); };
  auto ptr = [&](int b) -> void* {
    if (b == BUF_IN) return const_cast<void*>(pl.in);
    if (b == BUF_OUT) return pl.out;
    if (b < 0) return nullptr;
    return ws + L.off[b];
  };
  (void)elt;
  size_t flags_off = L.flags_off;
  bool flags_cleared = false;
  for (size_t oi = 0; oi < g.ops.size(); ++oi) {
    const OpDecl& op = g.ops[oi];
    if (f16 && op.kind == OP_CONV_TC) {
      if (const ChainDecl* ch = chain_at(e, g, (int)oi, B, H, W)) {
        if (!flags_cleared) {   // one memset for the halo flags of every chain of the forward
          void* fp = ws + L.flags_off;
          const size_t fb = L.flags_bytes;
          pl.launches.insert(pl.launches.begin(), Launch{"memset:chain_flags", [fp, fb](cudaStream_t s) { return cudaMemsetAsync(fp, 0, fb, s); }});
          flags_cleared = true;
        }
        int rc = plan_chain(e, dg, *ch, L, flags_off, pl, nullptr);
        if (rc) return rc;
        flags_off += (chain_flag_ints(B, H, W, ch->n_ops) * sizeof(int) + 1023) / 1024 * 1024;
        double fl = 0;
        const int nops = ch->n_ops + ((e->opt_chain_pw && ch->pw_tc >= 0) ? 1 : 0);
        for (int k = 0; k < nops; ++k) fl += op_flops(g.ops[oi + k], B, H, W);
        pl.launches.back().flops = fl;
        oi += nops - 1;
        continue;
      }
    }
    switch (op.kind) {
      case OP_HEAD:
      case OP_BSRN_HEAD: {
        const long long npix = (long long)B * H * W;
        // head: 256-pixel row segments per block, or 128-pixel ones when those would not give every SM two blocks
        // (measured at 256x256, batch 1: 19.3 us with 128-pixel blocks against 14.3 us with 256-pixel ones: the 4-pixel
        // register tile is what keeps the kernel off the shared-memory limit; kept as a template parameter, not used)
        const bool head_small = false;
        const dim3 grid(op.kind == OP_HEAD ? (unsigned)((long long)B * H * (head_small ? (W + 127) / 128 : (W + 255) / 256))
                                           : (unsigned)((npix + 127) / 128));
        const void* in = pl.in;
        void* out = ptr(op.out);
        const int stride = g.bufs[op.out].C;
        const float* w = dg.d_params + dg.tables[op.tab].off_w;
        const float* b = dg.d_params + dg.tables[op.tab].off_b;
        const bool u8 = pl.u8;
        const float in_div = (float)(255.0 / (double)pl.data_range);   // uint2tensor4: .div(255. / data_range)
        if (op.kind == OP_HEAD) {
          pl.launches.push_back(Launch{"head:" + op.name, [=](cudaStream_t s) {
            if (u8 && f16 && head_small)
              return launch_k(k_head_conv<uint8_t, __half, float, 2>, grid, dim3(256), 0, s, (const uint8_t*)in, (__half*)out, w, b, B, H, W, stride, 64, in_div);
            if (u8 && f16)
              return launch_k(k_head_conv<uint8_t, __half, float, 4>, grid, dim3(256), 0, s, (const uint8_t*)in, (__half*)out, w, b, B, H, W, stride, 64, in_div);
            if (u8)
              return launch_k(k_head_conv<uint8_t, float, double, 4>, grid, dim3(256), 0, s, (const uint8_t*)in, (float*)out, w, b, B, H, W, stride, 64, in_div);
            if (f16 && head_small)
              return launch_k(k_head_conv<__half, __half, float, 2>, grid, dim3(256), 0, s, (const __half*)in, (__half*)out, w, b, B, H, W, stride, 64, 1.f);
            if (f16)
              return launch_k(k_head_conv<__half, __half, float, 4>, grid, dim3(256), 0, s, (const __half*)in, (__half*)out, w, b, B, H, W, stride, 64, 1.f);
            return launch_k(k_head_conv<float, float, double, 4>, grid, dim3(256), 0, s, (const float*)in, (float*)out, w, b, B, H, W, stride, 64, 1.f);
          }});
        } else {
          const float* wd = dg.d_params + dg.tables[op.tab2].off_w;
          const float* bd = dg.d_params + dg.tables[op.tab2].off_b;
          pl.launches.push_back(Launch{"bsrn_head:" + op.name, [=](cudaStream_t s) {
            if (u8 && f16)
              return launch_k(k_bsrn_head<uint8_t, __half, float>, grid, dim3(128), 0, s, (const uint8_t*)in, (__half*)out, w, b, wd, bd, B, H, W, stride, 64, in_div);
            if (u8)
              return launch_k(k_bsrn_head<uint8_t, float, double>, grid, dim3(128), 0, s, (const uint8_t*)in, (float*)out, w, b, wd, bd, B, H, W, stride, 64, in_div);
            if (f16)
              return launch_k(k_bsrn_head<__half, __half, float>, grid, dim3(128), 0, s, (const __half*)in, (__half*)out, w, b, wd, bd, B, H, W, stride, 64, 1.f);
            return launch_k(k_bsrn_head<float, float, double>, grid, dim3(128), 0, s, (const float*)in, (float*)out, w, b, wd, bd, B, H, W, stride, 64, 1.f);
          }});
        }
        break;
      }
      case OP_CONV: {
        const Table& t = dg.tables[op.tab];
        ConvGenericParams p;
        memset(&p, 0, sizeof(p));
        p.in = ptr(op.in); p.in_stride = g.bufs[op.in].C; p.in_coff = op.in_coff; p.cin8 = t.cin8;
        p.out = ptr(op.out); p.out_coff = op.out_coff; p.cout16 = t.cout16;
        p.out_stride = op.out >= 0 ? g.bufs[op.out].C : 0;
        p.w = dg.d_params + t.off_w; p.bias = dg.d_params + t.off_b;
        p.res = op.res >= 0 ? ptr(op.res) : nullptr;
        p.res_stride = op.res >= 0 ? g.bufs[op.res].C : 0;
        p.res_coff = op.res_coff; p.res_after = op.res_after;
        p.act = op.act; p.slope = op.slope;
        p.B = B; p.Hin = L.H[op.in]; p.Win = L.W[op.in];
        p.ksize = op.ksize; p.stride = op.stride; p.pad = op.pad;
        p.Hout = (p.Hin + 2 * op.pad - op.ksize) / op.stride + 1;
        p.Wout = (p.Win + 2 * op.pad - op.ksize) / op.stride + 1;
        if (op.out >= 0 && (p.Hout != L.H[op.out] || p.Wout != L.W[op.out]))
          return fail(e, ESR_E_INVALID, op.name + ": internal shape mismatch");
        p.ps_mode = op.ps ? 1 : 0;
        p.ps_fp32 = f16 ? 0 : 1;
        const long long npix = (long long)B * p.Hout * p.Wout;
        const dim3 grid((unsigned)((npix + 127) / 128), (unsigned)(t.cout16 / 16));
        const size_t smem = (size_t)t.k * t.k * t.cin8 * 16 * sizeof(float);
        if (smem > 48 * 1024) return fail(e, ESR_E_INVALID, op.name + ": generic conv weight tile too large");
        const bool in32 = is_f32(op.in), out32 = op.ps ? in32 : is_f32(op.out);
        const bool dbl = !f16;
        if (t.cin8 == 16 && t.cout16 == 16 && !op.ps) {  // ESA branch: small-conv kernel
          const long long pairs = (long long)B * p.Hout * ((p.Wout + 1) / 2);
          const dim3 grid16((unsigned)((pairs * 2 + 127) / 128));
          pl.launches.push_back(Launch{"conv16:" + op.name, [=](cudaStream_t s) {
            return dbl ? launch_conv16<double>(in32, out32, grid16, s, p) : launch_conv16<float>(in32, out32, grid16, s, p);
          }});
          break;
        }
        pl.launches.push_back(Launch{"conv_generic:" + op.name, [=](cudaStream_t s) {
          return dbl ? launch_conv_generic<double>(in32, out32, grid, smem, s, p)
                     : launch_conv_generic<float>(in32, out32, grid, smem, s, p);
        }});
        break;
      }
      case OP_DW: {
        const Table& t = dg.tables[op.tab];
        DwParams p;
        memset(&p, 0, sizeof(p));
        p.in = ptr(op.in); p.in_stride = g.bufs[op.in].C; p.in_coff = op.in_coff;
        p.out = ptr(op.out); p.out_stride = g.bufs[op.out].C; p.out_coff = op.out_coff;
        p.res = op.res >= 0 ? ptr(op.res) : nullptr;
        p.res_stride = op.res >= 0 ? g.bufs[op.res].C : 0;
        p.res_coff = op.res_coff;
        p.w = dg.d_params + t.off_w; p.bias = dg.d_params + t.off_b;
        p.c8 = t.cout16; p.act = op.act; p.slope = op.slope;
        p.B = B; p.H = L.H[op.in]; p.W = L.W[op.in];
        const long long total = (long long)B * p.H * p.W * (p.c8 / 8);
        const dim3 grid((unsigned)((total + 127) / 128));
        const bool in32 = is_f32(op.in), out32 = is_f32(op.out);
        const bool dbl = !f16;
        if (!in32 && !out32 && !dbl && p.c8 <= 64 && (op.res < 0 || !is_f32(op.res)) && p.W >= 64) {
          // full-resolution fp16 layer: 4 pixels x 8 channels per thread
          const long long t4 = (long long)B * p.H * ((p.W + 3) / 4) * (p.c8 / 8);
          const dim3 grid4((unsigned)((t4 + 255) / 256));
          pl.launches.push_back(Launch{"dwconv:" + op.name, [=](cudaStream_t s) {
            return launch1(k_dwconv3x3_h4, grid4, dim3(256), 0, s, p);
          }});
          break;
        }
        pl.launches.push_back(Launch{"dwconv:" + op.name, [=](cudaStream_t s) {
          return dbl ? launch_dw<double>(in32, out32, grid, s, p) : launch_dw<float>(in32, out32, grid, s, p);
        }});
        break;
      }
      case OP_POOL: {
        const float* in = (const float*)ptr(op.in);
        float* out = (float*)ptr(op.out);
        const int Hin = L.H[op.in], Win = L.W[op.in], Ho = L.H[op.out], Wo = L.W[op.out];
        const long long total = (long long)B * Ho * Wo * 4;
        const dim3 grid((unsigned)((total + 127) / 128));
        pl.launches.push_back(Launch{"maxpool:" + op.name, [=](cudaStream_t s) {
          return launch_k(k_maxpool7s3, grid, dim3(128), 0, s, in, out, B, Hin, Win, Ho, Wo);
        }});
        break;
      }
      case OP_ESA_APPLY: {
        EsaApplyParams p;
        memset(&p, 0, sizeof(p));
        p.x = ptr(op.in); p.x_stride = g.bufs[op.in].C; p.x_coff = op.in_coff;
        p.c1 = ptr(op.c1); p.c1_stride = g.bufs[op.c1].C; p.c1_coff = op.c1_coff;
        p.c3 = (const float*)ptr(op.c3); p.H3 = L.H[op.c3]; p.W3 = L.W[op.c3];
        p.out = ptr(op.out); p.out_stride = g.bufs[op.out].C; p.out_coff = op.out_coff;
        p.wf = dg.d_params + dg.tables[op.tab].off_w; p.bf = dg.d_params + dg.tables[op.tab].off_b;
        p.w4 = dg.d_params + dg.tables[op.tab2].off_w; p.b4 = dg.d_params + dg.tables[op.tab2].off_b;
        p.B = B; p.H = H; p.W = W; p.f = op.f; p.cgroups = op.cgroups; p.cf_ready = op.cf_ready;
        const long long total = (long long)B * H * W * op.cgroups;
        const dim3 grid((unsigned)((total + 127) / 128));
        pl.launches.push_back(Launch{"esa_apply:" + op.name, [=](cudaStream_t s) {
          if (f16) return launch_k(k_esa_apply<__half, float>, grid, dim3(128), 0, s, p);
          return launch_k(k_esa_apply<float, double>, grid, dim3(128), 0, s, p);
        }});
        break;
      }
      case OP_ESA_FRONT: {
        if (!f16) return fail(e, ESR_E_INVALID, "fused ESA front is fp16 only");
        EsaFrontParams p;
        memset(&p, 0, sizeof(p));
        const Table& t = dg.tables[op.tab];
        p.in = ptr(op.in); p.in_stride = g.bufs[op.in].C; p.in_coff = op.in_coff;
        p.w = dg.d_params + t.off_w; p.bias = dg.d_params + t.off_b;
        p.out = (float*)ptr(op.out);
        p.B = B; p.H = H; p.W = W;
        esa_dims(H, W, p.H2, p.W2, p.H3, p.W3);
        const int nblk = B * ((p.H3 + 3) / 4) * ((p.W3 + 3) / 4);
        const int f4 = op.f > 0 ? (op.f + 3) / 4 : 4;    // 4-channel groups of the ESA width
        const int variant = e->opt_esa_front_old ? 0 : f4;
        pl.launches.push_back(Launch{"esa_conv2_pool:" + op.name, [=](cudaStream_t s) {
          if (variant == 3) return launch_k(k_esa_conv2_pool4<3>, dim3(nblk), dim3(192), 0, s, p);
          if (variant == 4) return launch_k(k_esa_conv2_pool4<4>, dim3(nblk), dim3(256), 0, s, p);
          return launch_k(k_esa_conv2_pool<__half>, dim3(nblk), dim3(128), 0, s, p);
        }});
        break;
      }
      case OP_ESA_CHAIN: {
        if (!f16) return fail(e, ESR_E_INVALID, "fused ESA chain is fp16 only");
        EsaChainParams p;
        memset(&p, 0, sizeof(p));
        p.in = (const float*)ptr(op.in); p.out = (float*)ptr(op.out);
        p.npre = op.npre;
        if (op.npre > 0) { p.wpre0 = dg.d_params + dg.tables[op.tab].off_w; p.bpre0 = dg.d_params + dg.tables[op.tab].off_b; }
        if (op.npre > 1) { p.wpre1 = dg.d_params + dg.tables[op.tab2].off_w; p.bpre1 = dg.d_params + dg.tables[op.tab2].off_b; }
        p.wl = dg.d_params + dg.tables[op.tab3].off_w; p.bl = dg.d_params + dg.tables[op.tab3].off_b;
        p.B = B; p.H3 = L.H[op.in]; p.W3 = L.W[op.in];
        // 6x6 output tiles; 4x4 when those would leave most SMs idle (256x256 at batch 1: 49 tiles vs 121).  The
        // chain is latency bound there and the smaller tiles shorten every layer (fewer work items per thread)
        const int n6 = B * ((p.H3 + 5) / 6) * ((p.W3 + 5) / 6);
        const bool small = n6 < e->num_sms;
        const int ts = small ? 4 : 6;
        const int nblk = B * ((p.H3 + ts - 1) / ts) * ((p.W3 + ts - 1) / ts);
        const size_t smem = (size_t)(9 * 16 * 64 + 2 * 9 * 256 + 144 * 16 + 100 * 16) * sizeof(float);
        if (!e->attr_esa) {
          CUDA_TRY(e, cudaFuncSetAttribute(k_esa_chain<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          CUDA_TRY(e, cudaFuncSetAttribute(k_esa_chain<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          CUDA_TRY(e, cudaFuncSetAttribute(k_esa_chain<6, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          CUDA_TRY(e, cudaFuncSetAttribute(k_esa_chain<4, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          e->attr_esa = true;
        }
        const bool f3 = !e->opt_esa_front_old && op.f > 0 && op.f <= 12;   // ESA width 10 / 12: three 4-channel groups
        pl.launches.push_back(Launch{"esa_chain:" + op.name, [=](cudaStream_t s) {
          if (f3) return small ? launch_k(k_esa_chain<4, 3>, dim3(nblk), dim3(256), smem, s, p)
                               : launch_k(k_esa_chain<6, 3>, dim3(nblk), dim3(256), smem, s, p);
          if (small) return launch_k(k_esa_chain<4>, dim3(nblk), dim3(256), smem, s, p);
          return launch_k(k_esa_chain<6>, dim3(nblk), dim3(256), smem, s, p);
        }});
        break;
      }
      case OP_ESA_APPLY2: {
        if (!f16) return fail(e, ESR_E_INVALID, "commuted ESA tail is fp16 only");
        EsaApply2Params p;
        memset(&p, 0, sizeof(p));
        p.x = ptr(op.in); p.x_stride = g.bufs[op.in].C; p.x_coff = op.in_coff;
        p.cf = ptr(op.c1); p.cf_stride = g.bufs[op.c1].C; p.cf_coff = op.c1_coff;
        p.m3 = (const float*)ptr(op.c3); p.H3 = L.H[op.c3]; p.W3 = L.W[op.c3]; p.m3_stride = g.bufs[op.c3].C;
        p.out = ptr(op.out); p.out_stride = g.bufs[op.out].C; p.out_coff = op.out_coff;
        p.B = B; p.H = H; p.W = W; p.cg8 = op.cgroups;
        const long long total = (long long)B * H * W * op.cgroups;
        const int nblk = (int)std::min<long long>((total + 255) / 256, (long long)e->num_sms * 8);
        const bool small_idx = !e->opt_esa_front_old && total + (long long)nblk * 256 < (1ll << 31);
        pl.launches.push_back(Launch{"esa_apply2:" + op.name, [=](cudaStream_t s) {
          if (!small_idx) return launch_k(k_esa_apply2<__half>, dim3(nblk), dim3(256), 0, s, p);
          return launch_k(k_esa_apply2<__half, unsigned>, dim3(nblk), dim3(256), 0, s, p);
        }});
        break;
      }
      case OP_CONV_TC: {
        if (!f16) return fail(e, ESR_E_INVALID, "tcgen05 path is fp16 only");
        int rc = plan_tc(e, dg, g.tc[op.tc], op.name, L, pl);
        if (rc) return rc;
        break;
      }
      default: return fail(e, ESR_E_INVALID, "unknown op");
    }
    pl.launches.back().flops = op_flops(op, B, H, W);
  }
  if (pl.u8 && !pl.u8_folded) {   // tensor2uint of the reference, on the engine's NCHW output (folded into the tail chain's epilogue when there is one)
    const void* src = pl.out;
    uint8_t* dst = reinterpret_cast<uint8_t*>(pl.u8_out);
    const float dr = pl.data_range;
    const long long total = (long long)B * 16 * H * W;
    const int nblk = (int)std::min<long long>((total + 255) / 256, (long long)e->num_sms * 16);
    pl.launches.push_back(Launch{"tensor2uint:out", [=](cudaStream_t s) {
      if (f16) return launch_k(k_tensor2uint<__half>, dim3(nblk), dim3(256), 0, s, (const __half*)src, dst, B, 4 * H, 4 * W, dr);
      return launch_k(k_tensor2uint<float>, dim3(nblk), dim3(256), 0, s, (const float*)src, dst, B, 4 * H, 4 * W, dr);
    }});
  }
  return ESR_OK;
}

static int check_shape(Engine* e, int B, int H, int W, int dtype) {
  if (B < 1 || H < 1 || W < 1) return fail(e, ESR_E_INVALID, "B, H, W must be positive");
  if (dtype != ESR_DTYPE_F32 && dtype != ESR_DTYPE_F16) return fail(e, ESR_E_INVALID, "unknown dtype");
  if (e->arch != ESR_ARCH_IMDN && e->arch != ESR_ARCH_FMEN) {   // every other network has an ESA branch
    int H2, W2, H3, W3;
    esa_dims(H, W, H2, W2, H3, W3);
    if (H2 < 7 || W2 < 7 || H < 3 || W < 3)
      return fail(e, ESR_E_INVALID,
                  "input " + std::to_string(H) + "x" + std::to_string(W) +
                      " too small: ESA max_pool2d(7,3) needs H,W >= 15 (the reference raises here too)");
  }
  return ESR_OK;
}

static int graph_id(Engine* e, int dtype) { return (dtype == ESR_DTYPE_F16 && e->opt_tc) ? 1 : 0; }

static int ensure_graph(Engine* e, int gid) {
  if (e->graphs[gid].built) return ESR_OK;
  const std::string err = build_dev_graph(*e, gid);
  if (!err.empty()) return fail(e, ESR_E_WEIGHTS, err);
  return ESR_OK;
}

struct IoMode {   // uint8 HWC I/O of esr_forward_u8 (u8 = false: the plain NCHW float tensors of esr_forward)
  bool u8 = false;
  float data_range = 1.f;
  void* u8_out = nullptr;
};

static Plan* get_plan(Engine* e, const void* in, void* out, int B, int H, int W, int dtype, void* ws, int& rc,
                      const IoMode& io = IoMode()) {
  const int gid = graph_id(e, dtype);
  for (auto it = e->plans.begin(); it != e->plans.end(); ++it)
    if (it->B == B && it->H == H && it->W == W && it->dtype == dtype && it->gid == gid && it->in == in && it->out == out &&
        it->ws == ws && it->u8 == io.u8 && it->u8_out == io.u8_out && (!io.u8 || it->data_range == io.data_range)) {
      e->plans.splice(e->plans.begin(), e->plans, it);
      rc = ESR_OK;
      return &e->plans.front();
    }
  rc = ensure_graph(e, gid);
  if (rc) return nullptr;
  Plan pl;
  pl.B = B; pl.H = H; pl.W = W; pl.dtype = dtype; pl.gid = gid; pl.in = in; pl.out = out; pl.ws = ws;
  pl.u8 = io.u8; pl.data_range = io.data_range; pl.u8_out = io.u8_out;
  rc = build_plan(e, pl);
  if (rc) return nullptr;
  e->plans.push_front(std::move(pl));
  while (e->plans.size() > 48) {
    if (e->plans.back().gexec) cudaGraphExecDestroy(e->plans.back().gexec);
    e->plans.pop_back();
  }
  return &e->plans.front();
}

static void drop_plans(Engine* e) {
  for (auto& p : e->plans)
    if (p.gexec) cudaGraphExecDestroy(p.gexec);
  e->plans.clear();
}

}  // namespace esr

using namespace esr;

struct esr_engine : public esr::Engine {};

extern "C" {

const char* esr_version(void) { return "esr_b200 0.1 (sm_100a)"; }

int esr_device_ok(int device) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) {
    cudaGetLastError();
    return 0;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 0;
  return prop.major == 10 ? 1 : 0;
}

int esr_create(esr_handle** out, int arch, int nf, int nblocks, int device) {
  if (!out) return ESR_E_INVALID;
  *out = nullptr;
  if (arch < 0 || arch >= kNumArch) return ESR_E_INVALID;
  esr_engine* e = new esr_engine();
  e->arch = arch;
  static const int def_nf[kNumArch] = {64, 50, 46, 48, 40, 50}, def_nb[kNumArch] = {8, 4, 4, 5, 4, 4};
  e->nf = nf > 0 ? nf : def_nf[arch];
  e->nblocks = nblocks > 0 ? nblocks : def_nb[arch];
  const bool ok_cfg = (arch == ESR_ARCH_IMDN && e->nf == 64 && e->nblocks <= 16) ||
                      ((arch == ESR_ARCH_RFDN || arch == ESR_ARCH_RFDN_PRUNED) && e->nf >= 16 && e->nf <= 64 && e->nf % 4 == 0 &&
                       e->nblocks <= 4) ||
                      (arch == ESR_ARCH_RLFN && e->nf >= 16 && e->nf <= 48 && e->nblocks <= 8) ||
                      (arch == ESR_ARCH_BSRN && e->nf == 48 && e->nblocks <= 5) ||
                      (arch == ESR_ARCH_FMEN && e->nf >= 16 && e->nf <= 64 && e->nblocks <= 8);
  if (!ok_cfg && !(arch == ESR_ARCH_RFDN && e->nf == 50)) {
    delete e;
    return ESR_E_INVALID;
  }
  e->device = device;
  if (device >= 0) {
    if (!esr_device_ok(device)) {
      delete e;
      return ESR_E_NOGPU;
    }
    e->has_gpu = true;
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    e->num_sms = prop.multiProcessorCount;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      delete e;
      return ESR_E_CUDA;
    }
    e->encode = reinterpret_cast<PFN_encodeTiled>(fn);
  }
  *out = e;
  return ESR_OK;
}

int esr_load_weights(esr_handle* h, const char* name, const float* host_ptr, const int64_t* shape, int ndim) {
  if (!h) return ESR_E_INVALID;
  if (!name || !host_ptr || ndim < 0 || ndim > 8 || (ndim > 0 && !shape)) return fail(h, ESR_E_INVALID, "bad argument");
  if (h->finalized) return fail(h, ESR_E_STATE, "esr_load_weights after esr_finalize");
  HostTensor t;
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    if (shape[i] < 0) return fail(h, ESR_E_INVALID, "negative dimension");
    t.shape.push_back(shape[i]);
    n *= (size_t)shape[i];
  }
  t.data.assign(host_ptr, host_ptr + n);
  h->weights[name] = std::move(t);
  return ESR_OK;
}

int esr_finalize(esr_handle* h) {
  if (!h) return ESR_E_INVALID;
  if (h->finalized) return fail(h, ESR_E_STATE, "esr_finalize called twice");
  DeviceGuard guard(h->has_gpu ? h->device : -1);
  // both flavours are built now so that a bad state-dict is reported here, not at the first forward
  for (int gid = 0; gid < 2; ++gid) {
    int rc = ensure_graph(h, gid);
    if (rc) return rc;
  }
  h->finalized = true;
  return ESR_OK;
}

size_t esr_workspace_bytes(esr_handle* h, int B, int H, int W, int dtype) {
  if (!h) return 0;
  if (!h->finalized) { fail(h, ESR_E_STATE, "esr_workspace_bytes before esr_finalize"); return 0; }
  if (check_shape(h, B, H, W, dtype)) return 0;
  size_t m = 0;
  for (int gid = 0; gid < 2; ++gid) m = std::max(m, ws_layout(h->graphs[gid].g, B, H, W, dtype).total);
  return m;
}

static size_t u8_nchw_bytes(int B, int H, int W, int dtype) {
  const size_t b = (size_t)B * 3 * 16 * H * W * (dtype == ESR_DTYPE_F16 ? 2 : 4);
  return (b + 1023) / 1024 * 1024;
}

size_t esr_workspace_bytes_u8(esr_handle* h, int B, int H, int W, int dtype) {
  const size_t base = esr_workspace_bytes(h, B, H, W, dtype);
  return base ? base + u8_nchw_bytes(B, H, W, dtype) : 0;
}

static int forward_impl(esr_handle* h, const void* in_nchw, void* out_nchw, int B, int H, int W, int dtype, void* workspace,
                        size_t workspace_bytes, void* stream, IoMode io) {
  if (!h) return ESR_E_INVALID;
  if (!h->finalized) return fail(h, ESR_E_STATE, "esr_forward before esr_finalize");
  if (!h->has_gpu) return fail(h, ESR_E_NOGPU, "no sm_100 device bound to this handle (the engine has no CPU fallback)");
  int rc = check_shape(h, B, H, W, dtype);
  if (rc) return rc;
  if (!in_nchw || !out_nchw || !workspace) return fail(h, ESR_E_INVALID, "null device pointer");
  const size_t need = esr_workspace_bytes(h, B, H, W, dtype);
  if (workspace_bytes < need + (io.u8 ? u8_nchw_bytes(B, H, W, dtype) : 0))
    return fail(h, ESR_E_INVALID, "workspace too small: need " +
                                      std::to_string(need + (io.u8 ? u8_nchw_bytes(B, H, W, dtype) : 0)) + " bytes");
  if (!io.u8 && ((reinterpret_cast<uintptr_t>(in_nchw) & 15) || (reinterpret_cast<uintptr_t>(out_nchw) & 15)))
    return fail(h, ESR_E_INVALID, "input / output must be 16-byte aligned");
  if (io.u8 && !(io.data_range > 0.f)) return fail(h, ESR_E_INVALID, "data_range must be positive");
  // internal buffers need 1024-byte alignment (TMA, swizzle atoms); esr_workspace_bytes includes the slack
  workspace = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  if (io.u8) {   // the network's NCHW output lives behind the activation buffers; tensor2uint reads it
    io.u8_out = out_nchw;
    out_nchw = reinterpret_cast<uint8_t*>(workspace) + (need - 1024);
  }
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, ESR_E_CUDA, "cudaSetDevice failed");
  bool zeroed = false;
  const int zgid = graph_id(h, dtype);
  for (auto& z : h->ws_zeroed)
    zeroed = zeroed || (z.p == workspace && z.need == need && z.B == B && z.H == H && z.W == W && z.dtype == dtype && z.gid == zgid);
  if (!zeroed) {
    // padded channel lanes are never written by some layers and are multiplied by zero weights later: they must hold
    // finite values.  Another shape / dtype / graph lays the buffers out differently (stale fp32 bytes read as fp16
    // can be Inf / NaN), so the cache is keyed on the whole layout; the workspace must not be modified between calls.
    CUDA_TRY(h, cudaMemsetAsync(workspace, 0, need - 1024, s));
    for (auto it = h->ws_zeroed.begin(); it != h->ws_zeroed.end();)
      it = it->p == workspace ? h->ws_zeroed.erase(it) : it + 1;
    if (h->ws_zeroed.size() >= 8) h->ws_zeroed.erase(h->ws_zeroed.begin());
    h->ws_zeroed.push_back(Engine::WsKey{workspace, need, B, H, W, dtype, zgid});
  }
  Plan* pl = get_plan(h, in_nchw, out_nchw, B, H, W, dtype, workspace, rc, io);
  if (!pl) return rc;
  // use_pdl: 1 = every launch, 2 = only the small ESA kernels (their weight prologue and launch latency overlap the
  // predecessor's tail; the 200 KB tcgen05 CTAs cannot co-reside with their predecessor, so they gain nothing)
  struct PdlScope { ~PdlScope() { g_pdl = false; } } pdl_scope;
  const int pdl_mode = h->opt_pdl;
  auto set_pdl = [pdl_mode](const Launch& l) { g_pdl = pdl_mode == 1 || (pdl_mode == 2 && l.name.compare(0, 4, "esa_") == 0); };
  if (h->opt_use_graph && !pl->graph_failed && ++pl->hits >= 2) {
    if (!pl->gexec) {
      cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
      cudaStreamIsCapturing(s, &st);
      if (st == cudaStreamCaptureStatusNone) {
        cudaStream_t cs;
        CUDA_TRY(h, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          for (auto& l : pl->launches) {
            set_pdl(l);
            if (l.fn(cs) != cudaSuccess) { ok = false; break; }
          }
          if (cudaStreamEndCapture(cs, &graph) != cudaSuccess) ok = false;
        }
        if (ok && cudaGraphInstantiate(&pl->gexec, graph, 0) != cudaSuccess) { ok = false; pl->gexec = nullptr; }
        if (graph) cudaGraphDestroy(graph);
        cudaStreamDestroy(cs);
        if (!ok) { cudaGetLastError(); pl->graph_failed = true; }
      }
    }
    if (pl->gexec) {
      CUDA_TRY(h, cudaGraphLaunch(pl->gexec, s));
      return ESR_OK;
    }
  }
  for (auto& l : pl->launches) {
    set_pdl(l);
    cudaError_t err = l.fn(s);
    if (err != cudaSuccess) return fail(h, ESR_E_CUDA, l.name + ": " + cudaGetErrorString(err));
  }
  return ESR_OK;
}

int esr_forward(esr_handle* h, const void* in_nchw, void* out_nchw, int B, int H, int W, int dtype, void* workspace,
                size_t workspace_bytes, void* stream) {
  return forward_impl(h, in_nchw, out_nchw, B, H, W, dtype, workspace, workspace_bytes, stream, IoMode());
}

int esr_forward_u8(esr_handle* h, const uint8_t* in_hwc, uint8_t* out_hwc, int B, int H, int W, float data_range, int dtype,
                   void* workspace, size_t workspace_bytes, void* stream) {
  IoMode io;
  io.u8 = true;
  io.data_range = data_range;
  return forward_impl(h, in_hwc, out_hwc, B, H, W, dtype, workspace, workspace_bytes, stream, io);
}

int esr_host_wait(esr_handle* h, long long ticket) {
  if (!h) return ESR_E_INVALID;
  if (!h->has_gpu) return fail(h, ESR_E_NOGPU, "no sm_100 device bound to this handle (the engine has no CPU fallback)");
  DeviceGuard guard(h->device);
  for (auto& sl : h->hslot) {
    if (!sl.busy) continue;
    if (ticket >= 0 && sl.ticket > ticket) continue;   // requests complete in order: wait for everything up to `ticket`
    CUDA_TRY(h, cudaEventSynchronize(sl.ev_done));
    sl.busy = false;
  }
  return ESR_OK;
}

static int forward_host_impl(esr_handle* h, const void* in_host, void* out_host, int B, int H, int W, int dtype,
                             long long* ticket_out, IoMode io) {
  if (!h) return ESR_E_INVALID;
  if (!h->finalized) return fail(h, ESR_E_STATE, "esr_forward_host before esr_finalize");
  if (!h->has_gpu) return fail(h, ESR_E_NOGPU, "no sm_100 device bound to this handle (the engine has no CPU fallback)");
  int rc = check_shape(h, B, H, W, dtype);
  if (rc) return rc;
  if (!in_host || !out_host) return fail(h, ESR_E_INVALID, "null host pointer");
  DeviceGuard guard(h->device);
  if (!guard.ok) return fail(h, ESR_E_CUDA, "cudaSetDevice failed");
  const size_t elt = io.u8 ? 1 : (dtype == ESR_DTYPE_F16 ? 2 : 4);
  const size_t in_b = (size_t)B * 3 * H * W * elt, out_b = in_b * 16;
  const size_t ws_b = io.u8 ? esr_workspace_bytes_u8(h, B, H, W, dtype) : esr_workspace_bytes(h, B, H, W, dtype);
  if (!h->s_h2d) {
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->s_h2d, cudaStreamNonBlocking));
    CUDA_TRY(h, cudaStreamCreateWithFlags(&h->s_d2h, cudaStreamNonBlocking));
    for (auto& sl : h->hslot) {
      CUDA_TRY(h, cudaStreamCreateWithFlags(&sl.s_cmp, cudaStreamNonBlocking));
      CUDA_TRY(h, cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
      CUDA_TRY(h, cudaEventCreateWithFlags(&sl.ev_fwd, cudaEventDisableTiming));
      CUDA_TRY(h, cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    }
  }
  const long long ticket = h->next_ticket++;
  esr::Engine::HostSlot& sl = h->hslot[ticket % esr::Engine::kHostSlots];
  if (sl.busy) {   // the slot's previous request must have left the device buffers
    CUDA_TRY(h, cudaEventSynchronize(sl.ev_done));
    sl.busy = false;
  }
  auto grow = [&](void*& p, size_t& have, size_t need) -> cudaError_t {
    if (have >= need) return cudaSuccess;
    cudaError_t err = cudaDeviceSynchronize();   // nothing may still use the old buffer
    if (err != cudaSuccess) return err;
    if (p) cudaFree(p);
    p = nullptr; have = 0;
    err = cudaMalloc(&p, need);
    if (err == cudaSuccess) have = need;
    return err;
  };
  CUDA_TRY(h, grow(sl.d_in, sl.in_sz, in_b));
  CUDA_TRY(h, grow(sl.d_out, sl.out_sz, out_b));
  if (sl.ws_sz < ws_b) {   // a regrown workspace is a new allocation: forget what was cleared at the old address
    for (auto it = h->ws_zeroed.begin(); it != h->ws_zeroed.end();)
      it = it->p == reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(sl.ws) + 1023) & ~uintptr_t(1023)) ? h->ws_zeroed.erase(it) : it + 1;
  }
  CUDA_TRY(h, grow(sl.ws, sl.ws_sz, ws_b));
  CUDA_TRY(h, cudaMemcpyAsync(sl.d_in, in_host, in_b, cudaMemcpyHostToDevice, h->s_h2d));
  CUDA_TRY(h, cudaEventRecord(sl.ev_in, h->s_h2d));
  CUDA_TRY(h, cudaStreamWaitEvent(sl.s_cmp, sl.ev_in, 0));
  rc = forward_impl(h, sl.d_in, sl.d_out, B, H, W, dtype, sl.ws, sl.ws_sz, sl.s_cmp, io);
  if (rc) return rc;
  CUDA_TRY(h, cudaEventRecord(sl.ev_fwd, sl.s_cmp));
  CUDA_TRY(h, cudaStreamWaitEvent(h->s_d2h, sl.ev_fwd, 0));
  CUDA_TRY(h, cudaMemcpyAsync(out_host, sl.d_out, out_b, cudaMemcpyDeviceToHost, h->s_d2h));
  CUDA_TRY(h, cudaEventRecord(sl.ev_done, h->s_d2h));
  sl.busy = true;
  sl.ticket = ticket;
  if (ticket_out) *ticket_out = ticket;
  return ESR_OK;
}

int esr_forward_host_async(esr_handle* h, const void* in_host, void* out_host, int B, int H, int W, int dtype,
                           long long* ticket_out) {
  return forward_host_impl(h, in_host, out_host, B, H, W, dtype, ticket_out, IoMode());
}

int esr_forward_host_u8_async(esr_handle* h, const uint8_t* in_host_hwc, uint8_t* out_host_hwc, int B, int H, int W,
                              float data_range, int dtype, long long* ticket_out) {
  IoMode io;
  io.u8 = true;
  io.data_range = data_range;
  return forward_host_impl(h, in_host_hwc, out_host_hwc, B, H, W, dtype, ticket_out, io);
}

int esr_forward_host_u8(esr_handle* h, const uint8_t* in_host_hwc, uint8_t* out_host_hwc, int B, int H, int W, float data_range,
                        int dtype) {
  long long ticket = -1;
  int rc = esr_forward_host_u8_async(h, in_host_hwc, out_host_hwc, B, H, W, data_range, dtype, &ticket);
  if (rc) return rc;
  return esr_host_wait(h, ticket);
}

int esr_forward_host(esr_handle* h, const void* in_host, void* out_host, int B, int H, int W, int dtype) {
  long long ticket = -1;
  int rc = esr_forward_host_async(h, in_host, out_host, B, H, W, dtype, &ticket);
  if (rc) return rc;
  return esr_host_wait(h, ticket);
}

static Plan* dry_plan(esr_handle* h, int B, int H, int W, int dtype, Plan& tmp) {
  if (!h || !h->finalized || check_shape(h, B, H, W, dtype)) return nullptr;
  const int gid = graph_id(h, dtype);
  for (auto& p : h->plans)
    if (p.B == B && p.H == H && p.W == W && p.dtype == dtype && p.gid == gid) return &p;
  // names only: one launch per op
  tmp.launches.clear();
  const Graph& dgr = h->graphs[gid].g;
  bool any_chain = false;
  for (size_t oi = 0; oi < dgr.ops.size(); ++oi) {
    const OpDecl& op = dgr.ops[oi];
    if (dtype == ESR_DTYPE_F16 && op.kind == OP_CONV_TC) {
      if (const ChainDecl* ch = chain_at(h, dgr, (int)oi, B, H, W)) {
        std::string name;
        double fl = 0;
        const int nops = ch->n_ops + ((h->opt_chain_pw && ch->pw_tc >= 0) ? 1 : 0);
        for (int k = 0; k < nops; ++k) {
          name += (k ? " | " : "") + dgr.ops[oi + k].name;
          fl += op_flops(dgr.ops[oi + k], B, H, W);
        }
        tmp.launches.push_back(Launch{"conv_chain:" + name, nullptr, fl});
        any_chain = true;
        oi += nops - 1;
        continue;
      }
    }
    static const char* kn[] = {"head", "bsrn_head", "conv_generic", "dwconv", "maxpool", "esa_apply", "conv_tc", "esa_apply2",
                               "esa_conv2_pool", "esa_chain"};
    std::string kname = kn[op.kind];
    if (op.kind == OP_CONV && !op.ps && h->graphs[gid].tables[op.tab].cin8 == 16 && h->graphs[gid].tables[op.tab].cout16 == 16)
      kname = "conv16";
    tmp.launches.push_back(Launch{kname + ":" + op.name, nullptr, op_flops(op, B, H, W)});
  }
  if (any_chain) tmp.launches.insert(tmp.launches.begin(), Launch{"memset:chain_flags", nullptr, 0.0});
  return &tmp;
}

int esr_launch_count(esr_handle* h, int B, int H, int W, int dtype) {
  Plan tmp;
  Plan* p = dry_plan(h, B, H, W, dtype, tmp);
  return p ? (int)p->launches.size() : 0;
}

const char* esr_launch_name(esr_handle* h, int B, int H, int W, int dtype, int i) {
  static thread_local std::string name;
  Plan tmp;
  Plan* p = dry_plan(h, B, H, W, dtype, tmp);
  if (!p || i < 0 || i >= (int)p->launches.size()) return nullptr;
  name = p->launches[i].name;
  return name.c_str();
}

double esr_launch_flops(esr_handle* h, int B, int H, int W, int dtype, int i) {
  Plan tmp;
  Plan* p = dry_plan(h, B, H, W, dtype, tmp);
  if (!p || i < 0 || i >= (int)p->launches.size()) return 0.0;
  return p->launches[i].flops;
}

int esr_profile_launches(esr_handle* h, const void* in_nchw, void* out_nchw, int B, int H, int W, int dtype,
                         void* workspace, size_t workspace_bytes, int reps, float* ms_out, int n, void* stream) {
  if (!h) return ESR_E_INVALID;
  if (!ms_out || reps < 1) return fail(h, ESR_E_INVALID, "bad argument");
  // one ordinary forward first: validates the arguments, zeroes the workspace, builds the plan
  const int saved = h->opt_use_graph;
  h->opt_use_graph = 0;
  int rc = esr_forward(h, in_nchw, out_nchw, B, H, W, dtype, workspace, workspace_bytes, stream);
  h->opt_use_graph = saved;
  if (rc) return rc;
  Plan* pl = get_plan(h, in_nchw, out_nchw, B, H, W, dtype, workspace, rc);
  if (!pl) return rc;
  if ((int)pl->launches.size() > n) return fail(h, ESR_E_INVALID, "ms_out too small");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  cudaEvent_t e0, e1;
  CUDA_TRY(h, cudaEventCreate(&e0));
  CUDA_TRY(h, cudaEventCreate(&e1));
  cudaStream_t cs;
  CUDA_TRY(h, cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  int rc_out = (int)pl->launches.size();
  for (size_t i = 0; i < pl->launches.size(); ++i) {
    // every launch is idempotent (no op writes a buffer it reads), so repeating it in place is safe.  The
    // repeats are captured into a CUDA graph so that the CPU launch rate (~10 us per cudaLaunchKernelEx with
    // tensor-map arguments) does not hide the GPU time of short kernels.
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t gexec = nullptr;
    cudaError_t err = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
    // a fused chain needs its halo flags re-armed before every run (the memset is the plan's first launch)
    const bool is_chain = pl->launches[i].name.rfind("conv_chain", 0) == 0;
    for (int r = 0; r < reps && err == cudaSuccess; ++r) {
      if (is_chain) err = pl->launches[0].fn(cs);
      if (err == cudaSuccess) err = pl->launches[i].fn(cs);
    }
    cudaError_t err2 = cudaStreamEndCapture(cs, &graph);
    if (err == cudaSuccess) err = err2;
    if (err == cudaSuccess) err = cudaGraphInstantiate(&gexec, graph, 0);
    if (err == cudaSuccess) err = cudaGraphLaunch(gexec, s);   // warm
    if (err == cudaSuccess) err = cudaEventRecord(e0, s);
    if (err == cudaSuccess) err = cudaGraphLaunch(gexec, s);
    if (err == cudaSuccess) err = cudaEventRecord(e1, s);
    if (err == cudaSuccess) err = cudaEventSynchronize(e1);
    if (gexec) cudaGraphExecDestroy(gexec);
    if (graph) cudaGraphDestroy(graph);
    if (err != cudaSuccess) {
      rc_out = fail(h, ESR_E_CUDA, pl->launches[i].name + ": " + cudaGetErrorString(err));
      break;
    }
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    ms_out[i] = ms / reps;
  }
  cudaStreamDestroy(cs);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc_out;
}

int esr_set_option(esr_handle* h, const char* key, int value) {
  if (!h || !key) return ESR_E_INVALID;
  const std::string k = key;
  if (k == "tc_enable") h->opt_tc = value ? 1 : 0;
  else if (k == "tc_shift_mode") h->opt_shift_mode = value ? 1 : 0;
  else if (k == "use_graph") h->opt_use_graph = value ? 1 : 0;
  else if (k == "tc_rows_per_item") h->opt_rows_per_item = value;
  else if (k == "tc_timeline") h->opt_timeline = value ? 1 : 0;
  else if (k == "tc_dbg_flags") h->opt_dbg_flags = value;
  else if (k == "tc_acc_slots") h->opt_acc_slots = value == 4 ? 4 : 2;
  else if (k == "esa_front_old") h->opt_esa_front_old = value ? 1 : 0;
  else if (k == "use_pdl") h->opt_pdl = value < 0 || value > 2 ? 0 : value;
  else if (k == "chain_enable") h->opt_chain = value == 2 ? 2 : (value ? 1 : 0);
  else if (k == "chain_store_all") h->opt_chain_store_all = value ? 1 : 0;
  else if (k == "chain_mask") h->opt_chain_mask = value;
  else if (k == "chain_pw") h->opt_chain_pw = value ? 1 : 0;
  else return fail(h, ESR_E_INVALID, "unknown option: " + k);
  drop_plans(h);
  return ESR_OK;
}

int esr_debug_timeline(esr_handle* h, long long* out, int n_launches) {
  if (!h || !out) return ESR_E_INVALID;
  if (!h->d_timeline) return fail(h, ESR_E_STATE, "tc_timeline option was not enabled");
  if (n_launches > 256) n_launches = 256;
  CUDA_TRY(h, cudaDeviceSynchronize());
  CUDA_TRY(h, cudaMemcpy(out, h->d_timeline, (size_t)n_launches * 128 * sizeof(long long), cudaMemcpyDeviceToHost));
  return ESR_OK;
}

int esr_debug_tc_layer(esr_handle* h, int index, char* name, int name_cap, int32_t* meta, int32_t* entries, int32_t* groups,
                       float* bias, float* bias9, uint8_t* blob, size_t blob_cap) {
  if (!h) return ESR_E_INVALID;
  if (!h->finalized) return fail(h, ESR_E_STATE, "esr_debug_tc_layer before esr_finalize");
  const DevGraph& dg = h->graphs[1];
  const int n = (int)dg.g.tc.size();
  if (index < 0) return n;                     // query: number of tcgen05 layers of the fp16 graph
  if (index >= n || !meta || !entries || !groups || !bias || !bias9) return fail(h, ESR_E_INVALID, "bad argument");
  const TcConv& c = dg.g.tc[index];
  if (name && name_cap > 0) {
    name[0] = 0;
    for (auto& op : dg.g.ops)
      if (op.kind == OP_CONV_TC && op.tc == index) snprintf(name, (size_t)name_cap, "%s", op.name.c_str());
  }
  if (c.entries.size() > (size_t)TC_MAX_ENTRIES || c.groups.size() > (size_t)TC_MAX_GROUPS)
    return fail(h, ESR_E_INVALID, "layer exceeds the kernel's table sizes");
  meta[0] = c.nchunks; meta[1] = c.halo; meta[2] = c.acc_cols; meta[3] = (int)c.entries.size();
  meta[4] = (int)c.groups.size(); meta[5] = (int)c.blob.size(); meta[6] = 0; meta[7] = 0;
  for (int i = 0; i < 4; ++i) meta[8 + i] = c.chunk_c0[i];
  for (size_t i = 0; i < c.entries.size(); ++i) {
    const TcPlaneEntry& e = c.entries[i];
    int32_t* d = entries + 8 * i;
    d[0] = e.dy; d[1] = e.dx; d[2] = e.chunk; d[3] = e.nsteps; d[4] = e.n; d[5] = e.dcol; d[6] = e.first; d[7] = (int32_t)e.b_off;
  }
  // biases live in the parameter tables; the group records hold their arena offsets
  auto table_b = [&](size_t off) -> const std::vector<float>* {
    for (auto& t : dg.tables)
      if (t.off_b == off && !t.b.empty()) return &t.b;
    return nullptr;
  };
  for (size_t gi = 0; gi < c.groups.size(); ++gi) {
    const TcGroupDecl& gd = c.groups[gi];
    int32_t* d = groups + 8 * gi;
    d[0] = gd.col0; d[1] = gd.ncols; d[2] = gd.act; d[3] = gd.res != BUF_NONE ? 1 : 0; d[4] = gd.res_after; d[5] = gd.mode;
    memcpy(&d[6], &gd.slope, 4);
    d[7] = gd.off_bias9 >= 0 ? 1 : 0;
    const std::vector<float>* b = table_b(gd.off_bias);
    for (int j = 0; j < 64; ++j) bias[gi * 64 + j] = (b && j < (int)b->size()) ? (*b)[j] : 0.f;
    if (gi == 0 && gd.off_bias9 >= 0) {
      const std::vector<float>* b9 = table_b((size_t)gd.off_bias9);
      for (int j = 0; j < 9 * 64; ++j) bias9[j] = (b9 && j < (int)b9->size()) ? (*b9)[j] : 0.f;
      meta[7] = 1;
    }
  }
  if (blob) {
    if (blob_cap < c.blob.size()) return fail(h, ESR_E_INVALID, "blob buffer too small");
    memcpy(blob, c.blob.data(), c.blob.size());
  }
  return ESR_OK;
}

int esr_debug_chain(esr_handle* h, int index, int32_t* meta, int32_t* layers, uint8_t* blob, size_t blob_cap) {
  if (!h) return ESR_E_INVALID;
  if (!h->finalized) return fail(h, ESR_E_STATE, "esr_debug_chain before esr_finalize");
  const DevGraph& dg = h->graphs[1];
  const int n = (int)dg.g.chains.size();
  if (index < 0) return n;                     // query: number of fused chains of the fp16 graph
  if (index >= n || !meta || !layers) return fail(h, ESR_E_INVALID, "bad argument");
  const ChainDecl& ch = dg.g.chains[index];
  meta[0] = (int)ch.layers.size(); meta[1] = (int)ch.blob.size(); meta[2] = ch.layers.empty() ? -1 : ch.layers[0].tc;
  meta[3] = ch.pw_tc >= 0 ? 1 : 0;
  for (size_t l = 0; l < ch.layers.size(); ++l) {
    const ChainLayerDecl& d = ch.layers[l];
    int32_t* o = layers + 12 * l;
    o[0] = d.tc; o[1] = d.np; o[2] = d.ksteps; o[3] = d.ctr_n; o[4] = d.part_bytes; o[5] = (int32_t)d.w_goff; o[6] = d.res_smem;
    o[7] = d.n0; o[8] = d.n1; o[9] = d.g1_ctr; o[10] = d.col1; o[11] = 0;
  }
  if (blob) {
    if (blob_cap < ch.blob.size()) return fail(h, ESR_E_INVALID, "blob buffer too small");
    memcpy(blob, ch.blob.data(), ch.blob.size());
  }
  return ESR_OK;
}

const char* esr_last_error(esr_handle* h) { return h ? h->err.c_str() : "null handle"; }

void esr_destroy(esr_handle* h) {
  if (!h) return;
  if (h->has_gpu) {
    DeviceGuard guard(h->device);
    drop_plans(h);
    for (auto& dg : h->graphs) {
      if (dg.d_params) cudaFree(dg.d_params);
      if (dg.d_blobs) cudaFree(dg.d_blobs);
      if (dg.d_ident) cudaFree(dg.d_ident);
    }
    cudaDeviceSynchronize();
    for (auto& sl : h->hslot) {
      if (sl.d_in) cudaFree(sl.d_in);
      if (sl.d_out) cudaFree(sl.d_out);
      if (sl.ws) cudaFree(sl.ws);
      if (sl.s_cmp) cudaStreamDestroy(sl.s_cmp);
      if (sl.ev_in) cudaEventDestroy(sl.ev_in);
      if (sl.ev_fwd) cudaEventDestroy(sl.ev_fwd);
      if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    if (h->s_h2d) cudaStreamDestroy(h->s_h2d);
    if (h->s_d2h) cudaStreamDestroy(h->s_d2h);
    if (h->d_timeline) cudaFree(h->d_timeline);
  }
  delete h;
}

}  // extern "C"
