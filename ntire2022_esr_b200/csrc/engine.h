// Internal (non-ABI) declarations of the SR engine: host weight store, packed tables, the
// shape-independent op graph of each network, and the per-call launch plan.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/esr_b200.h"

namespace esr {

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
  bool used = false;
};

// Dense conv weights packed for the CUDA-core kernels: w[tap][cin8][cout16], bias[cout16].
// Depthwise tables use the same struct with w[9][c8], cin8 = 1, cout16 = c8.
struct Table {
  int k = 1, cin8 = 0, cout16 = 0;
  std::vector<float> w, b;
  size_t off_w = 0, off_b = 0;  // float offsets inside the engine's device parameter arena
};

// Spatial class of a workspace buffer: full LR resolution, after ESA conv2 (3x3 stride 2 pad 0),
// after the 7/3 max-pool.
enum BufKind { BK_FULL = 0, BK_S2 = 1, BK_S3 = 2 };
struct BufDecl {
  int kind;
  int C;      // channels per pixel (pixel stride)
  bool f32;   // true: always fp32; false: the call's storage dtype
};
enum { BUF_IN = -1, BUF_OUT = -2, BUF_NONE = -3 };

enum OpKind { OP_HEAD = 0, OP_BSRN_HEAD, OP_CONV, OP_DW, OP_POOL, OP_ESA_APPLY, OP_CONV_TC, OP_ESA_APPLY2, OP_ESA_FRONT, OP_ESA_CHAIN };

// ---- tcgen05 convolution, shape independent part ------------------------------------------------
struct TcPlaneEntry {   // one [n x 64] B block = one (tap, chunk, column segment)
  int dy, dx, chunk, nsteps, n, dcol, first;
  size_t b_off;
};
struct TcGroupDecl {
  int col0 = 0, ncols = 0, act = 0;
  float slope = 0.f;
  int res = BUF_NONE, res_coff = 0, res_after = 0;
  int mode = 0;                 // 0: NHWC store, 1: pixel-shuffle into the network output
  int out = BUF_NONE, out_coff = 0;
  size_t off_bias = 0;          // float offset in the parameter arena
  long long off_bias9 = -1;     // >= 0: border-class bias table [9][64] (first group only), see TcOutGroup::bias9
};
struct TcDensePlane {   // one tap of a layer before packing: w[cin (64 per chunk)][accumulator column]
  int dy = 0, dx = 0;
  bool identity = false;   // the exact `+ x` plane of a block residual
  std::vector<float> w;
};
struct TcConv {
  int in = BUF_NONE;
  int nchunks = 1;
  int chunk_c0[4] = {0, 0, 0, 0};
  int halo = 0;
  int acc_cols = 0;
  std::vector<TcPlaneEntry> entries;
  std::vector<TcGroupDecl> groups;
  std::vector<uint8_t> blob;    // pre-swizzled fp16 B blocks
  size_t off_blob = 0;          // byte offset in the device blob arena
  // unpacked form, kept for the fused-chain packing (conv_chain.cuh)
  int accP = 0;
  std::vector<TcDensePlane> dense;
  std::vector<std::pair<int, int>> segs;
  std::vector<float> bias;      // [accP]
};

// A run of consecutive 3x3 tcgen05 layers that conv_chain_kernel executes as one launch
struct ChainLayerDecl {
  int tc = -1;                  // index into Graph::tc
  int np = 0, ksteps = 0, ctr_n = 0, part_bytes = 0, res_smem = 0;
  int n0 = 0, n1 = 0, g1_ctr = 0, col1 = 0;
  size_t w_goff = 0;            // offset of the layer inside the chain blob
};
struct ChainDecl {
  int first_op = 0, n_ops = 0;  // n_ops = number of 3x3 layers
  int pw_tc = -1;               // >= 0: the op after them is a pointwise layer executed as the chain's last stage (Graph::tc index)
  size_t pw_bias_goff = 0;      // 160 bias floats (per accumulator column) inside the blob
  int n_ops_total() const { return n_ops + (pw_tc >= 0 ? 1 : 0); }
  std::vector<ChainLayerDecl> layers;
  std::vector<uint8_t> blob;    // per layer: 3 dy parts (atoms interleaved over dx), centre block, 128 bias floats
  size_t off_blob = 0;
};

struct OpDecl {
  int kind = OP_CONV;
  std::string name;
  int in = BUF_NONE, in_coff = 0;
  int out = BUF_NONE, out_coff = 0;
  int res = BUF_NONE, res_coff = 0, res_after = 0;
  int tab = -1, tab2 = -1;
  int act = 0;
  float slope = 0.f;
  int ksize = 1, stride = 1, pad = 0;
  bool ps = false;
  // ESA apply
  int c1 = BUF_NONE, c1_coff = 0, c3 = BUF_NONE, f = 0, cgroups = 0, cf_ready = 0;
  int tc = -1;  // OP_CONV_TC: index into Graph::tc
  int tab3 = -1, npre = 0;  // OP_ESA_CHAIN: tab/tab2 = the 16->16 layers, tab3 = last layer (16->64)
  double macs_pp = 0;  // algorithmic (unpadded) multiply-accumulates per output pixel
  int macs_res = BK_FULL;  // spatial class the MAC count applies to
};

struct Graph {
  std::vector<BufDecl> bufs;
  std::vector<OpDecl> ops;
  std::vector<TcConv> tc;
  std::vector<ChainDecl> chains;
  int chain_buf[3] = {BUF_NONE, BUF_NONE, BUF_NONE};   // global copies of the intermediate rows of a chain (halo exchange)
};

}  // namespace esr
