// Host side of the fused chain kernel (conv_chain.cuh): finds the runs of consecutive 3x3 tcgen05 layers of a
// graph that can execute as one launch and packs their weights into the stacked layout the kernel expects.
#pragma once
#include <cstring>

#include "conv_chain.cuh"
#include "engine.h"
#include "kernels_generic.cuh"

namespace esr {

// Can this layer be a member of a chain?  Fills the shape-independent layer record.
inline bool chain_layer_ok(const TcConv& c, ChainLayerDecl& d) {
  if (c.halo != 1 || c.nchunks != 1 || c.dense.empty() || c.segs.empty() || c.groups.empty()) return false;
  if (c.segs.size() > 2 || c.groups.size() > 2) return false;
  if (c.segs[0].first != 0) return false;
  const int n_conv = c.segs[0].second;
  if (n_conv % 16 != 0 || n_conv > 64) return false;
  for (auto& gd : c.groups) {
    if (gd.off_bias9 >= 0) return false;                      // border-class bias (BSRN): not in the fused kernel
    if (gd.act != ACT_NONE && gd.act != ACT_RELU && gd.act != ACT_LRELU) return false;
  }
  d.np = n_conv;
  d.part_bytes = n_conv * 384;
  d.res_smem = 0;
  int last_k = -1;
  for (auto& pl : c.dense) {
    if (pl.identity) {
      // the identity plane must be exactly "+ input channel c on column c": the epilogue adds the centre pixel instead
      for (int k = 0; k < 64; ++k)
        for (int n = 0; n < c.accP; ++n) {
          const float v = pl.w[(size_t)k * c.accP + n];
          if (v != ((k == n && v != 0.f) ? 1.f : 0.f)) return false;
        }
      d.res_smem = 1;
      continue;
    }
    for (int k = 0; k < 64; ++k)
      for (int n = 0; n < n_conv; ++n)
        if (pl.w[(size_t)k * c.accP + n] != 0.f) last_k = std::max(last_k, k);
  }
  d.ksteps = last_k < 0 ? 1 : last_k / 16 + 1;
  d.ctr_n = 0;
  if (c.segs.size() == 2) {
    const int col0 = c.segs[1].first, n = c.segs[1].second;
    if (col0 != n_conv || n % 16 != 0 || n > 32) return false;
    for (auto& pl : c.dense) {
      if (pl.identity) continue;
      int lk = -1;
      for (int k = 0; k < 64; ++k)
        for (int j = 0; j < n; ++j)
          if (pl.w[(size_t)k * c.accP + col0 + j] != 0.f) lk = std::max(lk, k);
      if (lk >= 0 && (pl.dy != 0 || pl.dx != 0)) return false;   // must be a centre-tap-only block
      if (lk >= 0) d.ksteps = std::max(d.ksteps, lk / 16 + 1);
    }
    d.ctr_n = n;
  }
  // output groups: group 0 = accumulator columns [0, n0); group 1 = the centre block or a column range of the conv part
  const TcGroupDecl& g0 = c.groups[0];
  if (g0.col0 != 0 || g0.ncols > n_conv) return false;
  if (d.res_smem && g0.res != BUF_NONE) return false;
  d.n0 = g0.ncols;
  d.n1 = 0; d.g1_ctr = 0; d.col1 = 0;
  if (c.groups.size() == 2) {
    const TcGroupDecl& g1 = c.groups[1];
    if (g1.res != BUF_NONE || g1.mode != 0) return false;
    if (g1.ncols % 16 != 0 || g1.ncols > 32) return false;
    d.n1 = g1.ncols;
    if (d.ctr_n > 0 && g1.col0 == n_conv && g1.ncols <= d.ctr_n) d.g1_ctr = 1;
    else if (g1.col0 % 16 == 0 && g1.col0 >= g0.ncols && g1.col0 + g1.ncols <= n_conv) d.col1 = g1.col0;
    else return false;
  } else if (d.ctr_n > 0) {
    return false;   // a centre block nobody stores
  }
  return true;
}

inline void chain_pack_layer(const TcConv& c, ChainLayerDecl& d, std::vector<uint8_t>& blob) {
  d.w_goff = (blob.size() + 127) / 128 * 128;
  blob.resize(d.w_goff + 3 * (size_t)d.part_bytes + CH_CTR_BYTES + 128 * sizeof(float), 0);
  uint8_t* base = blob.data() + d.w_goff;
  for (auto& pl : c.dense) {
    if (pl.identity) continue;
    const int dyi = 1 - pl.dy, dxi = pl.dx + 1;   // parts are ordered dy = +1, 0, -1
    for (int n = 0; n < d.np; ++n)
      for (int k = 0; k < 64; ++k) {
        const float v = pl.w[(size_t)k * c.accP + n];
        if (v == 0.f) continue;
        const int a = n >> 3, r = n & 7;
        const size_t off = (size_t)dyi * d.part_bytes + (size_t)(a * 3 + dxi) * 1024 + r * 128 + ((((k >> 3) ^ r) & 7) << 4) + (k & 7) * 2;
        const __half h = __float2half_rn(v);
        memcpy(base + off, &h, 2);
      }
    if (d.ctr_n > 0 && pl.dy == 0 && pl.dx == 0)
      for (int n = 0; n < d.ctr_n; ++n)
        for (int k = 0; k < 64; ++k) {
          const float v = pl.w[(size_t)k * c.accP + d.np + n];
          if (v == 0.f) continue;
          const __half h = __float2half_rn(v);
          memcpy(base + 3 * (size_t)d.part_bytes + sw128_offset((uint32_t)n, (uint32_t)k), &h, 2);
        }
  }
  float* bias = reinterpret_cast<float*>(base + 3 * (size_t)d.part_bytes + CH_CTR_BYTES);
  for (int cc = 0; cc < d.n0; ++cc) bias[cc] = c.bias[cc];
  const int b1 = d.g1_ctr ? d.np : d.col1;
  for (int cc = 0; cc < d.n1; ++cc) bias[64 + cc] = c.bias[b1 + cc];
}

// Can this layer run as the pointwise last stage of a chain?  One MMA per (K chunk, K step) over all its columns,
// 64-channel output groups (staged in ring slots) plus at most one 16-channel group (staging buffer).
inline bool chain_pw_ok(const TcConv& c) {
  if (c.halo != 0 || c.nchunks < 1 || c.nchunks > 2 || c.groups.empty() || c.groups.size() > 3) return false;
  const int n = (c.acc_cols + 15) / 16 * 16;
  if (n > 144 || (int)c.entries.size() != c.nchunks) return false;
  for (int ch = 0; ch < c.nchunks; ++ch) {
    const TcPlaneEntry& e = c.entries[ch];
    if (e.dy != 0 || e.dx != 0 || e.chunk != ch || e.n != n || e.dcol != 0 || e.b_off != (size_t)ch * n * 128) return false;
    if (c.chunk_c0[ch] != c.chunk_c0[0] + 64 * ch) return false;
  }
  int n64 = 0, n16 = 0, col = 0;
  for (size_t gi = 0; gi < c.groups.size(); ++gi) {
    const TcGroupDecl& gd = c.groups[gi];
    if (gd.col0 != col || gd.mode != 0 || gd.off_bias9 >= 0) return false;
    if (gd.act != ACT_NONE && gd.act != ACT_RELU && gd.act != ACT_LRELU) return false;
    if (gd.res != BUF_NONE && gi != 0) return false;
    if (gd.ncols == 64) ++n64; else if (gd.ncols == 16) ++n16; else return false;
    col += gd.ncols;
  }
  return n64 <= 2 && n16 <= 1 && col <= n;
}

// Is buffer `buf` (as written by op `writer`) read by any op after op index `from` before it is written again?
// (Inputs, residual / gate operands of tensor-core groups and of the CUDA-core ops, the ESA operands.)
inline bool chain_buffer_read_later(const Graph& g, size_t from, int buf) {
  for (size_t k = from; k < g.ops.size(); ++k) {
    const OpDecl& op = g.ops[k];
    bool writes = false;
    if (op.kind == OP_CONV_TC) {
      const TcConv& c = g.tc[op.tc];
      if (c.in == buf) return true;
      for (auto& gd : c.groups) {
        if (gd.res == buf) return true;
        writes = writes || gd.out == buf;
      }
    } else {
      if (op.in == buf || op.res == buf || op.c1 == buf || op.c3 == buf) return true;
      writes = op.out == buf;
    }
    if (writes) return false;
  }
  return false;
}

// Longest chain the planner forms: validated on hardware up to this depth (three global row buffers for the halo
// exchange, 16 of the 20 tensor maps); CH_MAX_LAYERS is the kernel's table size.
constexpr int CH_CHAIN_CAP = 4;

// Scans the op list for maximal runs of chainable layers where each layer feeds the next one.
inline void find_chains(Graph& g) {
  g.chains.clear();
  size_t i = 0;
  while (i < g.ops.size()) {
    if (g.ops[i].kind != OP_CONV_TC) { ++i; continue; }
    ChainDecl ch;
    ch.first_op = (int)i;
    size_t j = i;
    while (j < g.ops.size() && g.ops[j].kind == OP_CONV_TC && (int)ch.layers.size() < CH_CHAIN_CAP) {
      const TcConv& c = g.tc[g.ops[j].tc];
      ChainLayerDecl d;
      d.tc = g.ops[j].tc;
      if (!chain_layer_ok(c, d)) break;
      if (!ch.layers.empty()) {
        const TcConv& prev = g.tc[ch.layers.back().tc];
        const TcGroupDecl& pg = prev.groups[0];
        // the previous layer's group 0 must be exactly this layer's input, stored as plain NHWC, and its padded width
        // must cover this layer's K extent
        if (pg.mode != 0 || pg.out != c.in || pg.out_coff != c.chunk_c0[0] || d.ksteps * 16 > 64) break;
        // a chain never writes a buffer one of its earlier layers reads as residual / gate operand, nor its own input
        // (the bands of a chain run at different paces)
        const TcGroupDecl& og = c.groups[0];
        auto overlaps = [&](int buf, int c0, int n) { return buf == og.out && c0 < og.out_coff + og.ncols && og.out_coff < c0 + n; };
        const TcConv& first = g.tc[ch.layers[0].tc];
        bool clash = overlaps(first.in, first.chunk_c0[0], 64);
        for (auto& dl : ch.layers)
          for (auto& gd : g.tc[dl.tc].groups) clash = clash || (gd.res != BUF_NONE && overlaps(gd.res, gd.res_coff, gd.ncols));
        if (clash) break;
      }
      ch.layers.push_back(d);
      ++j;
      // a layer whose group 0 is the pixel-shuffle store, or that has two stored groups at the end, closes the chain
      if (c.groups[0].mode != 0) break;
      // ... and so does a layer whose output somebody other than the next layer reads later (FMEN: the input of a
      // high-frequency attention block is also its gate): only the last layer of a chain stores every row
      {
        const int ob = c.groups[0].out;                                       // (j is already the next op)
        bool other_reader = false, rewritten = false;
        if (j < g.ops.size() && g.ops[j].kind == OP_CONV_TC) {                // the next layer may read it as its input only
          for (auto& gd : g.tc[g.ops[j].tc].groups) {
            other_reader = other_reader || gd.res == ob;
            rewritten = rewritten || gd.out == ob;
          }
        }
        if (other_reader || (!rewritten && chain_buffer_read_later(g, j + 1, ob))) break;
      }
    }
    // the last layer of a chain stores group 0 itself: it must be its only group
    while (!ch.layers.empty() && g.tc[ch.layers.back().tc].groups.size() != 1) ch.layers.pop_back();
    if (ch.layers.size() >= 2) {
      ch.n_ops = (int)ch.layers.size();
      for (auto& d : ch.layers) chain_pack_layer(g.tc[d.tc], d, ch.blob);
      // a pointwise layer right behind the chain (RFDB c5 + ESA entry, IMDB conv1x1) becomes its last stage
      const size_t nxt = (size_t)ch.first_op + ch.n_ops;
      if (nxt < g.ops.size() && g.ops[nxt].kind == OP_CONV_TC && g.tc[ch.layers.back().tc].groups[0].mode == 0 &&
          chain_pw_ok(g.tc[g.ops[nxt].tc])) {
        const TcConv& pc = g.tc[g.ops[nxt].tc];
        ch.pw_tc = g.ops[nxt].tc;
        ch.pw_bias_goff = (ch.blob.size() + 127) / 128 * 128;
        ch.blob.resize(ch.pw_bias_goff + 160 * sizeof(float), 0);
        float* pb = reinterpret_cast<float*>(ch.blob.data() + ch.pw_bias_goff);
        for (int cc = 0; cc < pc.accP && cc < 160; ++cc) pb[cc] = pc.bias[cc];
      }
      i = ch.first_op + ch.n_ops;     // (the pointwise op stays in the op list: whether it is fused is a per-call option)
      g.chains.push_back(std::move(ch));
    } else {
      ++i;
    }
  }
}

}  // namespace esr
