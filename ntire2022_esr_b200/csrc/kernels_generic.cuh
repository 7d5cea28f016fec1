// CUDA-core kernels of the SR engine (NHWC activations; fp32 or fp16 storage).
// They carry the whole fp32 mode (fp64 accumulation, so the result sits at the centre of the
// reference's own fp32 rounding cloud) and, in fp16 mode, the small / irregular layers around the
// tcgen05 convolutions (3-channel head, ESA attention branch, depthwise convs of BSRN).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

namespace esr {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

enum Act : int { ACT_NONE = 0, ACT_LRELU = 1, ACT_RELU = 2, ACT_GELU = 3 };

// Exact-erf GELU of the reference (nn.GELU(), approximate='none') for the fp16 engine: erf through the
// Abramowitz-Stegun 7.1.26 rational form (|error| <= 1.5e-7) with the fast exp / reciprocal; the result is then
// rounded to fp16 (relative 4.9e-4), so it is indistinguishable from erff() there at a third of the instructions.
// The fp32 parity mode accumulates in double and uses erf().
__device__ __forceinline__ float gelu_fast(float v) {
  const float x = fabsf(v) * 0.70710678118654752440f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, x, 1.f));
  float q = fmaf(t, 1.061405429f, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  const float e = 1.f - q * t * __expf(-x * x);   // erf(|v| / sqrt 2)
  return 0.5f * v * (1.f + copysignf(e, v));
}
__device__ __forceinline__ float apply_act(float v, int act, float slope) {
  switch (act) {
    case ACT_LRELU: return v >= 0.f ? v : v * slope;
    case ACT_RELU: return fmaxf(v, 0.f);
    case ACT_GELU: return gelu_fast(v);
    default: return v;
  }
}
__device__ __forceinline__ double apply_act(double v, int act, float slope) {
  switch (act) {
    case ACT_LRELU: return v >= 0.0 ? v : v * (double)slope;
    case ACT_RELU: return fmax(v, 0.0);
    case ACT_GELU: return 0.5 * v * (1.0 + erf(v * 0.70710678118654752440));
    default: return v;
  }
}
__device__ __forceinline__ float sigmoid_acc(float z) { return 1.f / (1.f + expf(-z)); }
__device__ __forceinline__ float sigmoid_gate(float z) { return 1.f / (1.f + expf(-z)); }
__device__ __forceinline__ double sigmoid_gate(double z) { return 1.0 / (1.0 + exp(-z)); }
__device__ __forceinline__ double sigmoid_acc(double z) { return 1.0 / (1.0 + exp(-z)); }

// ---- 8-channel vector load/store helpers ------------------------------------------------------
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const __half* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __half22float2(h[i]);
    v[2 * i] = f.x; v[2 * i + 1] = f.y;
  }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(__half* p, const float (&v)[8]) {
  uint4 u;
  __half2* h = reinterpret_cast<__half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ float ldf(const float* p) { return *p; }
__device__ __forceinline__ float ldf(const __half* p) { return __half2float(*p); }
__device__ __forceinline__ void stf(float* p, float v) { *p = v; }
__device__ __forceinline__ void stf(__half* p, float v) { *p = __float2half_rn(v); }

// Network input accessor: planar NCHW fp32 / fp16 (the reference's tensor), or interleaved HWC uint8 with the
// scaling of the reference's uint2tensor4 folded into the load (utils/utils_image.py:190-193:
// float(u8) / (255 / data_range), `in_div` = 255 / data_range as fp32).  In the fp16 engine the scaled value
// is rounded to fp16 first, so the result is identical to feeding uint2tensor4(img).half().
template <typename TStore>
__device__ __forceinline__ float ldin(const float* in, int b, int ci, int y, int x, int H, int W, float) {
  return in[(((long long)b * 3 + ci) * H + y) * W + x];
}
template <typename TStore>
__device__ __forceinline__ float ldin(const __half* in, int b, int ci, int y, int x, int H, int W, float) {
  return __half2float(in[(((long long)b * 3 + ci) * H + y) * W + x]);
}
template <typename TStore>
__device__ __forceinline__ float ldin(const uint8_t* in, int b, int ci, int y, int x, int H, int W, float in_div) {
  const float v = (float)in[(((long long)b * H + y) * W + x) * 3 + ci] / in_div;
  if (std::is_same<TStore, __half>::value) return __half2float(__float2half_rn(v));
  return v;
}

// ---------------------------------------------------------------------------------------------
// Head: 3x3 conv (pad 1) on the NCHW 3-channel input, NHWC output with `cstore` channels
// (channels >= cout are written as zero so padded lanes stay finite).
// w: [27][64] fp32, index (ky*3+kx)*3+ci; bias [64].
// ---------------------------------------------------------------------------------------------
template <typename TIn, typename TOut, typename TAcc, int PX = 4>
__global__ void __launch_bounds__(256) k_head_conv(const TIn* __restrict__ in, TOut* __restrict__ out,
                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                   int B, int H, int W, int out_stride, int cstore, float in_div) {
  // block = up to 64 * PX consecutive pixels of one image row; thread = PX consecutive pixels x 16 output
  // channels (16 * PX accumulators): every weight fetched from shared memory (broadcast LDS.128) feeds PX FMAs
  // per lane, which is what keeps a CUDA-core convolution off the shared-memory bandwidth limit.  PX = 4 for large
  // workloads; PX = 2 doubles the number of blocks when a batch-1 image would otherwise leave the GPU half empty
  // (256 rows = 256 blocks of 8 warps on 148 SMs: two latency-bound waves).
  constexpr int BPX = 64 * PX;
  __shared__ __align__(16) float ws[27 * 64];
  __shared__ __align__(16) float bs[64];
  __shared__ float xs[3][3][BPX + 4];
  const int segs = (W + BPX - 1) / BPX;
  const int seg = blockIdx.x % segs;
  const int y = (blockIdx.x / segs) % H;
  const int b = blockIdx.x / (segs * H);
  const int x0 = seg * BPX;
  {
    const float4* src = reinterpret_cast<const float4*>(w);
    float4* dst = reinterpret_cast<float4*>(ws);
    for (int i = threadIdx.x; i < 27 * 16; i += 256) dst[i] = __ldg(src + i);
    if (threadIdx.x < 64) bs[threadIdx.x] = __ldg(bias + threadIdx.x);
    pdl_wait();
    for (int i = threadIdx.x; i < 3 * 3 * (BPX + 2); i += 256) {
      const int xx = i % (BPX + 2), r = (i / (BPX + 2)) % 3, ci = i / (3 * (BPX + 2));
      const int gy = y + r - 1, gx = x0 + xx - 1;
      const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
      xs[ci][r][xx] = ok ? ldin<TOut>(in, b, ci, gy, gx, H, W, in_div) : 0.f;
    }
  }
  __syncthreads();
  const int quad = threadIdx.x >> 2, g = threadIdx.x & 3;
  const int xq = x0 + quad * PX;
  if (xq >= W || g * 16 >= cstore) return;
  TAcc acc[PX][16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const TAcc bv = (TAcc)bs[g * 16 + j];
#pragma unroll
    for (int px = 0; px < PX; ++px) acc[px][j] = bv;
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
      float xv[PX + 2];
#pragma unroll
      for (int j = 0; j < PX + 2; ++j) xv[j] = xs[ci][ky][quad * PX + j];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4* w4 = reinterpret_cast<const float4*>(ws + ((ky * 3 + kx) * 3 + ci) * 64 + g * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 ww = w4[q];
          const float wv[4] = {ww.x, ww.y, ww.z, ww.w};
#pragma unroll
          for (int px = 0; px < PX; ++px) {
            const TAcc xa = (TAcc)xv[px + kx];
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[px][4 * q + j] = fma(xa, (TAcc)wv[j], acc[px][4 * q + j]);
          }
        }
      }
    }
#pragma unroll
  for (int px = 0; px < PX; ++px) {
    if (xq + px >= W) break;
    const long long pix = ((long long)b * H + y) * W + xq + px;
    TOut* o = out + pix * out_stride + g * 16;
    float f0[8], f1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { f0[j] = (float)acc[px][j]; f1[j] = (float)acc[px][8 + j]; }
    store8(o, f0);
    store8(o + 8, f1);
  }
}

// ---------------------------------------------------------------------------------------------
// BSRN head: cat(x,x,x,x) -> Linear(12->48) -> depthwise 3x3 (zero pad applied AFTER the Linear,
// models/team18_bsrn.py:82-88,218).  wpw: [3][64] (the four replicas pre-summed), bpw[64],
// wdw: [9][64], bdw[64].
// ---------------------------------------------------------------------------------------------
template <typename TIn, typename TOut, typename TAcc>
__global__ void __launch_bounds__(128) k_bsrn_head(const TIn* __restrict__ in, TOut* __restrict__ out,
                                                   const float* __restrict__ wpw, const float* __restrict__ bpw,
                                                   const float* __restrict__ wdw, const float* __restrict__ bdw,
                                                   int B, int H, int W, int out_stride, int cstore, float in_div) {
  __shared__ float s_wpw[3 * 64], s_bpw[64], s_wdw[9 * 64], s_bdw[64];
  for (int i = threadIdx.x; i < 3 * 64; i += blockDim.x) s_wpw[i] = wpw[i];
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) s_wdw[i] = wdw[i];
  if (threadIdx.x < 64) { s_bpw[threadIdx.x] = bpw[threadIdx.x]; s_bdw[threadIdx.x] = bdw[threadIdx.x]; }
  pdl_wait();
  __syncthreads();
  const long long npix = (long long)B * H * W;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int x = (int)(pix % W);
  const int y = (int)((pix / W) % H);
  const int b = (int)(pix / ((long long)W * H));
  float xin[27];
  bool okt[9];
#pragma unroll
  for (int ky = 0; ky < 3; ++ky)
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int yy = y + ky - 1, xx = x + kx - 1;
      const bool ok = yy >= 0 && yy < H && xx >= 0 && xx < W;
      okt[ky * 3 + kx] = ok;
#pragma unroll
      for (int ci = 0; ci < 3; ++ci)
        xin[(ky * 3 + kx) * 3 + ci] = ok ? ldin<TOut>(in, b, ci, yy, xx, H, W, in_div) : 0.f;
    }
  TOut* o = out + pix * out_stride;
  for (int c0 = 0; c0 < cstore; c0 += 8) {
    TAcc acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = (TAcc)s_bdw[c0 + j];
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      if (!okt[t]) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        TAcc pw = (TAcc)s_bpw[c];
        pw = fma((TAcc)xin[t * 3 + 0], (TAcc)s_wpw[0 * 64 + c], pw);
        pw = fma((TAcc)xin[t * 3 + 1], (TAcc)s_wpw[1 * 64 + c], pw);
        pw = fma((TAcc)xin[t * 3 + 2], (TAcc)s_wpw[2 * 64 + c], pw);
        // the reference rounds the Linear output to the storage type before the depthwise conv
        pw = (TAcc)(float)pw;
        acc[j] = fma(pw, (TAcc)s_wdw[t * 64 + c], acc[j]);
      }
    }
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (float)acc[j];
    store8(o + c0, f);
  }
}

// ---------------------------------------------------------------------------------------------
// Generic dense conv (k = 1 or 3, stride 1 or 2, zero pad) on NHWC buffers.
//   thread = one output pixel x 16 output columns (blockIdx.y selects the 16-column group)
//   w:    [taps][cin8][cout16] fp32 (cin8 = Cin rounded up to 8, cout16 = columns rounded to 16;
//         rows/columns beyond the logical extent are zero)
//   out:  NHWC (16 columns of the group are stored) or pixel-shuffle x4 NCHW
// residual is added before (res_after=0) or after (res_after=1) the activation; res_after=2: out = sigmoid(act(v)) * res.
// ---------------------------------------------------------------------------------------------
struct ConvGenericParams {
  const void* in; int in_stride, in_coff, cin8;
  void* out; int out_stride, out_coff, cout16;
  const float* w; const float* bias;
  const void* res; int res_stride, res_coff, res_after;
  int act; float slope;
  int B, Hin, Win, Hout, Wout, ksize, stride, pad;
  int ps_mode;      // 1: out is (B,3,4H,4W) NCHW, columns are 16*c+4*i+j
  int ps_fp32;      // dtype of the pixel-shuffled output
};

template <typename TIn, typename TOut, typename TAcc>
__global__ void __launch_bounds__(128) k_conv_generic(const ConvGenericParams p) {
  extern __shared__ float wsm[];  // [taps][cin8][16]
  const int taps = p.ksize * p.ksize;
  const int g = blockIdx.y;
  const int nw = taps * p.cin8 * 16;
  {
    float4* dst = reinterpret_cast<float4*>(wsm);
#pragma unroll 4
    for (int i = threadIdx.x; i < (nw >> 2); i += 128) {
      const int r = i >> 2, c4 = i & 3;
      dst[i] = __ldg(reinterpret_cast<const float4*>(p.w + (long long)r * p.cout16 + g * 16) + c4);
    }
  }
  pdl_wait();
  __syncthreads();
  const long long npix = (long long)p.B * p.Hout * p.Wout;
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= npix) return;
  const int x = (int)(pix % p.Wout);
  const int y = (int)((pix / p.Wout) % p.Hout);
  const int b = (int)(pix / ((long long)p.Wout * p.Hout));
  TAcc acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = (TAcc)p.bias[g * 16 + j];
  const TIn* in = reinterpret_cast<const TIn*>(p.in);
  for (int ky = 0; ky < p.ksize; ++ky) {
    const int yy = y * p.stride + ky - p.pad;
    if (yy < 0 || yy >= p.Hin) continue;
    for (int kx = 0; kx < p.ksize; ++kx) {
      const int xx = x * p.stride + kx - p.pad;
      if (xx < 0 || xx >= p.Win) continue;
      const TIn* ip = in + (((long long)b * p.Hin + yy) * p.Win + xx) * p.in_stride + p.in_coff;
      const float* wt = wsm + (ky * p.ksize + kx) * p.cin8 * 16;
      for (int c0 = 0; c0 < p.cin8; c0 += 8) {
        float xv[8];
        load8(ip + c0, xv);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4* w4 = reinterpret_cast<const float4*>(wt + (c0 + i) * 16);
          const TAcc xa = (TAcc)xv[i];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 ww = w4[q];
            acc[4 * q + 0] = fma(xa, (TAcc)ww.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fma(xa, (TAcc)ww.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fma(xa, (TAcc)ww.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fma(xa, (TAcc)ww.w, acc[4 * q + 3]);
          }
        }
      }
    }
  }
  float rv[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) rv[j] = 0.f;
  if (p.res != nullptr) {
    const TOut* rp = reinterpret_cast<const TOut*>(p.res) + pix * p.res_stride + p.res_coff + g * 16;
    float a[8], c[8];
    load8(rp, a);
    load8(rp + 8, c);
#pragma unroll
    for (int j = 0; j < 8; ++j) { rv[j] = a[j]; rv[8 + j] = c[j]; }
  }
  float o16[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    TAcc v = acc[j];
    if (!p.res_after) v += (TAcc)rv[j];
    v = apply_act(v, p.act, p.slope);
    if (p.res_after == 2) v = (TAcc)rv[j] * sigmoid_gate(v);      // gate: sigmoid(v) * res
    else if (p.res_after) v += (TAcc)rv[j];
    o16[j] = (float)v;
  }
  if (!p.ps_mode) {
    TOut* o = reinterpret_cast<TOut*>(p.out) + pix * p.out_stride + p.out_coff + g * 16;
    float a[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = o16[j]; c[j] = o16[8 + j]; }
    store8(o, a);
    store8(o + 8, c);
  } else {
    // column 16*c + 4*i + j of pixel (y,x) -> out[b, c, 4y+i, 4x+j]; this thread owns channel c = g
    const int Ho = 4 * p.Hout, Wo = 4 * p.Wout;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const long long o = (((long long)b * 3 + g) * Ho + 4 * y + i) * Wo + 4 * x;
      if (p.ps_fp32) {
        *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o) =
            make_float4(o16[4 * i], o16[4 * i + 1], o16[4 * i + 2], o16[4 * i + 3]);
      } else {
        uint2 u;
        __half2* h = reinterpret_cast<__half2*>(&u);
        h[0] = __floats2half2_rn(o16[4 * i], o16[4 * i + 1]);
        h[1] = __floats2half2_rn(o16[4 * i + 2], o16[4 * i + 3]);
        *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.out) + o) = u;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Small dense conv for the ESA branch: <= 16 input and output channels (tables with cin8 = cout16 = 16),
// k = 1 or 3, stride 1 or 2.  thread = one output pixel x 4 output channels (4 consecutive threads
// share a pixel), weights read as broadcast LDS.128.  Same parameter block as k_conv_generic.
// ---------------------------------------------------------------------------------------------
template <typename TIn, typename TOut, typename TAcc>
__global__ void __launch_bounds__(128) k_conv16(const ConvGenericParams p) {
  // thread = 2 horizontally adjacent output pixels x 8 output channels (16 accumulators against
  // 2 broadcast LDS.128 of weights per input channel); 2 consecutive threads share a pixel pair
  __shared__ __align__(16) float wsm[9 * 16 * 16];
  const int taps = p.ksize * p.ksize;
  {
    const float4* src = reinterpret_cast<const float4*>(p.w);
    float4* dst = reinterpret_cast<float4*>(wsm);
#pragma unroll 5
    for (int i = threadIdx.x; i < taps * 64; i += 128) dst[i] = __ldg(src + i);
  }
  pdl_wait();
  __syncthreads();
  const int Wp = (p.Wout + 1) >> 1;   // pixel pairs per row
  const long long total = (long long)p.B * p.Hout * Wp * 2;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g8 = (int)(idx & 1);
  const long long pp = idx >> 1;
  const int xp = (int)(pp % Wp);
  const int y = (int)((pp / Wp) % p.Hout);
  const int b = (int)(pp / ((long long)Wp * p.Hout));
  const int xa = 2 * xp, xb = 2 * xp + 1;
  const bool has_b = xb < p.Wout;
  TAcc acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { acc[0][j] = (TAcc)p.bias[g8 * 8 + j]; acc[1][j] = acc[0][j]; }
  const TIn* in = reinterpret_cast<const TIn*>(p.in);
#pragma unroll 1
  for (int ky = 0; ky < p.ksize; ++ky) {
    const int yy = y * p.stride + ky - p.pad;
    if (yy < 0 || yy >= p.Hin) continue;
#pragma unroll 1
    for (int kx = 0; kx < p.ksize; ++kx) {
      const int xxa = xa * p.stride + kx - p.pad, xxb = xb * p.stride + kx - p.pad;
      const bool oka = xxa >= 0 && xxa < p.Win, okb = has_b && xxb >= 0 && xxb < p.Win;
      float va[16], vb[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) { va[j] = 0.f; vb[j] = 0.f; }
      const TIn* row = in + ((long long)b * p.Hin + yy) * p.Win * p.in_stride + p.in_coff;
      if (oka) {
        float a[8], c[8];
        load8(row + (long long)xxa * p.in_stride, a);
        load8(row + (long long)xxa * p.in_stride + 8, c);
#pragma unroll
        for (int j = 0; j < 8; ++j) { va[j] = a[j]; va[8 + j] = c[j]; }
      }
      if (okb) {
        float a[8], c[8];
        load8(row + (long long)xxb * p.in_stride, a);
        load8(row + (long long)xxb * p.in_stride + 8, c);
#pragma unroll
        for (int j = 0; j < 8; ++j) { vb[j] = a[j]; vb[8 + j] = c[j]; }
      }
      const float* wt = wsm + (ky * p.ksize + kx) * 256 + g8 * 8;
#pragma unroll
      for (int ci = 0; ci < 16; ++ci) {
        const float4 w0 = *reinterpret_cast<const float4*>(wt + ci * 16);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + ci * 16 + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        const TAcc xa_ = (TAcc)va[ci], xb_ = (TAcc)vb[ci];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[0][j] = fma(xa_, (TAcc)wv[j], acc[0][j]);
          acc[1][j] = fma(xb_, (TAcc)wv[j], acc[1][j]);
        }
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (h == 1 && !has_b) break;
    const long long pix = ((long long)b * p.Hout + y) * p.Wout + (h ? xb : xa);
    float o8[8];
    float rv[8];
    if (p.res != nullptr) load8(reinterpret_cast<const TOut*>(p.res) + pix * p.res_stride + p.res_coff + g8 * 8, rv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      TAcc v = acc[h][j];
      if (p.res != nullptr && !p.res_after) v += (TAcc)rv[j];
      v = apply_act(v, p.act, p.slope);
      if (p.res != nullptr && p.res_after == 2) v = (TAcc)rv[j] * sigmoid_gate(v);
      else if (p.res != nullptr && p.res_after) v += (TAcc)rv[j];
      o8[j] = (float)v;
    }
    store8(reinterpret_cast<TOut*>(p.out) + pix * p.out_stride + p.out_coff + g8 * 8, o8);
  }
}

// ---------------------------------------------------------------------------------------------
// Depthwise 3x3 (pad 1) + bias (+ residual) + activation; thread = pixel x 8 channels.
// w: [9][c8] fp32, bias [c8]
// ---------------------------------------------------------------------------------------------
struct DwParams {
  const void* in; int in_stride, in_coff;
  void* out; int out_stride, out_coff;
  const void* res; int res_stride, res_coff;
  const float* w; const float* bias;
  int c8; int act; float slope; int B, H, W;
};
template <typename TIn, typename TOut, typename TAcc>
__global__ void __launch_bounds__(128) k_dwconv3x3(const DwParams p) {
  pdl_wait();
  const int groups = p.c8 >> 3;
  const long long total = (long long)p.B * p.H * p.W * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const long long pix = idx / groups;
  const int x = (int)(pix % p.W);
  const int y = (int)((pix / p.W) % p.H);
  const int b = (int)(pix / ((long long)p.W * p.H));
  TAcc acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (TAcc)p.bias[g * 8 + j];
  const TIn* in = reinterpret_cast<const TIn*>(p.in);
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
    if (yy < 0 || yy >= p.H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = x + kx - 1;
      if (xx < 0 || xx >= p.W) continue;
      float xv[8];
      load8(in + (((long long)b * p.H + yy) * p.W + xx) * p.in_stride + p.in_coff + g * 8, xv);
      const float* wt = p.w + (ky * 3 + kx) * p.c8 + g * 8;
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fma((TAcc)xv[j], (TAcc)wt[j], acc[j]);
    }
  }
  if (p.res != nullptr) {
    float rv[8];
    load8(reinterpret_cast<const TOut*>(p.res) + pix * p.res_stride + p.res_coff + g * 8, rv);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += (TAcc)rv[j];
  }
  float f[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) f[j] = (float)apply_act(acc[j], p.act, p.slope);
  store8(reinterpret_cast<TOut*>(p.out) + pix * p.out_stride + p.out_coff + g * 8, f);
}

// fp16 fast path of the same layer: thread = 4 consecutive pixels of a row x 8 channels.  Per kernel row the six
// input pixels the four outputs touch are fetched once (18 x 16-byte loads per thread for 4 outputs instead of 36),
// the weights come from shared memory as broadcast 16-byte reads.  HBM-bound layer (read in + residual, write out).
__global__ void __launch_bounds__(256) k_dwconv3x3_h4(const DwParams p) {
  __shared__ __align__(16) float ws[9 * 64];
  __shared__ __align__(16) float bs[64];
  for (int i = threadIdx.x; i < 9 * p.c8; i += 256) ws[i] = p.w[i];
  if (threadIdx.x < p.c8) bs[threadIdx.x] = p.bias[threadIdx.x];
  pdl_wait();
  __syncthreads();
  const int groups = p.c8 >> 3;
  const int quads = (p.W + 3) >> 2;
  const long long total = (long long)p.B * p.H * quads * groups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % groups);
  const long long q = idx / groups;
  const int x0 = (int)(q % quads) * 4;
  const int y = (int)((q / quads) % p.H);
  const int b = (int)(q / ((long long)quads * p.H));
  const __half* in = reinterpret_cast<const __half*>(p.in) + p.in_coff + g * 8;
  float acc[4][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float bv = bs[g * 8 + j];
    acc[0][j] = bv; acc[1][j] = bv; acc[2][j] = bv; acc[3][j] = bv;
  }
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = y + ky - 1;
    if (yy < 0 || yy >= p.H) continue;
    const __half* row = in + ((long long)b * p.H + yy) * p.W * p.in_stride;
    uint4 raw[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int xx = x0 + i - 1;
      raw[i] = (xx >= 0 && xx < p.W) ? *reinterpret_cast<const uint4*>(row + (long long)xx * p.in_stride) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const float4 w0 = *reinterpret_cast<const float4*>(ws + (ky * 3 + kx) * p.c8 + g * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(ws + (ky * 3 + kx) * p.c8 + g * 8 + 4);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&raw[px + kx]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          acc[px][2 * j] = fmaf(f.x, wv[2 * j], acc[px][2 * j]);
          acc[px][2 * j + 1] = fmaf(f.y, wv[2 * j + 1], acc[px][2 * j + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int px = 0; px < 4; ++px) {
    const int x = x0 + px;
    if (x >= p.W) break;
    const long long pix = ((long long)b * p.H + y) * p.W + x;
    if (p.res != nullptr) {
      float rv[8];
      load8(reinterpret_cast<const __half*>(p.res) + pix * p.res_stride + p.res_coff + g * 8, rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[px][j] += rv[j];
    }
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = apply_act(acc[px][j], p.act, p.slope);
    store8(reinterpret_cast<__half*>(p.out) + pix * p.out_stride + p.out_coff + g * 8, f);
  }
}

// ---------------------------------------------------------------------------------------------
// max_pool2d(kernel 7, stride 3, no pad, floor) on fp32 NHWC with 16-channel pixels.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_maxpool7s3(const float* __restrict__ in, float* __restrict__ out, int B,
                                                    int Hin, int Win, int Hout, int Wout) {
  pdl_wait();
  const long long total = (long long)B * Hout * Wout * 4;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int q = (int)(idx & 3);
  const long long pix = idx >> 2;
  const int x = (int)(pix % Wout);
  const int y = (int)((pix / Wout) % Hout);
  const int b = (int)(pix / ((long long)Wout * Hout));
  float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
  for (int i = 0; i < 7; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(
          in + (((long long)b * Hin + 3 * y + i) * Win + 3 * x + j) * 16 + q * 4);
      m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
    }
  *reinterpret_cast<float4*>(out + pix * 16 + q * 4) = m;
}

// ---------------------------------------------------------------------------------------------
// ESA tail: y = x * sigmoid(conv4(bilinear(c3) + cf))                    (block.py:123-129)
//   x   : NHWC T, `x_stride` channels per pixel
//   cf  : conv_f(c1_) when cf_ready=1 (NHWC T, 16 ch), else c1_ itself and conv_f is applied here
//   c3  : NHWC fp32 16 ch at (H3,W3)
//   wf  : [16][16] (in,out) fp32, bf[16];  w4: [16][64] (in,out), b4[64]
//   thread = pixel x 16 output channels (cgroups threads per pixel)
// ---------------------------------------------------------------------------------------------
struct EsaApplyParams {
  const void* x; int x_stride, x_coff;
  const void* c1; int c1_stride, c1_coff;
  const float* c3; int H3, W3;
  void* out; int out_stride, out_coff;
  const float* wf; const float* bf; const float* w4; const float* b4;
  int B, H, W, f, cgroups;  // f: ESA channels (<=16); cgroups: number of 16-channel output groups
  int cf_ready;
};
template <typename T, typename TAcc>
__global__ void __launch_bounds__(128) k_esa_apply(const EsaApplyParams p) {
  __shared__ __align__(16) float s_wf[16 * 16], s_bf[16], s_w4[16 * 64], s_b4[64];
  {
    if (threadIdx.x < 64) reinterpret_cast<float4*>(s_wf)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(p.wf) + threadIdx.x);
    reinterpret_cast<float4*>(s_w4)[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(p.w4) + threadIdx.x);
    reinterpret_cast<float4*>(s_w4)[threadIdx.x + 128] = __ldg(reinterpret_cast<const float4*>(p.w4) + threadIdx.x + 128);
    if (threadIdx.x < 16) s_bf[threadIdx.x] = __ldg(p.bf + threadIdx.x);
    if (threadIdx.x < 64) s_b4[threadIdx.x] = __ldg(p.b4 + threadIdx.x);
  }
  pdl_wait();
  __syncthreads();
  const long long total = (long long)p.B * p.H * p.W * p.cgroups;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int g = (int)(idx % p.cgroups);
  const long long pix = idx / p.cgroups;
  const int x = (int)(pix % p.W);
  const int y = (int)((pix / p.W) % p.H);
  const int b = (int)(pix / ((long long)p.W * p.H));
  // bilinear source coordinates, align_corners=False (ATen area_pixel_compute_source_index)
  const float sy = fmaxf(((float)y + 0.5f) * ((float)p.H3 / (float)p.H) - 0.5f, 0.f);
  const float sx = fmaxf(((float)x + 0.5f) * ((float)p.W3 / (float)p.W) - 0.5f, 0.f);
  const int y0 = min((int)sy, p.H3 - 1), x0 = min((int)sx, p.W3 - 1);
  const int y1 = min(y0 + 1, p.H3 - 1), x1 = min(x0 + 1, p.W3 - 1);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  const float* c3b = p.c3 + (long long)b * p.H3 * p.W3 * 16;
  const float4* p00 = reinterpret_cast<const float4*>(c3b + ((long long)y0 * p.W3 + x0) * 16);
  const float4* p01 = reinterpret_cast<const float4*>(c3b + ((long long)y0 * p.W3 + x1) * 16);
  const float4* p10 = reinterpret_cast<const float4*>(c3b + ((long long)y1 * p.W3 + x0) * 16);
  const float4* p11 = reinterpret_cast<const float4*>(c3b + ((long long)y1 * p.W3 + x1) * 16);
  float c1v[16];
  {
    const T* cp = reinterpret_cast<const T*>(p.c1) + pix * p.c1_stride + p.c1_coff;
    float a[8], c[8];
    load8(cp, a);
    load8(cp + 8, c);
#pragma unroll
    for (int j = 0; j < 8; ++j) { c1v[j] = a[j]; c1v[8 + j] = c[j]; }
  }
  TAcc s[16];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float4 a00 = p00[q], a01 = p01[q], a10 = p10[q], a11 = p11[q];
    const float v00[4] = {a00.x, a00.y, a00.z, a00.w}, v01[4] = {a01.x, a01.y, a01.z, a01.w};
    const float v10[4] = {a10.x, a10.y, a10.z, a10.w}, v11[4] = {a11.x, a11.y, a11.z, a11.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = q * 4 + j;
      // same association as ATen: blend along x first, then along y
      const TAcc top = (TAcc)v00[j] * (TAcc)(1.f - lx) + (TAcc)v01[j] * (TAcc)lx;
      const TAcc bot = (TAcc)v10[j] * (TAcc)(1.f - lx) + (TAcc)v11[j] * (TAcc)lx;
      const TAcc v = top * (TAcc)(1.f - ly) + bot * (TAcc)ly;
      TAcc cf;
      if (p.cf_ready) {
        cf = (TAcc)c1v[k];
      } else {
        cf = (TAcc)s_bf[k];
#pragma unroll
        for (int i = 0; i < 16; ++i) cf = fma((TAcc)c1v[i], (TAcc)s_wf[i * 16 + k], cf);
      }
      s[k] = (k < p.f) ? (v + cf) : (TAcc)0;
    }
  }
  float xv[16];
  {
    const T* xp = reinterpret_cast<const T*>(p.x) + pix * p.x_stride + p.x_coff + g * 16;
    float a[8], c[8];
    load8(xp, a);
    load8(xp + 8, c);
#pragma unroll
    for (int j = 0; j < 8; ++j) { xv[j] = a[j]; xv[8 + j] = c[j]; }
  }
  TAcc z[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) z[j] = (TAcc)s_b4[g * 16 + j];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const float4* w4 = reinterpret_cast<const float4*>(s_w4 + k * 64 + g * 16);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 ww = w4[q];
      z[4 * q + 0] = fma(s[k], (TAcc)ww.x, z[4 * q + 0]);
      z[4 * q + 1] = fma(s[k], (TAcc)ww.y, z[4 * q + 1]);
      z[4 * q + 2] = fma(s[k], (TAcc)ww.z, z[4 * q + 2]);
      z[4 * q + 3] = fma(s[k], (TAcc)ww.w, z[4 * q + 3]);
    }
  }
  float o[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) o[j] = (float)((TAcc)xv[j] * sigmoid_acc(z[j]));
  T* op = reinterpret_cast<T*>(p.out) + pix * p.out_stride + p.out_coff + g * 16;
  float a[8], c[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { a[j] = o[j]; c[j] = o[8 + j]; }
  store8(op, a);
  store8(op + 8, c);
}

// ---------------------------------------------------------------------------------------------
// ESA tail, commuted form used by the fp16 path:  y = x * sigmoid(bilinear(M3) + cf')
// where M3 = conv4 o conv3_ evaluated on the pooled map (bilinear interpolation commutes with a 1x1 conv)
// and cf' = conv4(conv_f(conv1(.))) + b4 comes out of the c5 tensor-core GEMM.  Pure elementwise:
// 8 threads per pixel x 8 channels, 16-byte accesses.
// ---------------------------------------------------------------------------------------------
struct EsaApply2Params {
  const void* x; int x_stride, x_coff;
  const void* cf; int cf_stride, cf_coff;
  const float* m3; int H3, W3, m3_stride;
  void* out; int out_stride, out_coff;
  int B, H, W, cg8;   // cg8: 8-channel groups per pixel
};
// I = index type: unsigned when the work-item count fits 32 bits (the 64-bit divisions of the index decomposition cost
// more instructions than the rest of a work item)
template <typename T, typename I = long long>
__global__ void __launch_bounds__(256, 4) k_esa_apply2(const EsaApply2Params p) {
  pdl_wait();
  const I total = (I)((long long)p.B * p.H * p.W * p.cg8);
  for (I idx = (I)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (I)gridDim.x * blockDim.x) {
    const int g = (int)(idx % (I)p.cg8);
    const I pixi = idx / (I)p.cg8;
    const long long pix = (long long)pixi;
    const int x = (int)(pixi % (I)p.W);
    const int y = (int)((pixi / (I)p.W) % (I)p.H);
    const int b = (int)(pixi / ((I)p.W * (I)p.H));
    const float sy = fmaxf(((float)y + 0.5f) * ((float)p.H3 / (float)p.H) - 0.5f, 0.f);
    const float sx = fmaxf(((float)x + 0.5f) * ((float)p.W3 / (float)p.W) - 0.5f, 0.f);
    const int y0 = min((int)sy, p.H3 - 1), x0 = min((int)sx, p.W3 - 1);
    const int y1 = min(y0 + 1, p.H3 - 1), x1 = min(x0 + 1, p.W3 - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    const float* mb = p.m3 + (long long)b * p.H3 * p.W3 * p.m3_stride + g * 8;
    float a00[8], a01[8], a10[8], a11[8], xv[8], cf[8], o[8];
    load8(mb + ((long long)y0 * p.W3 + x0) * p.m3_stride, a00);
    load8(mb + ((long long)y0 * p.W3 + x1) * p.m3_stride, a01);
    load8(mb + ((long long)y1 * p.W3 + x0) * p.m3_stride, a10);
    load8(mb + ((long long)y1 * p.W3 + x1) * p.m3_stride, a11);
    load8(reinterpret_cast<const T*>(p.x) + pix * p.x_stride + p.x_coff + g * 8, xv);
    load8(reinterpret_cast<const T*>(p.cf) + pix * p.cf_stride + p.cf_coff + g * 8, cf);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float top = a00[j] * (1.f - lx) + a01[j] * lx;
      const float bot = a10[j] * (1.f - lx) + a11[j] * lx;
      const float z = top * (1.f - ly) + bot * ly + cf[j];
      o[j] = __fdividef(xv[j], 1.f + __expf(-z));
    }
    store8(reinterpret_cast<T*>(p.out) + pix * p.out_stride + p.out_coff + g * 8, o);
  }
}

// ---------------------------------------------------------------------------------------------
// Fused ESA front (fp16 path): conv2 (3x3, stride 2, pad 0, f->f) + max_pool2d(7, 3) in one kernel.
// block = 4x4 pooled outputs = a 16x16 tile of conv2 outputs kept in shared memory; 128 threads, each
// 4 horizontally adjacent conv2 pixels x 8 channels (register tile: every broadcast LDS.128 of weights
// feeds 16 FMAs).  Tiles overlap by 4 conv2 pixels (recompute 1.78x on a 37 MMAC layer) - the price
// for never writing / re-reading the 127x127 map and for one launch instead of two.
//   in : NHWC T (c1_), 16 channels used;  w: [9][16][16] fp32 (tap, cin, cout), bias[16]
//   out: pooled map NHWC fp32, 16 channels
// ---------------------------------------------------------------------------------------------
struct EsaFrontParams {
  const void* in; int in_stride, in_coff;
  const float* w; const float* bias;
  float* out;
  int B, H, W, H2, W2, H3, W3;
};
template <typename T>
__global__ void __launch_bounds__(128) k_esa_conv2_pool(const EsaFrontParams p) {
  __shared__ __align__(16) float wsm[9 * 256];
  __shared__ __align__(16) float tile[16][16][16];   // conv2 outputs [y][x][c]
  {
    const float4* src = reinterpret_cast<const float4*>(p.w);
    float4* dst = reinterpret_cast<float4*>(wsm);
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const int j = threadIdx.x + i * 128;
      if (j < 9 * 64) dst[j] = __ldg(src + j);
    }
  }
  pdl_wait();
  __syncthreads();
  const int tx3 = (p.W3 + 3) >> 2, ty3 = (p.H3 + 3) >> 2;
  const int bx = blockIdx.x % tx3, by = (blockIdx.x / tx3) % ty3, b = blockIdx.x / (tx3 * ty3);
  const int cy0 = by * 12, cx0 = bx * 12;   // conv2 origin of this tile (pooled origin * 3)
  {
    const int g8 = threadIdx.x & 1, quad = (threadIdx.x >> 1) & 3, ty = threadIdx.x >> 3;
    const int cy = cy0 + ty, cx = cx0 + quad * 4;
    float acc[4][8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float bv = __ldg(p.bias + g8 * 8 + j);
      acc[0][j] = bv; acc[1][j] = bv; acc[2][j] = bv; acc[3][j] = bv;
    }
    const T* in = reinterpret_cast<const T*>(p.in);
    if (cy < p.H2) {
      // per kernel row: the 9 input pixels (2*cx .. 2*cx+8) the 4 outputs x 3 taps touch are fetched once, raw,
      // in one batch of 18 x 16-byte loads (3 latency rounds per thread instead of 9, 25 % fewer loads)
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
        const T* row = in + ((long long)b * p.H + 2 * cy + ky) * p.W * p.in_stride + p.in_coff;
        uint4 raw[9][2];
#pragma unroll
        for (int ip = 0; ip < 9; ++ip) {
          const int ix = 2 * cx + ip;
          if (ix < p.W) {
            const uint4* src = reinterpret_cast<const uint4*>(row + (long long)ix * p.in_stride);
            raw[ip][0] = src[0];
            raw[ip][1] = src[1];
          } else {
            raw[ip][0] = make_uint4(0, 0, 0, 0);
            raw[ip][1] = make_uint4(0, 0, 0, 0);
          }
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* wt = wsm + (ky * 3 + kx) * 256 + g8 * 8;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {          // input channels 0-7 / 8-15
            float xv[4][8];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              const __half2* h2 = reinterpret_cast<const __half2*>(&raw[2 * px + kx][hh]);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(h2[j]);
                xv[px][2 * j] = f.x; xv[px][2 * j + 1] = f.y;
              }
            }
#pragma unroll
            for (int ci = 0; ci < 8; ++ci) {
              const float4 w0 = *reinterpret_cast<const float4*>(wt + (hh * 8 + ci) * 16);
              const float4 w1 = *reinterpret_cast<const float4*>(wt + (hh * 8 + ci) * 16 + 4);
              const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
              for (int px = 0; px < 4; ++px)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[px][j] = fmaf(xv[px][ci], wv[j], acc[px][j]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      // positions outside the conv2 map never win a max
      const bool ok = cy < p.H2 && cx + px < p.W2;
      float4* d = reinterpret_cast<float4*>(&tile[ty][quad * 4 + px][g8 * 8]);
      d[0] = ok ? make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      d[1] = ok ? make_float4(acc[px][4], acc[px][5], acc[px][6], acc[px][7]) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
  }
  __syncthreads();
  // 4x4 pooled outputs x 16 channels = 256 values, 2 per thread
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int o = threadIdx.x + it * 128;
    const int c = o & 15, pxl = (o >> 4) & 3, pyl = o >> 6;
    const int py = by * 4 + pyl, pxg = bx * 4 + pxl;
    if (py < p.H3 && pxg < p.W3) {
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int j = 0; j < 7; ++j) m = fmaxf(m, tile[3 * pyl + i][3 * pxl + j][c]);
      p.out[(((long long)b * p.H3 + py) * p.W3 + pxg) * 16 + c] = m;
    }
  }
}

// The same with the register tile cut to the channels that exist: F4 = number of 4-channel groups of the ESA width
// (3 for f = 10 / 12: RFDN, RFDN40, the pruned RFDN; 4 for f = 16).  thread = 4 conv2 pixels x 4 output channels, 64 * F4
// threads: f = 12 does 9 * 12 * 12 instead of 9 * 16 * 16 multiply-adds per pixel and runs 6 warps instead of 4 (the
// kernel is FMA-issue bound at one warp per scheduler).  Padded input channels are zero and padded weights are zero, so
// skipping them changes no bit of the result; padded output channels are written as the zeros they would come out as.
template <int F4>
__global__ void __launch_bounds__(64 * F4) k_esa_conv2_pool4(const EsaFrontParams p) {
  __shared__ __align__(16) float wsm[9 * 256];
  __shared__ __align__(16) float tile[16][16][16];   // conv2 outputs [y][x][c]
  constexpr int NT = 64 * F4;
  {
    const float4* src = reinterpret_cast<const float4*>(p.w);
    float4* dst = reinterpret_cast<float4*>(wsm);
    for (int j = threadIdx.x; j < 9 * 64; j += NT) dst[j] = __ldg(src + j);
  }
  pdl_wait();
  __syncthreads();
  const int tx3 = (p.W3 + 3) >> 2, ty3 = (p.H3 + 3) >> 2;
  const int bx = blockIdx.x % tx3, by = (blockIdx.x / tx3) % ty3, b = blockIdx.x / (tx3 * ty3);
  const int cy0 = by * 12, cx0 = bx * 12;   // conv2 origin of this tile (pooled origin * 3)
  {
    const int g4 = threadIdx.x % F4, quad = (threadIdx.x / F4) & 3, ty = threadIdx.x / (4 * F4);
    const int cy = cy0 + ty, cx = cx0 + quad * 4;
    float acc[4][4];
    {
      const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias) + g4);
#pragma unroll
      for (int px = 0; px < 4; ++px) { acc[px][0] = bv.x; acc[px][1] = bv.y; acc[px][2] = bv.z; acc[px][3] = bv.w; }
    }
    const __half* in = reinterpret_cast<const __half*>(p.in);
    if (cy < p.H2) {
#pragma unroll 1
      for (int ky = 0; ky < 3; ++ky) {
        const __half* row = in + ((long long)b * p.H + 2 * cy + ky) * p.W * p.in_stride + p.in_coff;
        uint4 raw[9][2];
#pragma unroll
        for (int ip = 0; ip < 9; ++ip) {
          const int ix = 2 * cx + ip;
          if (ix < p.W) {
            const uint4* src = reinterpret_cast<const uint4*>(row + (long long)ix * p.in_stride);
            raw[ip][0] = src[0];
            if (F4 > 2) raw[ip][1] = src[1];
          } else {
            raw[ip][0] = make_uint4(0, 0, 0, 0);
            raw[ip][1] = make_uint4(0, 0, 0, 0);
          }
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* wt = wsm + (ky * 3 + kx) * 256 + g4 * 4;
#pragma unroll
          for (int c4 = 0; c4 < F4; ++c4) {          // input channels 4 c4 .. 4 c4 + 3
            float xv[4][4];
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              const __half2* h2 = reinterpret_cast<const __half2*>(&raw[2 * px + kx][c4 >> 1]) + 2 * (c4 & 1);
              const float2 f0 = __half22float2(h2[0]), f1 = __half22float2(h2[1]);
              xv[px][0] = f0.x; xv[px][1] = f0.y; xv[px][2] = f1.x; xv[px][3] = f1.y;
            }
#pragma unroll
            for (int ci = 0; ci < 4; ++ci) {
              const float4 w0 = *reinterpret_cast<const float4*>(wt + (c4 * 4 + ci) * 16);
#pragma unroll
              for (int px = 0; px < 4; ++px) {
                acc[px][0] = fmaf(xv[px][ci], w0.x, acc[px][0]);
                acc[px][1] = fmaf(xv[px][ci], w0.y, acc[px][1]);
                acc[px][2] = fmaf(xv[px][ci], w0.z, acc[px][2]);
                acc[px][3] = fmaf(xv[px][ci], w0.w, acc[px][3]);
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int px = 0; px < 4; ++px) {
      // positions outside the conv2 map never win a max
      const bool ok = cy < p.H2 && cx + px < p.W2;
      *reinterpret_cast<float4*>(&tile[ty][quad * 4 + px][g4 * 4]) =
          ok ? make_float4(acc[px][0], acc[px][1], acc[px][2], acc[px][3]) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
  }
  __syncthreads();
  // 4x4 pooled outputs x 16 channels = 256 values
  for (int o = threadIdx.x; o < 256; o += NT) {
    const int c = o & 15, pxl = (o >> 4) & 3, pyl = o >> 6;
    const int py = by * 4 + pyl, pxg = bx * 4 + pxl;
    if (py < p.H3 && pxg < p.W3) {
      float m = 0.f;                            // channels beyond the ESA width: what bias 0 + zero weights give
      if (c < 4 * F4) {
        m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 7; ++i)
#pragma unroll
          for (int j = 0; j < 7; ++j) m = fmaxf(m, tile[3 * pyl + i][3 * pxl + j][c]);
      }
      p.out[(((long long)b * p.H3 + py) * p.W3 + pxg) * 16 + c] = m;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Fused ESA low-resolution chain (fp16 path) on the pooled map: [conv_max + ReLU, conv3 + ReLU] (RFDN;
// none for RLFN) followed by the last 3x3 composed with conv4 (f -> nf, no activation) - one launch,
// intermediates in shared memory.  block = TS x TS outputs (6, or 4 at small batch); halo = number of 3x3 layers.
//   in : pooled map NHWC fp32 16 ch;  wpre: npre x [9][16][16], bpre: npre x [16];  wl: [9][16][64], bl[64]
//   out: M3 NHWC fp32 64 ch.   Zero padding applies to every layer's input: intermediate values at positions
//   outside the map are forced to 0.
// ---------------------------------------------------------------------------------------------
struct EsaChainParams {
  const float* in; float* out;
  const float* wpre0; const float* bpre0; const float* wpre1; const float* bpre1; const float* wl; const float* bl;
  int npre, B, H3, W3;
};
// TS = output tile extent (6, or 4 when there are too few 6x6 tiles to fill the GPU: batch 1 has 49 of them)
// F4 = 4-channel groups of the ESA width: input channels (and, in the 16 -> 16 layers, output channels) beyond it are
// zero with zero weights and are skipped - bit-identical, 25-44 % fewer multiply-adds for f = 10 / 12
template <int TS, int F4 = 4>
__global__ void __launch_bounds__(256) k_esa_chain(const EsaChainParams p) {
  extern __shared__ __align__(16) float sm[];
  const int npre = p.npre;
  const int halo = npre + 1;
  const int T0 = TS + 2 * halo;                // input tile extent
  float* wl = sm;                              // [9][16][64]
  float* wpre = wl + 9 * 16 * 64;              // npre x [9][16][16]
  float* buf0 = wpre + 2 * 9 * 256;            // [16][12*12]  (channel-major: conflict-free window reads)
  float* buf1 = buf0 + 12 * 12 * 16;           // [16][10*10]
  {
    // all weight loads of a thread are issued before the first shared-memory store (one L2 round trip, not nine)
    const float4* s1 = reinterpret_cast<const float4*>(p.wl);
    float4* d1 = reinterpret_cast<float4*>(wl);
    float4 t1[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) t1[i] = __ldg(s1 + threadIdx.x + i * 256);
    float4 t2[3], t3[3];
    float4* d2 = reinterpret_cast<float4*>(wpre);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int j = threadIdx.x + i * 256;
      t2[i] = (npre > 0 && j < 9 * 64) ? __ldg(reinterpret_cast<const float4*>(p.wpre0) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
      t3[i] = (npre > 1 && j < 9 * 64) ? __ldg(reinterpret_cast<const float4*>(p.wpre1) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) d1[threadIdx.x + i * 256] = t1[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int j = threadIdx.x + i * 256;
      if (j < 9 * 64) { d2[j] = t2[i]; d2[9 * 64 + j] = t3[i]; }
    }
  }
  pdl_wait();
  const int tx = (p.W3 + TS - 1) / TS, ty = (p.H3 + TS - 1) / TS;
  const int bx = blockIdx.x % tx, by = (blockIdx.x / tx) % ty, b = blockIdx.x / (tx * ty);
  const int oy0 = by * TS, ox0 = bx * TS;
  // input tile with zero padding outside the map
  for (int i = threadIdx.x; i < T0 * T0 * 4; i += 256) {
    const int q = i & 3, pos = i >> 2;
    const int yy = oy0 - halo + pos / T0, xx = ox0 - halo + pos % T0;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < p.H3 && xx >= 0 && xx < p.W3)
      v = *reinterpret_cast<const float4*>(p.in + (((long long)b * p.H3 + yy) * p.W3 + xx) * 16 + q * 4);
    const int TP0 = T0 * T0;
    buf0[(q * 4 + 0) * TP0 + pos] = v.x;
    buf0[(q * 4 + 1) * TP0 + pos] = v.y;
    buf0[(q * 4 + 2) * TP0 + pos] = v.z;
    buf0[(q * 4 + 3) * TP0 + pos] = v.w;
  }
  __syncthreads();
  float* cur = buf0;
  float* nxt = buf1;
  int Tin = T0;
  // ---- the 16 -> 16 layers with ReLU: item = (2x2 outputs, 2 channels) so that all 256 threads have work on a
  // 10x10 / 8x8 tile (the chain is latency bound at batch 1)
  for (int l = 0; l < npre; ++l) {
    const int Tout = Tin - 2;
    const int q2 = Tout >> 1;                  // 2x2 blocks per row (Tout is even)
    const float* wt = wpre + l * 9 * 256;
    const int oy_base = oy0 - halo + (l + 1), ox_base = ox0 - halo + (l + 1);
    for (int item = threadIdx.x; item < q2 * q2 * (2 * F4); item += 256) {
      const int g2 = item % (2 * F4), blk = item / (2 * F4);
      const int y2 = (blk / q2) * 2, x2 = (blk % q2) * 2;
      float acc[4][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float bv = __ldg((l == 0 ? p.bpre0 : p.bpre1) + g2 * 2 + j);
        acc[0][j] = bv; acc[1][j] = bv; acc[2][j] = bv; acc[3][j] = bv;
      }
#pragma unroll 4
      for (int ci = 0; ci < 4 * F4; ++ci) {
        float win[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) win[a][c] = cur[ci * (Tin * Tin) + (y2 + a) * Tin + x2 + c];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float2 w2 = *reinterpret_cast<const float2*>(wt + ((ky * 3 + kx) * 16 + ci) * 16 + g2 * 2);
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
              for (int px = 0; px < 2; ++px) {
                acc[py * 2 + px][0] = fmaf(win[py + ky][px + kx], w2.x, acc[py * 2 + px][0]);
                acc[py * 2 + px][1] = fmaf(win[py + ky][px + kx], w2.y, acc[py * 2 + px][1]);
              }
          }
      }
#pragma unroll
      for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int px = 0; px < 2; ++px) {
          const int gy = oy_base + y2 + py, gx = ox_base + x2 + px;
          const bool inside = gy >= 0 && gy < p.H3 && gx >= 0 && gx < p.W3;
          const int pos = (y2 + py) * Tout + x2 + px;
#pragma unroll
          for (int j = 0; j < 2; ++j) nxt[(g2 * 2 + j) * (Tout * Tout) + pos] = inside ? fmaxf(acc[py * 2 + px][j], 0.f) : 0.f;
        }
    }
    __syncthreads();
    float* t = cur; cur = nxt; nxt = t;
    Tin = Tout;
  }
  // ---- last layer 16 -> 64 (composed with conv4), TS x TS outputs: item = (2x2 outputs, 2 channels)
  {
    constexpr int QL = TS / 2;
    for (int item = threadIdx.x; item < QL * QL * 32; item += 256) {
      const int g2 = item & 31, blk = item >> 5;
      const int y2 = (blk / QL) * 2, x2 = (blk % QL) * 2;
      float acc[4][2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float bv = __ldg(p.bl + g2 * 2 + j);
        acc[0][j] = bv; acc[1][j] = bv; acc[2][j] = bv; acc[3][j] = bv;
      }
#pragma unroll 4
      for (int ci = 0; ci < 4 * F4; ++ci) {
        float win[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int c = 0; c < 4; ++c) win[a][c] = cur[ci * (Tin * Tin) + (y2 + a) * Tin + x2 + c];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            const float2 w2 = *reinterpret_cast<const float2*>(wl + ((ky * 3 + kx) * 16 + ci) * 64 + g2 * 2);
#pragma unroll
            for (int py = 0; py < 2; ++py)
#pragma unroll
              for (int px = 0; px < 2; ++px) {
                acc[py * 2 + px][0] = fmaf(win[py + ky][px + kx], w2.x, acc[py * 2 + px][0]);
                acc[py * 2 + px][1] = fmaf(win[py + ky][px + kx], w2.y, acc[py * 2 + px][1]);
              }
          }
      }
#pragma unroll
      for (int py = 0; py < 2; ++py)
#pragma unroll
        for (int px = 0; px < 2; ++px) {
          const int gy = oy0 + y2 + py, gx = ox0 + x2 + px;
          if (gy < p.H3 && gx < p.W3)
            *reinterpret_cast<float2*>(p.out + (((long long)b * p.H3 + gy) * p.W3 + gx) * 64 + g2 * 2) =
                make_float2(acc[py * 2 + px][0], acc[py * 2 + px][1]);
        }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// tensor2uint of the reference (utils/utils_image.py:204-208) on the device: planar NCHW output of the network
// -> interleaved HWC uint8: clamp to [0, data_range], * 255 / data_range in fp32, round half to even (np.round).
// thread = one output pixel (three coalesced planar reads, three adjacent byte writes).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) k_tensor2uint(const T* __restrict__ in, uint8_t* __restrict__ out, int B, int Ho,
                                                     int Wo, float data_range) {
  pdl_wait();
  const long long plane = (long long)Ho * Wo, total = (long long)B * plane;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / plane, r = i - b * plane;
    const T* src = in + b * 3 * plane + r;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = fminf(fmaxf(ldf(src + c * plane), 0.f), data_range);
      v = __fdiv_rn(__fmul_rn(v, 255.0f), data_range);
      out[i * 3 + c] = (uint8_t)rintf(v);
    }
  }
}

}  // namespace esr
