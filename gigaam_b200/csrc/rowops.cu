// Row-wise kernels of the Conformer block (HBM / L2 bound):
//   * LayerNorm(768) fp32 -> fp16 GEMM operand                       (gigaam/encoder.py:447-471,481-497)
//   * LayerNorm + rotary embedding -> (u, rope(u)) fp16 operands      (gigaam/encoder.py:245-250, utils.py:83-100)
//   * norm_out LayerNorm fused with the next layer's first LayerNorm   (gigaam/encoder.py:497 -> :481)
//   * masked depthwise conv (k taps) + folded eval-BatchNorm + SiLU    (gigaam/encoder.py:400-407)
//   * masked depthwise conv + LayerNorm-over-channels + SiLU (v3 shape)
//   * stage-length recursion of the striding subsampling               (gigaam/encoder.py:77-90)
// One warp owns one row of D=768 floats (24 per lane, 6 x float4, fully coalesced).
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace gam {

namespace {

constexpr int kD = 768;
constexpr int kVec = kD / 128;  // float4 per lane

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// loads a 768-float row into registers (lane-strided float4) and returns (mean, rstd)
__device__ __forceinline__ void load_row_stats(const float* __restrict__ row, int lane, float4 (&v)[kVec],
                                               float& mean, float& rstd, float eps) {
  const float4* r4 = reinterpret_cast<const float4*>(row);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    v[i] = r4[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + eps);
}

// Row handled by warp `w` of block `blk`; `reverse` = the LAST rows first.  The residual GEMM that has just written x
// streamed ~120 MB through the 126 MB L2, so the rows it wrote first are evicted and the rows it wrote last are still
// resident: a LayerNorm that walks the rows in the OPPOSITE direction turns an LRU-pathological re-read (0 % hits) into
// hits on everything still cached (gam_api.cu alternates the direction kernel by kernel).
__device__ __forceinline__ int ln_row(int blk, int w, int rows, int reverse) {
  const int r = blk * 8 + w;
  return r >= rows ? -1 : (reverse ? rows - 1 - r : r);
}

__device__ __forceinline__ float4 ln_apply(float4 v, float mean, float rstd, float4 g, float4 b) {
  return make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y,
                     (v.z - mean) * rstd * g.z + b.z, (v.w - mean) * rstd * g.w + b.w);
}

__device__ __forceinline__ uint2 pack4(float4 v) {
  __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}

// ------------------------------------------------------------------ LN -> fp16
__global__ void __launch_bounds__(256) ln_f16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, __half* __restrict__ out,
                                                     int rows, int reverse, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = ln_row(blockIdx.x, threadIdx.x >> 5, rows, reverse);
  if (row < 0) return;
  float4 v[kVec];
  float mean, rstd;
  load_row_stats(x + static_cast<size_t>(row) * kD, lane, v, mean, rstd, eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* o = reinterpret_cast<uint2*>(out + static_cast<size_t>(row) * kD);
#pragma unroll
  for (int i = 0; i < kVec; ++i) o[lane + 32 * i] = pack4(ln_apply(v[i], mean, rstd, g4[lane + 32 * i], b4[lane + 32 * i]));
}

// ------------------------------------------------------------------ LN -> (u, rope(u)) fp16
// rope(u)[h*48+i]    = u[i]*cos[t,i] - u[i+24]*sin[t,i]          (i < 24)
// rope(u)[h*48+24+i] = u[i+24]*cos[t,i] + u[i]*sin[t,i]
// cos/sin tables: [max_len, 24] fp32, theta = t / base^(2i/48)   (gigaam/encoder.py:337-355)
__global__ void __launch_bounds__(256) ln_rope_f16_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ rope_cos,
                                                          const float* __restrict__ rope_sin, __half* __restrict__ out_u,
                                                          __half* __restrict__ out_r, int rows, int T, int half_dim,
                                                          int reverse, float eps) {
  __shared__ float srow[8][kD];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int row = ln_row(blockIdx.x, w, rows, reverse);
  if (row < 0) return;
  float4 v[kVec];
  float mean, rstd;
  load_row_stats(x + static_cast<size_t>(row) * kD, lane, v, mean, rstd, eps);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
  uint2* ou = reinterpret_cast<uint2*>(out_u + static_cast<size_t>(row) * kD);
  float4* s4 = reinterpret_cast<float4*>(srow[w]);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    float4 y = ln_apply(v[i], mean, rstd, g4[lane + 32 * i], b4[lane + 32 * i]);
    ou[lane + 32 * i] = pack4(y);
    s4[lane + 32 * i] = y;
  }
  __syncwarp();
  const int t = row % T;
  const float* cs = rope_cos + static_cast<size_t>(t) * half_dim;
  const float* sn = rope_sin + static_cast<size_t>(t) * half_dim;
  // head_dim and half_dim are multiples of 4, so a float4 never straddles the rotation boundary: the partner of
  // a float4 is the float4 half_dim/4 positions away (conflict-free 16-byte smem reads, aligned table reads)
  const int hq = half_dim >> 2;            // float4 per half head
  uint2* orr = reinterpret_cast<uint2*>(out_r + static_cast<size_t>(row) * kD);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const int f4 = lane + 32 * i;          // float4 index in the row
    const int q = f4 % (2 * hq);           // position inside the head
    const bool lo = q < hq;
    const float4 a = s4[f4];
    const float4 pt = s4[lo ? f4 + hq : f4 - hq];
    const float sg = lo ? -1.f : 1.f;
    const float4 c = __ldg(reinterpret_cast<const float4*>(cs) + (lo ? q : q - hq));
    const float4 s = __ldg(reinterpret_cast<const float4*>(sn) + (lo ? q : q - hq));
    orr[f4] = pack4(make_float4(fmaf(sg * pt.x, s.x, a.x * c.x), fmaf(sg * pt.y, s.y, a.y * c.y),
                                fmaf(sg * pt.z, s.z, a.z * c.z), fmaf(sg * pt.w, s.w, a.w * c.w)));
  }
}

// ------------------------------------------------------------------ x = LN_out(r) (fp32, may alias r);  y = LN_next(x) fp16
__global__ void __launch_bounds__(256) ln_out_ln_kernel(const float* __restrict__ r, const float* __restrict__ g_out,
                                                        const float* __restrict__ b_out, const float* __restrict__ g_next,
                                                        const float* __restrict__ b_next, float* __restrict__ x_out,
                                                        __half* __restrict__ y_out, int rows, int reverse, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = ln_row(blockIdx.x, threadIdx.x >> 5, rows, reverse);
  if (row < 0) return;
  float4 v[kVec];
  float mean, rstd;
  load_row_stats(r + static_cast<size_t>(row) * kD, lane, v, mean, rstd, eps);
  const float4* g4 = reinterpret_cast<const float4*>(g_out);
  const float4* b4 = reinterpret_cast<const float4*>(b_out);
  float4* xo = reinterpret_cast<float4*>(x_out + static_cast<size_t>(row) * kD);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    v[i] = ln_apply(v[i], mean, rstd, g4[lane + 32 * i], b4[lane + 32 * i]);
    xo[lane + 32 * i] = v[i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  if (y_out == nullptr) return;
  mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + eps);
  const float4* gn = reinterpret_cast<const float4*>(g_next);
  const float4* bn = reinterpret_cast<const float4*>(b_next);
  uint2* yo = reinterpret_cast<uint2*>(y_out + static_cast<size_t>(row) * kD);
#pragma unroll
  for (int i = 0; i < kVec; ++i) yo[lane + 32 * i] = pack4(ln_apply(v[i], mean, rstd, gn[lane + 32 * i], bn[lane + 32 * i]));
}

// ------------------------------------------------------------------ depthwise conv (+ folded BN) + SiLU
// g: [B, T, 768] fp16 (GLU output, NOT yet pad-masked: masked here on load, gigaam/encoder.py:400-401)
// w: [768, KW] fp32, b: [768] fp32 with eval BatchNorm folded in.  out = silu(conv) fp16.
// block = (channel tile of 128, time tile of 64, b); thread = 2 channels x 16 time steps.
constexpr int kDwTT = 32;          // time steps per block
constexpr int kDwCT = 128;         // channels per block
constexpr int kDwPerThread = 8;    // consecutive outputs per thread (x 2 channels)

template <int KW>
__global__ void __launch_bounds__(256) dwconv_bn_silu_kernel(const __half* __restrict__ g, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const int* __restrict__ len,
                                                             __half* __restrict__ out, int T) {
  constexpr int kHalo = (KW - 1) / 2;
  constexpr int kRows = kDwTT + KW - 1;
  __shared__ __align__(16) __half2 tile[kRows][kDwCT / 2];
  const int c0 = blockIdx.x * kDwCT;
  const int t0 = blockIdx.y * kDwTT;
  const int b = blockIdx.z;
  const int L = min(len[b], T);
  const __half* gb = g + static_cast<size_t>(b) * T * kD;
  // input tile: 16-byte loads, one 256-byte row segment per 16 threads; padded frames / halo -> 0
  for (int i = threadIdx.x; i < kRows * (kDwCT / 8); i += blockDim.x) {
    const int rr = i / (kDwCT / 8), c8 = i % (kDwCT / 8);
    const int t = t0 + rr - kHalo;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (t >= 0 && t < L) v = *reinterpret_cast<const uint4*>(gb + static_cast<size_t>(t) * kD + c0 + 8 * c8);
    *reinterpret_cast<uint4*>(&tile[rr][4 * c8]) = v;
  }
  __syncthreads();
  const int cp = threadIdx.x % (kDwCT / 2);      // channel pair
  const int tg = threadIdx.x / (kDwCT / 2);      // time group (0..3)
  const int ch = c0 + 2 * cp;
  const int rbase = tg * kDwPerThread;
  float2 x[kDwPerThread + KW - 1];
#pragma unroll
  for (int j = 0; j < kDwPerThread + KW - 1; ++j) x[j] = __half22float2(tile[rbase + j][cp]);
  float a0[kDwPerThread], a1[kDwPerThread];
  const float2 bb = __ldg(reinterpret_cast<const float2*>(bias + ch));
#pragma unroll
  for (int o = 0; o < kDwPerThread; ++o) { a0[o] = bb.x; a1[o] = bb.y; }
  // taps are stored transposed [KW][768]: one coalesced 8-byte load per tap, L1-resident across the block
  const float2* wt = reinterpret_cast<const float2*>(w + ch);
#pragma unroll
  for (int k = 0; k < KW; ++k) {
    const float2 wk = __ldg(wt + static_cast<size_t>(k) * (kD / 2));
#pragma unroll
    for (int o = 0; o < kDwPerThread; ++o) {
      a0[o] = fmaf(wk.x, x[o + k].x, a0[o]);
      a1[o] = fmaf(wk.y, x[o + k].y, a1[o]);
    }
  }
  __half* ob = out + static_cast<size_t>(b) * T * kD;
#pragma unroll
  for (int o = 0; o < kDwPerThread; ++o) {
    const int t = t0 + rbase + o;
    if (t < T) {
      // silu(a) = 0.5 a (1 + tanh(0.5 a)): one MUFU, no division
      const float h0 = 0.5f * a0[o], h1 = 0.5f * a1[o];
      float t0v, t1v;
      asm("tanh.approx.f32 %0, %1;" : "=f"(t0v) : "f"(h0));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t1v) : "f"(h1));
      *reinterpret_cast<__half2*>(ob + static_cast<size_t>(t) * kD + ch) = __floats2half2_rn(fmaf(h0, t0v, h0), fmaf(h1, t1v, h1));
    }
  }
}

// depthwise conv + LayerNorm over channels + SiLU (conv_norm_type == "layer_norm", gigaam/encoder.py:404-406)
// one warp per (b, t) row; lane owns 24 channels.
template <int KW>
__global__ void __launch_bounds__(256) dwconv_ln_silu_kernel(const __half* __restrict__ g, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, const int* __restrict__ len,
                                                             __half* __restrict__ out, int T, int rows, float eps) {
  constexpr int kHalo = (KW - 1) / 2;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / T, t = row % T;
  const int L = min(len[b], T);
  float acc[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = __ldg(bias + lane * 24 + i);
  for (int k = 0; k < KW; ++k) {
    const int tt = t + k - kHalo;
    if (tt < 0 || tt >= L) continue;
    const __half2* src = reinterpret_cast<const __half2*>(g + (static_cast<size_t>(b) * T + tt) * kD + lane * 24);
#pragma unroll
    for (int i = 0; i < 12; ++i) {
      const float2 x = __half22float2(src[i]);
      const float2 wk = __ldg(reinterpret_cast<const float2*>(w + static_cast<size_t>(k) * kD + lane * 24 + 2 * i));
      acc[2 * i] = fmaf(wk.x, x.x, acc[2 * i]);
      acc[2 * i + 1] = fmaf(wk.y, x.y, acc[2 * i + 1]);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) s += acc[i];
  const float mean = warp_sum(s) * (1.0f / kD);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 24; ++i) { const float d = acc[i] - mean; q += d * d; }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / kD) + eps);
  __half2* dst = reinterpret_cast<__half2*>(out + static_cast<size_t>(row) * kD + lane * 24);
#pragma unroll
  for (int i = 0; i < 12; ++i) {
    float y0 = (acc[2 * i] - mean) * rstd * __ldg(gamma + lane * 24 + 2 * i) + __ldg(beta + lane * 24 + 2 * i);
    float y1 = (acc[2 * i + 1] - mean) * rstd * __ldg(gamma + lane * 24 + 2 * i + 1) + __ldg(beta + lane * 24 + 2 * i + 1);
    y0 = y0 / (1.f + __expf(-y0));
    y1 = y1 / (1.f + __expf(-y1));
    dst[i] = __floats2half2_rn(y0, y1);
  }
}

// ------------------------------------------------------------------ subsampling length recursion
// len_k = floor((len_{k-1} + 2p - k) / 2 + 1), computed in float like the reference (encoder.py:86-90)
__global__ void sub_lengths_kernel(const long long* __restrict__ mel_len, int B, int pad2_minus_k, int max_T0, int* __restrict__ len0,
                                   int* __restrict__ len1, int* __restrict__ len2) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float l = static_cast<float>(mel_len[b]);
  len0[b] = static_cast<int>(min(mel_len[b], static_cast<long long>(max_T0)));
  l = floorf((l + pad2_minus_k) / 2.0f + 1.0f);
  len1[b] = static_cast<int>(l);
  l = floorf((l + pad2_minus_k) / 2.0f + 1.0f);
  len2[b] = static_cast<int>(l);
}

}  // namespace

// ------------------------------------------------------------------ launchers
void launch_ln_f16(const float* x, const float* g, const float* b, __half* out, int rows, int reverse, cudaStream_t s) {
  launch_k(ln_f16_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, x, g, b, out, rows, reverse, 1e-5f);
}
void launch_ln_rope_f16(const float* x, const float* g, const float* b, const float* rc, const float* rs, __half* out_u,
                        __half* out_r, int rows, int T, int half_dim, int reverse, cudaStream_t s) {
  launch_k(ln_rope_f16_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, x, g, b, rc, rs, out_u, out_r, rows, T, half_dim, reverse, 1e-5f);
}
void launch_ln_out_ln(const float* r, const float* g_out, const float* b_out, const float* g_next, const float* b_next,
                      float* x_out, __half* y_out, int rows, int reverse, cudaStream_t s) {
  launch_k(ln_out_ln_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, r, g_out, b_out, g_next, b_next, x_out, y_out, rows, reverse, 1e-5f);
}
int launch_dwconv_bn_silu(const __half* g, const float* w, const float* bias, const int* len, __half* out, int B, int T,
                          int kw, cudaStream_t s) {
  dim3 grid(kD / kDwCT, (T + kDwTT - 1) / kDwTT, B);
  if (kw == 31) launch_k(dwconv_bn_silu_kernel<31>, grid, dim3(256), 0, s, g, w, bias, len, out, T);
  else if (kw == 5) launch_k(dwconv_bn_silu_kernel<5>, grid, dim3(256), 0, s, g, w, bias, len, out, T);
  else return -1;
  return 0;
}
int launch_dwconv_ln_silu(const __half* g, const float* w, const float* bias, const float* gamma, const float* beta,
                          const int* len, __half* out, int B, int T, int kw, cudaStream_t s) {
  const int rows = B * T;
  if (kw == 5) dwconv_ln_silu_kernel<5><<<(rows + 7) / 8, 256, 0, s>>>(g, w, bias, gamma, beta, len, out, T, rows, 1e-5f);
  else if (kw == 31) dwconv_ln_silu_kernel<31><<<(rows + 7) / 8, 256, 0, s>>>(g, w, bias, gamma, beta, len, out, T, rows, 1e-5f);
  else return -1;
  return 0;
}
void launch_sub_lengths(const long long* mel_len, int B, int pad2_minus_k, int max_T0, int* len0, int* len1, int* len2,
                        cudaStream_t s) {
  sub_lengths_kernel<<<(B + 127) / 128, 128, 0, s>>>(mel_len, B, pad2_minus_k, max_T0, len0, len1, len2);
}

}  // namespace gam
