// Row-wise kernels of the Conformer block (HBM / L2 bound):
//   * LayerNorm(768) fp32 -> fp16 GEMM operand                       (gigaam/encoder.py:447-471,481-497)
//   * LayerNorm + rotary embedding -> (u, rope(u)) fp16 operands      (gigaam/encoder.py:245-250, utils.py:83-100)
//   * norm_out LayerNorm fused with the next layer's first LayerNorm   (gigaam/encoder.py:497 -> :481)
//   * masked depthwise conv (k taps) + folded eval-BatchNorm + SiLU    (gigaam/encoder.py:400-407)
//   * masked depthwise conv + LayerNorm-over-channels + SiLU (v3 shape)
//   * stage-length recursion of the striding subsampling               (gigaam/encoder.py:77-90) + the packed-row plan
// One warp owns one row of D=768 floats (24 per lane, 6 x float4, fully coalesced).
//
// Rows are PACKED (varlen): utterance b owns rows cu[b] .. cu[b] + plen[b] and the row count lives on the device
// (`rows_dev`); grids are sized for the padded maximum and the surplus warps leave at once.
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
// rows that exist: the device-side count when there is one (never more than the host's maximum)
__device__ __forceinline__ int live_rows(const int* rows_dev, int rows_max) {
  return rows_dev != nullptr ? min(max(__ldg(rows_dev), 0), rows_max) : rows_max;
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
                                                     int rows_max, const int* __restrict__ rows_dev, int reverse, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = ln_row(blockIdx.x, threadIdx.x >> 5, live_rows(rows_dev, rows_max), reverse);
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
                                                          __half* __restrict__ out_r, int rows_max,
                                                          const int* __restrict__ rows_dev, const int* __restrict__ row_t,
                                                          int T, int half_dim, int reverse, float eps) {
  __shared__ float srow[8][kD];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int row = ln_row(blockIdx.x, w, live_rows(rows_dev, rows_max), reverse);
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
  const int t = row_t != nullptr ? __ldg(row_t + row) : row % T;   // frame index inside its utterance
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
                                                        __half* __restrict__ y_out, int rows_max,
                                                        const int* __restrict__ rows_dev, int reverse, float eps) {
  const int lane = threadIdx.x & 31;
  const int row = ln_row(blockIdx.x, threadIdx.x >> 5, live_rows(rows_dev, rows_max), reverse);
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

// ------------------------------------------------------------------ packed rows -> the caller's padded [B, T, 768] fp32
// out[b, t] = LN(x[cu[b] + t]) (gamma != null: the last layer's norm_out, gigaam/encoder.py:497) or x[cu[b] + t] itself
// (pre_encode output); frames t >= plen[b] do not exist in the packed stream and are written as zeros.
__global__ void __launch_bounds__(256) unpack_rows_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const int* __restrict__ cu,
                                                          const int* __restrict__ plen, float* __restrict__ out, int B, int T,
                                                          int reverse, float eps) {
  const int lane = threadIdx.x & 31;
  const int orow = ln_row(blockIdx.x, threadIdx.x >> 5, B * T, reverse);
  if (orow < 0) return;
  const int b = orow / T, t = orow - b * T;
  float4* o = reinterpret_cast<float4*>(out + static_cast<size_t>(orow) * kD);
  if (t >= min(__ldg(plen + b), T)) {
#pragma unroll
    for (int i = 0; i < kVec; ++i) o[lane + 32 * i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const float* src = x + (static_cast<size_t>(__ldg(cu + b)) + t) * kD;
  float4 v[kVec];
  if (gamma != nullptr) {
    float mean, rstd;
    load_row_stats(src, lane, v, mean, rstd, eps);
    const float4* g4 = reinterpret_cast<const float4*>(gamma);
    const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
    for (int i = 0; i < kVec; ++i) o[lane + 32 * i] = ln_apply(v[i], mean, rstd, g4[lane + 32 * i], b4[lane + 32 * i]);
  } else {
#pragma unroll
    for (int i = 0; i < kVec; ++i) o[lane + 32 * i] = reinterpret_cast<const float4*>(src)[lane + 32 * i];
  }
}

// ------------------------------------------------------------------ depthwise conv (+ folded BN) + SiLU
// g: packed rows x 768 fp16 (GLU output, NOT yet pad-masked: masked here on load, gigaam/encoder.py:400-401); utterance b
// starts at row cu[b] and owns plen[b] rows (cu == null: row b*T, T rows); frames t >= len[b] read as zero (len < plen only
// for a batch of one, whose padded frames stay in the stream because the reference attends to them).
// w: [768, KW] fp32, b: [768] fp32 with eval BatchNorm folded in.  out = silu(conv) fp16.
// block = (channel tile of 128, time tile of 64, b); thread = 2 channels x 16 time steps.
constexpr int kDwTT = 32;          // time steps per block
constexpr int kDwCT = 128;         // channels per block
constexpr int kDwPerThread = 8;    // consecutive outputs per thread (x 2 channels)

template <int KW>
__global__ void __launch_bounds__(256) dwconv_bn_silu_kernel(const __half* __restrict__ g, const float* __restrict__ w,
                                                             const float* __restrict__ bias, const int* __restrict__ len,
                                                             const int* __restrict__ cu, const int* __restrict__ plen,
                                                             __half* __restrict__ out, int T) {
  constexpr int kHalo = (KW - 1) / 2;
  constexpr int kRows = kDwTT + KW - 1;
  __shared__ __align__(16) __half2 tile[kRows][kDwCT / 2];
  const int c0 = blockIdx.x * kDwCT;
  const int t0 = blockIdx.y * kDwTT;
  const int b = blockIdx.z;
  const int L = min(len[b], T);
  const int P = cu != nullptr ? min(__ldg(plen + b), T) : T;   // frames of this utterance that exist
  if (t0 >= P) return;
  const size_t base = cu != nullptr ? static_cast<size_t>(__ldg(cu + b)) : static_cast<size_t>(b) * T;
  const __half* gb = g + base * kD;
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
  __half* ob = out + base * kD;
#pragma unroll
  for (int o = 0; o < kDwPerThread; ++o) {
    const int t = t0 + rbase + o;
    if (t < P) {
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
                                                             const int* __restrict__ cu, const int* __restrict__ row_b,
                                                             const int* __restrict__ row_t, const int* __restrict__ rows_dev,
                                                             __half* __restrict__ out, int T, int rows_max, float eps) {
  constexpr int kHalo = (KW - 1) / 2;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= live_rows(rows_dev, rows_max)) return;
  const int b = cu != nullptr ? __ldg(row_b + row) : row / T;
  const int t = cu != nullptr ? __ldg(row_t + row) : row % T;
  const size_t base = cu != nullptr ? static_cast<size_t>(__ldg(cu + b)) : static_cast<size_t>(b) * T;
  const int L = min(len[b], T);
  float acc[24];
#pragma unroll
  for (int i = 0; i < 24; ++i) acc[i] = __ldg(bias + lane * 24 + i);
  for (int k = 0; k < KW; ++k) {
    const int tt = t + k - kHalo;
    if (tt < 0 || tt >= L) continue;
    const __half2* src = reinterpret_cast<const __half2*>(g + (base + tt) * kD + lane * 24);
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

// ------------------------------------------------------------------ subsampling length recursion + packed-row plan
// len_k = floor((len_{k-1} + 2p - k) / 2 + 1), computed in float like the reference (encoder.py:86-90).
// plen[b] = frames of utterance b kept in the packed stream = len2[b]; a batch of ONE keeps all T2 frames because the
// reference builds no attention mask for it (encoder.py:621-625) and its padded frames are attended to.
// cu = exclusive prefix sum of plen ([B + 1]; cu[B] = rows_dev[0] = packed row count); run1[b] = how many stage-1 frames
// of the conv2d subsampling the stage-2 conv of the kept frames can touch (stage 1 writes exactly those).
// One block; B is walked in chunks of blockDim.x with a carried prefix.
__global__ void __launch_bounds__(1024) pack_plan_kernel(const long long* __restrict__ mel_len, int B, int pad2_minus_k, int max_T0,
                                                         int T1, int T2, int* __restrict__ len0, int* __restrict__ len1,
                                                         int* __restrict__ len2, int* __restrict__ plen, int* __restrict__ run1,
                                                         int* __restrict__ cu, int* __restrict__ rows_dev) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int b0 = 0; b0 < B; b0 += blockDim.x) {
    const int b = b0 + threadIdx.x;
    int pl = 0;
    if (b < B) {
      float l = static_cast<float>(mel_len[b]);
      len0[b] = static_cast<int>(min(mel_len[b], static_cast<long long>(max_T0)));
      l = floorf((l + pad2_minus_k) / 2.0f + 1.0f);
      const int l1 = static_cast<int>(l);
      len1[b] = l1;
      l = floorf((l + pad2_minus_k) / 2.0f + 1.0f);
      const int l2 = static_cast<int>(l);
      len2[b] = l2;
      pl = B > 1 ? min(max(l2, 0), T2) : T2;
      plen[b] = pl;
      // conv2d: a kept stage-2 frame t2 reads stage-1 frames 2 t2 - 1 .. 2 t2 + 1 and stage 2 works in blocks of 8 frames, so
      // every stage-1 frame below 2 * roundup8(pl) must be defined (zeros past len1 included); nothing above it is read
      // by a block that stores anything
      run1[b] = B > 1 ? min(T1, 2 * ((pl + 7) / 8 * 8) + 2) : T1;
    }
    int v = pl;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += n;
    }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    if (w == 0) {
      int tv = lane < nw ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, tv, o);
        if (lane >= o) tv += n;
      }
      warp_tot[lane] = tv;   // inclusive totals of the warps
    }
    __syncthreads();
    const int carry = carry_s;
    const int incl = carry + v + (w > 0 ? warp_tot[w - 1] : 0);
    if (b < B) cu[b] = incl - pl;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    cu[B] = carry_s;
    rows_dev[0] = carry_s;
  }
}

// row -> (utterance, frame) of the packed stream, for the kernels that walk rows but need the position (RoPE, conv halo)
__global__ void row_map_kernel(const int* __restrict__ cu, const int* __restrict__ plen, int T, int* __restrict__ row_b,
                               int* __restrict__ row_t) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < min(plen[b], T)) {
    const int r = cu[b] + t;
    row_b[r] = b;
    row_t[r] = t;
  }
}

}  // namespace

// ------------------------------------------------------------------ launchers
void launch_ln_f16(const float* x, const float* g, const float* b, __half* out, int rows, const int* rows_dev, int reverse,
                   cudaStream_t s) {
  launch_k(ln_f16_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, x, g, b, out, rows, rows_dev, reverse, 1e-5f);
}
void launch_ln_rope_f16(const float* x, const float* g, const float* b, const float* rc, const float* rs, __half* out_u,
                        __half* out_r, int rows, const int* rows_dev, const int* row_t, int T, int half_dim, int reverse,
                        cudaStream_t s) {
  launch_k(ln_rope_f16_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, x, g, b, rc, rs, out_u, out_r, rows, rows_dev, row_t, T,
           half_dim, reverse, 1e-5f);
}
void launch_ln_out_ln(const float* r, const float* g_out, const float* b_out, const float* g_next, const float* b_next,
                      float* x_out, __half* y_out, int rows, const int* rows_dev, int reverse, cudaStream_t s) {
  launch_k(ln_out_ln_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, r, g_out, b_out, g_next, b_next, x_out, y_out, rows, rows_dev,
           reverse, 1e-5f);
}
void launch_unpack_rows(const float* x, const float* gamma, const float* beta, const int* cu, const int* plen, float* out, int B,
                        int T, int reverse, cudaStream_t s) {
  launch_k(unpack_rows_kernel, dim3((B * T + 7) / 8), dim3(256), 0, s, x, gamma, beta, cu, plen, out, B, T, reverse, 1e-5f);
}
int launch_dwconv_bn_silu(const __half* g, const float* w, const float* bias, const int* len, const int* cu, const int* plen,
                          __half* out, int B, int T, int kw, cudaStream_t s) {
  dim3 grid(kD / kDwCT, (T + kDwTT - 1) / kDwTT, B);
  if (kw == 31) launch_k(dwconv_bn_silu_kernel<31>, grid, dim3(256), 0, s, g, w, bias, len, cu, plen, out, T);
  else if (kw == 5) launch_k(dwconv_bn_silu_kernel<5>, grid, dim3(256), 0, s, g, w, bias, len, cu, plen, out, T);
  else return -1;
  return 0;
}
int launch_dwconv_ln_silu(const __half* g, const float* w, const float* bias, const float* gamma, const float* beta,
                          const int* len, const int* cu, const int* row_b, const int* row_t, const int* rows_dev, __half* out,
                          int B, int T, int kw, cudaStream_t s) {
  const int rows = B * T;
  if (kw == 5)
    dwconv_ln_silu_kernel<5><<<(rows + 7) / 8, 256, 0, s>>>(g, w, bias, gamma, beta, len, cu, row_b, row_t, rows_dev, out, T, rows, 1e-5f);
  else if (kw == 31)
    dwconv_ln_silu_kernel<31><<<(rows + 7) / 8, 256, 0, s>>>(g, w, bias, gamma, beta, len, cu, row_b, row_t, rows_dev, out, T, rows, 1e-5f);
  else return -1;
  return 0;
}
void launch_pack_plan(const long long* mel_len, int B, int pad2_minus_k, int max_T0, int T1, int T2, int* len0, int* len1,
                      int* len2, int* plen, int* run1, int* cu, int* rows_dev, int* row_b, int* row_t, cudaStream_t s) {
  const int threads = B >= 1024 ? 1024 : ((B + 31) / 32 * 32);
  pack_plan_kernel<<<1, threads, 0, s>>>(mel_len, B, pad2_minus_k, max_T0, T1, T2, len0, len1, len2, plen, run1, cu, rows_dev);
  row_map_kernel<<<dim3((T2 + 255) / 256, B), 256, 0, s>>>(cu, plen, T2, row_b, row_t);
}

}  // namespace gam
