// CTC head + greedy decode (gigaam/decoder.py:7-21, gigaam/decoding.py:56-96).
//   (1) head + argmax: labels[b,t] = argmax_c (W[c,:] . enc[b,t,:] + bias[c]) in fp32 - log_softmax is
//       argmax-invariant so it is never computed.  First maximal index wins (torch.argmax).
//   (2) collapse: keep (l != blank) && (t == 0 || l != l_{t-1}) && (t < len); one warp per utterance,
//       ballot + popc prefix compaction, results resident on device:
//       ids[B,T], frames[B,T], counts[B] (int32).
#include "kernels.h"

namespace gam {
namespace {

constexpr int kRowsPerBlock = 128;
constexpr int kKC = 64;       // K chunk
constexpr int kCT = 36;       // class tile held in registers

// enc: [R, D] fp32 row-major.  W: [V1, D], bias [V1].  labels: [R] int32.
__global__ void __launch_bounds__(kRowsPerBlock) ctc_argmax_kernel(const float* __restrict__ enc, const float* __restrict__ W,
                                                                   const float* __restrict__ bias, int* __restrict__ labels,
                                                                   int R, int D, int V1) {
  __shared__ float e_s[kKC][kRowsPerBlock + 1];
  __shared__ float w_s[kCT][kKC];
  const int row0 = blockIdx.x * kRowsPerBlock;
  const int row = row0 + threadIdx.x;
  float best = -INFINITY;
  int best_i = 0;
  for (int c0 = 0; c0 < V1; c0 += kCT) {
    float acc[kCT];
#pragma unroll
    for (int c = 0; c < kCT; ++c) acc[c] = 0.f;
    for (int k0 = 0; k0 < D; k0 += kKC) {
      __syncthreads();
      // enc tile: coalesced float4 reads along K, transposed into e_s[k][row]
      for (int i = threadIdx.x; i < kRowsPerBlock * (kKC / 4); i += kRowsPerBlock) {
        const int rr = i / (kKC / 4), k4 = (i % (kKC / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + rr < R) v = *reinterpret_cast<const float4*>(enc + static_cast<size_t>(row0 + rr) * D + k0 + k4);
        e_s[k4 + 0][rr] = v.x;
        e_s[k4 + 1][rr] = v.y;
        e_s[k4 + 2][rr] = v.z;
        e_s[k4 + 3][rr] = v.w;
      }
      for (int i = threadIdx.x; i < kCT * kKC; i += kRowsPerBlock) {
        const int c = i / kKC, k = i % kKC;
        w_s[c][k] = (c0 + c < V1) ? __ldg(W + static_cast<size_t>(c0 + c) * D + k0 + k) : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int k = 0; k < kKC; ++k) {
        const float x = e_s[k][threadIdx.x];
#pragma unroll
        for (int c = 0; c < kCT; ++c) acc[c] = fmaf(w_s[c][k], x, acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < kCT; ++c) {
      if (c0 + c < V1) {
        const float v = acc[c] + __ldg(bias + c0 + c);
        if (v > best) { best = v; best_i = c0 + c; }
      }
    }
  }
  if (row < R) labels[row] = best_i;
}

__global__ void __launch_bounds__(128) ctc_collapse_kernel(const int* __restrict__ labels, const int* __restrict__ len, int B,
                                                           int T, int blank, int* __restrict__ ids, int* __restrict__ frames,
                                                           int* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (b >= B) return;
  const int L = min(max(len[b], 0), T);
  const int* lab = labels + static_cast<size_t>(b) * T;
  int base = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    int l = blank, prev = -1;
    if (t < T) {
      l = lab[t];
      prev = t > 0 ? lab[t - 1] : -1;
    }
    const bool keep = (t < L) && (l != blank) && (t == 0 || l != prev);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int pos = base + __popc(mask & ((1u << lane) - 1u));
      ids[static_cast<size_t>(b) * T + pos] = l;
      frames[static_cast<size_t>(b) * T + pos] = t;
    }
    base += __popc(mask);
  }
  if (lane == 0) counts[b] = base;
}

}  // namespace

void launch_ctc_argmax(const float* enc, const float* W, const float* bias, int* labels, int R, int D, int V1,
                       cudaStream_t s) {
  ctc_argmax_kernel<<<(R + kRowsPerBlock - 1) / kRowsPerBlock, kRowsPerBlock, 0, s>>>(enc, W, bias, labels, R, D, V1);
}
void launch_ctc_collapse(const int* labels, const int* len, int B, int T, int blank, int* ids, int* frames, int* counts,
                         cudaStream_t s) {
  ctc_collapse_kernel<<<(B + 3) / 4, 128, 0, s>>>(labels, len, B, T, blank, ids, frames, counts);
}

}  // namespace gam
