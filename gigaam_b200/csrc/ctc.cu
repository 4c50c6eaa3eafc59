// CTC head + greedy decode (gigaam/decoder.py:7-21, gigaam/decoding.py:56-96).
//   (1) head + argmax: labels[b,t] = argmax_c (W[c,:] . enc[b,t,:] + bias[c]) in fp32 - log_softmax is
//       argmax-invariant so it is never computed.  First maximal index wins (torch.argmax).
//   (2) collapse: keep (l != blank) && (t == 0 || l != l_{t-1}) && (t < len); one warp per utterance,
//       ballot + popc prefix compaction, results resident on device:
//       ids[B,T], frames[B,T], counts[B] (int32).
#include "kernels.h"

namespace gam {
namespace {

constexpr int kRows = 32;      // rows (frames) per block: thread = (row, class group)
constexpr int kGroups = 4;     // class groups = warps per block
constexpr int kKC = 64;        // K chunk
constexpr int kCG = 9;         // classes per thread and class tile
constexpr int kCT = kGroups * kCG;   // class tile (36)

// enc: [R, D] fp32 row-major.  W: [V1, D], bias [V1].  labels: [R] int32.
// Block = 32 frames x 4 class groups (one warp per group: its W reads are broadcasts, its enc reads conflict-free):
// 502 blocks at the benchmark shape instead of 126 single-warp-per-SM blocks (ncu r2f: 98 us at 6 % warps active).
// Every (frame, class) sum still runs over k in ascending order, so the logits -- and the argmax -- are bit-identical.
__global__ void __launch_bounds__(kRows * kGroups) ctc_argmax_kernel(const float* __restrict__ enc, const float* __restrict__ W,
                                                                     const float* __restrict__ bias, int* __restrict__ labels,
                                                                     int R, int D, int V1) {
  __shared__ float e_s[kKC][kRows + 1];
  __shared__ float w_s[kCT][kKC];
  __shared__ float best_v[kGroups][kRows];
  __shared__ int best_c[kGroups][kRows];
  const int r = threadIdx.x & 31, cg = threadIdx.x >> 5;
  const int row0 = blockIdx.x * kRows;
  float best = -INFINITY;
  int best_i = 0;
  for (int c0 = 0; c0 < V1; c0 += kCT) {
    float acc[kCG];
#pragma unroll
    for (int c = 0; c < kCG; ++c) acc[c] = 0.f;
    for (int k0 = 0; k0 < D; k0 += kKC) {
      __syncthreads();
      // enc tile: coalesced float4 reads along K, transposed into e_s[k][row]
      for (int i = threadIdx.x; i < kRows * (kKC / 4); i += kRows * kGroups) {
        const int rr = i / (kKC / 4), k4 = (i % (kKC / 4)) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row0 + rr < R) v = *reinterpret_cast<const float4*>(enc + static_cast<size_t>(row0 + rr) * D + k0 + k4);
        e_s[k4 + 0][rr] = v.x;
        e_s[k4 + 1][rr] = v.y;
        e_s[k4 + 2][rr] = v.z;
        e_s[k4 + 3][rr] = v.w;
      }
      for (int i = threadIdx.x; i < kCT * kKC; i += kRows * kGroups) {
        const int c = i / kKC, k = i % kKC;
        w_s[c][k] = (c0 + c < V1) ? __ldg(W + static_cast<size_t>(c0 + c) * D + k0 + k) : 0.f;
      }
      __syncthreads();
#pragma unroll 8
      for (int k = 0; k < kKC; ++k) {
        const float x = e_s[k][r];
#pragma unroll
        for (int c = 0; c < kCG; ++c) acc[c] = fmaf(w_s[cg * kCG + c][k], x, acc[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < kCG; ++c) {
      const int cls = c0 + cg * kCG + c;
      if (cls < V1) {
        const float v = acc[c] + __ldg(bias + cls);
        if (v > best) { best = v; best_i = cls; }       // ascending classes, strict >: first maximal index wins
      }
    }
  }
  // merge the class groups; a group only ever holds classes c0 + cg * 9 + j, so "first maximal index" = smallest index
  // among equal values, decided explicitly
  best_v[cg][r] = best;
  best_c[cg][r] = best_i;
  __syncthreads();
  if (cg == 0 && row0 + r < R) {
#pragma unroll
    for (int g = 1; g < kGroups; ++g) {
      const float v = best_v[g][r];
      const int i = best_c[g][r];
      if (v > best || (v == best && i < best_i)) { best = v; best_i = i; }
    }
    labels[row0 + r] = best_i;
  }
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
  ctc_argmax_kernel<<<(R + kRows - 1) / kRows, kRows * kGroups, 0, s>>>(enc, W, bias, labels, R, D, V1);
}
void launch_ctc_collapse(const int* labels, const int* len, int B, int T, int blank, int* ids, int* frames, int* counts,
                         cudaStream_t s) {
  ctc_collapse_kernel<<<(B + 3) / 4, 128, 0, s>>>(labels, len, B, T, blank, ids, frames, counts);
}

}  // namespace gam
