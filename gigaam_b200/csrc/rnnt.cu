// RNN-T greedy decode, device resident (gigaam/decoding.py:128-207, gigaam/decoder.py:24-102).
//
// Per utterance the reference's batched Python loop is exactly this serial recurrence:
//   (label, h, c) = (blank -> zero embedding, 0, 0);  (g, h', c') = LSTM(embed(label), h, c)
//   for t < len:  repeat <= max_symbols:  k = argmax W_o relu(W_e e_t + b_e + W_p g + b_p) + b_o
//                   k == blank -> next t ;  else emit (k, t), (label,h,c) = (k,h',c'), re-run LSTM
// The LSTM only runs after an emission; the encoder projection W_e e_t + b_e is hoisted out of the
// loop as one fp32 GEMM over all frames; embed(k) W_ih^T + b_ih + b_hh is a lookup table built at
// load time.  All head arithmetic stays fp32 (the reference never casts the head to fp16,
// gigaam/__init__.py:188-189).  One CTA per utterance; recurrent weights stream from L2.
#include "kernels.h"

namespace gam {
namespace {

// ------------------------------------------------------------------ fp32 GEMM  C[M,N] = A[M,K] W[N,K]^T + bias
constexpr int kSgBM = 64, kSgBN = 64, kSgBK = 16;
__global__ void __launch_bounds__(256) sgemm_tn_bias_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                            const float* __restrict__ bias, float* __restrict__ C, int M,
                                                            int N, int K) {
  __shared__ float As[kSgBK][kSgBM + 4];
  __shared__ float Ws[kSgBK][kSgBN + 4];
  const int m0 = blockIdx.y * kSgBM, n0 = blockIdx.x * kSgBN;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += kSgBK) {
    {
      const int r = threadIdx.x / 4, k4 = (threadIdx.x % 4) * 4;
      float4 a = make_float4(0, 0, 0, 0), w = make_float4(0, 0, 0, 0);
      if (m0 + r < M) a = *reinterpret_cast<const float4*>(A + static_cast<size_t>(m0 + r) * K + k0 + k4);
      if (n0 + r < N) w = *reinterpret_cast<const float4*>(W + static_cast<size_t>(n0 + r) * K + k0 + k4);
      As[k4 + 0][r] = a.x; As[k4 + 1][r] = a.y; As[k4 + 2][r] = a.z; As[k4 + 3][r] = a.w;
      Ws[k4 + 0][r] = w.x; Ws[k4 + 1][r] = w.y; Ws[k4 + 2][r] = w.z; Ws[k4 + 3][r] = w.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kSgBK; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = As[k][ty * 4 + i]; w[i] = Ws[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < N) C[static_cast<size_t>(m) * N + n] = acc[i][j] + (bias ? bias[n] : 0.f);
    }
  }
}

// ------------------------------------------------------------------ greedy loop
constexpr int kRnntThreads = 512;
constexpr int kMaxH = 320;

struct RnntParams {
  const float* encproj;   // [B*T, H]  W_e e + b_e
  const int* len;         // [B]
  const float* emb_gates; // [V1, 4H]  embed(k) W_ih^T + b_ih + b_hh  (row blank = biases only)
  const float* whhT;      // [H, 4H]   W_hh^T
  const float* wpT;       // [H, H]    joint.pred weight^T
  const float* bp;        // [H]
  const float* wo;        // [V1, H]   joint_net.1 weight
  const float* bo;        // [V1]
  int T, H, V1, blank, max_symbols, max_out;
  int* ids;               // [B, max_out]
  int* frames;            // [B, max_out]
  int* counts;            // [B]
};

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kRnntThreads) rnnt_greedy_kernel(const RnntParams p) {
  __shared__ float h_s[kMaxH], c_s[kMaxH];          // committed state
  __shared__ float hn_s[kMaxH], cn_s[kMaxH];        // candidate state (after LSTM on current label)
  __shared__ float pg_s[kMaxH];                     // W_p g + b_p
  __shared__ float hid_s[kMaxH];
  __shared__ float gates_s[4 * kMaxH];
  __shared__ float best_v[kRnntThreads / 32];
  __shared__ int best_i[kRnntThreads / 32];
  __shared__ int k_s;

  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int H = p.H, G = 4 * p.H;
  const int L = min(max(p.len[b], 0), p.T);
  int* ids = p.ids + static_cast<size_t>(b) * p.max_out;
  int* frames = p.frames + static_cast<size_t>(b) * p.max_out;

  for (int j = tid; j < H; j += blockDim.x) { h_s[j] = 0.f; c_s[j] = 0.f; }
  __syncthreads();

  auto lstm_step = [&](int label) {
    // gates = emb_gates[label] + W_hh h
    for (int j = tid; j < G; j += blockDim.x) {
      float acc = __ldg(p.emb_gates + static_cast<size_t>(label) * G + j);
      const float* w = p.whhT + j;
#pragma unroll 8
      for (int i = 0; i < H; ++i) acc = fmaf(__ldg(w + static_cast<size_t>(i) * G), h_s[i], acc);
      gates_s[j] = acc;
    }
    __syncthreads();
    for (int j = tid; j < H; j += blockDim.x) {
      const float ig = sigmoid_acc(gates_s[j]);
      const float fg = sigmoid_acc(gates_s[H + j]);
      const float gg = tanhf(gates_s[2 * H + j]);
      const float og = sigmoid_acc(gates_s[3 * H + j]);
      const float cn = fg * c_s[j] + ig * gg;
      cn_s[j] = cn;
      hn_s[j] = og * tanhf(cn);
    }
    __syncthreads();
    for (int j = tid; j < H; j += blockDim.x) {
      float acc = __ldg(p.bp + j);
      const float* w = p.wpT + j;
#pragma unroll 8
      for (int i = 0; i < H; ++i) acc = fmaf(__ldg(w + static_cast<size_t>(i) * H), hn_s[i], acc);
      pg_s[j] = acc;
    }
    __syncthreads();
  };

  lstm_step(p.blank);
  int count = 0;
  for (int t = 0; t < L; ++t) {
    const float* ep = p.encproj + (static_cast<size_t>(b) * p.T + t) * H;
    for (int sidx = 0; sidx < p.max_symbols; ++sidx) {
      for (int j = tid; j < H; j += blockDim.x) hid_s[j] = fmaxf(ep[j] + pg_s[j], 0.f);
      __syncthreads();
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      for (int v = warp; v < p.V1; v += kRnntThreads / 32) {
        const float* w = p.wo + static_cast<size_t>(v) * H;
        float acc = 0.f;
        for (int j = lane; j < H; j += 32) acc = fmaf(__ldg(w + j), hid_s[j], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        acc += __ldg(p.bo + v);
        if (acc > bv) { bv = acc; bi = v; }  // v ascending within a warp: first max wins
      }
      if (lane == 0) { best_v[warp] = bv; best_i[warp] = bi; }
      __syncthreads();
      if (tid == 0) {
        float v0 = best_v[0];
        int i0 = best_i[0];
        for (int w = 1; w < kRnntThreads / 32; ++w) {
          if (best_v[w] > v0 || (best_v[w] == v0 && best_i[w] < i0)) { v0 = best_v[w]; i0 = best_i[w]; }
        }
        k_s = i0;
      }
      __syncthreads();
      const int k = k_s;
      if (k == p.blank) break;
      if (tid == 0 && count < p.max_out) { ids[count] = k; frames[count] = t; }
      ++count;
      // commit candidate state, then advance the prediction network on the new label
      for (int j = tid; j < H; j += blockDim.x) { h_s[j] = hn_s[j]; c_s[j] = cn_s[j]; }
      __syncthreads();
      lstm_step(k);
    }
  }
  if (tid == 0) p.counts[b] = min(count, p.max_out);
}

}  // namespace

void launch_sgemm_tn_bias(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, cudaStream_t s) {
  dim3 grid((N + kSgBN - 1) / kSgBN, (M + kSgBM - 1) / kSgBM);
  sgemm_tn_bias_kernel<<<grid, 256, 0, s>>>(A, W, bias, C, M, N, K);
}

int launch_rnnt_greedy(const float* encproj, const int* len, const float* emb_gates, const float* whhT, const float* wpT,
                       const float* bp, const float* wo, const float* bo, int B, int T, int H, int V1, int blank,
                       int max_symbols, int max_out, int* ids, int* frames, int* counts, cudaStream_t s) {
  if (H > kMaxH) return -1;
  RnntParams p;
  p.encproj = encproj; p.len = len; p.emb_gates = emb_gates; p.whhT = whhT; p.wpT = wpT; p.bp = bp; p.wo = wo; p.bo = bo;
  p.T = T; p.H = H; p.V1 = V1; p.blank = blank; p.max_symbols = max_symbols; p.max_out = max_out;
  p.ids = ids; p.frames = frames; p.counts = counts;
  rnnt_greedy_kernel<<<B, kRnntThreads, 0, s>>>(p);
  return 0;
}

}  // namespace gam
