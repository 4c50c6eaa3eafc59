// RNN-T greedy decode, cluster-resident variant (gigaam/decoding.py:128-207, gigaam/decoder.py:24-102).
//
// The serial recurrence per utterance (see rnnt.cu) is latency bound; what made the one-CTA-per-utterance kernel
// slow was streaming W_hh (1.6 MB fp32) and W_p through one SM's L2 port on every emission.  Here a thread-block
// cluster of 16 CTAs keeps the recurrent weights RESIDENT in distributed shared memory, sliced by hidden unit:
// CTA c owns units [c*H/16, (c+1)*H/16) -> its 4 gate rows of W_hh (80 x 320 fp32) and its rows of W_p (20 x 320).
// A cluster decodes a group of up to 8 utterances in lock-step (utterances are independent; every CTA replays
// the same control flow from the same exchanged argmax results):
//   LSTM phase  (only utterances that just emitted): own gate rows . h  -> c', h' slice -> DSMEM broadcast
//   pred phase  : own rows of W_p . h'                                  -> DSMEM broadcast
//   joint phase : hid = relu(W_e e_t + b_e + pg);  own slice of classes -> local (max, argmax) -> DSMEM all-to-all
// with one cluster barrier after each phase.  All arithmetic is fp32 as in the reference head.
#include <cooperative_groups.h>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace gam {
namespace {

constexpr int kCl = 16;           // CTAs per cluster
constexpr int kNU = 4;            // utterances decoded in lock-step per cluster
constexpr int kWoRows = 8;        // class rows of W_o kept in smem per CTA (covers V+1 <= 128)
constexpr int kH = 320;
constexpr int kHS = kH / kCl;     // hidden units owned per CTA (20)
constexpr int kThreads = 512;
constexpr int kPitch = kH + 1;    // conflict-free row walks

struct RnntClParams {
  const float* encproj;    // [B*T, H]
  const int* len;          // [B]
  const float* emb_gates;  // [V1, 4H]
  const float* whhT;       // [H, 4H]  (W_hh^T)
  const float* wpT;        // [H, H]   (W_p^T)
  const float* bp;
  const float* wo;         // [V1, H]
  const float* bo;
  int B, T, V1, blank, max_symbols, max_out, num_groups, nu;   // nu <= kNU utterances per group
  int* ids;
  int* frames;
  int* counts;
};

struct Smem {
  float whh[4 * kHS][kPitch];     // rows: gate g, unit j  ->  g*kHS + j
  float wp[kHS][kPitch];
  float h[kNU][kH];               // committed hidden state (full vector, replicated in every CTA)
  float hn[2][kNU][kH];           // candidate h' (full vector, assembled from all CTAs); double-buffered by round parity
  float pg[kNU][kH];              // W_p h' + b_p (full vector)
  float hid[kNU][kH];
  float c[kNU][kHS];              // committed cell state, own units
  float cn[kNU][kHS];
  float gates[kNU][4 * kHS];
  float best_v[2][kNU][kCl];      // per-CTA partial argmax, written by every CTA of the cluster (round parity)
  int best_i[2][kNU][kCl];
  float wbest_v[kThreads / 32][kNU];
  int wbest_i[kThreads / 32][kNU];
  // cluster.sync() invalidates L1, so everything a round needs from global memory is staged here once per round,
  // right after the decision, with a single exposed L2 latency
  float ep[kNU][kH];              // encoder projection row of the utterance's current frame
  float eg[kNU][4 * kHS];         // own gate rows of emb_gates[label]
  float wo[kWoRows][kPitch];      // own class rows of W_o (small vocabularies only)
  float bo[kWoRows];
};

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(kThreads, 1) rnnt_cluster_kernel(const RnntClParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cluster_id = blockIdx.x / kCl;
  const int num_clusters = gridDim.x / kCl;
  const int G = 4 * kH;

  // ---- resident weight slices (once per kernel)
  for (int i = tid; i < 4 * kHS * kH; i += kThreads) {
    const int r = i / kH, k = i % kH;                    // r = g*kHS + j
    const int g = r / kHS, j = r % kHS;
    s.whh[r][k] = __ldg(p.whhT + static_cast<size_t>(k) * G + g * kH + rank * kHS + j);
  }
  for (int i = tid; i < kHS * kH; i += kThreads) {
    const int j = i / kH, k = i % kH;
    s.wp[j][k] = __ldg(p.wpT + static_cast<size_t>(k) * kH + rank * kHS + j);
  }
  const int cls_per = (p.V1 + kCl - 1) / kCl;
  const int cls0 = rank * cls_per;
  const int cls1 = min(p.V1, cls0 + cls_per);
  const bool wo_smem = cls_per <= kWoRows;
  if (wo_smem) {
    for (int i = tid; i < (cls1 - cls0) * kH; i += kThreads) s.wo[i / kH][i % kH] = __ldg(p.wo + static_cast<size_t>(cls0 + i / kH) * kH + i % kH);
    for (int i = tid; i < cls1 - cls0; i += kThreads) s.bo[i] = __ldg(p.bo + cls0 + i);
  }
  __syncthreads();

  for (int group = cluster_id; group < p.num_groups; group += num_clusters) {
    // ---- per-utterance control state: identical in every thread of every CTA of the cluster
    int t_u[kNU], nsym[kNU], cnt[kNU], label[kNU], L[kNU];
    bool need_lstm[kNU];
    // hb[u]: which of the two hn buffers holds utterance u's CURRENT candidate state.  A candidate may be consumed
    // many rounds after it was computed (blank frames in between), so the buffer is chosen per utterance, not per
    // round; the next LSTM of u writes the other buffer, which keeps remote writes of round r+1 away from the
    // commit reads of round r.
    int hb[kNU];
#pragma unroll
    for (int u = 0; u < kNU; ++u) {
      const int ug = group * p.nu + u;
      L[u] = (u < p.nu && ug < p.B) ? min(max(p.len[ug], 0), p.T) : 0;
      t_u[u] = 0; nsym[u] = 0; cnt[u] = 0; label[u] = p.blank;
      need_lstm[u] = L[u] > 0;
      hb[u] = 0;
    }
    for (int i = tid; i < kNU * kH; i += kThreads) (&s.h[0][0])[i] = 0.f;
    for (int i = tid; i < kNU * kHS; i += kThreads) (&s.c[0][0])[i] = 0.f;
    for (int d = tid; d < kNU * kH; d += kThreads) {
      const int u = d / kH, j = d % kH;
      const int ug = group * p.nu + u;
      s.ep[u][j] = L[u] > 0 ? __ldg(p.encproj + static_cast<size_t>(ug) * p.T * kH + j) : 0.f;
    }
    for (int d = tid; d < kNU * 4 * kHS; d += kThreads) {
      const int r = d % (4 * kHS);
      s.eg[d / (4 * kHS)][r] = __ldg(p.emb_gates + static_cast<size_t>(p.blank) * G + (r / kHS) * kH + rank * kHS + r % kHS);
    }
    cluster.sync();
    int round = 0;

    while (true) {
      const int par = round & 1;
      ++round;
      bool any_active = false, any_lstm = false;
#pragma unroll
      for (int u = 0; u < kNU; ++u) {
        any_active |= t_u[u] < L[u];
        any_lstm |= need_lstm[u] && t_u[u] < L[u];
      }
      if (!any_active) break;

      if (any_lstm) {
        // ---------------- LSTM: own gate rows for every utterance that needs a new prediction-network state
        for (int d = tid; d < kNU * 4 * kHS; d += kThreads) {
          const int u = d / (4 * kHS), r = d % (4 * kHS);
          if (!(need_lstm[u] && t_u[u] < L[u])) continue;
          const int g = r / kHS, j = r % kHS;
          const float* w = s.whh[r];
          const float* hv = s.h[u];
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
          for (int k = 0; k < kH; k += 4) {
            a0 = fmaf(w[k], hv[k], a0);
            a1 = fmaf(w[k + 1], hv[k + 1], a1);
            a2 = fmaf(w[k + 2], hv[k + 2], a2);
            a3 = fmaf(w[k + 3], hv[k + 3], a3);
          }
          (void)g; (void)j;
          s.gates[u][r] = s.eg[u][r] + ((a0 + a1) + (a2 + a3));
        }
        __syncthreads();
        for (int d = tid; d < kNU * kHS; d += kThreads) {
          const int u = d / kHS, j = d % kHS;
          if (!(need_lstm[u] && t_u[u] < L[u])) continue;
          const float ig = sigm(s.gates[u][j]), fg = sigm(s.gates[u][kHS + j]);
          const float gg = tanhf(s.gates[u][2 * kHS + j]), og = sigm(s.gates[u][3 * kHS + j]);
          const float cn = fg * s.c[u][j] + ig * gg;
          s.cn[u][j] = cn;
          const float hn = og * tanhf(cn);
#pragma unroll
          for (int rr = 0; rr < kCl; ++rr) cluster.map_shared_rank(&s.hn[hb[u] ^ 1][u][rank * kHS + j], rr)[0] = hn;
        }
        cluster.sync();
        // ---------------- prediction projection: own rows of W_p
        for (int d = tid; d < kNU * kHS; d += kThreads) {
          const int u = d / kHS, j = d % kHS;
          if (!(need_lstm[u] && t_u[u] < L[u])) continue;
          const float* w = s.wp[j];
          const float* hv = s.hn[hb[u] ^ 1][u];
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
          for (int k = 0; k < kH; k += 4) {
            a0 = fmaf(w[k], hv[k], a0);
            a1 = fmaf(w[k + 1], hv[k + 1], a1);
            a2 = fmaf(w[k + 2], hv[k + 2], a2);
            a3 = fmaf(w[k + 3], hv[k + 3], a3);
          }
          const float v = __ldg(p.bp + rank * kHS + j) + ((a0 + a1) + (a2 + a3));
#pragma unroll
          for (int rr = 0; rr < kCl; ++rr) cluster.map_shared_rank(&s.pg[u][rank * kHS + j], rr)[0] = v;
        }
        cluster.sync();
#pragma unroll
        for (int u = 0; u < kNU; ++u)
          if (need_lstm[u] && t_u[u] < L[u]) hb[u] ^= 1;   // the fresh candidate is now the current one
      }

      // ---------------- joint: hid = relu(enc_proj[t] + pg), own class slice, local argmax
      for (int d = tid; d < kNU * kH; d += kThreads) {
        const int u = d / kH, j = d % kH;
        const int ug = group * p.nu + u;
        float v = 0.f;
        (void)ug;
        if (t_u[u] < L[u]) v = fmaxf(s.ep[u][j] + s.pg[u][j], 0.f);
        s.hid[u][j] = v;
      }
      __syncthreads();
      float bv[kNU];
      int bi[kNU];
#pragma unroll
      for (int u = 0; u < kNU; ++u) { bv[u] = -INFINITY; bi[u] = 0x7fffffff; }
      for (int cls = cls0 + warp; cls < cls1; cls += kThreads / 32) {
        const float* w = wo_smem ? s.wo[cls - cls0] : p.wo + static_cast<size_t>(cls) * kH;
        float acc[kNU];
#pragma unroll
        for (int u = 0; u < kNU; ++u) acc[u] = 0.f;
        for (int k = lane; k < kH; k += 32) {
          const float wv = w[k];
#pragma unroll
          for (int u = 0; u < kNU; ++u) acc[u] = fmaf(wv, s.hid[u][k], acc[u]);
        }
        const float bo = wo_smem ? s.bo[cls - cls0] : __ldg(p.bo + cls);
#pragma unroll
        for (int u = 0; u < kNU; ++u) {
          float a = acc[u];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
          a += bo;
          if (a > bv[u]) { bv[u] = a; bi[u] = cls; }   // classes ascend within a warp: first max wins
        }
      }
      if (lane == 0) {
#pragma unroll
        for (int u = 0; u < kNU; ++u) { s.wbest_v[warp][u] = bv[u]; s.wbest_i[warp][u] = bi[u]; }
      }
      __syncthreads();
      if (tid < kNU) {
        const int u = tid;
        float v0 = s.wbest_v[0][u];
        int i0 = s.wbest_i[0][u];
        for (int w = 1; w < kThreads / 32; ++w) {
          const float v = s.wbest_v[w][u];
          const int i = s.wbest_i[w][u];
          if (v > v0 || (v == v0 && i < i0)) { v0 = v; i0 = i; }
        }
        for (int rr = 0; rr < kCl; ++rr) {
          cluster.map_shared_rank(&s.best_v[par][u][rank], rr)[0] = v0;
          cluster.map_shared_rank(&s.best_i[par][u][rank], rr)[0] = i0;
        }
      }
      cluster.sync();

      // ---------------- every thread of every CTA replays the same decision
#pragma unroll
      for (int u = 0; u < kNU; ++u) {
        if (!(t_u[u] < L[u])) continue;
        float v0 = s.best_v[par][u][0];
        int k = s.best_i[par][u][0];
        for (int rr = 1; rr < kCl; ++rr) {
          const float v = s.best_v[par][u][rr];
          const int i = s.best_i[par][u][rr];
          if (v > v0 || (v == v0 && i < k)) { v0 = v; k = i; }
        }
        if (k == p.blank) {
          t_u[u] += 1;
          nsym[u] = 0;
          need_lstm[u] = false;
        } else {
          const int ug = group * p.nu + u;
          if (rank == 0 && tid == 0 && cnt[u] < p.max_out) {
            p.ids[static_cast<size_t>(ug) * p.max_out + cnt[u]] = k;
            p.frames[static_cast<size_t>(ug) * p.max_out + cnt[u]] = t_u[u];
          }
          cnt[u] += 1;
          label[u] = k;
          need_lstm[u] = true;   // commit (h', c') below and advance the prediction network on the new label
          nsym[u] += 1;
          if (nsym[u] >= p.max_symbols) { t_u[u] += 1; nsym[u] = 0; }
        }
      }
      __syncthreads();   // all reads of best_* / hn done before the commit below and the next round's writes
      // commit candidate state for utterances that emitted (their need_lstm was just set)
      for (int d = tid; d < kNU * kH; d += kThreads) {
        const int u = d / kH;
        if (need_lstm[u]) s.h[u][d % kH] = s.hn[hb[u]][u][d % kH];
      }
      for (int d = tid; d < kNU * kHS; d += kThreads) {
        const int u = d / kHS;
        if (need_lstm[u]) s.c[u][d % kHS] = s.cn[u][d % kHS];
      }
      // stage what the next round needs from global memory (one exposed L2 round trip per round)
      for (int d = tid; d < kNU * kH; d += kThreads) {
        const int u = d / kH, j = d % kH;
        if (t_u[u] < L[u]) s.ep[u][j] = __ldg(p.encproj + (static_cast<size_t>(group * p.nu + u) * p.T + t_u[u]) * kH + j);
      }
      for (int d = tid; d < kNU * 4 * kHS; d += kThreads) {
        const int u = d / (4 * kHS), r = d % (4 * kHS);
        if (need_lstm[u] && t_u[u] < L[u])
          s.eg[u][r] = __ldg(p.emb_gates + static_cast<size_t>(label[u]) * G + (r / kHS) * kH + rank * kHS + r % kHS);
      }
      __syncthreads();
    }
    if (rank == 0 && tid < kNU) {
      const int ug = group * p.nu + tid;
      if (tid < p.nu && ug < p.B) p.counts[ug] = min(cnt[tid], p.max_out);
    }
    cluster.sync();
  }
}

}  // namespace

// returns 0 on success, 1 if a 16-CTA cluster cannot be scheduled on this device (caller falls back to the
// per-utterance kernel), negative on error
int launch_rnnt_greedy_cluster(const float* encproj, const int* len, const float* emb_gates, const float* whhT, const float* wpT,
                               const float* bp, const float* wo, const float* bo, int B, int T, int H, int V1, int blank,
                               int max_symbols, int max_out, int* ids, int* frames, int* counts, cudaStream_t s) {
  if (H != kH) return 1;
  static int max_clusters = -1;
  const int smem = static_cast<int>(sizeof(Smem));
  cudaLaunchConfig_t cfg{};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (max_clusters < 0) {
    if (cudaFuncSetAttribute(rnnt_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(rnnt_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      max_clusters = 0;
    } else {
      cfg.gridDim = dim3(kCl);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, rnnt_cluster_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      max_clusters = n;
    }
  }
  if (max_clusters <= 0) return 1;
  RnntClParams p;
  p.encproj = encproj; p.len = len; p.emb_gates = emb_gates; p.whhT = whhT; p.wpT = wpT; p.bp = bp; p.wo = wo; p.bo = bo;
  p.B = B; p.T = T; p.V1 = V1; p.blank = blank; p.max_symbols = max_symbols; p.max_out = max_out;
  // spread utterances over as many clusters as can be resident: fewer lock-stepped utterances per cluster
  int nu = (B + max_clusters - 1) / max_clusters;
  if (nu > kNU) nu = kNU;
  p.nu = nu;
  p.num_groups = (B + nu - 1) / nu;
  const int nclusters = p.num_groups < max_clusters ? p.num_groups : max_clusters;
  p.ids = ids; p.frames = frames; p.counts = counts;
  cfg.gridDim = dim3(nclusters * kCl);
  if (cudaLaunchKernelEx(&cfg, rnnt_cluster_kernel, p) != cudaSuccess) return -2;
  return 0;
}

}  // namespace gam
