// RNN-T greedy decode, cluster-resident variant (gigaam/decoding.py:128-207, gigaam/decoder.py:24-102).
//
// The serial recurrence per utterance (see rnnt.cu) is latency bound; what made the one-CTA-per-utterance kernel
// slow was streaming W_hh (1.6 MB fp32) and W_p through one SM's L2 port on every emission.  Here a thread-block
// cluster of 16 CTAs keeps the weights RESIDENT in distributed shared memory: CTA c owns hidden units
// [c*H/16, (c+1)*H/16) -> its 4 gate rows of W_hh (80 x 320 fp32), its rows of W_p (20 x 320) and its slice of the
// output classes (rows of W_o: as many as fit next to the recurrent weights, the rest is prefetched from L2 into
// registers at the top of every joint phase).  A cluster decodes a group of up to 4 utterances in lock-step
// (utterances are independent; every CTA replays the same control flow from the same exchanged argmax results):
//   LSTM phase  (only utterances that just emitted): own gate rows . h  -> c', h' slice -> DSMEM broadcast
//   pred phase  : own rows of W_p . h'                                  -> DSMEM broadcast
//   joint phase : hid = relu(W_e e_t + b_e + pg);  own slice of classes -> local (max, argmax) -> DSMEM all-to-all
// with one cluster barrier after each phase.
//
// Layout rules that came out of measuring the first versions (profiles/r1d, r1e):
//  * every state vector is utterance-interleaved ([unit] -> float4 of the four utterances): a weight is read from
//    shared memory once per phase for all four utterances, and -- more important -- every DSMEM exchange is a
//    16-byte store issued by a different thread (672 remote stores per emission round instead of 2 700 4-byte ones;
//    remote stores are paid on the producer's shared-memory port, ~20 B/clk).
//  * the prediction-network state is double-buffered by a cluster-wide parity that flips on every LSTM round;
//    utterances that do not step in that round carry their state over inside the same float4, so a state that is
//    consumed many rounds after it was produced (blank frames) is always in the current buffer.
//  * everything a round needs from global memory is requested one phase (or one frame) ahead and parked in
//    registers: cluster.sync() invalidates L1, so every global load costs a full L2 round trip.
// All arithmetic is fp32 as in the reference head.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace cg = cooperative_groups;

namespace gam {
namespace {

constexpr int kCl = 16;           // CTAs per cluster
constexpr int kNU = 4;            // utterances decoded in lock-step per cluster (= components of a float4)
constexpr int kH = 320;
constexpr int kHS = kH / kCl;     // hidden units owned per CTA (20)
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kWhhP = 324;        // W_hh row pitch (words), 4 mod 32: (8 rows x 4 k-lanes) warps read conflict-free
constexpr int kWpP = 336;         // W_p row pitch, 16 mod 32: (2 rows x 16 k-lanes) warps read conflict-free
constexpr int kWoPitch = kH + 1;  // W_o row pitch: conflict-free 4-byte row walks
constexpr int kCB = 5;            // classes accumulated together per warp (one hid read serves all of them)
constexpr int kGP = 2;            // L2-resident class rows a warp prefetches into registers per round

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
  int rows_smem;           // class rows of W_o resident in shared memory per CTA
  int cls_pad;             // floats reserved for the bias slice
  int* ids;
  int* frames;
  int* counts;
};

struct Smem {
  float whh[4 * kHS][kWhhP];      // rows: gate g, unit j  ->  g*kHS + j
  float wp[kHS][kWpP];
  float4 h4[2][kH];               // prediction-network state h, [parity][unit] -> 4 utterances (replicated in every CTA)
  float4 pg4[kH];                 // W_p h' + b_p
  float4 hid4[kH];                // relu(enc_proj[t] + pg)
  float4 hnew4[kHS];              // own slice of the next state, staged for the 16-byte broadcast
  float c[2][kNU][kHS];           // cell state of the own units, same parity as h4
  float gates[kNU][4 * kHS];
  float4 best_v[2][kCl];          // per-CTA partial argmax of the 4 utterances, written by every CTA (round parity)
  int4 best_i[2][kCl];
  float wbest_v[kWarps][kNU];
  int wbest_i[kWarps][kNU];
  // followed by: float bo[cls_pad]; float wo[rows_smem][kWoPitch];
};

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float comp(const float4& v, int u) { return u == 0 ? v.x : (u == 1 ? v.y : (u == 2 ? v.z : v.w)); }

// sum of v over the warp; returns the total of component u = (lane >> 3) & 3  (6 shuffles instead of 20)
__device__ __forceinline__ float reduce4(const float4 v, int lane) {
  const bool hi = (lane & 16) != 0;
  float k0 = hi ? v.z : v.x, k1 = hi ? v.w : v.y;
  k0 += __shfl_xor_sync(0xffffffffu, hi ? v.x : v.z, 16);
  k1 += __shfl_xor_sync(0xffffffffu, hi ? v.y : v.w, 16);
  const bool mid = (lane & 8) != 0;
  float k = mid ? k1 : k0;
  k += __shfl_xor_sync(0xffffffffu, mid ? k0 : k1, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}

__device__ __forceinline__ void fma4(float4& a, float w, const float4& h) {
  a.x = fmaf(w, h.x, a.x); a.y = fmaf(w, h.y, a.y); a.z = fmaf(w, h.z, a.z); a.w = fmaf(w, h.w, a.w);
}

#ifdef GAM_RNNT_DBG
// phase timing of cluster 0 / CTA 0 / thread 0 (tools only; never compiled into the shipped library)
__device__ long long g_rnnt_dbg[16];
#define DBG_T(i) do { if (dbg_on) { const long long t_now = clock64(); dbg_acc[i] += t_now - dbg_last; dbg_last = t_now; } } while (0)
#else
#define DBG_T(i) do { } while (0)
#endif

__global__ void __launch_bounds__(kThreads, 1) rnnt_cluster_kernel(const RnntClParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  float* s_bo = reinterpret_cast<float*>(smem_raw + sizeof(Smem));
  float* s_wo = s_bo + p.cls_pad;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cluster_id = blockIdx.x / kCl;
  const int num_clusters = gridDim.x / kCl;
  const int G = 4 * kH;
#ifdef GAM_RNNT_DBG
  const bool dbg_on = blockIdx.x == 0 && threadIdx.x == 0;
  long long dbg_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  long long dbg_last = clock64();
#endif

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
  const int cls0 = min(p.V1, rank * cls_per);
  const int ncls = min(p.V1, cls0 + cls_per) - cls0;       // classes owned by this CTA
  const int nsm = min(ncls, p.rows_smem);                  // ... of which resident in shared memory
  for (int i = tid; i < nsm * kH; i += kThreads) s_wo[(i / kH) * kWoPitch + i % kH] = __ldg(p.wo + static_cast<size_t>(cls0 + i / kH) * kH + i % kH);
  for (int i = tid; i < ncls; i += kThreads) s_bo[i] = __ldg(p.bo + cls0 + i);
  // LSTM-phase role of this thread: gate row r, k-lane q; lane q of a quad also finishes utterance q of that row
  const int lq = lane & 3, lr_row = warp * 8 + (lane >> 2);
  const bool gate_thread = warp < 4 * kHS / 8;
  const size_t eg_off = gate_thread ? static_cast<size_t>((lr_row / kHS) * kH + rank * kHS + lr_row % kHS) : 0;
  // pred-phase role: row pj, k-lane pq
  const int pq = lane & 15, pj = warp * 2 + (lane >> 4);
  const bool pred_thread = warp < kHS / 2;
  const float my_bp = pred_thread ? __ldg(p.bp + rank * kHS + pj) : 0.f;
  __syncthreads();

  for (int group = cluster_id; group < p.num_groups; group += num_clusters) {
    // ---- per-utterance control state: identical in every thread of every CTA of the cluster
    int t_u[kNU], nsym[kNU], cnt[kNU], label[kNU], L[kNU];
    bool need_lstm[kNU];
    int gb = 0;   // parity of the h4 / c buffer that holds the current prediction-network state of all utterances
#pragma unroll
    for (int u = 0; u < kNU; ++u) {
      const int ug = group * p.nu + u;
      L[u] = (u < p.nu && ug < p.B) ? min(max(p.len[ug], 0), p.T) : 0;
      t_u[u] = 0; nsym[u] = 0; cnt[u] = 0; label[u] = p.blank;
      need_lstm[u] = L[u] > 0;
    }
    for (int i = tid; i < 2 * kH; i += kThreads) (&s.h4[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < kH; i += kThreads) s.pg4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 2 * kNU * kHS; i += kThreads) (&s.c[0][0][0])[i] = 0.f;
    // encoder projection of the current frame (ep) and of the next one (epn): thread k < H keeps the four
    // utterances' values in registers; a frame advance promotes epn and requests the frame after it
    float4 ep = make_float4(0.f, 0.f, 0.f, 0.f), epn = ep;
    const float* ep_base = p.encproj + static_cast<size_t>(group * p.nu) * p.T * kH + (tid < kH ? tid : 0);
    if (tid < kH) {
      if (L[0] > 0) ep.x = __ldg(ep_base);
      if (L[1] > 0) ep.y = __ldg(ep_base + static_cast<size_t>(1) * p.T * kH);
      if (L[2] > 0) ep.z = __ldg(ep_base + static_cast<size_t>(2) * p.T * kH);
      if (L[3] > 0) ep.w = __ldg(ep_base + static_cast<size_t>(3) * p.T * kH);
      if (L[0] > 1) epn.x = __ldg(ep_base + kH);
      if (L[1] > 1) epn.y = __ldg(ep_base + (static_cast<size_t>(1) * p.T + 1) * kH);
      if (L[2] > 1) epn.z = __ldg(ep_base + (static_cast<size_t>(2) * p.T + 1) * kH);
      if (L[3] > 1) epn.w = __ldg(ep_base + (static_cast<size_t>(3) * p.T + 1) * kH);
    }
    // embedding contribution to this thread's gate (row lr_row, utterance lq); reloaded after every emission
    float eg = gate_thread ? __ldg(p.emb_gates + static_cast<size_t>(p.blank) * G + eg_off) : 0.f;
    cluster.sync();
    DBG_T(6);
    int round = 0;

    while (true) {
      const int par = round & 1;
      ++round;
      bool act[kNU];             // still decoding
      int run_m = 0;             // bit u: utterance u needs an LSTM step this round
      bool any_active = false;
#pragma unroll
      for (int u = 0; u < kNU; ++u) {
        act[u] = t_u[u] < L[u];
        if (need_lstm[u] && act[u]) run_m |= 1 << u;
        any_active |= act[u];
      }
      if (!any_active) break;
      DBG_T(9);
#ifdef GAM_RNNT_DBG
      dbg_acc[7] += 1;
      if (run_m != 0) dbg_acc[8] += 1;
#endif

      if (run_m != 0) {
        // ---------------- gates: warp = 8 own rows x 4 k-lanes (k = 16 i + 4 e + q)
        if (gate_thread) {
          const float* wrow = &s.whh[lr_row][lq];
          const float4* hv = &s.h4[gb][lq];
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
          for (int i = 0; i < kH / 16; ++i) {
            fma4(a, wrow[16 * i], hv[16 * i]);
            fma4(a, wrow[16 * i + 4], hv[16 * i + 4]);
            fma4(a, wrow[16 * i + 8], hv[16 * i + 8]);
            fma4(a, wrow[16 * i + 12], hv[16 * i + 12]);
          }
#pragma unroll
          for (int o = 1; o <= 2; o <<= 1) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
          }
          s.gates[lq][lr_row] = eg + comp(a, lq);   // lane q of the quad finishes utterance q
        }
        __syncthreads();
        DBG_T(0);
        if (tid < kNU * kHS) {
          const int u = tid / kHS, j = tid % kHS;
          const float c_old = s.c[gb][u][j];
          float cn = c_old, hn = comp(s.h4[gb][rank * kHS + j], u);   // utterances that do not step carry their state over
          if ((run_m >> u) & 1) {
            const float ig = sigm(s.gates[u][j]), fg = sigm(s.gates[u][kHS + j]);
            const float gg = tanhf(s.gates[u][2 * kHS + j]), og = sigm(s.gates[u][3 * kHS + j]);
            cn = fg * c_old + ig * gg;
            hn = og * tanhf(cn);
          }
          s.c[gb ^ 1][u][j] = cn;
          reinterpret_cast<float*>(&s.hnew4[j])[u] = hn;
        }
        __syncthreads();
        if (tid < kCl * kHS) {   // one 16-byte remote store per thread: unit j -> CTA tid / kHS
          const int j = tid % kHS;
          *cluster.map_shared_rank(&s.h4[gb ^ 1][rank * kHS + j], tid / kHS) = s.hnew4[j];
        }
        DBG_T(1);
        cluster.sync();
        DBG_T(2);
        // ---------------- prediction projection: own rows of W_p on the new state (warp = 2 rows x 16 k-lanes)
        if (pred_thread) {
          const float* wrow = &s.wp[pj][pq];
          const float4* hv = &s.h4[gb ^ 1][pq];
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < kH / 64; ++i) {
            fma4(a, wrow[64 * i], hv[64 * i]);
            fma4(a, wrow[64 * i + 16], hv[64 * i + 16]);
            fma4(a, wrow[64 * i + 32], hv[64 * i + 32]);
            fma4(a, wrow[64 * i + 48], hv[64 * i + 48]);
          }
#pragma unroll
          for (int o = 1; o <= 8; o <<= 1) {
            a.x += __shfl_xor_sync(0xffffffffu, a.x, o); a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
            a.z += __shfl_xor_sync(0xffffffffu, a.z, o); a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
          }
          const float4 old = s.pg4[rank * kHS + pj];
          const float4 v = make_float4((run_m & 1) ? my_bp + a.x : old.x, (run_m & 2) ? my_bp + a.y : old.y,
                                       (run_m & 4) ? my_bp + a.z : old.z, (run_m & 8) ? my_bp + a.w : old.w);
          *cluster.map_shared_rank(&s.pg4[rank * kHS + pj], pq) = v;   // lane q of the row's 16 serves CTA q
        }
        DBG_T(3);
        cluster.sync();
        DBG_T(2);
        gb ^= 1;
      }

      // ---------------- joint: hid = relu(enc_proj[t] + pg), own class slice, local argmax
      // class rows that do not fit in shared memory: issue their L2 loads now, consume them after the smem rows
      float wg[kGP][kH / 32];
      int gcls[kGP];
#pragma unroll
      for (int gi = 0; gi < kGP; ++gi) {
        const int lr = nsm + ((warp - nsm) & (kWarps - 1)) + kWarps * gi;   // this warp's gi-th row at or after nsm
        gcls[gi] = lr < ncls ? lr : -1;
        if (gcls[gi] >= 0) {
          const float* w = p.wo + static_cast<size_t>(cls0 + lr) * kH;
#pragma unroll
          for (int kk = 0; kk < kH / 32; ++kk) wg[gi][kk] = __ldcg(w + lane + 32 * kk);
        }
      }
      if (tid < kH) {
        const float4 g4 = s.pg4[tid];
        s.hid4[tid] = make_float4(act[0] ? fmaxf(ep.x + g4.x, 0.f) : 0.f, act[1] ? fmaxf(ep.y + g4.y, 0.f) : 0.f,
                                  act[2] ? fmaxf(ep.z + g4.z, 0.f) : 0.f, act[3] ? fmaxf(ep.w + g4.w, 0.f) : 0.f);
      }
      __syncthreads();
      DBG_T(13);
      const int myu = (lane >> 3) & 3;       // the utterance whose logits this lane ends up holding (reduce4)
      float bv = -INFINITY;
      int bi = 0x7fffffff;
      // shared-memory rows: local rows warp, warp+16, ... (ascending, so the first maximum wins as in torch.argmax)
      for (int lr0 = warp; lr0 < nsm; lr0 += kWarps * kCB) {
        float4 acc[kCB];
#pragma unroll
        for (int c = 0; c < kCB; ++c) acc[c] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
        for (int kk = 0; kk < kH / 32; ++kk) {
          const int k = lane + 32 * kk;
          const float4 hv = s.hid4[k];
#pragma unroll
          for (int c = 0; c < kCB; ++c)
            if (lr0 + kWarps * c < nsm) fma4(acc[c], s_wo[(lr0 + kWarps * c) * kWoPitch + k], hv);
        }
#pragma unroll
        for (int c = 0; c < kCB; ++c) {
          const int lr = lr0 + kWarps * c;
          if (lr < nsm) {
            const float a = reduce4(acc[c], lane) + s_bo[lr];
            if (a > bv) { bv = a; bi = cls0 + lr; }
          }
        }
      }
      // L2 rows (registers), then anything beyond the prefetch depth straight from L2
#pragma unroll
      for (int gi = 0; gi < kGP; ++gi) {
        if (gcls[gi] >= 0) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int kk = 0; kk < kH / 32; ++kk) fma4(acc, wg[gi][kk], s.hid4[lane + 32 * kk]);
          const float a = reduce4(acc, lane) + s_bo[gcls[gi]];
          if (a > bv) { bv = a; bi = cls0 + gcls[gi]; }
        }
      }
      for (int lr = nsm + ((warp - nsm) & (kWarps - 1)) + kWarps * kGP; lr < ncls; lr += kWarps) {
        const float* w = p.wo + static_cast<size_t>(cls0 + lr) * kH;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int kk = 0; kk < kH / 32; ++kk) fma4(acc, __ldcg(w + lane + 32 * kk), s.hid4[lane + 32 * kk]);
        const float a = reduce4(acc, lane) + s_bo[lr];
        if (a > bv) { bv = a; bi = cls0 + lr; }
      }
      DBG_T(14);
      if ((lane & 7) == 0) { s.wbest_v[warp][myu] = bv; s.wbest_i[warp][myu] = bi; }
      __syncthreads();
      DBG_T(15);
      if (tid < kCl) {   // thread rr reduces the 16 warps for all 4 utterances and serves CTA rr with two 16-byte stores
        float v0[kNU];
        int i0[kNU];
#pragma unroll
        for (int u = 0; u < kNU; ++u) { v0[u] = s.wbest_v[0][u]; i0[u] = s.wbest_i[0][u]; }
        for (int w = 1; w < kWarps; ++w) {
#pragma unroll
          for (int u = 0; u < kNU; ++u) {
            const float v = s.wbest_v[w][u];
            const int i = s.wbest_i[w][u];
            if (v > v0[u] || (v == v0[u] && i < i0[u])) { v0[u] = v; i0[u] = i; }
          }
        }
        *cluster.map_shared_rank(&s.best_v[par][rank], tid) = make_float4(v0[0], v0[1], v0[2], v0[3]);
        *cluster.map_shared_rank(&s.best_i[par][rank], tid) = make_int4(i0[0], i0[1], i0[2], i0[3]);
      }
      DBG_T(4);
      cluster.sync();
      DBG_T(5);

      // ---------------- every thread of every CTA replays the same decision
      float4 bvv = s.best_v[par][0];
      int4 bii = s.best_i[par][0];
      for (int rr = 1; rr < kCl; ++rr) {
        const float4 v = s.best_v[par][rr];
        const int4 i = s.best_i[par][rr];
        if (v.x > bvv.x || (v.x == bvv.x && i.x < bii.x)) { bvv.x = v.x; bii.x = i.x; }
        if (v.y > bvv.y || (v.y == bvv.y && i.y < bii.y)) { bvv.y = v.y; bii.y = i.y; }
        if (v.z > bvv.z || (v.z == bvv.z && i.z < bii.z)) { bvv.z = v.z; bii.z = i.z; }
        if (v.w > bvv.w || (v.w == bvv.w && i.w < bii.w)) { bvv.w = v.w; bii.w = i.w; }
      }
      DBG_T(10);
      const int kbest[kNU] = {bii.x, bii.y, bii.z, bii.w};
      bool moved[kNU];
      int emitted_m = 0;
#pragma unroll
      for (int u = 0; u < kNU; ++u) {
        moved[u] = false;
        if (!act[u]) continue;
        const int k = kbest[u];
        if (k == p.blank) {
          t_u[u] += 1;
          nsym[u] = 0;
          need_lstm[u] = false;
          moved[u] = true;
        } else {
          const int ug = group * p.nu + u;
          if (rank == 0 && tid == 0 && cnt[u] < p.max_out) {
            p.ids[static_cast<size_t>(ug) * p.max_out + cnt[u]] = k;
            p.frames[static_cast<size_t>(ug) * p.max_out + cnt[u]] = t_u[u];
          }
          cnt[u] += 1;
          label[u] = k;
          need_lstm[u] = true;   // the state that produced this token becomes the input of the next LSTM step
          emitted_m |= 1 << u;
          nsym[u] += 1;
          if (nsym[u] >= p.max_symbols) { t_u[u] += 1; nsym[u] = 0; moved[u] = true; }
        }
      }
      DBG_T(11);
      // what the next rounds need from global memory, requested now and consumed a phase (or a frame) later
      if (tid < kH) {
        if (moved[0]) { ep.x = epn.x; if (t_u[0] + 1 < L[0]) epn.x = __ldcg(ep_base + static_cast<size_t>(t_u[0] + 1) * kH); }
        if (moved[1]) { ep.y = epn.y; if (t_u[1] + 1 < L[1]) epn.y = __ldcg(ep_base + (static_cast<size_t>(1) * p.T + t_u[1] + 1) * kH); }
        if (moved[2]) { ep.z = epn.z; if (t_u[2] + 1 < L[2]) epn.z = __ldcg(ep_base + (static_cast<size_t>(2) * p.T + t_u[2] + 1) * kH); }
        if (moved[3]) { ep.w = epn.w; if (t_u[3] + 1 < L[3]) epn.w = __ldcg(ep_base + (static_cast<size_t>(3) * p.T + t_u[3] + 1) * kH); }
      }
      if (gate_thread && ((emitted_m >> lq) & 1)) {
        int lab = label[0];
#pragma unroll
        for (int uu = 1; uu < kNU; ++uu)
          if (lq == uu) lab = label[uu];
        eg = __ldcg(p.emb_gates + static_cast<size_t>(lab) * G + eg_off);
      }
      DBG_T(12);
    }
    DBG_T(6);
    if (rank == 0 && tid < kNU) {
      const int ug = group * p.nu + tid;
      int c = cnt[0];
#pragma unroll
      for (int uu = 1; uu < kNU; ++uu)
        if (tid == uu) c = cnt[uu];
      if (tid < p.nu && ug < p.B) p.counts[ug] = min(c, p.max_out);
    }
    cluster.sync();
  }
#ifdef GAM_RNNT_DBG
  if (dbg_on)
    for (int i = 0; i < 16; ++i) g_rnnt_dbg[i] = dbg_acc[i];
#endif
}

}  // namespace

#ifdef GAM_RNNT_DBG
extern "C" int gam_rnnt_debug_read(long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_rnnt_dbg, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
#endif

// returns 0 on success, 1 if a 16-CTA cluster cannot be scheduled on this device (caller falls back to the
// per-utterance kernel), negative on error
int launch_rnnt_greedy_cluster(const float* encproj, const int* len, const float* emb_gates, const float* whhT, const float* wpT,
                               const float* bp, const float* wo, const float* bo, int B, int T, int H, int V1, int blank,
                               int max_symbols, int max_out, int* ids, int* frames, int* counts, cudaStream_t s) {
  if (H != kH) return 1;
  static int max_clusters = -1, smem_cap = 0;
  if (max_clusters < 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  // W_o slice: as many class rows as fit next to the recurrent weights
  const int cls_per = (V1 + kCl - 1) / kCl;
  const int cls_pad = (cls_per + 3) & ~3;
  const int fixed = static_cast<int>(sizeof(Smem)) + cls_pad * 4;
  int rows_smem = (smem_cap - fixed) / (kWoPitch * 4);
  if (rows_smem < 0) return 1;
  if (rows_smem > cls_per) rows_smem = cls_per;
  const int smem = fixed + rows_smem * kWoPitch * 4;
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
  static int smem_set = 0;
  if (max_clusters < 0 || smem > smem_set) {
    if (cudaFuncSetAttribute(rnnt_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(rnnt_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      max_clusters = 0;
    } else {
      smem_set = smem;
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
  p.rows_smem = rows_smem;
  p.cls_pad = cls_pad;
  // spread utterances over as many clusters as can be resident: fewer lock-stepped utterances per cluster
  int nu = (B + max_clusters - 1) / max_clusters;
  if (nu > kNU) nu = kNU;
  p.nu = nu;
  p.num_groups = (B + nu - 1) / nu;
  const int nclusters = p.num_groups < max_clusters ? p.num_groups : max_clusters;
  p.ids = ids; p.frames = frames; p.counts = counts;
  cfg.gridDim = dim3(nclusters * kCl);
  static int info = -1;
  if (info < 0) {
    const char* e = getenv("GAM_RNNT_INFO");
    info = (e && e[0] == '1') ? 1 : 0;
  }
  if (info)
    fprintf(stderr, "[gam] rnnt cluster kernel: %d clusters of %d CTAs resident at most, %d groups of %d utterances on %d clusters, "
                    "%d of %d class rows per CTA in shared memory (%d B)\n", max_clusters, kCl, p.num_groups, nu, nclusters, rows_smem,
            cls_per, smem);
  if (cudaLaunchKernelEx(&cfg, rnnt_cluster_kernel, p) != cudaSuccess) return -2;
  return 0;
}

}  // namespace gam
