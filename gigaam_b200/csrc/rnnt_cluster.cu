// RNN-T greedy decode, cluster-resident variant (gigaam/decoding.py:128-207, gigaam/decoder.py:24-102).
//
// The serial recurrence per utterance (see rnnt.cu) is latency bound; what made the one-CTA-per-utterance kernel
// slow was streaming W_hh (1.6 MB fp32) and W_p through one SM's L2 port on every emission.  Here a thread-block
// cluster of 16 CTAs keeps the weights RESIDENT in distributed shared memory: CTA c owns hidden units
// [c*H/16, (c+1)*H/16) -> its 4 gate rows of W_hh (80 x 320 fp32), its rows of W_p (20 x 320) and its slice of the
// output classes (rows of W_o: as many as fit next to the recurrent weights, the rest is prefetched from L2 into
// registers at the top of every joint phase).  A cluster decodes a group of up to 8 utterances in lock-step
// (utterances are independent; every CTA derives the same control flow from the same exchanged argmax results):
//   LSTM phase  (only utterances that just emitted): own gate rows . h  -> c', h' slice -> DSMEM all-to-all
//   pred phase  : own rows of W_p . h'                                  -> DSMEM all-to-all
//   joint phase : hid = relu(W_e e_t + b_e + pg);  own slice of classes -> local (max, argmax) -> DSMEM all-to-all
//
// What measuring the earlier versions taught (tools/rnnt_phase_probe.py, profiles/r1e_rnnt_phases.md):
//  * cg::cluster.sync() compiles to MEMBAR.ALL.GPU + ERRBAR + cluster barrier + CCTL.IVALL (~1 000 cycles, and it
//    drains every prefetch).  The round loop therefore has NO cluster barrier: every exchange is a set of 16-byte
//    st.async stores whose arrival is counted on an mbarrier of the RECEIVING CTA (two barriers per exchange type,
//    alternating, so bytes of consecutive exchanges can never mix); a CTA waits only on its own barriers.
//  * scalar control flow replicated in 16 warps costs 4x its single-warp time (4 warps per scheduler): the per-utterance
//    state lives in shared memory and ONE warp (one lane per utterance) takes the decision for the CTA.
//  * a B200 can keep only 7 clusters of 16 CTAs resident, so 32 utterances in groups of 4 need two passes: groups hold
//    up to 8 utterances (two float4 halves), all state vectors are utterance-interleaved ([half][unit] -> float4), a
//    weight is read from shared memory once per phase for all of them, and exchanges are 16-byte stores.
//  * the prediction-network state is double-buffered by a cluster-wide parity that flips on every LSTM round;
//    utterances that do not step in that round carry their state over inside the same float4.
// All arithmetic is fp32 as in the reference head.
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace gam {
namespace {

constexpr int kCl = 16;           // CTAs per cluster
constexpr int kMaxU = 8;          // utterances decoded in lock-step per cluster (two float4 halves)
constexpr int kH = 320;
constexpr int kHS = kH / kCl;     // hidden units owned per CTA (20)
constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kWhhP = 324;        // W_hh row pitch (words), 4 mod 32: (8 rows x 4 k-lanes) warps read conflict-free
constexpr int kWpP = 336;         // W_p row pitch, 16 mod 32: (2 rows x 16 k-lanes) warps read conflict-free
constexpr int kWoPitch = kH + 1;  // W_o row pitch: conflict-free 4-byte row walks
constexpr int kCB = 4;            // classes accumulated together per warp (one hid read serves all of them)
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
  int B, T, V1, blank, max_symbols, max_out, num_groups, nu;   // nu <= 4 * NH utterances per group
  int rows_smem;           // class rows of W_o resident in shared memory per CTA
  int cls_pad;             // floats reserved for the bias slice
  int* ids;
  int* frames;
  int* counts;
};

// per-utterance decoding state (gigaam/decoding.py:150-205), owned by warp 0 of every CTA (identical in all of them)
struct Ctl {
  int t[kMaxU], nsym[kMaxU], cnt[kMaxU], label[kMaxU], L[kMaxU], need[kMaxU];
  int act_m, run_m, moved_m, emit_m;   // bit u: still decoding / needs an LSTM step / frame advanced / emitted a token
};

template <int NH>
struct Smem {
  float whh[4 * kHS][kWhhP];      // rows: gate g, unit j  ->  g*kHS + j
  float wp[kHS][kWpP];
  float4 h4[2][NH][kH];           // prediction-network state h, [parity][half][unit] -> 4 utterances (replicated per CTA)
  float4 pg4[NH][kH];             // W_p h' + b_p
  float4 hid4[NH][kH];            // relu(enc_proj[t] + pg)
  float4 hnew4[NH][kHS];          // own slice of the next state, staged for the 16-byte all-to-all
  float c[2][4 * NH][kHS];        // cell state of the own units, same parity as h4
  float gates[4 * NH][4 * kHS];
  float4 best_v[2][kCl][NH];      // per-CTA partial argmax, [exchange parity][source CTA]
  int4 best_i[2][kCl][NH];
  float wbest_v[kWarps][4 * NH];
  int wbest_i[kWarps][4 * NH];
  float4 my_v[NH];
  int4 my_i[NH];
  Ctl ctl;
  uint64_t bar_h[2], bar_pg[2], bar_best[2];   // arrival of the three all-to-all exchanges (alternating pairs)
  // followed by: float bo[cls_pad]; float wo[rows_smem][kWoPitch];
};

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float comp(const float4& v, int u) { return u == 0 ? v.x : (u == 1 ? v.y : (u == 2 ? v.z : v.w)); }
__device__ __forceinline__ void set_comp(float4& v, int u, float x) {
  if (u == 0) v.x = x; else if (u == 1) v.y = x; else if (u == 2) v.z = x; else v.w = x;
}

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

// 16-byte counted store to the same shared-memory location (and barrier) in CTA `cta` of the cluster
__device__ __forceinline__ void push16(const void* local_dst, uint64_t* local_bar, uint32_t cta, uint32_t a, uint32_t b, uint32_t c,
                                       uint32_t d) {
  ptx::st_async_v4(ptx::mapa_u32(ptx::smem_u32(local_dst), cta), ptx::mapa_u32(ptx::smem_u32(local_bar), cta), a, b, c, d);
}
__device__ __forceinline__ void push16(const void* local_dst, uint64_t* local_bar, uint32_t cta, const float4& v) {
  push16(local_dst, local_bar, cta, __float_as_uint(v.x), __float_as_uint(v.y), __float_as_uint(v.z), __float_as_uint(v.w));
}

#ifdef GAM_RNNT_DBG
// phase timing of cluster 0 / CTA 0 / thread 0 (tools/rnnt_phase_probe.py; never compiled into the shipped library)
__device__ long long g_rnnt_dbg[16];
#define DBG_T(i) do { if (dbg_on) { const long long t_now = clock64(); dbg_acc[i] += t_now - dbg_last; dbg_last = t_now; } } while (0)
#else
#define DBG_T(i) do { } while (0)
#endif

// NH: float4 halves of utterances per group (4 or 8 utterances).  GLOB: some class rows stay in L2 (large vocabularies);
// compiled out otherwise (the round loop is executed once per step by every warp; 16 KB less code).
template <int NH, bool GLOB>
__global__ void __launch_bounds__(kThreads, 1) rnnt_cluster_kernel(const RnntClParams p) {
  constexpr int NU = 4 * NH;
  using SM = Smem<NH>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SM& s = *reinterpret_cast<SM*>(smem_raw);
  float* s_bo = reinterpret_cast<float*>(smem_raw + sizeof(SM));
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
  // LSTM-phase role of this thread: gate row lr_row, k-lane lq; lane q of a quad also finishes utterances q, q+4
  const int lq = lane & 3, lr_row = warp * 8 + (lane >> 2);
  const bool gate_thread = warp < 4 * kHS / 8;
  const size_t eg_off = gate_thread ? static_cast<size_t>((lr_row / kHS) * kH + rank * kHS + lr_row % kHS) : 0;
  // pred-phase role: row pj, k-lane pq (which is also the CTA this lane serves in the all-to-all)
  const int pq = lane & 15, pj = warp * 2 + (lane >> 4);
  const bool pred_thread = warp < kHS / 2;
  const float my_bp = pred_thread ? __ldg(p.bp + rank * kHS + pj) : 0.f;
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s.bar_h[i], 1);
      ptx::mbar_init(&s.bar_pg[i], 1);
      ptx::mbar_init(&s.bar_best[i], 1);
    }
    ptx::fence_mbar_init();
  }
  __syncthreads();
  uint32_t n_h = 0, n_pg = 0, n_b = 0;   // exchanges done so far (barrier = n & 1, phase parity = (n >> 1) & 1)
  constexpr uint32_t kStateBytes = kCl * kHS * NH * 16;
  constexpr uint32_t kBestBytes = kCl * NH * 32;

  for (int group = cluster_id; group < p.num_groups; group += num_clusters) {
    // ---- control state (warp 0: lane u = utterance u)
    if (warp == 0) {
      int Lu = 0;
      if (lane < kMaxU) {
        const int ug = group * p.nu + lane;
        Lu = (lane < p.nu && lane < NU && ug < p.B) ? min(max(p.len[ug], 0), p.T) : 0;
        s.ctl.t[lane] = 0; s.ctl.nsym[lane] = 0; s.ctl.cnt[lane] = 0; s.ctl.label[lane] = p.blank;
        s.ctl.L[lane] = Lu; s.ctl.need[lane] = Lu > 0;
      }
      const unsigned am = __ballot_sync(0xffffffffu, Lu > 0);
      if (lane == 0) { s.ctl.act_m = static_cast<int>(am); s.ctl.run_m = static_cast<int>(am); s.ctl.moved_m = 0; s.ctl.emit_m = 0; }
    }
    for (int i = tid; i < 2 * NH * kH; i += kThreads) (&s.h4[0][0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < NH * kH; i += kThreads) (&s.pg4[0][0])[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int i = tid; i < 2 * NU * kHS; i += kThreads) (&s.c[0][0][0])[i] = 0.f;
    __syncthreads();
    // encoder projection of the current frame (ep) and of the next one (epn): thread k < H keeps all utterances'
    // values in registers; a frame advance promotes epn and requests the frame after it
    float4 ep[NH], epn[NH];
    const float* ep_base = p.encproj + static_cast<size_t>(group * p.nu) * p.T * kH + (tid < kH ? tid : 0);
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) {
      ep[hh] = make_float4(0.f, 0.f, 0.f, 0.f);
      epn[hh] = ep[hh];
      if (tid < kH) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int u = 4 * hh + cc;
          const int Lu = s.ctl.L[u];
          if (Lu > 0) set_comp(ep[hh], cc, __ldg(ep_base + static_cast<size_t>(u) * p.T * kH));
          if (Lu > 1) set_comp(epn[hh], cc, __ldg(ep_base + (static_cast<size_t>(u) * p.T + 1) * kH));
        }
      }
    }
    // embedding contribution to this thread's gates (row lr_row, utterances lq + 4 hh); reloaded after an emission
    float eg[NH];
#pragma unroll
    for (int hh = 0; hh < NH; ++hh) eg[hh] = gate_thread ? __ldg(p.emb_gates + static_cast<size_t>(p.blank) * G + eg_off) : 0.f;
    int gb = 0;   // parity of the h4 / c buffer that holds the current prediction-network state of all utterances
    cluster.sync();   // every CTA's buffers and barriers are initialised before the first remote store can arrive
    DBG_T(6);

    while (true) {
      const int act_m = s.ctl.act_m, run_m = s.ctl.run_m;
      if (act_m == 0) break;
      DBG_T(9);
#ifdef GAM_RNNT_DBG
      dbg_acc[7] += 1;
      if (run_m != 0) dbg_acc[8] += 1;
#endif

      if (run_m != 0) {
        if (tid == 0) ptx::mbar_arrive_expect_tx(&s.bar_h[n_h & 1], kStateBytes);
        // ---------------- gates: warp = 8 own rows x 4 k-lanes (k = 16 i + 4 e + q)
        if (gate_thread) {
          const float* wrow = &s.whh[lr_row][lq];
          float4 a[NH];
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) a[hh] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
          for (int i = 0; i < kH / 16; ++i) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float w = wrow[16 * i + 4 * e];
#pragma unroll
              for (int hh = 0; hh < NH; ++hh) fma4(a[hh], w, s.h4[gb][hh][16 * i + 4 * e + lq]);
            }
          }
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) {
#pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
              a[hh].x += __shfl_xor_sync(0xffffffffu, a[hh].x, o); a[hh].y += __shfl_xor_sync(0xffffffffu, a[hh].y, o);
              a[hh].z += __shfl_xor_sync(0xffffffffu, a[hh].z, o); a[hh].w += __shfl_xor_sync(0xffffffffu, a[hh].w, o);
            }
            s.gates[4 * hh + lq][lr_row] = eg[hh] + comp(a[hh], lq);   // lane q of the quad finishes utterance 4 hh + q
          }
        }
        __syncthreads();
        DBG_T(0);
        if (tid < NU * kHS) {
          const int u = tid / kHS, j = tid % kHS;
          const float c_old = s.c[gb][u][j];
          float cn = c_old, hn = comp(s.h4[gb][u >> 2][rank * kHS + j], u & 3);   // utterances that do not step carry over
          if ((run_m >> u) & 1) {
            const float ig = sigm(s.gates[u][j]), fg = sigm(s.gates[u][kHS + j]);
            const float gg = tanhf(s.gates[u][2 * kHS + j]), og = sigm(s.gates[u][3 * kHS + j]);
            cn = fg * c_old + ig * gg;
            hn = og * tanhf(cn);
          }
          s.c[gb ^ 1][u][j] = cn;
          reinterpret_cast<float*>(&s.hnew4[u >> 2][j])[u & 3] = hn;
        }
        __syncthreads();
        for (int i = tid; i < kCl * NH * kHS; i += kThreads) {   // 16-byte counted stores: (CTA, half, unit)
          const int cta = i / (NH * kHS), hh = (i / kHS) % NH, j = i % kHS;
          push16(&s.h4[gb ^ 1][hh][rank * kHS + j], &s.bar_h[n_h & 1], cta, s.hnew4[hh][j]);
        }
        DBG_T(1);
        if (pred_thread) ptx::mbar_wait(&s.bar_h[n_h & 1], (n_h >> 1) & 1);
        ++n_h;
        DBG_T(2);
        // ---------------- prediction projection: own rows of W_p on the new state (warp = 2 rows x 16 k-lanes)
        if (tid == 0) ptx::mbar_arrive_expect_tx(&s.bar_pg[n_pg & 1], kStateBytes);
        if (pred_thread) {
          const float* wrow = &s.wp[pj][pq];
          float4 a[NH];
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) a[hh] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < kH / 64; ++i) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float w = wrow[64 * i + 16 * e];
#pragma unroll
              for (int hh = 0; hh < NH; ++hh) fma4(a[hh], w, s.h4[gb ^ 1][hh][64 * i + 16 * e + pq]);
            }
          }
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) {
#pragma unroll
            for (int o = 1; o <= 8; o <<= 1) {
              a[hh].x += __shfl_xor_sync(0xffffffffu, a[hh].x, o); a[hh].y += __shfl_xor_sync(0xffffffffu, a[hh].y, o);
              a[hh].z += __shfl_xor_sync(0xffffffffu, a[hh].z, o); a[hh].w += __shfl_xor_sync(0xffffffffu, a[hh].w, o);
            }
            const float4 old = s.pg4[hh][rank * kHS + pj];
            const int rm = run_m >> (4 * hh);
            const float4 v = make_float4((rm & 1) ? my_bp + a[hh].x : old.x, (rm & 2) ? my_bp + a[hh].y : old.y,
                                         (rm & 4) ? my_bp + a[hh].z : old.z, (rm & 8) ? my_bp + a[hh].w : old.w);
            push16(&s.pg4[hh][rank * kHS + pj], &s.bar_pg[n_pg & 1], pq, v);   // lane q of the row's 16 serves CTA q
          }
        }
        DBG_T(3);
        if (pred_thread) ptx::mbar_wait(&s.bar_pg[n_pg & 1], (n_pg >> 1) & 1);   // warps 0-9 also build hid4 below
        ++n_pg;
        gb ^= 1;
        DBG_T(2);
      }

      // ---------------- joint: hid = relu(enc_proj[t] + pg), own class slice, local argmax
      if (tid == 0) ptx::mbar_arrive_expect_tx(&s.bar_best[n_b & 1], kBestBytes);
      // class rows that do not fit in shared memory: issue their L2 loads now, consume them after the smem rows
      [[maybe_unused]] float wg[kGP][kH / 32];
      [[maybe_unused]] int gcls[kGP];
      if constexpr (GLOB) {
#pragma unroll
        for (int gi = 0; gi < kGP; ++gi) {
          const int lr = nsm + ((warp - nsm) & (kWarps - 1)) + kWarps * gi;   // this warp's gi-th row at or after nsm
          gcls[gi] = lr < ncls ? lr : -1;
          if (gcls[gi] >= 0) {
            const float* w = p.wo + static_cast<size_t>(cls0 + lr) * kH;
#pragma unroll
            for (int kk = 0; kk < kH / 32; ++kk) wg[gi][kk] = __ldg(w + lane + 32 * kk);
          }
        }
      }
      if (tid < kH) {
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          const float4 g4 = s.pg4[hh][tid];
          const int am = act_m >> (4 * hh);
          s.hid4[hh][tid] = make_float4((am & 1) ? fmaxf(ep[hh].x + g4.x, 0.f) : 0.f, (am & 2) ? fmaxf(ep[hh].y + g4.y, 0.f) : 0.f,
                                        (am & 4) ? fmaxf(ep[hh].z + g4.z, 0.f) : 0.f, (am & 8) ? fmaxf(ep[hh].w + g4.w, 0.f) : 0.f);
        }
      }
      __syncthreads();
      DBG_T(13);
      const int myu = (lane >> 3) & 3;       // the utterance (within a half) whose logits this lane ends up holding
      float bv[NH];
      int bi[NH];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) { bv[hh] = -INFINITY; bi[hh] = 0x7fffffff; }
      // shared-memory rows: local rows warp, warp+16, ... (ascending, so the first maximum wins as in torch.argmax)
      for (int lr0 = warp; lr0 < nsm; lr0 += kWarps * kCB) {
        float4 acc[kCB][NH];
#pragma unroll
        for (int c = 0; c < kCB; ++c)
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) acc[c][hh] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int kk = 0; kk < kH / 32; ++kk) {
          const int k = lane + 32 * kk;
          float4 hv[NH];
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) hv[hh] = s.hid4[hh][k];
#pragma unroll
          for (int c = 0; c < kCB; ++c)
            if (lr0 + kWarps * c < nsm) {
              const float w = s_wo[(lr0 + kWarps * c) * kWoPitch + k];
#pragma unroll
              for (int hh = 0; hh < NH; ++hh) fma4(acc[c][hh], w, hv[hh]);
            }
        }
#pragma unroll
        for (int c = 0; c < kCB; ++c) {
          const int lr = lr0 + kWarps * c;
          if (lr < nsm) {
#pragma unroll
            for (int hh = 0; hh < NH; ++hh) {
              const float a = reduce4(acc[c][hh], lane) + s_bo[lr];
              if (a > bv[hh]) { bv[hh] = a; bi[hh] = cls0 + lr; }
            }
          }
        }
      }
      // L2 rows (registers), then anything beyond the prefetch depth straight from L2
      if constexpr (GLOB) {
#pragma unroll
      for (int gi = 0; gi < kGP; ++gi) {
        if (gcls[gi] >= 0) {
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int kk = 0; kk < kH / 32; ++kk) fma4(acc, wg[gi][kk], s.hid4[hh][lane + 32 * kk]);
            const float a = reduce4(acc, lane) + s_bo[gcls[gi]];
            if (a > bv[hh]) { bv[hh] = a; bi[hh] = cls0 + gcls[gi]; }
          }
        }
      }
      for (int lr = nsm + ((warp - nsm) & (kWarps - 1)) + kWarps * kGP; lr < ncls; lr += kWarps) {
        const float* w = p.wo + static_cast<size_t>(cls0 + lr) * kH;
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int kk = 0; kk < kH / 32; ++kk) fma4(acc, __ldg(w + lane + 32 * kk), s.hid4[hh][lane + 32 * kk]);
          const float a = reduce4(acc, lane) + s_bo[lr];
          if (a > bv[hh]) { bv[hh] = a; bi[hh] = cls0 + lr; }
        }
      }
      }
      DBG_T(14);
      if ((lane & 7) == 0) {
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) { s.wbest_v[warp][4 * hh + myu] = bv[hh]; s.wbest_i[warp][4 * hh + myu] = bi[hh]; }
      }
      __syncthreads();
      DBG_T(15);

      // ---------------- warp 0: CTA argmax -> all-to-all -> decision (lane = utterance u + NU * sub-lane)
      if (warp == 0) {
        constexpr int kSub = 32 / NU;
        const int u = lane % NU, sub = lane / NU;
        const int par = n_b & 1;
        float v0 = -INFINITY;
        int i0 = 0x7fffffff;
        for (int w = sub; w < kWarps; w += kSub) {
          const float v = s.wbest_v[w][u];
          const int i = s.wbest_i[w][u];
          if (v > v0 || (v == v0 && i < i0)) { v0 = v; i0 = i; }
        }
#pragma unroll
        for (int o = NU; o < 32; o <<= 1) {
          const float v = __shfl_xor_sync(0xffffffffu, v0, o);
          const int i = __shfl_xor_sync(0xffffffffu, i0, o);
          if (v > v0 || (v == v0 && i < i0)) { v0 = v; i0 = i; }
        }
        if (lane < NU) { reinterpret_cast<float*>(s.my_v)[u] = v0; reinterpret_cast<int*>(s.my_i)[u] = i0; }
        __syncwarp();
        if (lane < kCl) {   // lane = destination CTA
#pragma unroll
          for (int hh = 0; hh < NH; ++hh) {
            const int4 ii = s.my_i[hh];
            push16(&s.best_v[par][rank][hh], &s.bar_best[par], lane, s.my_v[hh]);
            push16(&s.best_i[par][rank][hh], &s.bar_best[par], lane, static_cast<uint32_t>(ii.x), static_cast<uint32_t>(ii.y),
                   static_cast<uint32_t>(ii.z), static_cast<uint32_t>(ii.w));
          }
        }
        DBG_T(4);
        ptx::mbar_wait(&s.bar_best[par], (n_b >> 1) & 1);
        DBG_T(5);
        v0 = -INFINITY;
        i0 = 0x7fffffff;
        for (int r = sub; r < kCl; r += kSub) {   // source CTAs ascend with the class index
          const float v = reinterpret_cast<const float*>(&s.best_v[par][r][0])[u];
          const int i = reinterpret_cast<const int*>(&s.best_i[par][r][0])[u];
          if (v > v0 || (v == v0 && i < i0)) { v0 = v; i0 = i; }
        }
#pragma unroll
        for (int o = NU; o < 32; o <<= 1) {
          const float v = __shfl_xor_sync(0xffffffffu, v0, o);
          const int i = __shfl_xor_sync(0xffffffffu, i0, o);
          if (v > v0 || (v == v0 && i < i0)) { v0 = v; i0 = i; }
        }
        DBG_T(10);
        // lane u < NU decides for utterance u (gigaam/decoding.py:176-205)
        bool act_new = false, run_new = false, moved = false, emitted = false;
        if (lane < NU) {
          int t = s.ctl.t[u];
          const int Lu = s.ctl.L[u];
          int need = s.ctl.need[u];
          if (t < Lu) {
            if (i0 == p.blank) {
              t += 1;
              s.ctl.nsym[u] = 0;
              need = 0;
              moved = true;
            } else {
              const int cnt = s.ctl.cnt[u];
              if (rank == 0 && cnt < p.max_out) {
                const size_t o = static_cast<size_t>(group * p.nu + u) * p.max_out + cnt;
                p.ids[o] = i0;
                p.frames[o] = t;
              }
              s.ctl.cnt[u] = cnt + 1;
              s.ctl.label[u] = i0;
              need = 1;   // the state that produced this token becomes the input of the next LSTM step
              emitted = true;
              int ns = s.ctl.nsym[u] + 1;
              if (ns >= p.max_symbols) { t += 1; ns = 0; moved = true; }
              s.ctl.nsym[u] = ns;
            }
            s.ctl.t[u] = t;
            s.ctl.need[u] = need;
          }
          act_new = t < Lu;
          run_new = act_new && need != 0;
        }
        const unsigned am = __ballot_sync(0xffffffffu, act_new), rm = __ballot_sync(0xffffffffu, run_new);
        const unsigned mm = __ballot_sync(0xffffffffu, moved), em = __ballot_sync(0xffffffffu, emitted);
        if (lane == 0) {
          s.ctl.act_m = static_cast<int>(am); s.ctl.run_m = static_cast<int>(rm);
          s.ctl.moved_m = static_cast<int>(mm); s.ctl.emit_m = static_cast<int>(em);
        }
        DBG_T(11);
      }
      ++n_b;
      __syncthreads();
      // what the next rounds need from global memory, requested now and consumed a phase (or a frame) later
      const int moved_m = s.ctl.moved_m, emit_m = s.ctl.emit_m;
      if (tid < kH && moved_m != 0) {
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const int u = 4 * hh + cc;
            if ((moved_m >> u) & 1) {
              set_comp(ep[hh], cc, comp(epn[hh], cc));
              const int t1 = s.ctl.t[u] + 1;
              if (t1 < s.ctl.L[u]) set_comp(epn[hh], cc, __ldg(ep_base + (static_cast<size_t>(u) * p.T + t1) * kH));
            }
          }
        }
      }
      if (gate_thread && emit_m != 0) {
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          const int u = 4 * hh + lq;
          if ((emit_m >> u) & 1) eg[hh] = __ldg(p.emb_gates + static_cast<size_t>(s.ctl.label[u]) * G + eg_off);
        }
      }
      DBG_T(12);
    }
    if (rank == 0 && warp == 0 && lane < NU) {
      const int ug = group * p.nu + lane;
      if (lane < p.nu && ug < p.B) p.counts[ug] = min(s.ctl.cnt[lane], p.max_out);
    }
    __syncthreads();   // ctl is re-initialised by warp 0 at the top of the next group
  }
  cluster.sync();      // no CTA may exit while a peer can still store into its shared memory
#ifdef GAM_RNNT_DBG
  if (dbg_on)
    for (int i = 0; i < 16; ++i) g_rnnt_dbg[i] = dbg_acc[i];
#endif
}

struct LaunchState {
  int max_clusters = -1;
  int smem_set = 0;
};

template <int NH, bool GLOB>
int launch_nh(RnntClParams& p, int B, int V1, int smem_cap, bool info, cudaStream_t s) {
  static LaunchState per_device[64];   // function attributes and cluster occupancy are per device
  int dev_index = 0;
  cudaGetDevice(&dev_index);
  LaunchState& st = per_device[dev_index & 63];
  const int cls_per = (V1 + kCl - 1) / kCl;
  const int cls_pad = (cls_per + 3) & ~3;
  const int fixed = static_cast<int>(sizeof(Smem<NH>)) + cls_pad * 4;
  int rows_smem = (smem_cap - fixed) / (kWoPitch * 4);
  if (rows_smem < 0) return 1;
  if (rows_smem > cls_per) rows_smem = cls_per;
  if (!GLOB && rows_smem < cls_per) return 1;   // caller picked the wrong variant
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
  if (st.max_clusters < 0 || smem > st.smem_set) {
    if (cudaFuncSetAttribute(rnnt_cluster_kernel<NH, GLOB>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
        cudaFuncSetAttribute(rnnt_cluster_kernel<NH, GLOB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) {
      cudaGetLastError();
      st.max_clusters = 0;
    } else {
      st.smem_set = smem;
      cfg.gridDim = dim3(kCl);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, rnnt_cluster_kernel<NH, GLOB>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
      st.max_clusters = n;
    }
  }
  if (st.max_clusters <= 0) return 1;
  // spread utterances over as many clusters as can be resident: fewer lock-stepped utterances per cluster
  int nu = (B + st.max_clusters - 1) / st.max_clusters;
  if (nu > 4 * NH) nu = 4 * NH;
  p.nu = nu;
  p.num_groups = (B + nu - 1) / nu;
  p.rows_smem = rows_smem;
  p.cls_pad = cls_pad;
  const int nclusters = p.num_groups < st.max_clusters ? p.num_groups : st.max_clusters;
  cfg.gridDim = dim3(nclusters * kCl);
  if (info)
    fprintf(stderr, "[gam] rnnt cluster kernel<%d,%d>: at most %d clusters of %d CTAs resident; %d groups of %d utterances on %d clusters; "
                    "%d of %d class rows per CTA in shared memory (%d B)\n", NH, GLOB ? 1 : 0, st.max_clusters, kCl, p.num_groups, nu, nclusters,
            rows_smem, cls_per, smem);
  if (cudaLaunchKernelEx(&cfg, rnnt_cluster_kernel<NH, GLOB>, p) != cudaSuccess) return -2;
  return 0;
}

}  // namespace

#ifdef GAM_RNNT_DBG
extern "C" int gam_rnnt_debug_read(long long* out16) {
  return cudaMemcpyFromSymbol(out16, g_rnnt_dbg, sizeof(long long) * 16) == cudaSuccess ? 0 : -1;
}
#endif

// returns 0 on success, 1 if the shape is unsupported (pred_hidden != 320) or a 16-CTA cluster cannot be scheduled on
// this device, negative on a launch error
int launch_rnnt_greedy_cluster(const float* encproj, const int* len, const float* emb_gates, const float* whhT, const float* wpT,
                               const float* bp, const float* wo, const float* bo, int B, int T, int H, int V1, int blank,
                               int max_symbols, int max_out, int* ids, int* frames, int* counts, cudaStream_t s) {
  if (H != kH) return 1;
  static int smem_cap = 0, clusters_hint = 0;
  constexpr int info = 0;
  if (smem_cap == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&smem_cap, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    clusters_hint = sms / kCl - 2;   // GPCs rarely hold more than one 16-CTA cluster each (7 on a 148-SM B200)
    if (clusters_hint < 1) clusters_hint = 1;
  }
  RnntClParams p;
  p.encproj = encproj; p.len = len; p.emb_gates = emb_gates; p.whhT = whhT; p.wpT = wpT; p.bp = bp; p.wo = wo; p.bo = bo;
  p.B = B; p.T = T; p.V1 = V1; p.blank = blank; p.max_symbols = max_symbols; p.max_out = max_out;
  p.ids = ids; p.frames = frames; p.counts = counts;
  // groups of up to 4 utterances while every group still gets its own cluster, else groups of up to 8
  const int cls_per = (V1 + kCl - 1) / kCl;
  const bool small = B <= 4 * clusters_hint;
  const int fixed = static_cast<int>(small ? sizeof(Smem<1>) : sizeof(Smem<2>)) + ((cls_per + 3) & ~3) * 4;
  const bool glob = (smem_cap - fixed) / (kWoPitch * 4) < cls_per;   // some class rows have to stay in L2
  if (small) return glob ? launch_nh<1, true>(p, B, V1, smem_cap, info != 0, s) : launch_nh<1, false>(p, B, V1, smem_cap, info != 0, s);
  return glob ? launch_nh<2, true>(p, B, V1, smem_cap, info != 0, s) : launch_nh<2, false>(p, B, V1, smem_cap, info != 0, s);
}

}  // namespace gam
