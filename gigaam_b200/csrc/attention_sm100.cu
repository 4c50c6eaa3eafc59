// tcgen05 attention for the Conformer block (head_dim 48, T' <= 768, key-padding mask by length).
// Replaces F.scaled_dot_product_attention / flash_attn_varlen in
// gigaam/encoder.py:258-277 (RotaryPositionMultiHeadAttention) on the q/k/v produced by the fused
// LN+RoPE -> GEMM kernels.
//
//   qkv : [rows, 2304] fp16 = [q(768) | k(768) | v(768)], head h at columns h*48 .. h*48+47 of each part
//   out : [rows, 768]  fp16
//
// Rows are PACKED (the varlen contract of gigaam/utils.py:103-155, apply_masked_flash_attn): utterance b owns rows
// cu[b] .. cu[b] + klen[b]; only the query tiles and key blocks that hold one of its frames are computed, a tile that
// reaches past the utterance reads its neighbour's rows (masked as keys, never stored as queries).  With cu == null the
// layout is the padded [B, T] one (unit tests): row b*T, every query row stored.  The rows of the last key block past
// klen are zeroed in shared memory before P.V: their P is 0, but 0 x (stale inf / NaN bits) would not be.
//
// ONE kernel, one sweep: every score is read out of tensor memory once.  A softmax thread (one per query row) pulls its
// 128 scores of a key block into registers with four tcgen05.ld, releases the block at once (the tensor core computes
// the next S while this one is exponentiated), and runs an online softmax on the registers:
//   * LAZY reference point: exact block maximum on the first block; afterwards it moves only when a block maximum exceeds
//     it by more than 2^8 (P then stays <= 256, far inside fp16).  Moving it rescales the O accumulator in tensor memory
//     (after the previous P.V has retired); on real score rows that is rare;
//   * the softmax denominator comes from the tensor core: column dk of every V tile is set to 1.0 in shared memory, so
//     O[:, dk] accumulates the row sum of exactly the fp16 P that P.V used (P.V multiplies all 64 columns of the tile
//     anyway) and follows every rescaling of O by construction;
//   * P goes back to tensor memory as packed fp16 (tcgen05.st) into its own 64 columns and O += P V runs as a TS-mode
//     tcgen05.mma with V straight from its [key, d] layout (MN-major B).  S, P and O have separate columns, so S of
//     block k+1, the softmax of block k and P.V of block k-1 are in flight together.
// K / V blocks are 128 keys x 64 columns (SWIZZLE_128B; the 16 columns past the 48 real ones belong to the next head:
// QK^T runs K = 3 x 16 and never touches them; of the 16 extra output columns of P.V, column dk is the row sum and the
// other 15 are dropped).
//
// Persistent CTA per SM over (utterance, head) items; two query-tile streams per CTA (a softmax warpgroup + an MMA
// issuer warp + 256 TMEM columns each: S [0,128) | P [128,192) | O [192,256)) take alternate query tiles of the item;
// a fourth warpgroup turns finished O tiles into output rows for both streams.
// The item's K / V blocks are resident in shared memory and shared by both streams; with up to 3 key blocks (T' <= 384)
// a second set of K / V buffers lets the producer fetch the next item while this one is computed.  Output rows leave
// through a per-warp staging tile as runs of whole rows (see the epilogue).
//
// Measured (tools/attn_probe.py, profiles/r2n_attn_probe.txt), one launch with a flushed L2: 47 us at the c2 shape
// (64 x 251; the two-sweep kernels this replaces: 56), 55 us at c3 (32 x 376; before: 89), 131 us at 32 x 626 (before: 191).
// Tried and dropped: walking a block in 32-score chunks with the next tcgen05.ld in flight (slower: per-chunk reference
// logic), and moving a quarter / a half of the exponentials to the FMA pipe (Cody-Waite + cubic: no change) -- the softmax
// warps are bound by the latency of their serial block chain (ld -> max -> exp -> st), not by MUFU or issue throughput.
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace gam {
namespace {

constexpr int kMaxKB = 6;              // up to 768 keys (30 s segments of the reference's VAD, gigaam/vad_utils.py:85)
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 fp16
constexpr int kThreads = 512;          // warp 0 TMA, 1 / 2 MMA issuers, 3 TMEM owner, 4-7 / 8-11 softmax warpgroups, 12-15 output
constexpr uint32_t kPCol = 128, kOCol = 192;
constexpr float kLazyLog2 = 8.0f;      // the softmax reference point trails the running maximum by at most 2^8
constexpr int kBarBytes = 512;         // mbarriers + TMEM base slot
constexpr int kStagePitch = 112;       // bytes per staged output row (96 used): conflict-free 16-byte accesses
constexpr int kStageBytes = 4 * 32 * kStagePitch;   // one 32-row staging tile per output warp

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// geometry of one (utterance, head) item; identical in every warp role
struct ItemGeom {
  int row0;   // first row of the utterance
  int klen;   // valid keys
  int qlim;   // query rows that are stored
  int nq;     // query tiles that are computed (0: the item is skipped by every role)
  int nk;     // key blocks that are multiplied (>= 1)
};
__device__ __forceinline__ ItemGeom item_geom(const int* klen_p, const int* cu, int b, int T, int nkb) {
  ItemGeom g;
  g.klen = klen_p != nullptr ? min(max(__ldg(klen_p + b), 0), T) : T;
  const int kb = (g.klen + 127) >> 7;
  g.nk = max(1, kb);
  if (cu != nullptr) {
    g.row0 = __ldg(cu + b);
    g.qlim = g.klen;
    g.nq = kb;
  } else {
    g.row0 = b * T;
    g.qlim = T;
    g.nq = nkb;
  }
  return g;
}

// rows of a key block past klen -> 0 in the V tile (SWIZZLE_128B permutes 16-byte chunks inside a 128-byte row, so
// clearing whole rows needs no address arithmetic); thread = row.  Followed by the generic -> async proxy fence.
__device__ __forceinline__ void zero_v_tail(uint8_t* v_tile, int r, int valid_rows) {
  if (r >= valid_rows) {
    uint4* row = reinterpret_cast<uint4*>(v_tile + r * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) row[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  ptx::fence_proxy_async_smem();
}

struct AttnParams {
  int T, nkb, B, H;
  int ring;          // K / V buffer sets (2 when they fit)
  int stage;         // output rows go through a shared-memory staging tile (whenever it fits)
  const int* klen;
  const int* cu;
  __half* out;
  int ld_out, dk;
  float scale_log2;
};

// Barrier phases.  Every barrier is used as a sequence of completions 0, 1, 2, ...; "wait for completion j" is
// mbar_wait(bar, j & 1), and waiting for completion j - 1 with j = 0 returns at once on a fresh barrier.
//   per K / V set : kv_full[set][kb] (TMA bytes of K and V of block kb), kv_empty[set] (both streams are done with the item)
//   per stream    : q_full / q_empty (Q tile), s_full (S block in TMEM) / s_empty (softmax has it in registers),
//                   p_full (P block in TMEM) / p_empty (its P.V has retired: P free, O current),
//                   o_full (last P.V of a query tile has retired) / o_empty (the output warpgroup has read O out)
__global__ void __launch_bounds__(kThreads, 1) attention_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nkb = p.nkb;
  const int set_bytes = 2 * nkb * kTileBytes;                 // [K blocks | V blocks] of one item
  uint8_t* sQ = smem + p.ring * set_bytes;                    // [2] one per stream
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + 2 * kTileBytes);
  uint64_t* kv_full = bars;                     // [2][kMaxKB]
  uint64_t* kv_empty = bars + 2 * kMaxKB;       // [2]
  uint64_t* q_full = kv_empty + 2;              // [2] ...
  uint64_t* q_empty = q_full + 2;
  uint64_t* s_full = q_full + 4;
  uint64_t* s_empty = q_full + 6;
  uint64_t* p_full = q_full + 8;
  uint64_t* p_empty = q_full + 10;
  uint64_t* o_empty = q_full + 12;
  uint64_t* o_full = q_full + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 16);
  uint8_t* sStage = reinterpret_cast<uint8_t*>(bars) + kBarBytes;

  const int warp_idx = threadIdx.x >> 5;
  const int n_items = p.B * p.H;
  const int dmodel = p.ld_out;

  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmap_qkv);
    for (int i = 0; i < 2 * kMaxKB; ++i) ptx::mbar_init(&kv_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&kv_empty[i], 2);   // one arrival per stream
      ptx::mbar_init(&q_full[i], 1);
      ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_empty[i], 4);    // one arrival per softmax warp
      ptx::mbar_init(&p_full[i], 4);
      ptx::mbar_init(&p_empty[i], 1);
      ptx::mbar_init(&o_empty[i], 4);    // one arrival per output warp
      ptx::mbar_init(&o_full[i], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp_idx == 3) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // 512 threads x 128 registers fill the register file; a softmax thread holds a whole 128-score block, so warps 0-3
  // (TMA / MMA issue / TMEM owner: a few dozen live values) and the output warpgroup hand their surplus to the two
  // softmax warpgroups: 128 x 56 + 128 x 88 + 256 x 184 = 65 536
  // (each role's branch opens with its own setmaxnreg: ptxas budgets a branch by the instruction that dominates it)
  if (warp_idx == 0) {
    // ===================================================== TMA producer
    ptx::setmaxnreg_dec<56>();
    if (ptx::elect_one()) {
      int it = 0;                       // items that use the buffers (an utterance without frames is skipped by every role)
      uint32_t n_q[2] = {0u, 0u};       // Q tiles loaded per stream
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / p.H, h = item % p.H;
        const ItemGeom g = item_geom(p.klen, p.cu, b, p.T, nkb);
        if (g.nq == 0) continue;
        const int set = it % p.ring;
        const uint32_t use = static_cast<uint32_t>(it / p.ring);
        ++it;
        uint8_t* sK = smem + set * set_bytes;
        uint8_t* sV = sK + nkb * kTileBytes;
        ptx::mbar_wait(&kv_empty[set], (use & 1) ^ 1);
        for (int kb = 0; kb < nkb; ++kb) {
          uint64_t* bar = &kv_full[set * kMaxKB + kb];
          if (kb < g.nk) {
            ptx::mbar_arrive_expect_tx(bar, 2 * kTileBytes);
            ptx::tma_load_2d(sK + kb * kTileBytes, &tmap_qkv, bar, dmodel + h * p.dk, g.row0 + kb * 128);
            ptx::tma_load_2d(sV + kb * kTileBytes, &tmap_qkv, bar, 2 * dmodel + h * p.dk, g.row0 + kb * 128);
          } else {
            ptx::mbar_arrive(bar);   // block without a valid key: nothing to load, but every kv_full completes once per use
          }
        }
        for (int qt = 0; qt < g.nq; ++qt) {
          const int wg = qt & 1;
          ptx::mbar_wait(&q_empty[wg], (n_q[wg] & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&q_full[wg], kTileBytes);
          ptx::tma_load_2d(sQ + wg * kTileBytes, &tmap_qkv, &q_full[wg], h * p.dk, g.row0 + qt * 128);
          ++n_q[wg];
        }
      }
    }
  } else if (warp_idx == 1 || warp_idx == 2) {
    // ===================================================== MMA issuer of stream wg
    ptx::setmaxnreg_dec<56>();
    const int wg = warp_idx - 1;
    constexpr uint32_t kIdescS = ptx::make_idesc_f16(128, 128, 0, 0);
    constexpr uint32_t kIdescPV = ptx::make_idesc_f16(128, 64, 0, 1);   // B (= V) MN-major; A (= P) from TMEM
    const int ksteps_qk = p.dk / 16;
    const uint32_t t_s = tmem_base + wg * 256;
    const uint32_t qa = ptx::smem_u32(sQ + wg * kTileBytes);
    uint32_t n_q = 0, n_s = 0, n_p = 0;   // query tiles completed, S blocks and P.V products issued by this stream so far

    // ---- the stream's query tiles in order, across items (items without a tile for this stream are released on the way)
    struct Tile {
      int valid, nk, set, first_q, last_q;
      uint32_t use;
    };
    int it = 0, item = blockIdx.x, qt_next = 0, cur_nq = 0, cur_nk = 0, cur_set = 0;
    uint32_t cur_use = 0;
    const int item_stride = gridDim.x;
    // geometry of the next item not yet opened, fetched one item ahead (two dependent global loads off the issue path)
    ItemGeom g_ahead = item_geom(p.klen, p.cu, min(item, n_items - 1) / p.H, p.T, nkb);
    auto next_tile = [&]() -> Tile {
      for (;;) {
        if (qt_next < cur_nq) {
          Tile t{1, cur_nk, cur_set, qt_next == wg, qt_next + 2 >= cur_nq, cur_use};
          qt_next += 2;
          return t;
        }
        if (item >= n_items) return Tile{0, 0, 0, 0, 0, 0u};
        const ItemGeom g = g_ahead;
        item += item_stride;
        g_ahead = item_geom(p.klen, p.cu, min(item, n_items - 1) / p.H, p.T, nkb);
        cur_nq = 0;
        if (g.nq == 0) continue;
        cur_set = it % p.ring;
        cur_use = static_cast<uint32_t>(it / p.ring);
        ++it;
        if (wg >= g.nq) {
          // no query tile of this item for this stream: release the buffers, but only after the item's first load has
          // completed, so that a stream can never arrive twice within one phase of kv_empty
          ptx::mbar_wait(&kv_full[cur_set * kMaxKB], cur_use & 1);
          if (ptx::elect_one()) ptx::mbar_arrive(&kv_empty[cur_set]);
          __syncwarp();
          continue;
        }
        cur_nq = g.nq;
        cur_nk = g.nk;
        qt_next = wg;
      }
    };
    // S(kb) of tile t: the softmax must have pulled the previous S block into registers
    auto issue_s = [&](const Tile& t, int kb) {
      ptx::mbar_wait(&s_empty[wg], (n_s & 1) ^ 1);
      if (t.first_q) ptx::mbar_wait(&kv_full[t.set * kMaxKB + kb], t.use & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t ka = ptx::smem_u32(smem + t.set * set_bytes + kb * kTileBytes);
        for (int k = 0; k < ksteps_qk; ++k)
          ptx::mma_f16_ss(t_s, ptx::make_smem_desc_sw128(qa + k * 32, 16, 1024), ptx::make_smem_desc_sw128(ka + k * 32, 16, 1024), kIdescS,
                          k != 0 ? 1u : 0u);
        ptx::mma_commit(&s_full[wg]);
        if (kb == t.nk - 1) ptx::mma_commit(&q_empty[wg]);   // last use of this Q tile
      }
      __syncwarp();
      ++n_s;
    };
    // O (+)= P(pb) V(pb) of tile t
    auto issue_pv = [&](const Tile& t, int pb) {
      ptx::mbar_wait(&p_full[wg], n_p & 1);
      if (pb == 0) ptx::mbar_wait(&o_empty[wg], (n_q & 1) ^ 1);   // the previous query tile's O has been read out
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t va = ptx::smem_u32(smem + t.set * set_bytes + (nkb + pb) * kTileBytes);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)
          ptx::mma_f16_ts(t_s + kOCol, t_s + kPCol + ks * 8, ptx::make_smem_desc_sw128(va + ks * 2048, 1024, 1024), kIdescPV,
                          (pb | ks) != 0 ? 1u : 0u);
        ptx::mma_commit(&p_empty[wg]);
        if (pb == t.nk - 1) ptx::mma_commit(&o_full[wg]);                 // the query tile's O is complete
        if (pb == t.nk - 1 && t.last_q) ptx::mma_commit(&kv_empty[t.set]);   // this stream is done with the item
      }
      __syncwarp();
      ++n_p;
    };

    // Issue order per tile: S(0) | S(1) PV(0) | S(2) PV(1) | ... | S'(0) PV(nk-1): every P.V goes out behind the NEXT S, so
    // the tensor core never waits for a softmax -- including across tiles: S'(0) of the stream's next tile is issued ahead of
    // this tile's last P.V (its softmax then finds its scores waiting).  Looking ahead may cross into the NEXT item only,
    // and only when that item has its own K / V buffers (ring == 2) and a tile for this stream: the loads of any item
    // that reuses this item's buffers wait for this tile's last P.V, which would then wait for them.
    Tile t = next_tile();
    bool s0_issued = false;
    while (t.valid) {
      if (!s0_issued) {
        ptx::mbar_wait(&q_full[wg], n_q & 1);
        issue_s(t, 0);
      }
      for (int kb = 1; kb < t.nk; ++kb) {
        issue_s(t, kb);
        issue_pv(t, kb - 1);
      }
      Tile n{0, 0, 0, 0, 0, 0u};
      const bool look = qt_next < cur_nq || (p.ring == 2 && item < n_items && g_ahead.nq > wg);
      s0_issued = false;
      if (look) {
        n = next_tile();
        if (n.valid) {
          ptx::mbar_wait(&q_full[wg], (n_q + 1) & 1);
          issue_s(n, 0);
          s0_issued = true;
        }
      }
      issue_pv(t, t.nk - 1);
      ++n_q;
      if (!look) n = next_tile();
      t = n;
    }
  } else if (warp_idx == 3) {
    ptx::setmaxnreg_dec<56>();   // TMEM owner: idles until the teardown, but its warpgroup releases registers as one
  } else if (warp_idx < 12) {
    // ===================================================== softmax warpgroup of stream wg (thread = query row)
    ptx::setmaxnreg_inc<184>();
    const int wg = (warp_idx - 4) >> 2;
    const int quad = warp_idx & 3;
    const int lane = threadIdx.x & 31;
    const int r = quad * 32 + lane;
    const uint32_t t_s = tmem_base + wg * 256 + (static_cast<uint32_t>(quad * 32) << 16);
    const int item_stride = gridDim.x;
    int it = 0;
    uint32_t n_s = 0;   // S blocks consumed by this stream so far
    // the geometry of the NEXT item (two dependent global loads) is fetched while this one is computed
    int item = blockIdx.x;
    ItemGeom g_next = item_geom(p.klen, p.cu, min(item, n_items - 1) / p.H, p.T, nkb);
    for (; item < n_items; item += item_stride) {
      const ItemGeom g = g_next;
      g_next = item_geom(p.klen, p.cu, min(item + item_stride, n_items - 1) / p.H, p.T, nkb);
      if (g.nq == 0) continue;
      const int set = it % p.ring;
      ++it;
      uint8_t* sV = smem + set * set_bytes + nkb * kTileBytes;
      for (int qt = wg; qt < g.nq; qt += 2) {
        float mc = 0.f;    // softmax reference point x scale (log2 domain)
        for (int kb = 0; kb < g.nk; ++kb, ++n_s) {
          const int nvalid = max(min(g.klen - kb * 128, 128), 0);
          ptx::mbar_wait(&s_full[wg], n_s & 1);
          ptx::tc_fence_after();
          uint32_t s[128];
          ptx::tmem_ld_32x32b_x32(t_s, s);
          ptx::tmem_ld_32x32b_x32(t_s + 32, s + 32);
          ptx::tmem_ld_32x32b_x32(t_s + 64, s + 64);
          ptx::tmem_ld_32x32b_x32(t_s + 96, s + 96);
          // S of block kb exists => its K / V have landed.  Once per stream and item (both streams write the same bytes;
          // ordered before this stream's P.V of the block by the p_full arrival below), while the loads above are in flight:
          //   * V rows past klen -> 0 (their P is 0, but 0 x stale inf / NaN bits would not be);
          //   * V column dk of the valid rows -> 1.0: P.V multiplies all 64 columns of the tile anyway, so O[:, dk] becomes
          //     the row sum of the fp16 P that the tensor core actually used -- the softmax denominator costs no
          //     instruction here and follows every rescaling of O by construction
          if (qt == wg) {
            uint8_t* v_tile = sV + kb * kTileBytes;
            if (r < nvalid) *reinterpret_cast<__half*>(v_tile + r * 128 + ((((p.dk >> 3)) ^ (r & 7)) << 4)) = __float2half_rn(1.0f);
            zero_v_tail(v_tile, r, nvalid);
          }
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&s_empty[wg]);   // the block is in registers: S(kb + 1) may overwrite it
          // ---- block maximum (keys past klen hold whatever the neighbouring rows produced: never looked at)
          float bm;
          if (nvalid == 128) {
            float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
            for (int j = 0; j < 128; j += 8) {
              m0 = fmaxf(m0, fmaxf(__uint_as_float(s[j]), __uint_as_float(s[j + 1])));
              m1 = fmaxf(m1, fmaxf(__uint_as_float(s[j + 2]), __uint_as_float(s[j + 3])));
              m2 = fmaxf(m2, fmaxf(__uint_as_float(s[j + 4]), __uint_as_float(s[j + 5])));
              m3 = fmaxf(m3, fmaxf(__uint_as_float(s[j + 6]), __uint_as_float(s[j + 7])));
            }
            bm = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
          } else {
            bm = -INFINITY;
#pragma unroll
            for (int j = 0; j < 128; ++j)
              if (j < nvalid) bm = fmaxf(bm, __uint_as_float(s[j]));
          }
          const float bml = bm * p.scale_log2;
          // ---- reference point: exact on the first block, afterwards moved only when it trails by more than 2^kLazyLog2
          float corr = 1.f;
          bool move = false;
          if (kb == 0) {
            mc = bm == -INFINITY ? 0.f : bml;
          } else if (bml > mc + kLazyLog2) {
            move = true;
            corr = ex2(mc - bml);
            mc = bml;
          }
          // P.V of the previous block has retired: the P columns are free and O (with its row-sum column) is current
          ptx::mbar_wait(&p_empty[wg], (n_s & 1) ^ 1);
          ptx::tc_fence_after();
          if (kb > 0 && __any_sync(0xffffffffu, move)) {
#pragma unroll 1
            for (int c = 0; c < 64; c += 16) {
              uint32_t o[16];
              ptx::tmem_ld_32x32b_x16(t_s + kOCol + c, o);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * corr);
              ptx::tmem_st_32x32b_x16(t_s + kOCol + c, o);
            }
          }
          // ---- p = exp2(s * scale - reference) -> packed fp16, 32 keys (16 columns) per store
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t pk[16];
            if ((c + 1) * 32 <= nvalid) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float x0 = fmaf(__uint_as_float(s[c * 32 + j]), p.scale_log2, -mc);
                const float x1 = fmaf(__uint_as_float(s[c * 32 + j + 1]), p.scale_log2, -mc);
                const float p0 = ex2(x0);
                const float p1 = ex2(x1);
                __half2 hh = __floats2half2_rn(p0, p1);
                pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float p0 = (c * 32 + j < nvalid) ? ex2(fmaf(__uint_as_float(s[c * 32 + j]), p.scale_log2, -mc)) : 0.f;
                const float p1 = (c * 32 + j + 1 < nvalid) ? ex2(fmaf(__uint_as_float(s[c * 32 + j + 1]), p.scale_log2, -mc)) : 0.f;
                __half2 hh = __floats2half2_rn(p0, p1);
                pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
              }
            }
            ptx::tmem_st_32x32b_x16(t_s + kPCol + c * 16, pk);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&p_full[wg]);
        }
        // (the query tile's output is the output warpgroup's business: this stream goes straight to its next tile)
      }
    }
  } else {
    // ===================================================== output warpgroup: O / row sum -> fp16 rows, for both streams.
    // The denominator sits in column dk of O (see the softmax warps), so nothing is handed over in registers: the
    // softmax warps never wait for the last P.V of a tile, the O read-out or the stores.
    ptx::setmaxnreg_dec<88>();
    const int quad = warp_idx & 3;
    const int lane = threadIdx.x & 31;
    uint8_t* stage_w = sStage + quad * (32 * kStagePitch);
    const int nch = p.dk >> 3;   // 16-byte chunks per output row
    const int item_stride = gridDim.x;
    uint32_t n_o0 = 0, n_o1 = 0;   // query tiles read out per stream
    int item = blockIdx.x;
    ItemGeom g_next = item_geom(p.klen, p.cu, min(item, n_items - 1) / p.H, p.T, nkb);
    for (; item < n_items; item += item_stride) {
      const int h = item % p.H;
      const ItemGeom g = g_next;
      g_next = item_geom(p.klen, p.cu, min(item + item_stride, n_items - 1) / p.H, p.T, nkb);
      for (int qt = 0; qt < g.nq; ++qt) {
        const int wg = qt & 1;
        const uint32_t t_o = tmem_base + wg * 256 + kOCol + (static_cast<uint32_t>(quad * 32) << 16);
        ptx::mbar_wait(&o_full[wg], (wg ? n_o1 : n_o0) & 1);
        if (wg) ++n_o1; else ++n_o0;
        ptx::tc_fence_after();
        uint32_t v[16];
        ptx::tmem_ld_32x32b_x16(t_o + p.dk, v);   // column dk: the row sum of P (a row without a valid key: 0 -> zeros)
        ptx::tmem_ld_wait();
        const float sum = __uint_as_float(v[0]);
        const float inv = sum > 0.f ? 1.0f / sum : 0.f;
        const int q0 = qt * 128 + quad * 32;
        __half* dst0 = p.out + (static_cast<size_t>(g.row0) + q0) * p.ld_out + h * p.dk;
        if (p.stage) __syncwarp();   // the previous tile's copy out of the staging tile is complete
        for (int c = 0; c < p.dk; c += 16) {
          ptx::tmem_ld_32x32b_x16(t_o + c, v);
          ptx::tmem_ld_wait();
          if (c + 16 >= p.dk) {   // O is in registers: the stream's next tile may overwrite it
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&o_empty[wg]);
          }
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            __half2 hh = __floats2half2_rn(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
            o[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          }
          if (p.stage) {
            uint4* d = reinterpret_cast<uint4*>(stage_w + lane * kStagePitch + c * 2);
            d[0] = make_uint4(o[0], o[1], o[2], o[3]);
            d[1] = make_uint4(o[4], o[5], o[6], o[7]);
          } else if (q0 + lane < g.qlim) {
            uint4* d = reinterpret_cast<uint4*>(dst0 + static_cast<size_t>(lane) * p.ld_out + c);
            d[0] = make_uint4(o[0], o[1], o[2], o[3]);
            d[1] = make_uint4(o[4], o[5], o[6], o[7]);
          }
        }
        if (p.stage) {
          // A thread owns a ROW (96 bytes, 1 536 bytes from its neighbour's): stored directly, every instruction of the warp
          // touches 32 lines with 16 bytes each.  Through the warp's staging tile the same bytes leave as runs of whole
          // rows: lane l of store i writes chunk (32 i + l) of the tile in row-major order.
          __syncwarp();
          for (int f = lane; f < 32 * nch; f += 32) {
            const int row = nch == 6 ? f / 6 : f / nch;
            const int ch = f - row * nch;
            if (q0 + row < g.qlim)
              *reinterpret_cast<uint4*>(dst0 + static_cast<size_t>(row) * p.ld_out + ch * 8) =
                  *reinterpret_cast<const uint4*>(stage_w + row * kStagePitch + ch * 16);
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 3) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

int launch_attention(const CUtensorMap* tmap_qkv, const int* klen, const int* cu, __half* out, int B, int T, int H, int dk,
                     int d_model, int num_sms, cudaStream_t s) {
  const int nkb = (T + 127) / 128;
  if (nkb > kMaxKB || dk % 16 != 0 || dk > 48 || (cu != nullptr && klen == nullptr)) return -1;
  AttnParams p;
  p.T = T;
  p.nkb = nkb;
  p.B = B;
  p.H = H;
  // shared memory: `ring` sets of K / V blocks + two Q tiles + barriers (+ the output staging tiles where they still fit:
  // T' <= 256 and 384 < T' <= 640).  Measured on the c3 shape (T' = 376), a second K / V set is worth more than staging
  p.ring = nkb <= 3 ? 2 : 1;
  p.stage = (p.ring * 2 * nkb + 2) * kTileBytes + kBarBytes + kStageBytes + 1024 <= 227 * 1024 ? 1 : 0;
  p.klen = klen;
  p.cu = cu;
  p.out = out;
  p.ld_out = d_model;
  p.dk = dk;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(dk));
  static PerDeviceOnce attr_once;
  if (attr_once.first()) cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  const int items = B * H;
  const int grid = items < num_sms ? items : num_sms;
  const int smem = (p.ring * 2 * nkb + 2) * kTileBytes + kBarBytes + (p.stage ? kStageBytes : 0) + 1024;
  return launch_k(attention_kernel, dim3(grid), dim3(kThreads), smem, s, *tmap_qkv, p) == cudaSuccess ? 0 : -2;
}

}  // namespace gam
