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
// Common to both kernels below: K / V blocks are 128 keys x 64 columns (SWIZZLE_128B; the 16 columns past the 48
// real ones belong to the next head and are never multiplied: QK^T runs K = 3 x 16 and the 16 extra output columns
// of P.V are dropped); S = Q K^T (128 x 128 x 48) lands in TMEM; the softmax warps (one thread per query row,
// tcgen05.ld 32x32b) write P back over the consumed S columns as packed fp16 (tcgen05.st) and O += P V runs as a
// TS-mode tcgen05.mma with V straight from its [key, d] layout (MN-major B).
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace gam {
namespace {

constexpr int kMaxKB = 6;              // up to 768 keys (30 s segments of the reference's VAD, gigaam/vad_utils.py:85)
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 fp16

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

// rows of key block kb_last past klen -> 0 in the V tile (SWIZZLE_128B permutes 16-byte chunks inside a 128-byte row, so
// clearing whole rows needs no address arithmetic); thread = row.  Followed by the generic -> async proxy fence.
__device__ __forceinline__ void zero_v_tail(uint8_t* v_tile, int r, int valid_rows) {
  if (r >= valid_rows) {
    uint4* row = reinterpret_cast<uint4*>(v_tile + r * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) row[j] = make_uint4(0u, 0u, 0u, 0u);
  }
  ptx::fence_proxy_async_smem();
}

// ---------------------------------------------------------------------------------------------------------
// Persistent variant for T <= 256 (the headline config, T' = 251): one CTA per SM loops over (utterance, head)
// work items.  Q/K/V of item i+1 stream into the other half of a double-buffered smem ring while item i is in
// its softmax, P never touches shared memory (softmax warps write it as packed fp16 straight over the consumed
// part of their S rows in TMEM with tcgen05.st, and P.V runs as a TS-mode tcgen05.mma with A in TMEM), and O
// lands in the dead upper half of the tile's S region.  TMEM: 2 query tiles x 256 columns.
struct AttnPersParams {
  int T, nkb, B, H;
  const int* klen;
  const int* cu;
  __half* out;
  int ld_out, dk;
  float scale_log2;
};

__global__ void __launch_bounds__(128 + 256, 1) attention_persistent_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                          const AttnPersParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int nkb = p.nkb;                    // 1 or 2: key blocks == query tiles
  const int item_bytes = 3 * nkb * kTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * item_bytes);
  uint64_t* kv_full = bars;        // [2] ring buffers
  uint64_t* kv_empty = bars + 2;   // [2]
  uint64_t* s_full = bars + 4;     // [2] query tiles
  uint64_t* p_full = bars + 6;     // [2]
  uint64_t* o_full = bars + 8;     // [2]
  uint64_t* o_empty = bars + 10;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp_idx = threadIdx.x >> 5;
  const int n_items = p.B * p.H;
  const int dmodel = p.ld_out;

  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmap_qkv);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], p.nkb);   // one commit per query-tile stream
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], 4);
      ptx::mbar_init(&o_full[i], 1);
      ptx::mbar_init(&o_empty[i], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp_idx == 3) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    // ===================================================== TMA producer
    if (ptx::elect_one()) {
      int it = 0;   // items that use the ring (an utterance without frames is skipped by every role)
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int b = item / p.H, h = item % p.H;
        const ItemGeom g = item_geom(p.klen, p.cu, b, p.T, nkb);
        if (g.nq == 0) continue;
        const int buf = it & 1;
        const uint32_t par = (it >> 1) & 1;
        ++it;
        const int row0 = g.row0;
        uint8_t* sQ = smem + buf * item_bytes;
        uint8_t* sK = sQ + nkb * kTileBytes;
        uint8_t* sV = sK + nkb * kTileBytes;
        ptx::mbar_wait(&kv_empty[buf], par ^ 1);
        ptx::mbar_arrive_expect_tx(&kv_full[buf], item_bytes);
        for (int t = 0; t < nkb; ++t) {
          ptx::tma_load_2d(sQ + t * kTileBytes, &tmap_qkv, &kv_full[buf], h * p.dk, row0 + t * 128);
          ptx::tma_load_2d(sK + t * kTileBytes, &tmap_qkv, &kv_full[buf], dmodel + h * p.dk, row0 + t * 128);
          ptx::tma_load_2d(sV + t * kTileBytes, &tmap_qkv, &kv_full[buf], 2 * dmodel + h * p.dk, row0 + t * 128);
        }
      }
    }
  } else if (warp_idx == 1 || warp_idx == 2) {
    // ===================================================== MMA issuers: one per query-tile stream, so the two streams
    // drift apart and fill each other's hand-off bubbles (S -> softmax -> P.V -> O read is a serial chain per tile)
    const int qt = warp_idx - 1;
    if (qt < nkb) {
      constexpr uint32_t kIdescS = ptx::make_idesc_f16(128, 128, 0, 0);
      constexpr uint32_t kIdescPV = ptx::make_idesc_f16(128, 64, 0, 1);   // B (= V) MN-major; A (= P) from TMEM
      const int ksteps_qk = p.dk / 16;
      int it = 0;        // ring position (shared by both streams)
      uint32_t n_mine = 0;   // items this stream has computed: parity of its own s / p / o barriers
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const ItemGeom g = item_geom(p.klen, p.cu, item / p.H, p.T, nkb);
        if (g.nq == 0) continue;
        const int buf = it & 1;
        const uint32_t kv_par = (it >> 1) & 1;
        ++it;
        uint8_t* sQ = smem + buf * item_bytes;
        uint8_t* sK = sQ + nkb * kTileBytes;
        uint8_t* sV = sK + nkb * kTileBytes;
        ptx::mbar_wait(&kv_full[buf], kv_par);
        if (qt >= g.nq) {
          // this stream's query tile holds no frame of the utterance: only release the ring slot (after the wait above, so
          // that a stream can never arrive twice within one phase of kv_empty)
          if (ptx::elect_one()) ptx::mbar_arrive(&kv_empty[buf]);
          __syncwarp();
          continue;
        }
        const uint32_t ipar = n_mine & 1;
        ++n_mine;
        ptx::mbar_wait(&o_empty[qt], ipar ^ 1);   // previous item's O has been read out of this tile's region
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t qa = ptx::smem_u32(sQ + qt * kTileBytes);
          for (int kb = 0; kb < g.nk; ++kb) {
            const uint32_t ka = ptx::smem_u32(sK + kb * kTileBytes);
            for (int k = 0; k < ksteps_qk; ++k)
              ptx::mma_f16_ss(tmem_base + qt * 256 + kb * 128, ptx::make_smem_desc_sw128(qa + k * 32, 16, 1024),
                              ptx::make_smem_desc_sw128(ka + k * 32, 16, 1024), kIdescS, k != 0 ? 1u : 0u);
          }
          ptx::mma_commit(&s_full[qt]);
        }
        __syncwarp();
        ptx::mbar_wait(&p_full[qt], ipar);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          for (int kb = 0; kb < g.nk; ++kb) {
            const uint32_t va = ptx::smem_u32(sV + kb * kTileBytes);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks)
              ptx::mma_f16_ts(tmem_base + qt * 256 + 128, tmem_base + qt * 256 + (kb * 8 + ks) * 8,
                              ptx::make_smem_desc_sw128(va + ks * 2048, 1024, 1024), kIdescPV, (kb | ks) != 0 ? 1u : 0u);
          }
          ptx::mma_commit(&o_full[qt]);
          ptx::mma_commit(&kv_empty[buf]);   // this stream is done with the ring slot (count = nkb streams)
        }
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4 && warp_idx < 4 + 4 * nkb) {
    // ===================================================== softmax + output (one warpgroup per query tile)
    const int qt = (warp_idx - 4) >> 2;
    const int quad = warp_idx & 3;
    const int lane = threadIdx.x & 31;
    const int r = quad * 32 + lane;
    const uint32_t t_s = tmem_base + qt * 256 + (static_cast<uint32_t>(quad * 32) << 16);
    int it = 0;
    uint32_t n_mine = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int b = item / p.H, h = item % p.H;
      const ItemGeom g = item_geom(p.klen, p.cu, b, p.T, nkb);
      if (g.nq == 0) continue;
      const int buf = it & 1;
      ++it;
      if (qt >= g.nq) continue;
      const uint32_t ipar = n_mine & 1;
      ++n_mine;
      const int klen = g.klen;
      const int nchunks = (klen + 31) >> 5;
      ptx::mbar_wait(&s_full[qt], ipar);   // S complete => this item's Q / K / V have all landed (one kv_full transaction)
      ptx::tc_fence_after();
      if (klen < g.nk * 128)
        zero_v_tail(smem + buf * item_bytes + (2 * nkb + g.nk - 1) * kTileBytes, r, klen - (g.nk - 1) * 128);
      // Both sweeps are software-pipelined over two register buffers: the TMEM load of chunk c+1 is in flight while
      // chunk c is reduced / exponentiated (tcgen05.wait::ld waits for every outstanding load, so the next load is
      // issued right after the wait and before the math).
      float m = -INFINITY;
      uint32_t va[32], vb[32];
      // only the chunk that straddles klen pays for per-key predicates (the softmax is issue-bound: ~10 instructions
      // per score before this split)
      const int full_chunks = klen >> 5;
      auto row_max = [&](const uint32_t* v, int c) {
        if (c < full_chunks) {
          float m0 = m, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            m0 = fmaxf(m0, __uint_as_float(v[j]));
            m1 = fmaxf(m1, __uint_as_float(v[j + 1]));
            m2 = fmaxf(m2, __uint_as_float(v[j + 2]));
            m3 = fmaxf(m3, __uint_as_float(v[j + 3]));
          }
          m = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < klen) m = fmaxf(m, __uint_as_float(v[j]));
        }
      };
      if (nchunks > 0) ptx::tmem_ld_32x32b_x32(t_s, va);
#pragma unroll 1
      for (int c = 0; c < nchunks; c += 2) {
        ptx::tmem_ld_wait();
        if (c + 1 < nchunks) ptx::tmem_ld_32x32b_x32(t_s + (c + 1) * 32, vb);
        row_max(va, c);
        if (c + 1 < nchunks) {
          ptx::tmem_ld_wait();
          if (c + 2 < nchunks) ptx::tmem_ld_32x32b_x32(t_s + (c + 2) * 32, va);
          row_max(vb, c + 1);
        }
      }
      if (m == -INFINITY) m = 0.f;
      const float mc = m * p.scale_log2;
      float sum = 0.f;
      // P chunk c (32 keys = 16 packed fp16 columns) overwrites columns [16c, 16c+16) of this row: part of S chunk c/2,
      // already consumed, and below every chunk whose load may still be in flight
      float sum1 = 0.f;
      auto exp_store = [&](const uint32_t* v, int c) {
        uint32_t pk[16];
        if (c < full_chunks) {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, -mc));
            const float p1 = ex2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -mc));
            sum += p0;
            sum1 += p1;
            __half2 hh = __floats2half2_rn(p0, p1);
            pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = (c * 32 + j < klen) ? ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, -mc)) : 0.f;
            const float p1 = (c * 32 + j + 1 < klen) ? ex2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -mc)) : 0.f;
            sum += p0;
            sum1 += p1;
            __half2 hh = __floats2half2_rn(p0, p1);
            pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
          }
        }
        ptx::tmem_st_32x32b_x16(t_s + c * 16, pk);
      };
      if (nchunks > 0) ptx::tmem_ld_32x32b_x32(t_s, va);
#pragma unroll 1
      for (int c = 0; c < nchunks; c += 2) {
        ptx::tmem_ld_wait();
        if (c + 1 < nchunks) ptx::tmem_ld_32x32b_x32(t_s + (c + 1) * 32, vb);
        exp_store(va, c);
        if (c + 1 < nchunks) {
          ptx::tmem_ld_wait();
          if (c + 2 < nchunks) ptx::tmem_ld_32x32b_x32(t_s + (c + 2) * 32, va);
          exp_store(vb, c + 1);
        }
      }
      {
        uint32_t zero[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) zero[j] = 0u;
        for (int c = nchunks; c < g.nk * 4; ++c) ptx::tmem_st_32x32b_x16(t_s + c * 16, zero);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[qt]);
      ptx::mbar_wait(&o_full[qt], ipar);
      ptx::tc_fence_after();
      uint32_t ov[48];
      ptx::tmem_ld_32x32b_x16(t_s + 128, ov);
      ptx::tmem_ld_32x32b_x16(t_s + 144, ov + 16);
      ptx::tmem_ld_32x32b_x16(t_s + 160, ov + 32);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&o_empty[qt]);   // region free for the next item's S
      sum += sum1;
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      const int q = qt * 128 + r;
      if (q < g.qlim) {
        __half* dst = p.out + (static_cast<size_t>(g.row0) + q) * p.ld_out + h * p.dk;
#pragma unroll
        for (int c = 0; c < 48; c += 8) {
          if (c < p.dk) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              __half2 hh = __floats2half2_rn(__uint_as_float(ov[c + j]) * inv, __uint_as_float(ov[c + j + 1]) * inv);
              o[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
            *reinterpret_cast<uint4*>(dst + c) = make_uint4(o[0], o[1], o[2], o[3]);
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

// ---------------------------------------------------------------------------------------------------------
// Long-sequence variant (any T <= 768, used for T > 256): one CTA per (utterance, head) keeps ALL K / V blocks of
// the head resident in shared memory and walks the query tiles; two softmax warpgroups (each with its own MMA
// issuer warp, S slot, P alias and O accumulator in TMEM) process alternate query tiles concurrently.  A score
// row no longer fits TMEM, so each query tile makes two sweeps over the key blocks (max, then exp / P.V) with
// S recomputed in the second sweep: tensor time is cheap here, the exp throughput (MUFU) is the bound.
struct AttnLongParams {
  int T, nkb, H;
  const int* klen;
  const int* cu;
  __half* out;
  int ld_out, dk;
  float scale_log2;
};

constexpr int kLongThreads = 384;   // warp 0 TMA, 1/2 MMA issuers, 3 TMEM owner, 4-7 / 8-11 softmax warpgroups

__global__ void __launch_bounds__(kLongThreads, 1) attention_long_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                       const AttnLongParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const ItemGeom g = item_geom(p.klen, p.cu, blockIdx.x / p.H, p.T, p.nkb);
  if (g.nq == 0) return;                       // packed rows: the utterance has no frame (uniform for the CTA)
  const int nkb = g.nk;                        // key blocks with a valid key (>= 1); shared memory is sized for p.nkb
  uint8_t* sK = smem;                          // [nkb]
  uint8_t* sV = sK + nkb * kTileBytes;         // [nkb]
  uint8_t* sQ = sV + nkb * kTileBytes;         // [2] one per warpgroup
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + 2 * kTileBytes);
  uint64_t* kv_full = bars;          // [kMaxKB]
  uint64_t* q_full = bars + kMaxKB;  // [2]
  uint64_t* q_empty = q_full + 2;    // [2]
  uint64_t* s_full = q_full + 4;     // [2]
  uint64_t* s_empty = q_full + 6;    // [2]
  uint64_t* p_full = q_full + 8;     // [2]
  uint64_t* o_full = q_full + 10;    // [2]
  uint64_t* o_empty = q_full + 12;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 14);

  const int warp_idx = threadIdx.x >> 5;
  const int h = blockIdx.x % p.H;
  const int row0 = g.row0;
  const int dmodel = p.ld_out;
  const int nqt = g.nq;

  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmap_qkv);
    for (int i = 0; i < kMaxKB; ++i) ptx::mbar_init(&kv_full[i], 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&q_full[i], 1);
      ptx::mbar_init(&q_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&s_empty[i], 4);
      ptx::mbar_init(&p_full[i], 4);
      ptx::mbar_init(&o_full[i], 1);
      ptx::mbar_init(&o_empty[i], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp_idx == 3) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    if (ptx::elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_arrive_expect_tx(&kv_full[kb], 2 * kTileBytes);
        ptx::tma_load_2d(sK + kb * kTileBytes, &tmap_qkv, &kv_full[kb], dmodel + h * p.dk, row0 + kb * 128);
        ptx::tma_load_2d(sV + kb * kTileBytes, &tmap_qkv, &kv_full[kb], 2 * dmodel + h * p.dk, row0 + kb * 128);
      }
      for (int qt = 0; qt < nqt; ++qt) {
        const int wg = qt & 1, it = qt >> 1;
        ptx::mbar_wait(&q_empty[wg], (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&q_full[wg], kTileBytes);
        ptx::tma_load_2d(sQ + wg * kTileBytes, &tmap_qkv, &q_full[wg], h * p.dk, row0 + qt * 128);
      }
    }
  } else if (warp_idx == 1 || warp_idx == 2) {
    // ===================================================== MMA issuer of warpgroup wg
    const int wg = warp_idx - 1;
    constexpr uint32_t kIdescS = ptx::make_idesc_f16(128, 128, 0, 0);
    constexpr uint32_t kIdescPV = ptx::make_idesc_f16(128, 64, 0, 1);
    const int ksteps_qk = p.dk / 16;
    const uint32_t t_s = tmem_base + wg * 256;        // S slot [0,128) (P aliases [0,64)), O at [128,192)
    const uint32_t qa = ptx::smem_u32(sQ + wg * kTileBytes);
    int it = 0;
    for (int qt = wg; qt < nqt; qt += 2, ++it) {
      ptx::mbar_wait(&q_full[wg], it & 1);
      ptx::mbar_wait(&o_empty[wg], (it & 1) ^ 1);
      for (int pass = 0; pass < 2; ++pass) {
        for (int kb = 0; kb < nkb; ++kb) {
          // The S slot may be overwritten once the softmax warps have read the previous max-sweep tile (they arrive
          // on s_empty only in sweep 0: phase index it*nkb + kb').  In sweep 1 the slot also holds P, but the next S
          // MMA is issued by this same thread after the P.V MMAs and the tensor pipe executes them in order.
          if (pass == 0 && kb > 0) ptx::mbar_wait(s_empty + wg, (it * nkb + kb - 1) & 1);
          if (pass == 1 && kb == 0) ptx::mbar_wait(s_empty + wg, (it * nkb + nkb - 1) & 1);
          if (it == 0 && pass == 0) ptx::mbar_wait(&kv_full[kb], 0);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t ka = ptx::smem_u32(sK + kb * kTileBytes);
            for (int k = 0; k < ksteps_qk; ++k)
              ptx::mma_f16_ss(t_s, ptx::make_smem_desc_sw128(qa + k * 32, 16, 1024), ptx::make_smem_desc_sw128(ka + k * 32, 16, 1024),
                              kIdescS, k != 0 ? 1u : 0u);
            ptx::mma_commit(&s_full[wg]);
          }
          __syncwarp();
          if (pass == 1) {
            ptx::mbar_wait(&p_full[wg], (it * nkb + kb) & 1);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint32_t va = ptx::smem_u32(sV + kb * kTileBytes);
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::mma_f16_ts(t_s + 128, t_s + ks * 8, ptx::make_smem_desc_sw128(va + ks * 2048, 1024, 1024), kIdescPV,
                                (kb | ks) != 0 ? 1u : 0u);
              if (kb == nkb - 1) {
                ptx::mma_commit(&o_full[wg]);
                ptx::mma_commit(&q_empty[wg]);
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================================================== softmax warpgroups
    const int wg = (warp_idx - 4) >> 2;
    const int quad = warp_idx & 3;
    const int lane = threadIdx.x & 31;
    const int r = quad * 32 + lane;
    const uint32_t t_s = tmem_base + wg * 256 + (static_cast<uint32_t>(quad * 32) << 16);
    const int klen = g.klen;
    uint32_t n_s = 0;
    int it = 0;
    for (int qt = wg; qt < nqt; qt += 2, ++it) {
      float m = -INFINITY;
      for (int kb = 0; kb < nkb; ++kb, ++n_s) {
        ptx::mbar_wait(&s_full[wg], n_s & 1);
        ptx::tc_fence_after();
        // S of block kb exists => kv_full[kb] completed => V[kb] has landed: clear its rows past klen once per warpgroup,
        // long before this warpgroup's first P.V over the block (ordered by its later p_full arrivals)
        if (it == 0 && kb == nkb - 1 && klen < nkb * 128) zero_v_tail(sV + kb * kTileBytes, r, klen - kb * 128);
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int key0 = kb * 128 + c * 32;
          if (key0 >= klen) break;
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(t_s + c * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (key0 + j < klen) m = fmaxf(m, __uint_as_float(v[j]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(s_empty + wg);
      }
      if (m == -INFINITY) m = 0.f;
      const float mc = m * p.scale_log2;
      float sum = 0.f;
      for (int kb = 0; kb < nkb; ++kb, ++n_s) {
        ptx::mbar_wait(&s_full[wg], n_s & 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          const int key0 = kb * 128 + c * 32;
          uint32_t pk[16];
          if (key0 < klen) {
            uint32_t v[32];
            ptx::tmem_ld_32x32b_x32(t_s + c * 32, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = (key0 + j < klen) ? ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, -mc)) : 0.f;
              const float p1 = (key0 + j + 1 < klen) ? ex2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -mc)) : 0.f;
              sum += p0 + p1;
              __half2 hh = __floats2half2_rn(p0, p1);
              pk[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[j] = 0u;
          }
          ptx::tmem_st_32x32b_x16(t_s + c * 16, pk);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[wg]);   // (the P.V commit, not us, releases the slot in this sweep)
      }
      ptx::mbar_wait(&o_full[wg], it & 1);
      ptx::tc_fence_after();
      uint32_t ov[48];
      ptx::tmem_ld_32x32b_x16(t_s + 128, ov);
      ptx::tmem_ld_32x32b_x16(t_s + 144, ov + 16);
      ptx::tmem_ld_32x32b_x16(t_s + 160, ov + 32);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&o_empty[wg]);
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      const int q = qt * 128 + r;
      if (q < g.qlim) {
        __half* dst = p.out + (static_cast<size_t>(row0) + q) * p.ld_out + h * p.dk;
#pragma unroll
        for (int c = 0; c < 48; c += 8) {
          if (c < p.dk) {
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 8; j += 2) {
              __half2 hh = __floats2half2_rn(__uint_as_float(ov[c + j]) * inv, __uint_as_float(ov[c + j + 1]) * inv);
              o[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
            }
            *reinterpret_cast<uint4*>(dst + c) = make_uint4(o[0], o[1], o[2], o[3]);
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

static int launch_attention_long(const CUtensorMap* tmap_qkv, const int* klen, const int* cu, __half* out, int B, int T, int H,
                                 int dk, int d_model, cudaStream_t s) {
  AttnLongParams p;
  p.T = T;
  p.nkb = (T + 127) / 128;
  p.H = H;
  p.klen = klen;
  p.cu = cu;
  p.out = out;
  p.ld_out = d_model;
  p.dk = dk;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(dk));
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncSetAttribute(attention_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int smem = (2 * p.nkb + 2) * kTileBytes + 256 + 1024;
  attention_long_kernel<<<B * H, kLongThreads, smem, s>>>(*tmap_qkv, p);
  return 0;
}

static int launch_attention_persistent(const CUtensorMap* tmap_qkv, const int* klen, const int* cu, __half* out, int B, int T,
                                       int H, int dk, int d_model, int num_sms, cudaStream_t s) {
  AttnPersParams p;
  p.T = T;
  p.nkb = (T + 127) / 128;
  p.B = B;
  p.H = H;
  p.klen = klen;
  p.cu = cu;
  p.out = out;
  p.ld_out = d_model;
  p.dk = dk;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(dk));
  static PerDeviceOnce attr_once;
  if (attr_once.first()) {
    cudaFuncSetAttribute(attention_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  }
  const int items = B * H;
  const int grid = items < num_sms ? items : num_sms;
  const int smem = 2 * 3 * p.nkb * kTileBytes + 256 + 1024;
  return launch_k(attention_persistent_kernel, dim3(grid), dim3(128 + 128 * p.nkb), smem, s, *tmap_qkv, p) == cudaSuccess ? 0 : -2;
}
int launch_attention(const CUtensorMap* tmap_qkv, const int* klen, const int* cu, __half* out, int B, int T, int H, int dk,
                     int d_model, int num_sms, cudaStream_t s) {
  const int nkb = (T + 127) / 128;
  if (nkb > kMaxKB || dk % 16 != 0 || dk > 64 || (cu != nullptr && klen == nullptr)) return -1;
  if (nkb <= 2) return launch_attention_persistent(tmap_qkv, klen, cu, out, B, T, H, dk, d_model, num_sms, s);
  return launch_attention_long(tmap_qkv, klen, cu, out, B, T, H, dk, d_model, s);
}

}  // namespace gam
