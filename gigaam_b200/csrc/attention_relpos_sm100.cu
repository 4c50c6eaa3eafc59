// tcgen05 attention with Transformer-XL relative positions (v1_* checkpoints, self_attention_model == "rel_pos").
// Replaces RelPositionMultiHeadAttention.forward (gigaam/encoder.py:208-228) + forward_attention (:173-188) on
// the output of ONE projection GEMM whose weight is [W_q ; W_q ; W_k ; W_v] and whose bias carries pos_bias_u / _v:
//
//   qkv : [rows, 4*768] fp16 = [q+u | q+v | k | v], head h at columns h*48 .. h*48+47 of each part; rows are packed
//         (utterance b = rows cu[b] .. cu[b] + klen[b]; cu == null: the padded [B, T] layout of the unit tests)
//   pos : [2*kRelPosMaxT-1, 768] fp16 = W_pos pe(r) for r = kRelPosMaxT-1 ... -(kRelPosMaxT-1)   (row = kRelPosMaxT-1-r)
//   out : [rows, 768] fp16
//
//   s[i, j] = ((q_i+u) . k_j + (q_i+v) . p_{i-j}) / sqrt(d_k)          p_r = pos row for relative position r
//
// The reference materialises (q+v) P^T for all 2T-1 positions and re-indexes it with the pad/view "rel_shift"
// (:202-206).  Here, for a tile of 128 queries x 128 keys, the positions that can occur are the 255 consecutive
// table rows i0-j0-127 .. i0-j0+127, so ONE extra MMA (128 x 256 x 48) against that window gives every position score
// of the tile, and the shift becomes "row r reads column c - r + 127".  tcgen05.ld cannot take a per-lane column, so
// each softmax thread bounces a 48-column window of its row through a private shared-memory row and reads it back
// at its own offset (conflict-free: 52-word pitch).
//
// One CTA per (128-query tile, head, utterance); single sweep with a running maximum (O is rescaled in TMEM):
//   warp 0    : TMA   - Q(u), Q(v); per key block a 2-stage ring of {K, V, 256 position rows}
//   warp 1    : MMA   - S = Qu K^T  (TMEM cols 0..127), BD = Qv Pw^T (cols 256..511), O += P V (cols 128..191)
//   warps 2-5 : softmax, one thread per query row
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace gam {
namespace {

constexpr int kThreads = 192;
constexpr int kMaxKB = 6;                // up to 768 keys
constexpr int kTile = 128 * 128;         // bytes of a 128-row x 64-column fp16 tile
constexpr int kStageBytes = 4 * kTile;   // K, V, 2 position tiles
constexpr int kSkewPitch = 52;           // words per private row (48 used)
constexpr int kSkewBytes = 4 * 32 * kSkewPitch * 4;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kOCol = 128, kBDCol = 256;

struct RelParams {
  int T;
  const int* klen;   // may be null
  const int* cu;     // may be null (padded rows)
  __half* out;
  int ld_out;        // d_model
  int dk;
  int pos_center;    // table row of relative position 0 (= kRelPosMaxT - 1)
  float scale_log2;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// the skew bounce: same-thread store -> load through shared memory, kept in program order by the memory clobbers
__device__ __forceinline__ void skew_store4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float skew_load(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

__global__ void __launch_bounds__(kThreads, 1) attention_relpos_kernel(const __grid_constant__ CUtensorMap tmap_qkv,
                                                                      const __grid_constant__ CUtensorMap tmap_pos,
                                                                      const RelParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQu = smem;
  uint8_t* sQv = smem + kTile;
  uint8_t* sP = smem + 2 * kTile;                 // 2 chunks of 64 keys
  uint8_t* sStage = smem + 4 * kTile;             // 2 x {K, V, Pw0, Pw1}
  float* sSkew = reinterpret_cast<float*>(sStage + 2 * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sSkew) + kSkewBytes);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;    // [2]
  uint64_t* kv_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_empty = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* p_empty = bars + 8;
  uint64_t* o_full = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp_idx = threadIdx.x >> 5;
  const int q0 = blockIdx.x * 128;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  int klen = p.T;
  if (p.klen != nullptr) klen = min(max(p.klen[b], 0), p.T);
  // packed rows: a query tile past the utterance's last frame has nothing to compute or store (uniform for the CTA)
  if (p.cu != nullptr && q0 >= klen) return;
  const int row0 = p.cu != nullptr ? __ldg(p.cu + b) : b * p.T;
  const int qlim = p.cu != nullptr ? klen : p.T;   // query rows that are stored
  const int dmodel = p.ld_out;
  const int nkb = (klen + 127) >> 7;   // key blocks that hold at least one valid key

  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmap_qkv);
    ptx::prefetch_tmap(&tmap_pos);
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(s_empty, 4);
    ptx::mbar_init(p_full, 4);
    ptx::mbar_init(p_empty, 1);
    ptx::mbar_init(o_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp_idx == 1) ptx::tmem_alloc<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_idx == 0) {
    if (ptx::elect_one() && nkb > 0) {
      ptx::mbar_arrive_expect_tx(q_full, 2 * kTile);
      ptx::tma_load_2d(sQu, &tmap_qkv, q_full, h * p.dk, row0 + q0);
      ptx::tma_load_2d(sQv, &tmap_qkv, q_full, dmodel + h * p.dk, row0 + q0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int st = kb & 1;
        if (kb >= 2) ptx::mbar_wait(&kv_empty[st], ((kb >> 1) - 1) & 1);
        uint8_t* base = sStage + st * kStageBytes;
        ptx::mbar_arrive_expect_tx(&kv_full[st], kStageBytes);
        ptx::tma_load_2d(base, &tmap_qkv, &kv_full[st], 2 * dmodel + h * p.dk, row0 + kb * 128);
        ptx::tma_load_2d(base + kTile, &tmap_qkv, &kv_full[st], 3 * dmodel + h * p.dk, row0 + kb * 128);
        // window row w holds relative position (q0 + 127 - kb*128) - w  ->  score(r, c) sits at w = c - r + 127
        const int prow = p.pos_center - (q0 + 127) + kb * 128;
        ptx::tma_load_2d(base + 2 * kTile, &tmap_pos, &kv_full[st], h * p.dk, prow);
        ptx::tma_load_2d(base + 3 * kTile, &tmap_pos, &kv_full[st], h * p.dk, prow + 128);
      }
    }
  } else if (warp_idx == 1) {
    if (nkb > 0) {
      constexpr uint32_t kIdescS = ptx::make_idesc_f16(128, 128, 0, 0);
      constexpr uint32_t kIdescBD = ptx::make_idesc_f16(128, 256, 0, 0);
      constexpr uint32_t kIdescPV = ptx::make_idesc_f16(128, 64, 0, 1);   // B (= V) is MN-major
      const int ksteps = p.dk / 16;
      ptx::mbar_wait(q_full, 0);
      for (int kb = 0; kb <= nkb; ++kb) {
        if (kb < nkb) {
          const int st = kb & 1;
          ptx::mbar_wait(&kv_full[st], (kb >> 1) & 1);
          if (kb > 0) ptx::mbar_wait(s_empty, (kb - 1) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t qu = ptx::smem_u32(sQu), qv = ptx::smem_u32(sQv);
            const uint32_t ka = ptx::smem_u32(sStage + st * kStageBytes);
            const uint32_t pw = ka + 2 * kTile;
            for (int k = 0; k < ksteps; ++k)
              ptx::mma_f16_ss(tmem_base, ptx::make_smem_desc_sw128(qu + k * 32, 16, 1024),
                              ptx::make_smem_desc_sw128(ka + k * 32, 16, 1024), kIdescS, k != 0 ? 1u : 0u);
            for (int k = 0; k < ksteps; ++k)
              ptx::mma_f16_ss(tmem_base + kBDCol, ptx::make_smem_desc_sw128(qv + k * 32, 16, 1024),
                              ptx::make_smem_desc_sw128(pw + k * 32, 16, 1024), kIdescBD, k != 0 ? 1u : 0u);
            ptx::mma_commit(s_full);
          }
          __syncwarp();
        }
        if (kb > 0) {
          const int pb = kb - 1, st = pb & 1;
          ptx::mbar_wait(p_full, pb & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t pa = ptx::smem_u32(sP);
            const uint32_t va = ptx::smem_u32(sStage + st * kStageBytes + kTile);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t da = ptx::make_smem_desc_sw128(pa + (ks >> 2) * kTile + (ks & 3) * 32, 16, 1024);
              const uint64_t db = ptx::make_smem_desc_sw128(va + ks * 2048, 1024, 1024);
              ptx::mma_f16_ss(tmem_base + kOCol, da, db, kIdescPV, (pb | ks) != 0 ? 1u : 0u);
            }
            ptx::mma_commit(p_empty);
            ptx::mma_commit(&kv_empty[st]);
            if (pb == nkb - 1) ptx::mma_commit(o_full);
          }
          __syncwarp();
        }
      }
    }
  } else {
    const int quad = warp_idx & 3;
    const int lane = threadIdx.x & 31;
    const int r = quad * 32 + lane;
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const uint32_t my_row = ptx::smem_u32(sSkew + (quad * 32 + lane) * kSkewPitch);   // quad is a bijection of warps 2..5
    const uint32_t my_read = my_row + 4u * static_cast<uint32_t>(31 - lane);
    const int q = q0 + r;
    __half* dst = p.out + static_cast<size_t>(row0 + q) * p.ld_out + h * p.dk;
    float m = -INFINITY, sum = 0.f;
    for (int kb = 0; kb < nkb; ++kb) {
      ptx::mbar_wait(s_full, kb & 1);
      ptx::tc_fence_after();
      const int nvalid = min(klen - kb * 128, 128);
      // S of this block exists => its K / V / position tiles have landed: clear the V rows past klen (their P is 0, but
      // 0 x stale inf / NaN bits would not be); made visible to the tensor core by the proxy fence before p_full below
      if (nvalid < 128 && r >= nvalid) {
        uint4* vrow = reinterpret_cast<uint4*>(sStage + (kb & 1) * kStageBytes + kTile + r * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) vrow[j] = make_uint4(0u, 0u, 0u, 0u);
      }
      // ---- sweep A: s = ac + shifted bd, written back over the ac columns; block maximum
      float bm = -INFINITY;
#pragma unroll 1
      for (int stp = 0; stp < 8; ++stp) {
        if (stp * 16 >= nvalid) break;
        uint32_t ac[16], bd[48];
        const uint32_t wb = static_cast<uint32_t>(stp * 16 + 96 - 32 * quad);
        ptx::tmem_ld_32x32b_x16(t_s + stp * 16, ac);
        ptx::tmem_ld_32x32b_x16(t_s + kBDCol + wb, bd);
        ptx::tmem_ld_32x32b_x16(t_s + kBDCol + wb + 16, bd + 16);
        ptx::tmem_ld_32x32b_x16(t_s + kBDCol + wb + 32, bd + 32);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 12; ++j) skew_store4(my_row + 16 * j, bd[4 * j], bd[4 * j + 1], bd[4 * j + 2], bd[4 * j + 3]);
        uint32_t sv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float s = __uint_as_float(ac[j]) + skew_load(my_read + 4 * j);
          if (stp * 16 + j < nvalid) bm = fmaxf(bm, s);
          sv[j] = __float_as_uint(s);
        }
        ptx::tmem_st_32x32b_x16(t_s + stp * 16, sv);
      }
      ptx::tmem_st_wait();
      const float m_new = fmaxf(m, bm);
      const float mc = m_new * p.scale_log2;
      // ---- sweep B: p = exp2((s - m) * scale) -> fp16
      float bsum = 0.f;
      uint32_t pk[64];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c * 32 < nvalid) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(t_s + c * 32, v);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const float p0 = (c * 32 + j < nvalid) ? ex2(fmaf(__uint_as_float(v[j]), p.scale_log2, -mc)) : 0.f;
            const float p1 = (c * 32 + j + 1 < nvalid) ? ex2(fmaf(__uint_as_float(v[j + 1]), p.scale_log2, -mc)) : 0.f;
            bsum += p0 + p1;
            __half2 hh = __floats2half2_rn(p0, p1);
            pk[c * 16 + (j >> 1)] = *reinterpret_cast<uint32_t*>(&hh);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[c * 16 + j] = 0u;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_empty);
      const float corr = ex2(fmaf(m, p.scale_log2, -mc));   // m == -inf on the first block -> 0
      sum = sum * corr + bsum;
      m = m_new;
      ptx::mbar_wait(p_empty, (kb & 1) ^ 1);                // P V of the previous block has completed
      if (kb > 0) {
        ptx::tc_fence_after();
        for (int c = 0; c < p.dk; c += 16) {
          uint32_t o[16];
          ptx::tmem_ld_32x32b_x16(t_s + kOCol + c, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * corr);
          ptx::tmem_st_32x32b_x16(t_s + kOCol + c, o);
        }
        ptx::tmem_st_wait();
      }
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          uint4 val = make_uint4(pk[ch * 32 + j * 4], pk[ch * 32 + j * 4 + 1], pk[ch * 32 + j * 4 + 2], pk[ch * 32 + j * 4 + 3]);
          *reinterpret_cast<uint4*>(sP + ch * kTile + r * 128 + ((j ^ (r & 7)) << 4)) = val;
        }
      }
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    // ---- epilogue: O / sum -> fp16 (an utterance without a single valid key gets zeros)
    if (nkb > 0) {
      ptx::mbar_wait(o_full, 0);
      ptx::tc_fence_after();
    }
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
    for (int c = 0; c < p.dk; c += 16) {
      uint32_t v[16];
      if (nkb > 0) {
        ptx::tmem_ld_32x32b_x16(t_s + kOCol + c, v);
        ptx::tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = 0u;
      }
      if (q < qlim) {
        uint32_t o[8];
#pragma unroll
        for (int j = 0; j < 16; j += 2) {
          __half2 hh = __floats2half2_rn(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
          o[j >> 1] = *reinterpret_cast<uint32_t*>(&hh);
        }
        uint4* d4 = reinterpret_cast<uint4*>(dst + c);
        d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
      }
    }
    ptx::tc_fence_before();
  }

  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

constexpr int kSmemBytes = 4 * kTile + 2 * kStageBytes + kSkewBytes + 128 + 1024;

}  // namespace

int launch_attention_relpos(const CUtensorMap* tmap_qkv, const CUtensorMap* tmap_pos, const int* klen, const int* cu, __half* out,
                            int B, int T, int H, int dk, int d_model, cudaStream_t s) {
  const int nkb = (T + 127) / 128;
  if (nkb > kMaxKB || T > kRelPosMaxT || dk % 16 != 0 || dk > 64 || (cu != nullptr && klen == nullptr)) return -1;
  static PerDeviceOnce attr_once;
  if (attr_once.first() &&
      cudaFuncSetAttribute(attention_relpos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess)
    return -2;
  RelParams p;
  p.T = T;
  p.klen = klen;
  p.cu = cu;
  p.out = out;
  p.ld_out = d_model;
  p.dk = dk;
  p.pos_center = kRelPosMaxT - 1;
  p.scale_log2 = 1.4426950408889634f / sqrtf(static_cast<float>(dk));
  dim3 grid(nkb, H, B);
  attention_relpos_kernel<<<grid, kThreads, kSmemBytes, s>>>(*tmap_qkv, *tmap_pos, p);
  return 0;
}

}  // namespace gam
