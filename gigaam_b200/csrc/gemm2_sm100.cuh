// CTA-pair (cta_group::2) tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
// Two CTAs on the two SMs of a TPC compute one 256 x 256 tile: each CTA stages its own 128 rows of A and HALF
// (128 rows) of the W tile, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), and each CTA ends up with
// the accumulators of its 128 rows (x 256 columns) in its own TMEM.  Per k-block each SM moves
// 16 KB (A) + 16 KB (W half) through shared memory instead of 16 + 32 KB: the single-CTA kernel
// (gemm_sm100.cuh) was bound by shared-memory / L2 operand traffic at ~50 % tensor-pipe.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader CTA only) + TMEM owner,
// warps 2..9 = epilogue (two warps per TMEM lane quadrant, each owning 128 of the 256 accumulator columns).
// 5-stage smem ring, double-buffered TMEM accumulators (2 x 256 columns).
// Barriers: full[s] lives in the leader (both CTAs' TMA bytes + both producers' arrivals land there),
// empty[s] / tmem_full[a] are per CTA and signalled with multicast tcgen05.commit, tmem_empty[a] lives in the
// leader and collects one arrival per epilogue warp of both CTAs.  Remote arrivals are plain
// (CTA-scope) arrives: a cluster-scope release fence per k-block serialised the producer (measured).
//
// Epilogue: TMEM -> registers (thread = row) -> per-warp smem staging tile (pitch 36 words: conflict-free both
// ways) -> coalesced 16-byte global accesses (a warp instruction covers 4 full 128-byte rows), where bias /
// residual / activation are applied.  Residual reads are software-pipelined one chunk ahead; with eight warps
// that keeps 16 chunk loads in flight per SM, enough to hide L2 latency behind the next tile's main loop.
#pragma once
#include "gemm_params.cuh"

namespace gam {

constexpr int kG2EpiWarps = 8;                // epilogue warps per SET (one set drains one accumulator)
constexpr int kG2ABytes = 128 * 64 * 2;       // 16 KB: this CTA's 128 rows of A
constexpr int kG2BBytes = 128 * 64 * 2;       // 16 KB: this CTA's half of the 256-row W tile
constexpr int kG2StageBytes = kG2ABytes + kG2BBytes;
constexpr int kG2WarpStage = 32 * 36 * 4;     // 4608 B staging tile per epilogue warp
constexpr int kG2WarpBias = 128 * 4;          // per-warp copy of its 128 bias values
constexpr int kG2BarBytes = 256;

// The fp32 residual epilogues (read x, write x, + the LayerNorm tail) move 4-6x the bytes of the fp16 ones through
// 8 latency-bound warps and bound the N = 768 GEMMs (K = 768: 39 us per launch against a 10 us main loop).  They run with
// TWO epilogue sets: set s drains accumulator s, i.e. the tiles of local parity s, so two tiles are in their epilogue at
// once and the bytes in flight per SM double.  Cost: a 4-stage operand ring, and 20 warps (warpgroup 0 = producer, MMA
// issuer and two idle warps; warpgroups 1-4 = the two sets) whose register file is re-split with setmaxnreg: 24 per
// thread for warpgroup 0, 112 for the epilogue warpgroups (the .inc draws on the CTA pool the .dec fills: 128 x 72 released >= 512 x 16 claimed).
template <int EPI>
struct G2Cfg {
  static constexpr int kSets = (EPI == EPI_BIAS_RES_F32 || EPI == EPI_BIAS_RES_LN_F32) ? 2 : 1;
  static constexpr int kStages = kSets == 2 ? 4 : 5;
  static constexpr int kEpiWarp0 = kSets == 2 ? 4 : 2;          // first epilogue warp
  static constexpr int kThreads = 32 * (kEpiWarp0 + kG2EpiWarps * kSets);
  static constexpr int kEpiBytes = kSets * kG2EpiWarps * (kG2WarpStage + kG2WarpBias);
  static constexpr int kSmem = kStages * kG2StageBytes + kEpiBytes + kG2BarBytes + 1024;
};

// ------------------------------------------------------------------------------------------------------------------
// LayerNorm tail of EPI_BIAS_RES_LN_F32 (LnFuse in gemm_params.cuh).  Called by one epilogue warp after it has stored
// x = res + scale * (acc + bias) for its 32 rows x 128 columns; thread layout as in the store loop: lane covers rows
// i * 4 + (lane >> 3), i = 0..7, and 4 consecutive columns (lane & 7) * 4 of every 32-column chunk.
constexpr int kLnSlots = 6;          // 3 n-tiles x 2 column halves of a 768-wide row
constexpr int kLnD = 768;

// sum of `s` / `q` over the 8 lanes that share a row (lane & 7 varies); lanes with (lane & 7) == i keep row i's totals
__device__ __forceinline__ void ln_row_add(float s, float q, int i, int lane, float& ps, float& pq) {
#pragma unroll
  for (int o = 1; o < 8; o <<= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if ((lane & 7) == i) { ps += s; pq += q; }
}

// lane l publishes (sum, sum of squares) of row (l & 7) * 4 + (l >> 3) over this warp's 128 columns, then the warp's
// arrival is counted with release semantics
__device__ __forceinline__ void ln_publish(float ps, float pq, float2* stats, unsigned int* cnt, long long warp_row0, int rows_valid,
                                           int group, int slot, int lane) {
  const int pr = (lane & 7) * 4 + (lane >> 3);
  if (pr < rows_valid) stats[(warp_row0 + pr) * kLnSlots + slot] = make_float2(ps, pq);
  __syncwarp();
  if (lane == 0) {
    __threadfence();              // this warp's x and stats stores (ordered before by the __syncwarp) become visible GPU-wide ...
    atomicAdd(&cnt[group], 1u);   // ... before the arrival is counted
  }
}

// wait until all six column slices of this warp's rows have been published.  A peer that never arrives (the grid was
// not co-resident) traps after ~1-2 s instead of hanging the device
__device__ __forceinline__ void ln_wait(const unsigned int* cnt, int group, int lane) {
  if (lane == 0) {
    unsigned int v = 0;
    long long spins = 0;
    while (true) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cnt + group) : "memory");
      if (v >= static_cast<unsigned int>(kLnSlots)) break;
      __nanosleep(64);
      if (++spins > (1ll << 24)) __trap();
    }
  }
  __syncwarp();
}

__device__ __forceinline__ uint2 ln_pack4(float4 v) {
  return make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
}
__device__ __forceinline__ float4 ln_affine(float4 v, float mean, float rstd, float4 g, float4 b) {
  return make_float4((v.x - mean) * rstd * g.x + b.x, (v.y - mean) * rstd * g.y + b.y, (v.z - mean) * rstd * g.z + b.z,
                     (v.w - mean) * rstd * g.w + b.w);
}

// The passes below are latency bound (8 epilogue warps per SM against a ~1 us loaded L2 round trip), so each one works on
// half of the thread's rows at a time and issues ALL loads of that half -- 16 x 16 bytes per thread, 64 KB in flight per
// SM -- before it touches the first value, and never stores between two loads: xin / out16 / xout are not __restrict__
// (xout really aliases xin in mode 3), so a store in between would serialise the loads at one L2 round trip each.
// Inlined into the epilogue (after its setmaxnreg.inc): the accumulator / residual registers of pass 1 are dead by then.
struct LnRows {
  float mean[4], rstd[4];
};

// rows i = 4 * h + j (j = 0..3) of this thread: global row warp_row0 + i * 4 + (lane >> 3)
__device__ __forceinline__ void ln_half_stats(const float2* stats, long long warp_row0, int rows_valid, int lane, float eps, int h,
                                              LnRows& st) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = (4 * h + j) * 4 + (lane >> 3);
    float s = 0.f, q = 0.f;
    if (r < rows_valid) {
      const float2* sp = stats + (warp_row0 + r) * kLnSlots;
#pragma unroll
      for (int k = 0; k < kLnSlots; ++k) {   // fixed order: bit-reproducible
        const float2 v = __ldcg(sp + k);
        s += v.x;
        q += v.y;
      }
    }
    st.mean[j] = s * (1.0f / kLnD);
    st.rstd[j] = rsqrtf(fmaxf(q * (1.0f / kLnD) - st.mean[j] * st.mean[j], 0.f) + eps);
  }
}

__device__ __forceinline__ void ln_load_half(const float* base, size_t pitch, long long warp_row0, int rows_valid, int col0, int lane,
                                             int h, float4 (&v)[4][4]) {
#pragma unroll
  for (int ci = 0; ci < 4; ++ci)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = (4 * h + j) * 4 + (lane >> 3);
      v[ci][j] = r < rows_valid ? __ldcg(reinterpret_cast<const float4*>(base + static_cast<size_t>(warp_row0 + r) * pitch + col0 + ci * 32))
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
}

// mode 1: out16 = LN(x)
__device__ __forceinline__ void ln_mode1(const GemmParams& p, long long warp_row0, int rows_valid, int col0, int lane) {
  const LnFuse& f = p.ln;
  const float* xin = reinterpret_cast<const float*>(p.out);
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    float4 v[4][4];
    ln_load_half(xin, static_cast<size_t>(p.ldo), warp_row0, rows_valid, col0, lane, h, v);
    LnRows st;
    ln_half_stats(f.stats, warp_row0, rows_valid, lane, f.eps, h, st);
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const int col = col0 + ci * 32;
      const float4 g = __ldg(reinterpret_cast<const float4*>(f.g + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(f.b + col));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = (4 * h + j) * 4 + (lane >> 3);
        if (r < rows_valid)
          *reinterpret_cast<uint2*>(f.out16 + static_cast<size_t>(warp_row0 + r) * kLnD + col) =
              ln_pack4(ln_affine(v[ci][j], st.mean[j], st.rstd[j], g, b));
      }
    }
  }
}

// mode 2: out16 = u = LN(x), rope16 = rotary(u).  The rotary partner of a float4 is the float4 24 columns away inside the
// same 48-wide head (utils.py:83-100); it may belong to another CTA's tile -- visible, because all six slots of these
// rows have been published.
__device__ __forceinline__ void ln_mode2(const GemmParams& p, long long warp_row0, int rows_valid, int col0, int lane) {
  const LnFuse& f = p.ln;
  const float* xin = reinterpret_cast<const float*>(p.out);
  const size_t ldx = static_cast<size_t>(p.ldo);
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    LnRows st;
    ln_half_stats(f.stats, warp_row0, rows_valid, lane, f.eps, h, st);
    int trow[4];           // frame index t = row mod T of each row
#pragma unroll
    for (int j = 0; j < 4; ++j) trow[j] = static_cast<int>((warp_row0 + (4 * h + j) * 4 + (lane >> 3)) % f.T);
#pragma unroll 1
    for (int cp = 0; cp < 2; ++cp) {     // two chunks at a time: 8 own + 8 partner loads in flight
      float4 v[2][4], vp[2][4];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = col0 + (cp * 2 + c) * 32;
        const int colp = (col % 48) < 24 ? col + 24 : col - 24;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = (4 * h + j) * 4 + (lane >> 3);
          const bool ok = r < rows_valid;
          const float* rowp = xin + static_cast<size_t>(warp_row0 + r) * ldx;
          v[c][j] = ok ? __ldcg(reinterpret_cast<const float4*>(rowp + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          vp[c][j] = ok ? __ldcg(reinterpret_cast<const float4*>(rowp + colp)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int col = col0 + (cp * 2 + c) * 32;
        const int q = col % 48;
        const bool lo = q < 24;
        const int colp = lo ? col + 24 : col - 24;
        const float4 g = __ldg(reinterpret_cast<const float4*>(f.g + col));
        const float4 b = __ldg(reinterpret_cast<const float4*>(f.b + col));
        const float4 gp = __ldg(reinterpret_cast<const float4*>(f.g + colp));
        const float4 bp = __ldg(reinterpret_cast<const float4*>(f.b + colp));
        const float sg = lo ? -1.f : 1.f;
        const int qo = lo ? q : q - 24;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int r = (4 * h + j) * 4 + (lane >> 3);
          if (r < rows_valid) {
            const size_t row = static_cast<size_t>(warp_row0 + r);
            const float4 u = ln_affine(v[c][j], st.mean[j], st.rstd[j], g, b);
            const float4 up = ln_affine(vp[c][j], st.mean[j], st.rstd[j], gp, bp);
            const float4 cs = __ldg(reinterpret_cast<const float4*>(f.rope_cos + static_cast<size_t>(trow[j]) * f.half_dim + qo));
            const float4 sn = __ldg(reinterpret_cast<const float4*>(f.rope_sin + static_cast<size_t>(trow[j]) * f.half_dim + qo));
            *reinterpret_cast<uint2*>(f.out16 + row * kLnD + col) = ln_pack4(u);
            *reinterpret_cast<uint2*>(f.rope16 + row * kLnD + col) =
                ln_pack4(make_float4(fmaf(sg * up.x, sn.x, u.x * cs.x), fmaf(sg * up.y, sn.y, u.y * cs.y),
                                     fmaf(sg * up.z, sn.z, u.z * cs.z), fmaf(sg * up.w, sn.w, u.w * cs.w)));
          }
        }
      }
    }
  }
}

// mode 3: xout = LN(x) in fp32 (norm_out, encoder.py:497), then -- unless this is the last layer -- the next layer's
// first LayerNorm of that result, with a second statistics round
__device__ __forceinline__ void ln_mode3(const GemmParams& p, long long warp_row0, int rows_valid, int group, int slot, int col0, int lane) {
  const LnFuse& f = p.ln;
  const float* xin = reinterpret_cast<const float*>(p.out);
  float ps = 0.f, pq = 0.f;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    float4 v[4][4];
    ln_load_half(xin, static_cast<size_t>(p.ldo), warp_row0, rows_valid, col0, lane, h, v);
    LnRows st;
    ln_half_stats(f.stats, warp_row0, rows_valid, lane, f.eps, h, st);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = (4 * h + j) * 4 + (lane >> 3);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int ci = 0; ci < 4; ++ci) {
        const int col = col0 + ci * 32;
        if (r < rows_valid) {
          const float4 y = ln_affine(v[ci][j], st.mean[j], st.rstd[j], __ldg(reinterpret_cast<const float4*>(f.g + col)),
                                     __ldg(reinterpret_cast<const float4*>(f.b + col)));
          *reinterpret_cast<float4*>(f.xout + static_cast<size_t>(warp_row0 + r) * kLnD + col) = y;
          s1 += (y.x + y.y) + (y.z + y.w);
          s2 = fmaf(y.x, y.x, fmaf(y.y, y.y, fmaf(y.z, y.z, fmaf(y.w, y.w, s2))));
        }
      }
      ln_row_add(s1, s2, 4 * h + j, lane, ps, pq);
    }
  }
  if (f.g2 == nullptr) return;
  ln_publish(ps, pq, f.stats2, f.cnt2, warp_row0, rows_valid, group, slot, lane);
  ln_wait(f.cnt2, group, lane);
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    float4 v[4][4];
    ln_load_half(f.xout, kLnD, warp_row0, rows_valid, col0, lane, h, v);
    LnRows st;
    ln_half_stats(f.stats2, warp_row0, rows_valid, lane, f.eps, h, st);
#pragma unroll
    for (int ci = 0; ci < 4; ++ci) {
      const int col = col0 + ci * 32;
      const float4 g = __ldg(reinterpret_cast<const float4*>(f.g2 + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(f.b2 + col));
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = (4 * h + j) * 4 + (lane >> 3);
        if (r < rows_valid)
          *reinterpret_cast<uint2*>(f.out16 + static_cast<size_t>(warp_row0 + r) * kLnD + col) =
              ln_pack4(ln_affine(v[ci][j], st.mean[j], st.rstd[j], g, b));
      }
    }
  }
}

// everything after the first publish: wait for the row statistics, then the normalisation pass(es)
__device__ __forceinline__ void ln_tail(const GemmParams& p, long long warp_row0, int rows_valid, int group, int n_blk, int half,
                                        int lane) {
  const LnFuse& f = p.ln;
  const int slot = 2 * n_blk + half;
  const int col0 = n_blk * 256 + half * 128 + (lane & 7) * 4;   // + 32 * chunk
  if (rows_valid <= 0) {
    if (f.mode == 3 && f.g2 != nullptr && lane == 0) atomicAdd(&f.cnt2[group], 1u);   // keeps the second round's count complete
    return;
  }
  if (!(f.dbg & 1)) ln_wait(f.cnt, group, lane);
  if (f.dbg & 2) return;
  if (f.mode == 1) ln_mode1(p, warp_row0, rows_valid, col0, lane);
  else if (f.mode == 2) ln_mode2(p, warp_row0, rows_valid, col0, lane);
  else ln_mode3(p, warp_row0, rows_valid, group, slot, col0, lane);
}

template <int EPI, int AMODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2Cfg<EPI>::kThreads, 1)
gemm2_f16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                    const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
  constexpr int BN = 256;
  constexpr int kG2Stages = G2Cfg<EPI>::kStages;
  constexpr int kSets = G2Cfg<EPI>::kSets;
  constexpr int kG2EpiBytes = G2Cfg<EPI>::kEpiBytes;
  constexpr uint32_t kTmemCols = 512;
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(256, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kG2Stages * kG2ABytes;
  uint8_t* smem_epi = smem + kG2Stages * kG2StageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + kG2EpiBytes);
  uint64_t* full_bar = bars;                      // [stages]  (leader's copy is the live one)
  uint64_t* empty_bar = bars + kG2Stages;         // [stages]
  uint64_t* tmem_full = bars + 2 * kG2Stages;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;           // [2]  (leader's copy is the live one)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;   // num_m_tiles counts 256-row pair tiles
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_a2);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < kG2Stages; ++s) {
      ptx::mbar_init(&full_bar[s], 2);   // one arrival per CTA's producer (+ the transaction bytes of both)
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 2 * kG2EpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp_idx == 1) ptx::tmem_alloc_2sm<kTmemCols>(tmem_base_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // Two sets: the register file is re-split per WARPGROUP -- all four warps of a warpgroup execute the SAME setmaxnreg
  // instruction (one .dec for warpgroup 0, one .inc for the epilogue warpgroups), each at the top of its own branch,
  // which is also what lets ptxas size the code of that branch for the new limit.
  if (warp_idx < G2Cfg<EPI>::kEpiWarp0) {
   if constexpr (kSets == 2) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
   if (warp_idx == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int m_pair = tile / p.num_n_tiles;
        const int n_blk = tile % p.num_n_tiles;
        const int m_blk = m_pair * 2 + static_cast<int>(rank);   // this CTA's 128-row block
        const CUtensorMap* ta = (p.a1_nblks > 0 && n_blk >= p.a1_nblks) ? &tmap_a2 : &tmap_a;
        int conv_b = 0, conv_t0 = 0;
        if constexpr (AMODE == A_CONV) {
          conv_b = m_blk / p.conv_tiles_per_utt;   // == B for the idle block of an odd count: OOB -> zero fill
          conv_t0 = (m_blk % p.conv_tiles_per_utt) * 8;
        }
        if constexpr (AMODE == A_CONV1D) {
          conv_b = m_blk / p.conv_tiles_per_utt;
          conv_t0 = (m_blk % p.conv_tiles_per_utt) * 128;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * kG2StageBytes);
          else ptx::mbar_arrive_cluster(&full_bar[stage], 0);
          if constexpr (AMODE == A_2D) {
            ptx::tma_load_2d_2sm(smem_a + stage * kG2ABytes, ta, &full_bar[stage], kb * kGemmBK, m_blk * 128);
          } else if constexpr (AMODE == A_CONV1D) {
            const int tap = kb / p.conv_kchunks;
            const int c0 = (kb % p.conv_kchunks) * kGemmBK;
            // 128 output frames t0.. read input frames 2 t + tap - pad: box of 256 input frames traversed with stride 2
            ptx::tma_load_3d_2sm(smem_a + stage * kG2ABytes, &tmap_a, &full_bar[stage], c0, 2 * conv_t0 + tap - p.conv_pad,
                                 conv_b);
          } else {
            const int tap = kb / p.conv_kchunks;
            const int c0 = (kb % p.conv_kchunks) * kGemmBK;
            const int kt = tap / 3, kf = tap % 3;
            ptx::tma_load_4d_2sm(smem_a + stage * kG2ABytes, &tmap_a, &full_bar[stage], c0, kf - 1, 2 * conv_t0 + kt - 1,
                                 conv_b);
          }
          ptx::tma_load_2d_2sm(smem_b + stage * kG2BBytes, &tmap_w, &full_bar[stage], kb * kGemmBK,
                               n_blk * BN + static_cast<int>(rank) * 128);
          if (++stage == kG2Stages) { stage = 0; phase ^= 1; }
        }
      }
    }
   } else if (warp_idx == 1) {
    // ===================================================== MMA issuer (leader CTA only)
    if (leader) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t a_addr = ptx::smem_u32(smem_a + stage * kG2ABytes);
            const uint32_t b_addr = ptx::smem_u32(smem_b + stage * kG2BBytes);
#pragma unroll
            for (int k = 0; k < kGemmBK / 16; ++k) {
              const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
              const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
              ptx::mma_f16_ss_2sm(tmem_d, da, db, kIdesc, (kb | k) != 0 ? 1u : 0u);
            }
            ptx::mma_commit_2sm(&empty_bar[stage], 3);   // frees the slot in BOTH CTAs
            if (kb == p.num_k_blocks - 1) ptx::mma_commit_2sm(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == kG2Stages) { stage = 0; phase ^= 1; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
   }   // (two sets: warps 2 and 3 of warpgroup 0 have no role)
  } else {
    // ===================================================== epilogue (warps 2..9, or 4..19 with two sets; both CTAs)
    if constexpr (kSets == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int quad = warp_idx & 3;          // TMEM lane quadrant
    const int lane = threadIdx.x & 31;
    const int ewg = warp_idx - G2Cfg<EPI>::kEpiWarp0;   // epilogue warp index over all sets: owns one staging tile
    const int set = ewg >> 3;               // kSets == 2: this warp's set drains accumulator `set` (tiles of that local parity)
    const int ew = ewg & 7;
    const int half = ew >> 2;               // which 128 accumulator columns (GLU: which 64 value/gate columns)
    float* stg = reinterpret_cast<float*>(smem_epi + ewg * (kG2WarpStage + kG2WarpBias));
    float* bias_s = reinterpret_cast<float*>(smem_epi + ewg * (kG2WarpStage + kG2WarpBias) + kG2WarpStage);
    int acc = kSets == 2 ? set : 0;
    uint32_t acc_phase = 0;
    for (int tile = pair + (kSets == 2 ? set * npairs : 0); tile < num_tiles; tile += kSets * npairs) {
      const int m_pair = tile / p.num_n_tiles;
      const int n_blk = tile % p.num_n_tiles;
      const int m_blk = m_pair * 2 + static_cast<int>(rank);
      // this warp's 128 bias values -> smem for broadcast reads
      if constexpr (EPI != EPI_POWER_F32) {
        const float* bsrc = p.bias + n_blk * BN;
        float4 bv;
        if constexpr (EPI == EPI_BIAS_GLU_F16) {
          // [0,64) = value bias of columns half*64.., [64,128) = gate bias of the matching columns
          const int o = (lane < 16) ? (half * 64 + lane * 4) : (128 + half * 64 + (lane - 16) * 4);
          bv = __ldg(reinterpret_cast<const float4*>(bsrc + o));
        } else {
          bv = __ldg(reinterpret_cast<const float4*>(bsrc + half * 128 + lane * 4));
        }
        reinterpret_cast<float4*>(bias_s)[lane] = bv;
      }

      long long warp_row0;     // global output row of row 0 of this warp's 32
      int rows_valid;          // how many of the 32 rows exist
      bool row_live = true;    // conv: thread's own row lies inside the utterance's valid length
      [[maybe_unused]] int rows_live = 32;   // conv1d: how many of the warp's 32 rows are inside the valid length
      if constexpr (AMODE == A_2D) {
        warp_row0 = static_cast<long long>(m_blk) * 128 + quad * 32;
        const long long rem = static_cast<long long>(p.M) - warp_row0;
        rows_valid = rem >= 32 ? 32 : (rem > 0 ? static_cast<int>(rem) : 0);
      } else if constexpr (AMODE == A_CONV1D) {
        const bool blk_ok = m_blk < p.conv_num_blocks;
        const int b = blk_ok ? m_blk / p.conv_tiles_per_utt : 0;
        const int t0 = (m_blk % p.conv_tiles_per_utt) * 128 + quad * 32;   // this warp: 32 consecutive output frames
        warp_row0 = static_cast<long long>(b) * p.conv_T2 + t0;
        const int tv = p.conv_T2 - t0;
        rows_valid = blk_ok ? (tv >= 32 ? 32 : (tv > 0 ? tv : 0)) : 0;
        const int lv = blk_ok ? p.conv_len2[b] - t0 : 0;
        rows_live = lv >= 32 ? 32 : (lv > 0 ? lv : 0);
        row_live = lane < rows_live;
      } else {
        const bool blk_ok = m_blk < p.conv_num_blocks;   // odd block count: the pair's second CTA idles on the last tile
        const int b = blk_ok ? m_blk / p.conv_tiles_per_utt : 0;
        const int t0 = (m_blk % p.conv_tiles_per_utt) * 8 + quad * 2;   // this warp: 2 time steps x 16 freq bins
        warp_row0 = (static_cast<long long>(b) * p.conv_T2 + t0) * 16;
        const int tv = p.conv_T2 - t0;
        rows_valid = blk_ok ? (tv >= 2 ? 32 : (tv > 0 ? 16 : 0)) : 0;
        row_live = blk_ok && (t0 + (lane >> 4)) < p.conv_len2[b];
      }

      constexpr bool kRes = EPI == EPI_BIAS_RES_F32 || EPI == EPI_BIAS_RES_LN_F32;
      constexpr bool kLn = EPI == EPI_BIAS_RES_LN_F32;
      [[maybe_unused]] const int c4 = (lane & 7) * 4;
      if constexpr (kRes) {
        // ---- fp32 residual epilogue (two sets, 112 registers): x = res + scale * (acc + bias), residual of the next
        // 32-column chunk in flight while this one is transposed and stored; accumulator read 16 columns at a time
        float* outp = reinterpret_cast<float*>(p.out);
        float4 rr[2][8];
        float ps = 0.f, pq = 0.f;                 // kLn: row sums of the row this lane publishes
        {
          const size_t col = static_cast<size_t>(n_blk) * BN + half * 128 + c4;
#pragma unroll
          for (int i = 0; i < 8; ++i) {           // first chunk's residual goes out before we even wait for the accumulator
            const int r = i * 4 + (lane >> 3);
            rr[0][i] = r < rows_valid ? *reinterpret_cast<const float4*>(p.res + static_cast<size_t>(warp_row0 + r) * p.ldo + col)
                                      : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        ptx::mbar_wait(&tmem_full[acc], acc_phase);
        ptx::tc_fence_after();
        __syncwarp();
        const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = half * 128 + ci * 32;
          const size_t col = static_cast<size_t>(n_blk) * BN + c + c4;
          if (ci + 1 < 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3);
              rr[(ci + 1) & 1][i] = r < rows_valid
                                        ? *reinterpret_cast<const float4*>(p.res + static_cast<size_t>(warp_row0 + r) * p.ldo + col + 32)
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          float4* srow = reinterpret_cast<float4*>(stg + lane * 36);
#pragma unroll
          for (int hv = 0; hv < 2; ++hv) {
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(taddr + c + hv * 16, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int q = 0; q < 4; ++q)
              srow[hv * 4 + q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                             __uint_as_float(v[4 * q + 3]));
          }
          __syncwarp();
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + ci * 32 + c4);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            float s1 = 0.f, s2 = 0.f;
            if (r < rows_valid) {
              float4 a = *reinterpret_cast<const float4*>(stg + r * 36 + c4);
              const float4 x = rr[ci & 1][i];
              a.x = fmaf(p.scale, a.x + bv.x, x.x); a.y = fmaf(p.scale, a.y + bv.y, x.y);
              a.z = fmaf(p.scale, a.z + bv.z, x.z); a.w = fmaf(p.scale, a.w + bv.w, x.w);
              *reinterpret_cast<float4*>(outp + static_cast<size_t>(warp_row0 + r) * p.ldo + col) = a;
              s1 = (a.x + a.y) + (a.z + a.w);
              s2 = fmaf(a.x, a.x, fmaf(a.y, a.y, fmaf(a.z, a.z, a.w * a.w)));
            }
            if constexpr (kLn) ln_row_add(s1, s2, i, lane, ps, pq);
          }
          __syncwarp();
        }
        // the accumulator is drained: hand it back to the MMA warp (before the cross-CTA exchange of the LayerNorm tail)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);
        if constexpr (kLn) {
          ln_publish(ps, pq, p.ln.stats, p.ln.cnt, warp_row0, rows_valid, m_blk * 4 + quad, 2 * n_blk + half, lane);
          ln_tail(p, warp_row0, rows_valid, m_blk * 4 + quad, n_blk, half, lane);
        }
        acc_phase ^= 1;                             // this set always drains the same accumulator
        continue;
      }

      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      __syncwarp();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);

      if constexpr (EPI == EPI_BIAS_F32 || EPI == EPI_CONV_RELU_MASK_F32) {
        float* outp = reinterpret_cast<float*>(p.out);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = half * 128 + ci * 32;
          const size_t col = static_cast<size_t>(n_blk) * BN + c + c4;
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(taddr + c, v);
          ptx::tmem_ld_wait();
          float4* srow = reinterpret_cast<float4*>(stg + lane * 36);
#pragma unroll
          for (int q = 0; q < 8; ++q)
            srow[q] = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                  __uint_as_float(v[4 * q + 3]));
          __syncwarp();
          const float4 bv = *reinterpret_cast<const float4*>(bias_s + ci * 32 + c4);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            if (r < rows_valid) {
              float4 a = *reinterpret_cast<const float4*>(stg + r * 36 + c4);
              a.x += bv.x; a.y += bv.y; a.z += bv.z; a.w += bv.w;
              if constexpr (EPI == EPI_CONV_RELU_MASK_F32) {
                const bool lv = r < rows_live;
                a.x = lv ? fmaxf(a.x, 0.f) : 0.f; a.y = lv ? fmaxf(a.y, 0.f) : 0.f;
                a.z = lv ? fmaxf(a.z, 0.f) : 0.f; a.w = lv ? fmaxf(a.w, 0.f) : 0.f;
              }
              *reinterpret_cast<float4*>(outp + static_cast<size_t>(warp_row0 + r) * p.ldo + col) = a;
            }
          }
          __syncwarp();
        }
      } else if constexpr (EPI == EPI_POWER_F32) {
        // |X|^2 of a DFT whose cos rows fill accumulator columns [0,128) and sin rows [128,256) of the tile
        float* outp = reinterpret_cast<float*>(p.out);
#pragma unroll 1
        for (int ci = 0; ci < 2; ++ci) {
          const int c = half * 64 + ci * 32;
          uint32_t va[32], vb[32];
          ptx::tmem_ld_32x32b_x32(taddr + c, va);
          ptx::tmem_ld_32x32b_x32(taddr + 128 + c, vb);
          ptx::tmem_ld_wait();
          float4* srow = reinterpret_cast<float4*>(stg + lane * 36);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float4 o;
            o.x = __uint_as_float(va[4 * q]) * __uint_as_float(va[4 * q]) + __uint_as_float(vb[4 * q]) * __uint_as_float(vb[4 * q]);
            o.y = __uint_as_float(va[4 * q + 1]) * __uint_as_float(va[4 * q + 1]) + __uint_as_float(vb[4 * q + 1]) * __uint_as_float(vb[4 * q + 1]);
            o.z = __uint_as_float(va[4 * q + 2]) * __uint_as_float(va[4 * q + 2]) + __uint_as_float(vb[4 * q + 2]) * __uint_as_float(vb[4 * q + 2]);
            o.w = __uint_as_float(va[4 * q + 3]) * __uint_as_float(va[4 * q + 3]) + __uint_as_float(vb[4 * q + 3]) * __uint_as_float(vb[4 * q + 3]);
            o.x *= p.scale; o.y *= p.scale; o.z *= p.scale; o.w *= p.scale;   // undo the operand pre-scaling
            srow[q] = o;
          }
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            if (r < rows_valid)
              *reinterpret_cast<float4*>(outp + static_cast<size_t>(warp_row0 + r) * p.ldo + n_blk * 128 + c + c4) =
                  *reinterpret_cast<const float4*>(stg + r * 36 + c4);
          }
          __syncwarp();
        }
      } else if constexpr (EPI == EPI_BIAS_GLU_F16) {
        __half* outp = reinterpret_cast<__half*>(p.out);
        uint32_t* stw = reinterpret_cast<uint32_t*>(stg);
#pragma unroll 1
        for (int ci = 0; ci < 2; ++ci) {
          const int c = half * 64 + ci * 32;   // value columns c.., gate columns 128 + c..
          uint32_t va[32], vb[32];
          ptx::tmem_ld_32x32b_x32(taddr + c, va);
          ptx::tmem_ld_32x32b_x32(taddr + 128 + c, vb);
          ptx::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 ba = *reinterpret_cast<const float4*>(bias_s + ci * 32 + j);
            const float4 bb = *reinterpret_cast<const float4*>(bias_s + 64 + ci * 32 + j);
            const float g0 = (__uint_as_float(va[j]) + ba.x) * sigmoid_f(__uint_as_float(vb[j]) + bb.x);
            const float g1 = (__uint_as_float(va[j + 1]) + ba.y) * sigmoid_f(__uint_as_float(vb[j + 1]) + bb.y);
            const float g2 = (__uint_as_float(va[j + 2]) + ba.z) * sigmoid_f(__uint_as_float(vb[j + 2]) + bb.z);
            const float g3 = (__uint_as_float(va[j + 3]) + ba.w) * sigmoid_f(__uint_as_float(vb[j + 3]) + bb.w);
            pk[j >> 1] = pack_half2(g0, g1);
            pk[(j >> 1) + 1] = pack_half2(g2, g3);
          }
          uint4* srow = reinterpret_cast<uint4*>(stw + lane * 20);   // 64 B of data, pitch 80 B
#pragma unroll
          for (int q = 0; q < 4; ++q) srow[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r = i * 8 + (lane >> 2);
            if (r < rows_valid) {
              const uint4 a = *reinterpret_cast<const uint4*>(stw + r * 20 + (lane & 3) * 4);
              *reinterpret_cast<uint4*>(outp + static_cast<size_t>(warp_row0 + r) * p.ldo + n_blk * 128 + c + (lane & 3) * 8) = a;
            }
          }
          __syncwarp();
        }
      } else {
        // fp16 outputs, 64 accumulator columns per pass (one full 128-byte output row segment per thread)
        __half* outp = reinterpret_cast<__half*>(p.out);
        uint32_t* stw = reinterpret_cast<uint32_t*>(stg);
#pragma unroll 1
        for (int ci = 0; ci < 2; ++ci) {
          const int c = half * 128 + ci * 64;
          uint32_t v[64];
          ptx::tmem_ld_32x32b_x32(taddr + c, v);
          ptx::tmem_ld_32x32b_x32(taddr + c + 32, v + 32);
          ptx::tmem_ld_wait();
          uint32_t pk[32];
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 bv = *reinterpret_cast<const float4*>(bias_s + ci * 64 + j);
            float x0 = __uint_as_float(v[j]) + bv.x, x1 = __uint_as_float(v[j + 1]) + bv.y;
            float x2 = __uint_as_float(v[j + 2]) + bv.z, x3 = __uint_as_float(v[j + 3]) + bv.w;
            if constexpr (EPI == EPI_BIAS_SILU_F16) { x0 = silu_f(x0); x1 = silu_f(x1); x2 = silu_f(x2); x3 = silu_f(x3); }
            if constexpr (EPI == EPI_CONV_RELU_MASK_F16) {
              x0 = row_live ? fmaxf(x0, 0.f) : 0.f; x1 = row_live ? fmaxf(x1, 0.f) : 0.f;
              x2 = row_live ? fmaxf(x2, 0.f) : 0.f; x3 = row_live ? fmaxf(x3, 0.f) : 0.f;
            }
            pk[j >> 1] = pack_half2(x0, x1);
            pk[(j >> 1) + 1] = pack_half2(x2, x3);
          }
          uint4* srow = reinterpret_cast<uint4*>(stw + lane * 36);
#pragma unroll
          for (int q = 0; q < 8; ++q) srow[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = i * 4 + (lane >> 3);
            if (r < rows_valid) {
              const uint4 a = *reinterpret_cast<const uint4*>(stw + r * 36 + (lane & 7) * 4);
              *reinterpret_cast<uint4*>(outp + static_cast<size_t>(warp_row0 + r) * p.ldo + n_blk * BN + c + (lane & 7) * 8) = a;
            }
          }
          __syncwarp();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_2sm<kTmemCols>(tmem_base);
  }
}

}  // namespace gam
