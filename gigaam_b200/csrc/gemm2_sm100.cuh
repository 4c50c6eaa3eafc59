// CTA-pair (cta_group::2) tcgen05 GEMM for sm_100a:  D[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
// Two CTAs on the two SMs of a TPC compute one 256 x 256 tile: each CTA stages its own 128 rows of A and HALF
// (128 rows) of the W tile, the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), and each CTA ends up with
// the accumulators of its 128 rows (x 256 columns) in its own TMEM.  Per k-block each SM moves
// 16 KB (A) + 16 KB (W half) through shared memory instead of 16 + 32 KB: the single-CTA kernel
// of round 1 was bound by shared-memory / L2 operand traffic at ~50 % tensor-pipe.
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

constexpr int kG2Threads = 320;
constexpr int kG2EpiWarps = 8;
constexpr int kG2Stages = 5;
constexpr int kG2ABytes = 128 * 64 * 2;       // 16 KB: this CTA's 128 rows of A
constexpr int kG2BBytes = 128 * 64 * 2;       // 16 KB: this CTA's half of the 256-row W tile
constexpr int kG2StageBytes = kG2ABytes + kG2BBytes;
constexpr int kG2WarpStage = 32 * 36 * 4;     // 4608 B staging tile per epilogue warp
constexpr int kG2WarpBias = 128 * 4;          // per-warp copy of its 128 bias values
constexpr int kG2EpiBytes = kG2EpiWarps * (kG2WarpStage + kG2WarpBias);
constexpr int kG2BarBytes = 256;
constexpr int kG2Smem = kG2Stages * kG2StageBytes + kG2EpiBytes + kG2BarBytes + 1024;

template <int EPI, int AMODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kG2Threads, 1)
gemm2_f16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                    const __grid_constant__ CUtensorMap tmap_w, const GemmParams p) {
  constexpr int BN = 256;
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
  // packed rows: the row count of an A_2D product may live on the device (GemmParams::m_dev); every role derives the same
  // tile list from it, and CTA pairs the host sized for the padded maximum simply find no tile
  int m_rows = p.M, num_m_tiles = p.num_m_tiles;          // num_m_tiles counts 256-row pair tiles
  if constexpr (AMODE == A_2D) {
    if (p.m_dev != nullptr) {
      m_rows = min(max(__ldg(p.m_dev), 0), p.M);
      num_m_tiles = (m_rows + 255) >> 8;
    }
  }
  const int num_tiles = num_m_tiles * p.num_n_tiles;
  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;
  // conv modes with packed output: a 128-row block whose first frame lies past the utterance's length produces nothing,
  // and a pair tile made of two such blocks is skipped by all three roles (same predicate, same tile list)
  [[maybe_unused]] auto conv_tile_dead = [&](int m_pair) -> bool {
    if constexpr (AMODE == A_2D) {
      return false;
    } else {
      if (p.conv_cu == nullptr) return false;
      constexpr int kFramesPerBlock = (AMODE == A_CONV) ? 8 : 128;
      bool dead = true;
#pragma unroll
      for (int hblk = 0; hblk < 2; ++hblk) {
        const int mb = m_pair * 2 + hblk;
        if (mb < p.conv_num_blocks)
          dead = dead && (mb % p.conv_tiles_per_utt) * kFramesPerBlock >= __ldg(p.conv_plen + mb / p.conv_tiles_per_utt);
      }
      return dead;
    }
  };

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

  if (warp_idx == 0) {
    // ===================================================== TMA producer (both CTAs)
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair; tile < num_tiles; tile += npairs) {
        const int tl = p.reverse ? num_tiles - 1 - tile : tile;
        const int m_pair = tl / p.num_n_tiles;
        const int n_blk = tl % p.num_n_tiles;
        if (conv_tile_dead(m_pair)) continue;
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
        if constexpr (AMODE != A_2D) {
          if (conv_tile_dead((p.reverse ? num_tiles - 1 - tile : tile) / p.num_n_tiles)) continue;
        }
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
  } else {
    // ===================================================== epilogue (warps 2..9, both CTAs)
    const int quad = warp_idx & 3;          // TMEM lane quadrant
    const int lane = threadIdx.x & 31;
    const int ew = warp_idx - 2;
    const int half = ew >> 2;               // which 128 accumulator columns (GLU: which 64 value/gate columns)
    float* stg = reinterpret_cast<float*>(smem_epi + ew * (kG2WarpStage + kG2WarpBias));
    float* bias_s = reinterpret_cast<float*>(smem_epi + ew * (kG2WarpStage + kG2WarpBias) + kG2WarpStage);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = pair; tile < num_tiles; tile += npairs) {
      const int tl = p.reverse ? num_tiles - 1 - tile : tile;
      const int m_pair = tl / p.num_n_tiles;
      const int n_blk = tl % p.num_n_tiles;
      if (conv_tile_dead(m_pair)) continue;
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
        const long long rem = static_cast<long long>(m_rows) - warp_row0;
        rows_valid = rem >= 32 ? 32 : (rem > 0 ? static_cast<int>(rem) : 0);
      } else if constexpr (AMODE == A_CONV1D) {
        const bool blk_ok = m_blk < p.conv_num_blocks;
        const int b = blk_ok ? m_blk / p.conv_tiles_per_utt : 0;
        const int t0 = (m_blk % p.conv_tiles_per_utt) * 128 + quad * 32;   // this warp: 32 consecutive output frames
        const bool packed = p.conv_cu != nullptr;
        warp_row0 = (packed ? static_cast<long long>(__ldg(p.conv_cu + b)) : static_cast<long long>(b) * p.conv_T2) + t0;
        const int tv = (packed ? min(__ldg(p.conv_plen + b), p.conv_T2) : p.conv_T2) - t0;
        rows_valid = blk_ok ? (tv >= 32 ? 32 : (tv > 0 ? tv : 0)) : 0;
        const int lv = blk_ok ? p.conv_len2[b] - t0 : 0;
        rows_live = lv >= 32 ? 32 : (lv > 0 ? lv : 0);
        row_live = lane < rows_live;
      } else {
        const bool blk_ok = m_blk < p.conv_num_blocks;   // odd block count: the pair's second CTA idles on the last tile
        const int b = blk_ok ? m_blk / p.conv_tiles_per_utt : 0;
        const int t0 = (m_blk % p.conv_tiles_per_utt) * 8 + quad * 2;   // this warp: 2 time steps x 16 freq bins
        const bool packed = p.conv_cu != nullptr;
        warp_row0 = ((packed ? static_cast<long long>(__ldg(p.conv_cu + b)) : static_cast<long long>(b) * p.conv_T2) + t0) * 16;
        const int tv = (packed ? min(__ldg(p.conv_plen + b), p.conv_T2) : p.conv_T2) - t0;
        rows_valid = blk_ok ? (tv >= 2 ? 32 : (tv > 0 ? 16 : 0)) : 0;
        row_live = blk_ok && (t0 + (lane >> 4)) < p.conv_len2[b];
      }

      [[maybe_unused]] float4 rr[2][8];
      [[maybe_unused]] const int c4 = (lane & 7) * 4;
      if constexpr (EPI == EPI_BIAS_RES_F32) {
        // first chunk's residual goes out before we even wait for the accumulator
        const size_t col = static_cast<size_t>(n_blk) * BN + half * 128 + c4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = i * 4 + (lane >> 3);
          rr[0][i] = r < rows_valid ? *reinterpret_cast<const float4*>(p.res + static_cast<size_t>(warp_row0 + r) * p.ldo + col)
                                    : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // and the residual of this warp's NEXT tile is pulled into L2 a whole main loop ahead (no registers involved): the
        // residual stream is fp32 and mostly DRAM-resident, and 8 epilogue warps cannot cover DRAM latency with the two
        // chunks (16 loads per thread) they keep in flight -- the N = 768 GEMMs are bound by exactly these loads
        if constexpr (AMODE == A_2D) {
          if (tile + npairs < num_tiles) {
            const int nt = p.reverse ? num_tiles - 1 - (tile + npairs) : tile + npairs;
            const long long row = (static_cast<long long>(nt / p.num_n_tiles) * 2 + static_cast<long long>(rank)) * 128 + quad * 32 + lane;
            if (row < m_rows) {
              const float* src = p.res + static_cast<size_t>(row) * p.ldo + static_cast<size_t>(nt % p.num_n_tiles) * BN + half * 128;
#pragma unroll
              for (int j = 0; j < 4; ++j) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + j * 32));
            }
          }
        }
      }

      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      __syncwarp();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);

      if constexpr (EPI == EPI_BIAS_RES_F32 || EPI == EPI_BIAS_F32 || EPI == EPI_CONV_RELU_MASK_F32) {
        float* outp = reinterpret_cast<float*>(p.out);
#pragma unroll
        for (int ci = 0; ci < 4; ++ci) {
          const int c = half * 128 + ci * 32;
          const size_t col = static_cast<size_t>(n_blk) * BN + c + c4;
          if constexpr (EPI == EPI_BIAS_RES_F32) {
            if (ci + 1 < 4) {   // residual of the next chunk in flight while this one is transposed and stored
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int r = i * 4 + (lane >> 3);
                rr[(ci + 1) & 1][i] = r < rows_valid
                                          ? *reinterpret_cast<const float4*>(p.res + static_cast<size_t>(warp_row0 + r) * p.ldo + col + 32)
                                          : make_float4(0.f, 0.f, 0.f, 0.f);
              }
            }
          }
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
              if constexpr (EPI == EPI_BIAS_RES_F32) {
                const float4 x = rr[ci & 1][i];
                a.x = fmaf(p.scale, a.x, x.x); a.y = fmaf(p.scale, a.y, x.y);
                a.z = fmaf(p.scale, a.z, x.z); a.w = fmaf(p.scale, a.w, x.w);
              }
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
