// Persistent warp-specialised tcgen05 GEMM for sm_100a:   D[M,N] = A[M,K] * W[N,K]^T  (+ fused epilogue)
//
//   * A and W are fp16, K-contiguous ("TN"): exactly the layout of an activation matrix [rows, feat]
//     and of an nn.Linear / Conv1d(k=1) weight [out, in] (reference: gigaam/encoder.py:145-148,
//     378,393,418-420), so no operand is ever transposed in memory.
//   * TMA (cp.async.bulk.tensor, SWIZZLE_128B) -> smem ring -> tcgen05.mma (M=128, N=BN, K=16) ->
//     fp32 accumulators in TMEM (double buffered) -> tcgen05.ld -> fused epilogue -> global.
//   * warp 0 = TMA producer, warp 1 = MMA issuer (+ TMEM owner), warps 2..5 = epilogue.
//   * persistent: grid = min(#tiles, #SMs); tile t -> (m = t / n_tiles, n = t % n_tiles) so CTAs
//     running together share the A row-panel through L2.
//   * A_CONV mode: the A operand is the implicit im2col of a channels-last activation
//     [B, T1, F1, C] for a 3x3 / stride-2 / pad-1 convolution (reference: gigaam/encoder.py:59-70),
//     fetched tap by tap with a 4-D strided TMA box (elementStrides = 2 on T and F, OOB = zero fill
//     = the conv's zero padding).  Nothing is ever materialised as an im2col matrix.
#pragma once
#include "ptx.cuh"

namespace gam {

enum GemmEpilogue : int {
  EPI_BIAS_F16 = 0,        // out16 = acc + bias
  EPI_BIAS_SILU_F16 = 1,   // out16 = silu(acc + bias)
  EPI_BIAS_GLU_F16 = 2,    // out16[:, n] = (acc_a + bias_a) * sigmoid(acc_b + bias_b), tile = [a|b]
  EPI_BIAS_RES_F32 = 3,    // out32 = res + scale * (acc + bias)
  EPI_BIAS_F32 = 4,        // out32 = acc + bias
  EPI_CONV_RELU_MASK_F16 = 5,  // out16 = t2 < len2[b] ? relu(acc + bias) : 0   (A_CONV / A_CONV1D row mapping)
  EPI_CONV_RELU_MASK_F32 = 6,  // out32 = t < len[b] ? relu(acc + bias) : 0      (A_CONV1D, last subsampling stage)
  EPI_POWER_F32 = 7,           // out32[:, n] = re^2 + im^2, tile = [128 re | 128 im]  (DFT power spectrum, no bias)
};

// A_CONV  : implicit im2col of a 3x3 / stride-2 conv2d over channels-last [B, T1, F1, C]  (4-D strided TMA)
// A_CONV1D: implicit im2col of a k-tap / stride-2 conv1d over time-major [B, T_in, C]      (3-D strided TMA);
//           a 128-row block = 128 consecutive output frames of one utterance
enum GemmAMode : int { A_2D = 0, A_CONV = 1, A_CONV1D = 2 };

struct GemmParams {
  int M;             // valid rows of D (A_2D) ; unused for A_CONV
  int N;             // columns of the accumulator matrix (= rows of W)
  int num_m_tiles;
  int num_n_tiles;
  int num_k_blocks;  // K / 64   (A_CONV: 9 taps * C/64)
  const float* bias;  // [N] in accumulator column order
  const float* res;   // fp32 residual, row pitch ldo (EPI_BIAS_RES_F32)
  void* out;
  int ldo;            // output row pitch in elements
  float scale;
  // A_CONV only
  int conv_T2;            // output time steps per utterance
  int conv_tiles_per_utt; // ceil(T2 / 8)
  int conv_kchunks;       // C / 64
  int conv_num_blocks;    // B * conv_tiles_per_utt  (128-row blocks that exist)
  int conv_pad;           // A_CONV1D: (taps - 1) / 2
  const int* conv_len2;   // [B] valid output time steps
  // pair kernel, A_2D only: n-tiles [0, a1_nblks) read A through tmap_a, the rest through tmap_a2 (0 = tmap_a for all).
  // Lets two GEMMs that share M, K and the output buffer but not the A operand (W_qk on rope(u), W_v on u) run as one launch.
  int a1_nblks;
};

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;
constexpr int kGemmThreads = 192;

template <int BN>
struct GemmSmem {
  static constexpr int kABytes = kGemmBM * kGemmBK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kGemmBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = kStages * kStageBytes + kBarrierBytes + 1024;  // + align slack
};

__device__ __forceinline__ float fast_tanh(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// silu(x) = x * sigmoid(x) = 0.5x * (1 + tanh(0.5x))
__device__ __forceinline__ float silu_f(float x) {
  float h = 0.5f * x;
  return fmaf(h, fast_tanh(h), h);
}
__device__ __forceinline__ float sigmoid_f(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BN, int EPI, int AMODE>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_f16_tn_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                   const GemmParams p) {
  using S = GemmSmem<BN>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = 2 * BN;  // double-buffered accumulator (256 or 512 columns)
  constexpr uint32_t kIdesc = ptx::make_idesc_f16(kGemmBM, BN, 0, 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                              // kStages * 16 KB
  uint8_t* smem_b = smem + kStages * S::kABytes;       // kStages * BN*128 B
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::kStageBytes);
  uint64_t* full_bar = bars;                 // [kStages]
  uint64_t* empty_bar = bars + kStages;      // [kStages]
  uint64_t* tmem_full = bars + 2 * kStages;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;      // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles;

  if (warp_idx == 0 && ptx::elect_one()) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      ptx::mbar_init(&tmem_empty[s], 4);  // one arrive per epilogue warp
    }
    ptx::fence_mbar_init();
  }
  if (warp_idx == 1) ptx::tmem_alloc<kTmemCols>(tmem_base_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  if (warp_idx == 0) {
    // ===================================================== TMA producer
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / p.num_n_tiles;
        const int n_blk = tile % p.num_n_tiles;
        int conv_b = 0, conv_t0 = 0;
        if constexpr (AMODE == A_CONV) {
          conv_b = m_blk / p.conv_tiles_per_utt;
          conv_t0 = (m_blk % p.conv_tiles_per_utt) * 8;
        }
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
          ptx::mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          if constexpr (AMODE == A_2D) {
            ptx::tma_load_2d(smem_a + stage * S::kABytes, &tmap_a, &full_bar[stage], kb * kGemmBK,
                             m_blk * kGemmBM);
          } else {
            const int tap = kb / p.conv_kchunks;
            const int c0 = (kb % p.conv_kchunks) * kGemmBK;
            const int kt = tap / 3, kf = tap % 3;
            ptx::tma_load_4d(smem_a + stage * S::kABytes, &tmap_a, &full_bar[stage], c0, kf - 1,
                             2 * conv_t0 + kt - 1, conv_b);
          }
          // W descriptors carry 128-row boxes (shared with the CTA-pair kernel): BN / 128 loads per stage
#pragma unroll
          for (int h = 0; h < BN / 128; ++h)
            ptx::tma_load_2d(smem_b + stage * S::kBBytes + h * 128 * 128, &tmap_w, &full_bar[stage], kb * kGemmBK,
                             n_blk * BN + h * 128);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================================================== MMA issuer
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          const uint32_t a_addr = ptx::smem_u32(smem_a + stage * S::kABytes);
          const uint32_t b_addr = ptx::smem_u32(smem_b + stage * S::kBBytes);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            const uint64_t da = ptx::make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = ptx::make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            ptx::mma_f16_ss(tmem_d, da, db, kIdesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::mma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (kb == p.num_k_blocks - 1) ptx::mma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == kStages) { stage = 0; phase ^= 1; }
      }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================================================== epilogue (warps 2..5)
    const int quad = warp_idx & 3;  // TMEM lane quadrant this warp may read
    const int lane = threadIdx.x & 31;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.num_n_tiles;
      const int n_blk = tile % p.num_n_tiles;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quad * 32) << 16);
      const int r_in_tile = quad * 32 + lane;

      long long out_row;
      bool row_valid;
      bool row_live = true;  // A_CONV: time step inside the utterance's valid length
      if constexpr (AMODE == A_2D) {
        out_row = static_cast<long long>(m_blk) * kGemmBM + r_in_tile;
        row_valid = out_row < p.M;
      } else {
        const int b = m_blk / p.conv_tiles_per_utt;
        const int t2 = (m_blk % p.conv_tiles_per_utt) * 8 + (r_in_tile >> 4);
        row_valid = t2 < p.conv_T2;
        row_live = t2 < p.conv_len2[b];
        out_row = (static_cast<long long>(b) * p.conv_T2 + t2) * 16 + (r_in_tile & 15);
      }

      if constexpr (EPI == EPI_BIAS_GLU_F16) {
        constexpr int kHalf = BN / 2;
        __half* out = reinterpret_cast<__half*>(p.out) + out_row * p.ldo + n_blk * kHalf;
        const float* bias = p.bias + n_blk * BN;
#pragma unroll 1
        for (int c = 0; c < kHalf; c += 32) {
          uint32_t va[32], vb[32];
          ptx::tmem_ld_32x32b_x32(taddr + c, va);
          ptx::tmem_ld_32x32b_x32(taddr + kHalf + c, vb);
          ptx::tmem_ld_wait();
          if (row_valid) {
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float a0 = __uint_as_float(va[j]) + __ldg(bias + c + j);
              float a1 = __uint_as_float(va[j + 1]) + __ldg(bias + c + j + 1);
              float b0 = __uint_as_float(vb[j]) + __ldg(bias + kHalf + c + j);
              float b1 = __uint_as_float(vb[j + 1]) + __ldg(bias + kHalf + c + j + 1);
              pk[j >> 1] = pack_half2(a0 * sigmoid_f(b0), a1 * sigmoid_f(b1));
            }
            uint4* dst = reinterpret_cast<uint4*>(out + c);
#pragma unroll
            for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
      } else {
        const float* bias = p.bias + n_blk * BN;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
          uint32_t v[32];
          ptx::tmem_ld_32x32b_x32(taddr + c, v);
          ptx::tmem_ld_wait();
          if (row_valid) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + __ldg(bias + c + j);
            if constexpr (EPI == EPI_BIAS_F16 || EPI == EPI_BIAS_SILU_F16 || EPI == EPI_CONV_RELU_MASK_F16) {
              uint32_t pk[16];
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                float x0 = f[j], x1 = f[j + 1];
                if constexpr (EPI == EPI_BIAS_SILU_F16) { x0 = silu_f(x0); x1 = silu_f(x1); }
                if constexpr (EPI == EPI_CONV_RELU_MASK_F16) {
                  x0 = row_live ? fmaxf(x0, 0.f) : 0.f;
                  x1 = row_live ? fmaxf(x1, 0.f) : 0.f;
                }
                pk[j >> 1] = pack_half2(x0, x1);
              }
              __half* out = reinterpret_cast<__half*>(p.out) + out_row * p.ldo + n_blk * BN + c;
              uint4* dst = reinterpret_cast<uint4*>(out);
#pragma unroll
              for (int q = 0; q < 4; ++q) dst[q] = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            } else {
              float* out = reinterpret_cast<float*>(p.out) + out_row * p.ldo + n_blk * BN + c;
              if constexpr (EPI == EPI_BIAS_RES_F32) {
                const float4* res = reinterpret_cast<const float4*>(p.res + out_row * p.ldo + n_blk * BN + c);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                  float4 r = res[q];
                  f[4 * q + 0] = fmaf(p.scale, f[4 * q + 0], r.x);
                  f[4 * q + 1] = fmaf(p.scale, f[4 * q + 1], r.y);
                  f[4 * q + 2] = fmaf(p.scale, f[4 * q + 2], r.z);
                  f[4 * q + 3] = fmaf(p.scale, f[4 * q + 3], r.w);
                }
              }
              float4* dst = reinterpret_cast<float4*>(out);
#pragma unroll
              for (int q = 0; q < 8; ++q) dst[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
            }
          }
        }
      }
      // accumulator drained -> hand the TMEM buffer back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp_idx == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kTmemCols>(tmem_base);
  }
}

}  // namespace gam
