// Instantiations and host launchers of the tcgen05 GEMM (gemm_sm100.cuh).
#include "gemm_sm100.cuh"
#include "kernels.h"

namespace gam {
namespace {

constexpr int kBN = 256;

template <int EPI, int AMODE>
int launch_one(const CUtensorMap* ta, const CUtensorMap* tw, const GemmParams& p, int num_sms, cudaStream_t s) {
  auto kern = gemm_f16_tn_kernel<kBN, EPI, AMODE>;
  const int tiles = p.num_m_tiles * p.num_n_tiles;
  const int grid = tiles < num_sms ? tiles : num_sms;
  if (grid <= 0) return 0;
  kern<<<grid, kGemmThreads, GemmSmem<kBN>::kTotal, s>>>(*ta, *tw, p);
  return cudaPeekAtLastError() == cudaSuccess ? 0 : -2;
}

template <int EPI, int AMODE>
int set_attr() {
  auto kern = gemm_f16_tn_kernel<kBN, EPI, AMODE>;
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmSmem<kBN>::kTotal) == cudaSuccess ? 0 : -1;
}

}  // namespace

int gemm_init() {
  int rc = 0;
  rc |= set_attr<EPI_BIAS_F16, A_2D>();
  rc |= set_attr<EPI_BIAS_SILU_F16, A_2D>();
  rc |= set_attr<EPI_BIAS_GLU_F16, A_2D>();
  rc |= set_attr<EPI_BIAS_RES_F32, A_2D>();
  rc |= set_attr<EPI_BIAS_F32, A_2D>();
  rc |= set_attr<EPI_CONV_RELU_MASK_F16, A_CONV>();
  return rc;
}

int launch_gemm(int kind, const CUtensorMap* ta, const CUtensorMap* tw, int M, int N, int K, const float* bias,
                const float* res, void* out, int ldo, float scale, int num_sms, cudaStream_t s) {
  if (N % kBN != 0 || K % kGemmBK != 0 || M <= 0) return -1;
  GemmParams p{};
  p.M = M;
  p.N = N;
  p.num_m_tiles = (M + kGemmBM - 1) / kGemmBM;
  p.num_n_tiles = N / kBN;
  p.num_k_blocks = K / kGemmBK;
  p.bias = bias;
  p.res = res;
  p.out = out;
  p.ldo = ldo;
  p.scale = scale;
  switch (kind) {
    case GEMM_BIAS_F16: return launch_one<EPI_BIAS_F16, A_2D>(ta, tw, p, num_sms, s);
    case GEMM_BIAS_SILU_F16: return launch_one<EPI_BIAS_SILU_F16, A_2D>(ta, tw, p, num_sms, s);
    case GEMM_BIAS_GLU_F16: return launch_one<EPI_BIAS_GLU_F16, A_2D>(ta, tw, p, num_sms, s);
    case GEMM_BIAS_RES_F32: return launch_one<EPI_BIAS_RES_F32, A_2D>(ta, tw, p, num_sms, s);
    case GEMM_BIAS_F32: return launch_one<EPI_BIAS_F32, A_2D>(ta, tw, p, num_sms, s);
    default: return -1;
  }
}

int launch_gemm_conv(const CUtensorMap* ta4, const CUtensorMap* tw, int B, int T2, int C, int N, const float* bias,
                     const int* len2, void* out, int ldo, int num_sms, cudaStream_t s) {
  if (N % kBN != 0 || C % kGemmBK != 0) return -1;
  GemmParams p{};
  p.M = 0;
  p.N = N;
  p.conv_T2 = T2;
  p.conv_tiles_per_utt = (T2 + 7) / 8;
  p.conv_kchunks = C / kGemmBK;
  p.conv_len2 = len2;
  p.num_m_tiles = B * p.conv_tiles_per_utt;
  p.num_n_tiles = N / kBN;
  p.num_k_blocks = 9 * p.conv_kchunks;
  p.bias = bias;
  p.res = nullptr;
  p.out = out;
  p.ldo = ldo;
  p.scale = 1.f;
  return launch_one<EPI_CONV_RELU_MASK_F16, A_CONV>(ta4, tw, p, num_sms, s);
}

}  // namespace gam
