// Internal launcher declarations shared by the translation units of libgigaam_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gam {

// rowops.cu
// Rows are packed: `rows` is the padded maximum the grid is sized for, `rows_dev` (may be null) the live count on the device.
// `reverse`: walk the rows from the last to the first (see GemmParams::reverse: consecutive kernels alternate direction)
void launch_ln_f16(const float* x, const float* g, const float* b, __half* out, int rows, const int* rows_dev, int reverse,
                   cudaStream_t s);
// row_t (may be null: row % T): frame index of every packed row inside its utterance = its rotary position
void launch_ln_rope_f16(const float* x, const float* g, const float* b, const float* rope_cos, const float* rope_sin,
                        __half* out_u, __half* out_r, int rows, const int* rows_dev, const int* row_t, int T, int half_dim,
                        int reverse, cudaStream_t s);
void launch_ln_out_ln(const float* r, const float* g_out, const float* b_out, const float* g_next, const float* b_next,
                      float* x_out, __half* y_out, int rows, const int* rows_dev, int reverse, cudaStream_t s);
// packed fp32 rows -> padded [B, T, 768] (LayerNorm on the way when gamma != null); frames that do not exist become zeros
void launch_unpack_rows(const float* x, const float* gamma, const float* beta, const int* cu, const int* plen, float* out, int B,
                        int T, int reverse, cudaStream_t s);
// cu / plen null: padded layout (utterance b at row b*T, T rows)
int launch_dwconv_bn_silu(const __half* g, const float* w, const float* bias, const int* len, const int* cu, const int* plen,
                          __half* out, int B, int T, int kw, cudaStream_t s);
int launch_dwconv_ln_silu(const __half* g, const float* w, const float* bias, const float* gamma, const float* beta,
                          const int* len, const int* cu, const int* row_b, const int* row_t, const int* rows_dev, __half* out,
                          int B, int T, int kw, cudaStream_t s);
// stage lengths of the subsampling + the packed-row plan (plen, cu [B+1], rows_dev [1], run1, row -> (b, t) maps)
void launch_pack_plan(const long long* mel_len, int B, int pad2_minus_k, int max_T0, int T1, int T2, int* len0, int* len1,
                      int* len2, int* plen, int* run1, int* cu, int* rows_dev, int* row_b, int* row_t, cudaStream_t s);

// frontend.cu
int launch_logmel(const float* wav, int B, int n_samples, int n_frames, const float* window, const float* tcos,
                  const float* tsin, const float* fb, float* mel, int n_fft, int hop, int center, int n_mels,
                  cudaStream_t s);
// run1 (may be null): stage-1 frames per utterance that are produced at all (pack_plan_kernel)
int launch_subsample_conv1(const float* mel, const int* len0, const int* len1, const int* run1, const float* w, const float* bias,
                           __half* out, int B, int M, int F, int T1, int F1, int C, cudaStream_t s);

// attention_sm100.cu
// klen (may be null: T): keys / queries of utterance b that exist; cu (may be null): packed rows, utterance b starts at row
// cu[b] and only its klen[b] query rows are computed and stored (null: row b*T, all T query rows stored)
int launch_attention(const CUtensorMap* tmap_qkv, const int* klen, const int* cu, __half* out, int B, int T, int H, int dk,
                     int d_model, int num_sms, cudaStream_t s);

// attention_relpos_sm100.cu: qkv [B*T, 4*d_model] = [q+u | q+v | k | v]; pos = projected position table of
// 2*kRelPosMaxT-1 rows (GAM_REL_POS_MAX_T in the public header)
constexpr int kRelPosMaxT = 768;
int launch_attention_relpos(const CUtensorMap* tmap_qkv, const CUtensorMap* tmap_pos, const int* klen, const int* cu, __half* out,
                            int B, int T, int H, int dk, int d_model, cudaStream_t s);

// ctc.cu
void launch_ctc_argmax(const float* enc, const float* W, const float* bias, int* labels, int R, int D, int V1,
                       cudaStream_t s);
void launch_ctc_collapse(const int* labels, const int* len, int B, int T, int blank, int* ids, int* frames, int* counts,
                         cudaStream_t s);

// words.cu: (token id, frame) pairs -> word records (first frame, last frame + 1, first token, tokens) per utterance
void launch_group_words(const int* ids, const int* frames, const int* counts, const unsigned char* flags, int B, int V, int max_out,
                        int max_words, int* w_start, int* w_end, int* w_first, int* w_ntok, int* n_words, cudaStream_t s);

// rnnt.cu
void launch_sgemm_tn_bias(const float* A, const float* W, const float* bias, float* C, int M, int N, int K, cudaStream_t s);

// rnnt_cluster.cu: returns 0 ok, 1 = 16-CTA clusters unavailable / unsupported shape, <0 error
int launch_rnnt_greedy_cluster(const float* encproj, const int* len, const float* emb_gates, const float* whhT, const float* wpT,
                               const float* bp, const float* wo, const float* bo, int B, int T, int H, int V1, int blank,
                               int max_symbols, int max_out, int* ids, int* frames, int* counts, cudaStream_t s);

// gemm.cu
struct GemmParams;
enum GemmKind : int {
  GEMM_BIAS_F16 = 0,
  GEMM_BIAS_SILU_F16 = 1,
  GEMM_BIAS_GLU_F16 = 2,
  GEMM_BIAS_RES_F32 = 3,
  GEMM_BIAS_F32 = 4,
  GEMM_CONV_RELU_MASK_F16 = 5,
};
// 2-D operand GEMM  D[M,N] = A[M,K] W[N,K]^T with fused epilogue `kind`; N % 256 == 0, K % 64 == 0.
int launch_gemm(int kind, const CUtensorMap* tmap_a, const CUtensorMap* tmap_w, int M, int N, int K, const float* bias,
                const float* res, void* out, int ldo, float scale, int num_sms, cudaStream_t s, int reverse = 0,
                const int* m_dev = nullptr);   // m_dev: row count on the device (<= M), see GemmParams::m_dev
// one launch for two GEMMs that share M, K, W's row space and the output buffer but read different A operands:
// columns [0, n1) from tmap_a1, [n1, N) from tmap_a2 (bias -> fp16).  
int launch_gemm_dual_a(const CUtensorMap* tmap_a1, const CUtensorMap* tmap_a2, int n1, const CUtensorMap* tmap_w, int M, int N,
                       int K, const float* bias, void* out, int ldo, int num_sms, cudaStream_t s, int reverse = 0,
                       const int* m_dev = nullptr);
// implicit-GEMM 3x3/s2 conv over channels-last [B,T1,F1,C] (tmap_a 4-D strided), output [rows*16, N] fp16.
// cu / plen (both or neither): packed output rows, frame (b, t < plen[b]) -> row cu[b] + t; null: row b*T2 + t, all frames
int launch_gemm_conv(const CUtensorMap* tmap_a4d, const CUtensorMap* tmap_w, int B, int T2, int C, int N, const float* bias,
                     const int* len2, const int* cu, const int* plen, void* out, int ldo, int num_sms, cudaStream_t s);
// implicit-GEMM k-tap/s2 conv1d over time-major [B,T_in,C_in] (tmap_a 3-D strided); out [B*T_out, N] fp16 or fp32
int launch_gemm_conv1d(const CUtensorMap* tmap_a3d, const CUtensorMap* tmap_w, int B, int T_out, int C_in, int taps, int N,
                       const float* bias, const int* len_out, const int* cu, const int* plen, void* out, int ldo, int f32_out,
                       int num_sms, cudaStream_t s);
int launch_gemm_power(const CUtensorMap* tmap_a, const CUtensorMap* tmap_w, int M, int N, int K, float* out, int ldo, int num_sms,
                      cudaStream_t s);
// tensor-core front end helpers (frontend.cu)
void launch_frames_split(const float* wav, int B, int n_samples, int n_frames, const float* window, __half* A, int n_fft, int Kp,
                         int hop, int center, cudaStream_t s);
void launch_mel_log(const float* P, int ldp, int B, int n_frames, int nbins, const float* fb, const int* mel_lo, const int* mel_hi,
                    float* mel, int n_mels, cudaStream_t s);
int gemm_init();
// mel [B, F, M] f32 -> time-major fp16 [B, M, F] with frames >= len zeroed (conv1d subsampling input)
void launch_mel_to_tmajor_f16(const float* mel, const int* len0, __half* out, int B, int F, int M, cudaStream_t s);

// tensor maps (gam_api.cu)
int make_tmap_2d_f16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                     uint32_t box_cols);

}  // namespace gam
