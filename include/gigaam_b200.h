/*
 * gigaam_b200 -- C ABI of the B200-native GigaAM hot path (libgigaam_b200.so).
 *
 * The reference (salute-developers/GigaAM) has no FFI of its own: the path is reached through Python
 * nn.Modules.  Each entry point below replaces the torch-op body of one of those modules; the Python
 * mirror classes in gigaam_b200/ (same names, constructor kwargs and state_dict keys as the
 * reference) bind them with ctypes.  See INTEGRATION.md for the reference-side binding.
 *
 *   gam_logmel        <- gigaam/preprocess.py:53-98   FeatureExtractor.forward (MelSpectrogram + log)
 *   gam_encode        <- gigaam/encoder.py:605-647    ConformerEncoder.forward (subsampling + N layers)
 *   gam_ctc_greedy    <- gigaam/decoder.py:18-21 + gigaam/decoding.py:56-96  CTCHead + CTCGreedyDecoding
 *   gam_rnnt_greedy   <- gigaam/decoder.py:41-47,85-102 + gigaam/decoding.py:128-207
 *
 * Conventions: every pointer marked "device" is a CUDA device pointer on the handle's device; the
 * library never allocates or frees caller memory in the hot calls (the caller passes a workspace of
 * gam_workspace_bytes()); all work is enqueued on `stream` (a cudaStream_t passed as void*) and the
 * call returns without synchronising; int return, 0 = ok, negative = error (gam_last_error()).
 * A handle is bound to one device and is not thread-safe.  No CPU fallback exists.
 */
#ifndef GIGAAM_B200_H_
#define GIGAAM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gam_handle gam_handle;

typedef struct gam_config {
  /* preprocessor (gigaam/preprocess.py:60-65) */
  int32_t sample_rate, n_mels, n_fft, win_length, hop_length, center;
  /* encoder (gigaam/encoder.py:510-526) */
  int32_t feat_in, n_layers, d_model, n_heads, d_ff;
  int32_t subsampling;      /* 0 = conv2d, 1 = conv1d */
  int32_t subs_kernel_size; /* 3 (conv2d) */
  int32_t conv_kernel_size; /* depthwise taps: 31 or 5 */
  int32_t conv_norm;        /* 0 = batch_norm (folded into the depthwise conv), 1 = layer_norm */
  int32_t self_attention;   /* 0 = rotary, 1 = rel_pos (v1 checkpoints) */
  int32_t pos_emb_max_len;
  /* head (gigaam/decoder.py) */
  int32_t head;        /* 0 = none (ssl), 1 = ctc, 2 = rnnt */
  int32_t num_classes; /* V + 1, blank id = V */
  int32_t pred_hidden, joint_hidden, max_symbols;
} gam_config;

/* One Conformer layer; all pointers device.  "h" = fp16 row-major [out, in]; "f" = fp32. */
typedef struct gam_layer_weights {
  const float *ln_ff1_g, *ln_ff1_b;
  const void* ff1_w1; /* h [d_ff, d] */
  const float* ff1_b1;
  const void* ff1_w2; /* h [d, d_ff] */
  const float* ff1_b2;
  const float *ln_att_g, *ln_att_b;
  const void* w_qk; /* h [2d, d] = [linear_q ; linear_k] */
  const float* b_qk;
  const void* w_v; /* h [d, d].  When w_v sits directly behind w_qk (one [3d, d] matrix) and b_v directly behind b_qk, the
                    * q, k and v projections run as a single launch */
  const float* b_v;
  const void* w_o; /* h [d, d] */
  const float* b_o;
  const float *ln_conv_g, *ln_conv_b;
  const void* pw1_w; /* h [2d, d], rows permuted so every 256-row tile is [128 value | 128 gate] */
  const float* pw1_b; /* f [2d], same permutation */
  const float* dw_w;  /* f [k, d]  depthwise taps, tap-major (eval BatchNorm folded in when conv_norm == 0) */
  const float* dw_b;  /* f [d] */
  const float *cn_g, *cn_b; /* conv LayerNorm affine (conv_norm == 1), else NULL */
  const void* pw2_w;  /* h [d, d] */
  const float* pw2_b;
  const float *ln_ff2_g, *ln_ff2_b;
  const void* ff2_w1;
  const float* ff2_b1;
  const void* ff2_w2;
  const float* ff2_b2;
  const float *ln_out_g, *ln_out_b;
  /* self_attention == 1 (rel_pos, gigaam/encoder.py:191-228) only, else NULL; w_qk / w_v are then unused */
  const void* w_qkv_rel;  /* h [4d, d] = [linear_q ; linear_q ; linear_k ; linear_v] */
  const float* b_qkv_rel; /* f [4d]    = [b_q + pos_bias_u ; b_q + pos_bias_v ; b_k ; b_v] */
  const void* pos_proj;   /* h [2*GAM_REL_POS_MAX_T-1, d]: linear_pos(pe(r)), row GAM_REL_POS_MAX_T-1-r for relative
                           * position r (pe = gigaam/encoder.py:318-326) */
} gam_layer_weights;

#define GAM_REL_POS_MAX_T 768 /* longest T' the attention kernels serve (6 key blocks of 128 = 30.7 s of audio) */

typedef struct gam_weights {
  /* front end */
  const float* window;  /* f [n_fft]  (checkpoint buffer preprocessor.featurizer.0.spectrogram.window) */
  const float* dft_cos; /* f [n_fft/2+1, n_fft/2+1]  cos(2 pi k n / n_fft), row n, column k */
  const float* dft_sin; /* f [n_fft/2+1, n_fft/2+1]  sin(2 pi k n / n_fft) */
  const float* mel_fb;  /* f [n_fft/2+1, n_mels]     (checkpoint buffer ...mel_scale.fb) */
  /* subsampling (conv2d) */
  const float* sub1_w; /* f [C, 9]       encoder.pre_encode.conv.0.weight */
  const float* sub1_b; /* f [C] */
  const void* sub2_w;  /* h [C, 9*C]     conv.2.weight permuted to (out, kt, kf, in) */
  const float* sub2_b; /* f [C] */
  const void* sub_out_w; /* h [d, F2*C]  pre_encode.out.weight with K permuted from (c, f) to (f, c) */
  const float* sub_out_b;
  /* rotary tables f [pos_emb_max_len, d_k/2] */
  const float* rope_cos;
  const float* rope_sin;
  const gam_layer_weights* layers; /* HOST array of n_layers structs holding device pointers */
  /* CTC head, fp32 */
  const float* ctc_w; /* f [V+1, d] */
  const float* ctc_b;
  /* RNN-T head, fp32 */
  const float* rnnt_enc_w;     /* f [joint_hidden, d]   joint.enc.weight */
  const float* rnnt_enc_b;
  const float* rnnt_emb_gates; /* f [V+1, 4H]  embed(k) W_ih^T + b_ih + b_hh */
  const float* rnnt_whh_t;     /* f [H, 4H]    lstm.weight_hh_l0^T */
  const float* rnnt_wp_t;      /* f [H, joint_hidden]  joint.pred.weight^T */
  const float* rnnt_bp;
  const float* rnnt_wo;        /* f [V+1, joint_hidden] joint.joint_net.1.weight */
  const float* rnnt_bo;
  /* subsampling (conv1d, v3 checkpoints): weights permuted to (out, tap, in) */
  const void* c1d_w1;  /* h [d, k * feat_in]  encoder.pre_encode.conv.0.weight */
  const float* c1d_b1; /* f [d] */
  const void* c1d_w2;  /* h [d, k * d]        encoder.pre_encode.conv.2.weight */
  const float* c1d_b2; /* f [d] */
  /* tensor-core front end: split-precision DFT basis h [512, 3*Kp], Kp = n_fft rounded up to 64; every 256-row tile is
   * [128 cos rows | 128 sin rows] of bins tile*128.., K blocks [hi | hi | lo]; and the bin range of every mel filter */
  const void* dft_w;
  const int32_t* mel_lo; /* i32 [n_mels] first bin with a non-zero weight */
  const int32_t* mel_hi; /* i32 [n_mels] one past the last */
} gam_weights;

int gam_create(const gam_config* cfg, const gam_weights* w, int device, gam_handle** out);
void gam_destroy(gam_handle* h);
const char* gam_last_error(const gam_handle* h);
int gam_version(void);

/* frames of log-mel for n_samples (gigaam/preprocess.py:78-92) and encoder frames for M mel frames
 * (gigaam/encoder.py:77-90) -- host arithmetic */
int64_t gam_logmel_frames(const gam_handle* h, int64_t n_samples);
int64_t gam_encoded_frames(const gam_handle* h, int64_t mel_frames);

/* bytes of scratch gam_encode / gam_*_greedy need for a batch of B utterances of M mel frames */
int64_t gam_workspace_bytes(const gam_handle* h, int32_t B, int64_t mel_frames);

/* bytes of scratch gam_ctc_greedy / gam_rnnt_greedy need on their own (B utterances of T encoder frames) */
int64_t gam_decode_workspace_bytes(const gam_handle* h, int32_t B, int32_t T);

/* wav: device f32 [B, n_samples]  ->  mel: device f32 [B, n_mels, M] */
int gam_logmel(gam_handle* h, const float* wav, int32_t B, int64_t n_samples, float* mel, void* stream);

/* Same result as gam_logmel on the tensor cores: frames -> fp16 (hi, lo) split -> one K-concatenated tcgen05 GEMM
 * against the split DFT basis with a |X|^2 epilogue -> sparse mel projection + log.  Needs scratch
 * (gam_logmel_workspace_bytes); gam_logmel (one fused CUDA-core kernel) needs none. */
int64_t gam_logmel_workspace_bytes(const gam_handle* h, int32_t B, int64_t n_samples);
int gam_logmel_tc(gam_handle* h, const float* wav, int32_t B, int64_t n_samples, float* mel, void* workspace,
                  int64_t workspace_bytes, void* stream);

/* Varlen execution: lengths stay on the device (no host synchronisation, same launch sequence for every mix of lengths, so a
 * captured CUDA graph of the call is valid for all of them).  From the second subsampling stage on only the frames that
 * exist are computed (rows of the utterances packed back to back, cu_seqlens built by the first kernel of the call -- the
 * contract of apply_masked_flash_attn, gigaam/utils.py:103-155, applied to the whole Conformer block); frames t >= enc_len[b]
 * of `enc` are written as zeros.  A batch of ONE keeps its padded frames, like the reference (no attention mask for B == 1).
 * mel: device f32 [B, feat_in, M]; mel_len: device i64 [B]
 * -> enc: device f32 [B, T', d_model] (row-major; the reference's [B, d, T'] is its transpose(1,2) view)
 *    enc_len: device i32 [B].  n_layers_run < 0 runs the full stack; 0..n_layers stops early (tests). */
int gam_encode(gam_handle* h, const float* mel, const int64_t* mel_len, int32_t B, int64_t M, void* workspace,
               int64_t workspace_bytes, float* enc, int32_t* enc_len, int32_t n_layers_run, void* stream);

/* enc: device f32 [B, T, d_model]; enc_len: device i32 [B]
 * -> ids / frames: device i32 [B, max_out], counts: device i32 [B] */
int gam_ctc_greedy(gam_handle* h, const float* enc, const int32_t* enc_len, int32_t B, int32_t T, void* workspace,
                   int64_t workspace_bytes, int32_t* ids, int32_t* frames, int32_t* counts, int32_t max_out,
                   void* stream);
int gam_rnnt_greedy(gam_handle* h, const float* enc, const int32_t* enc_len, int32_t B, int32_t T, void* workspace,
                    int64_t workspace_bytes, int32_t* ids, int32_t* frames, int32_t* counts, int32_t max_out,
                    void* stream);

/* ---- the one multi-GPU exchange of the path: utterances are sharded over ranks, one process per GPU, and the
 * device-resident hypotheses are all-gathered ONCE over NCCL (NVLink / NVSwitch) when the batch was actually split.
 * gam_comm_unique_id: rank 0 fills 128 bytes, the host ships them to every rank by any channel (torch.distributed, MPI,
 * a file); gam_comm_init: collective over all ranks, binds an NCCL communicator to the handle; gam_gather_hyps: every
 * rank passes ONE packed int32 device buffer [ids B_local x W | frames B_local x W | counts B_local] of n_int32 elements
 * (same n on every rank: pad short shards with counts = 0) and receives gathered[r * n_int32 ...] = rank r's buffer.
 * Stream-ordered, capturable in a CUDA graph, no host synchronisation.  NCCL is bound at run time (dlopen): hosts without
 * it keep every other entry point; these three then return an error. */
int gam_comm_unique_id(uint8_t* out128);
int gam_comm_init(gam_handle* h, const uint8_t* id128, int32_t rank, int32_t nranks);
int32_t gam_comm_nccl_version(void);
int gam_gather_hyps(gam_handle* h, const int32_t* packed, int64_t n_int32, int32_t* gathered, void* stream);

/* Word grouping of hypotheses on the device  <- gigaam/timestamps_utils.py:13-53 frames_to_words (gigaam/model.py:104-124).
 * ids / frames / counts as produced by gam_*_greedy (row pitch max_out); token_flags: device u8 [V], bit 0 = the piece is
 * " " (delimiter), bit 1 = the piece starts with U+2581 (opens a new word), bit 2 = the piece (prefix removed) is empty
 * after strip().  Per utterance b and word w < n_words[b] (row pitch max_words; max_words >= max_out is always enough):
 * word_start = frame of the first piece, word_end = frame of the last piece + 1, and the pieces are tokens
 * [word_first_token, word_first_token + word_tokens) of ids[b].  Times = frame * frame_shift on the host. */
int gam_group_words(gam_handle* h, const int32_t* ids, const int32_t* frames, const int32_t* counts, int32_t B, int32_t max_out,
                    const uint8_t* token_flags, int32_t V, int32_t max_words, int32_t* word_start, int32_t* word_end,
                    int32_t* word_first_token, int32_t* word_tokens, int32_t* n_words, void* stream);

/* ---- unit entry points (parity tests of the individual kernels) ---- */
/* D[M,N] = A[M,K] W[N,K]^T with epilogue `kind` (0 bias->f16, 1 bias+silu->f16, 2 bias+glu->f16 [N/2 cols],
 * 3 res + scale*(acc+bias) -> f32, 4 bias -> f32).  A, W fp16 device; N % 256 == 0; K % 64 == 0. */
int gam_test_gemm(gam_handle* h, int32_t kind, const void* A, const void* W, const float* bias, const float* res, void* out,
                  int32_t M, int32_t N, int32_t K, int32_t ldo, float scale, void* stream);
/* qkv: f16 [B*T, 3*d_model]; klen i32 [B] or NULL -> out f16 [B*T, d_model] */
int gam_test_attention(gam_handle* h, const void* qkv, const int32_t* klen, void* out, int32_t B, int32_t T, void* stream);
/* rel_pos variant: qkv f16 [B*T, 4*d_model] = [q+u | q+v | k | v]; pos f16 [2*GAM_REL_POS_MAX_T-1, d_model] */
int gam_test_attention_relpos(gam_handle* h, const void* qkv, const void* pos, const int32_t* klen, void* out, int32_t B,
                              int32_t T, void* stream);
/* packed-row (varlen) form of the two calls above, the one gam_encode uses: qkv f16 [rows, 3*d_model] (pos == NULL, rotary)
 * or [rows, 4*d_model] (pos != NULL, rel_pos); utterance b owns rows cu[b] .. cu[b] + klen[b] (cu: i32 [B + 1], klen: i32
 * [B], klen[b] <= T) -> out f16 [rows, d_model]; rows that belong to no utterance are left untouched.  This is the
 * cu_seqlens contract of flash_attn_varlen_func in gigaam/utils.py:103-155 (apply_masked_flash_attn). */
int gam_test_attention_varlen(gam_handle* h, const void* qkv, const void* pos, const int32_t* klen, const int32_t* cu, void* out,
                              int32_t B, int32_t T, int32_t rows, void* stream);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t gam_launch_count(const gam_handle* h);

/* Optional per-launch timing with CUDA events on the launching stream (bench.py's roofline leg).
 * gam_profile_begin() arms it; every kernel launched through this handle until gam_profile_end() is bracketed
 * by an event pair; gam_profile_end() synchronises the events and returns summed milliseconds and launch
 * counts per kernel class (gam_profile_class_name()).  Must not be armed during CUDA-graph capture. */
int gam_profile_begin(gam_handle* h);
int gam_profile_end(gam_handle* h, double* ms_per_class, int64_t* launches_per_class, int32_t n_classes);
int gam_profile_class_count(void);
const char* gam_profile_class_name(int32_t cls);

#ifdef __cplusplus
}
#endif
#endif /* GIGAAM_B200_H_ */
