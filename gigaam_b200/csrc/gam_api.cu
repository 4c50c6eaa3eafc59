// C-ABI + host engine of libgigaam_b200.so: weight/plan bookkeeping, TMA descriptor construction,
// and the kernel sequence of the path.  No compute lives here and nothing here falls back to a CPU
// or library implementation: every stage is one of the hand-written kernels in this directory.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gigaam_b200.h"
#include "comm.h"
#include "gemm_params.cuh"
#include "kernels.h"
#include "launch.cuh"

namespace gam {

// ------------------------------------------------------------------ driver entry point for TMA descriptors
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int init_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || fn == nullptr ||
      q != cudaDriverEntryPointSuccess)
    return -1;
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return 0;
}

// 2-D fp16 row-major tensor [rows, cols] with row pitch ld_elems; box = [box_rows, box_cols], SWIZZLE_128B
int make_tmap_2d_f16(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems, uint32_t box_rows,
                     uint32_t box_cols) {
  if (init_encode() != 0) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

// 4-D channels-last activation [B, T1, F1, C] fp16 for the stride-2 3x3 conv: box = 8 time x 16 freq x 64 ch,
// traversal stride 2 on time and freq (so boxDim is 16 / 32 elements in tensor coordinates).
static int make_tmap_conv4d(CUtensorMap* m, const void* base, uint64_t B, uint64_t T1, uint64_t F1, uint64_t C) {
  if (init_encode() != 0) return -1;
  cuuint64_t dims[4] = {C, F1, T1, B};
  cuuint64_t strides[3] = {C * 2, F1 * C * 2, T1 * F1 * C * 2};
  cuuint32_t box[4] = {64, 32, 16, 1};
  cuuint32_t estr[4] = {1, 2, 2, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

// 3-D time-major activation [B, T, C] fp16 for the stride-2 conv1d: box = 128 output frames (256 input frames
// traversed with stride 2) x 64 channels
static int make_tmap_conv3d(CUtensorMap* m, const void* base, uint64_t B, uint64_t T, uint64_t C) {
  if (init_encode() != 0) return -1;
  cuuint64_t dims[3] = {C, T, B};
  cuuint64_t strides[2] = {C * 2, T * C * 2};
  cuuint32_t box[3] = {64, 256, 1};
  cuuint32_t estr[3] = {1, 2, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct LayerMaps {
  CUtensorMap ff1_w1, ff1_w2, w_qk, w_v, w_o, pw1, pw2, ff2_w1, ff2_w2;
  CUtensorMap w_qkv_rel, pos_proj;   // rel_pos attention only
  CUtensorMap w_qkv;                 // rotary: [W_q ; W_k ; W_v] when the caller packed them contiguously
  bool qkv_merged = false;
};

struct Plan {
  int B = 0;
  int64_t M = 0;
  void* ws = nullptr;
  // geometry
  int T1 = 0, F1 = 0, T2 = 0, F2 = 0, R = 0;
  // workspace carve-up
  int *len0 = nullptr, *len1 = nullptr, *len2 = nullptr;
  // packed-row plan (pack_plan_kernel): frames kept per utterance, their prefix sum [B + 1], the live row count, the
  // stage-1 frames that are produced, and row -> (utterance, frame)
  int *plen = nullptr, *cu = nullptr, *rows_dev = nullptr, *run1 = nullptr, *row_b = nullptr, *row_t = nullptr;
  __half *s1 = nullptr, *s2 = nullptr, *a16 = nullptr, *r16 = nullptr, *big16 = nullptr, *o16 = nullptr, *g16 = nullptr;
  __half* melT = nullptr;   // conv1d subsampling: time-major fp16 copy of the log-mel
  CUtensorMap m_melT, m_s1_3d;
  float* x = nullptr;
  int64_t bytes = 0;
  CUtensorMap m_s1, m_s2, m_a16, m_r16, m_hid, m_qkv, m_qkv4, m_o16;
};

}  // namespace gam

using namespace gam;

struct gam_handle {
  gam_config cfg;
  gam_weights w;
  std::vector<gam_layer_weights> layers;
  std::vector<LayerMaps> lmaps;
  CUtensorMap m_sub2_w, m_sub_out_w;
  CUtensorMap m_dft_w, m_lm_a;   // tensor-core front end: split DFT basis / frame matrix (cached per workspace)
  const void* lm_A = nullptr;
  int64_t lm_F = 0;
  int device = 0;
  int num_sms = 148;
  int64_t launches = 0;
  void* comm = nullptr;      // ncclComm_t of gam_comm_init
  int comm_rank = 0, comm_nranks = 1;
  std::string err;
  std::vector<Plan*> plans;
  // optional per-launch CUDA-event timing (bench.py's roofline leg); never active during graph capture
  bool prof = false;
  std::vector<cudaEvent_t> prof_ev;   // start/stop pairs
  std::vector<int> prof_cls;
};

// kernel classes reported by gam_profile_end
enum ProfClass : int {
  PC_LOGMEL = 0, PC_SUB_CONV1, PC_GEMM_CONV2, PC_GEMM_SUBOUT, PC_GEMM_FFN_UP, PC_GEMM_FFN_DOWN, PC_GEMM_QKV, PC_GEMM_PROJ,
  PC_GEMM_GLU, PC_LAYERNORM, PC_ATTENTION, PC_DWCONV, PC_CTC_ARGMAX, PC_CTC_COLLAPSE, PC_RNNT_ENCPROJ, PC_RNNT_GREEDY,
  PC_MISC, PC_COUNT
};

struct ProfScope {
  gam_handle* h;
  cudaStream_t s;
  ProfScope(gam_handle* h_, int cls, cudaStream_t s_) : h(h_), s(s_) {
    h->launches += 1;
    if (!h->prof) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    h->prof_ev.push_back(a);
    h->prof_ev.push_back(b);
    h->prof_cls.push_back(cls);
    cudaEventRecord(a, s);
  }
  ~ProfScope() {
    if (h->prof) cudaEventRecord(h->prof_ev.back(), s);
  }
};
#define PROF(cls) ProfScope prof_scope__(h, cls, s)

namespace {

int fail(gam_handle* h, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (h) h->err = buf;
  return code;
}

int sub_out_len(int len, int k, int pad) {
  // floor((len + 2p - k) / 2 + 1) with float semantics of the reference (encoder.py:86-90)
  float l = static_cast<float>(len);
  l = floorf((l + (2 * pad - k)) / 2.0f + 1.0f);
  return static_cast<int>(l);
}

void plan_geometry(const gam_handle* h, int B, int64_t M, Plan* p) {
  const gam_config& c = h->cfg;
  const int k = c.subs_kernel_size, pad = (k - 1) / 2;
  p->B = B;
  p->M = M;
  p->T1 = sub_out_len(static_cast<int>(M), k, pad);
  p->T2 = sub_out_len(p->T1, k, pad);
  p->F1 = c.subsampling == 0 ? sub_out_len(c.feat_in, k, pad) : 1;
  p->F2 = c.subsampling == 0 ? sub_out_len(p->F1, k, pad) : 1;
  p->R = B * p->T2;
}

// carve the workspace; returns total bytes.  base may be null (size query).
int64_t plan_carve(const gam_handle* h, Plan* p, uint8_t* base) {
  const gam_config& c = h->cfg;
  const int64_t d = c.d_model, R = p->R, B = p->B;
  const int64_t nqkv = (c.self_attention == 1 ? 4 : 3) * d;   // rel_pos carries q twice (q+u, q+v)
  const int64_t wide = (c.d_ff > nqkv ? c.d_ff : nqkv);
  int64_t off = 0;
  auto take = [&](int64_t bytes) -> uint8_t* {
    uint8_t* ptr = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return ptr;
  };
  p->len0 = reinterpret_cast<int*>(take(B * 4));
  p->len1 = reinterpret_cast<int*>(take(B * 4));
  p->len2 = reinterpret_cast<int*>(take(B * 4));
  p->plen = reinterpret_cast<int*>(take(B * 4));
  p->run1 = reinterpret_cast<int*>(take(B * 4));
  p->cu = reinterpret_cast<int*>(take((B + 1) * 4));
  p->rows_dev = reinterpret_cast<int*>(take(4));
  p->row_b = reinterpret_cast<int*>(take(R * 4));
  p->row_t = reinterpret_cast<int*>(take(R * 4));
  p->x = reinterpret_cast<float*>(take(R * d * 4));
  p->a16 = reinterpret_cast<__half*>(take(R * d * 2));
  p->r16 = reinterpret_cast<__half*>(take(R * d * 2));
  p->big16 = reinterpret_cast<__half*>(take(R * wide * 2));
  p->o16 = reinterpret_cast<__half*>(take(R * d * 2));
  p->g16 = reinterpret_cast<__half*>(take(R * d * 2));
  p->s2 = reinterpret_cast<__half*>(take(R * static_cast<int64_t>(p->F2) * d * 2));
  p->s1 = reinterpret_cast<__half*>(take(B * static_cast<int64_t>(p->T1) * p->F1 * d * 2));
  p->melT = reinterpret_cast<__half*>(take(c.subsampling == 1 ? B * p->M * static_cast<int64_t>(c.feat_in) * 2 : 0));
  return off;
}

int64_t decode_ws_bytes(const gam_handle* h, int B, int T) {
  // CTC: labels [B*T] i32 ; RNNT: encproj [B*T, joint_hidden] f32
  const int64_t R = static_cast<int64_t>(B) * T;
  int64_t a = align_up(R * 4, 1024);
  int64_t b = align_up(R * (h->cfg.joint_hidden > 0 ? h->cfg.joint_hidden : 1) * 4, 1024);
  return a + b;
}

Plan* get_plan(gam_handle* h, int B, int64_t M, void* ws, int64_t ws_bytes) {
  for (Plan* p : h->plans)
    if (p->B == B && p->M == M && p->ws == ws) return p;
  Plan* p = new Plan();
  plan_geometry(h, B, M, p);
  p->ws = ws;
  p->bytes = plan_carve(h, p, static_cast<uint8_t*>(ws));
  if (p->bytes > ws_bytes) {
    fail(h, -1, "workspace too small: need %lld bytes, got %lld", (long long)p->bytes, (long long)ws_bytes);
    delete p;
    return nullptr;
  }
  const gam_config& c = h->cfg;
  const uint64_t d = c.d_model, R = p->R;
  int rc = 0;
  if (c.subsampling == 0) {
    rc |= make_tmap_conv4d(&p->m_s1, p->s1, B, p->T1, p->F1, d);
    rc |= make_tmap_2d_f16(&p->m_s2, p->s2, R, static_cast<uint64_t>(p->F2) * d, static_cast<uint64_t>(p->F2) * d, 128, 64);
  } else {
    rc |= make_tmap_conv3d(&p->m_melT, p->melT, B, static_cast<uint64_t>(p->M), c.feat_in);
    rc |= make_tmap_conv3d(&p->m_s1_3d, p->s1, B, p->T1, d);
  }
  rc |= make_tmap_2d_f16(&p->m_a16, p->a16, R, d, d, 128, 64);
  rc |= make_tmap_2d_f16(&p->m_r16, p->r16, R, d, d, 128, 64);
  rc |= make_tmap_2d_f16(&p->m_hid, p->big16, R, c.d_ff, c.d_ff, 128, 64);
  rc |= make_tmap_2d_f16(&p->m_qkv, p->big16, R, 3 * d, 3 * d, 128, 64);
  rc |= make_tmap_2d_f16(&p->m_qkv4, p->big16, R, 4 * d, 4 * d, 128, 64);
  rc |= make_tmap_2d_f16(&p->m_o16, p->o16, R, d, d, 128, 64);
  if (rc != 0) {
    fail(h, -2, "cuTensorMapEncodeTiled failed for activation maps (rc=%d)", rc);
    delete p;
    return nullptr;
  }
  if (h->plans.size() >= 16) {
    delete h->plans.front();
    h->plans.erase(h->plans.begin());
  }
  h->plans.push_back(p);
  return p;
}

#define GAM_CHECK_LAUNCH(h, what)                                                             \
  do {                                                                                        \
    cudaError_t e__ = cudaGetLastError(); /* reads AND clears: one failed launch must not poison later calls */ \
    if (e__ != cudaSuccess) return fail(h, -3, "%s: %s", what, cudaGetErrorString(e__));      \
  } while (0)

}  // namespace

extern "C" {

int gam_version(void) { return 100; }

const char* gam_last_error(const gam_handle* h) { return h ? h->err.c_str() : "null handle"; }

int64_t gam_launch_count(const gam_handle* h) { return h ? h->launches : 0; }

int gam_create(const gam_config* cfg, const gam_weights* w, int device, gam_handle** out) {
  if (!cfg || !w || !out) return -1;
  *out = nullptr;
  gam_handle* h = new gam_handle();
  h->cfg = *cfg;
  h->w = *w;
  h->device = device;
  *out = h;  // returned even on failure so the caller can read gam_last_error()
  const gam_config& c = h->cfg;
  if (c.subsampling != 0 && c.subsampling != 1) return fail(h, -10, "unknown subsampling type %d", c.subsampling);
  if (c.subsampling == 1 && (c.feat_in % 64 != 0 || (c.subs_kernel_size & 1) == 0))
    return fail(h, -10, "conv1d subsampling needs feat_in %% 64 == 0 and an odd kernel size");
  if (c.self_attention != 0 && c.self_attention != 1) return fail(h, -10, "unknown self_attention type %d", c.self_attention);
  if (c.d_model != 768 || c.d_model % c.n_heads != 0 || (c.d_model / c.n_heads) % 16 != 0)
    return fail(h, -10, "unsupported d_model/n_heads (%d/%d): kernels are specialised for d_model 768, d_k %% 16 == 0",
                c.d_model, c.n_heads);
  if (c.d_ff % 256 != 0 || (c.subsampling == 0 && c.subs_kernel_size != 3)) return fail(h, -10, "unsupported d_ff / subs_kernel_size");
  if (c.win_length != c.n_fft) return fail(h, -10, "win_length != n_fft is not supported");
  if (cudaSetDevice(device) != cudaSuccess) return fail(h, -11, "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(h, -11, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return fail(h, -12, "sm_100a kernels need a Blackwell (cc 10.x) device, found cc %d.%d", prop.major, prop.minor);
  h->num_sms = prop.multiProcessorCount;
  if (init_encode() != 0) return fail(h, -13, "cuTensorMapEncodeTiled entry point not available");
  if (gemm_init() != 0) return fail(h, -14, "cudaFuncSetAttribute failed for the GEMM kernels: %s", cudaGetErrorString(cudaGetLastError()));
  h->layers.assign(w->layers, w->layers + c.n_layers);
  h->w.layers = h->layers.data();
  h->lmaps.resize(c.n_layers);
  const uint64_t d = c.d_model, ff = c.d_ff;
  int rc = 0;
  for (int l = 0; l < c.n_layers; ++l) {
    const gam_layer_weights& lw = h->layers[l];
    LayerMaps& lm = h->lmaps[l];
    rc |= make_tmap_2d_f16(&lm.ff1_w1, lw.ff1_w1, ff, d, d, 128, 64);
    rc |= make_tmap_2d_f16(&lm.ff1_w2, lw.ff1_w2, d, ff, ff, 128, 64);
    if (c.self_attention == 0) {
      rc |= make_tmap_2d_f16(&lm.w_qk, lw.w_qk, 2 * d, d, d, 128, 64);
      rc |= make_tmap_2d_f16(&lm.w_v, lw.w_v, d, d, d, 128, 64);
      // W_v directly behind W_qk (and b_v behind b_qk): q, k and v projections run as one launch (launch_gemm_dual_a)
      lm.qkv_merged = static_cast<const char*>(lw.w_v) == static_cast<const char*>(lw.w_qk) + 2 * d * d * 2 && lw.b_v == lw.b_qk + 2 * d;
      if (lm.qkv_merged) rc |= make_tmap_2d_f16(&lm.w_qkv, lw.w_qk, 3 * d, d, d, 128, 64);
    } else {
      if (!lw.w_qkv_rel || !lw.b_qkv_rel || !lw.pos_proj) return fail(h, -10, "layer %d: rel_pos weights missing", l);
      rc |= make_tmap_2d_f16(&lm.w_qkv_rel, lw.w_qkv_rel, 4 * d, d, d, 128, 64);
      rc |= make_tmap_2d_f16(&lm.pos_proj, lw.pos_proj, 2 * kRelPosMaxT - 1, d, d, 128, 64);
    }
    rc |= make_tmap_2d_f16(&lm.w_o, lw.w_o, d, d, d, 128, 64);
    rc |= make_tmap_2d_f16(&lm.pw1, lw.pw1_w, 2 * d, d, d, 128, 64);
    rc |= make_tmap_2d_f16(&lm.pw2, lw.pw2_w, d, d, d, 128, 64);
    rc |= make_tmap_2d_f16(&lm.ff2_w1, lw.ff2_w1, ff, d, d, 128, 64);
    rc |= make_tmap_2d_f16(&lm.ff2_w2, lw.ff2_w2, d, ff, ff, 128, 64);
  }
  const int pad = (c.subs_kernel_size - 1) / 2;
  const int F1 = sub_out_len(c.feat_in, c.subs_kernel_size, pad), F2 = sub_out_len(F1, c.subs_kernel_size, pad);
  if (c.subsampling == 0) {
    if (F1 != 32 || F2 != 16) return fail(h, -10, "conv2d subsampling kernels are specialised for feat_in 64 (F1=32,F2=16)");
    rc |= make_tmap_2d_f16(&h->m_sub2_w, w->sub2_w, d, 9 * d, 9 * d, 128, 64);
    rc |= make_tmap_2d_f16(&h->m_sub_out_w, w->sub_out_w, d, static_cast<uint64_t>(F2) * d, static_cast<uint64_t>(F2) * d, 128, 64);
  } else {
    const uint64_t k1 = static_cast<uint64_t>(c.subs_kernel_size) * c.feat_in, k2 = static_cast<uint64_t>(c.subs_kernel_size) * d;
    rc |= make_tmap_2d_f16(&h->m_sub2_w, w->c1d_w1, d, k1, k1, 128, 64);        // stage 1: [d, taps * feat_in]
    rc |= make_tmap_2d_f16(&h->m_sub_out_w, w->c1d_w2, d, k2, k2, 128, 64);     // stage 2: [d, taps * d]
  }
  if (w->dft_w != nullptr) {
    const uint64_t kk = 3 * static_cast<uint64_t>((c.n_fft + 63) / 64 * 64);
    rc |= make_tmap_2d_f16(&h->m_dft_w, w->dft_w, 512, kk, kk, 128, 64);
  }
  if (rc != 0) return fail(h, -2, "cuTensorMapEncodeTiled failed for weight maps (rc=%d)", rc);
  return 0;
}

void gam_destroy(gam_handle* h) {
  if (!h) return;
  for (Plan* p : h->plans) delete p;
  comm_destroy(h->comm);
  delete h;
}

int64_t gam_logmel_frames(const gam_handle* h, int64_t n) {
  const gam_config& c = h->cfg;
  if (c.center) return n / c.hop_length + 1;
  return n < c.win_length ? 0 : (n - c.win_length) / c.hop_length + 1;
}

int64_t gam_encoded_frames(const gam_handle* h, int64_t M) {
  Plan p;
  plan_geometry(h, 1, M, &p);
  return p.T2;
}

int64_t gam_decode_workspace_bytes(const gam_handle* h, int32_t B, int32_t T) { return decode_ws_bytes(h, B, T) + 1024; }

int64_t gam_workspace_bytes(const gam_handle* h, int32_t B, int64_t M) {
  Plan p;
  plan_geometry(h, B, M, &p);
  return plan_carve(h, &p, nullptr) + decode_ws_bytes(h, B, p.T2) + 4096;
}

int gam_logmel(gam_handle* h, const float* wav, int32_t B, int64_t n_samples, float* mel, void* stream) {
  const gam_config& c = h->cfg;
  const int64_t M = gam_logmel_frames(h, n_samples);
  if (M <= 0) return fail(h, -1, "waveform too short: %lld samples", (long long)n_samples);
  if (c.center && n_samples <= c.n_fft / 2) return fail(h, -1, "reflect padding needs more than n_fft/2 samples");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    PROF(PC_LOGMEL);
    if (launch_logmel(wav, B, static_cast<int>(n_samples), static_cast<int>(M), h->w.window, h->w.dft_cos, h->w.dft_sin,
                      h->w.mel_fb, mel, c.n_fft, c.hop_length, c.center, c.n_mels, s) != 0)
      return fail(h, -1, "logmel: unsupported n_fft/n_mels (%d/%d)", c.n_fft, c.n_mels);
  }
  GAM_CHECK_LAUNCH(h, "logmel");
  return 0;
}

static inline int logmel_kp(const gam_config& c) { return (c.n_fft + 63) / 64 * 64; }

int64_t gam_logmel_workspace_bytes(const gam_handle* h, int32_t B, int64_t n_samples) {
  const int64_t F = static_cast<int64_t>(B) * gam_logmel_frames(h, n_samples);
  return align_up(F * 3 * logmel_kp(h->cfg) * 2, 1024) + align_up(F * 256 * 4, 1024) + 2048;
}

int gam_logmel_tc(gam_handle* h, const float* wav, int32_t B, int64_t n_samples, float* mel, void* workspace,
                  int64_t workspace_bytes, void* stream) {
  const gam_config& c = h->cfg;
  if (!h->w.dft_w || !h->w.mel_lo || !h->w.mel_hi) return fail(h, -1, "logmel_tc: split DFT basis not provided");
  if (c.n_fft / 2 + 1 > 256) return fail(h, -1, "logmel_tc: n_fft %d exceeds 256 bins", c.n_fft);
  const int64_t M = gam_logmel_frames(h, n_samples);
  if (M <= 0) return fail(h, -1, "waveform too short: %lld samples", (long long)n_samples);
  if (c.center && n_samples <= c.n_fft / 2) return fail(h, -1, "reflect padding needs more than n_fft/2 samples");
  if (workspace_bytes < gam_logmel_workspace_bytes(h, B, n_samples)) return fail(h, -1, "logmel_tc: workspace too small");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int Kp = logmel_kp(c);
  const int64_t F = static_cast<int64_t>(B) * M;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  __half* A = reinterpret_cast<__half*>(ws);
  float* P = reinterpret_cast<float*>(ws + align_up(F * 3 * Kp * 2, 1024));
  if (h->lm_A != A || h->lm_F != F) {
    if (make_tmap_2d_f16(&h->m_lm_a, A, F, 3 * Kp, 3 * Kp, 128, 64) != 0) return fail(h, -2, "logmel_tc: tensor map encode failed");
    h->lm_A = A;
    h->lm_F = F;
  }
  {
    PROF(PC_LOGMEL);
    launch_frames_split(wav, B, static_cast<int>(n_samples), static_cast<int>(M), h->w.window, A, c.n_fft, Kp, c.hop_length, c.center, s);
  }
  {
    PROF(PC_LOGMEL);
    if (launch_gemm_power(&h->m_lm_a, &h->m_dft_w, static_cast<int>(F), 512, 3 * Kp, P, 256, h->num_sms, s) != 0)
      return fail(h, -4, "logmel_tc: DFT GEMM launch rejected");
  }
  {
    PROF(PC_LOGMEL);
    launch_mel_log(P, 256, B, static_cast<int>(M), c.n_fft / 2 + 1, h->w.mel_fb, h->w.mel_lo, h->w.mel_hi, mel, c.n_mels, s);
  }
  GAM_CHECK_LAUNCH(h, "logmel_tc");
  return 0;
}

int gam_encode(gam_handle* h, const float* mel, const int64_t* mel_len, int32_t B, int64_t M, void* workspace,
               int64_t workspace_bytes, float* enc, int32_t* enc_len, int32_t n_layers_run, void* stream) {
  const gam_config& c = h->cfg;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (B <= 0 || M <= 0) return fail(h, -1, "encode: empty batch (B=%d, M=%lld); callers skip the call for an empty shard", B, (long long)M);
  Plan* p = get_plan(h, B, M, workspace, workspace_bytes);
  if (!p) return -1;
  if (p->T2 <= 0) return fail(h, -1, "input too short for the subsampling (M=%lld)", (long long)M);
  if (p->T2 > GAM_REL_POS_MAX_T)
    return fail(h, -1, "T'=%d exceeds the attention kernels' %d-frame limit (%.1f s of audio); cut the recording into segments "
                "(transcribe_longform does)", p->T2, GAM_REL_POS_MAX_T, GAM_REL_POS_MAX_T * 0.04);
  const int d = c.d_model, R = p->R, nsm = h->num_sms;
  const int L = (n_layers_run < 0 || n_layers_run > c.n_layers) ? c.n_layers : n_layers_run;
  int rc = 0;

  // Varlen execution: from the second subsampling stage on, the path keeps only the frames that exist.  Utterance b owns rows
  // cu[b] .. cu[b] + plen[b] of every activation matrix (plen = its subsampled length; a batch of one keeps all T' frames,
  // like the reference's mask-free single-utterance attention); the live row count stays on the device (rows_dev), so the
  // same launch sequence -- and a captured CUDA graph of it -- serves every mix of lengths.  Grids are sized for R = B*T'.
  {
    PROF(PC_MISC);
    launch_pack_plan(reinterpret_cast<const long long*>(mel_len), B, 2 * ((c.subs_kernel_size - 1) / 2) - c.subs_kernel_size,
                     static_cast<int>(M), p->T1, p->T2, p->len0, p->len1, p->len2, p->plen, p->run1, p->cu, p->rows_dev, p->row_b,
                     p->row_t, s);
  }
  const int* rdev = p->rows_dev;
  if (c.subsampling == 0) {
    {
      PROF(PC_SUB_CONV1);
      rc |= launch_subsample_conv1(mel, p->len0, p->len1, p->run1, h->w.sub1_w, h->w.sub1_b, p->s1, B, static_cast<int>(M), c.feat_in,
                                   p->T1, p->F1, d, s);
    }
    {
      PROF(PC_GEMM_CONV2);
      rc |= launch_gemm_conv(&p->m_s1, &h->m_sub2_w, B, p->T2, d, d, h->w.sub2_b, p->len2, p->cu, p->plen, p->s2, d, nsm, s);
    }
    GAM_CHECK_LAUNCH(h, "subsampling");
    if (rc) return fail(h, -4, "subsampling launch rejected (rc=%d)", rc);
    {
      PROF(PC_GEMM_SUBOUT);
      rc |= launch_gemm(GEMM_BIAS_F32, &p->m_s2, &h->m_sub_out_w, R, d, p->F2 * d, h->w.sub_out_b, nullptr, p->x, d, 1.f, nsm, s, 0, rdev);
    }
  } else {
    // conv1d subsampling (gigaam/encoder.py:59-70 with Conv1d): two k-tap / stride-2 implicit GEMMs over time-major data
    {
      PROF(PC_SUB_CONV1);
      launch_mel_to_tmajor_f16(mel, p->len0, p->melT, B, c.feat_in, static_cast<int>(M), s);
    }
    {
      PROF(PC_GEMM_CONV2);
      rc |= launch_gemm_conv1d(&p->m_melT, &h->m_sub2_w, B, p->T1, c.feat_in, c.subs_kernel_size, d, h->w.c1d_b1, p->len1, nullptr,
                               nullptr, p->s1, d, 0, nsm, s);   // stage 1 stays [B, T1]: stage 2 fetches it with 3-D TMA boxes
    }
    {
      PROF(PC_GEMM_SUBOUT);
      rc |= launch_gemm_conv1d(&p->m_s1_3d, &h->m_sub_out_w, B, p->T2, d, c.subs_kernel_size, d, h->w.c1d_b2, p->len2, p->cu,
                               p->plen, p->x, d, 1, nsm, s);
    }
    GAM_CHECK_LAUNCH(h, "subsampling");
    if (rc) return fail(h, -4, "conv1d subsampling launch rejected (rc=%d)", rc);
  }
  // Direction plan of a layer (GemmParams::reverse): every kernel walks the rows in the direction OPPOSITE to the one its
  // main input was written in, so that it starts on what is still in L2.  U = ascending, D = descending:
  //   LN_ff1 D | FF1-up U | FF1-down D | LN_att U | QKV D | attention U | proj D | LN_conv U | pw1 D | depthwise U | pw2 D |
  //   LN_ff2 U | FF2-up D | FF2-down U | LN_out (+ next LN_ff1) D
  // Measured on c2 (same box, alternating runs, profiles/r2f_zigzag_ab.md): 12.47 ms per step one-directional, 12.27 ms alternating.
  constexpr int zz = 1;
  if (L == 0) {   // n_layers_run == 0: the caller wants the pre_encode output
    PROF(PC_MISC);
    launch_unpack_rows(p->x, nullptr, nullptr, p->cu, p->plen, enc, B, p->T2, 0, s);
  } else {
    PROF(PC_LAYERNORM);
    launch_ln_f16(p->x, h->layers[0].ln_ff1_g, h->layers[0].ln_ff1_b, p->a16, R, rdev, zz, s);
  }
  const int dk = d / c.n_heads;
  for (int l = 0; l < L; ++l) {
    const gam_layer_weights& w = h->layers[l];
    const LayerMaps& m = h->lmaps[l];
    // x += 0.5 * FF1(LN(x))                                     (encoder.py:480-483)
    { PROF(PC_GEMM_FFN_UP);
      rc |= launch_gemm(GEMM_BIAS_SILU_F16, &p->m_a16, &m.ff1_w1, R, c.d_ff, d, w.ff1_b1, nullptr, p->big16, c.d_ff, 1.f, nsm, s, 0, rdev); }
    { PROF(PC_GEMM_FFN_DOWN);
      rc |= launch_gemm(GEMM_BIAS_RES_F32, &p->m_hid, &m.ff1_w2, R, d, c.d_ff, w.ff1_b2, p->x, p->x, d, 0.5f, nsm, s, zz, rdev); }
    // x += W_o attn(q = W_q rope(u), k = W_k rope(u), v = W_v u), u = LN(x)   (encoder.py:485-487, 236-277)
    if (c.self_attention == 0) {
      { PROF(PC_LAYERNORM);
        launch_ln_rope_f16(p->x, w.ln_att_g, w.ln_att_b, h->w.rope_cos, h->w.rope_sin, p->a16, p->r16, R, rdev, p->row_t, p->T2,
                           dk / 2, 0, s); }
      bool merged = false;
      if (m.qkv_merged) {
        PROF(PC_GEMM_QKV);
        merged = launch_gemm_dual_a(&p->m_r16, &p->m_a16, 2 * d, &m.w_qkv, R, 3 * d, d, w.b_qk, p->big16, 3 * d, nsm, s, zz, rdev) == 0;
      }
      if (!merged) {
        { PROF(PC_GEMM_QKV);
          rc |= launch_gemm(GEMM_BIAS_F16, &p->m_r16, &m.w_qk, R, 2 * d, d, w.b_qk, nullptr, p->big16, 3 * d, 1.f, nsm, s, zz, rdev); }
        { PROF(PC_GEMM_QKV);
          rc |= launch_gemm(GEMM_BIAS_F16, &p->m_a16, &m.w_v, R, d, d, w.b_v, nullptr, p->big16 + 2 * d, 3 * d, 1.f, nsm, s, zz, rdev); }
      }
      { PROF(PC_ATTENTION);
        rc |= launch_attention(&p->m_qkv, p->plen, p->cu, p->o16, B, p->T2, c.n_heads, dk, d, nsm, s); }
    } else {
      // rel_pos (encoder.py:208-228): one projection GEMM -> [q+u | q+v | k | v], position scores inside the kernel
      { PROF(PC_LAYERNORM);
        launch_ln_f16(p->x, w.ln_att_g, w.ln_att_b, p->a16, R, rdev, 0, s); }
      { PROF(PC_GEMM_QKV);
        rc |= launch_gemm(GEMM_BIAS_F16, &p->m_a16, &m.w_qkv_rel, R, 4 * d, d, w.b_qkv_rel, nullptr, p->big16, 4 * d, 1.f, nsm, s, zz, rdev); }
      { PROF(PC_ATTENTION);
        rc |= launch_attention_relpos(&p->m_qkv4, &m.pos_proj, p->plen, p->cu, p->o16, B, p->T2, c.n_heads, dk, d, s); }
    }
    { PROF(PC_GEMM_PROJ);
      rc |= launch_gemm(GEMM_BIAS_RES_F32, &p->m_o16, &m.w_o, R, d, d, w.b_o, p->x, p->x, d, 1.f, nsm, s, zz, rdev); }
    // x += Conv(LN(x))                                           (encoder.py:489-491, 396-409)
    { PROF(PC_LAYERNORM);
      launch_ln_f16(p->x, w.ln_conv_g, w.ln_conv_b, p->a16, R, rdev, 0, s); }
    { PROF(PC_GEMM_GLU);
      rc |= launch_gemm(GEMM_BIAS_GLU_F16, &p->m_a16, &m.pw1, R, 2 * d, d, w.pw1_b, nullptr, p->g16, d, 1.f, nsm, s, zz, rdev); }
    { PROF(PC_DWCONV);
      if (c.conv_norm == 0)
        rc |= launch_dwconv_bn_silu(p->g16, w.dw_w, w.dw_b, p->len2, p->cu, p->plen, p->o16, B, p->T2, c.conv_kernel_size, s);
      else
        rc |= launch_dwconv_ln_silu(p->g16, w.dw_w, w.dw_b, w.cn_g, w.cn_b, p->len2, p->cu, p->row_b, p->row_t, rdev, p->o16, B, p->T2,
                                    c.conv_kernel_size, s); }
    { PROF(PC_GEMM_PROJ);
      rc |= launch_gemm(GEMM_BIAS_RES_F32, &p->m_o16, &m.pw2, R, d, d, w.pw2_b, p->x, p->x, d, 1.f, nsm, s, zz, rdev); }
    // x += 0.5 * FF2(LN(x))                                      (encoder.py:493-495)
    { PROF(PC_LAYERNORM);
      launch_ln_f16(p->x, w.ln_ff2_g, w.ln_ff2_b, p->a16, R, rdev, 0, s); }
    { PROF(PC_GEMM_FFN_UP);
      rc |= launch_gemm(GEMM_BIAS_SILU_F16, &p->m_a16, &m.ff2_w1, R, c.d_ff, d, w.ff2_b1, nullptr, p->big16, c.d_ff, 1.f, nsm, s, zz, rdev); }
    { PROF(PC_GEMM_FFN_DOWN);
      rc |= launch_gemm(GEMM_BIAS_RES_F32, &p->m_hid, &m.ff2_w2, R, d, c.d_ff, w.ff2_b2, p->x, p->x, d, 0.5f, nsm, s, 0, rdev); }
    // x = LN_out(x) (+ next layer's first LN fused)                (encoder.py:497)
    { PROF(PC_LAYERNORM);
      if (l + 1 < L)
        launch_ln_out_ln(p->x, w.ln_out_g, w.ln_out_b, h->layers[l + 1].ln_ff1_g, h->layers[l + 1].ln_ff1_b, p->x, p->a16, R, rdev, zz, s);
      else   // last layer: norm_out straight into the caller's padded [B, T', d] (frames that do not exist -> 0)
        launch_unpack_rows(p->x, w.ln_out_g, w.ln_out_b, p->cu, p->plen, enc, B, p->T2, zz, s); }
    if (rc) return fail(h, -4, "layer %d: a launch was rejected (rc=%d): %s", l, rc, cudaGetErrorString(cudaGetLastError()));
  }
  cudaMemcpyAsync(enc_len, p->len2, B * sizeof(int), cudaMemcpyDeviceToDevice, s);
  GAM_CHECK_LAUNCH(h, "encode");
  return 0;
}

int gam_ctc_greedy(gam_handle* h, const float* enc, const int32_t* enc_len, int32_t B, int32_t T, void* workspace,
                   int64_t workspace_bytes, int32_t* ids, int32_t* frames, int32_t* counts, int32_t max_out, void* stream) {
  const gam_config& c = h->cfg;
  if (c.head != 1) return fail(h, -1, "model has no CTC head");
  if (max_out < T) return fail(h, -1, "max_out (%d) must be >= T (%d)", max_out, T);
  if (max_out != T) return fail(h, -1, "ids/frames row pitch must equal T for the CTC path");
  const int64_t R = static_cast<int64_t>(B) * T;
  if (workspace_bytes < R * 4) return fail(h, -1, "workspace too small for CTC labels");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int* labels = static_cast<int*>(workspace);
  { PROF(PC_CTC_ARGMAX);
    launch_ctc_argmax(enc, h->w.ctc_w, h->w.ctc_b, labels, static_cast<int>(R), c.d_model, c.num_classes, s); }
  { PROF(PC_CTC_COLLAPSE);
    launch_ctc_collapse(labels, enc_len, B, T, c.num_classes - 1, ids, frames, counts, s); }
  GAM_CHECK_LAUNCH(h, "ctc_greedy");
  return 0;
}

int gam_rnnt_greedy(gam_handle* h, const float* enc, const int32_t* enc_len, int32_t B, int32_t T, void* workspace,
                    int64_t workspace_bytes, int32_t* ids, int32_t* frames, int32_t* counts, int32_t max_out, void* stream) {
  const gam_config& c = h->cfg;
  if (c.head != 2) return fail(h, -1, "model has no RNN-T head");
  if (c.pred_hidden != c.joint_hidden) return fail(h, -1, "pred_hidden != joint_hidden is not supported");
  const int64_t R = static_cast<int64_t>(B) * T;
  if (workspace_bytes < R * c.joint_hidden * 4) return fail(h, -1, "workspace too small for the RNN-T encoder projection");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  float* encproj = static_cast<float*>(workspace);
  { PROF(PC_RNNT_ENCPROJ);
    launch_sgemm_tn_bias(enc, h->w.rnnt_enc_w, h->w.rnnt_enc_b, encproj, static_cast<int>(R), c.joint_hidden, c.d_model, s); }
  PROF(PC_RNNT_GREEDY);
  const int rc = launch_rnnt_greedy_cluster(encproj, enc_len, h->w.rnnt_emb_gates, h->w.rnnt_whh_t, h->w.rnnt_wp_t, h->w.rnnt_bp,
                                            h->w.rnnt_wo, h->w.rnnt_bo, B, T, c.pred_hidden, c.num_classes, c.num_classes - 1,
                                            c.max_symbols, max_out, ids, frames, counts, s);
  if (rc > 0)
    return fail(h, -1, "rnnt: the greedy kernel is specialised for pred_hidden = joint_hidden = 320 and needs 16-CTA clusters "
                "(pred_hidden %d)", c.pred_hidden);
  if (rc < 0) return fail(h, -4, "rnnt cluster kernel launch failed: %s", cudaGetErrorString(cudaGetLastError()));
  GAM_CHECK_LAUNCH(h, "rnnt_greedy");
  return 0;
}

int gam_group_words(gam_handle* h, const int32_t* ids, const int32_t* frames, const int32_t* counts, int32_t B, int32_t max_out,
                    const uint8_t* token_flags, int32_t V, int32_t max_words, int32_t* word_start, int32_t* word_end,
                    int32_t* word_first_token, int32_t* word_tokens, int32_t* n_words, void* stream) {
  if (B < 0 || max_out <= 0 || max_words <= 0 || V <= 0) return fail(h, -1, "group_words: bad sizes");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  { PROF(PC_MISC);
    launch_group_words(ids, frames, counts, token_flags, B, V, max_out, max_words, word_start, word_end, word_first_token, word_tokens,
                       n_words, s); }
  GAM_CHECK_LAUNCH(h, "group_words");
  return 0;
}

int gam_comm_unique_id(uint8_t* out128) { return comm_unique_id(out128); }

int gam_comm_init(gam_handle* h, const uint8_t* id128, int32_t rank, int32_t nranks) {
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(h, -1, "comm_init: bad rank %d of %d", rank, nranks);
  if (h->comm) { comm_destroy(h->comm); h->comm = nullptr; }
  if (cudaSetDevice(h->device) != cudaSuccess) return fail(h, -11, "cudaSetDevice(%d) failed", h->device);
  const char* err = "";
  const int rc = comm_init(&h->comm, id128, rank, nranks, &err);
  if (rc != 0) return fail(h, -20, "NCCL communicator init failed (rank %d of %d): %s", rank, nranks, err);
  h->comm_rank = rank;
  h->comm_nranks = nranks;
  return 0;
}

int32_t gam_comm_nccl_version(void) { return comm_nccl_version(); }

int gam_gather_hyps(gam_handle* h, const int32_t* packed, int64_t n_int32, int32_t* gathered, void* stream) {
  if (!h->comm) return fail(h, -20, "gather_hyps: no communicator (call gam_comm_init first)");
  if (n_int32 <= 0) return fail(h, -1, "gather_hyps: empty payload");
  const char* err = "";
  { cudaStream_t s = static_cast<cudaStream_t>(stream);
    PROF(PC_MISC);
    if (comm_all_gather_i32(h->comm, packed, gathered, n_int32, s, &err) != 0) return fail(h, -20, "ncclAllGather failed: %s", err); }
  return 0;
}

int gam_profile_begin(gam_handle* h) {
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  h->prof_ev.clear();
  h->prof_cls.clear();
  h->prof = true;
  return 0;
}

int gam_profile_end(gam_handle* h, double* ms_per_class, int64_t* launches_per_class, int32_t n_classes) {
  h->prof = false;
  for (int i = 0; i < n_classes; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
  int rc = 0;
  for (size_t i = 0; i < h->prof_cls.size(); ++i) {
    float ms = 0.f;
    if (cudaEventSynchronize(h->prof_ev[2 * i + 1]) != cudaSuccess ||
        cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]) != cudaSuccess) { rc = -1; continue; }
    const int cls = h->prof_cls[i];
    if (cls < n_classes) { ms_per_class[cls] += ms; launches_per_class[cls] += 1; }
  }
  for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
  h->prof_ev.clear();
  h->prof_cls.clear();
  if (rc) return fail(h, -5, "profile: an event could not be read: %s", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

int gam_profile_class_count(void) { return PC_COUNT; }

const char* gam_profile_class_name(int32_t cls) {
  static const char* names[PC_COUNT] = {"logmel", "subsample_conv1", "gemm_conv2_implicit", "gemm_subsample_out", "gemm_ffn_up_silu",
                                        "gemm_ffn_down_res", "gemm_qkv", "gemm_proj_res", "gemm_pw1_glu", "layernorm", "attention",
                                        "dwconv_bn_silu", "ctc_head_argmax", "ctc_collapse", "rnnt_enc_proj", "rnnt_greedy", "misc"};
  return (cls >= 0 && cls < PC_COUNT) ? names[cls] : "?";
}

int gam_test_gemm(gam_handle* h, int32_t kind, const void* A, const void* W, const float* bias, const float* res, void* out,
                  int32_t M, int32_t N, int32_t K, int32_t ldo, float scale, void* stream) {
  CUtensorMap ta, tw;
  int rc = make_tmap_2d_f16(&ta, A, M, K, K, 128, 64);
  rc |= make_tmap_2d_f16(&tw, W, N, K, K, 128, 64);
  if (rc) return fail(h, -2, "tensor map encode failed (rc=%d)", rc);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    PROF(PC_MISC);
    rc = launch_gemm(kind, &ta, &tw, M, N, K, bias, res, out, ldo, scale, h->num_sms, s);
  }
  if (rc) return fail(h, -4, "gemm launch rejected (rc=%d): %s", rc, cudaGetErrorString(cudaGetLastError()));
  GAM_CHECK_LAUNCH(h, "test_gemm");
  return 0;
}

int gam_test_attention(gam_handle* h, const void* qkv, const int32_t* klen, void* out, int32_t B, int32_t T, void* stream) {
  const gam_config& c = h->cfg;
  CUtensorMap tq;
  const uint64_t d = c.d_model;
  int rc = make_tmap_2d_f16(&tq, qkv, static_cast<uint64_t>(B) * T, 3 * d, 3 * d, 128, 64);
  if (rc) return fail(h, -2, "tensor map encode failed (rc=%d)", rc);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    PROF(PC_ATTENTION);
    rc = launch_attention(&tq, klen, nullptr, static_cast<__half*>(out), B, T, c.n_heads, c.d_model / c.n_heads, c.d_model, h->num_sms, s);
  }
  if (rc) return fail(h, -4, "attention launch rejected (T=%d)", T);
  GAM_CHECK_LAUNCH(h, "test_attention");
  return 0;
}

int gam_test_attention_relpos(gam_handle* h, const void* qkv, const void* pos, const int32_t* klen, void* out, int32_t B,
                              int32_t T, void* stream) {
  const gam_config& c = h->cfg;
  CUtensorMap tq, tp;
  const uint64_t d = c.d_model;
  int rc = make_tmap_2d_f16(&tq, qkv, static_cast<uint64_t>(B) * T, 4 * d, 4 * d, 128, 64);
  rc |= make_tmap_2d_f16(&tp, pos, 2 * kRelPosMaxT - 1, d, d, 128, 64);
  if (rc) return fail(h, -2, "tensor map encode failed (rc=%d)", rc);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    PROF(PC_ATTENTION);
    rc = launch_attention_relpos(&tq, &tp, klen, nullptr, static_cast<__half*>(out), B, T, c.n_heads, c.d_model / c.n_heads, c.d_model, s);
  }
  if (rc) return fail(h, -4, "rel_pos attention launch rejected (T=%d, rc=%d)", T, rc);
  GAM_CHECK_LAUNCH(h, "test_attention_relpos");
  return 0;
}

int gam_test_attention_varlen(gam_handle* h, const void* qkv, const void* pos, const int32_t* klen, const int32_t* cu, void* out,
                              int32_t B, int32_t T, int32_t rows, void* stream) {
  const gam_config& c = h->cfg;
  if (!klen || !cu || rows <= 0) return fail(h, -1, "attention_varlen: klen, cu and rows are required");
  CUtensorMap tq, tp;
  const uint64_t d = c.d_model, parts = pos ? 4 : 3;
  int rc = make_tmap_2d_f16(&tq, qkv, static_cast<uint64_t>(rows), parts * d, parts * d, 128, 64);
  if (pos) rc |= make_tmap_2d_f16(&tp, pos, 2 * kRelPosMaxT - 1, d, d, 128, 64);
  if (rc) return fail(h, -2, "tensor map encode failed (rc=%d)", rc);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  {
    PROF(PC_ATTENTION);
    rc = pos ? launch_attention_relpos(&tq, &tp, klen, cu, static_cast<__half*>(out), B, T, c.n_heads, c.d_model / c.n_heads, c.d_model, s)
             : launch_attention(&tq, klen, cu, static_cast<__half*>(out), B, T, c.n_heads, c.d_model / c.n_heads, c.d_model, h->num_sms, s);
  }
  if (rc) return fail(h, -4, "varlen attention launch rejected (T=%d, rc=%d)", T, rc);
  GAM_CHECK_LAUNCH(h, "test_attention_varlen");
  return 0;
}

}  // extern "C"
