// Whole-model sequencer: the complete kernel sequence of one STTran / DSG-DETR forward, fused loss and backward pass is
// enqueued by ONE host call (include/nlv_b200.h, "Whole-model sequencer").
//
// Reference control flow replaced: lib/sttran.py:375-411 (STTran.forward), :173-184 (object classifier, sgdet / wks),
// :381-399 (pair tokens), lib/transformer_wk.py:130-217 (pad / window / 'latter' plumbing: four python loops with device
// syncs per frame), lib/transformer.py:20-30,49-58 (layers), lib/dsg_detr.py:514-572, tools/train_STTran.py:169-189 (losses).
//
// Memory: one caller-provided workspace, bump-allocated from both ends — tensors that live until the end of the step
// (saved activations, layer outputs, gradients in flight) grow from the bottom, per-layer temporaries from the top and are
// released when the layer returns.  A dry run of the same code (no launches) sizes the workspace.
#include <string.h>

#include <vector>

#include "common.cuh"

namespace nlv {
namespace {

constexpr int D = 1936, HEADS = 8, HD = 242, DFF = 2048;
constexpr int K_ = NLV_MAJOR_K, MN_ = NLV_MAJOR_MN;

struct T {   // row-major matrix view
  void* p = nullptr;
  int dt = NLV_F32;
  long long rows = 0;
  int cols = 0, ld = 0;
  int esz() const { return dt == NLV_BF16 ? 2 : 4; }
  bool ok() const { return p != nullptr; }
  float* f() const { return reinterpret_cast<float*>(p); }
  T cs(int c0, int n) const { T t = *this; t.p = (char*)p + (size_t)c0 * esz(); t.cols = n; return t; }
  T rs(long long r0, long long n) const { T t = *this; t.p = (char*)p + (size_t)r0 * ld * esz(); t.rows = n; return t; }
  T view(long long r, int c) const { T t = *this; t.rows = r; t.cols = c; t.ld = c; return t; }   // contiguous only
};

inline T mk(const void* p, int dt, long long rows, int cols) {
  T t; t.p = const_cast<void*>(p); t.dt = dt; t.rows = rows; t.cols = cols; t.ld = cols; return t;
}

struct EncCtx { T xop, qkv, o, lse, y1, m1, r1, x1op, h, y2, m2, r2; };
struct DecCtx { T xop, xpop, qkv, o, lse, y, m3, r3, top, h; };
struct OcCtx { T objfeat, cs, pos_bn, mean0, var0, h1, h2, mean1, var1; };
struct PtCtx { T feat_op, uf_op, col1, c1, mean2, var2, arg, xmax, col2, c2, mean6, var6, vr_in; };

// ---- optional per-call timing (bench.py roofline legs): CUDA events around every kernel entry the sequencer makes --------
struct ProfRec { char name[40]; cudaEvent_t e0, e1; double flops; double units; int m, n, k; int dt; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
double g_next_flops = 0, g_next_units = 0;
int g_next_m = 0, g_next_n = 0, g_next_k = 0, g_next_dt = 0;
struct ProfScope {
  bool on; ProfRec r; cudaStream_t st;
  ProfScope(void* stream, const char* call) : on(g_prof_on), st((cudaStream_t)stream) {
    if (!on) return;
    int i = 0;
    for (; call[i] && call[i] != '(' && i < 39; ++i) r.name[i] = call[i];
    r.name[i] = 0;
    cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
    cudaEventRecord(r.e0, st);
  }
  ~ProfScope() {
    if (!on) return;
    cudaEventRecord(r.e1, st);
    r.flops = g_next_flops; r.units = g_next_units; r.m = g_next_m; r.n = g_next_n; r.k = g_next_k; r.dt = g_next_dt;
    g_next_flops = g_next_units = 0; g_next_m = g_next_n = g_next_k = g_next_dt = 0;
    g_prof.push_back(r);
  }
};

}  // namespace
int set_seg(int* p, int rows, cudaStream_t s);   // util.cu
}  // namespace nlv

using namespace nlv;

struct nlv_session {
  // ---- arena -----------------------------------------------------------------------------------------
  char* base = nullptr;
  long long size = 0, lo = 0, hi = 0, peak = 0;
  bool dry = true;
  bool overflow = false;
  void* st = nullptr;
  // ---- copies of the caller's descriptors --------------------------------------------------------------
  nlv_model M{};
  nlv_batch B{};
  std::vector<const float*> params;
  std::vector<const void*> pop;
  std::vector<long long> goff;
  // called (host side, between kernel enqueues) when every gradient of the temporal / global layers — the tail of the
  // gradient buffer — has been enqueued: the data-parallel trainer starts their all-reduce under the rest of the backward
  void (*tail_hook)(void*) = nullptr;
  void* tail_hook_user = nullptr;
  int flags = 0;
  bool want_ctx = false, training = false, have_fwd = false, xf_only = false;
  int AD = NLV_BF16;
  // ---- per-step state ----------------------------------------------------------------------------------
  std::vector<T> wop;
  T w_c0, w_c4, w_c4t, w_vr, w26, b26;
  OcCtx oc;
  PtCtx pt;
  bool pool_fused = false;
  std::vector<EncCtx> enc;    // STTran: spatial encoder layers; DSG: [local, global0..2]
  std::vector<DecCtx> dec;
  T masks, rel, local_out, xf_out, logits26, obj_logits, att, spa, con, loss, d26, dobj, dx_in;
  long long N = 0, R = 0, Mg = 0;

  // ---- arena ops ---------------------------------------------------------------------------------------
  void arena(void* ws, long long bytes, bool dry_) {
    dry = dry_;
    if (dry) { base = reinterpret_cast<char*>(4096); size = 1ll << 60; }
    else { base = reinterpret_cast<char*>(ws); size = bytes; }
    const long long mis = (256 - (reinterpret_cast<uintptr_t>(base) & 255)) & 255;
    lo = mis;
    hi = (long long)((reinterpret_cast<uintptr_t>(base) + (uintptr_t)size) & ~(uintptr_t)255) - (long long)reinterpret_cast<uintptr_t>(base);
    peak = 0; overflow = false;
  }
  void* take_lo(long long bytes) {
    bytes = (bytes + 255) & ~255ll;
    void* p = base + lo;
    lo += bytes;
    if (lo > hi) overflow = true;
    if (lo + (size - hi) > peak) peak = lo + (size - hi);
    return overflow ? nullptr : p;
  }
  void* take_hi(long long bytes) {
    bytes = (bytes + 255) & ~255ll;
    hi -= bytes;
    if (lo > hi) overflow = true;
    if (lo + (size - hi) > peak) peak = lo + (size - hi);
    return overflow ? nullptr : base + hi;
  }
  // keep: lives to the end of the step; tmp: released by the enclosing Scope
  T keep(long long rows, int cols, int dt, int ld = 0) {
    T t; t.dt = dt; t.rows = rows; t.cols = cols; t.ld = ld ? ld : cols;
    t.p = take_lo((long long)rows * t.ld * t.esz() + 16);
    return t;
  }
  T tmp(long long rows, int cols, int dt, int ld = 0) {
    T t; t.dt = dt; t.rows = rows; t.cols = cols; t.ld = ld ? ld : cols;
    t.p = take_hi((long long)rows * t.ld * t.esz() + 16);
    return t;
  }
  T ctx(long long rows, int cols, int dt, int ld = 0) { return want_ctx ? keep(rows, cols, dt, ld) : tmp(rows, cols, dt, ld); }
  struct Scope {
    nlv_session* s; long long m;
    explicit Scope(nlv_session* s_) : s(s_), m(s_->hi) {}
    ~Scope() { s->hi = m; }
  };

  const float* P(int slot) const { return params[slot]; }
  float* G(int slot) const { return (slot < (int)goff.size() && goff[slot] >= 0) ? M.grad_base + goff[slot] : nullptr; }
  static int LS(int layer, int which) { return NLV_P_LAYER0 + layer * NLV_P_LAYER_STRIDE + which; }

  int run_forward();
  int run_loss();
  int run_backward(const float* d26_in, const float* dobj_in);
  int prepare_weights();
  int mm(const T& a, int am, const T& b, int bm, const T& out, const float* bias = nullptr, const T* residual = nullptr,
         bool relu = false, const T* gate = nullptr, bool exact = false, const nlv_dropout* drop = nullptr, float gate_scale = 1.f);
  bool dropping = false;
  // dropout site of a layer: 0 = attention-output residual, 1 = FFN inner (after ReLU), 2 = FFN-output residual, 3 = attention weights;
  // 8000 = object-classifier pos_embed, 8001 = DSG-DETR positional encoder
  nlv_dropout site(int layer, int which) const {
    nlv_dropout d;
    memset(&d, 0, sizeof(d));
    if (dropping) {
      d.thr16 = (unsigned)(M.dropout_p * 65536.f + 0.5f);
      d.scale = 1.f / (1.f - M.dropout_p);
      d.seed_lo = (unsigned)M.seed; d.seed_hi = (unsigned)(M.seed >> 32);
      d.stream = (unsigned)(layer * 8 + which);
    }
    return d;
  }
  int opnd(const T& x, T* out);
  int tma_ready(const T& x, T* out);
  int split3(const T& x, int major, int pattern, T* out);
  T W(int slot, long long rows, int cols) const {
    if (AD == NLV_BF16 && wop[slot].ok()) return wop[slot];
    return mk(params[slot], NLV_F32, rows, cols);
  }
  // bf16 path: the data gradient of the 3x3 conv is an implicit GEMM (NLV_CONV_IMPLICIT=0: column-gradient product + col2im)
  bool implicit_dgrad() const {
    static const int env = [] { const char* e = getenv("NLV_CONV_IMPLICIT"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
    return AD == NLV_BF16 && env != 0;
  }
  // bf16 path: the first stage of the mask branch runs on the fused kernels of maskconv.cu (NLV_MASKCONV=0: the im2col + GEMM route)
  bool fused_mask_conv() const {
    static const int env = [] { const char* e = getenv("NLV_MASKCONV"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
    return AD == NLV_BF16 && env != 0;
  }
  int bn_fwd(const T& x, const int* seg, const int* row_seg, int row_div, int slot_w, float momentum, bool relu, const T& y, T* mean, T* var);
  int bn_bwd(const T& dy, const T& x, const T* yout, const int* seg, const int* row_seg, int row_div, int slot_w, const T& mean, const T& var,
             bool gate_by_x, const T& dx, float* dx_colsum = nullptr);
  int lin_grads(const T& dy_op, const T& x_op, const T& dy_bias, int wslot, int bslot);
  int encoder_fwd(int layer, const T& x, const T& xop, const int* work, int n_work, bool out_op, T* x2, T* x2op, EncCtx* c);
  int encoder_bwd(int layer, const EncCtx& c, const T& dx2, const int* work, int n_work, bool need_dx, T* dx);
  int decoder_fwd(int layer, const T& x, const T& xop, const T& xpop, T* out, DecCtx* c);
  int decoder_bwd(int layer, const DecCtx& c, const T& dout, T* dx);
  int object_classifier_fwd();
  int object_classifier_bwd(const T& dlogits);
  int pair_tokens_fwd(const T& feat_op);
  int pair_tokens_bwd(const T& drel);
  int sttran_transformer_fwd(const T& rel_in, T* out);
  int sttran_transformer_bwd(const T& dout, T* drel);
  int dsg_transformer_fwd(const T& rel_in, T* out);
  int dsg_transformer_bwd(const T& dout, T* drel);
  int heads_fwd(const T& x);
  int heads_bwd(const T& x, const T& dlogits, T* dx);
  int setup(const nlv_model* model, const nlv_batch* batch, int flags_);
};

// every kernel entry goes through this: skipped in the dry run, error propagated otherwise
#define RUN(call)                                  \
  do {                                             \
    if (!dry) {                                    \
      nlv::ProfScope _ps(st, #call);               \
      const int _rc = (call);                      \
      if (_rc != NLV_OK) return _rc;               \
    }                                              \
  } while (0)
#define CK(call)                                   \
  do {                                             \
    const int _rc = (call);                        \
    if (_rc != NLV_OK) return _rc;                 \
  } while (0)
#define OOM_CHECK()                                                                                         \
  do {                                                                                                      \
    if (!dry && overflow) {                                                                                 \
      nlv::set_error("sequencer workspace too small (%lld bytes given; call nlv_session_plan first)", size); \
      return NLV_ERR_INVALID_ARGUMENT;                                                                      \
    }                                                                                                       \
  } while (0)

// ================================================================================================================
// GEMM policy (precision modes) — the same rules for every product of the model
// ================================================================================================================
int nlv_session::opnd(const T& x, T* out) {
  if (M.precision == NLV_PREC_BF16 && x.dt != NLV_BF16) {
    T y = tmp(x.rows, x.cols, NLV_BF16, (x.cols + 7) / 8 * 8);
    RUN(nlv_convert(x.p, x.dt, x.ld, y.p, y.dt, y.ld, x.rows, x.cols, st));
    *out = y;
    return NLV_OK;
  }
  *out = x;
  return NLV_OK;
}

int nlv_session::tma_ready(const T& x, T* out) {
  T y = x;
  if ((x.dt != NLV_BF16) || (x.ld & 7) != 0 || (reinterpret_cast<uintptr_t>(x.p) & 15) != 0) {
    y = tmp(x.rows, x.cols, NLV_BF16, (x.cols + 7) / 8 * 8);
    if (y.ld != y.cols) RUN(nlv_zero_bytes(y.p, (long long)y.rows * y.ld * 2, st));
    RUN(nlv_convert(x.p, x.dt, x.ld, y.p, y.dt, y.ld, x.rows, x.cols, st));
  }
  *out = y;
  return NLV_OK;
}

// fp32 [r,c] -> three bf16 blocks along K (K-major operands: columns; MN-major operands: rows)
int nlv_session::split3(const T& x, int major, int pattern, T* out) {
  T src = x;
  if (x.dt != NLV_F32) {
    src = tmp(x.rows, x.cols, NLV_F32);
    RUN(nlv_convert(x.p, x.dt, x.ld, src.p, NLV_F32, src.ld, x.rows, x.cols, st));
  }
  T y;
  if (major == K_) y = tmp(src.rows, 3 * src.cols, NLV_BF16, (3 * src.cols + 7) / 8 * 8);
  else y = tmp(3 * src.rows, src.cols, NLV_BF16, (src.cols + 7) / 8 * 8);
  RUN(nlv_split3(src.f(), src.ld, src.rows, src.cols, y.p, y.ld, major == K_ ? 1 : 0, pattern, st));
  *out = y;
  return NLV_OK;
}

int nlv_session::mm(const T& a_in, int am, const T& b_in, int bm, const T& out, const float* bias, const T* residual, bool relu,
                    const T* gate, bool exact, const nlv_dropout* drop, float gate_scale) {
  T a = a_in, b = b_in;
  const long long m = am == K_ ? a.rows : a.cols;
  const long long kdim = am == K_ ? a.cols : a.rows;
  const long long n = bm == K_ ? b.rows : b.cols;
  if (m == 0 || n == 0) return NLV_OK;
  int force_simt = 0;
  if (exact && M.precision == NLV_PREC_BF16 && a.dt == NLV_F32 && b.dt == NLV_F32 && kdim >= 256) {
    CK(split3(a, am, 0, &a));   // fp32-grade product on the tensor cores (three bf16 terms per operand, error ~2^-17)
    CK(split3(b, bm, 1, &b));
  } else if (exact || M.precision == NLV_PREC_FP32) {
    if (a.dt != NLV_F32) { T t = tmp(a.rows, a.cols, NLV_F32); RUN(nlv_convert(a.p, a.dt, a.ld, t.p, NLV_F32, t.ld, a.rows, a.cols, st)); a = t; }
    if (b.dt != NLV_F32) { T t = tmp(b.rows, b.cols, NLV_F32); RUN(nlv_convert(b.p, b.dt, b.ld, t.p, NLV_F32, t.ld, b.rows, b.cols, st)); b = t; }
  } else if (M.precision == NLV_PREC_BF16X3) {
    CK(split3(a, am, 0, &a));
    CK(split3(b, bm, 1, &b));
  } else {
    CK(tma_ready(a, &a));
    CK(tma_ready(b, &b));
  }
  OOM_CHECK();
  nlv_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.a = a.p; g.b = b.p; g.d = out.p; g.bias = bias;
  g.m = (int)m; g.n = (int)n; g.k = (int)(am == K_ ? a.cols : a.rows);
  g.lda = a.ld; g.ldb = b.ld; g.ldd = out.ld;
  g.a_major = am; g.b_major = bm;
  g.ab_dtype = a.dt | force_simt; g.d_dtype = out.dt;
  g.relu = relu ? 1 : 0;
  if (residual != nullptr) { g.residual = residual->p; g.ldr = residual->ld; g.r_dtype = residual->dt; }
  if (gate != nullptr) { g.gate = gate->p; g.ldg = gate->ld; g.gate_dtype = gate->dt; g.gate_scale = gate_scale; }
  if (drop != nullptr) g.drop = *drop;
  if (g_prof_on) {   // algorithmic FLOPs of the product as the model states it (a bf16x3 product still counts 2mnk)
    g_next_flops = 2.0 * (double)m * (double)n * (double)kdim; g_next_m = (int)m; g_next_n = (int)n; g_next_k = (int)kdim; g_next_dt = a.dt;
  }
  RUN(nlv_gemm(&g, st));
  return NLV_OK;
}

// dW = dY^T X (both operands MN-major) straight into the gradient slot; db += column sums
int nlv_session::lin_grads(const T& dy_op, const T& x_op, const T& dy_bias, int wslot, int bslot) {
  const T gw = mk(G(wslot), NLV_F32, dy_op.cols, x_op.cols);
  CK(mm(dy_op, MN_, x_op, MN_, gw));
  if (bslot < 0) return NLV_OK;       // the bias gradient was accumulated by the kernel that produced dy (fused LayerNorm backward)
  g_next_units = (double)dy_bias.rows * dy_bias.cols * dy_bias.esz();
  RUN(nlv_colsum(dy_bias.p, dy_bias.dt, dy_bias.ld, dy_bias.rows, dy_bias.cols, nullptr, 1, G(bslot), st));
  return NLV_OK;
}

// ================================================================================================================
// transformer layers
// ================================================================================================================
// Post-norm encoder layer (lib/transformer.py:20-30; nn.TransformerEncoderLayer of lib/dsg_detr.py:502-506)
int nlv_session::encoder_fwd(int layer, const T& x, const T& xop, const int* work, int n_work, bool out_op, T* x2_out, T* x2op_out,
                             EncCtx* c) {
  const long long Mr = x.rows;
  const bool b16 = AD == NLV_BF16;
  // outputs first (bottom of the arena), temporaries of this layer are released on return
  T x2 = keep(Mr, D, NLV_F32);
  T x2op = (b16 && out_op) ? keep(Mr, D, NLV_BF16) : T();
  Scope sc(this);
  T qkv = ctx(Mr, 3 * D, AD), o = ctx(Mr, D, AD), lse = want_ctx ? keep(Mr * HEADS, 1, NLV_F32) : T();
  T y1 = ctx(Mr, D, NLV_F32), m1 = ctx(Mr, 1, NLV_F32), r1 = ctx(Mr, 1, NLV_F32);
  T x1 = tmp(Mr, D, NLV_F32), x1op = b16 ? ctx(Mr, D, NLV_BF16) : T();
  if (!b16 && want_ctx) x1 = keep(Mr, D, NLV_F32);
  T h = ctx(Mr, DFF, AD), y2 = ctx(Mr, D, NLV_F32), m2 = ctx(Mr, 1, NLV_F32), r2 = ctx(Mr, 1, NLV_F32);
  OOM_CHECK();
  const nlv_dropout d0 = site(layer, 0), d1 = site(layer, 1), d2 = site(layer, 2), d3 = site(layer, 3);
  CK(mm(xop, K_, W(LS(layer, NLV_L_INPROJ_W), 3 * D, D), K_, qkv, P(LS(layer, NLV_L_INPROJ_B))));
  g_next_units = 4.0 * (double)Mr * D * qkv.esz(); g_next_dt = qkv.dt;   // algorithmic bytes: Q, K, V in, O out
  // additive_mask (STTran spatial encoder, inference): the frame's padded keys stay in the softmax (lib/transformer_wk.py:154 under
  // torch 1.10.1); their key / value are the K / V slices of the in-projection bias
  const bool padkeys = M.additive_mask != 0 && M.arch == NLV_ARCH_STTRAN && work == B.local_work;
  const float* inb = P(LS(layer, NLV_L_INPROJ_B));
  RUN(nlv_attn_fwd_padkeys(qkv.p, qkv.ld, (char*)qkv.p + (size_t)D * qkv.esz(), qkv.ld, (char*)qkv.p + (size_t)2 * D * qkv.esz(), qkv.ld,
                           qkv.dt, HD, HEADS, 1.0f / sqrtf((float)HD), work, n_work, o.p, o.ld, o.dt, lse.ok() ? lse.f() : nullptr, &d3,
                           padkeys ? inb + D : nullptr, padkeys ? inb + 2 * D : nullptr, st));
  CK(mm(o, K_, W(LS(layer, NLV_L_OUTPROJ_W), D, D), K_, y1, P(LS(layer, NLV_L_OUTPROJ_B)), &x, false, nullptr, false, &d0));
  g_next_units = (double)Mr * D * (4 + 4 + (b16 ? 2 : 0));
  RUN(nlv_layernorm_fwd(y1.f(), Mr, D, P(LS(layer, NLV_L_NORMA_W)), P(LS(layer, NLV_L_NORMA_B)), 1e-5f, x1.f(),
                        b16 ? x1op.p : nullptr, NLV_BF16, m1.f(), r1.f(), st));
  const T& x1o = b16 ? x1op : x1;
  CK(mm(x1o, K_, W(LS(layer, NLV_L_LIN1_W), DFF, D), K_, h, P(LS(layer, NLV_L_LIN1_B)), nullptr, true, nullptr, false, &d1));
  CK(mm(h, K_, W(LS(layer, NLV_L_LIN2_W), D, DFF), K_, y2, P(LS(layer, NLV_L_LIN2_B)), &x1, false, nullptr, false, &d2));
  g_next_units = (double)Mr * D * (4 + 4 + (b16 ? 2 : 0));
  RUN(nlv_layernorm_fwd(y2.f(), Mr, D, P(LS(layer, NLV_L_NORMB_W)), P(LS(layer, NLV_L_NORMB_B)), 1e-5f, x2.f(),
                        x2op.ok() ? x2op.p : nullptr, NLV_BF16, m2.f(), r2.f(), st));
  if (c != nullptr) { c->xop = xop; c->qkv = qkv; c->o = o; c->lse = lse; c->y1 = y1; c->m1 = m1; c->r1 = r1; c->x1op = x1o; c->h = h;
                      c->y2 = y2; c->m2 = m2; c->r2 = r2; }
  *x2_out = x2;
  *x2op_out = x2op.ok() ? x2op : x2;
  return NLV_OK;
}

int nlv_session::encoder_bwd(int layer, const EncCtx& c, const T& dx2, const int* work, int n_work, bool need_dx, T* dx_out) {
  const long long Mr = dx2.rows;
  const bool b16 = AD == NLV_BF16;
  T dx = need_dx ? keep(Mr, D, NLV_F32) : T();
  Scope sc(this);
  const nlv_dropout d0 = site(layer, 0), d1 = site(layer, 1), d2 = site(layer, 2), d3 = site(layer, 3);
  T dy2 = tmp(Mr, D, NLV_F32), dy2op = b16 ? tmp(Mr, D, NLV_BF16) : T();
  OOM_CHECK();
  g_next_units = (double)Mr * D * (4 + 4 + 4 + (b16 ? 2 : 0));
  // dy2 = gradient of the residual stream; dy2op = its copy behind the FFN-output dropout (mask * dy2 / (1 - p) when dropping)
  RUN(nlv_layernorm_bwd_fused(dx2.f(), c.y2.f(), c.m2.f(), c.r2.f(), P(LS(layer, NLV_L_NORMB_W)), Mr, D, dy2.f(), b16 ? dy2op.p : nullptr,
                              NLV_BF16, G(LS(layer, NLV_L_NORMB_W)), G(LS(layer, NLV_L_NORMB_B)), G(LS(layer, NLV_L_LIN2_B)), &d2, st));
  const T& dy2o = b16 ? dy2op : dy2;
  CK(lin_grads(dy2o, c.h, dy2, LS(layer, NLV_L_LIN2_W), -1));
  T dh = tmp(Mr, DFF, AD);
  CK(mm(dy2o, K_, W(LS(layer, NLV_L_LIN2_W), D, DFF), MN_, dh, nullptr, nullptr, false, &c.h, false, nullptr,
        dropping ? d1.scale : 1.f));   // ReLU (+ inner dropout) backward fused: h > 0 <=> kept and active
  CK(lin_grads(dh, c.x1op, dh, LS(layer, NLV_L_LIN1_W), LS(layer, NLV_L_LIN1_B)));
  T dx1 = tmp(Mr, D, NLV_F32);
  CK(mm(dh, K_, W(LS(layer, NLV_L_LIN1_W), DFF, D), MN_, dx1, nullptr, &dy2));
  T dy1 = tmp(Mr, D, NLV_F32), dy1op = b16 ? tmp(Mr, D, NLV_BF16) : T();
  OOM_CHECK();
  g_next_units = (double)Mr * D * (4 + 4 + 4 + (b16 ? 2 : 0));
  RUN(nlv_layernorm_bwd_fused(dx1.f(), c.y1.f(), c.m1.f(), c.r1.f(), P(LS(layer, NLV_L_NORMA_W)), Mr, D, dy1.f(), b16 ? dy1op.p : nullptr,
                              NLV_BF16, G(LS(layer, NLV_L_NORMA_W)), G(LS(layer, NLV_L_NORMA_B)), G(LS(layer, NLV_L_OUTPROJ_B)), &d0, st));
  const T& dy1o = b16 ? dy1op : dy1;
  CK(lin_grads(dy1o, c.o, dy1, LS(layer, NLV_L_OUTPROJ_W), -1));
  T d_o = tmp(Mr, D, AD);
  CK(mm(dy1o, K_, W(LS(layer, NLV_L_OUTPROJ_W), D, D), MN_, d_o));
  T dqkv = tmp(Mr, 3 * D, AD), delta = tmp(Mr * HEADS, 1, NLV_F32);
  OOM_CHECK();
  const T& q = c.qkv;
  const size_t e = q.esz();
  g_next_units = 8.0 * (double)Mr * D * q.esz(); g_next_dt = q.dt;   // Q, K, V, O, dO in; dQ, dK, dV out
  const int n_long = !B.work_sorted ? -1 : (work == B.local_work ? B.n_local_long : (work == B.cls_work ? B.n_cls_long : -1));
  RUN(nlv_attn_bwd_sorted(q.p, q.ld, (char*)q.p + D * e, q.ld, (char*)q.p + 2 * D * e, q.ld, q.dt, HD, HEADS, 1.0f / sqrtf((float)HD), work, n_work,
                        n_long, c.o.p, c.o.ld, c.o.dt, d_o.p, d_o.ld, d_o.dt, c.lse.f(), delta.f(), dqkv.p, dqkv.ld, (char*)dqkv.p + D * e, dqkv.ld,
                        (char*)dqkv.p + 2 * D * e, dqkv.ld, dqkv.dt, &d3, st));
  CK(lin_grads(dqkv, c.xop, dqkv, LS(layer, NLV_L_INPROJ_W), LS(layer, NLV_L_INPROJ_B)));
  if (need_dx) {
    CK(mm(dqkv, K_, W(LS(layer, NLV_L_INPROJ_W), 3 * D, D), MN_, dx, nullptr, &dy1));
    *dx_out = dx;
  }
  return NLV_OK;
}

// Temporal decoder layer (lib/transformer.py:49-58): q = k = x + pos, v = x; LayerNorm after the attention residual, plain
// residual after the FFN.  xpop = operand form of x + pos.
int nlv_session::decoder_fwd(int layer, const T& x, const T& xop, const T& xpop, T* out_, DecCtx* c) {
  const long long Mr = x.rows;
  const bool b16 = AD == NLV_BF16;
  T out = keep(Mr, D, NLV_F32);
  Scope sc(this);
  T qkv = ctx(Mr, 3 * D, AD), o = ctx(Mr, D, AD), lse = want_ctx ? keep(Mr * HEADS, 1, NLV_F32) : T();
  T y = ctx(Mr, D, NLV_F32), m3 = ctx(Mr, 1, NLV_F32), r3 = ctx(Mr, 1, NLV_F32);
  T t = tmp(Mr, D, NLV_F32), top = b16 ? ctx(Mr, D, NLV_BF16) : T();
  if (!b16 && want_ctx) t = keep(Mr, D, NLV_F32);
  T h = ctx(Mr, DFF, AD);
  OOM_CHECK();
  const nlv_dropout d0 = site(layer, 0), d1 = site(layer, 1), d2 = site(layer, 2), d3 = site(layer, 3);
  const T win = W(LS(layer, NLV_L_INPROJ_W), 3 * D, D);
  const float* bin = P(LS(layer, NLV_L_INPROJ_B));
  CK(mm(xpop, K_, win.rs(0, 2 * D), K_, qkv.cs(0, 2 * D), bin));
  CK(mm(xop, K_, win.rs(2 * D, D), K_, qkv.cs(2 * D, D), bin + 2 * D));
  g_next_units = 4.0 * (double)Mr * D * qkv.esz(); g_next_dt = qkv.dt;   // algorithmic bytes: Q, K, V in, O out
  RUN(nlv_attn_fwd_drop(qkv.p, qkv.ld, (char*)qkv.p + (size_t)D * qkv.esz(), qkv.ld, (char*)qkv.p + (size_t)2 * D * qkv.esz(), qkv.ld, qkv.dt,
                        HD, HEADS, 1.0f / sqrtf((float)HD), B.glob_work, B.n_glob_work, o.p, o.ld, o.dt, lse.ok() ? lse.f() : nullptr, &d3, st));
  CK(mm(o, K_, W(LS(layer, NLV_L_OUTPROJ_W), D, D), K_, y, P(LS(layer, NLV_L_OUTPROJ_B)), &x, false, nullptr, false, &d0));
  g_next_units = (double)Mr * D * (4 + 4 + (b16 ? 2 : 0));
  RUN(nlv_layernorm_fwd(y.f(), Mr, D, P(LS(layer, NLV_L_NORMA_W)), P(LS(layer, NLV_L_NORMA_B)), 1e-5f, t.f(), b16 ? top.p : nullptr,
                        NLV_BF16, m3.f(), r3.f(), st));
  const T& to = b16 ? top : t;
  CK(mm(to, K_, W(LS(layer, NLV_L_LIN1_W), DFF, D), K_, h, P(LS(layer, NLV_L_LIN1_B)), nullptr, true, nullptr, false, &d1));
  CK(mm(h, K_, W(LS(layer, NLV_L_LIN2_W), D, DFF), K_, out, P(LS(layer, NLV_L_LIN2_B)), &t, false, nullptr, false, &d2));
  if (c != nullptr) { c->xop = xop; c->xpop = xpop; c->qkv = qkv; c->o = o; c->lse = lse; c->y = y; c->m3 = m3; c->r3 = r3; c->top = to; c->h = h; }
  *out_ = out;
  return NLV_OK;
}

// dx = dqkv W_in + dy (one K = 5808 product: d(x+pos) W_qk and dv W_v share the output); the position-embedding gradient
// is the per-slot column sum of dq|dk times W_qk — a [2 x 3872] x [3872 x 1936] product instead of a pass over [Mg, 1936]
int nlv_session::decoder_bwd(int layer, const DecCtx& c, const T& dout, T* dx_out) {
  const long long Mr = dout.rows;
  const bool b16 = AD == NLV_BF16;
  T dx = keep(Mr, D, NLV_F32);
  Scope sc(this);
  const nlv_dropout d0 = site(layer, 0), d1 = site(layer, 1), d2 = site(layer, 2), d3 = site(layer, 3);
  T doutop;
  if (dropping) {     // operand behind the FFN-output dropout: mask * dout / (1 - p)
    doutop = tmp(Mr, D, AD);
    OOM_CHECK();
    RUN(nlv_dropout_apply(dout.p, dout.dt, dout.ld, doutop.p, doutop.dt, doutop.ld, Mr, D, &d2, st));
  } else {
    CK(opnd(dout, &doutop));
  }
  CK(lin_grads(doutop, c.h, dropping ? doutop : dout, LS(layer, NLV_L_LIN2_W), LS(layer, NLV_L_LIN2_B)));
  T dh = tmp(Mr, DFF, AD);
  CK(mm(doutop, K_, W(LS(layer, NLV_L_LIN2_W), D, DFF), MN_, dh, nullptr, nullptr, false, &c.h, false, nullptr, dropping ? d1.scale : 1.f));
  CK(lin_grads(dh, c.top, dh, LS(layer, NLV_L_LIN1_W), LS(layer, NLV_L_LIN1_B)));
  T dt = tmp(Mr, D, NLV_F32);
  CK(mm(dh, K_, W(LS(layer, NLV_L_LIN1_W), DFF, D), MN_, dt, nullptr, &dout));
  T dy = tmp(Mr, D, NLV_F32), dyop = b16 ? tmp(Mr, D, NLV_BF16) : T();
  OOM_CHECK();
  g_next_units = (double)Mr * D * (4 + 4 + 4 + (b16 ? 2 : 0));
  RUN(nlv_layernorm_bwd_fused(dt.f(), c.y.f(), c.m3.f(), c.r3.f(), P(LS(layer, NLV_L_NORMA_W)), Mr, D, dy.f(), b16 ? dyop.p : nullptr, NLV_BF16,
                              G(LS(layer, NLV_L_NORMA_W)), G(LS(layer, NLV_L_NORMA_B)), G(LS(layer, NLV_L_OUTPROJ_B)), &d0, st));
  const T& dyo = b16 ? dyop : dy;
  CK(lin_grads(dyo, c.o, dy, LS(layer, NLV_L_OUTPROJ_W), -1));
  T d_o = tmp(Mr, D, AD);
  CK(mm(dyo, K_, W(LS(layer, NLV_L_OUTPROJ_W), D, D), MN_, d_o));
  T dqkv = tmp(Mr, 3 * D, AD), delta = tmp(Mr * HEADS, 1, NLV_F32);
  OOM_CHECK();
  const T& q = c.qkv;
  const size_t e = q.esz();
  g_next_units = 8.0 * (double)Mr * D * q.esz(); g_next_dt = q.dt;   // Q, K, V, O, dO in; dQ, dK, dV out
  RUN(nlv_attn_bwd_sorted(q.p, q.ld, (char*)q.p + D * e, q.ld, (char*)q.p + 2 * D * e, q.ld, q.dt, HD, HEADS, 1.0f / sqrtf((float)HD), B.glob_work,
                        B.n_glob_work, B.work_sorted ? B.n_glob_long : -1, c.o.p, c.o.ld, c.o.dt, d_o.p, d_o.ld, d_o.dt, c.lse.f(), delta.f(), dqkv.p, dqkv.ld, (char*)dqkv.p + D * e,
                        dqkv.ld, (char*)dqkv.p + 2 * D * e, dqkv.ld, dqkv.dt, &d3, st));
  const T gw = mk(G(LS(layer, NLV_L_INPROJ_W)), NLV_F32, 3 * D, D);
  CK(mm(dqkv.cs(0, 2 * D), MN_, c.xpop, MN_, gw.rs(0, 2 * D)));
  CK(mm(dqkv.cs(2 * D, D), MN_, c.xop, MN_, gw.rs(2 * D, D)));
  // per-slot column sums of dqkv: their sum is the in_proj bias gradient, their q|k part gives the position-embedding gradient
  T ssum = tmp(2, 3 * D, NLV_F32);
  OOM_CHECK();
  RUN(nlv_zero_bytes(ssum.p, 2ll * 3 * D * 4, st));
  RUN(nlv_colsum(dqkv.p, dqkv.dt, dqkv.ld, Mr, 3 * D, B.stream_slot, 2, ssum.f(), st));
  RUN(nlv_add(ssum.f(), ssum.f() + 3 * D, 3 * D, G(LS(layer, NLV_L_INPROJ_B)), st));
  const T gpos = mk(G(NLV_P_POS), NLV_F32, 2, D);
  CK(mm(ssum.cs(0, 2 * D), K_, W(LS(layer, NLV_L_INPROJ_W), 3 * D, D).rs(0, 2 * D), MN_, gpos, nullptr, &gpos));
  CK(mm(dqkv, K_, W(LS(layer, NLV_L_INPROJ_W), 3 * D, D), MN_, dx, nullptr, &dy));
  *dx_out = dx;
  return NLV_OK;
}

// ================================================================================================================
// object classifier + pair tokens
// ================================================================================================================
// Training: per-video batch statistics, running stats updated in video order.  slot_w: BN weight slot (b, rm, rv follow).
int nlv_session::bn_fwd(const T& x, const int* seg, const int* row_seg, int row_div, int slot_w, float momentum, bool relu, const T& y, T* mean,
                        T* var) {
  const int c = x.cols;
  const float *w = P(slot_w), *b = P(slot_w + 1);
  float *rm = const_cast<float*>(P(slot_w + 2)), *rv = const_cast<float*>(P(slot_w + 3));
  if (training) {
    T mu = ctx(B.nv, c, NLV_F32), va = ctx(B.nv, c, NLV_F32);
    T ws = tmp((long long)B.nv * 2 * c, 1, NLV_F32, 2);   // double[nv*2*c]
    OOM_CHECK();
    g_next_units = (double)x.rows * c * x.esz();                    // algorithmic bytes: x read once
    RUN(nlv_bn_stats(x.p, x.dt, x.ld, seg, B.nv, x.rows, c, momentum, reinterpret_cast<double*>(ws.p), mu.f(), va.f(), rm, rv, st));
    g_next_units = (double)x.rows * c * (x.esz() + y.esz());        // x in, y out
    RUN(nlv_bn_apply(x.p, x.dt, x.ld, B.nv > 1 ? row_seg : nullptr, row_div, mu.f(), va.f(), w, b, 1e-5f, relu ? 1 : 0, x.rows, c, y.p, y.dt, y.ld,
                     nullptr, 0, c, st));
    *mean = mu; *var = va;
  } else {
    OOM_CHECK();
    RUN(nlv_bn_apply(x.p, x.dt, x.ld, nullptr, 1, rm, rv, w, b, 1e-5f, relu ? 1 : 0, x.rows, c, y.p, y.dt, y.ld, nullptr, 0, c, st));
    *mean = mk(rm, NLV_F32, 1, c); *var = mk(rv, NLV_F32, 1, c);
  }
  return NLV_OK;
}

// BatchNorm backward (+ the ReLU that follows (yout) or precedes (gate_by_x) it).  Training: per-video statistics.
// Eval (running statistics): one segment over all rows, dx = w * rstd * dy.
int nlv_session::bn_bwd(const T& dy, const T& x, const T* yout, const int* seg, const int* row_seg, int row_div, int slot_w, const T& mean, const T& var,
                        bool gate_by_x, const T& dx, float* dx_colsum) {
  const int c = x.cols;
  int nseg = B.nv;
  if (!training) {
    T one = tmp(1, 4, NLV_F32);
    OOM_CHECK();
    RUN(set_seg(reinterpret_cast<int*>(one.p), (int)x.rows, (cudaStream_t)st));
    seg = reinterpret_cast<const int*>(one.p);
    row_seg = nullptr;
    nseg = 1;
  } else if (B.nv == 1) {
    row_seg = nullptr;
  }
  T ws = tmp((long long)nseg * 2 * c, 1, NLV_F32, 2);   // double[nseg*2*c]
  OOM_CHECK();
  // algorithmic bytes: dy, x (and the ReLU gate) read for the sums and again for dx (the sums are a full reduction), dx out
  g_next_units = (double)x.rows * c * (2.0 * (dy.esz() + x.esz() + (yout ? yout->esz() : 0)) + dx.esz());
  RUN(nlv_bn_bwd_colsum(dy.p, dy.dt, dy.ld, x.p, x.dt, x.ld, yout ? yout->p : nullptr, yout ? yout->dt : 0, yout ? yout->ld : 0, seg, row_seg, row_div,
                        nseg, mean.f(), var.f(), P(slot_w), 1e-5f, training ? 1 : 0, gate_by_x ? 1 : 0, x.rows, c, reinterpret_cast<double*>(ws.p), dx.p,
                        dx.dt, dx.ld, G(slot_w), G(slot_w + 1), dx_colsum, st));
  return NLV_OK;
}

// sgdet / is_wks branch of lib/sttran.py:173-184 (the sgcls training branch :95-104 is the same computation)
int nlv_session::object_classifier_fwd() {
  const bool b16 = AD == NLV_BF16;
  obj_logits = keep(N, 37, NLV_F32);
  T objfeat = keep(N, 2376, AD);     // its first 2048 columns are the feature operand of the pair stage
  oc.objfeat = objfeat;
  Scope sc(this);
  T cs = ctx(N, 4, NLV_F32), pos_bn = ctx(N, 4, NLV_F32), h1 = ctx(N, 1024, NLV_F32), h2 = ctx(N, 1024, NLV_F32);
  OOM_CHECK();
  RUN(nlv_convert(B.features, B.feat_dtype, 2048, objfeat.p, objfeat.dt, objfeat.ld, N, 2048, st));
  CK(mm(mk(B.distribution, NLV_F32, N, 36), K_, mk(P(NLV_P_OC_EMBED), NLV_F32, 36, 200), MN_, objfeat.cs(2048, 200), nullptr, nullptr, false,
        nullptr, true));
  RUN(nlv_center_size(B.boxes, N, cs.f(), st));
  CK(bn_fwd(cs, B.box_seg, B.box_row, 1, NLV_P_OC_BN0_W, 0.01f / 10.0f, false, pos_bn, &oc.mean0, &oc.var0));
  CK(mm(pos_bn, K_, mk(P(NLV_P_OC_LIN1_W), NLV_F32, 128, 4), K_, objfeat.cs(2248, 128), P(NLV_P_OC_LIN1_B), nullptr, true, nullptr, true));
  if (dropping) {     // nn.Dropout(0.1) that closes pos_embed (lib/sttran.py:46)
    const nlv_dropout dp = site(1000, 0);
    const T pe = objfeat.cs(2248, 128);
    RUN(nlv_dropout_apply(pe.p, pe.dt, pe.ld, pe.p, pe.dt, pe.ld, N, 128, &dp, st));
  }
  CK(mm(objfeat, K_, W(NLV_P_OC_DEC0_W, 1024, 2376), K_, h1, P(NLV_P_OC_DEC0_B)));
  CK(bn_fwd(h1, B.box_seg, B.box_row, 1, NLV_P_OC_BN1_W, 0.1f, true, h2, &oc.mean1, &oc.var1));
  CK(mm(h2, K_, mk(P(NLV_P_OC_DEC3_W), NLV_F32, 37, 1024), K_, obj_logits, P(NLV_P_OC_DEC3_B), nullptr, false, nullptr, true));
  oc.cs = cs; oc.pos_bn = pos_bn; oc.h1 = h1; oc.h2 = h2;
  (void)b16;
  return NLV_OK;
}

int nlv_session::object_classifier_bwd(const T& dlogits) {
  Scope sc(this);
  CK(mm(dlogits, MN_, oc.h2, MN_, mk(G(NLV_P_OC_DEC3_W), NLV_F32, 37, 1024), nullptr, nullptr, false, nullptr, true));
  RUN(nlv_colsum(dlogits.p, dlogits.dt, dlogits.ld, N, 37, nullptr, 1, G(NLV_P_OC_DEC3_B), st));
  T dh2 = tmp(N, 1024, NLV_F32);
  CK(mm(dlogits, K_, mk(P(NLV_P_OC_DEC3_W), NLV_F32, 37, 1024), MN_, dh2, nullptr, nullptr, false, nullptr, true));
  T dh1 = tmp(N, 1024, AD);
  CK(bn_bwd(dh2, oc.h1, &oc.h2, B.box_seg, B.box_row, 1, NLV_P_OC_BN1_W, oc.mean1, oc.var1, false, dh1));
  CK(lin_grads(dh1, oc.objfeat, dh1, NLV_P_OC_DEC0_W, NLV_P_OC_DEC0_B));
  const T w0 = W(NLV_P_OC_DEC0_W, 1024, 2376);
  T dtail = tmp(N, 328, NLV_F32);                                   // [N, 200 + 128]
  CK(mm(dh1, K_, w0.cs(2048, 328), MN_, dtail));
  CK(mm(mk(B.distribution, NLV_F32, N, 36), MN_, dtail.cs(0, 200), MN_, mk(G(NLV_P_OC_EMBED), NLV_F32, 36, 200), nullptr, nullptr, false,
        nullptr, true));
  T dpos = tmp(N, 128, NLV_F32);
  OOM_CHECK();
  const T gatev = oc.objfeat.cs(2248, 128);
  const T dt128 = dtail.cs(200, 128);
  RUN(nlv_relu_mask(dt128.p, dt128.dt, dt128.ld, gatev.p, gatev.dt, gatev.ld, N, 128, dpos.p, dpos.dt, dpos.ld, st));
  if (dropping) {     // kept entries (the only non-zero ones after the gate) pick up the 1 / (1 - p) of the forward dropout
    const nlv_dropout dp = site(1000, 0);
    RUN(nlv_dropout_apply(dpos.p, dpos.dt, dpos.ld, dpos.p, dpos.dt, dpos.ld, N, 128, &dp, st));
  }
  CK(mm(dpos, MN_, oc.pos_bn, MN_, mk(G(NLV_P_OC_LIN1_W), NLV_F32, 128, 4), nullptr, nullptr, false, nullptr, true));
  RUN(nlv_colsum(dpos.p, dpos.dt, dpos.ld, N, 128, nullptr, 1, G(NLV_P_OC_LIN1_B), st));
  T dposbn = tmp(N, 4, NLV_F32), dcs = tmp(N, 4, NLV_F32);
  CK(mm(dpos, K_, mk(P(NLV_P_OC_LIN1_W), NLV_F32, 128, 4), MN_, dposbn, nullptr, nullptr, false, nullptr, true));
  CK(bn_bwd(dposbn, oc.cs, nullptr, B.box_seg, B.box_row, 1, NLV_P_OC_BN0_W, oc.mean0, oc.var0, false, dcs));
  return NLV_OK;
}

// 1936-d relation tokens (lib/sttran.py:381-399): [subj_fc | obj_fc | vr_fc(union 1x1 conv + mask conv stack) | emb | emb2]
int nlv_session::pair_tokens_fwd(const T& feat_op) {
  rel = keep(R, D, NLV_F32);
  pt.feat_op = feat_op;
  Scope sc(this);
  T fo = tmp(N, 1024, NLV_F32);
  CK(mm(feat_op, K_, W(NLV_P_SUBJ_W, 512, 2048), K_, fo.cs(0, 512), P(NLV_P_SUBJ_B)));
  CK(mm(feat_op, K_, W(NLV_P_OBJ_W, 512, 2048), K_, fo.cs(512, 512), P(NLV_P_OBJ_B)));
  // union features as [R*49, 2048] rows (operand of the 1x1 conv)
  T uf;
  if (B.union_rows == 2 || B.union_rows == 3) {          // zero-suppressed rows from a packed feature file
    NLV_CHECK_ARG(B.union_bitmap != nullptr && B.union_off != nullptr, "session: sparse union features need bitmap and offsets");
    NLV_CHECK_ARG(B.union_rows == 2 || (B.union_hx != nullptr && B.union_base != nullptr), "session: 12-bit union features need codes and row bases");
    T d = AD == NLV_BF16 ? ctx(R * 49, 2048, NLV_BF16) : tmp(R * 49, 2048, NLV_BF16);
    OOM_CHECK();
    g_next_units = (double)R * 49 * (256 + 4 + 4096);   // + the bytes of the stored values (added by the reader of the profile)
    if (B.union_rows == 3) {
      RUN(nlv_union_unpack12(B.union_bitmap, B.union_off, B.union_feat, B.union_hx, B.union_base, R * 49, d.p, st));
      if (B.n_union_exc > 0) RUN(nlv_union_patch(d.p, B.union_exc_pos, B.union_exc_val, B.n_union_exc, st));
    } else {
      RUN(nlv_union_unpack(B.union_bitmap, B.union_off, B.union_feat, R * 49, d.p, st));
    }
    uf = d;
    if (AD != NLV_BF16) {
      T t = ctx(R * 49, 2048, AD);
      OOM_CHECK();
      RUN(nlv_convert(d.p, d.dt, d.ld, t.p, t.dt, t.ld, R * 49, 2048, st));
      uf = t;
    }
  } else if (B.union_rows) {
    uf = mk(B.union_feat, B.union_dtype, R * 49, 2048);
    if (uf.dt != AD) {
      T t = ctx(R * 49, 2048, AD);
      OOM_CHECK();
      RUN(nlv_convert(uf.p, uf.dt, uf.ld, t.p, t.dt, t.ld, R * 49, 2048, st));
      uf = t;
    }
  } else {
    uf = ctx(R * 49, 2048, AD);
    OOM_CHECK();
    const size_t in_row = (size_t)2048 * 49 * (B.union_dtype == NLV_BF16 ? 2 : 4);
    for (long long s = 0; s < R; s += 32768) {   // grid.y limit of the transposing kernel
      const int n = (int)((R - s) < 32768 ? (R - s) : 32768);
      g_next_units = (double)n * 49 * 2048 * ((B.union_dtype == NLV_BF16 ? 2 : 4) + uf.esz());
      RUN(nlv_nchw_to_rows((const char*)B.union_feat + s * in_row, B.union_dtype, n, 2048, 49, (char*)uf.p + (size_t)s * 49 * 2048 * uf.esz(),
                           uf.dt, st));
    }
  }
  T col1, c1 = ctx(R * 196, 128, AD), xmax_keep;
  T p1 = tmp(R * 49, 128, AD), arg = ctx(R * 49, 128, NLV_BF16 /*u8 payload*/, 64);
  arg.dt = NLV_BF16;
  if (fused_mask_conv()) {
    // conv 7x7 s2 + ReLU (+ BatchNorm statistics) straight from the masks, then BatchNorm + pooling in one pass (maskconv.cu)
    T ws = tmp((long long)B.nv * 2 * 128, 1, NLV_F32, 2);   // double[nv*2*128]
    T mu = training ? ctx(B.nv, 128, NLV_F32) : T(), va = training ? ctx(B.nv, 128, NLV_F32) : T();
    pool_fused = want_ctx && training;                       // the backward then runs nlv_pool_bn_bwd (batch statistics only)
    T xmax = pool_fused ? ctx(R * 49, 128, AD) : T();        // the activation at every argmax: input of its reduction pass
    OOM_CHECK();
    float *rm = const_cast<float*>(P(NLV_P_BN2_W + 2)), *rv = const_cast<float*>(P(NLV_P_BN2_W + 3));
    g_next_units = (double)R * (2 * 27 * 27 * 4 + 196 * 128 * 2);
    RUN(nlv_mask_conv1_fwd(masks.f(), P(NLV_P_CONV0_W), P(NLV_P_CONV0_B), R, B.pair_row, B.seg196, B.nv, c1.p, 0.01f,
                           reinterpret_cast<double*>(ws.p), training ? mu.f() : nullptr, training ? va.f() : nullptr, rm, rv, st));
    g_next_units = (double)R * (196 * 128 * 2 + 49 * 128 * (pool_fused ? 5 : 3));
    if (training) {
      RUN(nlv_bn_apply_maxpool(c1.p, B.nv > 1 ? B.pair_row : nullptr, mu.f(), va.f(), P(NLV_P_BN2_W), P(NLV_P_BN2_W + 1), 1e-5f, R, p1.p,
                               reinterpret_cast<uint8_t*>(arg.p), xmax.p, st));
      pt.mean2 = mu; pt.var2 = va;
    } else {
      RUN(nlv_bn_apply_maxpool(c1.p, nullptr, rm, rv, P(NLV_P_BN2_W), P(NLV_P_BN2_W + 1), 1e-5f, R, p1.p, reinterpret_cast<uint8_t*>(arg.p), xmax.p,
                               st));
      pt.mean2 = mk(rm, NLV_F32, 1, 128); pt.var2 = mk(rv, NLV_F32, 1, 128);
    }
    xmax_keep = xmax;
  } else {
    pool_fused = false;
    col1 = ctx(R * 196, 104, AD);
    T b1 = tmp(R * 196, 128, AD);
    OOM_CHECK();
    RUN(nlv_im2col_mask(masks.f(), (int)R, col1.p, col1.dt, 104, st));
    CK(mm(col1, K_, w_c0, K_, c1, P(NLV_P_CONV0_B), nullptr, true));                                     // conv 7x7 s2 + ReLU
    CK(bn_fwd(c1, B.seg196, B.pair_row, 196, NLV_P_BN2_W, 0.01f, false, b1, &pt.mean2, &pt.var2));
    OOM_CHECK();
    RUN(nlv_maxpool_fwd(b1.p, b1.dt, (int)R, 128, p1.p, p1.dt, reinterpret_cast<uint8_t*>(arg.p), st));
  }
  T col2 = ctx(R * 49, 1152, AD), c2 = ctx(R * 49, 256, AD), b2 = tmp(R * 49, 256, AD);
  OOM_CHECK();
  RUN(nlv_im2col_3x3(p1.p, p1.dt, (int)R, 7, 7, 128, col2.p, col2.dt, st));
  CK(mm(col2, K_, w_c4, K_, c2, P(NLV_P_CONV4_B), nullptr, true));
  CK(bn_fwd(c2, B.seg49, B.pair_row, 49, NLV_P_BN6_W, 0.01f, false, b2, &pt.mean6, &pt.var6));
  T vr_in = ctx(R * 49, 256, AD);                                                                      // = [R, 12544] in (hw, c) order
  OOM_CHECK();
  CK(mm(uf, K_, W(NLV_P_UNION_W, 256, 2048), K_, vr_in, P(NLV_P_UNION_B), &b2));
  CK(mm(vr_in.view(R, 12544), K_, w_vr, K_, rel.cs(1024, 512), P(NLV_P_VR_B)));
  RUN(nlv_assemble_tokens(fo.f(), B.pair_idx, B.labels, P(NLV_P_EMB1), P(NLV_P_EMB2), R, rel.f(), st));
  pt.uf_op = uf; pt.col1 = col1; pt.c1 = c1; pt.arg = arg; pt.xmax = xmax_keep; pt.col2 = col2; pt.c2 = c2; pt.vr_in = vr_in;
  return NLV_OK;
}

int nlv_session::pair_tokens_bwd(const T& drel) {
  Scope sc(this);
  const bool b16 = AD == NLV_BF16;
  T dfo = tmp(N, 1024, NLV_F32);
  OOM_CHECK();
  RUN(nlv_zero_bytes(dfo.p, (long long)N * 1024 * 4, st));
  RUN(nlv_assemble_tokens_bwd(drel.f(), B.pair_idx, B.labels, R, dfo.f(), G(NLV_P_EMB1), G(NLV_P_EMB2), st));
  T dvr;
  CK(opnd(drel.cs(1024, 512), &dvr));
  const T vr2d = pt.vr_in.view(R, 12544);
  {
    Scope s2(this);
    T gperm = tmp(512, 12544, NLV_F32);
    CK(mm(dvr, MN_, vr2d, MN_, gperm));
    OOM_CHECK();
    RUN(nlv_permute_021(gperm.p, NLV_F32, 512, 49, 256, G(NLV_P_VR_W), NLV_F32, st));   // (hw, c) -> (c, hw)
  }
  const T dvrb = drel.cs(1024, 512);
  RUN(nlv_colsum(dvrb.p, dvrb.dt, dvrb.ld, R, 512, nullptr, 1, G(NLV_P_VR_B), st));
  T dvr_in2d = tmp(R, 12544, AD);   // activation dtype: feeds a GEMM, a column sum and BN backward
  CK(mm(dvr, K_, w_vr, MN_, dvr_in2d));
  const T dvr_in = dvr_in2d.view(R * 49, 256);
  CK(mm(dvr_in, MN_, pt.uf_op, MN_, mk(G(NLV_P_UNION_W), NLV_F32, 256, 2048)));
  T dc2 = tmp(R * 49, 256, AD);
  // BN backward + ReLU backward fused; the kernel also accumulates the column sums of dc2 = the bias gradient of the 3x3 conv
  CK(bn_bwd(dvr_in, pt.c2, nullptr, B.seg49, B.pair_row, 49, NLV_P_BN6_W, pt.mean6, pt.var6, true, dc2, G(NLV_P_CONV4_B)));
  // vr = union_func1(x) + BatchNorm(conv(masks)): the two bias gradients are the same column sums of dvr_in, and BatchNorm's
  // backward has just produced them (one 256-float add instead of another pass over the 300 MB map)
  RUN(nlv_add(G(NLV_P_UNION_B), G(NLV_P_BN6_W + 1), 256, G(NLV_P_UNION_B), st));
  {
    Scope s2(this);
    T gtap = tmp(256, 1152, NLV_F32);
    CK(mm(dc2, MN_, pt.col2, MN_, gtap));
    OOM_CHECK();
    RUN(nlv_permute_021(gtap.p, NLV_F32, 256, 9, 128, G(NLV_P_CONV4_W), NLV_F32, st));  // (tap, c) -> (c, tap)
  }
  T dp1 = tmp(R * 49, 128, NLV_F32);
  if (implicit_dgrad()) {
    // data gradient of the 3x3 conv as one implicit GEMM over the nine shifted views of dc2 (4-D TMA boxes, zero halo)
    OOM_CHECK();
    g_next_flops = 2.0 * (double)R * 49 * 128 * 2304; g_next_m = (int)(R * 49); g_next_n = 128; g_next_k = 2304; g_next_dt = 1;
    RUN(nlv_conv3x3_dgrad(dc2.p, R, 256, w_c4t.p, 128, dp1.p, dp1.dt, st));
  } else {
    Scope s3(this);
    T dcol2 = tmp(R * 49, 1152, AD);
    CK(mm(dc2, K_, w_c4, MN_, dcol2));
    OOM_CHECK();
    RUN(nlv_col2im_3x3(dcol2.p, dcol2.dt, (int)R, 7, 7, 128, dp1.f(), st));
  }
  T dc1 = tmp(R * 196, 128, AD);
  if (fused_mask_conv() && pool_fused) {
    // pooling + BatchNorm + ReLU backward in one sweep over the pooled gradient (maskconv.cu); + the 7x7 conv bias gradient
    T ws = tmp((long long)B.nv * 2 * 128, 1, NLV_F32, 2);
    OOM_CHECK();
    g_next_units = (double)R * (49 * 128 * (4 + 2) + 49 * 128 * (4 + 1) + 196 * 128 * (2 + 2));
    RUN(nlv_pool_bn_bwd(dp1.f(), reinterpret_cast<const uint8_t*>(pt.arg.p), pt.c1.p, pt.xmax.p, B.nv > 1 ? B.pair_row : nullptr, B.seg196, B.seg49,
                        B.nv, pt.mean2.f(), pt.var2.f(), P(NLV_P_BN2_W), 1e-5f, training ? 1 : 0, R, reinterpret_cast<double*>(ws.p), dc1.p,
                        G(NLV_P_BN2_W), G(NLV_P_BN2_W + 1), G(NLV_P_CONV0_B), st));
  } else {
    T db1 = tmp(R * 196, 128, AD);
    OOM_CHECK();
    RUN(nlv_maxpool_bwd(dp1.f(), reinterpret_cast<const uint8_t*>(pt.arg.p), (int)R, 128, db1.p, db1.dt, st));
    CK(bn_bwd(db1, pt.c1, nullptr, B.seg196, B.pair_row, 196, NLV_P_BN2_W, pt.mean2, pt.var2, true, dc1, G(NLV_P_CONV0_B)));   // + 7x7 conv bias gradient
  }
  if (fused_mask_conv()) {
    Scope s2(this);
    T ws = tmp(nlv_mask_conv1_dw_ws_floats(), 1, NLV_F32);
    OOM_CHECK();
    g_next_units = (double)R * (2 * 27 * 27 * 4 + 196 * 128 * 2);
    RUN(nlv_mask_conv1_dw(dc1.p, masks.f(), R, ws.f(), G(NLV_P_CONV0_W), st));
  } else {
    Scope s2(this);
    T g0 = tmp(128, 104, NLV_F32);
    CK(mm(dc1, MN_, pt.col1, MN_, g0));
    OOM_CHECK();
    RUN(nlv_convert(g0.p, NLV_F32, 104, G(NLV_P_CONV0_W), NLV_F32, 98, 128, 98, st));
  }
  T dfo_op;
  CK(opnd(dfo, &dfo_op));
  CK(mm(dfo_op.cs(0, 512), MN_, pt.feat_op, MN_, mk(G(NLV_P_SUBJ_W), NLV_F32, 512, 2048)));
  CK(mm(dfo_op.cs(512, 512), MN_, pt.feat_op, MN_, mk(G(NLV_P_OBJ_W), NLV_F32, 512, 2048)));
  RUN(nlv_colsum(dfo.p, dfo.dt, dfo.ld, N, 512, nullptr, 1, G(NLV_P_SUBJ_B), st));
  const T dfo2 = dfo.cs(512, 512);
  RUN(nlv_colsum(dfo2.p, dfo2.dt, dfo2.ld, N, 512, nullptr, 1, G(NLV_P_OBJ_B), st));
  (void)b16;
  return NLV_OK;
}

// ================================================================================================================
// spatio-temporal transformers
// ================================================================================================================
// transformer_wk.forward (mode='latter') on the concatenated batch
int nlv_session::sttran_transformer_fwd(const T& rel_in, T* out_) {
  const bool b16 = AD == NLV_BF16;
  T x = rel_in, xop;
  if (b16) { xop = ctx(x.rows, D, NLV_BF16); OOM_CHECK(); RUN(nlv_convert(x.p, x.dt, x.ld, xop.p, xop.dt, xop.ld, x.rows, D, st)); }
  else xop = x;
  enc.assign(M.n_enc, EncCtx());
  dec.assign(M.n_dec, DecCtx());
  for (int i = 0; i < M.n_enc; ++i) {
    T x2, x2op;
    CK(encoder_fwd(i, x, xop, B.local_work, B.n_local_work, false, &x2, &x2op, want_ctx ? &enc[i] : nullptr));
    x = x2;
    if (i + 1 < M.n_enc) CK(opnd(x2, &xop));   // (only with more than one spatial layer)
  }
  local_out = x;
  if (Mg == 0) { *out_ = local_out; return NLV_OK; }
  const float* pe = P(NLV_P_POS);
  // window stream: g = local_out[stream_src]; operand copies of g and g + pos[slot]
  T g = keep(Mg, D, NLV_F32);
  for (int i = 0; i < M.n_dec; ++i) {
    Scope sc(this);
    T gop = b16 ? ctx(Mg, D, NLV_BF16) : T(), gpop = ctx(Mg, D, AD);
    OOM_CHECK();
    if (i == 0) {
      RUN(nlv_gather_rows(local_out.p, NLV_F32, local_out.ld, B.stream_src, nullptr, nullptr, 0, Mg, D, g.p, NLV_F32, D, b16 ? gop.p : nullptr,
                          NLV_BF16, D, st));
      RUN(nlv_gather_rows(local_out.p, NLV_F32, local_out.ld, B.stream_src, pe, B.stream_slot, D, Mg, D, nullptr, 0, D, gpop.p, gpop.dt, D, st));
    } else {
      if (b16) RUN(nlv_convert(g.p, NLV_F32, D, gop.p, NLV_BF16, D, Mg, D, st));
      RUN(nlv_gather_rows(g.p, NLV_F32, D, nullptr, pe, B.stream_slot, D, Mg, D, nullptr, 0, D, gpop.p, gpop.dt, D, st));
    }
    T gn;
    CK(decoder_fwd(M.n_enc + i, g, b16 ? gop : g, gpop, &gn, want_ctx ? &dec[i] : nullptr));
    g = gn;
  }
  T out = keep(R, D, NLV_F32);
  OOM_CHECK();
  if (M.transformer_both) {      // mean over the windows a token appears in (lib/transformer_wk.py:197-207)
    NLV_CHECK_ARG(dry || B.both_w != nullptr, "session: transformer mode 'both' needs the per-token window weights");
    RUN(nlv_gather_sum_rows(g.f(), D, B.inv, 2, R, D, out.f(), D, 0, st));
    RUN(nlv_scale_rows(out.f(), D, B.both_w, R, D, out.f(), D, st));
  } else {
    RUN(nlv_gather_rows(g.p, NLV_F32, D, B.out_src, nullptr, nullptr, 0, R, D, out.p, NLV_F32, D, nullptr, 0, D, st));
  }
  if (B.has_passthrough)
    RUN(nlv_gather_sum_rows(local_out.f(), D, B.passthrough, 1, R, D, out.f(), D, 1, st));
  *out_ = out;
  return NLV_OK;
}

int nlv_session::sttran_transformer_bwd(const T& dout, T* drel) {
  T dlocal = dout;
  if (Mg != 0) {
    T dg = keep(Mg, D, NLV_F32);
    OOM_CHECK();
    if (M.transformer_both) {    // every stream row takes its token's gradient, weighted by 1 / (windows of the token)
      T dw = tmp(R, D, NLV_F32);
      OOM_CHECK();
      RUN(nlv_scale_rows(dout.f(), dout.ld, B.both_w, R, D, dw.f(), D, st));
      RUN(nlv_gather_rows(dw.p, NLV_F32, D, B.stream_src, nullptr, nullptr, 0, Mg, D, dg.p, NLV_F32, D, nullptr, 0, D, st));
    } else {
      RUN(nlv_gather_rows(dout.p, NLV_F32, dout.ld, B.out_inv, nullptr, nullptr, 0, Mg, D, dg.p, NLV_F32, D, nullptr, 0, D, st));
    }
    for (int i = M.n_dec - 1; i >= 0; --i) {
      T dn;
      CK(decoder_bwd(M.n_enc + i, dec[i], dg, &dn));
      dg = dn;
    }
    if (!dry && tail_hook != nullptr) tail_hook(tail_hook_user);
    dlocal = keep(R, D, NLV_F32);
    OOM_CHECK();
    RUN(nlv_gather_sum_rows(dg.f(), D, B.inv, 2, R, D, dlocal.f(), D, 0, st));
    if (B.has_passthrough) RUN(nlv_gather_sum_rows(dout.f(), dout.ld, B.passthrough, 1, R, D, dlocal.f(), D, 1, st));
  }
  for (int i = M.n_enc - 1; i >= 0; --i) {
    T dn;
    CK(encoder_bwd(i, enc[i], dlocal, B.local_work, B.n_local_work, true, &dn));
    dlocal = dn;
  }
  *drel = dlocal;
  return NLV_OK;
}

// DSG-DETR (lib/dsg_detr.py:536-564): spatial encoder over frames, then 3 encoder layers over per-class sequences of the
// class-sorted token stream with the sinusoidal encoding of the frame rank added
int nlv_session::dsg_transformer_fwd(const T& rel_in, T* out_) {
  const bool b16 = AD == NLV_BF16;
  enc.assign(4, EncCtx());
  T xop;
  if (b16) { xop = ctx(R, D, NLV_BF16); OOM_CHECK(); RUN(nlv_convert(rel_in.p, rel_in.dt, rel_in.ld, xop.p, xop.dt, xop.ld, R, D, st)); }
  else xop = rel_in;
  T x, xo;
  CK(encoder_fwd(0, rel_in, xop, B.local_work, B.n_local_work, false, &x, &xo, want_ctx ? &enc[0] : nullptr));
  T g = keep(R, D, NLV_F32), gop = b16 ? ctx(R, D, NLV_BF16) : T();
  OOM_CHECK();
  RUN(nlv_gather_rows(x.p, NLV_F32, D, B.cls_perm, P(NLV_P_POS), B.cls_pos, D, R, D, g.p, NLV_F32, D, (b16 && !dropping) ? gop.p : nullptr,
                      NLV_BF16, D, st));
  if (dropping) {     // PositionalEncoding.dropout (lib/dsg_detr.py:28,48) on x + pe
    const nlv_dropout dp = site(1001, 0);
    RUN(nlv_dropout_apply(g.p, g.dt, g.ld, g.p, g.dt, g.ld, R, D, &dp, st));
    if (b16) RUN(nlv_convert(g.p, g.dt, g.ld, gop.p, gop.dt, gop.ld, R, D, st));
  }
  if (!b16) gop = g;
  for (int i = 0; i < 3; ++i) {
    T g2, g2op;
    CK(encoder_fwd(1 + i, g, gop, B.cls_work, B.n_cls_work, i < 2, &g2, &g2op, want_ctx ? &enc[1 + i] : nullptr));
    g = g2; gop = g2op;
  }
  T out = keep(R, D, NLV_F32);
  OOM_CHECK();
  RUN(nlv_gather_rows(g.p, NLV_F32, D, B.cls_iperm, nullptr, nullptr, 0, R, D, out.p, NLV_F32, D, nullptr, 0, D, st));
  *out_ = out;
  return NLV_OK;
}

int nlv_session::dsg_transformer_bwd(const T& dout, T* drel) {
  T dg = keep(R, D, NLV_F32);
  OOM_CHECK();
  RUN(nlv_gather_rows(dout.p, NLV_F32, dout.ld, B.cls_perm, nullptr, nullptr, 0, R, D, dg.p, NLV_F32, D, nullptr, 0, D, st));
  for (int i = 2; i >= 0; --i) {
    T dn;
    CK(encoder_bwd(1 + i, enc[1 + i], dg, B.cls_work, B.n_cls_work, true, &dn));
    dg = dn;
  }
  if (!dry && tail_hook != nullptr) tail_hook(tail_hook_user);
  T dx = keep(R, D, NLV_F32);   // the encoding is a constant buffer
  OOM_CHECK();
  if (dropping) {
    const nlv_dropout dp = site(1001, 0);
    RUN(nlv_dropout_apply(dg.p, dg.dt, dg.ld, dg.p, dg.dt, dg.ld, R, D, &dp, st));
  }
  RUN(nlv_gather_rows(dg.p, NLV_F32, D, B.cls_iperm, nullptr, nullptr, 0, R, D, dx.p, NLV_F32, D, nullptr, 0, D, st));
  CK(encoder_bwd(0, enc[0], dx, B.local_work, B.n_local_work, true, drel));
  return NLV_OK;
}

// lib/sttran.py:404-406 as one [R,1936] x [26,1936]^T product at fp32 grade -> logits [R,26]
int nlv_session::heads_fwd(const T& x) {
  logits26 = keep(R, 26, NLV_F32);
  Scope sc(this);
  OOM_CHECK();
  CK(mm(x, K_, w26, K_, logits26, b26.f(), nullptr, false, nullptr, true));
  return NLV_OK;
}

int nlv_session::heads_bwd(const T& x, const T& dlogits, T* dx_out) {
  T dx = keep(R, D, NLV_F32);
  Scope sc(this);
  T dw = tmp(26, D, NLV_F32), db = tmp(1, 32, NLV_F32);
  CK(mm(dlogits, MN_, x, MN_, dw, nullptr, nullptr, false, nullptr, true));
  OOM_CHECK();
  RUN(nlv_zero_bytes(db.p, 32 * 4, st));
  RUN(nlv_colsum(dlogits.p, dlogits.dt, dlogits.ld, R, 26, nullptr, 1, db.f(), st));
  const void* src[6] = {dw.f(), dw.f() + 3 * D, dw.f() + 9 * D, db.f(), db.f() + 3, db.f() + 9};
  void* dst[6] = {G(NLV_P_A_W), G(NLV_P_S_W), G(NLV_P_C_W), G(NLV_P_A_B), G(NLV_P_S_B), G(NLV_P_C_B)};
  const long long n[6] = {3ll * D, 6ll * D, 17ll * D, 3, 6, 17};
  RUN(nlv_convert_multi(src, dst, n, 6, NLV_F32, NLV_F32, st));
  CK(mm(dlogits, K_, w26, MN_, dx, nullptr, nullptr, false, nullptr, true));
  *dx_out = dx;
  return NLV_OK;
}

// ================================================================================================================
// per-step operand weights
// ================================================================================================================
int nlv_session::prepare_weights() {
  wop.assign(M.n_slots, T());
  const bool b16 = M.precision == NLV_PREC_BF16;
  const int n_layers = M.arch == NLV_ARCH_DSG ? 4 : M.n_enc + M.n_dec;
  struct Wd { int slot; long long rows; int cols; };
  std::vector<Wd> ws;
  if (!xf_only) {
    if (M.mode != NLV_MODE_PREDCLS) ws.push_back({NLV_P_OC_DEC0_W, 1024, 2376});
    ws.push_back({NLV_P_UNION_W, 256, 2048});
    ws.push_back({NLV_P_SUBJ_W, 512, 2048});
    ws.push_back({NLV_P_OBJ_W, 512, 2048});
  }
  for (int l = 0; l < n_layers; ++l) {
    ws.push_back({LS(l, NLV_L_INPROJ_W), 3 * D, D});
    ws.push_back({LS(l, NLV_L_OUTPROJ_W), D, D});
    ws.push_back({LS(l, NLV_L_LIN1_W), DFF, D});
    ws.push_back({LS(l, NLV_L_LIN2_W), D, DFF});
  }
  if (b16) {
    std::vector<const void*> src;
    std::vector<void*> dst;
    std::vector<long long> cnt;
    for (const Wd& w : ws) {
      if (!pop.empty() && pop[w.slot] != nullptr) { wop[w.slot] = mk(pop[w.slot], NLV_BF16, w.rows, w.cols); continue; }
      T t = keep(w.rows, w.cols, NLV_BF16);
      wop[w.slot] = t;
      src.push_back(params[w.slot]); dst.push_back(t.p); cnt.push_back(w.rows * w.cols);
    }
    OOM_CHECK();
    if (!src.empty()) RUN(nlv_convert_multi(src.data(), dst.data(), cnt.data(), (int)src.size(), NLV_F32, NLV_BF16, st));
  }
  if (xf_only) return NLV_OK;
  // derived operand layouts: conv.0 [128,98] padded to K = 104; conv.4 taps (ky,kx,c) to match the channels-innermost
  // im2col; vr_fc columns (hw,c) to match the NHWC union tensor
  w_c0 = keep(128, 104, AD);
  w_c4 = keep(256, 1152, AD);
  if (implicit_dgrad()) w_c4t = keep(128, 2304, AD);
  w_vr = keep(512, 12544, AD);
  w26 = keep(26, D, NLV_F32);
  b26 = keep(1, 32, NLV_F32);
  OOM_CHECK();
  RUN(nlv_zero_bytes(w_c0.p, 128ll * 104 * w_c0.esz(), st));
  RUN(nlv_convert(P(NLV_P_CONV0_W), NLV_F32, 98, w_c0.p, w_c0.dt, 104, 128, 98, st));
  RUN(nlv_permute_021(P(NLV_P_CONV4_W), NLV_F32, 256, 128, 9, w_c4.p, w_c4.dt, st));
  if (implicit_dgrad()) RUN(nlv_permute_021(P(NLV_P_CONV4_W), NLV_F32, 1, 256, 1152, w_c4t.p, w_c4t.dt, st));   // [co, ci*9] -> [ci][tap][co]
  RUN(nlv_permute_021(P(NLV_P_VR_W), NLV_F32, 512, 256, 49, w_vr.p, w_vr.dt, st));
  const void* src[6] = {P(NLV_P_A_W), P(NLV_P_S_W), P(NLV_P_C_W), P(NLV_P_A_B), P(NLV_P_S_B), P(NLV_P_C_B)};
  void* dst[6] = {w26.f(), w26.f() + 3 * D, w26.f() + 9 * D, b26.f(), b26.f() + 3, b26.f() + 9};
  const long long n[6] = {3ll * D, 6ll * D, 17ll * D, 3, 6, 17};
  RUN(nlv_convert_multi(src, dst, n, 6, NLV_F32, NLV_F32, st));
  return NLV_OK;
}

// ================================================================================================================
// whole model
// ================================================================================================================
int nlv_session::setup(const nlv_model* model, const nlv_batch* batch, int flags_) {
  NLV_CHECK_ARG(model != nullptr && batch != nullptr, "session: null model / batch");
  M = *model;
  B = *batch;
  flags = flags_;
  NLV_CHECK_ARG(M.arch == NLV_ARCH_STTRAN || M.arch == NLV_ARCH_DSG, "session: bad arch %d", M.arch);
  NLV_CHECK_ARG(M.mode >= NLV_MODE_PREDCLS && M.mode <= NLV_MODE_SGDET, "session: bad mode %d", M.mode);
  NLV_CHECK_ARG(M.precision >= NLV_PREC_BF16 && M.precision <= NLV_PREC_FP32, "session: bad precision %d", M.precision);
  const int n_layers = M.arch == NLV_ARCH_DSG ? 4 : M.n_enc + M.n_dec;
  NLV_CHECK_ARG(M.n_slots >= NLV_P_LAYER0 + n_layers * NLV_P_LAYER_STRIDE, "session: parameter table has %d slots, model needs %d", M.n_slots,
                NLV_P_LAYER0 + n_layers * NLV_P_LAYER_STRIDE);
  NLV_CHECK_ARG(M.params != nullptr, "session: null parameter table");
  params.assign(M.params, M.params + M.n_slots);
  if (M.params_op != nullptr) pop.assign(M.params_op, M.params_op + M.n_slots); else pop.clear();
  if (M.grad_offset != nullptr) goff.assign(M.grad_offset, M.grad_offset + M.n_slots); else goff.clear();
  M.params = nullptr; M.params_op = nullptr; M.grad_offset = nullptr;   // the caller's host arrays may not outlive the call
  want_ctx = (flags & NLV_RUN_CTX) != 0;
  training = M.training != 0;
  AD = M.precision == NLV_PREC_BF16 ? NLV_BF16 : NLV_F32;
  dropping = training && M.dropout_p > 0.f;
  NLV_CHECK_ARG(!dropping || M.precision == NLV_PREC_BF16,
                "session: dropout is implemented on the bf16 path (attention-weight masks live in the tensor-core attention kernels); "
                "run the fp32 / bf16x3 parity modes with dropout 0");
  NLV_CHECK_ARG(M.dropout_p >= 0.f && M.dropout_p < 1.f, "session: bad dropout_p");
  NLV_CHECK_ARG(!(M.additive_mask != 0 && (training || (flags & NLV_RUN_CTX))),
                "session: additive_mask (the torch-1.10.1 int key_padding_mask reading) is an inference-only compatibility mode");
  N = B.n_boxes; R = B.n_pairs; Mg = B.n_stream;
  NLV_CHECK_ARG(N >= 0 && R >= 0 && Mg >= 0 && B.nv >= 1, "session: bad batch sizes");
  return NLV_OK;
}

int nlv_session::run_forward() {
  CK(prepare_weights());
  // spatial masks: the caller's, or rasterised here (fused pair gather + draw_union_boxes - 0.5, lib/sttran.py:279-281)
  if (B.spatial_masks != nullptr) masks = mk(B.spatial_masks, NLV_F32, R, 2 * 27 * 27);
  else {
    masks = keep(R, 2 * 27 * 27, NLV_F32);
    OOM_CHECK();
    if (R > 0) RUN(nlv_union_mask_pairs(B.boxes, reinterpret_cast<const int64_t*>(B.pair_idx), (int)R, 27, -0.5f, masks.f(), st));
  }
  T feat_op;
  if (M.mode == NLV_MODE_PREDCLS) {
    obj_logits = T();
    feat_op = mk(B.features, B.feat_dtype, N, 2048);
    if (feat_op.dt != AD) {
      T t = ctx(N, 2048, AD);
      OOM_CHECK();
      RUN(nlv_convert(feat_op.p, feat_op.dt, 2048, t.p, t.dt, t.ld, N, 2048, st));
      feat_op = t;
    }
  } else {
    if (B.distribution == nullptr && B.dist_conf != nullptr && B.dist_idx != nullptr) {   // create_dis on device
      T d = keep(N, 36, NLV_F32);
      OOM_CHECK();
      RUN(nlv_create_dis(B.dist_conf, B.dist_other, B.dist_idx, N, d.f(), st));
      B.distribution = d.f();
    }
    NLV_CHECK_ARG(dry || B.distribution != nullptr, "session: sgdet / sgcls need entry['distribution']");
    CK(object_classifier_fwd());
    feat_op = oc.objfeat.cs(0, 2048);
    if (flags & NLV_RUN_OBJECT_ONLY) return NLV_OK;   // the sgcls test branch builds its pairs from these logits (lib/sttran.py:105-170)
  }
  CK(pair_tokens_fwd(feat_op));
  if (M.arch == NLV_ARCH_STTRAN) CK(sttran_transformer_fwd(rel, &xf_out));
  else CK(dsg_transformer_fwd(rel, &xf_out));
  CK(heads_fwd(xf_out));
  if (flags & NLV_RUN_ACTIVATIONS) {
    att = keep(R, 3, NLV_F32); spa = keep(R, 6, NLV_F32); con = keep(R, 17, NLV_F32);
    OOM_CHECK();
    RUN(nlv_heads_activation(logits26.f(), R, att.f(), spa.f(), con.f(), st));
  }
  return NLV_OK;
}

// tools/train_STTran.py:169-189 with bce_loss: CE(object) + CE(attention) + BCE(spatial) + BCE(contacting), rows weighted
// so that the batch loss is the mean over videos of the reference's per-video loss
int nlv_session::run_loss() {
  NLV_CHECK_ARG(B.lab_att && B.w_att && B.spa_bits && B.w_spa && B.con_bits && B.w_con, "session: loss needs the label arrays");
  loss = keep(1, 8, NLV_F32);
  const bool wg = want_ctx;
  d26 = wg ? keep(R, 26, NLV_F32) : T();
  dobj = (wg && M.mode != NLV_MODE_PREDCLS) ? keep(N, 37, NLV_F32) : T();
  OOM_CHECK();
  RUN(nlv_zero_bytes(loss.p, 32, st));
  if (M.mode != NLV_MODE_PREDCLS) {
    NLV_CHECK_ARG(B.w_obj != nullptr, "session: loss needs w_obj");
    RUN(nlv_ce_loss(obj_logits.f(), 37, 37, B.labels, B.w_obj, N, loss.f(), wg ? dobj.f() : nullptr, 37, st));
  }
  RUN(nlv_ce_loss(logits26.f(), 26, 3, B.lab_att, B.w_att, R, loss.f(), wg ? d26.f() : nullptr, 26, st));
  RUN(nlv_bce_sigmoid_loss(logits26.f() + 3, 26, 6, B.spa_bits, B.w_spa, R, loss.f(), wg ? d26.f() + 3 : nullptr, 26, st));
  RUN(nlv_bce_sigmoid_loss(logits26.f() + 9, 26, 17, B.con_bits, B.w_con, R, loss.f(), wg ? d26.f() + 9 : nullptr, 26, st));
  return NLV_OK;
}

int nlv_session::run_backward(const float* d26_in, const float* dobj_in) {
  NLV_CHECK_ARG(want_ctx, "session: backward needs a forward run with NLV_RUN_CTX");
  NLV_CHECK_ARG(M.grad_base != nullptr && !goff.empty(), "session: backward needs the gradient buffer and offsets");
  RUN(nlv_zero_bytes(M.grad_base, M.grad_elems * 4, st));
  const T dl = mk(d26_in != nullptr ? d26_in : d26.f(), NLV_F32, R, 26);
  NLV_CHECK_ARG(dry || dl.p != nullptr, "session: backward without logits gradient");
  T dx, drel;
  CK(heads_bwd(xf_out, dl, &dx));
  if (M.arch == NLV_ARCH_STTRAN) CK(sttran_transformer_bwd(dx, &drel));
  else CK(dsg_transformer_bwd(dx, &drel));
  CK(pair_tokens_bwd(drel));
  if (M.mode != NLV_MODE_PREDCLS) {
    const float* dob = dobj_in != nullptr ? dobj_in : dobj.f();
    if (dob != nullptr || dry) CK(object_classifier_bwd(mk(dob, NLV_F32, N, 37)));
  }
  return NLV_OK;
}

namespace {
void fill_outputs(nlv_session* s, nlv_outputs* out) {
  if (out == nullptr) return;
  memset(out, 0, sizeof(*out));
  out->obj_logits = s->obj_logits.f();
  out->logits26 = s->logits26.f();
  out->att = s->att.f(); out->spa = s->spa.f(); out->con = s->con.f();
  out->loss = s->loss.f();
  out->masks = s->masks.f();
  out->rel_tokens = s->rel.f();
  out->rel_out = s->xf_out.f();
  out->d26 = s->d26.f(); out->dobj = s->dobj.f();
}
void reset_step(nlv_session* s) {
  s->obj_logits = s->logits26 = s->att = s->spa = s->con = s->loss = s->d26 = s->dobj = s->masks = s->rel = s->xf_out = T();
  s->have_fwd = false;
}
}  // namespace

extern "C" {

// sizes of the structs that cross the ABI (the ctypes mirror in nlvsgg_b200/_C.py is checked against these)
int nlv_struct_sizes(int* out, int n) {
  const int v[4] = {(int)sizeof(nlv_model), (int)sizeof(nlv_batch), (int)sizeof(nlv_outputs), (int)sizeof(nlv_gemm_args)};
  for (int i = 0; i < n && i < 4; ++i) out[i] = v[i];
  return 4;
}

/* per-call timing of the sequencer's kernel entries: nlv_profile(1) starts recording (CUDA events around every entry),
 * nlv_profile_read synchronises and writes one line per call "name\tms\tflops\tunits\tm\tn\tk\tdtype\n" into buf (returns the
 * bytes needed; records are cleared).  Diagnostic only: adds two event records per call. */
int nlv_profile(int on) { nlv::g_prof_on = on != 0; return NLV_OK; }
long long nlv_profile_read(char* buf, long long cap) {
  cudaDeviceSynchronize();
  long long pos = 0;
  for (nlv::ProfRec& r : nlv::g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.e0, r.e1);
    char line[160];
    const int n = snprintf(line, sizeof(line), "%s\t%.6f\t%.0f\t%.0f\t%d\t%d\t%d\t%d\n", r.name, ms, r.flops, r.units, r.m, r.n, r.k, r.dt);
    if (buf != nullptr && pos + n < cap) memcpy(buf + pos, line, n);
    pos += n;
    cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
  }
  if (buf != nullptr && pos < cap) buf[pos] = 0;
  nlv::g_prof.clear();
  return pos + 1;
}

nlv_session* nlv_session_create(void) { return new nlv_session(); }
void nlv_session_destroy(nlv_session* s) { delete s; }

long long nlv_session_plan(nlv_session* s, const nlv_model* model, const nlv_batch* batch, int flags) {
  if (s == nullptr) { nlv::set_error("session: null handle"); return NLV_ERR_INVALID_ARGUMENT; }
  reset_step(s);
  s->xf_only = false;
  if (flags & NLV_RUN_BACKWARD) flags |= NLV_RUN_CTX;
  int rc = s->setup(model, batch, flags);
  if (rc != NLV_OK) return rc;
  s->arena(nullptr, 0, true);
  rc = s->run_forward();
  if (rc == NLV_OK && (flags & NLV_RUN_LOSS)) rc = s->run_loss();
  if (rc == NLV_OK && (flags & NLV_RUN_BACKWARD)) {
    if (s->goff.empty()) s->goff.assign(s->M.n_slots, 0);
    if (s->M.grad_base == nullptr) s->M.grad_base = reinterpret_cast<float*>(4096);
    rc = s->run_backward(nullptr, nullptr);
  }
  reset_step(s);
  if (rc != NLV_OK) return rc;
  return s->peak + 4096;
}

int nlv_session_forward(nlv_session* s, const nlv_model* model, const nlv_batch* batch, void* workspace, long long workspace_bytes,
                        int flags, nlv_outputs* out, void* stream) {
  NLV_CHECK_ARG(s != nullptr && workspace != nullptr, "session_forward: null handle / workspace");
  reset_step(s);
  s->xf_only = false;
  CK(s->setup(model, batch, flags));
  s->st = stream;
  s->arena(workspace, workspace_bytes, false);
  CK(s->run_forward());
  if (flags & NLV_RUN_LOSS) CK(s->run_loss());
  s->have_fwd = true;
  fill_outputs(s, out);
  return NLV_OK;
}

int nlv_session_set_gradients(nlv_session* s, float* grad_base, long long grad_elems, const long long* grad_offset, int n_slots) {
  NLV_CHECK_ARG(s != nullptr && grad_base != nullptr && grad_offset != nullptr, "session_set_gradients: null argument");
  NLV_CHECK_ARG(n_slots == s->M.n_slots, "session_set_gradients: %d offsets for a model of %d slots", n_slots, s->M.n_slots);
  s->M.grad_base = grad_base;
  s->M.grad_elems = grad_elems;
  s->goff.assign(grad_offset, grad_offset + n_slots);
  return NLV_OK;
}

int nlv_session_set_tail_hook(nlv_session* s, void (*fn)(void*), void* user) {
  NLV_CHECK_ARG(s != nullptr, "session_set_tail_hook: null handle");
  s->tail_hook = fn;
  s->tail_hook_user = user;
  return NLV_OK;
}

int nlv_session_backward(nlv_session* s, const float* d26, const float* dobj, void* stream) {
  NLV_CHECK_ARG(s != nullptr && s->have_fwd, "session_backward: no forward pass to differentiate");
  NLV_CHECK_ARG(!s->xf_only, "session_backward: the session ran the standalone transformer; use nlv_session_transformer_backward");
  s->st = stream;
  const int rc = s->run_backward(d26, dobj);
  s->have_fwd = false;
  return rc;
}

long long nlv_session_transformer_forward(nlv_session* s, const nlv_model* model, const nlv_batch* batch, const float* x, void* workspace,
                                          long long workspace_bytes, int flags, float** out, void* stream) {
  NLV_CHECK_ARG(s != nullptr, "transformer_forward: null handle");
  NLV_CHECK_ARG(model != nullptr && model->arch == NLV_ARCH_STTRAN, "transformer_forward: STTran layers only");
  reset_step(s);
  s->xf_only = true;
  const bool dry = workspace == nullptr;
  if (dry && (flags & NLV_RUN_BACKWARD)) flags |= NLV_RUN_CTX;
  CK(s->setup(model, batch, flags));
  s->st = stream;
  s->arena(workspace, workspace_bytes, dry);
  CK(s->prepare_weights());
  CK(s->sttran_transformer_fwd(mk(x, NLV_F32, s->R, D), &s->xf_out));
  if (dry) {
    if (flags & NLV_RUN_BACKWARD) {
      if (s->goff.empty()) s->goff.assign(s->M.n_slots, 0);
      if (s->M.grad_base == nullptr) s->M.grad_base = reinterpret_cast<float*>(4096);
      T d;
      CK(s->sttran_transformer_bwd(s->xf_out, &d));
    }
    const long long need = s->peak + 4096;
    reset_step(s);
    return need;
  }
  NLV_CHECK_ARG(out != nullptr, "transformer_forward: null out");
  s->have_fwd = true;
  *out = s->xf_out.f();
  return NLV_OK;
}

int nlv_session_transformer_backward(nlv_session* s, const float* dout, float** dx, void* stream) {
  NLV_CHECK_ARG(s != nullptr && s->have_fwd && s->xf_only, "transformer_backward: no transformer forward to differentiate");
  NLV_CHECK_ARG(s->want_ctx && s->M.grad_base != nullptr && !s->goff.empty(), "transformer_backward: needs NLV_RUN_CTX and the gradient buffer");
  s->st = stream;
  const bool dry = false;
  void* st = stream;
  RUN(nlv_zero_bytes(s->M.grad_base, s->M.grad_elems * 4, stream));
  T d;
  const int rc = s->sttran_transformer_bwd(mk(dout, NLV_F32, s->R, D), &d);
  s->have_fwd = false;
  if (rc != NLV_OK) return rc;
  if (dx != nullptr) *dx = d.f();
  return NLV_OK;
}

}  // extern "C"
