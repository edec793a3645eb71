// Recall@K triplet matching on device — lib/evaluation_recall.py:397-467 (evaluate_scene_graph), :209-236
// (with constraint), :321-353 (no constraint, top-100), :257-302 (semi constraint), :630-695 (evaluate_recall),
// :731-773 (_compute_pred_matches) and lib/fpn/box_intersections_cpu/bbox.pyx:21-61 (float64 IoU, +1 convention).
//
// One CTA per frame; everything a frame needs lives in shared memory.  Outputs are integers only: for every frame,
// protocol (with / no / semi constraint) and K in {10,20,50}, the 256-bit set of ground-truth relations matched by
// the K best-scored predicted triplets.  The host turns set sizes into the reference's per-frame floats
// (|set| / G) in the reference's order, so recall values are bit-identical.
//
// Score arithmetic follows numpy's types exactly: relation scores are float32 values widened to float64; the
// triplet score is (f64(s_sub) * f64(s_obj)) * f64(s_rel); the no-constraint ranking score is
// f64(f32(s_sub * s_obj)) * f64(s_rel).  Tie order is the canonical one of SURVEY.md §7: no-constraint selection —
// descending, ties by ascending flat index; evaluate_recall — descending, ties by descending position.
#include "common.cuh"

namespace nlv {
namespace {

constexpr int P_MAX = 40;            // pairs per frame
constexpr int G_MAX = 256;           // ground-truth relations per frame
constexpr int GB_MAX = 64;           // ground-truth boxes per frame
constexpr int C_MAX = 24 * P_MAX;    // candidates per protocol (semi: <= P + 6P + 17P)
constexpr int NPOS_MAX = 26 * P_MAX; // positive entries of the 3P x 26 block-diagonal score matrix
constexpr int TOPN = 100;
constexpr int THREADS = 128;

struct EvalArgs {
  int n_frames;
  const int *pair_off, *gtrel_off, *gtbox_off;                 // [F+1] each
  const int *pair_sub, *pair_obj;                              // [R] indices into the pred_* arrays
  const float *att, *spa, *con;                                // [R,3] (softmaxed), [R,6], [R,17]
  const float* obj_scores; const int* pred_cls; const float* pred_boxes;   // [NB], [NB], [NB,4]
  const int* gt_rel; const int* gt_cls; const float* gt_boxes;             // [G,3] (sub,obj,pred local), [GB], [GB,4] (f32-rounded)
  unsigned* out;                                               // [F,3,3,8]
};

__device__ __forceinline__ double iou_f64(const float* g, const float* q) {
  // bbox_overlaps(boxes=gt, query=pred): float64 arithmetic on float32-rounded coordinates
  const double gx1 = g[0], gy1 = g[1], gx2 = g[2], gy2 = g[3];
  const double qx1 = q[0], qy1 = q[1], qx2 = q[2], qy2 = q[3];
  const double qa = __dmul_rn(__dadd_rn(__dsub_rn(qx2, qx1), 1.0), __dadd_rn(__dsub_rn(qy2, qy1), 1.0));
  const double iw = __dadd_rn(__dsub_rn(fmin(gx2, qx2), fmax(gx1, qx1)), 1.0);
  if (!(iw > 0)) return 0.0;
  const double ih = __dadd_rn(__dsub_rn(fmin(gy2, qy2), fmax(gy1, qy1)), 1.0);
  if (!(ih > 0)) return 0.0;
  const double ga = __dmul_rn(__dadd_rn(__dsub_rn(gx2, gx1), 1.0), __dadd_rn(__dsub_rn(gy2, gy1), 1.0));
  const double inter = __dmul_rn(iw, ih);
  return __ddiv_rn(inter, __dsub_rn(__dadd_rn(ga, qa), inter));
}

struct Smem {
  // candidates of the protocol being processed
  int c_sub[C_MAX], c_obj[C_MAX], c_pred[C_MAX];
  double c_score[C_MAX], c_trip[C_MAX];
  // positive entries of the score matrix (no-constraint selection)
  double p_ov[NPOS_MAX];
  int p_flat[NPOS_MAX];
  int sel[TOPN];
  int row_cnt[3 * P_MAX + 1];
  // ground truth
  int g_sub[G_MAX], g_obj[G_MAX], g_pred[G_MAX];
  int gb_cls[GB_MAX];
  float gb_box[GB_MAX][4];
  unsigned mask[3][8];
  int n_pos, n_cand;
};

// value of the virtual [3P,26] block-diagonal matrix (evaluation_recall.py:433-442)
__device__ __forceinline__ float rel_value(const EvalArgs& a, int pr0, int P, int i, int c) {
  const int t = i / P, p = pr0 + (i - t * P);
  if (t == 0) return c < 3 ? a.att[(size_t)p * 3 + c] : 0.f;
  if (t == 1) return (c >= 3 && c < 9) ? a.spa[(size_t)p * 6 + (c - 3)] : 0.f;
  return c >= 9 ? a.con[(size_t)p * 17 + (c - 9)] : 0.f;
}
__device__ __forceinline__ void row_boxes(const EvalArgs& a, int pr0, int P, int i, int& sub, int& obj) {
  const int t = i / P, p = pr0 + (i - t * P);
  if (t == 1) { sub = a.pair_obj[p]; obj = a.pair_sub[p]; } else { sub = a.pair_sub[p]; obj = a.pair_obj[p]; }
}
// np.argmax / max over the full 26-wide row (first maximum wins; values are >= 0)
__device__ __forceinline__ void row_argmax(const EvalArgs& a, int pr0, int P, int i, int& arg, float& mx) {
  arg = 0; mx = rel_value(a, pr0, P, i, 0);
  for (int c = 1; c < 26; ++c) {
    const float v = rel_value(a, pr0, P, i, c);
    if (v > mx) { mx = v; arg = c; }
  }
}

// evaluate_recall on the candidates in S: rank by triplet score, match, OR the matched-GT sets of the top K.
__device__ void rank_and_match(const EvalArgs& a, Smem& S, int G) {
  const int n = S.n_cand;
  for (int i = threadIdx.x; i < 24; i += THREADS) S.mask[i / 8][i % 8] = 0u;
  for (int i = threadIdx.x; i < n; i += THREADS)
    S.c_trip[i] = __dmul_rn(__dmul_rn((double)a.obj_scores[S.c_sub[i]], (double)a.obj_scores[S.c_obj[i]]), S.c_score[i]);
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += THREADS) {
    const double ti = S.c_trip[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const double tj = S.c_trip[j];
      rank += (tj > ti || (tj == ti && j > i)) ? 1 : 0;
    }
    if (rank >= 50) continue;
    const int sub = S.c_sub[i], obj = S.c_obj[i];
    const int cs = a.pred_cls[sub], co = a.pred_cls[obj], cp = S.c_pred[i];
    for (int g = 0; g < G; ++g) {
      if (S.g_pred[g] != cp || S.gb_cls[S.g_sub[g]] != cs || S.gb_cls[S.g_obj[g]] != co) continue;
      if (iou_f64(S.gb_box[S.g_sub[g]], a.pred_boxes + 4 * (size_t)sub) >= 0.5 &&
          iou_f64(S.gb_box[S.g_obj[g]], a.pred_boxes + 4 * (size_t)obj) >= 0.5) {
        const unsigned bit = 1u << (g & 31);
        if (rank < 10) atomicOr(&S.mask[0][g >> 5], bit);
        if (rank < 20) atomicOr(&S.mask[1][g >> 5], bit);
        atomicOr(&S.mask[2][g >> 5], bit);
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(THREADS) recall_match_kernel(EvalArgs a) {
  __shared__ Smem S;
  for (int f = blockIdx.x; f < a.n_frames; f += gridDim.x) {
    const int pr0 = a.pair_off[f], P = a.pair_off[f + 1] - pr0;
    const int gr0 = a.gtrel_off[f], G = a.gtrel_off[f + 1] - gr0;
    const int gb0 = a.gtbox_off[f], GB = a.gtbox_off[f + 1] - gb0;
    unsigned* out = a.out + (size_t)f * 72;
    __syncthreads();
    for (int i = threadIdx.x; i < G; i += THREADS) {
      S.g_sub[i] = a.gt_rel[3 * (size_t)(gr0 + i)]; S.g_obj[i] = a.gt_rel[3 * (size_t)(gr0 + i) + 1];
      S.g_pred[i] = a.gt_rel[3 * (size_t)(gr0 + i) + 2];
    }
    for (int i = threadIdx.x; i < GB; i += THREADS) {
      S.gb_cls[i] = a.gt_cls[gb0 + i];
#pragma unroll
      for (int c = 0; c < 4; ++c) S.gb_box[i][c] = a.gt_boxes[4 * (size_t)(gb0 + i) + c];
    }
    if (threadIdx.x == 0) { S.n_pos = 0; S.n_cand = 0; }
    __syncthreads();
    if (P == 0 || G == 0) {  // evaluate_recall returns [[]] for an empty prediction set: no matches
      for (int i = threadIdx.x; i < 72; i += THREADS) out[i] = 0u;
      continue;
    }
    const int rows = 3 * P;

    // ---------------- with constraint (:209-236) ----------------
    for (int i = threadIdx.x; i < rows; i += THREADS) {
      int arg; float mx;
      row_argmax(a, pr0, P, i, arg, mx);
      row_boxes(a, pr0, P, i, S.c_sub[i], S.c_obj[i]);
      S.c_pred[i] = arg; S.c_score[i] = (double)mx;
    }
    if (threadIdx.x == 0) S.n_cand = rows;
    __syncthreads();
    rank_and_match(a, S, G);
    for (int i = threadIdx.x; i < 24; i += THREADS) out[i] = S.mask[i / 8][i % 8];
    __syncthreads();

    // ---------------- no constraint (:321-353): top-100 of f32(os*oo) * rel ----------------
    for (int e = threadIdx.x; e < rows * 26; e += THREADS) {
      const int i = e / 26, c = e - i * 26;
      const float v = rel_value(a, pr0, P, i, c);
      int sub, obj;
      row_boxes(a, pr0, P, i, sub, obj);
      const float per_rel = __fmul_rn(a.obj_scores[sub], a.obj_scores[obj]);
      const double ov = __dmul_rn((double)per_rel, (double)v);
      if (ov > 0.0) {
        const int slot = atomicAdd(&S.n_pos, 1);
        S.p_ov[slot] = ov; S.p_flat[slot] = e;
      }
    }
    __syncthreads();
    const int n_pos = S.n_pos;
    for (int i = threadIdx.x; i < TOPN; i += THREADS) S.sel[i] = -1;
    __syncthreads();
    for (int i = threadIdx.x; i < n_pos; i += THREADS) {
      const double oi = S.p_ov[i];
      const int fi = S.p_flat[i];
      int rank = 0;
      for (int j = 0; j < n_pos; ++j) rank += (S.p_ov[j] > oi || (S.p_ov[j] == oi && S.p_flat[j] < fi)) ? 1 : 0;
      if (rank < TOPN) S.sel[rank] = fi;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int n_sel = n_pos < TOPN ? n_pos : TOPN;
      // fewer than 100 positive entries: the remaining slots are the zero-score entries in flat order
      // (np.argsort(-x, kind='stable') keeps equal keys in index order; negative values cannot occur)
      for (int e = 0; e < rows * 26 && n_sel < TOPN; ++e) {
        const int i = e / 26, c = e - i * 26;
        int sub, obj;
        row_boxes(a, pr0, P, i, sub, obj);
        const double ov = __dmul_rn((double)__fmul_rn(a.obj_scores[sub], a.obj_scores[obj]), (double)rel_value(a, pr0, P, i, c));
        if (!(ov > 0.0)) S.sel[n_sel++] = e;
      }
      S.n_cand = n_sel;
    }
    __syncthreads();
    for (int r = threadIdx.x; r < S.n_cand; r += THREADS) {
      const int e = S.sel[r], i = e / 26, c = e - i * 26;
      row_boxes(a, pr0, P, i, S.c_sub[r], S.c_obj[r]);
      S.c_pred[r] = c; S.c_score[r] = (double)rel_value(a, pr0, P, i, c);
    }
    __syncthreads();
    rank_and_match(a, S, G);
    for (int i = threadIdx.x; i < 24; i += THREADS) out[24 + i] = S.mask[i / 8][i % 8];
    __syncthreads();

    // ---------------- semi constraint (:257-302) ----------------
    for (int i = threadIdx.x; i < rows; i += THREADS) {
      int cnt = 0;
      const double r0 = (double)rel_value(a, pr0, P, i, 0) + (double)rel_value(a, pr0, P, i, 1);
      if (r0 > 0) cnt = 1;
      else if ((double)rel_value(a, pr0, P, i, 3) + (double)rel_value(a, pr0, P, i, 4) > 0 ||
               (double)rel_value(a, pr0, P, i, 9) + (double)rel_value(a, pr0, P, i, 10) > 0) {
        for (int c = 0; c < 26; ++c) cnt += rel_value(a, pr0, P, i, c) > 0.5f ? 1 : 0;
      }
      S.row_cnt[i] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int acc = 0;
      for (int i = 0; i < rows; ++i) { const int c = S.row_cnt[i]; S.row_cnt[i] = acc; acc += c; }
      S.row_cnt[rows] = acc;
      S.n_cand = acc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rows; i += THREADS) {
      int pos = S.row_cnt[i];
      const int end = S.row_cnt[i + 1];
      if (pos == end) continue;
      int sub, obj;
      row_boxes(a, pr0, P, i, sub, obj);
      const double r0 = (double)rel_value(a, pr0, P, i, 0) + (double)rel_value(a, pr0, P, i, 1);
      if (r0 > 0) {
        int arg; float mx;
        row_argmax(a, pr0, P, i, arg, mx);
        S.c_sub[pos] = sub; S.c_obj[pos] = obj; S.c_pred[pos] = arg; S.c_score[pos] = (double)mx;
      } else {
        for (int c = 0; c < 26; ++c) {
          const float v = rel_value(a, pr0, P, i, c);
          if (v > 0.5f) { S.c_sub[pos] = sub; S.c_obj[pos] = obj; S.c_pred[pos] = c; S.c_score[pos] = (double)v; ++pos; }
        }
      }
    }
    __syncthreads();
    rank_and_match(a, S, G);
    for (int i = threadIdx.x; i < 24; i += THREADS) out[48 + i] = S.mask[i / 8][i % 8];
  }
}

}  // namespace
}  // namespace nlv

extern "C" {

/* Limits per frame: <= 40 pairs, <= 256 ground-truth relations, <= 64 ground-truth boxes (host checks, NLV_ERR_UNSUPPORTED). */
int nlv_recall_match(int n_frames, const int* pair_off, const int* gtrel_off, const int* gtbox_off, const int* pair_sub,
                     const int* pair_obj, const float* att, const float* spa, const float* con, const float* obj_scores,
                     const int* pred_cls, const float* pred_boxes, const int* gt_rel, const int* gt_cls,
                     const float* gt_boxes, unsigned* out, void* stream) {
  using namespace nlv;
  NLV_CHECK_ARG(n_frames >= 0, "recall_match: bad sizes");
  if (n_frames == 0) return NLV_OK;
  NLV_CHECK_ARG(pair_off && gtrel_off && gtbox_off && out, "recall_match: null pointer");
  EvalArgs a{n_frames, pair_off, gtrel_off, gtbox_off, pair_sub, pair_obj, att, spa, con, obj_scores, pred_cls, pred_boxes,
             gt_rel, gt_cls, gt_boxes, out};
  int grid = n_frames;
  const int cap = 16 * sm_count();
  if (grid > cap) grid = cap;
  recall_match_kernel<<<grid, THREADS, 0, (cudaStream_t)stream>>>(a);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int nlv_recall_limits(int* p_max, int* g_max, int* gb_max) {
  if (p_max) *p_max = nlv::P_MAX;
  if (g_max) *g_max = nlv::G_MAX;
  if (gb_max) *gb_max = nlv::GB_MAX;
  return NLV_OK;
}
}
