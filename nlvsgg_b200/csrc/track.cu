// DSG-DETR object tracking on the device: rectangular linear-sum assignment and the whole
// per-video `get_sequence(task="sgcls")` loop of lib/track.py:127-262 as ONE launch (one CTA per video).
//
// lsap_warp: shortest-augmenting-path assignment (Crouse 2016, the algorithm behind scipy.optimize.linear_sum_assignment,
// which lib/matcher.py:147-149 calls on the host) run by one warp with float64 duals.  The scan over the remaining columns
// is split across lanes; the column choice reproduces the sequential rule exactly (lowest reduced cost; among equal
// ones the last unassigned column in scan order, else the first), including the swap-remove order of the `remaining`
// list, so assignments are identical to scipy's even on ties.
//
// track_sequence_kernel: frames are inherently sequential (tracks evolve frame by frame), everything inside a frame is
// parallel: unit feature vectors of the live tracks, the (detections x tracks) cost matrix (class / feature cosine
// distances, L1, GIoU: lib/matcher.py:124-145), the assignment (warp 0), the accept / new-cluster bookkeeping in the
// reference's order (thread 0) and the running feature sums of the tracks.  The reference spends ~0.77 s per 20 frames in
// this loop (one .cpu() + scipy call + several torch.cat per frame); here a video is one kernel.
#include <float.h>

#include "common.cuh"

namespace nlv {
namespace {

constexpr int LSAP_MAX = 1024;      // larger side of one assignment problem (shared-memory state)
constexpr int TRACK_THREADS = 256;

struct LsapState {      // shared-memory arrays; nr <= nc (the problem is transposed when it has more rows than columns)
  double* spc;          // [nc] shortest path costs
  double* v;            // [nc] column duals
  double* u;            // [nr] row duals
  int* path;            // [nc]
  int* row4col;         // [nc]
  int* remaining;       // [nc]
  int* col4row;         // [nr]
  unsigned char* SC;    // [nc]
  unsigned char* SR;    // [nr]
};

__host__ __device__ inline size_t lsap_smem_bytes(int n) {
  return (size_t)n * (8 + 8 + 8 + 4 + 4 + 4 + 4 + 1 + 1) + 64;
}

__device__ inline LsapState lsap_carve(unsigned char* base, int n) {
  LsapState s;
  s.spc = reinterpret_cast<double*>(base); base += (size_t)n * 8;
  s.v = reinterpret_cast<double*>(base); base += (size_t)n * 8;
  s.u = reinterpret_cast<double*>(base); base += (size_t)n * 8;
  s.path = reinterpret_cast<int*>(base); base += (size_t)n * 4;
  s.row4col = reinterpret_cast<int*>(base); base += (size_t)n * 4;
  s.remaining = reinterpret_cast<int*>(base); base += (size_t)n * 4;
  s.col4row = reinterpret_cast<int*>(base); base += (size_t)n * 4;
  s.SC = base; base += n;
  s.SR = base;
  return s;
}

// One warp.  cost: float [n_rows_in x n_cols_in], row stride ld.  match_of_row[i] = assigned column of input row i or -1.
__device__ void lsap_warp(const float* __restrict__ cost, int n_rows_in, int n_cols_in, int ld, const LsapState& S,
                          int* __restrict__ match_of_row) {
  const int lane = threadIdx.x & 31;
  const bool tr = n_cols_in < n_rows_in;                     // scipy transposes when there are more rows than columns
  const int nr = tr ? n_cols_in : n_rows_in, nc = tr ? n_rows_in : n_cols_in;
  auto C = [&](int i, int j) -> double { return (double)(tr ? cost[(size_t)j * ld + i] : cost[(size_t)i * ld + j]); };
  for (int j = lane; j < nc; j += 32) { S.v[j] = 0.0; S.row4col[j] = -1; S.path[j] = -1; }
  for (int i = lane; i < nr; i += 32) { S.u[i] = 0.0; S.col4row[i] = -1; }
  __syncwarp();
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  for (int cur = 0; cur < nr; ++cur) {
    // ---- augmenting path from row `cur` ----
    for (int j = lane; j < nc; j += 32) { S.remaining[j] = nc - j - 1; S.SC[j] = 0; S.spc[j] = INF; }
    for (int i = lane; i < nr; i += 32) S.SR[i] = 0;
    __syncwarp();
    double min_val = 0.0;
    int num_remaining = nc, i = cur, sink = -1;
    while (sink == -1) {
      if (lane == 0) S.SR[i] = 1;
      const double ui = S.u[i];
      double best = INF;
      int first = 0x7fffffff, last_free = -1;
      for (int it = lane; it < num_remaining; it += 32) {
        const int j = S.remaining[it];
        const double r = ((min_val + C(i, j)) - ui) - S.v[j];
        double s = S.spc[j];
        if (r < s) { S.path[j] = i; S.spc[j] = r; s = r; }
        const bool free_col = S.row4col[j] == -1;
        if (s < best) { best = s; first = it; last_free = free_col ? it : -1; }
        else if (s == best) { if (first == 0x7fffffff) first = it; if (free_col) last_free = it; }
      }
      // warp reduction: the minimum, then over the lanes that hold it the first position and the last unassigned position
      double m = best;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_xor_sync(0xffffffffu, m, o); m = t < m ? t : m; }
      int f = (best == m) ? first : 0x7fffffff, lf = (best == m) ? last_free : -1;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        f = min(f, __shfl_xor_sync(0xffffffffu, f, o));
        lf = max(lf, __shfl_xor_sync(0xffffffffu, lf, o));
      }
      min_val = m;
      if (!(min_val < INF)) { sink = -2; break; }             // infeasible (no finite entry): leave the rest unassigned
      const int index = lf >= 0 ? lf : f;
      const int j = S.remaining[index];
      if (S.row4col[j] == -1) sink = j; else i = S.row4col[j];
      __syncwarp();
      if (lane == 0) { S.SC[j] = 1; S.remaining[index] = S.remaining[num_remaining - 1]; }
      --num_remaining;
      __syncwarp();
    }
    if (sink < 0) break;
    // ---- dual update ----
    if (lane == 0) S.u[cur] += min_val;
    for (int r = lane; r < nr; r += 32)
      if (S.SR[r] && r != cur) S.u[r] += min_val - S.spc[S.col4row[r]];
    for (int j = lane; j < nc; j += 32)
      if (S.SC[j]) S.v[j] -= min_val - S.spc[j];
    __syncwarp();
    // ---- augment ----
    if (lane == 0) {
      int j = sink;
      while (true) {
        const int r = S.path[j];
        S.row4col[j] = r;
        const int t = S.col4row[r];
        S.col4row[r] = j;
        j = t;
        if (r == cur) break;
      }
    }
    __syncwarp();
  }
  // result per INPUT row
  if (!tr) {
    for (int r = lane; r < n_rows_in; r += 32) match_of_row[r] = S.col4row[r];
  } else {
    for (int r = lane; r < n_rows_in; r += 32) match_of_row[r] = S.row4col[r];     // input rows are the columns of the transposed problem
  }
  __syncwarp();
}

__global__ void __launch_bounds__(32) lsap_kernel(const float* __restrict__ cost, int nr, int nc, int ld, int* __restrict__ match_of_row) {
  extern __shared__ __align__(16) unsigned char smem[];
  const LsapState S = lsap_carve(smem, max(nr, nc));
  lsap_warp(cost, nr, nc, ld, S, match_of_row);
}

// ---------------------------------------------------------------------------------------------------------------
struct TrackArgs {
  const float* boxes;        // [N,5] frame index (within its video), x1, y1, x2, y2
  const float* feats;        // [N,F]
  const int* cls;            // [N] arg-max class of the detection (the one-hot `dists` of lib/track.py:172-173)
  const int* det_off;        // [V+1] detections of video v
  const int* frame_off;      // [V+1] key frames of video v (into frame_start / frame_key)
  const int* frame_start;    // per video: T+1 offsets (relative to the video's first detection), concatenated with stride T_v + 1
  const int* frame_key;      // frame number of each key frame (track.py:175 int(name.split('/')[1].split('.')[0]))
  int F, n_cls;
  float img_w, img_h;        // `shape` argument: Z = [w, h, w, h]
  float wc, wf, wb, wg;
  int max_gap;               // 50 (track.py:55)
  // workspaces (per detection unless noted)
  float* unit;               // [N,F] unit feature vectors of the detections
  float* trk_sum;            // [N,F] running feature sum of the track created by detection slot i
  float* trk_unit;           // [N,F] unit vector of the mean feature of a live track
  int* trk_hist;             // [N,n_cls] class histogram of the track's members
  float* trk_box;            // [N,4] xywh
  int* trk_state;            // [N,4] key, cluster, member count, updated flag
  int* live;                 // [N] live track slots in list order
  float* cost;               // [sum over videos of max_det * N_v] cost / class / feature matrices, 3 planes
  long long cost_plane;      // elements per plane
  const long long* cost_off; // [V] offset of video v inside a plane
  int* match;                // [N] scratch: matched track position per detection of the current frame
  // outputs
  int* cluster_of_det;       // [N]
  int* n_clusters;           // [V]
  int* status;               // [V] 0 ok, 1 problem larger than LSAP_MAX
};

__device__ __forceinline__ bool outside_image(const float* p, float w, float h) {      // track.py:200,212,226
  return (p[0] + p[2] > h) || (p[1] + p[3] > w) || (p[0] < 0.f) || (p[1] < 0.f);
}

__global__ void __launch_bounds__(TRACK_THREADS) track_sequence_kernel(const TrackArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int s_nlive, s_ncluster, s_nops, s_bad;
  const int v = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = TRACK_THREADS / 32;
  const int d0 = a.det_off[v], N = a.det_off[v + 1] - d0;
  const int T = a.frame_off[v + 1] - a.frame_off[v];
  const int* fstart = a.frame_start + a.frame_off[v] + v;        // T + 1 entries per video
  const int* fkey = a.frame_key + a.frame_off[v];
  const int F = a.F;
  const float* boxes = a.boxes + (size_t)d0 * 5;
  const float* feats = a.feats + (size_t)d0 * F;
  const int* cls = a.cls + d0;
  float* unit = a.unit + (size_t)d0 * F;
  float* trk_sum = a.trk_sum + (size_t)d0 * F;
  float* trk_unit = a.trk_unit + (size_t)d0 * F;
  int* trk_hist = a.trk_hist + (size_t)d0 * a.n_cls;
  float* trk_box = a.trk_box + (size_t)d0 * 4;
  int* trk_state = a.trk_state + (size_t)d0 * 4;
  int* live = a.live + d0;
  int* match = a.match + d0;
  int* cluster_of = a.cluster_of_det + d0;
  float* cost = a.cost + a.cost_off[v];
  float* cost_d = cost + a.cost_plane;
  float* cost_f = cost + 2 * a.cost_plane;
  // ops recorded by the bookkeeping thread, applied by all threads: (detection, track slot, 1 = new track / 0 = append)
  int* ops = reinterpret_cast<int*>(smem);                       // [3 * max_det]; the LSAP state follows
  if (tid == 0) { s_nlive = 0; s_ncluster = 0; s_bad = 0; }
  // unit vectors of all detections (x / (|x| + 1e-12), matcher.py:74)
  for (int i = warp; i < N; i += nwarp) {
    const float* x = feats + (size_t)i * F;
    float s = 0.f;
    for (int k = lane; k < F; k += 32) s += x[k] * x[k];
    s = warp_sum(s);
    const float nrm = sqrtf(s) + 1e-12f;
    for (int k = lane; k < F; k += 32) unit[(size_t)i * F + k] = x[k] / nrm;
  }
  __syncthreads();
  int max_det = 0;
  for (int t = 0; t < T; ++t) max_det = max(max_det, fstart[t + 1] - fstart[t]);
  unsigned char* lsap_base = smem + (((size_t)3 * max_det * 4 + 15) & ~(size_t)15);

  for (int t = 0; t < T; ++t) {
    const int f0 = fstart[t], nd = fstart[t + 1] - f0, key = fkey[t];
    const int nl = s_nlive;
    __syncthreads();                                            // everyone has read the list length before thread 0 changes it
    const bool matching = nl > 0 && nd > 0;
    if (matching) {
      // unit mean features of the live tracks (mean = sum / count, then / (|mean| + 1e-12))
      for (int p = warp; p < nl; p += nwarp) {
        const int slot = live[p];
        const float cnt = (float)trk_state[slot * 4 + 2];
        const float* su = trk_sum + (size_t)slot * F;
        float s = 0.f;
        for (int k = lane; k < F; k += 32) { const float y = su[k] / cnt; s += y * y; }
        s = warp_sum(s);
        const float nrm = sqrtf(s) + 1e-12f;
        for (int k = lane; k < F; k += 32) trk_unit[(size_t)slot * F + k] = (su[k] / cnt) / nrm;
      }
      __syncthreads();
      // cost matrix [nd x nl]: one warp per (detection, track)
      for (int e = warp; e < nd * nl; e += nwarp) {
        const int i = e / nl, p = e % nl, slot = live[p], di = f0 + i;
        const float* x = unit + (size_t)di * F;
        const float* y = trk_unit + (size_t)slot * F;
        float dot = 0.f;
        for (int k = lane; k < F; k += 32) dot += x[k] * y[k];
        dot = warp_sum(dot);
        if (lane == 0) {
          const float cf = 1.f - dot;
          // class term: the detection is one-hot, the track holds the mean of one-hots -> hist / count, normalised
          const int* hist = trk_hist + (size_t)slot * a.n_cls;
          const float cnt = (float)trk_state[slot * 4 + 2];
          float hs = 0.f;
          for (int c = 0; c < a.n_cls; ++c) { const float y2 = (float)hist[c] / cnt; hs += y2 * y2; }
          const float cd = 1.f - ((float)hist[cls[di]] / cnt) / (sqrtf(hs) + 1e-12f);
          // boxes: xyxy -> xywh, / Z (track.py:181-186), cxcywh L1 + GIoU (matcher.py:124-142)
          const float* b = boxes + (size_t)di * 5 + 1;
          const float Z[4] = {a.img_w, a.img_h, a.img_w, a.img_h};
          const float pa[4] = {b[0] / Z[0], b[1] / Z[1], (b[2] - b[0]) / Z[2], (b[3] - b[1]) / Z[3]};
          const float* tb = trk_box + (size_t)slot * 4;
          const float pb[4] = {tb[0] / Z[0], tb[1] / Z[1], tb[2] / Z[2], tb[3] / Z[3]};
          const float ac[4] = {pa[0] + pa[2] / 2, pa[1] + pa[3] / 2, pa[2], pa[3]};
          const float bc[4] = {pb[0] + pb[2] / 2, pb[1] + pb[3] / 2, pb[2], pb[3]};
          const float l1 = fabsf(ac[0] - bc[0]) + fabsf(ac[1] - bc[1]) + fabsf(ac[2] - bc[2]) + fabsf(ac[3] - bc[3]);
          const float ax[4] = {ac[0] - 0.5f * ac[2], ac[1] - 0.5f * ac[3], ac[0] + 0.5f * ac[2], ac[1] + 0.5f * ac[3]};
          const float bx[4] = {bc[0] - 0.5f * bc[2], bc[1] - 0.5f * bc[3], bc[0] + 0.5f * bc[2], bc[1] + 0.5f * bc[3]};
          const float area1 = (ax[2] - ax[0]) * (ax[3] - ax[1]), area2 = (bx[2] - bx[0]) * (bx[3] - bx[1]);
          const float iw = fmaxf(fminf(ax[2], bx[2]) - fmaxf(ax[0], bx[0]), 0.f), ih = fmaxf(fminf(ax[3], bx[3]) - fmaxf(ax[1], bx[1]), 0.f);
          const float inter = iw * ih, uni = area1 + area2 - inter, iou = inter / uni;
          const float ew = fmaxf(fmaxf(ax[2], bx[2]) - fminf(ax[0], bx[0]), 0.f), eh = fmaxf(fmaxf(ax[3], bx[3]) - fminf(ax[1], bx[1]), 0.f);
          const float earea = ew * eh, giou = iou - (earea - uni) / earea;
          cost[(size_t)i * nl + p] = a.wc * cd + a.wf * cf + a.wb * l1 + a.wg * (-giou);
          cost_d[(size_t)i * nl + p] = cd;
          cost_f[(size_t)i * nl + p] = cf;
        }
      }
      __syncthreads();
      if (max(nd, nl) > LSAP_MAX) {
        if (tid == 0) s_bad = 1;
      } else if (warp == 0) {
        const LsapState S = lsap_carve(lsap_base, max(nd, nl));
        lsap_warp(cost, nd, nl, nl, S, match + f0);
      }
      __syncthreads();
      if (s_bad) break;
    }
    // ---- bookkeeping in the reference's order (track.py:189-231): matched detections by row, then the unmatched ones ----
    if (tid == 0) {
      int nops = 0, nlive = nl, ncl = s_ncluster;
      for (int p = 0; p < nl; ++p) trk_state[live[p] * 4 + 3] = 0;                // tracker.updated = False (:177-178)
      for (int pass = 0; pass < 2; ++pass) {
        for (int i = 0; i < nd; ++i) {
          const int di = f0 + i;
          const int p = matching ? match[di] : -1;
          if ((pass == 0) != (p >= 0)) continue;
          const float* b = boxes + (size_t)di * 5 + 1;
          const float pred[4] = {b[0], b[1], b[2] - b[0], b[3] - b[1]};            // xyxy -> xywh
          const bool out = outside_image(pred, a.img_w, a.img_h);
          bool accepted = false;
          if (p >= 0) accepted = (cost_d[(size_t)i * nl + p] < 0.5f) || (cost_f[(size_t)i * nl + p] < 0.5f);   // tau (:197)
          if (accepted) {
            const int slot = live[p];
            cluster_of[di] = trk_state[slot * 4 + 1];
            if (out) continue;
            ops[3 * nops] = di; ops[3 * nops + 1] = slot; ops[3 * nops + 2] = 0; ++nops;
            trk_state[slot * 4 + 2] += 1;
            trk_state[slot * 4 + 0] = key; trk_state[slot * 4 + 3] = 1;           // Tracker.update(box, key)
            trk_box[slot * 4 + 0] = pred[0]; trk_box[slot * 4 + 1] = pred[1]; trk_box[slot * 4 + 2] = pred[2]; trk_box[slot * 4 + 3] = pred[3];
          } else {
            cluster_of[di] = ncl;
            if (!out) {                                                          // a new track, stored in the slot of its first detection
              const int slot = di;
              ops[3 * nops] = di; ops[3 * nops + 1] = slot; ops[3 * nops + 2] = 1; ++nops;
              trk_state[slot * 4 + 0] = key; trk_state[slot * 4 + 1] = ncl; trk_state[slot * 4 + 2] = 1; trk_state[slot * 4 + 3] = 0;
              trk_box[slot * 4 + 0] = pred[0]; trk_box[slot * 4 + 1] = pred[1]; trk_box[slot * 4 + 2] = pred[2]; trk_box[slot * 4 + 3] = pred[3];
              live[nlive++] = slot;
            }
            ++ncl;
          }
        }
      }
      // tracks that were not updated survive while the gap stays below max_gap (:232-241, Tracker.update(None, key))
      int w = 0;
      for (int p = 0; p < nlive; ++p) {
        const int slot = live[p];
        const bool keep = trk_state[slot * 4 + 3] != 0 || (key - trk_state[slot * 4 + 0] < a.max_gap);
        if (keep) live[w++] = slot;
      }
      s_nlive = w; s_ncluster = ncl; s_nops = nops;
    }
    __syncthreads();
    // ---- feature sums / class histograms of the touched tracks ----
    const int nops = s_nops;
    for (int o = 0; o < nops; ++o) {
      const int di = ops[3 * o], slot = ops[3 * o + 1], fresh = ops[3 * o + 2];
      const float* x = feats + (size_t)di * F;
      float* su = trk_sum + (size_t)slot * F;
      for (int k = tid; k < F; k += TRACK_THREADS) su[k] = fresh ? x[k] : su[k] + x[k];
      int* hist = trk_hist + (size_t)slot * a.n_cls;
      for (int c = tid; c < a.n_cls; c += TRACK_THREADS) hist[c] = (fresh ? 0 : hist[c]) + (c == cls[di] ? 1 : 0);
    }
    __syncthreads();
  }
  if (tid == 0) { a.n_clusters[v] = s_ncluster; a.status[v] = s_bad; }
}

}  // namespace
}  // namespace nlv

using namespace nlv;
#define STREAM ((cudaStream_t)stream)

extern "C" {

int nlv_lsap(const float* cost, int n_rows, int n_cols, int ld, int* match_of_row, void* stream) {
  NLV_CHECK_ARG(n_rows >= 0 && n_cols >= 0 && ld >= n_cols, "lsap: bad sizes");
  if (n_rows == 0) return NLV_OK;
  NLV_CHECK_ARG(match_of_row != nullptr, "lsap: null pointer");
  if (n_cols == 0) { NLV_CHECK_CUDA(cudaMemsetAsync(match_of_row, 0xff, (size_t)n_rows * 4, STREAM)); return NLV_OK; }
  NLV_CHECK_ARG(cost != nullptr, "lsap: null pointer");
  const int n = n_rows > n_cols ? n_rows : n_cols;
  NLV_CHECK_ARG(n <= LSAP_MAX, "lsap: %d x %d exceeds the %d-wide shared-memory state", n_rows, n_cols, LSAP_MAX);
  lsap_kernel<<<1, 32, lsap_smem_bytes(n), STREAM>>>(cost, n_rows, n_cols, ld, match_of_row);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

long long nlv_track_sequence_workspace(long long n_det_total, int feat_dim, int n_cls, long long cost_elems) {
  // unit, trk_sum, trk_unit | trk_hist | trk_box, trk_state, live, match | 3 cost planes
  return n_det_total * (long long)feat_dim * 4 * 3 + n_det_total * (long long)n_cls * 4 + n_det_total * 4 * (4 + 4 + 1 + 1) +
         cost_elems * 4 * 3 + 256;
}

int nlv_track_sequence(const float* boxes, const float* feats, int feat_dim, const int* cls, int n_cls, const int* det_off,
                       const int* frame_off, const int* frame_start, const int* frame_key, int n_videos, long long n_det_total,
                       int max_det_per_frame, float img_w, float img_h, float w_class, float w_feat, float w_bbox, float w_giou,
                       int max_gap, const long long* cost_off, long long cost_elems, void* workspace, int* cluster_of_det,
                       int* n_clusters, int* status, void* stream) {
  NLV_CHECK_ARG(n_videos >= 0 && n_det_total >= 0 && feat_dim > 0 && n_cls > 0 && max_det_per_frame >= 0, "track_sequence: bad sizes");
  if (n_videos == 0) return NLV_OK;
  NLV_CHECK_ARG(boxes && feats && cls && det_off && frame_off && frame_start && frame_key && cost_off && workspace && cluster_of_det &&
                    n_clusters && status,
                "track_sequence: null pointer");
  NLV_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "track_sequence: workspace must be 16-byte aligned");
  TrackArgs a;
  a.boxes = boxes; a.feats = feats; a.cls = cls; a.det_off = det_off; a.frame_off = frame_off; a.frame_start = frame_start;
  a.frame_key = frame_key; a.F = feat_dim; a.n_cls = n_cls; a.img_w = img_w; a.img_h = img_h;
  a.wc = w_class; a.wf = w_feat; a.wb = w_bbox; a.wg = w_giou; a.max_gap = max_gap;
  char* p = reinterpret_cast<char*>(workspace);
  const size_t nf = (size_t)n_det_total * feat_dim * 4;
  a.unit = reinterpret_cast<float*>(p); p += nf;
  a.trk_sum = reinterpret_cast<float*>(p); p += nf;
  a.trk_unit = reinterpret_cast<float*>(p); p += nf;
  a.trk_hist = reinterpret_cast<int*>(p); p += (size_t)n_det_total * n_cls * 4;
  a.trk_box = reinterpret_cast<float*>(p); p += (size_t)n_det_total * 16;
  a.trk_state = reinterpret_cast<int*>(p); p += (size_t)n_det_total * 16;
  a.live = reinterpret_cast<int*>(p); p += (size_t)n_det_total * 4;
  a.match = reinterpret_cast<int*>(p); p += (size_t)n_det_total * 4;
  p = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
  a.cost = reinterpret_cast<float*>(p);
  a.cost_plane = cost_elems; a.cost_off = cost_off;
  a.cluster_of_det = cluster_of_det; a.n_clusters = n_clusters; a.status = status;
  const size_t smem = (((size_t)3 * max_det_per_frame * 4 + 15) & ~(size_t)15) + lsap_smem_bytes(LSAP_MAX);
  NLV_CHECK_ARG(smem <= 200 * 1024, "track_sequence: %d detections in one frame exceed the shared-memory budget", max_det_per_frame);
  static bool attr_set = false;
  if (!attr_set) {
    NLV_CHECK_CUDA(cudaFuncSetAttribute(track_sequence_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  track_sequence_kernel<<<n_videos, TRACK_THREADS, smem, STREAM>>>(a);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

}
