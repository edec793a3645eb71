// Tensor-core variant of the varlen attention kernels (attn.cu) for bf16 inputs / outputs.
//
// Same work decomposition (16 rows of one segment x one head) but a whole work item is ONE warp: the 16 x hd Q tile and
// 16-row K / V tiles are staged in shared memory with 4-byte cp.async (head slices start on 4-byte boundaries only:
// head_dim 242 -> 484-byte slices; every warp-copy is one contiguous 128-byte run), zero-padded to 256 columns, and all
// products run on mma.sync.m16n8k16 (bf16 in, fp32 accumulate):
//   forward   S = Q K^T (A, B by ldmatrix), online softmax on the accumulator fragments, O += P V (P re-used from the
//             accumulator registers as the A fragment, V by ldmatrix.trans)
//   backward  query side: S, dP = dO V^T, dS = P o (dP - delta), dQ += dS K
//             key side:   S^T = K Q^T, dP^T = V dO^T (operands swapped so P^T / dS^T come out as A fragments),
//                         dV += P^T dO, dK += dS^T Q; two warps per work item, each owning half of the head columns
//   delta = rowsum(dO o O) is obtained as rowsum(P o dP) inside the query-side kernel (no O tile needed).
// tcgen05 is not used here on purpose: the segments are 6-40 rows, far below the 128-row UMMA tile; the kernels are bound
// by moving Q/K/V/dO once (HBM), and mma.sync on 16-row tiles already makes the arithmetic a small fraction of the copy time.
// Replaces torch.nn.MultiheadAttention's core (lib/transformer.py:9-13,38-42, lib/dsg_detr.py:21-22) with
// key_padding_mask semantics folded into the segment bounds.
#include <initializer_list>
#include <stdlib.h>

#include "common.cuh"
#include "philox.cuh"

namespace nlv {
namespace {

typedef __nv_bfloat16 bf16;
constexpr int KP = 264;            // padded tile row, bf16 elements (256 + 8: ldmatrix rows land on different banks)
constexpr int ROWB = KP * 2;       // 528 bytes
constexpr int TILE_B = 16 * ROWB;  // 8448 bytes: one 16-row tile
constexpr int NT = 32;             // 8-column output tiles covering 256 columns

struct Lead { int lw; bool wide; };

struct MArgs {
  const bf16 *q, *k, *v;
  int ldq, ldk, ldv;
  int hd, heads;
  float scale;
  const int4* work;
  DropCfg drop;     // dropout on the attention weights (nn.MultiheadAttention(dropout=0.1), lib/transformer.py:9,38)
  int skip_short;   // backward: segments of <= 16 rows are taken by the fused single-tile kernel; the two-kernel path skips them
  int wide;         // every pointer / row stride / slice width is 16-byte aligned: tiles hold aligned supersets, 16-byte copies
};
__device__ __forceinline__ Lead lead_of(const MArgs& a, int col0) {
  Lead le;
  le.wide = a.wide != 0;
  le.lw = le.wide ? (int)(((unsigned)col0 * 2u) & 15u) >> 2 : 0;
  return le;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float quad_max(float x) {
  x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 1));
  return fmaxf(x, __shfl_xor_sync(0xffffffffu, x, 2));
}
__device__ __forceinline__ float quad_sum(float x) {
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  return x + __shfl_xor_sync(0xffffffffu, x, 2);
}

// Tile layout with 16-byte copies: a head slice starts 4 * (h % 4) bytes past a 16-byte boundary of its row (484-byte
// slices), so the tile holds the 16-byte ALIGNED superset of the slice: logical word w of the head sits at tile word w + lw
// (lw = lead words, 0..3) and one cp.async.16 per lane moves a whole row (31 chunks) instead of four 4-byte copies.  The
// words in front of / behind the slice inside the first / last chunk belong to the neighbouring heads: they are zeroed
// after the copy in the K and V tiles (`fix_junk`) — every product contracting over the columns has K or V as one operand
// — and are never stored.  lw = 0 with 4-byte copies when the pointers / strides are not 16-byte aligned.

// zero the words outside [lw, lw + nwords) of a 16-row tile that copies never touch (lw + nwords rounded up to a chunk .. 127).
// Lane l owns row l & 15 and every second 4-word group of the pad: no index divisions (the pad is 0..4 groups wide per row).
__device__ __forceinline__ void zero_pad(uint32_t tile, int nwords, int lane, Lead ld = Lead{0, false}) {
  const int first = ld.wide ? ((ld.lw + nwords + 3) & ~3) : nwords;
  const uint32_t row = tile + (lane & 15) * ROWB;
  for (int w = first + (lane >> 4) * 4; w < 128; w += 8) {
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (w + j < 128) asm volatile("st.shared.b32 [%0], %1;" ::"r"(row + (w + j) * 4), "r"(0u) : "memory");
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// rows [first, first + cnt) of `src` (already offset to the head's first column) -> tile rows [r0, r0 + nr); rows >= cnt are zeroed
__device__ __forceinline__ void load_rows(uint32_t tile, const bf16* src, int ld, long long first, int cnt, int r0, int nr, int nwords,
                                          int lane, Lead le = Lead{0, false}) {
  if (le.wide) {
    const int nchunks = (le.lw + nwords + 3) >> 2;      // 16-byte chunks of the aligned superset (31 for 242-wide heads)
    for (int r = r0; r < r0 + nr; ++r) {
      const uint32_t dst = tile + r * ROWB;
      if (r < cnt) {
        const uint8_t* s = reinterpret_cast<const uint8_t*>(src + (size_t)(first + r) * ld) - 4 * le.lw;
        if (lane < nchunks) cp_async16(dst + lane * 16, s + lane * 16);
      } else if (lane < nchunks) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst + lane * 16), "r"(0u) : "memory");
      }
    }
    return;
  }
  for (int r = r0; r < r0 + nr; ++r) {
    const uint32_t dst = tile + r * ROWB;
    if (r < cnt) {
      const bf16* s = src + (size_t)(first + r) * ld;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int w = lane + 32 * j;
        if (w < nwords) cp_async4(dst + w * 4, s + 2 * w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int w = lane + 32 * j;
        if (w < nwords) asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + w * 4), "r"(0u) : "memory");
      }
    }
  }
}
// after the copies of a K / V tile have landed: zero the neighbouring heads' words inside the first and last chunk
__device__ __forceinline__ void fix_junk(uint32_t tile, int nwords, int lane, Lead le, int r0 = 0, int nr = 16) {
  if (!le.wide) return;
  const int tail0 = le.lw + nwords, tail1 = (tail0 + 3) & ~3;
  const int per = le.lw + (tail1 - tail0);               // junk words per row (0..6)
  for (int i = lane; i < nr * per; i += 32) {
    const int r = r0 + i / per, j = i % per;
    const int w = j < le.lw ? j : tail0 + (j - le.lw);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile + r * ROWB + w * 4), "r"(0u) : "memory");
  }
}
// coalesced copy of tile rows [0, cnt) (the head's nwords words) to global rows
__device__ __forceinline__ void store_rows(uint32_t tile, bf16* dst, int ld, long long first, int cnt, int nwords, int lane, int r0 = 0,
                                           int rstep = 1, Lead le = Lead{0, false}) {
  if (le.wide) {
    const int w0 = lane * 4;                              // this lane's chunk: tile words w0 .. w0 + 3
    const int lo = le.lw, hi = le.lw + nwords;            // the head's words
    for (int r = r0; r < cnt; r += rstep) {
      uint8_t* d = reinterpret_cast<uint8_t*>(dst + (size_t)(first + r) * ld) - 4 * le.lw;
      if (w0 >= lo && w0 + 4 <= hi) {
        uint32_t v0, v1, v2, v3;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(tile + r * ROWB + w0 * 4));
        *reinterpret_cast<uint4*>(d + w0 * 4) = make_uint4(v0, v1, v2, v3);
      } else if (w0 < hi && w0 + 4 > lo) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int w = w0 + j;
          if (w >= lo && w < hi) {
            uint32_t v;
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(tile + r * ROWB + w * 4));
            *reinterpret_cast<uint32_t*>(d + w * 4) = v;
          }
        }
      }
    }
    return;
  }
  for (int r = r0; r < cnt; r += rstep) {
    bf16* d = dst + (size_t)(first + r) * ld;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int w = lane + 32 * j;
      if (w < nwords) {
        uint32_t v;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(tile + r * ROWB + w * 4));
        *reinterpret_cast<uint32_t*>(d + 2 * w) = v;
      }
    }
  }
}

// C[2][4] (16 rows x 16 cols) = A(16 x 256, row tile) * B(16 x 256, row tile)^T : both tiles row-major in shared memory
__device__ __forceinline__ void qk_product(uint32_t a_tile, uint32_t b_tile, int lane, float (&c)[2][4]) {
  const uint32_t a_addr = a_tile + (lane & 15) * ROWB + (lane >> 4) * 16;
  const uint32_t b_addr = b_tile + ((lane & 7) + ((lane >> 4) << 3)) * ROWB + ((lane >> 3) & 1) * 16;
#pragma unroll
  for (int ks = 0; ks < 16; ++ks) {
    uint32_t a[4], b[4];
    ldsm_x4(a, a_addr + ks * 32);
    ldsm_x4(b, b_addr + ks * 32);
    mma16816(c[0], a, b[0], b[1]);
    mma16816(c[1], a, b[2], b[3]);
  }
}
// acc[nt] (16 rows x 8 cols each, nt in [nt0, nt0 + N)) += A(16 x 16, register fragments) * B(16 rows x 256 cols tile, row-major)
template <int N>
__device__ __forceinline__ void pv_product(const uint32_t (&a)[4], uint32_t b_tile, int lane, int nt0, float (&acc)[N][4]) {
  const uint32_t b_addr = b_tile + ((lane & 7) + ((lane >> 3) & 1) * 8) * ROWB + (lane >> 4) * 16;
#pragma unroll
  for (int i = 0; i < N; i += 2) {
    uint32_t b[4];
    ldsm_x4_t(b, b_addr + (nt0 + i) * 16);
    mma16816(acc[i], a, b[0], b[1]);
    mma16816(acc[i + 1], a, b[2], b[3]);
  }
}
// accumulator tiles -> bf16 tile in shared memory (rows g / g+8, columns nt*8 + 2t, +1), scaled per row
template <int N>
__device__ __forceinline__ void acc_to_tile(uint32_t tile, int lane, int nt0, const float (&acc)[N][4], float s0, float s1) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const uint32_t c = (nt0 + i) * 16 + t * 4;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile + g * ROWB + c), "r"(pack2(acc[i][0] * s0, acc[i][1] * s0)) : "memory");
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile + (g + 8) * ROWB + c), "r"(pack2(acc[i][2] * s1, acc[i][3] * s1)) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// forward: CTA = 4 warps = 4 consecutive heads of one work item
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
attn_fwd_mma_kernel(MArgs a, bf16* __restrict__ o, int ldo, float* __restrict__ lse) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int hgroups = a.heads >> 2;
  const int4 w = a.work[blockIdx.x / hgroups];
  const int h = (blockIdx.x % hgroups) * 4 + warp, col0 = h * a.hd, nwords = a.hd >> 1;
  const long long seg0 = w.x;
  const int L = w.y, q0 = w.z, nq = min(16, L - q0);
  const uint32_t Qs = smem_addr(smem) + warp * 3 * TILE_B, Ks = Qs + TILE_B, Vs = Ks + TILE_B;
  const Lead le = lead_of(a, col0);
  zero_pad(Qs, nwords, lane, le); zero_pad(Ks, nwords, lane, le); zero_pad(Vs, nwords, lane, le);
  load_rows(Qs, a.q + col0, a.ldq, seg0 + q0, nq, 0, 16, nwords, lane, le);

  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  for (int k0 = 0; k0 < L; k0 += 16) {
    const int nk = min(16, L - k0);
    if (k0 > 0) __syncwarp();
    load_rows(Ks, a.k + col0, a.ldk, seg0 + k0, nk, 0, 16, nwords, lane, le);
    load_rows(Vs, a.v + col0, a.ldv, seg0 + k0, nk, 0, 16, nwords, lane, le);
    cp_async_wait_all();
    __syncwarp();
    if (le.wide) { fix_junk(Ks, nwords, lane, le); __syncwarp(); }      // V only feeds output columns here: its junk is never stored
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    qk_product(Qs, Ks, lane, s);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int key = n * 8 + 2 * t + (c & 1);
        s[n][c] = key < nk ? s[n][c] * a.scale : -INFINITY;
        if (c < 2) mx0 = fmaxf(mx0, s[n][c]); else mx1 = fmaxf(mx1, s[n][c]);
      }
    const float mn0 = fmaxf(m0, quad_max(mx0)), mn1 = fmaxf(m1, quad_max(mx1));
    const float corr0 = __expf(m0 - mn0), corr1 = __expf(m1 - mn1);   // first tile: exp(-inf) = 0
    m0 = mn0; m1 = mn1;
    float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      s[n][0] = __expf(s[n][0] - mn0); s[n][1] = __expf(s[n][1] - mn0);
      s[n][2] = __expf(s[n][2] - mn1); s[n][3] = __expf(s[n][3] - mn1);
      ps0 += s[n][0] + s[n][1]; ps1 += s[n][2] + s[n][3];
    }
    l0 = l0 * corr0 + ps0; l1 = l1 * corr1 + ps1;     // per-thread partial row sums; combined over the quad at the end
    if (a.drop.thr16 != 0u) {   // O accumulates the dropped weights; the normaliser l keeps all of them
      // one Philox call per lane = (query row lane & 15, key group lane >> 4); the fragment's rows / groups come by shuffle
      const uint32_t mine = keep8_attn(a.drop, seg0 + q0 + (lane & 15), a.heads, h, (k0 >> 3) + (lane >> 4));
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const uint32_t kp0 = __shfl_sync(0xffffffffu, mine, g + 16 * n) >> (2 * t);
        const uint32_t kp1 = __shfl_sync(0xffffffffu, mine, g + 8 + 16 * n) >> (2 * t);
        s[n][0] = (kp0 & 1u) ? s[n][0] * a.drop.scale : 0.f; s[n][1] = (kp0 & 2u) ? s[n][1] * a.drop.scale : 0.f;
        s[n][2] = (kp1 & 1u) ? s[n][2] * a.drop.scale : 0.f; s[n][3] = (kp1 & 2u) ? s[n][3] * a.drop.scale : 0.f;
      }
    }
    if (k0 > 0) {
#pragma unroll
      for (int i = 0; i < NT; ++i) { acc[i][0] *= corr0; acc[i][1] *= corr0; acc[i][2] *= corr1; acc[i][3] *= corr1; }
    }
    const uint32_t pa[4] = {pack2(s[0][0], s[0][1]), pack2(s[0][2], s[0][3]), pack2(s[1][0], s[1][1]), pack2(s[1][2], s[1][3])};
    pv_product<NT>(pa, Vs, lane, 0, acc);
  }
  l0 = quad_sum(l0); l1 = quad_sum(l1);
  __syncwarp();
  acc_to_tile<NT>(Qs, lane, 0, acc, 1.f / l0, 1.f / l1);   // the Q tile is dead: reuse it to transpose the output for row-contiguous stores
  __syncwarp();
  store_rows(Qs, o + col0, ldo, seg0 + q0, nq, nwords, lane, 0, 1, le);
  if (lse != nullptr && t == 0) {
    if (g < nq) lse[(seg0 + q0 + g) * a.heads + h] = m0 + __logf(l0);
    if (g + 8 < nq) lse[(seg0 + q0 + g + 8) * a.heads + h] = m1 + __logf(l1);
  }
}

// ------------------------------------------------------------------------------------------
// backward, query side: CTA = 2 warps = 2 consecutive heads of one work item; tiles Q, dO (fixed) and K, V (streamed)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
attn_bwd_dq_mma_kernel(MArgs a, const bf16* __restrict__ dout, int lddo, const float* __restrict__ lse, float* __restrict__ delta,
                       bf16* __restrict__ dq, int lddq) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int hgroups = a.heads >> 1;
  const int4 w = a.work[blockIdx.x / hgroups];
  const int h = (blockIdx.x % hgroups) * 2 + warp, col0 = h * a.hd, nwords = a.hd >> 1;
  const long long seg0 = w.x;
  const int L = w.y, q0 = w.z, nq = min(16, L - q0);
  if (a.skip_short && L <= 16) return;
  const uint32_t Qs = smem_addr(smem) + warp * 4 * TILE_B, Gs = Qs + TILE_B, Ks = Gs + TILE_B, Vs = Ks + TILE_B;
  const Lead le = lead_of(a, col0);
  zero_pad(Qs, nwords, lane, le); zero_pad(Gs, nwords, lane, le); zero_pad(Ks, nwords, lane, le); zero_pad(Vs, nwords, lane, le);
  load_rows(Qs, a.q + col0, a.ldq, seg0 + q0, nq, 0, 16, nwords, lane, le);
  load_rows(Gs, dout + col0, lddo, seg0 + q0, nq, 0, 16, nwords, lane, le);
  const long long r0 = seg0 + q0 + g, r1 = r0 + 8;
  const float ls0 = g < nq ? lse[r0 * a.heads + h] : 0.f, ls1 = g + 8 < nq ? lse[r1 * a.heads + h] : 0.f;

  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  // delta_i = sum_j P_ij dP_ij (= dO_i . O_i).  One key tile (the common case): computed on the fly; longer segments take
  // a first pass over the key tiles for delta and a second one for dQ.
  const bool single = L <= 16;
  float dl0 = 0.f, dl1 = 0.f;
  bool first_load = true;
  for (int pass = single ? 1 : 0; pass < 2; ++pass) {
    float e0 = 0.f, e1 = 0.f;
    for (int k0 = 0; k0 < L; k0 += 16) {
      const int nk = min(16, L - k0);
      if (!first_load) __syncwarp();
      first_load = false;
      load_rows(Ks, a.k + col0, a.ldk, seg0 + k0, nk, 0, 16, nwords, lane, le);
      load_rows(Vs, a.v + col0, a.ldv, seg0 + k0, nk, 0, 16, nwords, lane, le);
      cp_async_wait_all();
      __syncwarp();
      if (le.wide) { fix_junk(Ks, nwords, lane, le); fix_junk(Vs, nwords, lane, le); __syncwarp(); }
      float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      qk_product(Qs, Ks, lane, s);
      qk_product(Gs, Vs, lane, dp);
      if (a.drop.thr16 != 0u) {   // dP = mask * d(P_dropped) / (1 - p)
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const uint32_t kp0 = keep8_attn(a.drop, r0, a.heads, h, (k0 >> 3) + n) >> (2 * t);
          const uint32_t kp1 = keep8_attn(a.drop, r1, a.heads, h, (k0 >> 3) + n) >> (2 * t);
          dp[n][0] = (kp0 & 1u) ? dp[n][0] * a.drop.scale : 0.f; dp[n][1] = (kp0 & 2u) ? dp[n][1] * a.drop.scale : 0.f;
          dp[n][2] = (kp1 & 1u) ? dp[n][2] * a.drop.scale : 0.f; dp[n][3] = (kp1 & 2u) ? dp[n][3] * a.drop.scale : 0.f;
        }
      }
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int key = n * 8 + 2 * t + (c & 1);
          s[n][c] = key < nk ? __expf(s[n][c] * a.scale - (c < 2 ? ls0 : ls1)) : 0.f;      // P
          if (c < 2) e0 = fmaf(s[n][c], dp[n][c], e0); else e1 = fmaf(s[n][c], dp[n][c], e1);
        }
      if (pass == 0) continue;
      if (single) { dl0 = quad_sum(e0); dl1 = quad_sum(e1); }
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[n][c] *= dp[n][c] - (c < 2 ? dl0 : dl1);               // dS
      const uint32_t da[4] = {pack2(s[0][0], s[0][1]), pack2(s[0][2], s[0][3]), pack2(s[1][0], s[1][1]), pack2(s[1][2], s[1][3])};
      pv_product<NT>(da, Ks, lane, 0, acc);
    }
    if (pass == 0) { dl0 = quad_sum(e0); dl1 = quad_sum(e1); }
  }
  if (t == 0) {
    if (g < nq) delta[r0 * a.heads + h] = dl0;
    if (g + 8 < nq) delta[r1 * a.heads + h] = dl1;
  }
  __syncwarp();
  acc_to_tile<NT>(Qs, lane, 0, acc, a.scale, a.scale);
  __syncwarp();
  store_rows(Qs, dq + col0, lddq, seg0 + q0, nq, nwords, lane, 0, 1, le);
}

// ------------------------------------------------------------------------------------------
// backward, key side: CTA = 2 warps on ONE (work item, head): the item's 16 rows are KEYS; both warps share the K, V tiles
// (fixed) and the streamed Q, dO tiles, compute S^T / dP^T redundantly, and own one half of the head columns of dK, dV.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64)
attn_bwd_dkv_mma_kernel(MArgs a, const bf16* __restrict__ dout, int lddo, const float* __restrict__ lse, const float* __restrict__ delta,
                        bf16* __restrict__ dk, int lddk, bf16* __restrict__ dv, int lddv) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float lse_s[16], dl_s[16];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int4 w = a.work[blockIdx.x / a.heads];
  const int h = blockIdx.x % a.heads, col0 = h * a.hd, nwords = a.hd >> 1;
  const long long seg0 = w.x;
  const int L = w.y, k0 = w.z, nkeys = min(16, L - k0);
  if (a.skip_short && L <= 16) return;
  const uint32_t Ks = smem_addr(smem), Vs = Ks + TILE_B, Qs = Vs + TILE_B, Gs = Qs + TILE_B;
  // each warp prepares / loads 8 rows of every tile
  const Lead le = lead_of(a, col0);
  if (warp == 0) { zero_pad(Ks, nwords, lane, le); zero_pad(Qs, nwords, lane, le); } else { zero_pad(Vs, nwords, lane, le); zero_pad(Gs, nwords, lane, le); }
  load_rows(Ks, a.k + col0, a.ldk, seg0 + k0, nkeys, warp * 8, 8, nwords, lane, le);
  load_rows(Vs, a.v + col0, a.ldv, seg0 + k0, nkeys, warp * 8, 8, nwords, lane, le);

  constexpr int NH = NT / 2;
  const int nt0 = warp * NH;
  float accK[NH][4], accV[NH][4];
#pragma unroll
  for (int i = 0; i < NH; ++i) { accK[i][0] = accK[i][1] = accK[i][2] = accK[i][3] = 0.f; accV[i][0] = accV[i][1] = accV[i][2] = accV[i][3] = 0.f; }
  for (int q0 = 0; q0 < L; q0 += 16) {
    const int nq = min(16, L - q0);
    if (q0 > 0) __syncthreads();                      // both warps are done with the previous Q / dO tiles
    load_rows(Qs, a.q + col0, a.ldq, seg0 + q0, nq, warp * 8, 8, nwords, lane, le);
    load_rows(Gs, dout + col0, lddo, seg0 + q0, nq, warp * 8, 8, nwords, lane, le);
    if (threadIdx.x < 16) {
      const bool ok = (int)threadIdx.x < nq;
      lse_s[threadIdx.x] = ok ? lse[(seg0 + q0 + threadIdx.x) * a.heads + h] : 0.f;
      dl_s[threadIdx.x] = ok ? delta[(seg0 + q0 + threadIdx.x) * a.heads + h] : 0.f;
    }
    cp_async_wait_all();
    if (le.wide && q0 == 0) { __syncwarp(); fix_junk(Ks, nwords, lane, le, warp * 8, 8); fix_junk(Vs, nwords, lane, le, warp * 8, 8); }   // own rows
    __syncthreads();
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    qk_product(Ks, Qs, lane, s);      // S^T : rows = keys (g, g+8), columns = queries
    qk_product(Vs, Gs, lane, dp);     // dP^T
    float pt[2][4];
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int qi = n * 8 + 2 * t + (c & 1);
        const int key = g + (c >> 1) * 8;
        const float p = (qi < nq && key < nkeys) ? __expf(s[n][c] * a.scale - lse_s[qi]) : 0.f;
        float pd = p, dpv = dp[n][c];
        if (a.drop.thr16 != 0u) {
          const int ka = k0 + key;
          const bool kept = (keep8_attn(a.drop, seg0 + q0 + qi, a.heads, h, ka >> 3) >> (ka & 7)) & 1u;
          pd = kept ? p * a.drop.scale : 0.f;
          dpv = kept ? dpv * a.drop.scale : 0.f;
        }
        pt[n][c] = pd;                                   // dV += P_dropped^T dO
        s[n][c] = p * (dpv - dl_s[qi]);                  // dS^T
      }
    const uint32_t pa[4] = {pack2(pt[0][0], pt[0][1]), pack2(pt[0][2], pt[0][3]), pack2(pt[1][0], pt[1][1]), pack2(pt[1][2], pt[1][3])};
    const uint32_t da[4] = {pack2(s[0][0], s[0][1]), pack2(s[0][2], s[0][3]), pack2(s[1][0], s[1][1]), pack2(s[1][2], s[1][3])};
    pv_product<NH>(pa, Gs, lane, nt0, accV);
    pv_product<NH>(da, Qs, lane, nt0, accK);
  }
  __syncthreads();
  acc_to_tile<NH>(Qs, lane, nt0, accK, a.scale, a.scale);   // Q / dO tiles are dead: transpose dK / dV through them
  acc_to_tile<NH>(Gs, lane, nt0, accV, 1.f, 1.f);
  __syncthreads();
  store_rows(Qs, dk + col0, lddk, seg0 + k0, nkeys, nwords, lane, warp, 2, le);
  store_rows(Gs, dv + col0, lddv, seg0 + k0, nkeys, nwords, lane, warp, 2, le);
}

// ------------------------------------------------------------------------------------------
// backward, segments of at most 16 rows (one tile: every frame and nearly every 2-frame window): ONE kernel reads Q, K, V
// and dO once and writes dQ, dK, dV — the two-kernel form above reads all four tiles twice and passes delta through
// memory.  CTA = 3 warps on one (segment, head) sharing the four tiles; each warp owns one output and recomputes the
// 16 x 16 score products it needs in the orientation that makes its left operand an A fragment:
//   warp 0: S, dP (rows = queries)   -> delta, dS   -> dQ = dS K
//   warp 1: S^T, dP^T (rows = keys)  -> delta, dS^T -> dK = dS^T Q
//   warp 2: S^T                      -> P^T         -> dV = P^T dO
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(96, 3)
attn_bwd_fused_mma_kernel(MArgs a, const bf16* __restrict__ dout, int lddo, const float* __restrict__ lse, bf16* __restrict__ dq, int lddq,
                          bf16* __restrict__ dk, int lddk, bf16* __restrict__ dv, int lddv) {
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ float lse_s[16];
  __shared__ uint8_t keep_s[16][2];                     // [query row][key group]: keep bits of 8 keys
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int4 w = a.work[blockIdx.x / a.heads];
  const int L = w.y;
  if (L > 16) return;                                   // longer segments: the two-kernel path
  const int h = blockIdx.x % a.heads, col0 = h * a.hd, nwords = a.hd >> 1;
  const long long seg0 = w.x;
  const uint32_t Qs = smem_addr(smem), Ks = Qs + TILE_B, Vs = Ks + TILE_B, Gs = Vs + TILE_B;
  // warp w prepares / loads rows [r0, r0 + nr) of every tile
  const int r0 = warp == 0 ? 0 : (warp == 1 ? 6 : 11), nr = warp == 0 ? 6 : 5;
  const Lead le = lead_of(a, col0);
  if (warp == 0) { zero_pad(Qs, nwords, lane, le); zero_pad(Ks, nwords, lane, le); } else if (warp == 1) zero_pad(Vs, nwords, lane, le); else zero_pad(Gs, nwords, lane, le);
  load_rows(Qs, a.q + col0, a.ldq, seg0, L, r0, nr, nwords, lane, le);
  load_rows(Ks, a.k + col0, a.ldk, seg0, L, r0, nr, nwords, lane, le);
  load_rows(Vs, a.v + col0, a.ldv, seg0, L, r0, nr, nwords, lane, le);
  load_rows(Gs, dout + col0, lddo, seg0, L, r0, nr, nwords, lane, le);
  if (threadIdx.x < 16) lse_s[threadIdx.x] = (int)threadIdx.x < L ? lse[(seg0 + threadIdx.x) * a.heads + h] : 0.f;
  // dropout: the keep bits of all 16 x 16 weights, one Philox call per (query row, key group) — 32 calls on warp 2 instead of every
  // thread of every warp regenerating the bits of its own fragment elements (640 calls)
  if (a.drop.thr16 != 0u && warp == 2) keep_s[lane & 15][lane >> 4] = (uint8_t)keep8_attn(a.drop, seg0 + (lane & 15), a.heads, h, lane >> 4);
  cp_async_wait_all();
  if (le.wide) { __syncwarp(); fix_junk(Ks, nwords, lane, le, r0, nr); fix_junk(Vs, nwords, lane, le, r0, nr); }   // own rows
  __syncthreads();

  float acc[NT][4];
#pragma unroll
  for (int i = 0; i < NT; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
  if (warp == 0) {
    // rows = queries g, g + 8; columns = keys n * 8 + 2 t + (c & 1)
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    qk_product(Qs, Ks, lane, s);
    qk_product(Gs, Vs, lane, dp);
    const float ls0 = lse_s[g], ls1 = lse_s[g + 8];
    if (a.drop.thr16 != 0u) {
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const uint32_t kp0 = (uint32_t)keep_s[g][n] >> (2 * t);
        const uint32_t kp1 = (uint32_t)keep_s[g + 8][n] >> (2 * t);
        dp[n][0] = (kp0 & 1u) ? dp[n][0] * a.drop.scale : 0.f; dp[n][1] = (kp0 & 2u) ? dp[n][1] * a.drop.scale : 0.f;
        dp[n][2] = (kp1 & 1u) ? dp[n][2] * a.drop.scale : 0.f; dp[n][3] = (kp1 & 2u) ? dp[n][3] * a.drop.scale : 0.f;
      }
    }
    float e0 = 0.f, e1 = 0.f;
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int key = n * 8 + 2 * t + (c & 1);
        s[n][c] = key < L ? __expf(s[n][c] * a.scale - (c < 2 ? ls0 : ls1)) : 0.f;      // P
        if (c < 2) e0 = fmaf(s[n][c], dp[n][c], e0); else e1 = fmaf(s[n][c], dp[n][c], e1);
      }
    const float dl0 = quad_sum(e0), dl1 = quad_sum(e1);
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[n][c] *= dp[n][c] - (c < 2 ? dl0 : dl1);               // dS
    const uint32_t da[4] = {pack2(s[0][0], s[0][1]), pack2(s[0][2], s[0][3]), pack2(s[1][0], s[1][1]), pack2(s[1][2], s[1][3])};
    pv_product<NT>(da, Ks, lane, 0, acc);
  } else {
    // rows = keys g, g + 8; columns = queries n * 8 + 2 t + (c & 1)
    float s[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, dp[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    qk_product(Ks, Qs, lane, s);
    if (warp == 1) qk_product(Vs, Gs, lane, dp);
    float pd[2][4], e[2][2] = {{0.f, 0.f}, {0.f, 0.f}};      // e[n][c & 1]: per-query partial of sum_keys P dP
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int qi = n * 8 + 2 * t + (c & 1);
        const int key = g + (c >> 1) * 8;
        const float p = (qi < L && key < L) ? __expf(s[n][c] * a.scale - lse_s[qi]) : 0.f;
        float pdv = p, dpv = dp[n][c];
        if (a.drop.thr16 != 0u) {
          const bool kept = ((uint32_t)keep_s[qi][key >> 3] >> (key & 7)) & 1u;
          pdv = kept ? p * a.drop.scale : 0.f;
          dpv = kept ? dpv * a.drop.scale : 0.f;
        }
        pd[n][c] = pdv;                       // dropped, rescaled weights: dV += P_dropped^T dO
        s[n][c] = p;
        dp[n][c] = dpv;
        e[n][c & 1] = fmaf(p, dpv, e[n][c & 1]);
      }
    if (warp == 1) {
      // delta of query qi = sum over the 16 keys: this thread holds keys g and g + 8; the other keys sit in the lanes that
      // differ in g (lane bits 2..4)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          float v = e[n][b];
          v += __shfl_xor_sync(0xffffffffu, v, 4);
          v += __shfl_xor_sync(0xffffffffu, v, 8);
          v += __shfl_xor_sync(0xffffffffu, v, 16);
          e[n][b] = v;
        }
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[n][c] *= dp[n][c] - e[n][c & 1];                       // dS^T
      const uint32_t da[4] = {pack2(s[0][0], s[0][1]), pack2(s[0][2], s[0][3]), pack2(s[1][0], s[1][1]), pack2(s[1][2], s[1][3])};
      pv_product<NT>(da, Qs, lane, 0, acc);
    } else {
      const uint32_t pa[4] = {pack2(pd[0][0], pd[0][1]), pack2(pd[0][2], pd[0][3]), pack2(pd[1][0], pd[1][1]), pack2(pd[1][2], pd[1][3])};
      pv_product<NT>(pa, Gs, lane, 0, acc);
    }
  }
  __syncthreads();                                    // every warp has finished reading the four input tiles
  // each warp transposes its output through a dead tile and stores whole rows
  if (warp == 0) {
    acc_to_tile<NT>(Qs, lane, 0, acc, a.scale, a.scale);
    __syncwarp();
    store_rows(Qs, dq + col0, lddq, seg0, L, nwords, lane, 0, 1, le);
  } else if (warp == 1) {
    acc_to_tile<NT>(Ks, lane, 0, acc, a.scale, a.scale);
    __syncwarp();
    store_rows(Ks, dk + col0, lddk, seg0, L, nwords, lane, 0, 1, le);
  } else {
    acc_to_tile<NT>(Vs, lane, 0, acc, 1.f, 1.f);
    __syncwarp();
    store_rows(Vs, dv + col0, lddv, seg0, L, nwords, lane, 0, 1, le);
  }
}

template <typename K> int opt_in_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) NLV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return NLV_OK;
}

}  // namespace

bool attn_mma_supported(int hd, int heads, int ld_or) {
  // 4-byte copies need even strides / head widths; 256-column tiles; heads in groups of 4 for the forward CTA
  return hd >= 16 && hd <= 256 && (hd & 1) == 0 && (heads & 3) == 0 && (ld_or & 1) == 0;
}

// 16-byte tile copies need every head-0 pointer and every row stride on a 16-byte boundary and a slice of all heads that ends
// on one (then no aligned superset of a head leaves its row's slice).  NLV_ATTN_WIDE=0 forces the 4-byte copies.
static int wide_ok(int hd, int heads, std::initializer_list<const void*> ptrs, std::initializer_list<int> lds) {
  static const int env = [] { const char* e = getenv("NLV_ATTN_WIDE"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
  if (!env || ((hd * heads * 2) & 15) != 0 || ((hd * 2) & 3) != 0) return 0;
  for (const void* p : ptrs) if ((reinterpret_cast<uintptr_t>(p) & 15) != 0) return 0;
  for (int ld : lds) if (((ld * 2) & 15) != 0) return 0;
  return 1;
}

static DropCfg attn_cfg(const nlv_dropout* d) {
  DropCfg c = drop_off();
  if (d != nullptr && d->thr16 != 0u) { c.thr16 = d->thr16; c.scale = d->scale; c.seed_lo = d->seed_lo; c.seed_hi = d->seed_hi; c.stream = d->stream; }
  return c;
}

int launch_attn_fwd_mma(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int hd, int heads, float scale,
                        const void* work, int n_work, void* o, int ldo, float* lse, const nlv_dropout* drop, cudaStream_t s) {
  MArgs a{(const bf16*)q, (const bf16*)k, (const bf16*)v, ldq, ldk, ldv, hd, heads, scale, (const int4*)work, attn_cfg(drop), 0,
          wide_ok(hd, heads, {q, k, v, o}, {ldq, ldk, ldv, ldo})};
  const size_t smem = 4 * 3 * TILE_B;
  int rc = opt_in_smem(attn_fwd_mma_kernel, smem);
  if (rc != NLV_OK) return rc;
  attn_fwd_mma_kernel<<<(unsigned)n_work * (heads / 4), 128, smem, s>>>(a, (bf16*)o, ldo, lse);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

int launch_attn_bwd_mma(const void* q, int ldq, const void* k, int ldk, const void* v, int ldv, int hd, int heads, float scale,
                        const void* work, int n_work, int n_long, const void* dout, int lddo, const float* lse, float* delta, void* dq, int lddq, void* dk, int lddk, void* dv, int lddv, const nlv_dropout* drop, cudaStream_t s) {
  // NLV_ATTN_BWD_FUSED=0: every segment through the two-kernel path (debugging switch)
  static const int fused = [] { const char* e = getenv("NLV_ATTN_BWD_FUSED"); return (e != nullptr && e[0] == '0') ? 0 : 1; }();
  MArgs a{(const bf16*)q, (const bf16*)k, (const bf16*)v, ldq, ldk, ldv, hd, heads, scale, (const int4*)work, attn_cfg(drop), fused,
          wide_ok(hd, heads, {q, k, v, dout, dq, dk, dv}, {ldq, ldk, ldv, lddo, lddq, lddk, lddv})};
  const size_t s1 = 2 * 4 * TILE_B, s2 = 4 * TILE_B;
  int rc = opt_in_smem(attn_bwd_dq_mma_kernel, s1);
  if (rc != NLV_OK) return rc;
  if (fused) {
    MArgs af = a;
    af.wide = 0;      // measured: the fused kernel is 8% faster with the 4-byte copies (303 vs 328 us on the C2 decoder shape), the forward 6% slower
    attn_bwd_fused_mma_kernel<<<(unsigned)n_work * heads, 96, s2, s>>>(af, (const bf16*)dout, lddo, lse, (bf16*)dq, lddq, (bf16*)dk, lddk, (bf16*)dv, lddv);
    NLV_CHECK_LAUNCH();
  }
  // the two-kernel path only has work for segments of more than 16 rows: with a long-first work list (n_long >= 0, plan.py) it
  // is launched over those items alone — usually none — instead of over every item to exit at once
  const int n_two = (fused && n_long >= 0) ? (n_long < n_work ? n_long : n_work) : n_work;
  if (n_two == 0) return NLV_OK;
  attn_bwd_dq_mma_kernel<<<(unsigned)n_two * (heads / 2), 64, s1, s>>>(a, (const bf16*)dout, lddo, lse, delta, (bf16*)dq, lddq);
  NLV_CHECK_LAUNCH();
  attn_bwd_dkv_mma_kernel<<<(unsigned)n_two * heads, 64, s2, s>>>(a, (const bf16*)dout, lddo, lse, delta, (bf16*)dk, lddk, (bf16*)dv, lddv);
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

}  // namespace nlv
