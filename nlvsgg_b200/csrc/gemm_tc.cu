// tcgen05 / TMEM / TMA GEMM for sm_100a.
//
//   D[m,n] = act( sum_k A(m,k) * B(n,k) + bias[n] ) + residual[m,n]       (bf16 in, fp32 accumulate)
//
// One persistent CTA per SM, 320 threads:
//   warp 0 (lane 0) : TMA producer   — cp.async.bulk.tensor 2-D boxes into a 128B-swizzled smem ring
//   warp 1 (lane 0) : MMA issuer     — tcgen05.mma.kind::f16, accumulators double-buffered in TMEM (2 x BLOCK_N columns).
//                                      MODE 2 (default when m > 128): the two CTAs of a cluster form a CTA PAIR and the
//                                      leader issues tcgen05.mma.cta_group::2 — one 256 x BLOCK_N x 16 instruction over
//                                      both SMs, each CTA staging its own 128 A rows and HALF of the B tile (6-8 stages
//                                      instead of 4-6, half the B shared-memory reads per SM).  MODE 1: cta_group::1 with
//                                      the B tile multicast to both CTAs.  MODE 0: single CTA, 128 x BLOCK_N x 16.
//   warps 2..9      : epilogue       — tcgen05.ld 32x32b (one TMEM lane = one output row per thread); warps w and w+4
//                                      share a TMEM lane quarter and take alternate 32-column chunks (the epilogue of
//                                      small-K products is bound by memory instructions in flight per warp);
//                                      bias / ReLU / residual fused, fp32 or bf16 stores
// Both operands may be K-major ([MN,K] row-major) or MN-major ([K,MN] row-major), selected in the UMMA
// instruction/smem descriptors, so the forward (X W^T), input-gradient (dY W) and weight-gradient
// (dY^T X) products of a Linear layer all run here without transposed copies.
//
// Replaces (through cuBLAS) nn.Linear / MHA projections / 1x1 conv of lib/sttran.py:336-348,370-372 and
// lib/transformer.py:9-13,38-42.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "philox.cuh"

namespace nlv {
namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int NUM_EPI_WARPS = 8;               // two warps per TMEM lane quarter, alternating 32-column chunks
constexpr int NUM_THREADS = 64 + 32 * NUM_EPI_WARPS;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;  // 16 KiB
constexpr int ATOM_BYTES = 64 * BLOCK_K * 2;          // one [64 mn x 64 k] MN-major box, 8 KiB

template <int BLOCK_N, int MODE>
struct Cfg {
  static constexpr int B_ROWS = MODE == 2 ? BLOCK_N / 2 : BLOCK_N;     // B rows staged per CTA
  static constexpr int B_STAGE_BYTES = B_ROWS * BLOCK_K * 2;
  static constexpr int STAGES = MODE == 2 ? ((BLOCK_N == 256) ? 6 : 8) : ((BLOCK_N == 256) ? 4 : 6);
  static constexpr int ACC_STAGES = 512 / BLOCK_N;   // accumulator stages in TMEM: 2 x 256 or 4 x 128 columns (all 512)
  static constexpr int TMEM_COLS = ACC_STAGES * BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * (A_STAGE_BYTES + B_STAGE_BYTES) + 256 /*barriers*/ + NUM_EPI_WARPS * 4096 /*epilogue staging*/ +
                                    1024 /*align*/;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU — trap after ~4 s instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// multicast variant: the box lands at the same shared-memory offset in every CTA of `mask`, and completes bytes on the
// mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%4, %5}], [%2], %3;"
      ::"r"(dst), "l"(tmap), "r"(bar), "h"(mask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// ---- CTA pair (cta_group::2) ----
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr, uint32_t rank) {   // same offset in CTA `rank` of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
// the box lands in THIS CTA's shared memory, its bytes complete on the LEADER's mbarrier (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tmap, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
// 4-D boxes [channels, x, y, pair] of an NHWC map: the A operand of the implicit 3x3 convolution products.  Coordinates are signed;
// elements outside the map (the halo of a shifted tap) are zero-filled by the TMA unit.
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(uint32_t dst, const CUtensorMap* tmap, uint32_t leader_bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 8000000000LL) __trap();
  }
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar, uint16_t mask) {   // arrives on `bar`'s offset in every CTA of mask
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (cute::UMMA::SmemDescriptor bit layout, SWIZZLE_128B, version 1).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);            // start address  [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;       // leading byte offset [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;       // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                                 // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                                 // layout type: SWIZZLE_128B
  return d;
}

struct Params {
  void* d;
  const float* bias;
  const void* residual;
  int m, n, k;
  int ldd, ldr;
  int d_dtype, r_dtype;
  int relu;
  const void* gate;
  int ldg, gate_dtype;
  float gate_scale;
  DropCfg drop;
  int num_m_blocks, num_n_blocks;
  int k_splits, kb_per_split;  // split-K: work item = (tile, k range); partial sums are atomically added to a zeroed fp32 D
  // implicit 3x3 convolution (conv_c > 0; A K-major only): A is an NHWC map [pairs, 7, 7, conv_c]; an m-tile is TWO pairs = 98 rows
  // of the 128-row instruction (rows 98.. compute on stale shared memory and are never stored); k-block kb = tap kb / (conv_c / 64),
  // channels (kb % (conv_c / 64)) * 64, loaded as the 4-D box [64, 7, 7, 2] at (c0, 1 - kx, 1 - ky, pair) (conv_flip) or
  // (c0, kx - 1, ky - 1, pair).  tile_rows = rows of D per m-tile (128, or 98), a_tx = bytes of one A box.
  int conv_c, conv_flip, tile_rows;
  uint32_t a_tx;
};

// ---------------------------------------------------------------------------------------------
// Kernel
// ---------------------------------------------------------------------------------------------
// CM = CTAs per cluster along M (1 or 2).  With CM == 2 the two CTAs compute vertically adjacent 128 x BLOCK_N tiles and
// share the B tile: each loads half of it and multicasts it into both shared memories, cutting the L2->SM operand
// traffic per tile from (128 + BLOCK_N) to (128 + BLOCK_N/2) rows per k-block.
template <int BLOCK_N, bool A_MN, bool B_MN, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Params p) {
  using C = Cfg<BLOCK_N, MODE>;
  constexpr int CM = MODE == 0 ? 1 : 2;          // CTAs per cluster
  constexpr bool PAIR = MODE == 2;               // one cta_group::2 MMA over both CTAs
  constexpr int STAGES = C::STAGES;
  const uint32_t STAGE_TX = (p.a_tx + C::B_STAGE_BYTES) * (PAIR ? 2 : 1);   // pair: both CTAs' boxes complete on the leader's barrier

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a = smem_base;
  const uint32_t smem_b = smem_base + STAGES * A_STAGE_BYTES;
  const uint32_t bar_base = smem_b + STAGES * C::B_STAGE_BYTES;
  // barrier layout: full[STAGES], empty[STAGES], tmem_full[4], tmem_empty[4], then the TMEM base address word
  constexpr int ACC = C::ACC_STAGES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
  auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 4 + s); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), MODE == 1 ? 2 : 1);   // MODE 1: one tcgen05.commit arrival per CTA that wrote into this stage
    }
    for (int s = 0; s < ACC; ++s) {
      mbar_init(tmem_full_bar(s), 1);
      mbar_init(tmem_empty_bar(s), NUM_EPI_WARPS * (PAIR ? 2 : 1));  // one arrive per epilogue warp (pair: of both CTAs, on the leader)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_b) : "memory");
  }
  if (warp == 1) {  // whole warp: allocate TMEM columns (pair: the same columns in both CTAs)
    if constexpr (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)C::TMEM_COLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CM > 1) cluster_sync_all();   // peers' barriers are initialised before any remote arrival / multicast
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  const int num_k_blocks = (p.k + BLOCK_K - 1) / BLOCK_K;
  const int cta_rank = CM > 1 ? (int)cluster_ctarank() : 0;
  // the last n-block issues a narrower instruction (N rounded up to 16; 32 for a pair so that each CTA's half stays a multiple
  // of 16) instead of multiplying zero-filled columns: n = 1936 = 7 x 256 + 144 would otherwise spend 5.5% of its tensor
  // time on padding.  In pair mode CTA r stages the B rows [n0 + r * n_inst / 2, ...) — the halves of the instruction's N.
  auto n_inst_of = [&](int n0) -> int {
    const int n_left = p.n - n0;
    if (n_left >= BLOCK_N) return BLOCK_N;
    return PAIR ? ((n_left + 31) & ~31) : ((n_left + 15) & ~15);
  };
  const int num_mp = (p.num_m_blocks + CM - 1) / CM;          // groups of CM vertically adjacent tiles
  const int num_work = num_mp * p.num_n_blocks * p.k_splits;
  const int work0 = blockIdx.x / CM, work_stride = gridDim.x / CM;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int work = work0; work < num_work; work += work_stride) {
        const int tile = work / p.k_splits;
        const int kb0 = (work % p.k_splits) * p.kb_per_split;
        const int kb1 = min(num_k_blocks, kb0 + p.kb_per_split);
        // n-block fastest: the ~74 clusters running at any moment cover a band of ~9 m-pairs x all n-blocks, so an A
        // row block is fetched from DRAM once and re-read from L2 by the other n-blocks (m-fastest order streamed all
        // of A once per n-block: 3x the algorithmic DRAM traffic on the [22931 x 1936 x 1936] products)
        const int m0 = ((tile / p.num_n_blocks) * CM + cta_rank) * p.tile_rows;   // may lie beyond M for the odd tile of a pair: loads zero-fill, stores are skipped
        const int n0 = (tile % p.num_n_blocks) * BLOCK_N;
        const int nb0 = PAIR ? n0 + cta_rank * (n_inst_of(n0) / 2) : n0;     // first B row this CTA stages (pair mode)
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1u);
          const int k0 = kb * BLOCK_K;
          const uint32_t sa = smem_a + stage * A_STAGE_BYTES;
          const uint32_t sb = smem_b + stage * C::B_STAGE_BYTES;
          if constexpr (PAIR) {
            // both CTAs' boxes complete on the leader's full barrier; only the leader posts the expected byte count
            const uint32_t lbar = map_to_cta(full_bar(stage), 0);
            if (cta_rank == 0) mbar_arrive_expect_tx(full_bar(stage), STAGE_TX);
            if constexpr (A_MN) {
#pragma unroll
              for (int i = 0; i < BLOCK_M / 64; ++i) tma_load_2d_pair(sa + i * ATOM_BYTES, &tmap_a, lbar, m0 + 64 * i, k0);
            } else if (p.conv_c > 0) {
              const int kpt = p.conv_c >> 6, tap = kb / kpt, ky = tap / 3, kx = tap - 3 * ky;
              tma_load_4d_pair(sa, &tmap_a, lbar, (kb - tap * kpt) << 6, p.conv_flip ? 1 - kx : kx - 1, p.conv_flip ? 1 - ky : ky - 1, m0 / 49);
            } else {
              tma_load_2d_pair(sa, &tmap_a, lbar, k0, m0);
            }
            if constexpr (B_MN) {
#pragma unroll
              for (int i = 0; i < C::B_ROWS / 64; ++i) tma_load_2d_pair(sb + i * ATOM_BYTES, &tmap_b, lbar, nb0 + 64 * i, k0);
            } else {
              tma_load_2d_pair(sb, &tmap_b, lbar, k0, nb0);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
            continue;
          }
          mbar_arrive_expect_tx(full_bar(stage), STAGE_TX);
          if constexpr (A_MN) {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 64; ++i) tma_load_2d(sa + i * ATOM_BYTES, &tmap_a, full_bar(stage), m0 + 64 * i, k0);
          } else if (p.conv_c > 0) {
            const int kpt = p.conv_c >> 6, tap = kb / kpt, ky = tap / 3, kx = tap - 3 * ky;
            tma_load_4d(sa, &tmap_a, full_bar(stage), (kb - tap * kpt) << 6, p.conv_flip ? 1 - kx : kx - 1, p.conv_flip ? 1 - ky : ky - 1, m0 / 49);
          } else {
            tma_load_2d(sa, &tmap_a, full_bar(stage), k0, m0);
          }
          if constexpr (CM == 1) {
            if constexpr (B_MN) {
#pragma unroll
              for (int i = 0; i < BLOCK_N / 64; ++i) tma_load_2d(sb + i * ATOM_BYTES, &tmap_b, full_bar(stage), n0 + 64 * i, k0);
            } else {
              tma_load_2d(sb, &tmap_b, full_bar(stage), k0, n0);
            }
          } else {
            // this CTA fetches its half of the shared B tile and multicasts it to both CTAs of the pair
            constexpr uint16_t kMask = (1u << CM) - 1;
            if constexpr (B_MN) {
              constexpr int kPer = BLOCK_N / 64 / CM;
#pragma unroll
              for (int i = 0; i < kPer; ++i) {
                const int a = cta_rank * kPer + i;
                tma_load_2d_mc(sb + a * ATOM_BYTES, &tmap_b, full_bar(stage), n0 + 64 * a, k0, kMask);
              }
            } else {
              constexpr int kRows = BLOCK_N / CM;
              tma_load_2d_mc(sb + cta_rank * kRows * (BLOCK_K * 2), &tmap_b, full_bar(stage), k0, n0 + cta_rank * kRows, kMask);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && (!PAIR || cta_rank == 0)) {
      // ===================== MMA issuer (pair: the leader CTA only) =====================
      // instruction descriptor: D=f32, A=B=bf16, majors, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc0 = (1u << 4) | (1u << 7) | (1u << 10) | ((A_MN ? 1u : 0u) << 15) | ((B_MN ? 1u : 0u) << 16) |
                              ((uint32_t)((PAIR ? 2 * BLOCK_M : BLOCK_M) >> 4) << 24);
      // K-major  : 8-row groups 1024 B apart (SBO); 16 k-elements = +32 B on the start address
      // MN-major : 64-mn atoms 8 KiB apart (LBO), 8-k groups 1024 B apart (SBO); 16 k-elements = +2048 B
      constexpr uint32_t A_LBO = A_MN ? ATOM_BYTES : 16, A_KSTEP = A_MN ? 2048 : 32;
      constexpr uint32_t B_LBO = B_MN ? ATOM_BYTES : 16, B_KSTEP = B_MN ? 2048 : 32;
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int work = work0; work < num_work; work += work_stride, ++iter) {
        const int kb0 = (work % p.k_splits) * p.kb_per_split;
        const int kb1 = min(num_k_blocks, kb0 + p.kb_per_split);
        const uint32_t n_inst = (uint32_t)n_inst_of(((work / p.k_splits) % p.num_n_blocks) * BLOCK_N);
        const uint32_t idesc = idesc0 | ((n_inst >> 3) << 17);
        const int acc = iter % ACC;
        const uint32_t acc_phase = (iter / ACC) & 1;
        if constexpr (PAIR) mbar_wait_cluster(tmem_empty_bar(acc), acc_phase ^ 1u);   // arrivals come from both CTAs' epilogues
        else mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(full_bar(stage), phase);
          tc_fence_after();
          const uint32_t sa = smem_a + stage * A_STAGE_BYTES;
          const uint32_t sb = smem_b + stage * C::B_STAGE_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * A_KSTEP, A_LBO, 1024);
            const uint64_t bdesc = make_smem_desc(sb + k * B_KSTEP, B_LBO, 1024);
            if constexpr (PAIR) tc_mma_bf16_pair(tmem_d, adesc, bdesc, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
            else tc_mma_bf16(tmem_d, adesc, bdesc, idesc, (kb != kb0 || k != 0) ? 1u : 0u);
          }
          // frees the smem slot when these MMAs retire — in BOTH CTAs of a cluster (MODE 1: each also wrote the peer's copy;
          // pair: the instruction read both CTAs' stages); the accumulator-ready signal of a pair goes to both epilogues
          if constexpr (CM == 1) tc_commit(empty_bar(stage));
          else if constexpr (PAIR) tc_commit_pair(empty_bar(stage), (uint16_t)3);
          else tc_commit_mc(empty_bar(stage), (uint16_t)((1u << CM) - 1));
          if (kb == kb1 - 1) {
            if constexpr (PAIR) tc_commit_pair(tmem_full_bar(acc), (uint16_t)3);
            else tc_commit(tmem_full_bar(acc));
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int chunk0 = ((warp - 2) >> 2) * 32;   // first 32-column chunk of this warp
    int iter = 0;
    const bool d_bf16 = p.d_dtype == NLV_BF16;
    const bool vec_ok = d_bf16 ? ((p.ldd & 7) == 0 && ((uintptr_t)p.d & 15) == 0)
                               : ((p.ldd & 3) == 0 && ((uintptr_t)p.d & 15) == 0);
    const bool split = p.k_splits > 1;
    const bool bias_vec = ((uintptr_t)p.bias & 15) == 0;
    const bool res_fast = p.residual != nullptr && p.r_dtype != NLV_BF16 && !split && !d_bf16 && vec_ok && (p.ldr & 3) == 0 &&
                          ((uintptr_t)p.residual & 15) == 0;
    // bf16 residual rows are requested one chunk ahead too (same registers), each thread its own row
    const bool res_b16 = p.residual != nullptr && p.r_dtype == NLV_BF16 && !split && (p.ldr & 7) == 0 && ((uintptr_t)p.residual & 15) == 0 &&
                         (p.n & 7) == 0;
    uint8_t* const stg = smem_raw + (bar_base + 256u - smem_u32(smem_raw)) + (warp - 2) * 4096;   // this warp's staging tile
    for (int work = work0; work < num_work; work += work_stride, ++iter) {
      const int tile = work / p.k_splits;
      const int m0 = ((tile / p.num_n_blocks) * CM + cta_rank) * p.tile_rows;   // may lie beyond M for the odd tile of a pair: loads zero-fill, stores are skipped
      const int n0 = (tile % p.num_n_blocks) * BLOCK_N;
      const int acc = iter % ACC;
      const uint32_t acc_phase = (iter / ACC) & 1;
      const int row = m0 + quarter * 32 + lane;
      const int rows_here = min(p.tile_rows, p.m - m0) - quarter * 32;      // rows of this warp's quarter that exist in D (may be <= 0)
      const bool row_ok = lane < rows_here;
      constexpr int CSTEP = 32 * (NUM_EPI_WARPS / 4);
      // fp32 residual rows are requested one chunk AHEAD of their use — the first chunk's before the accumulator is even
      // ready — so that the DRAM latency of the residual stream overlaps the MMAs / the previous chunk's stores instead of
      // being paid once per chunk by every epilogue warp
      float4 rcur[8];
      // (mapping of the transposed store below: register `it` = row it * 4 + lane / 8 of the warp's 32, columns 4 (lane % 8) ..)
      auto prefetch_res = [&](int c0, float4 (&dst)[8]) -> bool {
        if (res_b16) {      // bf16 residual (the mask-branch map under the union 1x1 conv): this thread's own row, 32 columns = 64 bytes
          if (c0 >= BLOCK_N || n0 + c0 + 32 > p.n) return false;                 // warp-uniform
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + (size_t)row * p.ldr + n0 + c0);
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const uint4 t = row_ok ? src[it] : make_uint4(0u, 0u, 0u, 0u);
            dst[it] = make_float4(__uint_as_float(t.x), __uint_as_float(t.y), __uint_as_float(t.z), __uint_as_float(t.w));
          }
          return true;
        }
        if (!res_fast || c0 >= BLOCK_N || n0 + c0 + 32 > p.n) return false;      // warp-uniform
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int grow = m0 + quarter * 32 + it * 4 + (lane >> 3);
          dst[it] = (it * 4 + (lane >> 3) < rows_here) ? *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + (size_t)grow * p.ldr + n0 + c0 + (lane & 7) * 4)
                               : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        return true;
      };
      bool have_res = prefetch_res(chunk0, rcur);
      mbar_wait(tmem_full_bar(acc), acc_phase);
      tc_fence_after();
      const uint32_t taddr_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N;
      bool released = false;
#pragma unroll 1
      for (int c0 = chunk0; c0 < BLOCK_N; c0 += CSTEP) {
        if (n0 + c0 >= p.n) break;  // warp-uniform
        uint32_t v[32];
        tmem_ld32(taddr_row + c0, v);
        float4 rnext[8];
        const bool have_next = prefetch_res(c0 + CSTEP, rnext);
        const bool have_now = have_res;
        float4 rnow[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { rnow[j] = rcur[j]; rcur[j] = rnext[j]; }
        have_res = have_next;
        tmem_ld_wait();
        if (c0 + CSTEP >= BLOCK_N || n0 + c0 + CSTEP >= p.n) {
          // this warp's last chunk of the tile: its share of the accumulator stage goes back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (PAIR) mbar_arrive_cluster(map_to_cta(tmem_empty_bar(acc), 0));   // the leader's MMA warp owns both halves
            else mbar_arrive(tmem_empty_bar(acc));
          }
          released = true;
        }
        const int ncol = min(32, p.n - (n0 + c0));
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = row_ok ? __uint_as_float(v[j]) : 0.f;
        if (split) {  // partial sum of one k range: accumulate into the zero-initialised fp32 output
          if (row_ok) {
            float* o = reinterpret_cast<float*>(p.d) + (size_t)row * p.ldd + n0 + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) atomicAdd(o + j, f[j]);
          }
          continue;
        }
        if (p.bias != nullptr) {
          if (ncol == 32 && bias_vec) {   // same address in every lane: 8 broadcast 16-byte loads instead of 32 scalar ones
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 t = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
              f[j] += t.x; f[j + 1] += t.y; f[j + 2] += t.z; f[j + 3] += t.w;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) f[j] += __ldg(p.bias + n0 + c0 + j);
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        if (p.drop.thr16 != 0u) {   // dropout of the activated value (before the residual); groups of 8 columns share one Philox call
          const int gpr = (p.n + 7) >> 3;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            const uint32_t keep = keep8_matrix(p.drop, row, (n0 + c0 + j) >> 3, gpr);
#pragma unroll
            for (int q = 0; q < 8; ++q) f[j + q] = ((keep >> q) & 1u) ? f[j + q] * p.drop.scale : 0.f;
          }
        }
        if (p.gate != nullptr && row_ok) {
          if (p.gate_scale != 1.f) {
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] *= p.gate_scale;
          }
          const size_t goff = (size_t)row * p.ldg + n0 + c0;
          if (p.gate_dtype == NLV_BF16) {
            const __nv_bfloat16* gt = reinterpret_cast<const __nv_bfloat16*>(p.gate) + goff;
            if (ncol == 32 && (p.ldg & 7) == 0 && ((uintptr_t)p.gate & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                const uint4 t = *reinterpret_cast<const uint4*>(gt + j);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float2 g2 = __bfloat1622float2(h[q]);
                  if (!(g2.x > 0.f)) f[j + 2 * q] = 0.f;
                  if (!(g2.y > 0.f)) f[j + 2 * q + 1] = 0.f;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < ncol && !(__bfloat162float(gt[j]) > 0.f)) f[j] = 0.f;
            }
          } else {
            const float* gt = reinterpret_cast<const float*>(p.gate) + goff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol && !(gt[j] > 0.f)) f[j] = 0.f;
          }
        }
        // full, aligned chunks leave through the warp's 4 KB staging tile: the accumulator arrives one ROW per thread, a
        // row-per-thread store touches 32 rows x 16 bytes per instruction (half-filled sectors); transposed through shared
        // memory (16-byte slots XOR-swizzled by the row: conflict-free both ways) every store / residual load instruction
        // covers whole 128-byte (fp32) or 64-byte (bf16) row segments
        const bool fast = ncol == 32 && vec_ok;   // warp-uniform
        const bool res_late = fast && !d_bf16 && res_fast;          // fp32 residual added after the transposition (coalesced)
        if (p.residual != nullptr && !res_late && row_ok) {
          const size_t roff = (size_t)row * p.ldr + n0 + c0;
          if (p.r_dtype == NLV_BF16 && res_b16 && have_now) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const uint32_t wv[4] = {__float_as_uint(rnow[it].x), __float_as_uint(rnow[it].y), __float_as_uint(rnow[it].z), __float_as_uint(rnow[it].w)};
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wv[q]));
                f[it * 8 + 2 * q] += t.x; f[it * 8 + 2 * q + 1] += t.y;
              }
            }
          } else if (p.r_dtype == NLV_BF16) {
            const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(p.residual) + roff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) f[j] += __bfloat162float(r[j]);
          } else {
            const float* r = reinterpret_cast<const float*>(p.residual) + roff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) f[j] += r[j];
          }
        }
        if (fast && !d_bf16) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
          __syncwarp();
          const int sub = lane >> 3, piece = lane & 7;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + sub;
            float4 o = *reinterpret_cast<const float4*>(stg + rr * 128 + ((piece ^ (rr & 7)) << 4));
            const int grow = m0 + quarter * 32 + rr;
            if (rr < rows_here) {
              if (res_late) {
                float4 t;
                if (have_now) t = rnow[it];
                else t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + (size_t)grow * p.ldr + n0 + c0 + piece * 4);
                o.x += t.x; o.y += t.y; o.z += t.z; o.w += t.w;
              }
              *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.d) + (size_t)grow * p.ldd + n0 + c0 + piece * 4) = o;
            }
          }
          __syncwarp();
        } else if (fast) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint4 t;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(f[8 * j], f[8 * j + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(f[8 * j + 2], f[8 * j + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[8 * j + 4], f[8 * j + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(f[8 * j + 6], f[8 * j + 7]);
            t.x = *reinterpret_cast<uint32_t*>(&h0); t.y = *reinterpret_cast<uint32_t*>(&h1);
            t.z = *reinterpret_cast<uint32_t*>(&h2); t.w = *reinterpret_cast<uint32_t*>(&h3);
            *reinterpret_cast<uint4*>(stg + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4)) = t;
          }
          __syncwarp();
          const int sub = lane >> 2, piece = lane & 3;
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            const int rr = it * 8 + sub;
            const uint4 o = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((piece ^ ((rr >> 1) & 3)) << 4));
            const int grow = m0 + quarter * 32 + rr;
            if (rr < rows_here) *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.d) + (size_t)grow * p.ldd + n0 + c0 + piece * 8) = o;
          }
          __syncwarp();
        } else if (row_ok) {
          const size_t doff = (size_t)row * p.ldd + n0 + c0;
          if (d_bf16) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.d) + doff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) o[j] = __float2bfloat16_rn(f[j]);
          } else {
            float* o = reinterpret_cast<float*>(p.d) + doff;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < ncol) o[j] = f[j];
          }
        }
      }
      if (!released) {   // no column chunk fell to this warp (narrow last n-block)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster(map_to_cta(tmem_empty_bar(acc), 0));
          else mbar_arrive(tmem_empty_bar(acc));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (CM > 1) cluster_sync_all();   // no CTA exits while its peer may still multicast into it
  if (warp == 1) {
    if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)C::TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// Tensor map over a bf16 row-major matrix [rows, cols] (cols contiguous, row stride ld elements).
int make_tmap(CUtensorMap* tm, const void* ptr, long long rows, long long cols, long long ld, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return NLV_ERR_CUDA;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld box=%dx%d ptr=%p", (int)r, rows, cols, ld,
              box_cols, box_rows, ptr);
    return NLV_ERR_CUDA;
  }
  return NLV_OK;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int MODE>
int run_kernel(const CUtensorMap& ta, const CUtensorMap& tb, const Params& p, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, MODE>;
  constexpr int CM = MODE == 0 ? 1 : 2;
  auto kern = gemm_tc_kernel<BLOCK_N, A_MN, B_MN, MODE>;
  static bool attr_set = false;
  if (!attr_set) {
    NLV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    attr_set = true;
  }
  const int groups = cdiv(p.num_m_blocks, CM) * p.num_n_blocks * p.k_splits;
  const int max_groups = sm_count() / CM;
  const int grid = (groups < max_groups ? groups : max_groups) * CM;
  if (CM == 1) {
    kern<<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CM; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NLV_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  }
  NLV_CHECK_LAUNCH();
  return NLV_OK;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int MODE>
int launch(const nlv_gemm_args& g, cudaStream_t stream) {
  using C = Cfg<BLOCK_N, MODE>;
  constexpr int CM = MODE == 0 ? 1 : 2;
  CUtensorMap ta, tb;
  int rc;
  if (A_MN) rc = make_tmap(&ta, g.a, g.k, g.m, g.lda, 64, BLOCK_K);
  else      rc = make_tmap(&ta, g.a, g.m, g.k, g.lda, BLOCK_K, BLOCK_M);
  if (rc != NLV_OK) return rc;
  if (B_MN) rc = make_tmap(&tb, g.b, g.k, g.n, g.ldb, 64, BLOCK_K);
  else      rc = make_tmap(&tb, g.b, g.n, g.k, g.ldb, BLOCK_K, BLOCK_N / CM);   // MODE 1: the half this CTA multicasts; MODE 2: the half it stages
  if (rc != NLV_OK) return rc;
  Params p;
  p.d = g.d; p.bias = g.bias; p.residual = g.residual;
  p.m = g.m; p.n = g.n; p.k = g.k; p.ldd = g.ldd; p.ldr = g.ldr;
  p.d_dtype = g.d_dtype; p.r_dtype = g.r_dtype; p.relu = g.relu;
  p.gate = g.gate; p.ldg = g.ldg; p.gate_dtype = g.gate_dtype;
  p.gate_scale = g.gate_scale == 0.f ? 1.f : g.gate_scale;
  p.drop.thr16 = g.drop.thr16; p.drop.scale = g.drop.scale; p.drop.seed_lo = g.drop.seed_lo; p.drop.seed_hi = g.drop.seed_hi;
  p.drop.stream = g.drop.stream;
  p.num_m_blocks = cdiv(g.m, BLOCK_M);
  p.num_n_blocks = cdiv(g.n, BLOCK_N);
  p.conv_c = 0; p.conv_flip = 0; p.tile_rows = BLOCK_M; p.a_tx = A_STAGE_BYTES;
  // split-K when the output has too few tiles to fill the GPU and the reduction is long (weight gradients of the
  // conv / union layers reduce over R*49..R*196 rows into one or a handful of tiles)
  const int nkb = cdiv(g.k, BLOCK_K);
  const int tiles0 = p.num_m_blocks * p.num_n_blocks;
  p.k_splits = 1;
  p.kb_per_split = nkb;
  if (tiles0 * 2 <= sm_count() && nkb >= 32 && g.d_dtype == NLV_F32 && g.bias == nullptr && g.residual == nullptr && !g.relu && g.gate == nullptr &&
      g.drop.thr16 == 0u) {
    int want = sm_count() / tiles0;
    if (want > nkb / 8) want = nkb / 8;
    if (want > 1) {
      p.kb_per_split = cdiv(nkb, want);
      p.k_splits = cdiv(nkb, p.kb_per_split);
      { int zrc = zero_fill(reinterpret_cast<float*>(g.d), g.m, g.n, g.ldd, stream); if (zrc != NLV_OK) return zrc; }
    }
  }
  return run_kernel<BLOCK_N, A_MN, B_MN, MODE>(ta, tb, p, stream);
}

// 4-D tensor map over an NHWC bf16 map [pairs, 7, 7, c]: box = 64 channels x 7 x 7 x 2 pairs (98 rows of 128 bytes, 128B-swizzled)
int make_tmap_nhwc7(CUtensorMap* tm, const void* ptr, long long pairs, int c) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled entry point unavailable");
    return NLV_ERR_CUDA;
  }
  cuuint64_t gdim[4] = {(cuuint64_t)c, 7, 7, (cuuint64_t)pairs};
  cuuint64_t gstride[3] = {(cuuint64_t)c * 2, (cuuint64_t)c * 2 * 7, (cuuint64_t)c * 2 * 49};
  cuuint32_t box[4] = {64, 7, 7, 2};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D) failed (%d) pairs=%lld c=%d ptr=%p", (int)r, pairs, c, ptr);
    return NLV_ERR_CUDA;
  }
  return NLV_OK;
}

}  // namespace

// D[pairs*49, n] = sum over taps / channels of the shifted NHWC map x[pairs,7,7,c] (bf16) times B[n, 9*c] (bf16, K-major, k = tap*c + channel):
// flip = 0: x[p, y + ky - 1, x + kx - 1] (the forward 3x3 convolution, pad 1); flip = 1: x[p, y + 1 - ky, x + 1 - kx] (its data gradient).
// n <= 128.  No im2col / col2im matrices: the taps are 4-D TMA boxes with a zero-filled halo.
int gemm_tc_conv3x3(const void* x, long long pairs, int c, const void* b, int ldb, int n, int flip, void* d, int d_dtype, int ldd,
                    const float* bias, int relu, cudaStream_t stream) {
  NLV_CHECK_ARG(pairs >= 0 && c >= 64 && (c & 63) == 0 && n >= 8 && n <= 128 && (n & 7) == 0, "conv3x3: bad sizes (c %% 64, n <= 128)");
  NLV_CHECK_ARG(pairs * 49 < 0x7fffffffll, "conv3x3: too many rows");
  if (pairs == 0) return NLV_OK;
  NLV_CHECK_ARG(((uintptr_t)x & 15) == 0 && ((uintptr_t)b & 15) == 0 && (ldb & 7) == 0, "conv3x3: operands must be 16-byte aligned");
  CUtensorMap ta, tb;
  int rc = make_tmap_nhwc7(&ta, x, pairs, c);
  if (rc != NLV_OK) return rc;
  const bool pair_mode = pairs > 2;
  rc = make_tmap(&tb, b, n, 9ll * c, ldb, BLOCK_K, pair_mode ? 64 : 128);
  if (rc != NLV_OK) return rc;
  Params p;
  memset(&p, 0, sizeof(p));
  p.d = d; p.bias = bias; p.m = (int)(pairs * 49); p.n = n; p.k = 9 * c; p.ldd = ldd; p.d_dtype = d_dtype; p.r_dtype = NLV_F32; p.relu = relu;
  p.gate_scale = 1.f;
  p.num_m_blocks = cdiv(pairs, 2); p.num_n_blocks = 1; p.k_splits = 1; p.kb_per_split = cdiv(p.k, BLOCK_K);
  p.conv_c = c; p.conv_flip = flip; p.tile_rows = 98; p.a_tx = 98 * BLOCK_K * 2;
  return pair_mode ? run_kernel<128, false, false, 2>(ta, tb, p, stream) : run_kernel<128, false, false, 0>(ta, tb, p, stream);
}

namespace {
template <int BLOCK_N, int MODE>
int dispatch_major(const nlv_gemm_args& g, cudaStream_t s) {
  if (g.a_major == NLV_MAJOR_K && g.b_major == NLV_MAJOR_K) return launch<BLOCK_N, false, false, MODE>(g, s);
  if (g.a_major == NLV_MAJOR_K && g.b_major == NLV_MAJOR_MN) return launch<BLOCK_N, false, true, MODE>(g, s);
  if (g.a_major == NLV_MAJOR_MN && g.b_major == NLV_MAJOR_K) return launch<BLOCK_N, true, false, MODE>(g, s);
  return launch<BLOCK_N, true, true, MODE>(g, s);
}

}  // namespace

int gemm_tc(const nlv_gemm_args& g, cudaStream_t stream) {
  NLV_CHECK_ARG(((uintptr_t)g.a & 15) == 0 && ((uintptr_t)g.b & 15) == 0, "gemm(bf16): a and b must be 16-byte aligned");
  NLV_CHECK_ARG((g.lda & 7) == 0 && (g.ldb & 7) == 0, "gemm(bf16): lda=%d and ldb=%d must be multiples of 8", g.lda, g.ldb);
  // at least two row tiles: CTA pairs running cta_group::2 MMAs (NLV_GEMM_CLUSTER=2: the cta_group::1 kernel with a multicast
  // B tile instead; =1: single CTAs)
  static const int mode_env = [] { const char* e = getenv("NLV_GEMM_CLUSTER"); const int v = e == nullptr ? 0 : atoi(e); return v == 1 ? 0 : (v == 2 ? 1 : 2); }();
  // short reductions (k < 512: the 7x7 / 3x3-data-gradient convolutions) turn tiles over every few k-blocks; there the pair's
  // cross-CTA hand-shakes cost more than its shared-memory savings (measured: 157 vs 107 and 472 vs 422 TFLOP/s)
  int mode = g.m > BLOCK_M ? mode_env : 0;
  if (mode == 2 && g.k < 512) mode = 1;
  if (g.n > 128) return mode == 2 ? dispatch_major<256, 2>(g, stream) : mode == 1 ? dispatch_major<256, 1>(g, stream) : dispatch_major<256, 0>(g, stream);
  return mode == 2 ? dispatch_major<128, 2>(g, stream) : mode == 1 ? dispatch_major<128, 1>(g, stream) : dispatch_major<128, 0>(g, stream);
}

}  // namespace nlv
