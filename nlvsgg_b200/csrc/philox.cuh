// Counter-based dropout masks (Philox4x32-10, Salmon et al. 2011): the same (seed, stream, element) always yields the same
// bit, so forward and backward regenerate a mask instead of storing it.
//   matrix sites    : counter = (group index lo, hi, stream, 0), group = row * ceil(cols / 8) + col / 8; one call -> 8 keep bits
//   attention sites : counter = (row * heads + head lo, hi, stream, key / 8);                            one call -> 8 keep bits
// keep(i) = u16_i >= thr16 with thr16 = round(p * 65536): P(keep) = 1 - p up to 2^-16.
// Replaces torch's nn.Dropout / MultiheadAttention(dropout=) in lib/transformer.py:9-29,38-57, lib/sttran.py:46, lib/dsg_detr.py:28,48.
#pragma once
#include <stdint.h>

namespace nlv {

struct DropCfg {      // thr16 == 0: dropout off
  uint32_t thr16;
  float scale;        // 1 / (1 - p)
  uint32_t seed_lo, seed_hi;
  uint32_t stream;    // identifies the dropout site (layer * 8 + site)
};

__host__ __device__ inline DropCfg drop_off() { DropCfg d; d.thr16 = 0; d.scale = 1.f; d.seed_lo = d.seed_hi = d.stream = 0; return d; }

inline DropCfg make_drop(float p, unsigned long long seed, uint32_t stream) {
  DropCfg d = drop_off();
  if (p > 0.f) {
    d.thr16 = (uint32_t)(p * 65536.f + 0.5f);
    if (d.thr16 > 65535u) d.thr16 = 65535u;
    d.scale = 1.f / (1.f - p);
    d.seed_lo = (uint32_t)seed; d.seed_hi = (uint32_t)(seed >> 32); d.stream = stream;
  }
  return d;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// 8 keep bits (bit i = element i of the group is kept)
__device__ __forceinline__ uint32_t keep8(const DropCfg& d, uint32_t c0, uint32_t c1, uint32_t c3) {
  const uint4 r = philox4x32_10(c0, c1, d.stream, c3, d.seed_lo, d.seed_hi);
  uint32_t m = 0;
  m |= ((r.x & 0xffffu) >= d.thr16) ? 1u : 0u;   m |= ((r.x >> 16) >= d.thr16) ? 2u : 0u;
  m |= ((r.y & 0xffffu) >= d.thr16) ? 4u : 0u;   m |= ((r.y >> 16) >= d.thr16) ? 8u : 0u;
  m |= ((r.z & 0xffffu) >= d.thr16) ? 16u : 0u;  m |= ((r.z >> 16) >= d.thr16) ? 32u : 0u;
  m |= ((r.w & 0xffffu) >= d.thr16) ? 64u : 0u;  m |= ((r.w >> 16) >= d.thr16) ? 128u : 0u;
  return m;
}
// matrix site: keep bits of columns [8 g, 8 g + 8) of `row` in a matrix of `groups_per_row` = ceil(cols / 8) groups per row
__device__ __forceinline__ uint32_t keep8_matrix(const DropCfg& d, long long row, int g, int groups_per_row) {
  const unsigned long long idx = (unsigned long long)row * (unsigned)groups_per_row + (unsigned)g;
  return keep8(d, (uint32_t)idx, (uint32_t)(idx >> 32), 0u);
}
// attention site: keep bits of keys [8 kg, 8 kg + 8) (index inside the segment) for (global row, head)
__device__ __forceinline__ uint32_t keep8_attn(const DropCfg& d, long long row, int heads, int head, int kg) {
  const unsigned long long idx = (unsigned long long)row * (unsigned)heads + (unsigned)head;
  return keep8(d, (uint32_t)idx, (uint32_t)(idx >> 32), (uint32_t)kg);
}

}  // namespace nlv
