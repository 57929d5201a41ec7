// Shared host/device definitions: HBM table layouts, the rune-sequence hash and the
// tokenizer's device parameter block.  See DESIGN.md "Data layout in HBM".
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define JB_HD __host__ __device__ __forceinline__
#else
#define JB_HD inline
#endif

#define JB_MINF (-3.14e100)  // minFloat, /root/reference/tokenizer.go:19

// ---------------------------------------------------------------------------------------
// First-rune table, direct-indexed by BMP code point (65536 x 16 B = 1 MiB, L2 resident).
// Answers buildDag's first probe termFreq[string(iRune)] (tokenizer.go:468-472) in one load.
// ---------------------------------------------------------------------------------------
struct __attribute__((aligned(16))) JbFirst {
  double w;        // weight of the edge (i,i+1): log(freq)-log(size); key missing: -log(size)
                   // (tokenizer.go:515,519 tf=1.0); freq 0: -Inf (tokenizer.go:516-519)
  uint32_t info;   // bit0 GATE: key missing or freq 0 => only edge (i,i+1), no longer probes
                   // bits 8..15: max key length in runes among keys starting with this rune
  uint32_t child;  // 32-bit Bloom of the 2nd rune over all 2-rune keys starting with this rune
};
#define JB_FIRST_GATE 1u

// ---------------------------------------------------------------------------------------
// Rune-prefix hash: open addressing, linear probing, 32-byte entries (one L2 sector each).
// Key = exact rune sequence.  Inline form: up to 8 BMP runes packed 16 bits each in k0,k1
// (Han code units are never 0, so the length is implicit).  Long form (more than 8 runes, or
// any supplementary-plane rune): k0 = 64-bit hash, k1 = blob offset | length, verified against
// the key blob.
// ---------------------------------------------------------------------------------------
struct __attribute__((aligned(32))) JbEntry {
  uint64_t k0;
  uint64_t k1;
  double w;        // log(freq) - log(size)   (only meaningful when POSITIVE)
  uint32_t child;  // Bloom of the next rune over keys that extend this key by one rune
  uint32_t meta;   // bit0 USED, bit1 POSITIVE (freq > 0), bit2 LONG form, bits 8..15 length in runes
};
#define JB_E_USED 1u
#define JB_E_POS 2u
#define JB_E_LONG 4u

JB_HD uint32_t jb_hash_init(uint32_t r0) { return (r0 ^ 0x811C9DC5u) * 0x01000193u; }
JB_HD uint32_t jb_hash_step(uint32_t h, uint32_t r) { return (h ^ r) * 0x01000193u + 0x9E3779B9u; }
JB_HD uint32_t jb_hash_fin(uint32_t h) {
  h ^= h >> 15;
  h *= 0x2C1B3C6Du;
  h ^= h >> 12;
  h *= 0x297A2D39u;
  h ^= h >> 15;
  return h;
}
JB_HD uint32_t jb_bloom_bit(uint32_t r) { return ((r * 0x9E3779B1u) >> 27) & 31u; }
// second, independent 64-bit hash used as the long-form key tag
JB_HD uint64_t jb_hash64_step(uint64_t h, uint32_t r) {
  h ^= r;
  h *= 0x100000001B3ull;
  h ^= h >> 29;
  return h;
}
#define JB_HASH64_INIT 0xCBF29CE484222325ull

// ---------------------------------------------------------------------------------------
// Per-slot record (uint32), slot(p) = floor((p+2)/3) for the lead byte p of a Han rune: every
// rune of >= 3 bytes owns exactly one slot; a 4-byte rune may leave the following slot unused
// ("hole", record 0).  All candidate lengths are expressed as slot DELTAS.
//   after the DAG kernel : bits 0..29 candidate mask (bit d-1 <=> edge to slot k+d, ascending
//                          order = ascending word length), bit 31 = first rune of its block
//   after the route DP   : bits 0..7 chosen delta, bit 8 = chosen piece is a single rune,
//                          bit 31 kept
//   after Viterbi        : bits 16..23 four 2-bit back-pointers, bit 24 = state in {E,S}
// ---------------------------------------------------------------------------------------
#define JB_REC_START 0x80000000u
#define JB_REC_MASK 0x3FFFFFFFu
#define JB_MAX_DELTA 30
#define JB_REC_SINGLE 0x100u
#define JB_REC_ES 0x01000000u

#define JB_MAX_SUPP_RANGES 16

struct JbTables {
  const JbFirst* first;        // [65536]
  const JbEntry* entries;      // [hash_cap]
  uint32_t hash_mask;          // hash_cap - 1
  const uint32_t* key_blob;    // code points of long-form keys
  const double* emit;          // [65536][4] B,M,E,S; missing = JB_MINF (tokenizer.go:690-692)
  const uint32_t* emit_supp_rune;  // sorted supplementary-plane runes with an emission
  const double* emit_supp;         // [n][4]
  uint32_t n_emit_supp;
  const uint32_t* han_bits;    // [2048] bitmap of \p{Han} over the BMP
  uint32_t n_supp;             // supplementary-plane Han ranges
  uint32_t supp_lo[JB_MAX_SUPP_RANGES], supp_hi[JB_MAX_SUPP_RANGES];
  double neg_log_total;        // -log(size): weight of a missing single rune
  double start[4];             // startP
  double trans[4][2];          // trans[now][c] = transP[prev_c(now)][now], c-th entry of stateChange[now]
  uint32_t max_delta;          // largest slot delta of any Han key (<= JB_MAX_DELTA)
};
