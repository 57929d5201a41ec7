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
// Rune-prefix hash: open addressing, linear probing, 16-byte entries (two per 32-byte L2 sector),
// home slot = jb_hash_next fold over the key's runes (below).
// A key is stored as a TRIE EDGE: (id of the entry of the key minus its last rune, last rune).
// Matching a probe is two integer compares and is exact -- no key bytes, no fingerprints -- and it
// mirrors how buildDag reaches a key: only through all of its prefixes (tokenizer.go:473-482), so a key
// whose proper prefix is missing is unreachable in the reference too and is simply not stored.
//   parent id: slot index of the parent entry (< 0x80000000)
//              0x80000000 | r0   for 2-rune keys whose first rune r0 is in the BMP (first-rune table)
//              0xC0000000        for 1-rune keys outside the BMP (root)
//   rb       : bits 0..20 last rune, bits 21..30 a 10-bit Bloom filter of the next rune over the
//              keys that extend this key by one rune (a miss ends the loop without a memory access),
//              bit 31 CONT: some key whose home slot is THIS slot lives further down the probe sequence
//              (clear => a lookup that finds a foreign entry in its home slot can stop: the key is absent)
//   w        : log(freq) - log(size); -Inf marks a key with freq 0 (prefix-only, tokenizer.go:360)
// ---------------------------------------------------------------------------------------
struct __attribute__((aligned(16))) JbEntry {
  double w;
  uint32_t parent;
  uint32_t rb;
};
#define JB_PARENT_EMPTY 0xFFFFFFFFu
#define JB_PARENT_ROOT 0xC0000000u
#define JB_PARENT_FIRST(r0) (0x80000000u | (r0))
#define JB_RB_RUNE(rb) ((rb) & 0x1FFFFFu)
#define JB_RB_CONT 0x80000000u

// Slot hash of a key = a fold over its RUNES (not over the parent's slot), so that the slots of successive
// prefixes of a text position can all be computed -- and fetched -- before any of them has been looked at:
//   state after the first rune r0:  JB_PARENT_FIRST(r0) for a BMP rune, jb_hash_next(JB_PARENT_ROOT, r0) otherwise
//   state after one more rune r:    jb_hash_next(state, r);  the key's home slot is the TOP bits of the state
// The fold is two multiply-adds and the slot one shift (Fibonacci hashing of a polynomial in the runes): the probe
// kernels are bound by the integer ALU pipe (logic / shift / select operations issue at half rate, multiply-adds go to
// the FMA pipe), and the xor-shift finaliser this replaces cost five ALU operations per probe for the same number of
// displaced keys (85.4 k of 599 k on the benchmark dictionary either way).
JB_HD uint32_t jb_hash_next(uint32_t h, uint32_t rune) { return h * 0x9E3779B1u + rune * 0x85EBCA6Bu; }
JB_HD uint32_t jb_hash_slot(uint32_t h, uint32_t shift) { return h >> shift; }
JB_HD uint32_t jb_bloom_bit(uint32_t r) { return (r * 0x9E3779B1u) >> 27; }  // first-rune table, 32 bits
JB_HD uint32_t jb_bloom11(uint32_t r) {                                       // hash entries, 10 bits (21..30): floor(x * 10 / 2^32)
#if defined(__CUDA_ARCH__)
  return __umulhi(r * 0x9E3779B1u, 10u);
#else
  return (uint32_t)(((uint64_t)(uint32_t)(r * 0x9E3779B1u) * 10u) >> 32);
#endif
}
// +/-Inf test on the bits (w is -Inf exactly for freq-0 keys)
JB_HD bool jb_w_positive(double w) { return w > -1.0e308; }

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

#if defined(__CUDACC__)
// One trie-edge lookup: termFreq[prefix + rune] where `parent` identifies termFreq[prefix] and `hs` is the
// hash state of the prefix (updated to the state of prefix + rune).  Returns the slot (>= 0) and fills
// w / rb, or -1 when the key is missing.  One 16-byte load per step.
__device__ __forceinline__ int jb_probe_edge(const JbEntry* __restrict__ entries, uint32_t mask, uint32_t shift, uint32_t& hs,
                                             uint32_t parent, uint32_t rune, double* w, uint32_t* rb) {
  hs = jb_hash_next(hs, rune);
  uint32_t slot = jb_hash_slot(hs, shift);
  for (bool home = true;; home = false) {
    const uint4 e = __ldg(reinterpret_cast<const uint4*>(entries + slot));
    if (e.z == JB_PARENT_EMPTY) return -1;
    if (e.z == parent && JB_RB_RUNE(e.w) == rune) {
      *w = __longlong_as_double(((long long)e.y << 32) | (long long)e.x);
      *rb = e.w;
      return (int)slot;
    }
    if (home && !(e.w & JB_RB_CONT)) return -1;  // nothing with this home slot was displaced
    slot = (slot + 1) & mask;
  }
}
#endif

struct JbTables {
  const JbFirst* first;        // [65536]
  const JbEntry* entries;      // [hash_cap]
  uint32_t hash_mask;          // hash_cap - 1 (linear probing wraps with it)
  uint32_t hash_shift;         // 32 - log2(hash_cap): home slot = hash state >> hash_shift
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
