// Streaming fast path of the Cut pipeline: three kernels, no per-position state in HBM.
//
//   k_scan   (S, N)        one lane per 32-byte word of text: UTF-8 / \p{Han} classification as BITMASKS
//                          (SWAR byte predicates -> one bit per byte), block boundaries, cutNonZh tokens,
//                          the Han-block start bitmap and the list of Han-block ends.
//   k_route  (D, G, P, X)  one lane per Han block, lanes refilled as they finish: walks the block right to
//                          left and, per rune, does buildDag's prefix probes and calcDagProba's selector
//                          update in one state machine -- candidates never leave registers; only the
//                          chosen word length per rune (4 bits) goes to HBM.
//   k_emit   (W, H, V, O)  one lane per Han block: findDagPath walk, Viterbi over single-rune runs, token bits.
//                          Blocks of 512 runes or more are cut into 256-rune segments first (k_land, k_chain) and
//                          walked one lane per segment; their Viterbi runs go to k_runs, one lane per run.
//
//   k_wide   (all rows)    the rare Han blocks that contain a 4-byte rune: one lane restates the reference per block.
//
// Handed to the general kernels (jb_kernels.cu): a batch that overflows a list (C_FLAGS bit0).  Blocks of any
// length stay on this path.
#pragma once
#include "jb_kernels.cuh"

namespace jb {

constexpr int kScThreads = 256;               // lane 0 / 255 classify one halo word on either side
constexpr int kScOwnWords = kScThreads - 2;   // 254 words = 8128 bytes per tile
constexpr int kScTileBytes = kScOwnWords * 32;
constexpr int kScLeft = 48;                   // staged bytes before the tile: 16 pad + the 32-byte halo word
constexpr int kScRegion = kScLeft + kScTileBytes + 48;
constexpr uint32_t kWideBlock = 0xFFFFFFFFu;  // block length marker: ends with a 4-byte Han rune

struct ScanArgs {
  const uint8_t* text;
  uint32_t n;
  const uint32_t* ds_bits;
  uint32_t* s_bits;
  uint32_t* e_bits;
  uint32_t* tile_last_hs; // per tile: byte position of its last Han-block start, 0xFFFFFFFF if none
  uint8_t* tile_sum;
  uint32_t* counters;
  uint4* deferred;        // (byte pos, len, flags: 1 need fwd 2 need bwd, tile)
  uint32_t deferred_cap;
  uint2* blocks;          // (lead byte of the LAST rune of a Han block, its runes or 0 = began in an earlier tile)
  uint32_t blocks_cap;
};

struct RouteArgs {
  const uint8_t* text;
  const uint32_t* tile_last_hs;
  uint2* blocks;          // in: (last rune, runes or 0 when the block began in an earlier tile); out: (lead byte of the first rune, runes)
  uint32_t blocks_cap;
  uint32_t* counters;
  uint32_t* path;         // chosen word length - 1 per rune, index = lead byte / 3
  uint32_t* wide_list;    // lead byte of the last rune of every block with a 4-byte Han rune, for k_wide
  uint32_t wide_cap;
  uint32_t min_chunk;     // blocks a warp takes from the queue at least (few long blocks: lanes per warp vs warps per SM)
  double* dbg_R;          // optional (jb_debug_route): selected route value / word length per rune, index = lead byte / 3
  uint8_t* dbg_D;
};

struct EmitArgs {
  const uint8_t* text;
  const uint2* blocks;
  uint32_t blocks_cap;
  uint32_t* counters;
  const uint32_t* path;
  uint8_t* bp;            // Viterbi back-pointers / state flags per rune, index = lead byte / 3
  uint32_t* s_bits;
  uint32_t* e_bits;
  uint32_t* m_bits;       // single-rune pieces marked by k_emit<2> for k_runs (HMM)
  uint32_t min_chunk;
  // long blocks: segment list (first byte, runes), per segment the landing offsets of its first 16 runes, and per long
  // block (first segment, segments).  segs == nullptr: blocks of any length are walked by one lane.
  uint2* segs;
  uint32_t segs_cap;
  unsigned long long* land;
  uint2* longs;
  uint32_t longs_cap;
  uint32_t count_idx, cursor_idx;  // counters that hold the length of `blocks` and the work cursor (set by launch_emit)
};

struct WideArgs {
  const uint8_t* text;
  uint32_t n;
  const uint32_t* ds_bits;
  const uint32_t* wide_list;  // lead byte of the block's last rune
  uint32_t wide_cap;
  const uint32_t* counters;
  double* R;        // per-rune scratch, index = lead byte / 3: route value,
  uint8_t* len8;    //   chosen word length in bytes,
  uint8_t* code8;   //   Viterbi back-pointers / state flag
  uint32_t* s_bits;
  uint32_t* e_bits;
};

int launch_wide(const JbTables& T, const WideArgs& A, bool hmm, int num_sms, cudaStream_t st);
int launch_scan(const JbTables& T, const ScanArgs& A, cudaStream_t st);
int launch_route(const JbTables& T, const RouteArgs& A, int num_sms, cudaStream_t st);
// returns the number of kernels launched, or -1
int launch_emit(const JbTables& T, const EmitArgs& A, bool hmm, int num_sms, cudaStream_t st, uint32_t n, const uint32_t* ds_bits);
inline uint32_t scan_tiles(uint32_t n) { return (n + kScTileBytes - 1) / kScTileBytes; }

#if defined(__CUDACC__)
// 16-byte read-only load that asks L1 to keep the line (the first-rune table is the hottest data of the probe kernels)
__device__ __forceinline__ uint4 ldg_keep(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::evict_last.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}

struct BitAcc2 {  // token bits of one lane, flushed one 32-byte word at a time (positions only grow)
  uint32_t* bits;
  uint32_t w, m;
  __device__ __forceinline__ void init(uint32_t* b) {
    bits = b;
    w = 0xFFFFFFFFu;
    m = 0;
  }
  __device__ __forceinline__ void set(uint32_t p) {
    const uint32_t pw = p >> 5;
    if (pw != w) {
      if (m) atomicOr(&bits[w], m);
      w = pw;
      m = 0;
    }
    m |= 1u << (p & 31);
  }
  __device__ __forceinline__ void flush() {
    if (m) atomicOr(&bits[w], m);
    m = 0;
    w = 0xFFFFFFFFu;
  }
};
// bits of lo -> positions p .., bits of hi -> positions p + 48 .. (lo spans at most 48 bits, hi 24): straight to
// the bitmap, branch-free (a run of <= 24 runes touches at most four 32-byte words)
__device__ __forceinline__ void or_span(uint32_t* __restrict__ bits, uint32_t p, unsigned long long lo, unsigned long long hi) {
  const uint32_t sh = p & 31u, pw = p >> 5;
  const unsigned long long v0 = lo | (hi << 48), v1 = hi >> 16;  // the 72-bit value
  const unsigned long long s0 = v0 << sh, s1 = (v1 << sh) | (sh ? (v0 >> (64u - sh)) : 0ull);
  const uint32_t x0 = (uint32_t)s0, x1 = (uint32_t)(s0 >> 32), x2 = (uint32_t)s1, x3 = (uint32_t)(s1 >> 32);
  if (x0) atomicOr(&bits[pw], x0);
  if (x1) atomicOr(&bits[pw + 1u], x1);
  if (x2) atomicOr(&bits[pw + 2u], x2);
  if (x3) atomicOr(&bits[pw + 3u], x3);
}
// bit j of x (j < 16) -> bit 3j
__device__ __forceinline__ unsigned long long spread3(uint32_t x16) {
  unsigned long long x = x16 & 0xFFFFu;
  x = (x | (x << 16)) & 0x0000FF0000FFull;
  x = (x | (x << 8)) & 0x00F00F00F00Full;
  x = (x | (x << 4)) & 0x0C30C30C30C3ull;
  x = (x | (x << 2)) & 0x249249249249ull;
  return x;
}
#endif

}  // namespace jb
