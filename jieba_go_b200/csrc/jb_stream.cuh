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
  uint32_t* wide_list;    // indexes of blocks with a 4-byte Han rune, for k_wide
  uint32_t wide_cap;
  uint32_t min_chunk;     // blocks a warp takes from the queue at least (few long blocks: lanes per warp vs warps per SM)
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
  uint32_t min_chunk;
};

struct WideArgs {
  const uint8_t* text;
  uint32_t n;
  const uint32_t* ds_bits;
  const uint2* blocks;
  const uint32_t* wide_list;
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
int launch_emit(const JbTables& T, const EmitArgs& A, bool hmm, int num_sms, cudaStream_t st);
inline uint32_t scan_tiles(uint32_t n) { return (n + kScTileBytes - 1) / kScTileBytes; }

}  // namespace jb
