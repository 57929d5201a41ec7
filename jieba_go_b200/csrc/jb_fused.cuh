// Fused tile kernel: the fast path of the Cut pipeline.  One CTA owns a 3 KiB tile of text and
// does, entirely in shared memory: UTF-8 decode + classification (slot-centric for 3-byte runes,
// exact per-byte rules only where something else occurs), non-Han tokens (cutNonZh), Han block
// detection, the DAG probe of the HBM-resident rune-prefix hash (buildDag), the route DP with the
// reference's selector (calcDagProba/maxIndexProba), the path walk (findDagPath), the HMM glue and
// Viterbi (cutZh/viterbi/cutHMM), and finally token start/end bits.  Per-position intermediates
// (candidate masks, weights, route values) never leave the SM.
//
// What it does NOT handle is handed to the general kernels of jb_kernels.cu:
//   * a Han block that does not end inside the tile's 768-byte halo, or a tile whose candidate
//     weights overflow shared memory  -> "long block" list (k_long_extent, k_split<dag-only>,
//     k_route_dp, k_walk)
//   * a well-formed 4-byte Han rune anywhere in the batch -> batch flag, the whole batch is redone
//     by the general pipeline (rare: CJK extension B+)
//   * a gated non-Han token whose block reaches a tile edge without an alnum -> deferred list,
//     resolved after the tile-summary scan (k_resolve_deferred)
#pragma once
#include "jb_kernels.cuh"

namespace jb {

constexpr int kFtTileBytes = kTileBytes;  // 3072: same tiles as k_split, so k_tile_scan is shared
constexpr int kFtHaloBytes = 768;
constexpr int kFtLeft = 16;
constexpr int kFtRightPad = 16;
constexpr int kFtRegion = kFtLeft + kFtTileBytes + kFtHaloBytes + kFtRightPad;  // 3872
constexpr int kFtSlots = (kFtTileBytes + kFtHaloBytes) / 3;                    // 1280
constexpr int kFtWords = (kFtTileBytes + kFtHaloBytes) / 32;                   // 120
constexpr int kFtTileWords = kFtTileBytes / 32;                                // 96
constexpr int kFtWCap = 2560;                                                  // candidate weights per tile in smem
constexpr int kFtThreads = 640;  // 20 warps: 2 slots per thread, 3 CTAs = 60 warps per SM (the probe chains are latency-bound)
constexpr int kFtWorkList = 384;  // prefix chains longer than 3 runes per tile (more => the tile goes the general way)
constexpr int kFtMaxBlocks = 768;
constexpr int kFtMaxBlockLen = 256;  // longest Han block (runes) the block-DP kernel takes; longer ones go the general way  // owned Han blocks per tile (a block needs >= 4 bytes)

struct FusedArgs {
  const uint8_t* text;
  uint32_t n;
  const uint32_t* ds_bits;
  uint32_t* s_bits;
  uint32_t* e_bits;
  uint8_t* tile_sum;
  uint32_t* counters;
  uint32_t* long_seeds;   // byte positions of long-block starts
  uint32_t long_cap;
  uint4* deferred;        // (byte pos, len, flags: 1 need fwd 2 need bwd, tile)
  uint32_t deferred_cap;
  unsigned long long* stream;  // packed candidate records of the owned blocks
  uint32_t stream_cap;         // in 8-byte units
  uint4* fblocks;              // per packed block: (stream offset, byte position of its first rune, runes, 0)
  uint32_t fblk_cap;
};

struct BlockDpArgs {
  const uint8_t* text;
  unsigned long long* stream;
  const uint4* fblocks;
  uint32_t fblk_cap;
  uint32_t* counters;
  uint32_t* s_bits;
  uint32_t* e_bits;
};
int launch_block_dp(const JbTables& T, const BlockDpArgs& A, bool hmm, int num_sms, cudaStream_t st);

int launch_fused(const JbTables& T, const FusedArgs& A, uint32_t ntiles, bool hmm, cudaStream_t st);

}  // namespace jb
