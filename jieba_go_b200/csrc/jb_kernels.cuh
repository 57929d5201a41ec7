// Device-side pipeline interface (internal).  See DESIGN.md for the kernel list.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "jb_common.h"

namespace jb {

// Tile geometry of the split/DAG kernels: a tile is 1024 slots = 3072 bytes (a multiple of 3
// for the slot map and of 32 for the bitmaps).
constexpr int kTileSlots = 1024;
constexpr int kTileBytes = 3 * kTileSlots;
constexpr int kHaloL = 16;
constexpr int kHaloR = 144;  // >= longest Han key (90 B) + one rune + slack
constexpr int kRegion = kHaloL + kTileBytes + kHaloR;
// Token ranking tiles: 32 KiB of text = 1024 bitmap words
constexpr int kRankWords = 1024;
constexpr int kRankBytes = kRankWords * 32;

enum Counter {
  C_N_ENDS = 0,    // number of Han blocks (entries of `ends`)
  C_CUR_DP = 1,    // work cursor of the route-DP kernel
  C_CUR_WALK = 2,  // work cursor of the walk kernel
  C_STATUS = 3,    // bit0: candidate-weight buffer overflow
  C_N_TOKENS = 4,
  C_W_NEEDED = 5,  // max weights needed by any tile (to size a retry)
  C_N_WIDE = 6,    // Han blocks with a 4-byte rune, handed to k_wide
  C_N_DEFER = 7,   // gated non-Han tokens waiting for the tile-summary scan
  C_FLAGS = 8,     // bit0: the batch must be redone by the general pipeline
  C_N_SEG = 9,     // segments of long Han blocks (k_emit -> k_land / k_chain / k_emit)
  C_N_LONG = 10,   // long Han blocks
  C_CUR_SEG = 11,  // work cursor of k_emit over the segments
  C_N_BLK = 12,    // Han blocks listed by k_scan for k_route / k_emit
  C_CUR_ROUTE = 13,  // work cursors of k_route / k_emit
  C_CUR_EMIT = 14,
  C_NUM = 16
};

constexpr int kNumProfKernels = 8;
extern const char* const kProfKernelNames[kNumProfKernels];

struct Workspace {
  // per-kernel CUDA-event timing (jb_profile_*): ev[i] is recorded before kernel i, ev[9] after the last
  bool prof = false;
  bool prof_pending = false;
  cudaEvent_t ev[kNumProfKernels + 1] = {};
  double prof_ms[kNumProfKernels] = {};
  uint64_t prof_steps = 0;
  // side stream for the tile-summary scan (forked after k_scan, joined before the ranking)
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  // capacity
  uint64_t cap_bytes = 0;
  uint32_t w_per_tile = 0;  // candidate weights reserved per tile
  // buffers
  uint8_t* text = nullptr;      // staging for host-memory batches
  uint64_t* doc_off64 = nullptr; // staging for host-memory batches
  uint64_t cap_docs = 0;
  uint32_t* doc_off32 = nullptr;
  uint32_t* ds_bits = nullptr;   // document-start bitmap
  uint32_t* s_bits = nullptr;    // token-start bitmap
  uint32_t* e_bits = nullptr;    // token-end bitmap (bit at the token's last byte)
  uint32_t* m_bits = nullptr;    // single-rune pieces of the route (HMM: k_emit<2> -> k_runs)
  uint2* segs = nullptr;         // long Han blocks cut into segments: (first byte, runes)
  unsigned long long* land = nullptr;  // per segment: landing offsets of its first 16 runes
  uint2* longs = nullptr;        // per long block: (first segment, segments)
  uint32_t segs_cap = 0, longs_cap = 0;
  uint32_t* rec = nullptr;       // per-slot records
  uint32_t* gend = nullptr;      // per 32-slot group: end offset of its weights in wbuf
  double* wbuf = nullptr;        // candidate weights
  uint2* ends = nullptr;         // per Han block: (last rune slot, weight end offset)
  uint2* walks = nullptr;        // per Han block: (first rune slot, block end byte)
  uint8_t* tile_sum = nullptr;   // per split tile: has-boundary / alnum-before / alnum-after
  uint8_t* tile_ctx = nullptr;   // per split tile: bit0 fwd, bit1 bwd
  uint4* deferred = nullptr;
  uint32_t deferred_cap = 0;
  uint32_t* tile_last_hs = nullptr;  // per k_scan tile: last Han-block start (k_scan -> k_route)
  uint32_t* path = nullptr;      // chosen word length - 1 per rune (k_route -> k_emit)
  uint8_t* bp = nullptr;         // Viterbi back-pointers per rune (k_emit)
  uint32_t blocks_cap = 0;       // entries of `ends` usable as the stream path's block list
  uint32_t* wide_list = nullptr; // k_route -> k_wide
  uint32_t wide_cap = 0;
  uint32_t* tile_first_doc = nullptr;  // per rank tile: first document index with doc_off >= the tile's first byte
  uint32_t* rank_cnt = nullptr;  // per rank tile: token count, then exclusive prefix
  uint32_t* counters = nullptr;  // Counter
  // optional: the number of Han blocks goes to this (pinned) host word right after k_scan, ev_nblk is recorded behind it
  uint32_t* h_nblk = nullptr;
  cudaEvent_t ev_nblk = nullptr;
  double* dbg_proba = nullptr;   // optional: selected route value per slot (general path)
  double* dbg_R = nullptr;       // optional: selected route value / word length per rune (streaming path)
  uint8_t* dbg_D = nullptr;
  // outputs for host-memory batches
  uint32_t* out_start = nullptr;
  uint32_t* out_end = nullptr;
  uint64_t out_cap = 0;
  uint64_t* out_doc_tok = nullptr;
  uint64_t* out_ntok = nullptr;  // [2]
};

int workspace_reserve(Workspace& ws, uint64_t nbytes, uint64_t ndocs, double w_per_slot, bool host_staging);
void workspace_free(Workspace& ws);

// Where a batch's result goes (all device pointers).
struct PipeOut {
  uint32_t* d_start = nullptr;  // token (start,end), doc-relative, up to cap_tokens (both NULL: count only; run_scatter later)
  uint32_t* d_end = nullptr;
  uint64_t cap_tokens = 0;
  uint64_t* d_doc_tok_off = nullptr;  // [ndocs+1]: tok_base + rank of each document's first token
  uint64_t tok_base = 0;
  uint64_t* d_n_tokens = nullptr;     // [0] token count of the batch, [1] status word
  // BITMAP result: bit p of d_s_bits / d_e_bits = a token starts at / ends with byte p of the batch.  When given, these
  // buffers ((nbytes / 32 + 8) words each) are the pipeline's working bitmaps instead of the workspace's.
  uint32_t* d_s_bits = nullptr;
  uint32_t* d_e_bits = nullptr;
  bool bits_only = false;  // no (start,end) arrays at all: the bitmaps + doc_tok_off are the result (k_rank_scatter is not run)
  uint32_t pos0 = 0;       // the first document starts at byte pos0 (< 32) of d_text; bytes before it must be spaces
  bool no_general = false; // the caller guarantees that no list of the streaming path can overflow (small fixed-size
                           // batches with lists sized for the worst case): the flag-gated general kernels are not enqueued
};

// Enqueue the whole Cut pipeline for one batch on `stream`.
//   d_text[nbytes], d_doc_off[ndocs+1] (uint64, absolute; doc_off[0] is subtracted, out.pos0 added) on device.
//   path: PATH_DEFAULT k_scan -> k_route -> k_emit (the streaming path); PATH_GENERAL skip it and run the general
//   kernels on everything.
int run_pipeline(const JbTables& T, Workspace& ws, const uint8_t* d_text, uint32_t nbytes, const uint64_t* d_doc_off,
                 uint64_t ndocs, bool use_hmm, const PipeOut& out, cudaStream_t stream, int path = 0);
enum { PATH_DEFAULT = 0, PATH_GENERAL = 1 };

// Second phase when d_start/d_end were NULL in run_pipeline (count first, then scatter).
int run_scatter(Workspace& ws, uint32_t nbytes, uint64_t ndocs, uint32_t* d_start, uint32_t* d_end, uint64_t cap_tokens,
                uint64_t* d_doc_tok_off, uint64_t tok_base, cudaStream_t stream, const uint32_t* d_s_bits = nullptr,
                const uint32_t* d_e_bits = nullptr);

int debug_lookup(const JbTables& T, const uint32_t* runes_host, int L, int* kind, double* w);

uint64_t kernel_launch_count();
void profile_collect(Workspace& ws);

}  // namespace jb
