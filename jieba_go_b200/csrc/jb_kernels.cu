// Hand-written sm_100a kernels for jieba-go's Cut hot path (see DESIGN.md for the map from
// reference functions to kernels).  Everything here is integer / byte work plus float64
// add/compare evaluated in the reference's operation order; no tensor cores, no library calls.
//
//   k_docstart      doc offsets -> document-start bitmap
//   k_split<true>   per 3 KiB tile: UTF-8 decode + classify, block boundaries, alnum summary       (pass 1)
//   k_tile_scan     segmented OR-scan of tile summaries (does a non-Han block hold any [a-zA-Z0-9]?)
//   k_split<false>  pass 2: non-Han tokens (cutNonZh), Han block list, DAG probe of the rune-prefix
//                   hash with smem-staged windows (buildDag) -> per-slot candidate masks + weights
//   k_route_dp      right-to-left route DP with the reference's selector (calcDagProba/maxIndexProba)
//   k_walk          forward path walk (findDagPath) + BMES Viterbi over single-rune runs (cutZh,
//                   viterbi, cutHMM) -> token start/end bits
//   k_rank_*        bitmap rank + scatter: token offsets in document order (the appends of Cut)
#include "jb_kernels.cuh"
#include "jb_stream.cuh"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>

#include "../../include/jieba_b200.h"

namespace jb {

static std::atomic<uint64_t> g_launches{0};
uint64_t kernel_launch_count() { return g_launches.load(); }
#define JB_LAUNCH(kernel, grid, block, smem, stream, ...)            \
  do {                                                               \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);      \
    g_launches.fetch_add(1);                                         \
  } while (0)

#define FULL 0xFFFFFFFFu

// rune classes (3 bits) | length (3 bits) in one byte; 0 = not a rune start
enum : uint32_t { CL_HAN = 1, CL_ALNUM = 2, CL_SPACE = 3, CL_OTHER = 4, CL_INVALID = 5 };
#define CLS(c, len) (uint8_t)(((c) << 3) | (len))
#define CLS_PENDING 0xFFu

__device__ __forceinline__ bool d_is_alnum(uint32_t c) {
  return (c - '0' < 10u) || ((c | 0x20) - 'a' < 26u);
}
// unicode.IsSpace (tokenizer.go:302)
__device__ __forceinline__ bool d_is_space(uint32_t cp) {
  if (cp <= 0xFF) return (cp - 9u < 5u) || cp == 0x20 || cp == 0x85 || cp == 0xA0;
  return cp == 0x1680 || (cp - 0x2000u <= 0xAu) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F ||
         cp == 0x3000;
}
__device__ __forceinline__ bool d_is_han(uint32_t cp, const uint32_t* han_bits, const JbTables& T) {
  if (cp < 0x10000) return (han_bits[cp >> 5] >> (cp & 31)) & 1;
  for (uint32_t i = 0; i < T.n_supp; i++)
    if (cp >= T.supp_lo[i] && cp <= T.supp_hi[i]) return true;
  return false;
}
__device__ __forceinline__ uint32_t d_decode(const uint8_t* b, int len) {
  uint32_t b0 = b[0];
  if (len == 1) return b0;
  if (len == 2) return ((b0 & 0x1F) << 6) | (b[1] & 0x3F);
  if (len == 3) return ((b0 & 0x0F) << 12) | ((b[1] & 0x3Fu) << 6) | (b[2] & 0x3F);
  return ((b0 & 0x07) << 18) | ((b[1] & 0x3Fu) << 12) | ((b[2] & 0x3Fu) << 6) | (b[3] & 0x3F);
}
__device__ __forceinline__ uint32_t d_slot(uint32_t p) { return (p + 2u) / 3u; }

// ------------------------------------------------------------------------------------------
__global__ void k_docstart(const uint64_t* __restrict__ doc_off, uint64_t ndocs, uint32_t n, uint32_t pos0, uint32_t* __restrict__ doc_off32,
                           uint32_t* __restrict__ ds_bits, uint32_t* __restrict__ tile_first_doc) {
  uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d > ndocs) return;
  uint64_t base = doc_off[0] - pos0;  // (as if pos0 bytes came before the first document; wraps harmlessly)
  uint64_t p64 = doc_off[d] - base;
  uint32_t p = p64 > n ? n : (uint32_t)p64;
  doc_off32[d] = p;
  if (d < ndocs && p < n) atomicOr(&ds_bits[p >> 5], 1u << (p & 31));
  // lower_bound(doc_off32, t * kRankBytes) for every rank tile t: document d answers for the tiles whose first byte
  // lies in (doc_off[d-1], doc_off[d]]  (k_rank_scatter used to search for it, one thread per tile)
  uint32_t t_lo = 0;
  if (d) {
    uint64_t q64 = doc_off[d - 1] - base;
    t_lo = (q64 > n ? n : (uint32_t)q64) / (uint32_t)kRankBytes + 1u;
  }
  for (uint32_t t = t_lo; t <= p / (uint32_t)kRankBytes; t++) tile_first_doc[t] = (uint32_t)d;
}

// ------------------------------------------------------------------------------------------
// Shared tile front end: stage [t0-16, t0+3072+144) in shared memory and classify every byte
// under Go's UTF-8 decoding rules (ill-formed => U+FFFD of width 1), clipped at document
// boundaries.  Restates what regexp \p{Han}+ over []byte (tokenizer.go:21,154), `range s`
// (tokenizer.go:301) and unicode.IsSpace see.
// ------------------------------------------------------------------------------------------
struct TileSmem {
  uint8_t sb[kRegion];
  uint8_t cls[kRegion];
  uint32_t dsw[(kRegion + 63) / 32 + 1];
  uint32_t han[2048];
  uint32_t BND[kTileBytes / 32 + 1];
  uint32_t ALN[kTileBytes / 32 + 1];
};

struct TileCtx {
  const TileSmem* s;
  uint32_t t0, n;
  __device__ __forceinline__ bool ds_at(int i) const {  // is region index i a document start (or >= n)?
    int64_t P = (int64_t)t0 - kHaloL + i;
    if (P >= (int64_t)n) return true;
    if (P < 0) return false;
    return (s->dsw[(i + 16) >> 5] >> ((i + 16) & 31)) & 1;
  }
};

template <int NT>
__device__ void classify_tile(TileSmem& S, const uint8_t* __restrict__ text, uint32_t n, const uint32_t* __restrict__ ds_bits,
                              uint32_t t0, const JbTables& T) {
  const int tid = threadIdx.x;
  // bytes
  const bool aligned = ((reinterpret_cast<uintptr_t>(text) & 15) == 0);
  for (int c = tid; c < kRegion / 16; c += NT) {
    int64_t P = (int64_t)t0 - kHaloL + c * 16;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (aligned && P >= 0 && P + 16 <= (int64_t)n) {
      v = __ldg(reinterpret_cast<const uint4*>(text + P));
    } else {
      uint8_t* vb = reinterpret_cast<uint8_t*>(&v);
      for (int j = 0; j < 16; j++) {
        int64_t q = P + j;
        vb[j] = (q >= 0 && q < (int64_t)n) ? __ldg(text + q) : 0;
      }
    }
    *reinterpret_cast<uint4*>(&S.sb[c * 16]) = v;
  }
  // document-start words: local word j <-> global word t0/32 - 1 + j
  const int n_dsw = (kRegion + 63) / 32 + 1;
  const uint32_t nwords = (n + 31) / 32;
  for (int j = tid; j < n_dsw; j += NT) {
    int64_t gw = (int64_t)(t0 / 32) - 1 + j;
    S.dsw[j] = (gw >= 0 && gw < (int64_t)nwords) ? __ldg(ds_bits + gw) : 0;
  }
  for (int j = tid; j < 2048; j += NT) S.han[j] = __ldg(T.han_bits + j);
  __syncthreads();
  TileCtx cx{&S, t0, n};
  // phase 1: every non-continuation byte is a rune start (see DESIGN.md "UTF-8 without a scan")
  for (int i = tid; i < kRegion; i += NT) {
    uint32_t b = S.sb[i];
    uint8_t c;
    if (b < 0x80) {
      c = d_is_alnum(b) ? CLS(CL_ALNUM, 1) : (d_is_space(b) ? CLS(CL_SPACE, 1) : CLS(CL_OTHER, 1));
    } else if ((b & 0xC0) == 0x80) {
      c = CLS_PENDING;
    } else {
      int len = 0;
      if (b >= 0xC2 && b <= 0xDF) len = 2;
      else if (b >= 0xE0 && b <= 0xEF) len = 3;
      else if (b >= 0xF0 && b <= 0xF4) len = 4;
      bool ok = len != 0 && i + len <= kRegion;
      if (ok) {
        uint32_t b1 = S.sb[i + 1];
        uint32_t lo = 0x80, hi = 0xBF;
        if (b == 0xE0) lo = 0xA0;
        if (b == 0xED) hi = 0x9F;
        if (b == 0xF0) lo = 0x90;
        if (b == 0xF4) hi = 0x8F;
        ok = b1 >= lo && b1 <= hi && !cx.ds_at(i + 1);
        if (ok && len >= 3) ok = (S.sb[i + 2] & 0xC0) == 0x80 && !cx.ds_at(i + 2);
        if (ok && len == 4) ok = (S.sb[i + 3] & 0xC0) == 0x80 && !cx.ds_at(i + 3);
      }
      if (!ok) {
        c = CLS(CL_INVALID, 1);
      } else {
        uint32_t cp = d_decode(&S.sb[i], len);
        c = d_is_han(cp, S.han, T) ? CLS(CL_HAN, len) : (d_is_space(cp) ? CLS(CL_SPACE, len) : CLS(CL_OTHER, len));
      }
    }
    S.cls[i] = c;
  }
  __syncthreads();
  // phase 2: a continuation byte is interior iff the nearest non-continuation byte within 3
  // to its left starts a well-formed sequence that covers it; otherwise it stands alone as U+FFFD.
  for (int i = tid; i < kRegion; i += NT) {
    if (S.cls[i] != CLS_PENDING) continue;
    uint8_t c = CLS(CL_INVALID, 1);
    for (int k = 1; k <= 3 && i - k >= 0; k++) {
      if ((S.sb[i - k] & 0xC0) != 0x80) {
        if ((S.cls[i - k] & 7) > k) c = 0;
        break;
      }
    }
    S.cls[i] = c;
  }
  __syncthreads();
}

// class of the rune that ends right before region index i (i >= 4)
__device__ __forceinline__ uint32_t prev_rune_class(const TileSmem& S, int i) {
  for (int k = 1; k <= 4; k++) {
    uint8_t c = S.cls[i - k];
    if (c) return c >> 3;
  }
  return CL_OTHER;
}

struct SplitArgs {
  const uint8_t* text;
  uint32_t n;
  const uint32_t* ds_bits;
  uint32_t* s_bits;
  uint32_t* e_bits;
  uint32_t* rec;
  uint32_t* gend;
  double* wbuf;
  uint32_t w_per_tile;
  uint2* ends;
  uint32_t* counters;
  const uint8_t* tile_ctx;
  uint8_t* tile_sum;
  uint32_t ntiles;
  int mode;  // 0: whole general pipeline;
             // 2: whole general pipeline, but only if the batch was flagged for it (C_FLAGS bit0)
};

// Does the non-Han block around tile-local byte `pos` contain an ASCII alnum?  (cutNonZh's
// `len(alnumIdx) == 0 -> return []`, tokenizer.go:290-293.)
__device__ bool block_has_alnum(const uint32_t* BND, const uint32_t* ALN, int pos, bool fwd_in, bool bwd_in) {
  const int NW = kTileBytes / 32;
  int w = pos >> 5, b = pos & 31;
  uint32_t lowmask = (b == 31) ? 0xFFFFFFFFu : ((2u << b) - 1u);
  bool found = false;
  uint32_t m = BND[w] & lowmask;
  if (m) {
    int bb = 31 - __clz(m);
    if (ALN[w] & lowmask & ~((1u << bb) - 1u)) return true;
    found = true;
  } else {
    if (ALN[w] & lowmask) return true;
    for (int ww = w - 1; ww >= 0; --ww) {
      m = BND[ww];
      if (m) {
        int bb = 31 - __clz(m);
        if (ALN[ww] & ~((1u << bb) - 1u)) return true;
        found = true;
        break;
      } else if (ALN[ww])
        return true;
    }
  }
  if (!found && fwd_in) return true;
  uint32_t highmask = ~lowmask;
  found = false;
  m = BND[w] & highmask;
  if (m) {
    int bb = __ffs(m) - 1;
    if (ALN[w] & highmask & ((1u << bb) - 1u)) return true;
    found = true;
  } else {
    if (ALN[w] & highmask) return true;
    for (int ww = w + 1; ww < NW; ++ww) {
      m = BND[ww];
      if (m) {
        int bb = __ffs(m) - 1;
        if (ALN[ww] & ((1u << bb) - 1u)) return true;
        found = true;
        break;
      } else if (ALN[ww])
        return true;
    }
  }
  if (!found && bwd_in) return true;
  return false;
}

constexpr int kSplitThreads = 256;
struct SplitSmem {
  TileSmem t;
  uint32_t NSA[kTileBytes / 32 + 1];  // alnum-run token starts
  uint32_t NEA[kTileBytes / 32 + 1];  // alnum-run token ends
  uint32_t NSO[kTileBytes / 32 + 1];  // other single-rune token starts (gated on the block's alnum flag)
  uint32_t NEO[kTileBytes / 32 + 2];  // their ends (may spill 3 bytes into the next tile)
  uint2 ends_local[kTileSlots];
  uint32_t n_ends_local, ends_base, wcnt;
  uint32_t red[8][4];
};

template <bool SUMMARY>
__device__ void split_tile(const JbTables& T, const SplitArgs& A, SplitSmem& S, const uint32_t tile) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t t0 = tile * (uint32_t)kTileBytes;
  const uint32_t n = A.n;
  classify_tile<kSplitThreads>(S.t, A.text, n, A.ds_bits, t0, T);
  TileCtx cx{&S.t, t0, n};
  const int NW = kTileBytes / 32;
  if (!SUMMARY) {
    if (tid == 0) {
      S.n_ends_local = 0;
      S.wcnt = 0;
    }
    for (int j = tid; j < NW + 2; j += kSplitThreads) S.NEO[j] = 0;
    __syncthreads();
  }
  // ---- per-byte predicates -> bit words -------------------------------------------------
  for (int g = warp; g < NW; g += kSplitThreads / 32) {
    const int i = kHaloL + g * 32 + lane;
    const uint32_t P = t0 + g * 32 + lane;
    const uint8_t c = S.t.cls[i];
    const uint32_t cl = c >> 3, len = c & 7;
    const bool start = c != 0 && P < n;
    const bool han = cl == CL_HAN;
    bool bnd = false;
    if (start) bnd = P == 0 || cx.ds_at(i) || (han != (prev_rune_class(S.t, i) == CL_HAN));
    const bool aln = start && cl == CL_ALNUM;
    uint32_t wb = __ballot_sync(FULL, bnd), wa = __ballot_sync(FULL, aln);
    if (lane == 0) {
      S.t.BND[g] = wb;
      S.t.ALN[g] = wa;
    }
    if (!SUMMARY) {
      bool nsa = false, nea = false, nso = false;
      if (start && !han) {
        if (cl == CL_ALNUM) {  // alnum run = one token (tokenizer.go:298-299)
          nsa = P == 0 || cx.ds_at(i) || !d_is_alnum(S.t.sb[i - 1]);
          nea = P + 1 >= n || cx.ds_at(i + 1) || !d_is_alnum(S.t.sb[i + 1]);
        } else if (cl != CL_SPACE) {  // every other rune is its own token; spaces are skipped (tokenizer.go:301-306)
          nso = true;
          int q = g * 32 + lane + (int)len - 1;
          atomicOr(&S.NEO[q >> 5], 1u << (q & 31));
        }
      }
      uint32_t w1 = __ballot_sync(FULL, nsa), w2 = __ballot_sync(FULL, nea), w3 = __ballot_sync(FULL, nso);
      if (lane == 0) {
        S.NSA[g] = w1;
        S.NEA[g] = w2;
        S.NSO[g] = w3;
      }
    }
  }
  __syncthreads();
  if (SUMMARY) {
    // tile summary for the segmented scan: bit0 has boundary, bit1 alnum before the first boundary,
    // bit2 alnum at/after the last boundary
    if (warp == 0) {
      int firstw = NW, lastw = -1;
      for (int j = lane; j < NW; j += 32)
        if (S.t.BND[j]) {
          firstw = min(firstw, j);
          lastw = max(lastw, j);
        }
      firstw = __reduce_min_sync(FULL, firstw);
      lastw = __reduce_max_sync(FULL, lastw);
      bool pre = false, post = false;
      if (lastw < 0) {
        for (int j = lane; j < NW; j += 32) pre |= S.t.ALN[j] != 0;
        post = pre;
      } else {
        uint32_t fb = __ffs(S.t.BND[firstw]) - 1, lb = 31 - __clz(S.t.BND[lastw]);
        for (int j = lane; j < NW; j += 32) {
          uint32_t a = S.t.ALN[j];
          if (j < firstw) pre |= a != 0;
          if (j == firstw) pre |= (a & ((1u << fb) - 1u)) != 0;
          if (j > lastw) post |= a != 0;
          if (j == lastw) post |= (a & ~((1u << lb) - 1u)) != 0;
        }
      }
      pre = __any_sync(FULL, pre);
      post = __any_sync(FULL, post);
      if (lane == 0) A.tile_sum[tile] = (uint8_t)((lastw >= 0 ? 1 : 0) | (pre ? 2 : 0) | (post ? 4 : 0));
    }
    return;
  }
  // ---- cutNonZh gating: other-rune tokens survive only if their block has an alnum ---------
  const uint8_t ctx = A.tile_ctx[tile];
  const bool fwd_in = ctx & 1, bwd_in = ctx & 2;
  for (int g = warp; g < NW; g += kSplitThreads / 32) {
    uint32_t so = S.NSO[g];
    bool mine = (so >> lane) & 1;
    bool drop = false;
    if (mine) drop = !block_has_alnum(S.t.BND, S.t.ALN, g * 32 + lane, fwd_in, bwd_in);
    if (drop) {
      int len = S.t.cls[kHaloL + g * 32 + lane] & 7;
      int q = g * 32 + lane + len - 1;
      atomicAnd(&S.NEO[q >> 5], ~(1u << (q & 31)));
    }
    uint32_t dm = __ballot_sync(FULL, drop);
    if (lane == 0) S.NSO[g] = so & ~dm;
  }
  __syncthreads();
  const uint32_t w0 = t0 / 32;
  for (int j = tid; j < NW + 1; j += kSplitThreads) {
    uint32_t sbits = j < NW ? (S.NSA[j] | S.NSO[j]) : 0;
    uint32_t ebits = (j < NW ? S.NEA[j] : 0) | S.NEO[j];
    if (sbits) atomicOr(&A.s_bits[w0 + j], sbits);
    if (ebits) atomicOr(&A.e_bits[w0 + j], ebits);
  }
  // ---- Han slots: block start/end flags and the DAG probe (buildDag, tokenizer.go:462-497) ----
  const uint32_t slot0 = tile * (uint32_t)kTileSlots;
  double* wtile = A.wbuf + (uint64_t)tile * A.w_per_tile;
  for (int it = 0; it < kTileSlots / kSplitThreads; it++) {
    const int kk = it * kSplitThreads + tid;
    const uint32_t k = slot0 + kk;
    int i = -1;
#pragma unroll
    for (int c = -2; c <= 0; c++) {
      int idx = kHaloL + 3 * kk + c;
      int64_t P = (int64_t)t0 + 3 * kk + c;
      if (P >= 0 && P < (int64_t)n && (S.t.cls[idx] >> 3) == CL_HAN) i = idx;
    }
    uint32_t mask = 0, cnt = 0;
    bool bstart = false, bend = false;
    double wv[4];
    bool overflow4 = false;
    uint32_t r0 = 0;
    int len0 = 0;
    uint32_t P0 = 0;
    if (i >= 0) {
      P0 = t0 - kHaloL + i;
      len0 = S.t.cls[i] & 7;
      bstart = P0 == 0 || cx.ds_at(i) || prev_rune_class(S.t, i) != CL_HAN;
      int q = i + len0;
      bend = (P0 + len0 >= n) || cx.ds_at(q) || (S.t.cls[q] >> 3) != CL_HAN;
      r0 = d_decode(&S.t.sb[i], len0);
      uint32_t d1 = d_slot(P0 + len0) - k;
      mask = 1u << (d1 - 1);
      // first probe: termFreq[string(iRune)] (tokenizer.go:468-472)
      uint32_t info, child, parent;
      uint32_t hs = r0 < 0x10000 ? JB_PARENT_FIRST(r0) : JB_PARENT_ROOT;  // hash state of the prefix (jb_hash_next)
      if (r0 < 0x10000) {
        const uint4 f = __ldg(reinterpret_cast<const uint4*>(T.first + r0));
        wv[0] = __longlong_as_double(((long long)f.y << 32) | (long long)f.x);
        info = f.z;
        child = f.w;
        parent = JB_PARENT_FIRST(r0);
      } else {
        double pw;
        uint32_t prb;
        int ps = jb_probe_edge(T.entries, T.hash_mask, T.hash_shift, hs, JB_PARENT_ROOT, r0, &pw, &prb);
        if (ps >= 0 && jb_w_positive(pw)) {
          wv[0] = pw;
          info = (uint32_t)JB_MAX_DELTA << 8;
        } else {
          wv[0] = ps >= 0 ? pw : T.neg_log_total;  // freq 0: -Inf; missing: log(1)-total
          info = JB_FIRST_GATE;
        }
        child = ps >= 0 ? (prb >> 21) : 0u;
        parent = (uint32_t)ps;
      }
      cnt = 1;
      if (!(info & JB_FIRST_GATE) && !bend) {
        const uint32_t maxlen = (info >> 8) & 0xFF;
        uint32_t L = 1;
        int qi = q;
        // for j := range textRunes[i:] ... break on the first missing prefix (tokenizer.go:473-482)
        while (L < maxlen) {
          if (qi + 4 > kRegion || cx.ds_at(qi)) break;
          uint8_t c = S.t.cls[qi];
          if ((c >> 3) != CL_HAN) break;
          int len = c & 7;
          uint32_t rl = d_decode(&S.t.sb[qi], len);
          // Bloom of the next rune: 32 bits in the first-rune table, 11 bits in hash entries
          const bool may = (L == 1 && r0 < 0x10000) ? ((child >> jb_bloom_bit(rl)) & 1) : ((child >> jb_bloom11(rl)) & 1);
          if (!may) break;
          double pw;
          uint32_t prb;
          int ps = jb_probe_edge(T.entries, T.hash_mask, T.hash_shift, hs, parent, rl, &pw, &prb);
          if (ps < 0) break;  // !found -> break (tokenizer.go:476-478)
          L++;
          qi += len;
          if (jb_w_positive(pw)) {  // val > 0 -> edge (tokenizer.go:479-481)
            uint32_t d = d_slot(t0 - kHaloL + qi) - k;
            mask |= 1u << (d - 1);
            if (cnt < 4) wv[cnt] = pw;
            else overflow4 = true;
            cnt++;
          }
          parent = (uint32_t)ps;
          child = prb >> 21;
        }
      }
    }
    // ---- place this group's weights (ascending position, ascending length) ---------------
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    uint32_t total = __shfl_sync(FULL, incl, 31);
    uint32_t gbase = 0;
    if (lane == 0 && total) gbase = atomicAdd(&S.wcnt, total);
    gbase = __shfl_sync(FULL, gbase, 0);
    uint32_t excl = gbase + incl - cnt;
    const uint32_t cap = A.w_per_tile;
    if (i >= 0) {
      if (excl + cnt <= cap) {
        if (!overflow4) {
          for (uint32_t j = 0; j < cnt; j++) wtile[excl + j] = wv[j];
        } else {
          // rare: more than 4 candidates -- walk the chain again and store as we go
          wtile[excl] = wv[0];
          uint32_t o = 1;
          uint32_t parent = JB_PARENT_FIRST(r0);
          uint32_t hs2 = r0 < 0x10000 ? JB_PARENT_FIRST(r0) : JB_PARENT_ROOT;
          if (r0 >= 0x10000) {
            double pw;
            uint32_t prb;
            parent = (uint32_t)jb_probe_edge(T.entries, T.hash_mask, T.hash_shift, hs2, JB_PARENT_ROOT, r0, &pw, &prb);
          }
          int qi = i + len0;
          while (o < cnt) {
            int len = S.t.cls[qi] & 7;
            uint32_t rl = d_decode(&S.t.sb[qi], len);
            double pw;
            uint32_t prb;
            int ps = jb_probe_edge(T.entries, T.hash_mask, T.hash_shift, hs2, parent, rl, &pw, &prb);
            if (ps < 0) break;
            qi += len;
            if (jb_w_positive(pw)) wtile[excl + o++] = pw;
            parent = (uint32_t)ps;
          }
        }
      }
      A.rec[k] = mask | (bstart ? JB_REC_START : 0u);
      if (bend) {
        uint32_t e = atomicAdd(&S.n_ends_local, 1u);
        uint32_t wend = min(excl + cnt, cap + 1024u);
        S.ends_local[e] = make_uint2(k, (uint32_t)(tile * (uint64_t)A.w_per_tile) + wend);
      }
    } else {
      A.rec[k] = 0;
    }
    if (lane == 0) A.gend[k >> 5] = (uint32_t)(tile * (uint64_t)A.w_per_tile) + min(gbase + total, cap + 1024u);
  }
  __syncthreads();
  if (tid == 0) {
    S.ends_base = S.n_ends_local ? atomicAdd(&A.counters[C_N_ENDS], S.n_ends_local) : 0;
    if (S.wcnt > A.w_per_tile) {
      atomicOr(&A.counters[C_STATUS], 1u);
      atomicMax(&A.counters[C_W_NEEDED], S.wcnt);
    }
  }
  __syncthreads();
  for (uint32_t j = tid; j < S.n_ends_local; j += kSplitThreads) A.ends[S.ends_base + j] = S.ends_local[j];
}

// Persistent over tiles: when the batch is not flagged for the general pipeline (mode 2) the launch costs
// a few microseconds instead of one CTA per tile.
template <bool SUMMARY>
__global__ void __launch_bounds__(kSplitThreads) k_split(const JbTables T, const SplitArgs A) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  SplitSmem& S = *reinterpret_cast<SplitSmem*>(smem_raw);
  if (A.mode == 2 && !(A.counters[C_FLAGS] & 1u)) return;
  for (uint32_t tile = blockIdx.x; tile < A.ntiles; tile += gridDim.x) {
    split_tile<SUMMARY>(T, A, S, tile);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// Segmented OR-scan over tile summaries.  Monoid element (has, pre, post):
//   combine(a,b) = (a.has|b.has, a.has ? a.pre : a.pre|b.pre, b.has ? b.post : a.post|b.post)
// ctx[t] bit0 = alnum between the last boundary before tile t and its start,
//        bit1 = alnum between tile t's end and the next boundary.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t seg_combine(uint32_t a, uint32_t b) {
  uint32_t has = (a | b) & 1;
  uint32_t pre = (a & 1) ? (a & 2) : ((a | b) & 2);
  uint32_t post = (b & 1) ? (b & 4) : ((a | b) & 4);
  return has | pre | post;
}
// identity: has=0, pre=post=0 combined with x gives x when treated as "no alnum, no boundary"
__global__ void __launch_bounds__(1024) k_tile_scan(const uint8_t* __restrict__ sum, uint8_t* __restrict__ ctx, uint32_t nt,
                                                    const uint32_t* __restrict__ counters, int require_flag) {
  if (require_flag && !(counters[C_FLAGS] & 1u)) return;
  __shared__ uint32_t part[1024];
  __shared__ uint32_t pfx[1024], sfx[1024];
  const int tid = threadIdx.x;
  const uint32_t per = (nt + 1023) / 1024;
  const uint32_t lo = min(nt, tid * per), hi = min(nt, lo + per);
  uint32_t acc = 0;
  for (uint32_t t = lo; t < hi; t++) acc = seg_combine(acc, sum[t]);
  part[tid] = acc;
  __syncthreads();
  if (tid == 0) {  // 1024 sequential combines: negligible
    uint32_t a = 0;
    for (int j = 0; j < 1024; j++) {
      pfx[j] = a;  // combine of all chunks before j
      a = seg_combine(a, part[j]);
    }
    a = 0;
    for (int j = 1023; j >= 0; j--) {
      sfx[j] = a;  // combine of all chunks after j
      a = seg_combine(part[j], a);
    }
  }
  __syncthreads();
  // forward: F = post of the prefix (every prefix holds the boundary at position 0)
  uint32_t a = pfx[tid];
  for (uint32_t t = lo; t < hi; t++) {
    ctx[t] = (a & 4) ? 1 : 0;
    a = seg_combine(a, sum[t]);
  }
  __syncthreads();
  a = sfx[tid];
  for (uint32_t t = hi; t-- > lo;) {
    if (a & 2) ctx[t] |= 2;
    a = seg_combine(sum[t], a);
  }
}

// ------------------------------------------------------------------------------------------
// Route DP (calcDagProba + maxIndexProba, tokenizer.go:502-548, 565-578), one lane per Han
// block, lanes refilled from a warp-level work queue.  R[i] = selector over the candidates
// (w + R[i+d]) in ascending order; only R of the last RING slots is live (shared-memory ring).
// ------------------------------------------------------------------------------------------
struct DpArgs {
  const uint8_t* text;
  uint32_t* rec;
  const uint32_t* gend;
  const double* wbuf;
  const uint2* ends;
  uint2* walks;
  uint32_t* counters;
  double* dbg_proba;
  int count_idx, cursor_idx;  // which counters hold the block count and the work cursor
  int require_flag;           // run only if C_FLAGS bit0 is set
};

__device__ __forceinline__ uint32_t lead_of_slot(const uint8_t* __restrict__ text, uint32_t k) {
  uint32_t p = 3u * k;
  if ((text[p] & 0xC0) != 0x80) return p;
  if ((text[p - 1] & 0xC0) != 0x80) return p - 1;
  return p - 2;
}
__device__ __forceinline__ uint32_t han_len(uint8_t lead) { return lead >= 0xF0 ? 4u : 3u; }

constexpr int kDpThreads = 128;
constexpr int kQueueBatch = 32;

template <int RING>
__global__ void __launch_bounds__(kDpThreads) k_route_dp(const DpArgs A) {
  __shared__ double ring[RING * kDpThreads];
  const int tid = threadIdx.x, lane = tid & 31;
  if (A.require_flag && !(A.counters[C_FLAGS] & 1u)) return;
  const uint32_t nblocks = A.counters[A.count_idx];
  uint32_t qh = 0, qt = 0;
  bool exhausted = false, active = false;
  uint32_t k = 0, wp = 0, grp = 0, idx = 0, pend = 0;
  for (;;) {
    unsigned need = __ballot_sync(FULL, !active);
    if (need && !exhausted) {
      if (qh == qt) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&A.counters[A.cursor_idx], (uint32_t)kQueueBatch);
        base = __shfl_sync(FULL, base, 0);
        if (base >= nblocks) exhausted = true;
        else {
          qh = base;
          qt = min(base + (uint32_t)kQueueBatch, nblocks);
        }
      }
      if (!exhausted) {
        uint32_t give = min((uint32_t)__popc(need), qt - qh);
        uint32_t rank = __popc(need & ((1u << lane) - 1u));
        if (!active && rank < give) {
          idx = qh + rank;
          uint2 e = A.ends[idx];
          k = e.x;
          wp = e.y;
          grp = k >> 5;
          uint32_t r = A.rec[k];
          uint32_t d1 = __ffs(r & JB_REC_MASK);
          ring[((k + d1) & (RING - 1)) * kDpThreads + tid] = 0.0;  // {j, 0.0} for j == len (tokenizer.go:522)
          uint32_t p = lead_of_slot(A.text, k);
          pend = p + han_len(A.text[p]);
          active = true;
        }
        qh += give;
      }
    }
    if (!__any_sync(FULL, active)) {
      if (exhausted) break;
      continue;
    }
    if (active) {
      if ((k >> 5) != grp) {
        grp = k >> 5;
        wp = A.gend[grp];
      }
      uint32_t r = A.rec[k];
      uint32_t m = r & JB_REC_MASK;
      if (m == 0) {  // hole after a 4-byte rune
        if (k == 0) {  // cannot happen on consistent records; never run off the array
          atomicOr(&A.counters[C_STATUS], 2u);
          active = false;
        } else {
          k--;
        }
      } else {
        uint32_t ncand = __popc(m);
        wp -= ncand;
        // maxIndexProba (tokenizer.go:565-578): NOT an argmax -- each candidate is compared with
        // the PREVIOUS candidate; the last one that is >= its predecessor wins, else the last one.
        double prev = JB_MINF, best_v = 0.0, v = 0.0;
        uint32_t best_d = 0, d = 0;
        const uint32_t d1 = __ffs(m);
        for (uint32_t j = 0; j < ncand; j++) {
          d = __ffs(m);
          m &= m - 1;
          double w = A.wbuf[wp + j];
          v = w + ring[((k + d) & (RING - 1)) * kDpThreads + tid];  // pieceFreq + nextBestPiece.proba (tokenizer.go:529)
          if (v >= prev) {
            best_d = d;
            best_v = v;
          }
          prev = v;
        }
        if (best_d == 0) {  // best.index == -1 -> return prev
          best_d = d;
          best_v = v;
        }
        ring[(k & (RING - 1)) * kDpThreads + tid] = best_v;
        A.rec[k] = (r & JB_REC_START) | best_d | (best_d == d1 ? JB_REC_SINGLE : 0u);
        if (A.dbg_proba) A.dbg_proba[k] = best_v;
        if (r & JB_REC_START) {
          A.walks[idx] = make_uint2(k, pend);
          active = false;
        } else if (k == 0) {
          atomicOr(&A.counters[C_STATUS], 2u);
          A.walks[idx] = make_uint2(k, pend);
          active = false;
        } else {
          k--;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Path walk (findDagPath, tokenizer.go:552-562) + HMM glue (cutZh, tokenizer.go:221-255) +
// Viterbi (tokenizer.go:668-756) + cutHMM (tokenizer.go:273-285).  One lane per Han block.
// ------------------------------------------------------------------------------------------
struct WalkArgs {
  const uint8_t* text;
  uint32_t* rec;
  const uint2* walks;
  uint32_t* counters;
  uint32_t* s_bits;
  uint32_t* e_bits;
  int count_idx, cursor_idx, require_flag;
};

__device__ __forceinline__ void set_bit(uint32_t* bits, uint32_t p) { atomicOr(&bits[p >> 5], 1u << (p & 31)); }

__device__ __forceinline__ void load_emit(const JbTables& T, uint32_t cp, double e[4]) {
  if (cp < 0x10000) {
    const double2* p = reinterpret_cast<const double2*>(T.emit + (size_t)cp * 4);
    double2 a = __ldg(p), b = __ldg(p + 1);
    e[0] = a.x; e[1] = a.y; e[2] = b.x; e[3] = b.y;
    return;
  }
  e[0] = e[1] = e[2] = e[3] = JB_MINF;
  int lo = 0, hi = (int)T.n_emit_supp - 1;
  while (lo <= hi) {
    int mid = (lo + hi) >> 1;
    uint32_t r = __ldg(T.emit_supp_rune + mid);
    if (r == cp) {
      for (int s = 0; s < 4; s++) e[s] = __ldg(T.emit_supp + (size_t)mid * 4 + s);
      return;
    }
    if (r < cp) lo = mid + 1;
    else hi = mid - 1;
  }
}

__device__ __forceinline__ uint32_t next_rune_slot(const uint32_t* rec, uint32_t k) { return rec[k + 1] ? k + 1 : k + 2; }

// viterbi's tail + cutHMM for the run of single runes at slots [ks .. kl] (n runes), values V.
__device__ void flush_run(const WalkArgs& A, uint32_t ks, uint32_t kl, uint32_t nrun, const double V[4]) {
  if (nrun == 1) {  // viterbi returns ["S"] (tokenizer.go:672-674)
    uint32_t p = lead_of_slot(A.text, ks);
    set_bit(A.s_bits, p);
    set_bit(A.e_bits, p + han_len(A.text[p]) - 1);
    return;
  }
  const int PREV[4][2] = {{2, 3}, {0, 1}, {0, 1}, {2, 3}};  // stateChange (tokenizer.go:24-29)
  int st = V[2] > V[3] ? 2 : 3;                             // e > s ? E : S (tokenizer.go:723-729)
  // pass 1: back-trace = fullPath[st]; it stops early where route.from == "" (tokenizer.go:715-716)
  uint32_t k = kl, plen = 0;
  for (;;) {
    uint32_t r = A.rec[k];
    if (r == 0) {
      k--;
      continue;
    }
    A.rec[k] = (r & ~JB_REC_ES) | ((st >= 2) ? JB_REC_ES : 0u);
    plen++;
    if (k == ks) break;
    int c = (r >> (16 + 2 * st)) & 3;
    if (c == 0) break;
    st = PREV[st][c - 1];
    k--;
  }
  // pass 2: cutHMM iterates the PATH and indexes the text with the path index (tokenizer.go:277-283):
  // path[j] applies to rune j, j < len(path); a short path drops the run's tail.
  uint32_t src = ks;
  for (uint32_t j = 0; j < nrun - plen; j++) src = next_rune_slot(A.rec, src);
  uint32_t dst = ks;
  bool prev_es = true;
  for (uint32_t j = 0; j < plen; j++) {
    bool es = A.rec[src] & JB_REC_ES;
    uint32_t p = lead_of_slot(A.text, dst);
    if (prev_es) set_bit(A.s_bits, p);
    if (es) set_bit(A.e_bits, p + han_len(A.text[p]) - 1);
    prev_es = es;
    if (j + 1 < plen) {
      src = next_rune_slot(A.rec, src);
      dst = next_rune_slot(A.rec, dst);
    }
  }
}

constexpr int kWalkThreads = 128;

template <bool HMM>
__global__ void __launch_bounds__(kWalkThreads) k_walk(const JbTables T, const WalkArgs A) {
  const int lane = threadIdx.x & 31;
  if (A.require_flag && !(A.counters[C_FLAGS] & 1u)) return;
  const uint32_t nblocks = A.counters[A.count_idx];
  uint32_t qh = 0, qt = 0;
  bool exhausted = false, active = false;
  uint32_t k = 0, pend = 0, kvend = 0, pcur = 0;
  uint32_t run_n = 0, run_s = 0, run_l = 0;
  double V[4];
  for (;;) {
    unsigned need = __ballot_sync(FULL, !active);
    if (need && !exhausted) {
      if (qh == qt) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&A.counters[A.cursor_idx], (uint32_t)kQueueBatch);
        base = __shfl_sync(FULL, base, 0);
        if (base >= nblocks) exhausted = true;
        else {
          qh = base;
          qt = min(base + (uint32_t)kQueueBatch, nblocks);
        }
      }
      if (!exhausted) {
        uint32_t give = min((uint32_t)__popc(need), qt - qh);
        uint32_t rank = __popc(need & ((1u << lane) - 1u));
        if (!active && rank < give) {
          uint2 e = A.walks[qh + rank];
          k = e.x;
          pend = e.y;
          kvend = d_slot(pend);
          pcur = lead_of_slot(A.text, k);
          run_n = 0;
          active = true;
        }
        qh += give;
      }
    }
    if (!__any_sync(FULL, active)) {
      if (exhausted) break;
      continue;
    }
    if (active) {
      const uint32_t r = A.rec[k];
      const uint32_t d = r & 0xFF;
      if (d == 0) {  // cannot happen on consistent records; never spin
        atomicOr(&A.counters[C_STATUS], 2u);
        active = false;
        continue;
      }
      const uint32_t knext = k + d;
      const uint32_t pnext = knext >= kvend ? pend : lead_of_slot(A.text, knext);
      if (HMM && (r & JB_REC_SINGLE)) {
        // collect singletons for HMM segmentation (tokenizer.go:233-234): one Viterbi step per rune
        uint32_t cp = d_decode(A.text + pcur, (int)(pnext - pcur));
        double em[4];
        load_emit(T, cp, em);
        if (run_n == 0) {
          run_s = k;
#pragma unroll
          for (int s = 0; s < 4; s++) V[s] = T.start[s] + em[s];  // tokenizer.go:688-695
        } else {
          // stateTransitionRoute (tokenizer.go:736-756): strict > starting from minFloat, list order
          double W[4];
          uint32_t code = 0;
#pragma unroll
          for (int s = 0; s < 4; s++) {
            const int pa = (s == 0 || s == 3) ? 2 : 0, pb = (s == 0 || s == 3) ? 3 : 1;
            double r0 = V[pa] + T.trans[s][0], r1 = V[pb] + T.trans[s][1];
            double best = JB_MINF;
            uint32_t from = 0;
            if (r0 > best) {
              best = r0;
              from = 1;
            }
            if (r1 > best) {
              best = r1;
              from = 2;
            }
            W[s] = best + em[s];  // tokenizer.go:712
            code |= from << (2 * s);
          }
#pragma unroll
          for (int s = 0; s < 4; s++) V[s] = W[s];
          A.rec[k] = r | (code << 16);
        }
        run_l = k;
        run_n++;
      } else {
        if (HMM && run_n) {
          flush_run(A, run_s, run_l, run_n, V);
          run_n = 0;
        }
        set_bit(A.s_bits, pcur);
        set_bit(A.e_bits, pnext - 1);
      }
      if (knext >= kvend) {
        if (HMM && run_n) {
          flush_run(A, run_s, run_l, run_n, V);
          run_n = 0;
        }
        active = false;
      } else {
        k = knext;
        pcur = pnext;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// Token ranking: start/end bitmaps -> (start,end) arrays in document order, doc-relative.
// Rank tiles of 32 KiB of text (1024 bitmap words, one per thread).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kRankWords) k_rank_count(const uint32_t* __restrict__ s_bits, const uint32_t* __restrict__ e_bits,
                                                          uint32_t nwords, uint32_t* __restrict__ cnt) {
  __shared__ uint32_t red[2][kRankWords / 32];
  const uint32_t w = blockIdx.x * kRankWords + threadIdx.x;
  uint32_t c = w < nwords ? __popc(__ldg(s_bits + w)) : 0, e = w < nwords ? __popc(__ldg(e_bits + w)) : 0;
  c = __reduce_add_sync(FULL, c);
  e = __reduce_add_sync(FULL, e);
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = c;
    red[1][threadIdx.x >> 5] = e;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    c = __reduce_add_sync(FULL, red[0][threadIdx.x]);
    e = __reduce_add_sync(FULL, red[1][threadIdx.x]);
    if (threadIdx.x == 0) {
      cnt[2 * blockIdx.x] = c;
      cnt[2 * blockIdx.x + 1] = e;
    }
  }
}

__global__ void __launch_bounds__(1024) k_rank_scan(uint32_t* __restrict__ cnt, uint32_t nt, uint32_t* __restrict__ counters,
                                                    uint64_t* __restrict__ d_n_tokens, const uint32_t* __restrict__ doc_off32,
                                                    uint64_t ndocs, uint32_t n, uint64_t* __restrict__ doc_tok, uint64_t tok_base) {
  __shared__ uint32_t part[2][1024];
  __shared__ uint32_t total_s;
  const int tid = threadIdx.x;
  const uint32_t per = (nt + 1023) / 1024;
  const uint32_t lo = min(nt, tid * per), hi = min(nt, lo + per);
  uint32_t acc = 0, acce = 0;
  for (uint32_t t = lo; t < hi; t++) {
    acc += cnt[2 * t];
    acce += cnt[2 * t + 1];
  }
  part[0][tid] = acc;
  part[1][tid] = acce;
  __syncthreads();
  if (tid == 0) {
    uint32_t a = 0, e = 0;
    for (int j = 0; j < 1024; j++) {
      uint32_t v = part[0][j], ve = part[1][j];
      part[0][j] = a;
      part[1][j] = e;
      a += v;
      e += ve;
    }
    total_s = a;
    if (a != e) atomicOr(&counters[C_STATUS], 4u);  // every token has one start and one end bit
    counters[C_N_TOKENS] = a;
    d_n_tokens[0] = a;
    d_n_tokens[1] = counters[C_STATUS];
  }
  __syncthreads();
  uint32_t a = part[0][tid], e = part[1][tid];
  for (uint32_t t = lo; t < hi; t++) {
    uint32_t v = cnt[2 * t], ve = cnt[2 * t + 1];
    cnt[2 * t] = a;
    cnt[2 * t + 1] = e;
    a += v;
    e += ve;
  }
  // documents that start at or after the end of the text (empty tail documents) and the sentinel
  if (doc_tok) {
    uint64_t lo_d = 0, hi_d = ndocs + 1;  // lower_bound(doc_off32, n)
    while (lo_d < hi_d) {
      uint64_t mid = (lo_d + hi_d) >> 1;
      if (doc_off32[mid] < n) lo_d = mid + 1;
      else hi_d = mid;
    }
    for (uint64_t d = lo_d + tid; d <= ndocs; d += 1024) doc_tok[d] = tok_base + total_s;
  }
}

constexpr int kRankStage = 256;  // tokens a warp stages per pass
__global__ void __launch_bounds__(kRankWords) k_rank_scatter(const uint32_t* __restrict__ s_bits, const uint32_t* __restrict__ e_bits,
                                                            const uint32_t* __restrict__ ds_bits, uint32_t nwords, uint32_t n,
                                                            const uint32_t* __restrict__ tile_base, const uint32_t* __restrict__ doc_off32,
                                                            const uint32_t* __restrict__ tile_first_doc, uint64_t ndocs, uint32_t* __restrict__ out_start, uint32_t* __restrict__ out_end,
                                                            uint64_t cap, uint64_t* __restrict__ doc_tok, uint64_t tok_base) {
  __shared__ uint32_t sS[kRankWords], sPS[kRankWords];
  __shared__ uint32_t wsum[2][kRankWords / 32];
  __shared__ int32_t wmax[kRankWords / 32];
  __shared__ uint32_t stage[kRankWords / 32][kRankStage];
  __shared__ uint32_t s_dpos0;
  __shared__ uint64_t s_dlo;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t w = tile * kRankWords + tid;
  const uint32_t t0 = tile * (uint32_t)kRankBytes;
  const uint32_t S = w < nwords ? __ldg(s_bits + w) : 0, E = w < nwords ? __ldg(e_bits + w) : 0, D = w < nwords ? __ldg(ds_bits + w) : 0;
  if (tid == 0) {
    // first document index with doc_off >= t0 (from k_docstart), last document start <= t0
    uint64_t ub = tile_first_doc[tile];
    s_dlo = ub;
    while (ub <= ndocs && doc_off32[ub] == t0) ub++;  // (empty documents share an offset)
    s_dpos0 = ub ? doc_off32[ub - 1] : 0;
  }
  // exclusive prefix of popc(S), popc(E) and running "last doc start" over the tile's words
  const uint32_t cs = __popc(S), ce = __popc(E);
  const int32_t ld = D ? (int32_t)(tid * 32 + 31 - __clz(D)) : -1;  // tile-local position of the word's last doc start
  uint32_t is = cs, ie = ce;
  int32_t im = ld;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t a = __shfl_up_sync(FULL, is, o), b = __shfl_up_sync(FULL, ie, o);
    int32_t c = __shfl_up_sync(FULL, im, o);
    if (lane >= o) {
      is += a;
      ie += b;
      im = max(im, c);
    }
  }
  if (lane == 31) {
    wsum[0][warp] = is;
    wsum[1][warp] = ie;
    wmax[warp] = im;
  }
  __syncthreads();
  if (warp == 0) {  // exclusive prefix over the 32 warps
    uint32_t a = wsum[0][lane], b = wsum[1][lane];
    int32_t c = wmax[lane];
    uint32_t ia = a, ib = b;
    int32_t ic = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t x = __shfl_up_sync(FULL, ia, o), y = __shfl_up_sync(FULL, ib, o);
      int32_t z = __shfl_up_sync(FULL, ic, o);
      if (lane >= o) {
        ia += x;
        ib += y;
        ic = max(ic, z);
      }
    }
    int32_t pc = __shfl_up_sync(FULL, ic, 1);
    wsum[0][lane] = ia - a;
    wsum[1][lane] = ib - b;
    wmax[lane] = lane ? pc : -1;
  }
  __syncthreads();
  const uint32_t ls = wsum[0][warp] + is - cs, le = wsum[1][warp] + ie - ce;  // tile-local rank of this word's first start / end
  const uint32_t base_s = tile_base[2 * tile], base_e = tile_base[2 * tile + 1];
  // last doc start strictly before this word (tile-local), or -1
  int32_t prev_ld = __shfl_up_sync(FULL, im, 1);
  if (lane == 0) prev_ld = -1;
  prev_ld = max(prev_ld, wmax[warp]);
  sS[tid] = S;
  sPS[tid] = base_s + ls;
  const uint32_t dpos0 = s_dpos0;
  const uint32_t wpos = t0 + tid * 32;
  // The tokens of a warp's 32 words have consecutive ranks: stage them in the warp's slice of shared memory in
  // rank order and write them out 128 contiguous bytes per instruction (scattered 4-byte stores from the bit loops
  // cost one L1 wavefront per token).
  {
    uint32_t* const wst = stage[warp];
    const uint32_t dbase = prev_ld >= 0 ? t0 + (uint32_t)prev_ld : dpos0;  // last document start before this word
    const uint32_t wls = __shfl_sync(FULL, ls, 0), wle = __shfl_sync(FULL, le, 0);        // tile-local rank of the warp's first start / end
    const uint32_t wts = __shfl_sync(FULL, is, 31), wte = __shfl_sync(FULL, ie, 31);      // tokens of the warp
#pragma unroll 1
    for (int which = 0; which < 2; which++) {
      const uint32_t bits = which ? E : S, lr0 = (which ? le - wle : ls - wls), total = which ? wte : wts;
      const uint64_t gbase = (uint64_t)(which ? base_e + wle : base_s + wls);
      uint32_t* __restrict__ out = which ? out_end : out_start;
      for (uint32_t c0 = 0; c0 < total; c0 += kRankStage) {
        uint32_t m = bits, lr = lr0;
        if (D == 0) {  // no document starts in this word (almost always): one base for all its tokens
          const uint32_t rel = wpos + (uint32_t)which - dbase;
          while (m) {
            const uint32_t b = __ffs(m) - 1;
            m &= m - 1;
            if (lr - c0 < (uint32_t)kRankStage) wst[lr - c0] = rel + b;
            lr++;
          }
        }
        while (m) {
          const uint32_t b = __ffs(m) - 1;
          m &= m - 1;
          if (lr - c0 < (uint32_t)kRankStage) {
            const uint32_t dm = D & ((b == 31) ? 0xFFFFFFFFu : ((2u << b) - 1u));
            const uint32_t dpos = dm ? wpos + (31 - __clz(dm)) : dbase;
            wst[lr - c0] = wpos + b + (uint32_t)which - dpos;
          }
          lr++;
        }
        __syncwarp();
        const uint32_t cn = min((uint32_t)kRankStage, total - c0);
        for (uint32_t i = lane; i < cn; i += 32) {
          const uint64_t r = gbase + c0 + i;
          if (r < cap) out[r] = wst[i];
        }
        __syncwarp();
      }
    }
  }
  // doc_tok_off for the documents that start inside this tile
  if (doc_tok) {
    __syncthreads();
    const uint32_t t1 = min(n, t0 + (uint32_t)kRankBytes);
    for (uint64_t d = s_dlo + tid; d <= ndocs; d += kRankWords) {
      uint32_t p = doc_off32[d];
      if (p >= t1) break;
      uint32_t lw = (p - t0) >> 5, lb = (p - t0) & 31;
      doc_tok[d] = tok_base + sPS[lw] + __popc(sS[lw] & ((1u << lb) - 1u));
    }
  }
}

// doc_tok_off without the scatter (bitmap results): one warp per document, rank of its first byte among the token starts
__global__ void __launch_bounds__(256) k_doc_tok(const uint32_t* __restrict__ s_bits, const uint32_t* __restrict__ tile_base,
                                                 const uint32_t* __restrict__ doc_off32, uint64_t ndocs, uint32_t n,
                                                 uint64_t* __restrict__ doc_tok, uint64_t tok_base) {
  const uint64_t d = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (d >= ndocs) return;
  const uint32_t p = doc_off32[d];
  if (p >= n) return;  // (documents at the end of the text: k_rank_scan)
  const uint32_t tile = p / (uint32_t)kRankBytes, w0 = tile * (uint32_t)kRankWords, w1 = p >> 5;
  uint32_t c = 0;
  for (uint32_t w = w0 + lane; w < w1; w += 32) c += __popc(__ldg(s_bits + w));
  c = __reduce_add_sync(FULL, c);
  if (lane == 0) doc_tok[d] = tok_base + tile_base[2 * tile] + c + __popc(__ldg(s_bits + w1) & ((1u << (p & 31)) - 1u));
}

// ------------------------------------------------------------------------------------------
// Helpers around the streaming fast path (jb_stream.cu)
// ------------------------------------------------------------------------------------------
// gated non-Han tokens whose block left the tile: keep them iff the scan found an alnum on an open side
__global__ void k_resolve_deferred(const uint4* __restrict__ deferred, const uint32_t* __restrict__ counters, uint32_t cap,
                                   const uint8_t* __restrict__ ctx, uint32_t* __restrict__ s_bits, uint32_t* __restrict__ e_bits) {
  uint32_t nd = min(counters[C_N_DEFER], cap);
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nd; i += gridDim.x * blockDim.x) {
    uint4 d = deferred[i];
    uint8_t c = ctx[d.w];
    if (((d.z & 1u) && (c & 1)) || ((d.z & 2u) && (c & 2))) {
      atomicOr(&s_bits[d.x >> 5], 1u << (d.x & 31));
      uint32_t q = d.x + d.y - 1;
      atomicOr(&e_bits[q >> 5], 1u << (q & 31));
    }
  }
}

// the batch was flagged for the general pipeline: forget what the fast path produced
__global__ void k_fallback_reset(uint32_t* __restrict__ counters, uint32_t* __restrict__ s_bits, uint32_t* __restrict__ e_bits,
                                 uint32_t nwords) {
  if (!(counters[C_FLAGS] & 1u)) return;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += gridDim.x * blockDim.x) {
    s_bits[i] = 0;
    e_bits[i] = 0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    counters[C_N_ENDS] = 0;
    counters[C_CUR_DP] = 0;
    counters[C_CUR_WALK] = 0;
  }
}

// ------------------------------------------------------------------------------------------
// debug: one dictionary lookup through the device tables
// ------------------------------------------------------------------------------------------
__global__ void k_debug_lookup(const JbTables T, const uint32_t* runes, int L, int* kind, double* w) {
  // walks the key the way buildDag reaches it: through every prefix (a key whose prefix is missing is unreachable)
  uint32_t r0 = runes[0];
  uint32_t parent;
  uint32_t hs = r0 < 0x10000 ? JB_PARENT_FIRST(r0) : JB_PARENT_ROOT;
  double cw = 0;
  int k = 0;
  if (r0 < 0x10000) {
    JbFirst f = T.first[r0];
    if (!(f.info & JB_FIRST_GATE)) k = 2;
    else k = (f.w == T.neg_log_total) ? 0 : 1;
    cw = f.w;
    parent = JB_PARENT_FIRST(r0);
  } else {
    uint32_t rb;
    int ps = jb_probe_edge(T.entries, T.hash_mask, T.hash_shift, hs, JB_PARENT_ROOT, r0, &cw, &rb);
    k = ps < 0 ? 0 : (jb_w_positive(cw) ? 2 : 1);
    parent = (uint32_t)ps;
  }
  for (int j = 1; j < L && k != 0; j++) {
    uint32_t rb;
    int ps = jb_probe_edge(T.entries, T.hash_mask, T.hash_shift, hs, parent, runes[j], &cw, &rb);
    k = ps < 0 ? 0 : (jb_w_positive(cw) ? 2 : 1);
    parent = (uint32_t)ps;
  }
  *kind = k;
  *w = k ? cw : 0.0;
}

int debug_lookup(const JbTables& T, const uint32_t* runes_host, int L, int* kind, double* w) {
  uint32_t* d_r;
  int* d_k;
  double* d_w;
  if (cudaMalloc(&d_r, L * 4) != cudaSuccess) return JB_ECUDA;
  cudaMalloc(&d_k, 4);
  cudaMalloc(&d_w, 8);
  cudaMemcpy(d_r, runes_host, L * 4, cudaMemcpyHostToDevice);
  JB_LAUNCH(k_debug_lookup, 1, 1, 0, 0, T, d_r, L, d_k, d_w);
  cudaMemcpy(kind, d_k, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(w, d_w, 8, cudaMemcpyDeviceToHost);
  cudaFree(d_r);
  cudaFree(d_k);
  cudaFree(d_w);
  return cudaGetLastError() == cudaSuccess ? JB_OK : JB_ECUDA;
}

// ------------------------------------------------------------------------------------------
// Workspace + pipeline
// ------------------------------------------------------------------------------------------
template <typename T>
static bool dalloc(T*& p, uint64_t count) {
  if (p) cudaFree(p);
  p = nullptr;
  return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T)) == cudaSuccess;
}

void workspace_free(Workspace& ws) {
  void* ptrs[] = {ws.text, ws.doc_off64, ws.doc_off32, ws.ds_bits, ws.s_bits, ws.e_bits, ws.m_bits, ws.segs, ws.land, ws.longs, ws.rec, ws.gend, ws.wbuf, ws.ends,
                  ws.walks, ws.tile_sum, ws.tile_ctx, ws.deferred, ws.tile_last_hs, ws.tile_first_doc, ws.wide_list, ws.path, ws.bp, ws.rank_cnt, ws.counters, ws.dbg_proba, ws.out_start, ws.out_end,
                  ws.out_doc_tok, ws.out_ntok};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  for (auto& e : ws.ev)
    if (e) cudaEventDestroy(e);
  if (ws.aux_stream) cudaStreamDestroy(ws.aux_stream);
  if (ws.ev_fork) cudaEventDestroy(ws.ev_fork);
  if (ws.ev_join) cudaEventDestroy(ws.ev_join);
  ws = Workspace();
}

int workspace_reserve(Workspace& ws, uint64_t nbytes, uint64_t ndocs, double w_per_slot, bool host_staging) {
  uint32_t wpt = (uint32_t)(w_per_slot * kTileSlots + 0.5);
  if (wpt < (uint32_t)kTileSlots) wpt = kTileSlots;
  bool ok = true;
  if (nbytes > ws.cap_bytes || wpt != ws.w_per_tile) {
    uint64_t cap = ws.cap_bytes > nbytes ? ws.cap_bytes : nbytes;
    cap = (cap + kRankBytes) / kRankBytes * kRankBytes;
    uint64_t ntiles = cap / kTileBytes + 2;
    uint64_t nwords = cap / 32 + 8;
    ok = ok && dalloc(ws.ds_bits, nwords) && dalloc(ws.s_bits, nwords) && dalloc(ws.e_bits, nwords) && dalloc(ws.m_bits, nwords);
    // a long block has >= 512 runes and is cut every 256: at most 1.5 segments per 256 runes = 768 bytes
    ws.segs_cap = (uint32_t)(cap / 512 + 8);
    ws.longs_cap = (uint32_t)(cap / 1536 + 8);
    ok = ok && dalloc(ws.segs, (uint64_t)ws.segs_cap) && dalloc(ws.land, (uint64_t)ws.segs_cap) && dalloc(ws.longs, (uint64_t)ws.longs_cap);
    ok = ok && dalloc(ws.rec, ntiles * kTileSlots + 64) && dalloc(ws.gend, ntiles * (kTileSlots / 32) + 8);
    ok = ok && dalloc(ws.wbuf, ntiles * (uint64_t)wpt + 4096);
    ok = ok && dalloc(ws.ends, ntiles * kTileSlots + 8) && dalloc(ws.walks, ntiles * kTileSlots + 8);
    ok = ok && dalloc(ws.tile_sum, ntiles + 8) && dalloc(ws.tile_ctx, ntiles + 8);
    ws.deferred_cap = (uint32_t)(cap / 64 + 4096);
    if (cap <= 65536) ws.deferred_cap = (uint32_t)cap + 64;  // (one gated token per byte at most: cannot overflow)
    ok = ok && dalloc(ws.deferred, (uint64_t)ws.deferred_cap);
    ws.wide_cap = (uint32_t)(cap / 256 + 4096);
    ok = ok && dalloc(ws.wide_list, (uint64_t)ws.wide_cap);
    ok = ok && dalloc(ws.tile_last_hs, cap / 8128 + 8) && dalloc(ws.path, cap / 12 + 64) && dalloc(ws.bp, cap / 3 + 64);
    ws.blocks_cap = (uint32_t)std::min<uint64_t>(ntiles * kTileSlots + 8, 0xFFFFFFF0ull);
    ok = ok && dalloc(ws.rank_cnt, 2 * (cap / kRankBytes + 8)) && dalloc(ws.tile_first_doc, cap / kRankBytes + 8);
    if (!ws.counters) ok = ok && dalloc(ws.counters, (uint64_t)C_NUM);
    if (host_staging) ok = ok && dalloc(ws.text, cap + 64);
    if (!ws.out_ntok) ok = ok && dalloc(ws.out_ntok, 2);
    ws.cap_bytes = cap;
    ws.w_per_tile = wpt;
  } else if (host_staging && !ws.text) {
    ok = ok && dalloc(ws.text, ws.cap_bytes + 64);
  }
  if (ndocs + 1 > ws.cap_docs) {
    uint64_t cd = ndocs + 1 + (ndocs >> 2) + 16;
    ok = ok && dalloc(ws.doc_off32, cd);
    if (host_staging) ok = ok && dalloc(ws.doc_off64, cd) && dalloc(ws.out_doc_tok, cd);
    ws.cap_docs = cd;
  } else if (host_staging && !ws.doc_off64) {
    ok = ok && dalloc(ws.doc_off64, ws.cap_docs) && dalloc(ws.out_doc_tok, ws.cap_docs);
  }
  if (!ok) {
    workspace_free(ws);
    return JB_ENOMEM;
  }
  return JB_OK;
}

const char* const kProfKernelNames[kNumProfKernels] = {"k_docstart+memset",
                                                       "k_scan",
                                                       "k_route (+k_wide)",
                                                       "k_emit (k_tile_scan+k_resolve_deferred on a side stream)",
                                                       "general pipeline (flagged batches)",
                                                       "k_rank_count",
                                                       "k_rank_scan",
                                                       "k_rank_scatter"};

void profile_collect(Workspace& ws) {
  if (!ws.prof || !ws.prof_pending) return;
  cudaEventSynchronize(ws.ev[kNumProfKernels]);
  for (int i = 0; i < kNumProfKernels; i++) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ws.ev[i], ws.ev[i + 1]) == cudaSuccess) ws.prof_ms[i] += ms;
  }
  ws.prof_steps++;
  ws.prof_pending = false;
}
#define PROF(i)                                    \
  do {                                             \
    if (ws.prof) cudaEventRecord(ws.ev[i], st);    \
  } while (0)

static int g_num_sms = 0;

static bool g_attr_done = false;

int run_pipeline(const JbTables& T, Workspace& ws_in, const uint8_t* d_text, uint32_t n, const uint64_t* d_doc_off, uint64_t ndocs,
                 bool use_hmm, const PipeOut& out, cudaStream_t st, int path) {
  const bool force_general = path == PATH_GENERAL;
  // the caller's bitmaps stand in for the workspace's (a shallow copy of the pointer set: nothing is owned twice)
  Workspace wsv;
  Workspace* wsp = &ws_in;
  if (out.d_s_bits && out.d_e_bits) {
    wsv = ws_in;
    wsv.s_bits = out.d_s_bits;
    wsv.e_bits = out.d_e_bits;
    wsp = &wsv;
  }
  Workspace& ws = *wsp;
  uint32_t* const d_start = out.d_start;
  uint32_t* const d_end = out.d_end;
  uint64_t* const d_doc_tok_off = out.d_doc_tok_off;
  uint64_t* const d_n_tokens = out.d_n_tokens;
  const uint64_t tok_base = out.tok_base, cap_tokens = out.cap_tokens;
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  if (!g_attr_done) {
    cudaFuncSetAttribute(k_split<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplitSmem));
    cudaFuncSetAttribute(k_split<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SplitSmem));
    g_attr_done = true;
  }
  const uint32_t nwords = (n + 31) / 32;
  const uint32_t ntiles = (n + kTileBytes - 1) / kTileBytes;
  const uint32_t nrt = (n + kRankBytes - 1) / kRankBytes;
  if (ws.prof) {
    profile_collect(ws);
    for (int i = 0; i <= kNumProfKernels; i++)
      if (!ws.ev[i]) cudaEventCreate(&ws.ev[i]);
  }
  PROF(0);
  cudaMemsetAsync(ws.counters, 0, C_NUM * sizeof(uint32_t), st);
  cudaMemsetAsync(ws.ds_bits, 0, ((uint64_t)nwords + 4) * 4, st);
  cudaMemsetAsync(ws.s_bits, 0, ((uint64_t)nwords + 4) * 4, st);
  cudaMemsetAsync(ws.e_bits, 0, ((uint64_t)nwords + 4) * 4, st);
  JB_LAUNCH(k_docstart, (unsigned)((ndocs + 1 + 255) / 256), 256, 0, st, d_doc_off, ndocs, n, out.pos0, ws.doc_off32, ws.ds_bits, ws.tile_first_doc);
  PROF(1);
  if (n > 0) {
    const unsigned pgrid = (unsigned)g_num_sms * 8;
    const unsigned sgrid = std::min<unsigned>(ntiles, (unsigned)g_num_sms * 8);
    SplitArgs sa;
    sa.text = d_text;
    sa.n = n;
    sa.ds_bits = ws.ds_bits;
    sa.s_bits = ws.s_bits;
    sa.e_bits = ws.e_bits;
    sa.rec = ws.rec;
    sa.gend = ws.gend;
    sa.wbuf = ws.wbuf;
    sa.w_per_tile = ws.w_per_tile;
    sa.ends = ws.ends;
    sa.counters = ws.counters;
    sa.ntiles = ntiles;
    sa.tile_ctx = ws.tile_ctx;
    sa.tile_sum = ws.tile_sum;
    DpArgs da;
    da.text = d_text;
    da.rec = ws.rec;
    da.gend = ws.gend;
    da.wbuf = ws.wbuf;
    da.ends = ws.ends;
    da.walks = ws.walks;
    da.counters = ws.counters;
    da.dbg_proba = ws.dbg_proba;
    WalkArgs wa;
    wa.text = d_text;
    wa.rec = ws.rec;
    wa.walks = ws.walks;
    wa.counters = ws.counters;
    wa.s_bits = ws.s_bits;
    wa.e_bits = ws.e_bits;
    auto launch_dp = [&]() {
      if (T.max_delta + 1 <= 8) JB_LAUNCH(k_route_dp<8>, pgrid, kDpThreads, 0, st, da);
      else if (T.max_delta + 1 <= 16) JB_LAUNCH(k_route_dp<16>, pgrid, kDpThreads, 0, st, da);
      else JB_LAUNCH(k_route_dp<32>, pgrid, kDpThreads, 0, st, da);
    };
    auto launch_walk = [&]() {
      if (use_hmm) JB_LAUNCH(k_walk<true>, pgrid, kWalkThreads, 0, st, T, wa);
      else JB_LAUNCH(k_walk<false>, pgrid, kWalkThreads, 0, st, T, wa);
    };
    if (!force_general) {
      // ---- fast path: k_scan -> k_route -> k_emit (jb_stream.cu) ----------------------------------
      const uint32_t nt1 = scan_tiles(n);
      cudaMemsetAsync(ws.path, 0, ((uint64_t)n / 12 + 16) * 4, st);
      if (use_hmm && n >= 1536) cudaMemsetAsync(ws.m_bits, 0, ((uint64_t)nwords + 4) * 4, st);  // (only long blocks leave marks)
      ScanArgs sc;
      sc.text = d_text;
      sc.n = n;
      sc.ds_bits = ws.ds_bits;
      sc.s_bits = ws.s_bits;
      sc.e_bits = ws.e_bits;
      sc.tile_last_hs = ws.tile_last_hs;
      sc.tile_sum = ws.tile_sum;
      sc.counters = ws.counters;
      sc.deferred = ws.deferred;
      sc.deferred_cap = ws.deferred_cap;
      sc.blocks = ws.ends;
      sc.blocks_cap = ws.blocks_cap;
      launch_scan(T, sc, st);
      g_launches.fetch_add(1);
      if (ws.h_nblk && ws.ev_nblk) {  // the host sizes the following sub-batches by the mean block length
        cudaMemcpyAsync(ws.h_nblk, ws.counters + C_N_BLK, 4, cudaMemcpyDeviceToHost, st);
        cudaEventRecord(ws.ev_nblk, st);
      }
      PROF(2);
      // the gated non-Han tokens that k_scan deferred only need the tile summaries: a one-CTA scan and a tiny kernel,
      // run on a side stream under k_route instead of leaving the GPU idle for them
      bool forked = false;
      if (!ws.aux_stream) {
        if (cudaStreamCreateWithFlags(&ws.aux_stream, cudaStreamNonBlocking) != cudaSuccess) ws.aux_stream = nullptr;
        else if (cudaEventCreateWithFlags(&ws.ev_fork, cudaEventDisableTiming) != cudaSuccess ||
                 cudaEventCreateWithFlags(&ws.ev_join, cudaEventDisableTiming) != cudaSuccess) {
          cudaStreamDestroy(ws.aux_stream);
          ws.aux_stream = nullptr;
        }
      }
      {
        cudaStream_t sx = ws.aux_stream ? ws.aux_stream : st;
        if (ws.aux_stream) {
          cudaEventRecord(ws.ev_fork, st);
          cudaStreamWaitEvent(sx, ws.ev_fork, 0);
          forked = true;
        }
        JB_LAUNCH(k_tile_scan, 1, 1024, 0, sx, ws.tile_sum, ws.tile_ctx, nt1, ws.counters, 0);
        JB_LAUNCH(k_resolve_deferred, (unsigned)g_num_sms, 256, 0, sx, ws.deferred, ws.counters, ws.deferred_cap, ws.tile_ctx, ws.s_bits,
                  ws.e_bits);
        if (forked) cudaEventRecord(ws.ev_join, sx);
      }
      RouteArgs ra;
      ra.text = d_text;
      ra.tile_last_hs = ws.tile_last_hs;
      ra.blocks = ws.ends;
      ra.blocks_cap = ws.blocks_cap;
      ra.dbg_R = ws.dbg_R;
      ra.dbg_D = ws.dbg_D;
      ra.counters = ws.counters;
      ra.path = ws.path;
      ra.wide_list = ws.wide_list;
      ra.wide_cap = ws.wide_cap;
      ra.min_chunk = 32;  // few, long blocks: full warps (measured on 10k-rune blocks: 18.4 ms/GB with 32 lanes per warp, 28.6 with 16, 28.0 with 8: the cost of an iteration does not depend on how many lanes it serves)
      launch_route(T, ra, g_num_sms, st);
      g_launches.fetch_add(1);
      {
        WideArgs wa2;
        wa2.text = d_text;
        wa2.n = n;
        wa2.ds_bits = ws.ds_bits;
        wa2.wide_list = ws.wide_list;
        wa2.wide_cap = ws.wide_cap;
        wa2.counters = ws.counters;
        wa2.R = ws.wbuf;
        wa2.len8 = ws.bp;
        wa2.code8 = reinterpret_cast<uint8_t*>(ws.rec);
        wa2.s_bits = ws.s_bits;
        wa2.e_bits = ws.e_bits;
        launch_wide(T, wa2, use_hmm, g_num_sms, st);
        g_launches.fetch_add(1);
      }
      PROF(3);
      EmitArgs ea;
      ea.text = d_text;
      ea.blocks = ws.ends;
      ea.blocks_cap = ws.blocks_cap;
      ea.counters = ws.counters;
      ea.path = ws.path;
      ea.bp = ws.bp;
      ea.s_bits = ws.s_bits;
      ea.e_bits = ws.e_bits;
      ea.m_bits = ws.m_bits;
      ea.min_chunk = 1;
      ea.segs = ws.segs;
      ea.segs_cap = ws.segs_cap;
      ea.land = ws.land;
      ea.longs = ws.longs;
      ea.longs_cap = ws.longs_cap;
      {
        const int nl = launch_emit(T, ea, use_hmm, g_num_sms, st, n, ws.ds_bits);
        if (nl > 0) g_launches.fetch_add(nl);
      }
      if (forked) cudaStreamWaitEvent(st, ws.ev_join, 0);
      PROF(4);
      if (!out.no_general) JB_LAUNCH(k_fallback_reset, (unsigned)g_num_sms * 4, 256, 0, st, ws.counters, ws.s_bits, ws.e_bits, nwords + 4);
    } else {
      PROF(2);
      PROF(3);
      PROF(4);
    }
    sa.mode = force_general ? 0 : 2;
    da.count_idx = C_N_ENDS;
    da.cursor_idx = C_CUR_DP;
    da.require_flag = force_general ? 0 : 1;
    wa.count_idx = C_N_ENDS;
    wa.cursor_idx = C_CUR_WALK;
    wa.require_flag = force_general ? 0 : 1;
    if (force_general || !out.no_general) {
      JB_LAUNCH(k_split<true>, sgrid, kSplitThreads, sizeof(SplitSmem), st, T, sa);
      JB_LAUNCH(k_tile_scan, 1, 1024, 0, st, ws.tile_sum, ws.tile_ctx, ntiles, ws.counters, force_general ? 0 : 1);
      JB_LAUNCH(k_split<false>, sgrid, kSplitThreads, sizeof(SplitSmem), st, T, sa);
      launch_dp();
      launch_walk();
    }
    PROF(5);
    JB_LAUNCH(k_rank_count, nrt, kRankWords, 0, st, ws.s_bits, ws.e_bits, nwords, ws.rank_cnt);
  } else {
    for (int i = 2; i <= 5; i++) PROF(i);
  }
  PROF(6);
  JB_LAUNCH(k_rank_scan, 1, 1024, 0, st, ws.rank_cnt, n ? nrt : 0u, ws.counters, d_n_tokens, ws.doc_off32, ndocs, n, d_doc_tok_off,
            tok_base);
  PROF(7);
  int rc = JB_OK;
  if (d_start && d_end) {
    rc = run_scatter(ws, n, ndocs, d_start, d_end, cap_tokens, d_doc_tok_off, tok_base, st);
  } else {
    if (out.bits_only && d_doc_tok_off && n > 0 && ndocs > 0)
      JB_LAUNCH(k_doc_tok, (unsigned)((ndocs * 32 + 255) / 256), 256, 0, st, ws.s_bits, ws.rank_cnt, ws.doc_off32, ndocs, n, d_doc_tok_off, tok_base);
    PROF(8);
    if (ws.prof) ws.prof_pending = true;
    rc = cudaGetLastError() == cudaSuccess ? JB_OK : JB_ECUDA;
  }
  if (wsp == &wsv) {  // event / stream handles created during this call belong to the real workspace
    wsv.s_bits = ws_in.s_bits;
    wsv.e_bits = ws_in.e_bits;
    ws_in = wsv;
    wsv = Workspace();
  }
  return rc;
}

int run_scatter(Workspace& ws, uint32_t n, uint64_t ndocs, uint32_t* d_start, uint32_t* d_end, uint64_t cap_tokens,
                uint64_t* d_doc_tok_off, uint64_t tok_base, cudaStream_t st, const uint32_t* d_s_bits, const uint32_t* d_e_bits) {
  const uint32_t nwords = (n + 31) / 32;
  const uint32_t nrt = (n + kRankBytes - 1) / kRankBytes;
  if (n > 0)
    JB_LAUNCH(k_rank_scatter, nrt, kRankWords, 0, st, d_s_bits ? d_s_bits : ws.s_bits, d_e_bits ? d_e_bits : ws.e_bits, ws.ds_bits, nwords, n, ws.rank_cnt, ws.doc_off32, ws.tile_first_doc, ndocs,
              d_start, d_end, cap_tokens, d_doc_tok_off, tok_base);
  PROF(8);
  if (ws.prof) ws.prof_pending = true;
  return cudaGetLastError() == cudaSuccess ? JB_OK : JB_ECUDA;
}

}  // namespace jb
