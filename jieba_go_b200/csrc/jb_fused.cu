// Fused tile kernel (fast path) -- see jb_fused.cuh for the contract.
// Reference functions restated here: Cut/splitText T:151-210, cutNonZh T:289-310, buildDag T:462-497,
// calcDagProba T:502-548, maxIndexProba T:565-578, findDagPath T:552-562, cutZh T:221-255,
// viterbi T:668-730, stateTransitionRoute T:736-756, cutHMM T:273-285  (T = /root/reference/tokenizer.go).
#include "jb_fused.cuh"

#include "../../include/jieba_b200.h"

namespace jb {

#define FULL 0xFFFFFFFFu

// rune-info word per slot
#define RI_CP(x) ((x) & 0xFFFFu)
#define RI_PHI(x) (((x) >> 16) & 3u)
#define RI_CLS(x) (((x) >> 18) & 3u)  // 0 none, 1 Han, 2 other 3-byte rune, 3 3-byte space
#define RI_DS 0x00100000u             // a document starts at this rune's lead byte

enum : uint32_t { PCL_HAN = 1, PCL_ALNUM = 2, PCL_SPACE = 3, PCL_OTHER = 4, PCL_INVALID = 5 };
#define PCLS(c, len) (uint8_t)(((c) << 3) | (len))

struct FusedSmem {
  uint8_t sb[kFtRegion];
  uint32_t ri[kFtSlots + 8];  // index = local slot + 1 (slot -1 lives at 0)
  uint32_t dsw[kFtRegion / 32 + 3];
  uint32_t BND[kFtWords + 1], ALN[kFtWords + 1];
  uint32_t S[kFtWords + 2], E[kFtWords + 2];
  uint32_t GS[kFtTileWords + 1], GE[kFtTileWords + 2];
  uint32_t CW[4];
  uint32_t HS[kFtSlots / 32 + 1], HE[kFtSlots / 32 + 1];
  uint32_t cmask[kFtSlots];
  // per slot: [0] = wbuf index of its first candidates (stored back to back) | their number << 12 (0xFFFF = not packed);
  // [1..3] = wbuf indexes of candidates found later by the work-list warp.  Ascending length throughout.
  uint16_t xi[kFtSlots][4];
  uint2 wl[kFtWorkList];              // prefix chains still alive after the two straight-line probes
  uint16_t blk[kFtMaxBlocks];      // owned blocks: first slot
  int16_t blk_end[kFtMaxBlocks];   //               last slot, or -1 when the block does not end inside the region
  double wbuf[kFtWCap];
  uint8_t wcls[kFtThreads / 32][40];
  uint16_t pofs[kFtSlots + 2];  // exclusive prefix of stream record sizes
  uint32_t wsum[kFtThreads / 32];
  uint32_t nblk, wcnt, overflow, npacked, npacked2, sbase, bbase, wl_n, ovf_total, ovf_cur, n_o3;
  uint16_t o3[kFtTileBytes / 3 + 8];  // slots of the tile that hold a 3-byte non-Han, non-space rune (gated tokens)
  int first_ks, act_limit;
};

__device__ __forceinline__ bool f_is_alnum(uint32_t c) { return (c - '0' < 10u) || ((c | 0x20) - 'a' < 26u); }
__device__ __forceinline__ bool f_is_space(uint32_t cp) {
  if (cp <= 0xFF) return (cp - 9u < 5u) || cp == 0x20 || cp == 0x85 || cp == 0xA0;
  return cp == 0x1680 || (cp - 0x2000u <= 0xAu) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}
__device__ __forceinline__ bool f_is_han(uint32_t cp, const JbTables& T) {
  if (cp < 0x10000) return (__ldg(T.han_bits + (cp >> 5)) >> (cp & 31)) & 1;
  for (uint32_t i = 0; i < T.n_supp; i++)
    if (cp >= T.supp_lo[i] && cp <= T.supp_hi[i]) return true;
  return false;
}

struct FCtx {
  const FusedSmem* s;
  uint32_t t0, n;
  // is region index i a document start, or at/after the end of the text?
  __device__ __forceinline__ bool ds_at(int i) const {
    int64_t P = (int64_t)t0 - kFtLeft + i;
    if (P >= (int64_t)n) return true;
    if (P < 0) return false;
    return (s->dsw[(i + 16) >> 5] >> ((i + 16) & 31)) & 1;
  }
};

// validated length (2..4) of the UTF-8 sequence whose lead byte is at region index i, 0 if ill-formed
__device__ __forceinline__ int f_seqlen(const FusedSmem& S, const FCtx& cx, int i) {
  uint32_t b = S.sb[i];
  int len = 0;
  if (b >= 0xC2 && b <= 0xDF) len = 2;
  else if (b >= 0xE0 && b <= 0xEF) len = 3;
  else if (b >= 0xF0 && b <= 0xF4) len = 4;
  if (!len || i + len > kFtRegion) return 0;
  uint32_t b1 = S.sb[i + 1], lo = 0x80, hi = 0xBF;
  if (b == 0xE0) lo = 0xA0;
  if (b == 0xED) hi = 0x9F;
  if (b == 0xF0) lo = 0x90;
  if (b == 0xF4) hi = 0x8F;
  if (b1 < lo || b1 > hi || cx.ds_at(i + 1)) return 0;
  if (len >= 3 && ((S.sb[i + 2] & 0xC0) != 0x80 || cx.ds_at(i + 2))) return 0;
  if (len == 4 && ((S.sb[i + 3] & 0xC0) != 0x80 || cx.ds_at(i + 3))) return 0;
  return len;
}
__device__ __forceinline__ uint32_t f_decode(const uint8_t* b, int len) {
  uint32_t b0 = b[0];
  if (len == 1) return b0;
  if (len == 2) return ((b0 & 0x1F) << 6) | (b[1] & 0x3F);
  if (len == 3) return ((b0 & 0x0F) << 12) | ((b[1] & 0x3Fu) << 6) | (b[2] & 0x3F);
  return ((b0 & 0x07) << 18) | ((b[1] & 0x3Fu) << 12) | ((b[2] & 0x3Fu) << 6) | (b[3] & 0x3F);
}
// exact per-byte class under Go's decoding rules (same rules as classify_tile in jb_kernels.cu)
__device__ uint8_t f_pb_cls(const FusedSmem& S, const FCtx& cx, const JbTables& T, int i) {
  uint32_t b = S.sb[i];
  if (b < 0x80) return f_is_alnum(b) ? PCLS(PCL_ALNUM, 1) : (f_is_space(b) ? PCLS(PCL_SPACE, 1) : PCLS(PCL_OTHER, 1));
  if ((b & 0xC0) == 0x80) {
    for (int k = 1; k <= 3 && i - k >= 0; k++) {
      if ((S.sb[i - k] & 0xC0) != 0x80) return f_seqlen(S, cx, i - k) > k ? 0 : PCLS(PCL_INVALID, 1);
    }
    return PCLS(PCL_INVALID, 1);
  }
  int len = f_seqlen(S, cx, i);
  if (!len) return PCLS(PCL_INVALID, 1);
  uint32_t cp = f_decode(&S.sb[i], len);
  return f_is_han(cp, T) ? PCLS(PCL_HAN, len) : (f_is_space(cp) ? PCLS(PCL_SPACE, len) : PCLS(PCL_OTHER, len));
}

// Does the non-Han block around tile-local byte q hold an ASCII alnum (cutNonZh T:290-293)?
// returns 1 yes, 0 no, else (2 | need_fwd<<2 | need_bwd<<3) when the block leaves the staged region.
__device__ uint32_t f_block_alnum(const uint32_t* BND, const uint32_t* ALN, int q) {
  int w = q >> 5, b = q & 31;
  uint32_t lowmask = (b == 31) ? 0xFFFFFFFFu : ((2u << b) - 1u);
  bool found = false;
  uint32_t m = BND[w] & lowmask;
  if (m) {
    int bb = 31 - __clz(m);
    if (ALN[w] & lowmask & ~((1u << bb) - 1u)) return 1;
    found = true;
  } else {
    if (ALN[w] & lowmask) return 1;
    for (int ww = w - 1; ww >= 0; --ww) {
      m = BND[ww];
      if (m) {
        int bb = 31 - __clz(m);
        if (ALN[ww] & ~((1u << bb) - 1u)) return 1;
        found = true;
        break;
      } else if (ALN[ww])
        return 1;
    }
  }
  uint32_t need = found ? 0u : 4u;
  uint32_t highmask = ~lowmask;
  found = false;
  m = BND[w] & highmask;
  if (m) {
    int bb = __ffs(m) - 1;
    if (ALN[w] & highmask & ((1u << bb) - 1u)) return 1;
    found = true;
  } else {
    if (ALN[w] & highmask) return 1;
    for (int ww = w + 1; ww < kFtWords; ++ww) {
      m = BND[ww];
      if (m) {
        int bb = __ffs(m) - 1;
        if (ALN[ww] & ((1u << bb) - 1u)) return 1;
        found = true;
        break;
      } else if (ALN[ww])
        return 1;
    }
  }
  if (!found) need |= 8u;
  return need ? (2u | need) : 0u;
}

__device__ __forceinline__ void f_set(uint32_t* bits, int q) { atomicOr(&bits[q >> 5], 1u << (q & 31)); }

// first set bit of HE at index >= kk, or -1
__device__ __forceinline__ int f_find_end(const uint32_t* HE, int kk) {
  int w = kk >> 5;
  uint32_t m = HE[w] & ~((1u << (kk & 31)) - 1u);
  for (;;) {
    if (m) return w * 32 + __ffs(m) - 1;
    if (++w >= kFtSlots / 32) return -1;
    m = HE[w];
  }
}
// last set bit of HS at index <= kk, or -1
__device__ __forceinline__ int f_find_start(const uint32_t* HS, int kk) {
  int w = kk >> 5, b = kk & 31;
  uint32_t m = HS[w] & ((b == 31) ? 0xFFFFFFFFu : ((2u << b) - 1u));
  for (;;) {
    if (m) return w * 32 + 31 - __clz(m);
    if (--w < 0) return -1;
    m = HS[w];
  }
}

// exact per-byte rules for the 32-byte word g of the tile (one warp, one byte per lane)
template <int NT>
__device__ __forceinline__ void f_pb_word(FusedSmem& S, const FCtx& cx, const JbTables& T, const FusedArgs& A, uint32_t t0, uint32_t n,
                                          int g, int lane, int warp) {
    const int q = 32 * g + lane, i = q + kFtLeft;
    const uint8_t c = f_pb_cls(S, cx, T, i);
    S.wcls[warp][4 + lane] = c;
    if (lane < 4) S.wcls[warp][lane] = f_pb_cls(S, cx, T, i - 4);
    __syncwarp();
    const uint32_t P = t0 + q;
    const uint32_t cl = c >> 3, len = c & 7;
    const bool start = c != 0 && P < n;
    const bool han = cl == PCL_HAN;
    uint32_t pc = PCL_OTHER;
    for (int k = 1; k <= 4; k++) {
      uint8_t c2 = S.wcls[warp][4 + lane - k];
      if (c2) {
        pc = c2 >> 3;
        break;
      }
    }
    __syncwarp();
    if (start && han && len == 4) atomicOr(&A.counters[C_FLAGS], 1u);  // 4-byte Han: general pipeline redoes the batch
    const bool dsq = start && (P == 0 || cx.ds_at(i));
    const bool bnd = start && (dsq || (han != (pc == PCL_HAN)));
    const bool aln = start && cl == PCL_ALNUM;
    bool nsa = false, nea = false, gs = false;
    if (start && !han && q < kFtTileBytes) {
      if (cl == PCL_ALNUM) {  // alnum run = one token (T:298-299)
        nsa = dsq || !f_is_alnum(S.sb[i - 1]);
        nea = P + 1 >= n || cx.ds_at(i + 1) || !f_is_alnum(S.sb[i + 1]);
      } else if (cl != PCL_SPACE) {  // any other rune is its own token, spaces are skipped (T:301-306)
        gs = true;
        f_set(S.GE, q + (int)len - 1);
      }
    }
    uint32_t wb = __ballot_sync(FULL, bnd), wa = __ballot_sync(FULL, aln), w1 = __ballot_sync(FULL, nsa), w2 = __ballot_sync(FULL, nea),
             w3 = __ballot_sync(FULL, gs);
    if (lane == 0) {
      if (wb) atomicOr(&S.BND[g], wb);
      if (wa) atomicOr(&S.ALN[g], wa);
      if (g < kFtTileWords) {
        if (w1) atomicOr(&S.S[g], w1);
        if (w2) atomicOr(&S.E[g], w2);
        if (w3) atomicOr(&S.GS[g], w3);
      }
    }
}

template <bool HMM>
__global__ void __launch_bounds__(kFtThreads) k_fused(const JbTables T, const FusedArgs A) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  FusedSmem& S = *reinterpret_cast<FusedSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t t0 = tile * (uint32_t)kFtTileBytes;
  const uint32_t n = A.n;
  FCtx cx{&S, t0, n};

  // ---- stage bytes + document-start words, clear bit words --------------------------------
  {
    const bool aligned = ((reinterpret_cast<uintptr_t>(A.text) & 15) == 0);
    for (int c = tid; c < kFtRegion / 16; c += kFtThreads) {
      int64_t P = (int64_t)t0 - kFtLeft + c * 16;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (aligned && P >= 0 && P + 16 <= (int64_t)n) {
        v = __ldcs(reinterpret_cast<const uint4*>(A.text + P));  // read once: evict first
      } else if (P + 16 > 0 && P < (int64_t)n) {
        uint8_t* vb = reinterpret_cast<uint8_t*>(&v);
        for (int j = 0; j < 16; j++) {
          int64_t q = P + j;
          vb[j] = (q >= 0 && q < (int64_t)n) ? __ldg(A.text + q) : 0;
        }
      }
      *reinterpret_cast<uint4*>(&S.sb[c * 16]) = v;
    }
    const uint32_t nwords = (n + 31) / 32;
    for (int j = tid; j < kFtRegion / 32 + 3; j += kFtThreads) {
      int64_t gw = (int64_t)(t0 / 32) - 1 + j;
      S.dsw[j] = (gw >= 0 && gw < (int64_t)nwords) ? __ldg(A.ds_bits + gw) : 0;
    }
    for (int j = tid; j < kFtWords + 2; j += kFtThreads) {
      S.S[j] = 0;
      S.E[j] = 0;
      if (j <= kFtWords) {
        S.BND[j] = 0;
        S.ALN[j] = 0;
      }
      if (j <= kFtTileWords + 1) S.GE[j] = 0;
      if (j <= kFtTileWords) S.GS[j] = 0;
      if (j < 4) S.CW[j] = 0;
    }
    if (tid == 0) {
      S.nblk = 0;
      S.wcnt = 0;
      S.wl_n = 0;
      S.overflow = 0;
      S.npacked = 0;
      S.ovf_total = 0;
      S.n_o3 = 0;
      S.first_ks = kFtSlots;
      S.act_limit = -1;
    }
  }
  __syncthreads();

  // ---- A: slot-centric decode of 3-byte runes ---------------------------------------------
  // slot kk (kk = -1 .. kFtSlots-1) owns the rune whose lead byte is at tile-local 3kk-2 .. 3kk
  for (int s = tid; s <= kFtSlots; s += kFtThreads) {
    const int i0 = kFtLeft + 3 * (s - 1) - 2;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(S.sb + (i0 & ~3));
    const uint32_t lo = wp[0], hi = wp[1], sh = (i0 & 3) * 8;
    const uint32_t x0 = __funnelshift_r(lo, hi, sh), x1 = hi >> sh;
    uint32_t info = 0;
    uint32_t mm = (x0 & 0x00F0F0F0u) ^ 0x00E0E0E0u;
    uint32_t z = (mm - 0x00010101u) & ~mm & 0x00808080u;  // bytes equal to 0xE? (exact: nibble values)
    while (z) {
      const uint32_t j = (__ffs(z) - 1) >> 3;
      z &= z - 1;
      const uint32_t y = __funnelshift_r(x0, x1, 8 * j);
      const uint32_t L = y & 0xFF, c1 = (y >> 8) & 0xFF, c2 = (y >> 16) & 0xFF;
      bool valid = ((y & 0x00C0C000u) == 0x00808000u) && !(L == 0xE0 && c1 < 0xA0) && !(L == 0xED && c1 > 0x9F);
      if (valid) {
        const uint32_t cp = ((L & 0xF) << 12) | ((c1 & 0x3F) << 6) | (c2 & 0x3F);
        const uint32_t cl = f_is_han(cp, T) ? 1u : (f_is_space(cp) ? 3u : 2u);
        info = cp | (j << 16) | (cl << 18);
        break;
      }
    }
    S.ri[s] = info;
  }
  __syncthreads();
  // document starts: flag the rune that starts there, kill a "rune" that a boundary cuts through
  for (int j = tid; j < kFtRegion / 32 + 3; j += kFtThreads) {
    uint32_t m = S.dsw[j];
    while (m) {
      int b = __ffs(m) - 1;
      m &= m - 1;
      int q = 32 * j + b - 32;  // tile-local byte of the document start
      if (q < -4 || q >= kFtTileBytes + kFtHaloBytes) continue;
      int kk = (q + 5) / 3 - 1;  // floor((q+2)/3) for q >= -5
      for (int s = kk + 1; s >= kk && s >= 0; --s) {  // s = slot index + 1 of slots kk and kk-1
        if (s > kFtSlots) continue;
        uint32_t info = S.ri[s];
        if (!RI_CLS(info)) continue;
        int lead = 3 * (s - 1) - 2 + (int)RI_PHI(info);
        if (lead == q) atomicOr(&S.ri[s], RI_DS);
        else if (lead < q && q <= lead + 2) {
          S.ri[s] = 0;
          if (q >= 0) atomicOr(&S.CW[(q >> 5) >> 5], 1u << ((q >> 5) & 31));
        }
      }
    }
  }
  if (tid == 0 && n >= t0 && n - t0 < (uint32_t)(kFtTileBytes + kFtHaloBytes)) f_set(S.BND, (int)(n - t0));  // end of text
  __syncthreads();

  // ---- B: seams, block boundaries ------------------------------------------------------------
  for (int base = 0; base < kFtSlots; base += kFtThreads) {
    const int kk = base + tid;
    const uint32_t cur = S.ri[kk + 1], prev = S.ri[kk], next = (kk + 1 < kFtSlots) ? S.ri[kk + 2] : 0u;
    const uint32_t ccl = RI_CLS(cur), pcl = RI_CLS(prev), ncl = RI_CLS(next);
    const bool seam_prev = !(ccl && pcl && RI_PHI(cur) == RI_PHI(prev));
    const bool seam_next = !(ccl && ncl && RI_PHI(cur) == RI_PHI(next));
    const int q = 3 * kk - 2 + (int)RI_PHI(cur);
    if (seam_prev) {  // something other than back-to-back 3-byte runes: exact per-byte rules for these words
      int qa = 3 * kk - 5, qb = 3 * kk;
      if (qa < 0) qa = 0;
      if (qb >= 0 && qa < kFtWords * 32) {
        if (qb >= kFtWords * 32) qb = kFtWords * 32 - 1;
        atomicOr(&S.CW[(qa >> 5) >> 5], 1u << ((qa >> 5) & 31));
        atomicOr(&S.CW[(qb >> 5) >> 5], 1u << ((qb >> 5) & 31));
      }
    }
    const bool han = ccl == 1, ds = cur & RI_DS;
    const bool bstart = han && (ds || seam_prev || pcl != 1);
    bool bend = han && (seam_next || ncl != 1 || (next & RI_DS));
    if (kk == kFtSlots - 1) bend = false;  // the region ends here: an open block is "long"
    const bool bnd = ccl && (ds || (han && seam_prev) || (!seam_prev && ((pcl == 1) != han)));
    if (bnd && q >= 0 && q < kFtWords * 32) f_set(S.BND, q);
    uint32_t hs = __ballot_sync(FULL, bstart), he = __ballot_sync(FULL, bend);
    const bool o3 = ccl == 2 && q >= 0 && q < kFtTileBytes;
    const uint32_t om = __ballot_sync(FULL, o3);
    uint32_t ob = 0;
    if (lane == 0) {
      S.HS[kk >> 5] = hs;
      S.HE[kk >> 5] = he;
      if (om) ob = atomicAdd(&S.n_o3, (uint32_t)__popc(om));
    }
    ob = __shfl_sync(FULL, ob, 0);
    if (o3) S.o3[ob + __popc(om & ((1u << lane) - 1u))] = (uint16_t)kk;
  }
  __syncthreads();

  // ---- C: exact per-byte rules on the words that need them (ASCII, 2/4-byte runes, ill-formed) ----
  // complex words only: the n-th set bit of CW goes to warp n mod (number of warps)
  for (int cw = 0, seen = 0; cw < 4; cw++) {
    uint32_t bits = S.CW[cw];
    while (bits) {
      const int g = cw * 32 + __ffs(bits) - 1;
      bits &= bits - 1;
      if ((seen++ % (kFtThreads / 32)) != warp || g >= kFtWords) continue;
      f_pb_word<kFtThreads>(S, cx, T, A, t0, n, g, lane, warp);
    }
  }
  __syncthreads();
  // ---- D: gated single-rune tokens (cutNonZh drops a block without [a-zA-Z0-9], T:290-293) ----
  // (1) 3-byte non-Han runes in clean words, straight from the slot table
  for (uint32_t oi = tid; oi < S.n_o3; oi += kFtThreads) {
    const int kk = S.o3[oi];
    const uint32_t cur = S.ri[kk + 1];
    const int q = 3 * kk - 2 + (int)RI_PHI(cur);
    if ((S.CW[(q >> 5) >> 5] >> ((q >> 5) & 31)) & 1) continue;  // that word went through C
    uint32_t r = f_block_alnum(S.BND, S.ALN, q);
    if (r == 1) {
      f_set(S.S, q);
      f_set(S.E, q + 2);
    } else if (r & 2) {
      uint32_t idx = atomicAdd(&A.counters[C_N_DEFER], 1u);
      if (idx < A.deferred_cap) A.deferred[idx] = make_uint4(t0 + q, 3u, (r >> 2) & 3u, tile);
      else atomicOr(&A.counters[C_FLAGS], 1u);
    }
  }
  // (2) tokens found by the per-byte pass
  for (int g = warp; g < kFtTileWords; g += kFtThreads / 32) {
    const uint32_t gsw = S.GS[g];
    if (!gsw) continue;
    if ((gsw >> lane) & 1) {
      const int q = 32 * g + lane;
      // its end bit: first GE bit at or after q
      int w = q >> 5;
      uint32_t m = S.GE[w] & ~((1u << (q & 31)) - 1u);
      if (!m) m = S.GE[++w];
      const int qe = w * 32 + __ffs(m) - 1;
      uint32_t r = f_block_alnum(S.BND, S.ALN, q);
      if (r == 1) {
        f_set(S.S, q);
        f_set(S.E, qe);
      } else if (r & 2) {
        uint32_t idx = atomicAdd(&A.counters[C_N_DEFER], 1u);
        if (idx < A.deferred_cap) A.deferred[idx] = make_uint4(t0 + q, (uint32_t)(qe - q + 1), (r >> 2) & 3u, tile);
        else atomicOr(&A.counters[C_FLAGS], 1u);
      }
    }
  }
  // tile summary for k_tile_scan (same encoding as k_split<true>)
  if (warp == 0) {
    int firstw = kFtTileWords, lastw = -1;
    for (int j = lane; j < kFtTileWords; j += 32)
      if (S.BND[j]) {
        firstw = min(firstw, j);
        lastw = max(lastw, j);
      }
    firstw = __reduce_min_sync(FULL, firstw);
    lastw = __reduce_max_sync(FULL, lastw);
    bool pre = false, post = false;
    if (lastw < 0) {
      for (int j = lane; j < kFtTileWords; j += 32) pre |= S.ALN[j] != 0;
      post = pre;
    } else {
      uint32_t fb = __ffs(S.BND[firstw]) - 1, lb = 31 - __clz(S.BND[lastw]);
      for (int j = lane; j < kFtTileWords; j += 32) {
        uint32_t a = S.ALN[j];
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

  // ---- E: blocks owned by this tile (start rune's lead byte inside the tile) -------------------
  for (int base = 0; base < kFtSlots; base += kFtThreads) {
    const int kk = base + tid;
    const uint32_t cur = S.ri[kk + 1];
    const int q = 3 * kk - 2 + (int)RI_PHI(cur);
    const bool owned = ((S.HS[kk >> 5] >> (kk & 31)) & 1) && q >= 0 && q < kFtTileBytes;
    uint32_t bm = __ballot_sync(FULL, owned);
    uint32_t wbase = 0;
    if (lane == 0 && bm) wbase = atomicAdd(&S.nblk, (uint32_t)__popc(bm));
    wbase = __shfl_sync(FULL, wbase, 0);
    if (owned) {
      const int ke = f_find_end(S.HE, kk);
      const uint32_t bi = wbase + __popc(bm & ((1u << lane) - 1u));
      if (bi < (uint32_t)kFtMaxBlocks) {
        S.blk[bi] = (uint16_t)kk;
        S.blk_end[bi] = (int16_t)ke;
        atomicMin(&S.first_ks, kk);
        if (ke >= 0) atomicMax(&S.act_limit, ke);
      } else {
        atomicOr(&A.counters[C_FLAGS], 1u);  // absurdly fragmented tile: the general pipeline redoes the batch
      }
    }
  }
  __syncthreads();

  // ---- F: DAG probe (buildDag T:462-497).  Per Han slot of an owned, closed block: the first-rune
  // table answers termFreq[string(iRune)] (T:468-472); then up to two more runes are probed in
  // straight-line code (88 % of all prefix chains end within them); a chain that is still alive goes
  // on a short work list that one warp finishes.  This is the reference's
  // `for j := range textRunes[i:]` with its break on the first missing prefix (T:473-482). -------
  // work-list entry: x = slot | L<<11 | maxlen<<27, y = parent id
  {
    const int first_ks = S.first_ks, act_limit = S.act_limit;
    for (int base = 0; base < kFtSlots; base += kFtThreads) {
      const int kk = base + tid;
      const uint32_t cur = S.ri[kk + 1];
      const bool act = kk >= first_ks && kk <= act_limit && RI_CLS(cur) == 1;
      uint32_t mask = 0, cnt = 0, L = 1, parent = 0, maxlen = 0;
      double w0 = 0.0, w1 = 0.0, w2 = 0.0;
      bool alive = false;
      if (act) {
        const uint32_t r0 = RI_CP(cur);
        const uint4 f = __ldg(reinterpret_cast<const uint4*>(T.first + r0));
        w0 = __longlong_as_double(((long long)f.y << 32) | (long long)f.x);
        mask = 1;
        cnt = 1;
        maxlen = min((f.z >> 8) & 0xFFu, 31u);
        if (!(f.z & JB_FIRST_GATE) && maxlen > 1 && !((S.HE[kk >> 5] >> (kk & 31)) & 1)) {
          const uint32_t r1 = RI_CP(S.ri[kk + 2]);
          alive = (f.w >> jb_bloom_bit(r1)) & 1;
          parent = JB_PARENT_FIRST(r0);
        }
      }
#pragma unroll
      for (int step = 0; step < 2; step++) {
        if (alive) {
          const uint32_t rl = RI_CP(S.ri[kk + 1 + L]);
          double pw;
          uint32_t prb;
          const int ps = jb_probe_edge(T.entries, T.hash_mask, parent, rl, &pw, &prb);
          alive = false;
          if (ps >= 0) {  // !found -> break (T:476-478)
            L++;
            if (jb_w_positive(pw)) {  // val > 0 -> edge (T:479-481)
              mask |= 1u << (L - 1);
              if (cnt == 1) w1 = pw;
              else w2 = pw;
              cnt++;
            }
            const int last = kk + (int)L - 1;
            if (L < maxlen && !((S.HE[last >> 5] >> (last & 31)) & 1)) {
              parent = (uint32_t)ps;
              alive = ((prb >> 21) >> jb_bloom11(RI_CP(S.ri[kk + 1 + L]))) & 1;
            }
          }
        }
      }
      // weight slots for what was found so far; chains still alive go on the work list
      uint32_t incl = cnt;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      const uint32_t total = __shfl_sync(FULL, incl, 31);
      const uint32_t lm = __ballot_sync(FULL, alive);
      uint32_t wb = 0, lb = 0;
      if (lane == 0) {
        if (total) wb = atomicAdd(&S.wcnt, total);
        if (lm) lb = atomicAdd(&S.wl_n, (uint32_t)__popc(lm));
      }
      wb = __shfl_sync(FULL, wb, 0);
      lb = __shfl_sync(FULL, lb, 0);
      if (act) {
        const uint32_t wi = wb + incl - cnt;
        S.cmask[kk] = mask;
        if (wi + cnt <= (uint32_t)kFtWCap) {
          S.xi[kk][0] = (uint16_t)(wi | (cnt << 12));
          S.wbuf[wi] = w0;
          if (cnt > 1) S.wbuf[wi + 1] = w1;
          if (cnt > 2) S.wbuf[wi + 2] = w2;
        } else {
          S.xi[kk][0] = 0;
          S.overflow = 1;
        }
      }
      if (alive) {
        const uint32_t li = lb + __popc(lm & ((1u << lane) - 1u));
        if (li < (uint32_t)kFtWorkList) S.wl[li] = make_uint2((uint32_t)kk | (L << 11) | (maxlen << 27), parent);
        else S.overflow = 1;
      }
    }
    __syncthreads();
    // leftovers (about one chain in ten): one thread each, spread over all warps (they are chains of
    // dependent L2 probes: many warps with a lane or two each hide that latency, one warp would not)
    {
      const uint32_t nl = min(S.wl_n, (uint32_t)kFtWorkList);
      const uint32_t nwarps = kFtThreads / 32;
      for (uint32_t i = (uint32_t)lane * nwarps + warp; i < nl; i += kFtThreads) {
        const uint2 e = S.wl[i];
        const int kk = (int)(e.x & 0x7FFu);
        uint32_t L = (e.x >> 11) & 31u, parent = e.y;
        const uint32_t maxlen = e.x >> 27;
        for (;;) {
          const uint32_t rl = RI_CP(S.ri[kk + 1 + L]);
          double pw;
          uint32_t prb;
          const int ps = jb_probe_edge(T.entries, T.hash_mask, parent, rl, &pw, &prb);
          if (ps < 0) break;
          L++;
          if (jb_w_positive(pw)) {
            const uint32_t m = S.cmask[kk];
            const uint32_t x = __popc(m) - (S.xi[kk][0] >> 12);  // ordinal among the late candidates
            const uint32_t wi = atomicAdd(&S.wcnt, 1u);
            S.cmask[kk] = m | (1u << (L - 1));
            if (wi < (uint32_t)kFtWCap && x < 3u) {
              S.wbuf[wi] = pw;
              S.xi[kk][1 + x] = (uint16_t)wi;
            } else {
              S.overflow = 1;
            }
          }
          const int last = kk + (int)L - 1;
          if (L >= maxlen || ((S.HE[last >> 5] >> (last & 31)) & 1)) break;
          if (!(((prb >> 21) >> jb_bloom11(RI_CP(S.ri[kk + 1 + L]))) & 1)) break;
          parent = (uint32_t)ps;
        }
      }
    }
  }
  __syncthreads();

  // ---- G: hand the owned blocks to k_block_dp: pack every block's candidates into one stream -------
  // Stream unit = 8 bytes; one 32-byte RECORD (= one L2 sector) per position: header (bits 0..31
  // candidate-length mask, 32..47 the rune, 48..63 index of its overflow weights) + the first three
  // float64 weights in ascending length; the rare 4th+ weights go to an overflow area behind the tile's
  // records.  Records are fixed-size, so k_block_dp can prefetch them ahead of the dependent DP chain.
  // The tile's positions are stored in REVERSE slot order: a block's records are contiguous and start
  // with its last rune -- the order the right-to-left route DP consumes them.
  const uint32_t nblk = min(S.nblk, (uint32_t)kFtMaxBlocks);
  const bool tile_overflow = S.overflow != 0;
  for (uint32_t b = tid; b < nblk; b += kFtThreads) {
    const int ks = S.blk[b];
    const int ke = S.blk_end[b];
    if (ke < 0 || tile_overflow || ke - ks + 1 > kFtMaxBlockLen) {
      // long block, or this tile's weights did not fit in shared memory: the general kernels take it
      uint32_t idx = atomicAdd(&A.counters[C_N_LONG], 1u);
      if (idx < A.long_cap) A.long_seeds[idx] = t0 + 3 * ks - 2 + RI_PHI(S.ri[ks + 1]);
      else atomicOr(&A.counters[C_FLAGS], 1u);
      S.blk_end[b] = -1;
      if (ke >= 0 && !tile_overflow)
        for (int k = ks; k <= ke; k++) S.xi[k][0] = 0xFFFFu;  // not packed
    } else {
      atomicAdd(&S.npacked, 1u);
    }
  }
  __syncthreads();
  {
    // exclusive scan of "packed position" flags over the slots (consecutive slots per thread)
    const int first_ks = S.first_ks, act_limit = tile_overflow ? -1 : S.act_limit;
    uint32_t sz[kFtSlots / kFtThreads], sum = 0, ovs = 0;
#pragma unroll
    for (int j = 0; j < kFtSlots / kFtThreads; j++) {
      const int kk = tid * (kFtSlots / kFtThreads) + j;
      uint32_t v = 0;
      if (kk >= first_ks && kk <= act_limit && RI_CLS(S.ri[kk + 1]) == 1 && S.xi[kk][0] != 0xFFFFu) {
        v = (uint32_t)__popc(S.cmask[kk]);  // candidates of this position (>= 1)
        if (v > 3) ovs += v - 3;
      }
      sz[j] = v;
      sum += v ? 1u : 0u;
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) S.wsum[warp] = incl;
    ovs = __reduce_add_sync(FULL, ovs);
    if (lane == 0 && ovs) atomicAdd(&S.ovf_total, ovs);
    __syncthreads();
    uint32_t pre = incl - sum;
    {  // exclusive prefix of the warps' sums: one shuffle scan per warp
      uint32_t ws = lane < kFtThreads / 32 ? S.wsum[lane] : 0u, wi = ws;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t v = __shfl_up_sync(FULL, wi, o);
        if (lane >= o) wi += v;
      }
      pre += __shfl_sync(FULL, wi - ws, warp);
    }
#pragma unroll
    for (int j = 0; j < kFtSlots / kFtThreads; j++) {
      S.pofs[tid * (kFtSlots / kFtThreads) + j] = (uint16_t)pre;
      pre += sz[j] ? 1u : 0u;
    }
    if (tid == kFtThreads - 1) {
      S.pofs[kFtSlots] = (uint16_t)pre;
      // reserve stream space (4 units per position + the overflow weights, rounded to whole records so that
      // every record stays 32-byte aligned) and block descriptors for this tile
      const uint32_t total = (4u * pre + S.ovf_total + 3u) & ~3u, np = S.npacked;
      uint32_t sbase = 0, bbase = 0;
      if (np) {
        sbase = atomicAdd(&A.counters[C_STREAM], total);
        bbase = atomicAdd(&A.counters[C_N_FBLK], np);
        if (sbase + total > A.stream_cap || bbase + np > A.fblk_cap) {
          atomicOr(&A.counters[C_FLAGS], 1u);  // out of stream space: the general pipeline redoes the batch
          S.npacked = 0;
        }
      }
      S.sbase = sbase;
      S.bbase = bbase;
      S.npacked2 = 0;
      S.ovf_cur = 0;
    }
    __syncthreads();
    if (S.npacked) {
      const uint32_t npos_t = S.pofs[kFtSlots], sbase = S.sbase;
      unsigned long long* __restrict__ st = A.stream + sbase;
#pragma unroll
      for (int j = 0; j < kFtSlots / kFtThreads; j++) {
        const int kk = tid * (kFtSlots / kFtThreads) + j;
        const uint32_t cnt = sz[j];
        if (!cnt) continue;
        const uint32_t r = npos_t - S.pofs[kk + 1];  // record index: the tile's positions in reverse slot order
        uint32_t ovi = 0;
        if (cnt > 3) ovi = atomicAdd(&S.ovf_cur, cnt - 3);
        const uint32_t x0 = S.xi[kk][0], wi0 = x0 & 0xFFFu, n0 = x0 >> 12;
        unsigned long long u[4];
        u[0] = (unsigned long long)S.cmask[kk] | ((unsigned long long)RI_CP(S.ri[kk + 1]) << 32) | ((unsigned long long)(ovi & 0xFFFFu) << 48);
#pragma unroll
        for (uint32_t c = 0; c < 3; c++)
          u[1 + c] = c < cnt ? (unsigned long long)__double_as_longlong(S.wbuf[c < n0 ? wi0 + c : S.xi[kk][1 + c - n0]]) : 0ull;
        // one-touch data: streaming stores, so that the record stream does not evict the hash table from L2
        ulonglong2* rp = reinterpret_cast<ulonglong2*>(st + 4u * r);
        __stcs(rp, make_ulonglong2(u[0], u[1]));
        __stcs(rp + 1, make_ulonglong2(u[2], u[3]));
        for (uint32_t c = 3; c < cnt; c++)
          __stcs(st + 4u * npos_t + ovi + (c - 3), (unsigned long long)__double_as_longlong(S.wbuf[c < n0 ? wi0 + c : S.xi[kk][1 + c - n0]]));
      }
      for (uint32_t b = tid; b < nblk; b += kFtThreads) {
        const int ke = S.blk_end[b];
        if (ke < 0) continue;
        const int ks = S.blk[b];
        const uint32_t bi = S.bbase + atomicAdd(&S.npacked2, 1u);
        A.fblocks[bi] = make_uint4(sbase + 4u * (npos_t - S.pofs[ke + 1]), t0 + 3 * ks - 2 + RI_PHI(S.ri[ks + 1]), (uint32_t)(ke - ks + 1),
                                   sbase + 4u * npos_t);
      }
    }
  }
  __syncthreads();
  // ---- H: publish token bits ----------------------------------------------------------------------
  const uint32_t w0 = t0 / 32;
  for (int j = tid; j < kFtWords + 1; j += kFtThreads) {
    const uint32_t sbits = S.S[j], ebits = S.E[j];
    if (sbits) atomicOr(&A.s_bits[w0 + j], sbits);
    if (ebits) atomicOr(&A.e_bits[w0 + j], ebits);
  }
}

// ------------------------------------------------------------------------------------------
// k_block_dp: one lane per packed Han block, lanes refilled from a warp-level queue (every lane of
// every warp busy, whatever the block lengths).  Per block: route DP right to left over the packed
// stream (calcDagProba T:502-548 + maxIndexProba T:565-578), then the forward walk (findDagPath
// T:552-562), the HMM glue (cutZh T:221-255), Viterbi (T:668-756) and cutHMM (T:273-285).
// Shared memory per lane: a ring of RING route values and one chosen length per rune.
// ------------------------------------------------------------------------------------------
constexpr int kBdThreads = 128;
constexpr int kBdQueue = 32;

struct BitAcc {  // token bits of one lane, flushed one 32-byte word at a time (positions only grow)
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

template <bool HMM, int RING>
__global__ void __launch_bounds__(kBdThreads) k_block_dp(const JbTables T, const BlockDpArgs A) {
  extern __shared__ __align__(16) uint8_t bd_smem[];
  double* ring = reinterpret_cast<double*>(bd_smem);                           // [RING][kBdThreads]
  uint8_t* path = bd_smem + (size_t)RING * kBdThreads * sizeof(double);       // [kFtMaxBlockLen][kBdThreads]
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t nblocks = min(A.counters[C_N_FBLK], A.fblk_cap);
  uint32_t qh = 0, qt = 0;
  bool exhausted = false;
  BitAcc sa, ea;
  sa.init(A.s_bits);
  ea.init(A.e_bits);
  for (;;) {
    // ---- take the next block (one global atomic per 32 blocks per warp) ----
    uint32_t idx = 0xFFFFFFFFu;
    if (qh == qt && !exhausted) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&A.counters[C_CUR_FBLK], (uint32_t)kBdQueue);
      base = __shfl_sync(FULL, base, 0);
      if (base >= nblocks) exhausted = true;
      else {
        qh = base;
        qt = min(base + (uint32_t)kBdQueue, nblocks);
      }
    }
    if (exhausted) break;
    if (qh + lane < qt) idx = qh + lane;
    qh = qt;
    if (idx == 0xFFFFFFFFu) continue;
    // NOTE: a warp takes 32 blocks at a time and each lane runs its block to completion; lanes of a
    // warp therefore finish together only as well as their block lengths match (sentence-sized blocks).
    const uint4 desc = A.fblocks[idx];
    const ulonglong2* __restrict__ rp = reinterpret_cast<const ulonglong2*>(A.stream + desc.x);
    const unsigned long long* __restrict__ ovf = A.stream + desc.w;
    const uint32_t P0 = desc.y;
    const int npos = (int)desc.z;
    // ---- route DP, right to left: record i belongs to rune npos-1-i; records are prefetched two ahead ----
    ulonglong2 a0 = __ldcs(rp), b0 = __ldcs(rp + 1), a1 = a0, b1 = b0;
    if (npos > 1) {
      a1 = __ldcs(rp + 2);
      b1 = __ldcs(rp + 3);
    }
    for (int i = 0; i < npos; i++) {
      const int k = npos - 1 - i;
      ulonglong2 a2 = a1, b2 = b1;
      if (i + 2 < npos) {
        a2 = __ldcs(rp + 2 * (i + 2));
        b2 = __ldcs(rp + 2 * (i + 2) + 1);
      }
      uint32_t m = (uint32_t)a0.x;
      const uint32_t ovi = (uint32_t)(a0.x >> 48);
      double prev = JB_MINF, best_v = 0.0, v = 0.0;
      uint32_t best_d = 0, d = 0, j = 0;
      while (m) {
        d = __ffs(m);
        m &= m - 1;
        const unsigned long long wu = j == 0 ? a0.y : (j == 1 ? b0.x : (j == 2 ? b0.y : ovf[ovi + j - 3]));
        j++;
        const double nxt = (k + (int)d >= npos) ? 0.0 : ring[((k + d) & (RING - 1)) * kBdThreads + tid];  // {j,0.0} at the end (T:522)
        v = __longlong_as_double((long long)wu) + nxt;  // pieceFreq + nextBestPiece.proba (T:529)
        if (v >= prev) {  // maxIndexProba: compare with the PREVIOUS candidate (T:569)
          best_d = d;
          best_v = v;
        }
        prev = v;
      }
      if (best_d == 0) {  // best.index == -1 -> return prev (T:574-576)
        best_d = d;
        best_v = v;
      }
      ring[(k & (RING - 1)) * kBdThreads + tid] = best_v;
      path[k * kBdThreads + tid] = (uint8_t)best_d;
      a0 = a1;
      b0 = b1;
      a1 = a2;
      b1 = b2;
    }
    // ---- forward walk + HMM ----
    uint8_t* bp = reinterpret_cast<uint8_t*>(A.stream + desc.x);  // the block's records are dead now: Viterbi back-pointers
    int k = 0;
    uint32_t run_n = 0;
    int run_s = 0;
    double V[4];
    while (k < npos) {
      const uint32_t d = path[k * kBdThreads + tid] & 0x7F;
      const bool single = HMM && d == 1;
      if (single) {  // collect singletons (T:233-234): one Viterbi step per rune (T:688-719)
        const uint8_t* tp = A.text + P0 + 3u * (uint32_t)k;
        const uint32_t cp = ((tp[0] & 0xFu) << 12) | ((tp[1] & 0x3Fu) << 6) | (tp[2] & 0x3Fu);
        const double2* ep = reinterpret_cast<const double2*>(T.emit + (size_t)cp * 4);
        const double2 e0 = __ldg(ep), e1 = __ldg(ep + 1);
        const double em[4] = {e0.x, e0.y, e1.x, e1.y};
        if (run_n == 0) {
          run_s = k;
#pragma unroll
          for (int s = 0; s < 4; s++) V[s] = T.start[s] + em[s];
        } else {
          double W[4];
          uint32_t code = 0;
#pragma unroll
          for (int s = 0; s < 4; s++) {  // stateTransitionRoute (T:736-756): strict > from minFloat, list order
            const int pa = (s == 0 || s == 3) ? 2 : 0, pb = (s == 0 || s == 3) ? 3 : 1;
            const double r0 = V[pa] + T.trans[s][0], r1 = V[pb] + T.trans[s][1];
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
            W[s] = best + em[s];
            code |= from << (2 * s);
          }
#pragma unroll
          for (int s = 0; s < 4; s++) V[s] = W[s];
          bp[k] = (uint8_t)code;
        }
        run_n++;
      }
      if (!single || k + 1 >= npos) {
        if (HMM && run_n) {  // flush the run: viterbi's tail (T:723-729) + cutHMM (T:273-285)
          if (run_n == 1) {
            sa.set(P0 + 3u * run_s);
            ea.set(P0 + 3u * run_s + 2);
          } else {
            int st2 = V[2] > V[3] ? 2 : 3;
            int kb = run_s + (int)run_n - 1;
            uint32_t plen = 0;
            for (;;) {  // back-trace; stops early where route.from == "" (T:715-716)
              uint8_t& pk = path[kb * kBdThreads + tid];
              pk = (uint8_t)((pk & 0x7F) | (st2 >= 2 ? 0x80 : 0));
              plen++;
              if (kb == run_s) break;
              const int c = (bp[kb] >> (2 * st2)) & 3;
              if (c == 0) break;
              st2 = (st2 == 0 || st2 == 3) ? (c == 1 ? 2 : 3) : (c == 1 ? 0 : 1);
              kb--;
            }
            // path[j] applies to rune j (T:277-283): a short path drops the run's tail
            const int shift = (int)(run_n - plen);
            bool prev_es = true;
            for (uint32_t j2 = 0; j2 < plen; j2++) {
              const bool es = path[(run_s + shift + (int)j2) * kBdThreads + tid] & 0x80;
              const uint32_t qq = P0 + 3u * (uint32_t)(run_s + (int)j2);
              if (prev_es) sa.set(qq);
              if (es) ea.set(qq + 2);
              prev_es = es;
            }
          }
          run_n = 0;
        }
        if (!single) {
          sa.set(P0 + 3u * (uint32_t)k);
          ea.set(P0 + 3u * (uint32_t)(k + (int)d) - 1);
        }
      }
      k += d;
    }
    sa.flush();
    ea.flush();
  }
}

int launch_block_dp(const JbTables& T, const BlockDpArgs& A, bool hmm, int num_sms, cudaStream_t st) {
  static bool attr_done = false;
  const size_t sm16 = (size_t)16 * kBdThreads * 8 + (size_t)kFtMaxBlockLen * kBdThreads;
  const size_t sm32 = (size_t)32 * kBdThreads * 8 + (size_t)kFtMaxBlockLen * kBdThreads;
  if (!attr_done) {
    cudaFuncSetAttribute(k_block_dp<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm16);
    cudaFuncSetAttribute(k_block_dp<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm16);
    cudaFuncSetAttribute(k_block_dp<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm32);
    cudaFuncSetAttribute(k_block_dp<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm32);
    attr_done = true;
  }
  // RING >= max_delta is enough: R[k+d] is read before R[k] overwrites the same ring cell
  const bool r16 = T.max_delta <= 16;
  const unsigned grid = (unsigned)num_sms * (r16 ? 4u : 3u);
  if (r16) {
    if (hmm) k_block_dp<true, 16><<<grid, kBdThreads, sm16, st>>>(T, A);
    else k_block_dp<false, 16><<<grid, kBdThreads, sm16, st>>>(T, A);
  } else {
    if (hmm) k_block_dp<true, 32><<<grid, kBdThreads, sm32, st>>>(T, A);
    else k_block_dp<false, 32><<<grid, kBdThreads, sm32, st>>>(T, A);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_fused(const JbTables& T, const FusedArgs& A, uint32_t ntiles, bool hmm, cudaStream_t st) {
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(k_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
    cudaFuncSetAttribute(k_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FusedSmem));
    attr_done = true;
  }
  if (hmm) k_fused<true><<<ntiles, kFtThreads, sizeof(FusedSmem), st>>>(T, A);
  else k_fused<false><<<ntiles, kFtThreads, sizeof(FusedSmem), st>>>(T, A);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace jb
