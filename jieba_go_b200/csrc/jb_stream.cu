// Streaming fast path -- see jb_stream.cuh for the contract and DESIGN.md for the measurements.
// Reference functions restated here: Cut/splitText T:151-210, cutNonZh T:289-310, buildDag T:462-497,
// calcDagProba T:502-548, maxIndexProba T:565-578, findDagPath T:552-562, cutZh T:221-255,
// viterbi T:668-730, stateTransitionRoute T:736-756, cutHMM T:273-285  (T = /root/reference/tokenizer.go).
#include "jb_stream.cuh"

#include <stdlib.h>

#include "../../include/jieba_b200.h"

namespace jb {

#define FULL 0xFFFFFFFFu

// ==========================================================================================
// k_scan
// ==========================================================================================
enum : uint32_t { SC_HAN = 1, SC_ALNUM = 2, SC_SPACE = 3, SC_OTHER = 4, SC_INVALID = 5 };
#define SCLS(c, len) (uint8_t)(((c) << 3) | (len))

// ---- bulk asynchronous copy global -> shared (the TMA engine's 1-D form) with an mbarrier for completion ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

struct ScanSmem {
  uint8_t sb[kScRegion];
  uint32_t dsw[kScThreads + 2];  // dsw[j] <-> global word t0/32 - 2 + j; lane wj's own word is dsw[wj + 1]
  uint32_t HANL[kScThreads];     // Han rune starts
  uint32_t ALN[kScThreads];      // ASCII [a-zA-Z0-9] bytes
  uint32_t BND[kScThreads];      // first byte of every block (Han or not), end of text
  uint32_t HS[kScThreads];       // first rune of every Han block
  uint32_t HAN4[kScThreads];     // 4-byte Han rune starts (subset of HANL)
  uint32_t OTE[kScThreads + 1];  // last byte of every non-Han, non-alnum, non-space rune
  uint32_t S[kScThreads + 1], E[kScThreads + 1];
  uint32_t wsum[kScThreads / 32];
  uint32_t ends_base;
  int last_hs;  // tile-local byte of the last Han-block start in the tile, -1: none
  uint64_t mbar;  // completion of the tile's bulk copy
};
static_assert(kScRegion % 16 == 0 && kScLeft % 16 == 0 && kScTileBytes % 16 == 0, "bulk copies move multiples of 16 bytes between 16-byte aligned addresses");

__device__ __forceinline__ bool s_is_alnum(uint32_t c) { return (c - '0' < 10u) || ((c | 0x20) - 'a' < 26u); }
// unicode.IsSpace (T:302)
__device__ __forceinline__ bool s_is_space(uint32_t cp) {
  if (cp <= 0xFF) return (cp - 9u < 5u) || cp == 0x20 || cp == 0x85 || cp == 0xA0;
  if (cp > 0x3000u || cp < 0x1680u) return false;
  return cp == 0x1680 || (cp - 0x2000u <= 0xAu) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}
__device__ __forceinline__ bool s_is_han(uint32_t cp, const JbTables& T) {
  if (cp - 0x4E00u < 0x51A6u) return true;  // U+4E00..U+9FA5: Han in every Unicode version the tables cover
  if (cp < 0x10000) return (__ldg(T.han_bits + (cp >> 5)) >> (cp & 31)) & 1;
  for (uint32_t i = 0; i < T.n_supp; i++)
    if (cp >= T.supp_lo[i] && cp <= T.supp_hi[i]) return true;
  return false;
}

struct ScCtx {
  const ScanSmem* s;
  uint32_t t0, n;
  // is region index i a document start, or at/after the end of the text?
  __device__ __forceinline__ bool ds_at(int i) const {
    int64_t P = (int64_t)t0 - kScLeft + i;
    if (P >= (int64_t)n) return true;
    if (P < 0) return false;
    return (s->dsw[(i + 16) >> 5] >> ((i + 16) & 31)) & 1;
  }
};
// validated length (2..4) of the UTF-8 sequence whose lead byte is at region index i, 0 if ill-formed
// (Go's decoding rules: RFC 3629 ranges; a document boundary inside the sequence cuts it)
__device__ __forceinline__ int s_seqlen(const ScanSmem& S, const ScCtx& cx, int i) {
  uint32_t b = S.sb[i];
  int len = 0;
  if (b >= 0xC2 && b <= 0xDF) len = 2;
  else if (b >= 0xE0 && b <= 0xEF) len = 3;
  else if (b >= 0xF0 && b <= 0xF4) len = 4;
  if (!len || i + len > kScRegion) return 0;
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
// exact class of the byte at region index i: 0 = interior of a rune, else class << 3 | rune length
__device__ uint8_t s_byte_class(const ScanSmem& S, const ScCtx& cx, const JbTables& T, int i) {
  uint32_t b = S.sb[i];
  if (b < 0x80) return s_is_alnum(b) ? SCLS(SC_ALNUM, 1) : (s_is_space(b) ? SCLS(SC_SPACE, 1) : SCLS(SC_OTHER, 1));
  if ((b & 0xC0) == 0x80) {
    for (int k = 1; k <= 3 && i - k >= 0; k++)
      if ((S.sb[i - k] & 0xC0) != 0x80) return s_seqlen(S, cx, i - k) > k ? 0 : SCLS(SC_INVALID, 1);
    return SCLS(SC_INVALID, 1);
  }
  int len = s_seqlen(S, cx, i);
  if (!len) return SCLS(SC_INVALID, 1);
  const uint8_t* p = &S.sb[i];
  uint32_t cp;
  if (len == 2) cp = ((p[0] & 0x1Fu) << 6) | (p[1] & 0x3Fu);
  else if (len == 3) cp = ((p[0] & 0x0Fu) << 12) | ((p[1] & 0x3Fu) << 6) | (p[2] & 0x3Fu);
  else cp = ((p[0] & 0x07u) << 18) | ((p[1] & 0x3Fu) << 12) | ((p[2] & 0x3Fu) << 6) | (p[3] & 0x3Fu);
  return s_is_han(cp, T) ? SCLS(SC_HAN, len) : (s_is_space(cp) ? SCLS(SC_SPACE, len) : SCLS(SC_OTHER, len));
}

// 4 predicate bits (bit 7 of each byte of x) -> a nibble
__device__ __forceinline__ uint32_t s_mm4(uint32_t x) { return (((x >> 7) * 0x00204081u) >> 21) & 0xFu; }

struct ScWord {
  uint32_t c, l3, as, bad;  // nibbles: continuation bytes, E0..EF leads, ASCII; bad != 0: something else (or E0 / ED)
};
__device__ __forceinline__ ScWord s_word(uint32_t w) {
  const uint32_t M = 0x80808080u;
  const uint32_t t1 = w << 1, t2 = w << 2, t3 = w << 3;
  const uint32_t cont = w & ~t1 & M;
  const uint32_t l3 = w & t1 & t2 & ~t3 & M;
  uint32_t bad = w & t1 & (~t2 | t3) & M;  // 110xxxxx / 1111xxxx leads
  // E0 and ED leads restrict their second byte (and never start a Han rune): exact rules for those words.
  // (zero-byte test on the low nibbles; a false positive only sends a clean word the exact way)
  const uint32_t nib = w & 0x0F0F0F0Fu, nd = nib ^ 0x0D0D0D0Du;
  bad |= (((nib - 0x01010101u) & ~nib) | ((nd - 0x01010101u) & ~nd)) & l3;
  ScWord r;
  r.c = s_mm4(cont);
  r.l3 = s_mm4(l3);
  r.as = s_mm4(~w & M);
  r.bad = bad;
  return r;
}

// Does the non-Han block around bit b of word wj hold an ASCII alnum (cutNonZh T:290-293)?  Own words are
// 1..nown.  returns 1 yes, 0 no, else (2 | need_fwd<<2 | need_bwd<<3) when the block leaves the tile.
__device__ uint32_t s_block_alnum(const uint32_t* BND, const uint32_t* ALN, int w, int b, int nown) {
  uint32_t lowmask = (b == 31) ? 0xFFFFFFFFu : ((2u << b) - 1u);
  bool found = false;
  uint32_t m = BND[w] & lowmask;
  if (m) {
    int bb = 31 - __clz(m);
    if (ALN[w] & lowmask & ~((1u << bb) - 1u)) return 1;
    found = true;
  } else {
    if (ALN[w] & lowmask) return 1;
    for (int ww = w - 1; ww >= 1; --ww) {
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
    for (int ww = w + 1; ww <= nown; ++ww) {
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

__global__ void __launch_bounds__(kScThreads, 8) k_scan(const JbTables T, const ScanArgs A) {
  __shared__ __align__(16) ScanSmem S;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wj = tid;  // word index inside the staged region: 0 = halo before, 1..254 own, 255 = halo after
  const uint32_t tile = blockIdx.x;
  const uint32_t t0 = tile * (uint32_t)kScTileBytes;
  const uint32_t n = A.n;
  ScCtx cx{&S, t0, n};

  // ---- stage the bytes (+ halo) and the document-start words -----------------------------------
  // An interior tile of a 16-byte aligned text is ONE bulk asynchronous copy (cp.async.bulk, the TMA engine): a single
  // thread issues it, the mbarrier counts the bytes in, and meanwhile every thread fetches the document-start words and
  // clears its accumulators.  Tiles at either end of the text (and unaligned texts) are staged by the threads.
  const int64_t Pr = (int64_t)t0 - kScLeft;
  const bool bulk = ((reinterpret_cast<uintptr_t>(A.text) & 15) == 0) && Pr >= 0 && Pr + kScRegion <= (int64_t)n;
  if (bulk) {
    if (tid == 0) mbar_init(&S.mbar, 1);
    __syncthreads();
    if (tid == 0) bulk_load(S.sb, A.text + Pr, (uint32_t)kScRegion, &S.mbar);
  }
  {
    const bool aligned = ((reinterpret_cast<uintptr_t>(A.text) & 15) == 0);
    for (int c = tid; c < (bulk ? 0 : kScRegion / 16); c += kScThreads) {
      const int64_t P = (int64_t)t0 - kScLeft + c * 16;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (aligned && P >= 0 && P + 16 <= (int64_t)n) {
        v = __ldg(reinterpret_cast<const uint4*>(A.text + P));
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
    for (int j = tid; j < kScThreads + 2; j += kScThreads) {
      int64_t gw = (int64_t)(t0 / 32) - 2 + j;
      S.dsw[j] = (gw >= 0 && gw < (int64_t)nwords) ? __ldg(A.ds_bits + gw) : 0;
    }
    S.OTE[tid] = 0;
    S.S[tid] = 0;
    S.E[tid] = 0;
    if (tid == 0) {
      S.last_hs = -1;
      S.OTE[kScThreads] = 0;
      S.S[kScThreads] = 0;
      S.E[kScThreads] = 0;
    }
  }
  if (bulk) mbar_wait(&S.mbar, 0);
  __syncthreads();

  // ---- stage 1: per 32-byte word, bitmasks of rune starts / Han / alnum / other-token runes ----
  const int64_t Pw = (int64_t)t0 - 32 + 32 * (int64_t)wj;  // text position of this lane's word
  const int base = 16 + 32 * wj;                           // its region index
  const bool live = Pw >= 0 && Pw < (int64_t)n;
  const uint32_t D = S.dsw[wj + 1];
  uint32_t RS = 0, HANL = 0, HAN4 = 0, ALN = 0, OTH = 0;
  if (live) {
    const uint32_t VM = (Pw + 32 <= (int64_t)n) ? FULL : ((1u << (uint32_t)(n - Pw)) - 1u);
    const uint4 qa = *reinterpret_cast<const uint4*>(&S.sb[base]), qb = *reinterpret_cast<const uint4*>(&S.sb[base + 16]);
    const uint32_t wp = *reinterpret_cast<const uint32_t*>(&S.sb[base - 4]), wn = *reinterpret_cast<const uint32_t*>(&S.sb[base + 32]);
    const uint32_t ww[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
    uint32_t C = 0, L3 = 0, AS = 0, bad = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const ScWord r = s_word(ww[j]);
      C |= r.c << (4 * j);
      L3 |= r.l3 << (4 * j);
      AS |= r.as << (4 * j);
      bad |= r.bad;
    }
    const ScWord rp = s_word(wp), rn = s_word(wn);
    const uint32_t Cp = rp.c << 28, L3p = rp.l3 << 28, Cn = rn.c, Dn = S.dsw[wj + 2], Dp = S.dsw[wj];
    bad |= rp.bad | rn.bad;
    // every E? lead from byte -3 on is followed by two continuation bytes; every continuation byte is covered
    bad |= L3 & ~(__funnelshift_r(C, Cn, 1) & __funnelshift_r(C, Cn, 2));
    bad |= L3p & ~(__funnelshift_r(Cp, C, 1) & __funnelshift_r(Cp, C, 2)) & 0xE0000000u;
    bad |= C & ~(__funnelshift_l(L3p, L3, 1) | __funnelshift_l(L3p, L3, 2));
    // no document boundary inside a rune
    bad |= (D & C) | (Dn & Cn & 3u) | (Dp & Cp & 0x80000000u);
    if (Pw + 36 > (int64_t)n) bad = 1;
    if (!bad) {
      // ---- clean word: ASCII and well-formed 3-byte runes only ----
      RS = AS | L3;
      uint32_t SPC = 0;
      uint32_t m = L3, rest = 0;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int off = base + b;
        const uint32_t* wq = reinterpret_cast<const uint32_t*>(&S.sb[off & ~3]);
        const uint32_t x = __funnelshift_r(wq[0], wq[1], (off & 3) * 8);
        // U+4E00..U+9F7F straight on the bytes: (lead << 8 | second) in [E4 B8, E9 BD]
        const bool common = __byte_perm(x, 0u, 0x4401u) - 0xE4B8u <= 0xE9BDu - 0xE4B8u;
        HANL |= (common ? 1u : 0u) << b;
        rest |= (common ? 0u : 1u) << b;
      }
      // the other 3-byte runes (punctuation, other scripts, rare Han) in a loop of their own: a word has one or none,
      // so the lanes that have one meet here instead of holding up the loop above one at a time
      while (rest) {
        const int b = __ffs(rest) - 1;
        rest &= rest - 1;
        const int off = base + b;
        const uint32_t* wq = reinterpret_cast<const uint32_t*>(&S.sb[off & ~3]);
        const uint32_t x = __funnelshift_r(wq[0], wq[1], (off & 3) * 8);
        const uint32_t cp = ((x & 0xFu) << 12) | ((x >> 2) & 0xFC0u) | ((x >> 16) & 0x3Fu);
        if (s_is_han(cp, T)) HANL |= 1u << b;
        else if (s_is_space(cp)) SPC |= 1u << b;
      }
      m = AS;
      while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const uint32_t c = S.sb[base + b];
        if (s_is_alnum(c)) ALN |= 1u << b;
        else if (s_is_space(c)) SPC |= 1u << b;
      }
      OTH = RS & ~HANL & ~ALN & ~SPC;  // each its own token if the block holds an alnum (T:301-306)
      const uint32_t o3 = OTH & L3;
      const uint32_t ote = (OTH & AS) | (o3 << 2);
      if (ote) atomicOr(&S.OTE[wj], ote);
      if (o3 >> 30) atomicOr(&S.OTE[wj + 1], o3 >> 30);
    } else {
      // ---- anything else (2/4-byte runes, ill-formed bytes, runes cut by a document boundary, the
      // end of the text): Go's decoding rules byte by byte ----
      for (int b = 0; b < 32; b++) {
        if (!((VM >> b) & 1)) break;
        const uint8_t c = s_byte_class(S, cx, T, base + b);
        if (!c) continue;
        const uint32_t cl = c >> 3, len = c & 7;
        RS |= 1u << b;
        if (cl == SC_HAN) {
          HANL |= 1u << b;
          if (len == 4) HAN4 |= 1u << b;  // its block goes to k_wide
        } else if (cl == SC_ALNUM) {
          ALN |= 1u << b;
        } else if (cl != SC_SPACE) {
          OTH |= 1u << b;
          const int eb = b + (int)len - 1;
          atomicOr(&S.OTE[wj + (eb >> 5)], 1u << (eb & 31));
        }
      }
    }
  }
  S.HANL[wj] = HANL;
  S.HAN4[wj] = HAN4;
  S.ALN[wj] = ALN;
  __syncthreads();

  // ---- stage 2: block boundaries, Han block starts / ends, alnum-run tokens ----------------------
  const bool own = wj >= 1 && wj <= kScOwnWords && Pw <= (int64_t)n;
  uint32_t HE = 0;
  if (own) {
    const uint32_t Dn = S.dsw[wj + 2];
    // the rune right before is Han (3 or 4 bytes); a Han rune follows in the same document
    const uint32_t H4p = S.HAN4[wj - 1], NX = HANL & ~D, NXn = S.HANL[wj + 1] & ~Dn;
    const uint32_t prevHan = __funnelshift_l(S.HANL[wj - 1] & ~H4p, HANL & ~HAN4, 3) | __funnelshift_l(H4p, HAN4, 4);
    const uint32_t nextHan = (~HAN4 & __funnelshift_r(NX, NXn, 3)) | (HAN4 & __funnelshift_r(NX, NXn, 4));
    uint32_t BND = RS & (D | (HANL ^ prevHan));
    if ((int64_t)n < Pw + 32) BND |= 1u << (uint32_t)(n - Pw);  // end of the text
    const uint32_t HS = HANL & (D | ~prevHan);
    HE = HANL & ~nextHan;
    const uint32_t prevAl = __funnelshift_l(S.ALN[wj - 1], ALN, 1);
    const uint32_t nextAl = __funnelshift_r(ALN & ~D, S.ALN[wj + 1] & ~Dn, 1);
    S.BND[wj] = BND;
    S.S[wj] = ALN & (D | ~prevAl);  // an alnum run is one token (T:298-299)
    S.E[wj] = ALN & ~nextAl;
    S.HS[wj] = HS;
    if (HS) atomicMax(&S.last_hs, (wj - 1) * 32 + 31 - __clz(HS));
  } else {
    S.BND[wj] = 0;
    S.HS[wj] = 0;
  }
  // Han block ends -> list, in text order within the tile
  {
    const uint32_t cnt = __popc(HE);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) S.wsum[warp] = incl;
    __syncthreads();
    if (tid == 0) {
      uint32_t tot = 0;
#pragma unroll
      for (int j = 0; j < kScThreads / 32; j++) {
        uint32_t v = S.wsum[j];
        S.wsum[j] = tot;
        tot += v;
      }
      uint32_t b0 = 0;
      if (tot) {
        b0 = atomicAdd(&A.counters[C_N_BLK], tot);
        if (b0 + tot > A.blocks_cap) atomicOr(&A.counters[C_FLAGS], 1u);
      }
      S.ends_base = b0;
    }
    __syncthreads();
    uint32_t o = S.ends_base + S.wsum[warp] + incl - cnt;
    uint32_t m = HE;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      // the block's length in runes, when its first rune lies in this tile (else 0: k_route finds it in hs_bits)
      uint32_t nr = 0;
      uint32_t hm = S.HS[wj] & ((b == 31) ? FULL : ((2u << b) - 1u));
      int ws = wj;
      while (!hm && ws > 1) hm = S.HS[--ws];
      if (hm) nr = (uint32_t)((wj - ws) * 32 + b - (31 - __clz(hm))) / 3u + 1u;
      if ((HAN4 >> b) & 1) nr = kWideBlock;  // ends with a 4-byte rune: k_route hands it to k_wide at once
      if (o < A.blocks_cap) A.blocks[o] = make_uint2((uint32_t)Pw + b, nr);
      o++;
    }
  }

  // ---- stage 3: gated single-rune tokens (cutNonZh drops a block without [a-zA-Z0-9], T:290-293) ----
  if (own) {
    uint32_t m = OTH;
    while (m) {
      const int b = __ffs(m) - 1;
      m &= m - 1;
      // its last byte: first OTE bit at or after it
      int we = wj;
      uint32_t em = S.OTE[we] & ~((1u << b) - 1u);
      if (!em) em = S.OTE[++we];
      const int eb = __ffs(em) - 1;
      const uint32_t r = s_block_alnum(S.BND, S.ALN, wj, b, kScOwnWords);
      if (r == 1) {
        atomicOr(&S.S[wj], 1u << b);
        atomicOr(&S.E[we], 1u << eb);
      } else if (r & 2) {
        const uint32_t idx = atomicAdd(&A.counters[C_N_DEFER], 1u);
        if (idx < A.deferred_cap)
          A.deferred[idx] = make_uint4((uint32_t)Pw + b, (uint32_t)((we - wj) * 32 + eb - b + 1), (r >> 2) & 3u, tile);
        else atomicOr(&A.counters[C_FLAGS], 1u);
      }
    }
  }
  // tile summary for k_tile_scan: has a boundary / alnum before the first / alnum after the last
  if (warp == 0) {
    int firstw = kScThreads, lastw = -1;
    for (int j = 1 + lane; j <= kScOwnWords; j += 32)
      if (S.BND[j]) {
        firstw = min(firstw, j);
        lastw = max(lastw, j);
      }
    firstw = __reduce_min_sync(FULL, firstw);
    lastw = __reduce_max_sync(FULL, lastw);
    bool pre = false, post = false;
    if (lastw < 0) {
      for (int j = 1 + lane; j <= kScOwnWords; j += 32) pre |= S.ALN[j] != 0;
      post = pre;
    } else {
      const uint32_t fb = __ffs(S.BND[firstw]) - 1, lb = 31 - __clz(S.BND[lastw]);
      for (int j = 1 + lane; j <= kScOwnWords; j += 32) {
        const uint32_t a = S.ALN[j];
        if (j < firstw) pre |= a != 0;
        if (j == firstw) pre |= (a & ((1u << fb) - 1u)) != 0;
        if (j > lastw) post |= a != 0;
        if (j == lastw) post |= (a & ~((1u << lb) - 1u)) != 0;
      }
    }
    pre = __any_sync(FULL, pre);
    post = __any_sync(FULL, post);
    if (lane == 0) {
      A.tile_sum[tile] = (uint8_t)((lastw >= 0 ? 1 : 0) | (pre ? 2 : 0) | (post ? 4 : 0));
      A.tile_last_hs[tile] = S.last_hs >= 0 ? t0 + (uint32_t)S.last_hs : 0xFFFFFFFFu;
    }
  }
  __syncthreads();
  // ---- publish the non-Han token bits (a rune may end in the next tile's first word) --------------
  if (wj >= 1 && wj <= kScOwnWords + 1) {
    const int64_t gw = (int64_t)(t0 / 32) - 1 + wj;
    const uint32_t sbits = S.S[wj], ebits = S.E[wj];
    if (sbits) atomicOr(&A.s_bits[gw], sbits);
    if (ebits) atomicOr(&A.e_bits[gw], ebits);
  }
}

int launch_scan(const JbTables& T, const ScanArgs& A, cudaStream_t st) {
  k_scan<<<scan_tiles(A.n), kScThreads, 0, st>>>(T, A);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ==========================================================================================
// k_route: buildDag + calcDagProba + maxIndexProba, one lane per Han block, right to left,
// ONE POSITION PER LOOP ITERATION per lane, in straight-line code:
//   * when a lane advances to a rune it issues every load the position will need and looks at none of
//     them: the first-rune table entry (T:468-472), the hash entries of the 2- and 3-rune prefixes (their
//     slots depend on the runes only -- jb_hash_next -- so they are fetched speculatively, before the gate /
//     Bloom tests say whether buildDag would look at them) and the next 8 bytes of text;
//   * the next iteration consumes them: candidates in ascending length, pieceFreq + next.proba (T:519-529),
//     folded into maxIndexProba's running (prev, best) pair (T:565-578); a 4th rune or a hash collision
//     (about one position in eight) goes on in a short loop of dependent probes;
//   * commit: selected (length, value) -> route ring + 4-bit path entry; first rune of the block -> the block
//     is handed to k_emit and the lane takes the next block from the warp's queue.
// All 32 lanes stay busy whatever the block lengths.  Per-lane shared memory: the last RING route values and runes.
// ==========================================================================================
constexpr int kRtThreads = 128;
constexpr int kRtQueue = 32;

template <int RING, int PB, bool DBG>
__global__ void __launch_bounds__(kRtThreads, 8) k_route(const JbTables T, const RouteArgs A) {
  __shared__ double ring[RING][kRtThreads];
  __shared__ uint16_t rr[RING][kRtThreads];  // (3-byte runes: BMP) -- shared memory not used here is L1 for the tables
  constexpr uint32_t M = RING - 1;
  constexpr uint32_t PPW = 32 / PB;  // path entries per word
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  const uint32_t nblocks = min(A.counters[C_N_BLK], A.blocks_cap);
  if (A.counters[C_FLAGS] & 1u) return;  // the general pipeline redoes this batch
  // few blocks (long ones): spread them over all warps instead of filling a few warps
  const uint32_t nwarps = gridDim.x * (kRtThreads / 32);
  const uint32_t chunk = min((uint32_t)kRtQueue, max(A.min_chunk, (nblocks + nwarps - 1) / nwarps));
  // text is read through 8-byte aligned words: offsets are relative to the aligned base
  const uint32_t tmis = (uint32_t)(reinterpret_cast<uintptr_t>(A.text) & 7);
  const uint2* __restrict__ text8 = reinterpret_cast<const uint2*>(A.text - tmis);
  const uint4* __restrict__ first = reinterpret_cast<const uint4*>(T.first);
  const uint4* __restrict__ entries = reinterpret_cast<const uint4*>(T.entries);
  const uint32_t hmask = T.hash_mask, hshift = T.hash_shift;
  double* const sring = &ring[0][tid];   // this lane's ring cells: index * kRtThreads
  uint16_t* const srr = &rr[0][tid];
  uint32_t qh = 0, qt = 0;
  bool exhausted = false, active = false;
  uint32_t bi = 0, p = 0, kq = 0, e3i = 0, nr = 0;  // block index, lead byte of the current rune (+ tmis), runes to its right, end / 3, runes (0: unknown)
  uint32_t r0 = 0;                                  // the current rune
  // the three route values and runes to the right of the current one (a fresh lane needs nothing else from the rings)
  double R1 = 0.0, R2 = 0.0;
  uint32_t q12 = 0;  // rune 1 to the right | rune 2 to the right << 16
  uint4 f = make_uint4(0, 0, 0, 0), e2 = f, e3 = f; // in flight: first-rune entry, entries of the 2- and 3-rune prefixes
  uint32_t h2 = 0, h3 = 0;                          // their hash states
  uint32_t w0 = 0, w1 = 0, wn = 0, wc = 0xFFFFFFFFu;  // text window: 8-byte word wc in (w0, w1), first 4 bytes of word wc + 1 in wn
  uint2 tp = make_uint2(0, 0);                        // in flight: 8-byte word wc - 1
  uint32_t acc = 0, accw = 0xFFFFFFFFu;
  // A position that needs more probes than the two prefetched ones stays for further iterations (one dependent
  // probe each, in flight in e2); meanwhile its selector state is parked in the registers of f / e3 / h2 / h3.
  bool chain = false;
  uint32_t cs = 0;  // L | home-slot flag << 7 | best_d << 16

  // the lane now stands on the rune at p: decode it from the window, issue its table loads (first = true: last
  // rune of a block, nothing to its right)
  auto setup_pos = [&](const bool first_rune) -> bool {
    const uint32_t o = p & 7u;
    const uint32_t x = __funnelshift_r(o & 4u ? w1 : w0, o & 4u ? wn : w1, (o & 3u) * 8u);
    q12 = (q12 << 16) | r0;
    r0 = ((x & 0xFu) << 12) | ((x >> 2) & 0xFC0u) | ((x >> 16) & 0x3Fu);
    srr[(kq & M) * kRtThreads] = (uint16_t)r0;
    f = ldg_keep(first + r0);
    if (!first_rune) {  // (with one rune to the right the 3-rune slot is computed from a stale rune: loaded, never looked at)
      h2 = jb_hash_next(JB_PARENT_FIRST(r0), q12 & 0xFFFFu);
      e2 = __ldg(entries + (h2 >> hshift));
      h3 = jb_hash_next(h2, q12 >> 16);
      e3 = __ldg(entries + (h3 >> hshift));
    }
    return (x & 0xF0u) == 0xE0u;  // inside a Han block every rune has 3 bytes -- or 4 (k_wide)
  };

  for (;;) {
    // ---- refill idle lanes from the warp's queue of block indexes ----
    const uint32_t nm = __ballot_sync(FULL, !active);
    if (nm) {
      if (qh == qt && !exhausted) {
        uint32_t b0 = 0;
        if (lane == 0) b0 = atomicAdd(&A.counters[C_CUR_ROUTE], chunk);
        b0 = __shfl_sync(FULL, b0, 0);
        if (b0 >= nblocks) exhausted = true;
        else {
          qh = b0;
          qt = min(b0 + chunk, nblocks);
        }
      }
      if (!active) {
        const uint32_t mine = qh + __popc(nm & lt_mask);
        if (mine < qt) {
          bi = mine;
          const uint2 bd = A.blocks[bi];
          e3i = bd.x / 3u;
          nr = bd.y;
          if (nr == kWideBlock) {  // ends with a 4-byte rune
            const uint32_t wi = atomicAdd(&A.counters[C_N_WIDE], 1u);
            if (wi < A.wide_cap) A.wide_list[wi] = bd.x;
            else atomicOr(&A.counters[C_FLAGS], 1u);
            A.blocks[bi].y = 0;  // nothing for k_emit
          } else {
          if (nr == 0) {  // the block began in an earlier k_scan tile: the nearest tile with a block start holds it
            uint32_t t = bd.x / (uint32_t)kScTileBytes, sp;
            do sp = __ldg(A.tile_last_hs + --t);
            while (sp == 0xFFFFFFFFu);
            nr = (bd.x - sp) / 3u + 1u;
          }
          p = bd.x + tmis;
          kq = 0;
          sring[M * kRtThreads] = 0.0;  // R[-1]: {j, 0.0} at the end of the block (T:522) -- cell M is not written before rune M
          R1 = 0.0;
          wc = p >> 3;
          const uint2 t0 = __ldg(text8 + wc);
          w0 = t0.x;
          w1 = t0.y;
          if ((p & 7u) >= 6u) wn = __ldg(reinterpret_cast<const uint32_t*>(text8 + wc + 1));  // the rune's tail (never past the text)
          if (wc) tp = __ldg(text8 + wc - 1);
          setup_pos(true);
          chain = false;
          active = true;
          }
        }
      }
      qh = min(qt, qh + (uint32_t)__popc(nm));
      if (exhausted && __all_sync(FULL, !active)) break;
    }
    // ---- one straight-line, predicated section for every active lane.  A FRESH lane (it advanced last iteration)
    // holds its first-rune entry (T:468-472) and the entries of its 2- and 3-rune prefixes; a CHAINED lane holds
    // the next entry of its prefix chain in e2 and its selector state parked in f / e3 / h2 / h3 / cs.
    // Candidates in ascending length: pieceFreq + nextBestPiece.proba (T:519-529) into maxIndexProba's running
    // (prev, best) pair: each candidate is compared with the previous one, first with minFloat (T:565-578). ----
    double best_v, prev_v;  // (all of these are written before they are read whenever the lane is active)
    uint32_t best_d, L, parent, slot, hs;
    bool more = false, home = false;  // home: the next probe looks at the home slot of its key
    if (active) {
      const bool chained = chain, fresh = !chain;
      const uint32_t L0 = chained ? (cs & 0x7Fu) : 1u;  // runes of the prefix matched so far
      // A key of LA runes needs LA - 1 runes to the right of this one: kq of them exist (beyond the block's end the rings
      // hold stale runes).  Nothing else bounds the chain: a gated first rune has an empty Bloom filter (jb_host.cpp), and
      // where no longer key exists the entry's filter is empty too.
      const uint32_t parA = chained ? f.z : JB_PARENT_FIRST(r0);
      const uint32_t slotA = chained ? h3 : (h2 >> hshift), slot3 = h3 >> hshift;
      const bool homeA = fresh || (cs & 0x80u);
      // route values / runes 1, 2, 3 positions to the right: registers for a fresh lane; a chained lane (L0 >= 2) reads the
      // rings at L0 and L0 + 1
      double RA = R1, RB = R2;
      uint32_t rA = q12 & 0xFFFFu, rB = q12 >> 16;  // the runes after a prefix of L0, L0 + 1 (and 3: rC) runes
      const uint32_t kC = ((kq - 3u) & M) * kRtThreads;
      const double RC = sring[kC];
      const uint32_t rC = srr[kC];
      if (chained) {
        const uint32_t kA = ((kq - L0) & M) * kRtThreads, kB = ((kq - L0 - 1u) & M) * kRtThreads;
        RA = sring[kA];
        RB = sring[kB];
        rA = srr[kA];
        rB = srr[kB];
      }
      const double wt1 = __longlong_as_double(((long long)f.y << 32) | (long long)f.x);
      const double wtA = __longlong_as_double(((long long)e2.y << 32) | (long long)e2.x);
      const double wt3 = __longlong_as_double(((long long)e3.y << 32) | (long long)e3.x);
      // candidate (i, i+1) of a fresh lane, or the parked selector state.  (R[-1] = 0.0 sits in the ring: no special
      // case for a word that ends the block.)  maxIndexProba's "best.index == -1 -> return prev" (T:574-576) needs no
      // (last) pair here: weights are finite or -Inf and never NaN (negative counts are rejected), so the first
      // candidate fails v >= minFloat only when it is -Inf, any further candidate then passes v >= -Inf, and the
      // fallback can only fire for a position whose ONLY candidate is the single rune -- it returns that one.
      const double v1 = wt1 + RA;
      best_v = chained ? __hiloint2double((int)e3.y, (int)e3.x) : v1;
      prev_v = chained ? wt1 : v1;
      // (v1 >= minFloat fails only for -Inf: route values of real text stay above -1e11, nowhere near -3.14e100; the
      // test is done on the high word, as are the freq > 0 tests below: -Inf is the only weight with these bits)
      best_d = chained ? ((cs >> 16) & 0xFFu) : ((uint32_t)__double2hiint(v1) != 0xFFF00000u ? 1u : 0u);
      // which entries buildDag looks at (T:469-482), and what it finds
      const bool gA = chained || (kq >= 1u && ((f.w >> jb_bloom_bit(rA)) & 1u));
      const bool mA = gA && e2.z == parA && JB_RB_RUNE(e2.w) == rA;
      // a foreign entry: linear probing goes on (past the home slot only if a key was displaced from it); empty: break
      const bool xA = gA && !mA && e2.z != JB_PARENT_EMPTY && (!homeA || (e2.w & JB_RB_CONT));
      const uint32_t LA = L0 + 1u;
      const bool cA = mA && e2.y != 0xFFF00000u;  // val > 0 -> edge (T:479-481): a freq-0 key carries -Inf
      const double vA = wtA + RB;
      const bool bA = cA && vA >= prev_v;
      best_d = bA ? LA : best_d;
      best_v = bA ? vA : best_v;
      prev_v = cA ? vA : prev_v;
      const bool contA = mA && LA <= kq && (((e2.w >> 21) >> jb_bloom11(rB)) & 1u);  // some key extends the prefix by rB
      // the 3-rune prefix of a fresh lane is here already
      const bool gB = contA && fresh;
      const bool mB = gB && e3.z == slotA && JB_RB_RUNE(e3.w) == rB;
      const bool xB = gB && !mB && e3.z != JB_PARENT_EMPTY && (e3.w & JB_RB_CONT);
      const bool cB = mB && e3.y != 0xFFF00000u;
      const double vB = wt3 + RC;
      const bool bB = cB && vB >= prev_v;
      best_d = bB ? 3u : best_d;
      best_v = bB ? vB : best_v;
      prev_v = cB ? vB : prev_v;
      const bool contB = mB && kq >= 3u && (((e3.w >> 21) >> jb_bloom11(rC)) & 1u);
      // anything beyond goes on in later iterations, one dependent probe each
      const bool deeperA = contA && chained;
      more = xA || deeperA || xB || contB;
      home = deeperA || contB;
      L = mB ? 3u : (mA ? LA : L0);
      parent = contB ? slot3 : ((deeperA || xB) ? slotA : parA);
      const uint32_t hn = jb_hash_next(contB ? h3 : h2, contB ? rC : rB);
      hs = home ? hn : (xB ? h3 : h2);
      slot = home ? (hn >> hshift) : (((xB ? slot3 : slotA) + 1u) & hmask);
    }
    __syncwarp();
    if (active && more) {  // park the selector state, issue the next probe
      e3.x = (uint32_t)__double2loint(best_v);
      e3.y = (uint32_t)__double2hiint(best_v);
      f.x = (uint32_t)__double2loint(prev_v);
      f.y = (uint32_t)__double2hiint(prev_v);
      f.z = parent;
      h2 = hs;
      h3 = slot;
      cs = L | (home ? 0x80u : 0u) | (best_d << 16);
      e2 = __ldg(entries + slot);
      chain = true;
    }
    __syncwarp();
    if (active && !more) {
      chain = false;
      // ---- commit the position ----
      best_d = max(best_d, 1u);  // best.index == -1 -> return prev (T:574-576): the lone single-rune candidate, see above
      sring[(kq & M) * kRtThreads] = best_v;
      R2 = R1;
      R1 = best_v;
      const uint32_t idx = e3i - kq, pwd = idx / PPW;
      if (pwd != accw) {
        if (acc) atomicOr(&A.path[accw], acc);
        acc = 0;
        accw = pwd;
      }
      acc |= (best_d - 1u) << ((idx % PPW) * PB);
      if (DBG) {  // (jb_debug_route only: a separate instantiation, nothing of it in the production kernel)
        A.dbg_R[idx] = best_v;
        A.dbg_D[idx] = (uint8_t)best_d;
      }
      if (kq + 1u == nr) {  // first rune of the block
        if (acc) atomicOr(&A.path[accw], acc);
        acc = 0;
        accw = 0xFFFFFFFFu;
        A.blocks[bi] = make_uint2(p - tmis, kq + 1u);
        active = false;
      } else {
        // ---- advance to the rune on the left ----
        p -= 3u;
        kq++;
        if ((p >> 3) != wc) {  // into the next 8-byte word (fetched when the lane entered this one)
          wn = w0;
          w0 = tp.x;
          w1 = tp.y;
          wc--;
          if (wc) tp = __ldg(text8 + wc - 1);
        }
        if (!setup_pos(false)) {  // not the lead of a 3-byte rune: the lane stepped into a 4-byte Han rune -> k_wide
          const uint32_t wi = atomicAdd(&A.counters[C_N_WIDE], 1u);
          if (wi < A.wide_cap) A.wide_list[wi] = A.blocks[bi].x;  // (still the block's last rune)
          else atomicOr(&A.counters[C_FLAGS], 1u);
          A.blocks[bi].y = 0;  // nothing for k_emit
          if (acc) atomicOr(&A.path[accw], acc);  // (path entries of the abandoned part are never read)
          acc = 0;
          accw = 0xFFFFFFFFu;
          active = false;
        }
      }
    }
    __syncwarp();
  }
}

int launch_route(const JbTables& T, const RouteArgs& A, int num_sms, cudaStream_t st) {
  // Shared memory that the 8 resident CTAs do not take is L1 for the dictionary tables, and the kernel is bound by
  // the latency of its table loads: with 20.5 KB per CTA the driver's default carveout leaves ~84 KB of L1
  // (6.4 ms/GB); with 24.5 KB per CTA (32-bit runes) it had to take the whole array (9.0 ms/GB).  Forcing other
  // carveouts or fewer CTAs measured slower.
  const bool r16 = T.max_delta <= 16;
  const unsigned grid = (unsigned)num_sms * 8u;
  if (A.dbg_R) {  // jb_debug_route: the instantiation that also records every selected route value
    if (r16) k_route<16, 4, true><<<grid, kRtThreads, 0, st>>>(T, A);
    else k_route<32, 8, true><<<grid, kRtThreads, 0, st>>>(T, A);
  } else if (r16) k_route<16, 4, false><<<grid, kRtThreads, 0, st>>>(T, A);
  else k_route<32, 8, false><<<grid, kRtThreads, 0, st>>>(T, A);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

// ==========================================================================================
// k_emit: findDagPath + cutZh/viterbi/cutHMM, one lane per Han block -> token start / end bits
// ==========================================================================================
constexpr int kEmThreads = 128;

// MODE 0: HMM off.  MODE 1: HMM on, Viterbi inside the walk (one lane carries a block's runs one after the other).
// MODE 2: HMM on, the walk only -- multi-rune pieces become tokens, single-rune pieces are MARKED in m_bits (bit at the
// rune's lead byte) and k_runs routes every run of marked runes on its own lane afterwards.  This is how the SEGMENTS of
// long blocks are emitted (below): a run of single runes may cross a segment boundary, a mark may not care.
//
// Long blocks.  One lane per block walks a 10k-rune block hop by hop while most of the machine idles.  The walk is a
// pointer chase k -> k + d(k), but it is the same chase from wherever it is entered: k_emit<0/1> therefore only CUTS a
// block of kLongRunes runes or more into segments of kSegRunes runes (A.segs, A.longs); k_land computes, per segment and
// for each of its first 16 runes, where the walk entered there leaves the segment; k_chain threads the true entry
// through a block's segments; and a second k_emit launch (MODE 2 or 0) walks the segments, one lane each.
constexpr uint32_t kSegRunes = 256, kLongRunes = 2 * kSegRunes;

template <int MODE, int PB>
__global__ void __launch_bounds__(kEmThreads) k_emit(const JbTables T, const EmitArgs A) {
  constexpr bool HMM = MODE == 1;
  constexpr uint32_t PPW = 32 / PB, PMASK = (1u << PB) - 1u;
  constexpr uint32_t kRegRun = 24;  // the four best paths of runs up to this length are carried in registers
  const int lane = threadIdx.x & 31;
  const uint32_t lt_mask = (1u << lane) - 1u;
  if (A.counters[C_FLAGS] & 1u) return;
  const uint32_t nblocks = min(A.counters[A.count_idx], A.blocks_cap);
  // few blocks (long ones): spread them over all warps instead of filling a few warps
  const uint32_t nwarps = gridDim.x * (kEmThreads / 32);
  const uint32_t chunk = min(32u, max(A.min_chunk, (nblocks + nwarps - 1) / nwarps));
  const uintptr_t tbase = reinterpret_cast<uintptr_t>(A.text);
  BitAcc2 sa, ea, ma;  // HMM off: every lane emits a token per iteration, merged per 32-byte word
  sa.init(A.s_bits);
  ea.init(A.e_bits);
  ma.init(A.m_bits);
  // HMM on: the lanes that emit in a given iteration are few and spread over several code sites; the per-lane word
  // accumulator then costs more than it saves: straight atomics
  auto set_s = [&](uint32_t q) {
    if (HMM) atomicOr(&A.s_bits[q >> 5], 1u << (q & 31));
    else sa.set(q);
  };
  auto set_e = [&](uint32_t q) {
    if (HMM) atomicOr(&A.e_bits[q >> 5], 1u << (q & 31));
    else ea.set(q);
  };
  uint32_t qh = 0, qt = 0, qbase = 0, qpw = 0;
  uint2 qdesc = make_uint2(0u, 0u);
  bool exhausted = false, active = false;
  uint32_t P0 = 0, i0 = 0, npos = 0, k = 0;
  uint32_t pw = 0, pw2 = 0, pt = 0;
  uint32_t run_n = 0, run_s = 0;
  uint32_t pend = 0xFFFFFFFFu;  // MODE 2: a single rune whose right neighbour is not known yet
  bool in_run = true;           //         the previous piece was a marked single rune
  double V[4] = {0.0, 0.0, 0.0, 0.0};
  // viterbi's fullPath (T:715-716) in bit form, per state: bits 0..23 = which runes of its best path are E or S
  // (token ends), bits 24..31 = the path's length (a route with from == "" restarts it)
  uint32_t pm[4] = {0, 0, 0, 0};
  for (;;) {
    // ---- refill idle lanes from the warp's queue of block indexes ----
    const uint32_t nm = __ballot_sync(FULL, !active);
    if (nm) {
      if (qh == qt && !exhausted) {
        uint32_t b0 = 0;
        if (lane == 0) b0 = atomicAdd(&A.counters[A.cursor_idx], chunk);
        b0 = __shfl_sync(FULL, b0, 0);
        if (b0 >= nblocks) exhausted = true;
        else {
          qh = qbase = b0;
          qt = min(b0 + chunk, nblocks);
          // the chunk's descriptors in one coalesced load, and each block's first path word behind it: a lane that
          // takes a block later gets both by shuffle instead of waiting for two dependent loads per block
          qdesc = b0 + lane < qt ? A.blocks[b0 + lane] : make_uint2(0u, 0u);
          qpw = A.path[(qdesc.x / 3u) / PPW];
        }
      }
      const uint32_t mine = qh + __popc(nm & lt_mask);
      const uint32_t src = (mine - qbase) & 31u;
      const uint32_t dx = __shfl_sync(FULL, qdesc.x, src), dy = __shfl_sync(FULL, qdesc.y, src), dw = __shfl_sync(FULL, qpw, src);
      if (!active) {
        if (mine < qt) {
          P0 = dx;
          npos = dy;
          i0 = P0 / 3u;
          k = 0;
          run_n = 0;
          pt = i0 / PPW;
          pw = dw;
          pw2 = A.path[pt + 1u];
          pend = 0xFFFFFFFFu;
          in_run = true;
          active = npos != 0;  // 0: the block went to k_wide
          if (MODE != 2 && A.segs && npos >= kLongRunes) {  // a long block: cut it into segments for the second launch
            const uint32_t nseg = (npos + kSegRunes - 1) / kSegRunes;
            const uint32_t s0 = atomicAdd(&A.counters[C_N_SEG], nseg), li = atomicAdd(&A.counters[C_N_LONG], 1u);
            if (s0 + nseg <= A.segs_cap && li < A.longs_cap) {
              for (uint32_t j = 0; j < nseg; j++)
                A.segs[s0 + j] = make_uint2(P0 + 3u * kSegRunes * j, min(kSegRunes, npos - kSegRunes * j));
              A.longs[li] = make_uint2(s0, nseg);
              active = false;
            } else {
              atomicOr(&A.counters[C_FLAGS], 1u);  // (the lists are sized so that this cannot happen)
            }
          }
        }
      }
      qh = min(qt, qh + (uint32_t)__popc(nm));
      if (exhausted && __all_sync(FULL, !active)) break;
    }
    // ---- one piece of findDagPath's walk (T:552-562) per iteration ----
    if (active) {
      const uint32_t pi = i0 + k;
      if (pi / PPW != pt) {  // the next word of the path is fetched when the lane enters this one
        const uint32_t t = pi / PPW;
        pw = t == pt + 1u ? pw2 : A.path[t];
        pt = t;
        pw2 = A.path[pt + 1u];
      }
      const uint32_t d = ((pw >> ((pi % PPW) * PB)) & PMASK) + 1u;
      const bool single = HMM && d == 1;
      if (single) {  // collect singletons (T:233-234): one Viterbi step per rune (T:688-719)
        const uintptr_t ap = tbase + P0 + 3u * k, a4 = ap & ~(uintptr_t)3;
        const uint32_t xl = __ldg(reinterpret_cast<const uint32_t*>(a4));
        const uint32_t xh = (ap & 3) >= 2 ? __ldg(reinterpret_cast<const uint32_t*>(a4 + 4)) : 0u;  // (never a word past the text)
        const uint32_t x = __funnelshift_r(xl, xh, (uint32_t)(ap & 3) * 8u);
        const uint32_t cp = ((x & 0xFu) << 12) | ((x >> 2) & 0xFC0u) | ((x >> 16) & 0x3Fu);
        const double2* ep = reinterpret_cast<const double2*>(T.emit + (size_t)cp * 4);
        const double2 e0 = __ldg(ep), e1 = __ldg(ep + 1);
        const double em[4] = {e0.x, e0.y, e1.x, e1.y};
        if (run_n == 0) {
          run_s = k;
#pragma unroll
          for (int s = 0; s < 4; s++) V[s] = T.start[s] + em[s];
          pm[0] = pm[1] = 1u << 24;
          pm[2] = pm[3] = (1u << 24) | 1u;
        } else {
          double W[4];
          uint32_t code = 0, npm[4];
          const uint32_t step = (1u << 24) | (run_n < kRegRun ? (1u << run_n) : 0u);  // one more entry; E and S end a token
#pragma unroll
          for (int s = 0; s < 4; s++) {  // stateTransitionRoute (T:736-756): strict > from minFloat, list order
            const int pa = (s == 0 || s == 3) ? 2 : 0, pb = (s == 0 || s == 3) ? 3 : 1;
            const double r0 = V[pa] + T.trans[s][0], r1 = V[pb] + T.trans[s][1];
            const bool t0 = r0 > JB_MINF;   // from = 1
            const double b0 = t0 ? r0 : JB_MINF;
            const bool t1 = r1 > b0;        // from = 2
            W[s] = (t1 ? r1 : b0) + em[s];
            code |= (t1 ? 2u : (t0 ? 1u : 0u)) << (2 * s);
            // fullPath[s] = fullPath[route.from] + [s]; fullPath[""] is nil (T:715-716)
            npm[s] = (t1 ? pm[pb] : (t0 ? pm[pa] : 0u)) + (s >= 2 ? step : (1u << 24));
          }
#pragma unroll
          for (int s = 0; s < 4; s++) {
            V[s] = W[s];
            pm[s] = npm[s];
          }
          A.bp[pi] = (uint8_t)code;  // only read back for runs longer than the register window
        }
        run_n++;
      }
      if (HMM && run_n && (!single || k + 1 >= npos)) {  // flush the run: viterbi's tail (T:723-729) + cutHMM (T:273-285)
        if (run_n == 1) {
          set_s(P0 + 3u * run_s);
          set_e(P0 + 3u * run_s + 2);
        } else if (run_n <= kRegRun) {
          const uint32_t pf = V[2] > V[3] ? pm[2] : pm[3];  // T:723-729
          const uint32_t plen = pf >> 24;
          // path[j] applies to rune j (T:277-283): a short path drops the run's tail
          const uint32_t lm = (1u << plen) - 1u;
          const uint32_t es = ((pf & 0xFFFFFFu) >> (run_n - plen)) & lm;
          const uint32_t starts = ((es << 1) | 1u) & lm;
          const uint32_t q0 = P0 + 3u * run_s;
          // rune j starts at q0 + 3j and ends at q0 + 3j + 2: spread the masks by 3 and OR them into the bitmaps
          if (plen <= 10u) {  // almost every run: 30 bits, two words per bitmap
            uint32_t x = es;  // bit j -> bit 3j
            x = (x | (x << 16)) & 0x030000FFu;
            x = (x | (x << 8)) & 0x0300F00Fu;
            x = (x | (x << 4)) & 0x030C30C3u;
            x = (x | (x << 2)) & 0x09249249u;
            const uint32_t sm = ((x << 3) | 1u) & ((1u << (3u * plen)) - 1u);  // a token starts after every end, and at rune 0
            const uint32_t sh = q0 & 31u, wq = q0 >> 5;
            const unsigned long long ss = (unsigned long long)sm << sh, ee = ((unsigned long long)x << 2) << sh;
            if ((uint32_t)ss) atomicOr(&A.s_bits[wq], (uint32_t)ss);
            if ((uint32_t)(ss >> 32)) atomicOr(&A.s_bits[wq + 1u], (uint32_t)(ss >> 32));
            if ((uint32_t)ee) atomicOr(&A.e_bits[wq], (uint32_t)ee);
            if ((uint32_t)(ee >> 32)) atomicOr(&A.e_bits[wq + 1u], (uint32_t)(ee >> 32));
          } else {
            or_span(A.s_bits, q0, spread3(starts), spread3(starts >> 16));
            or_span(A.e_bits, q0 + 2u, spread3(es), spread3(es >> 16));
          }
        } else {
          int st2 = V[2] > V[3] ? 2 : 3;
          uint32_t kb = run_s + run_n - 1;
          uint32_t plen = 0;
          for (;;) {
            const uint8_t code = A.bp[i0 + kb];
            A.bp[i0 + kb] = (uint8_t)(st2 >= 2 ? 0x80 : 0);  // the state of this path entry is E or S
            plen++;
            if (kb == run_s) break;
            const int c = (code >> (2 * st2)) & 3;
            if (c == 0) break;
            st2 = (st2 == 0 || st2 == 3) ? (c == 1 ? 2 : 3) : (c == 1 ? 0 : 1);
            kb--;
          }
          const uint32_t shift = run_n - plen;
          bool prev_es = true;
          for (uint32_t j2 = 0; j2 < plen; j2++) {
            const bool es = A.bp[i0 + run_s + shift + j2] & 0x80;
            const uint32_t qq = P0 + 3u * (run_s + j2);
            if (prev_es) set_s(qq);
            if (es) set_e(qq + 2);
            prev_es = es;
          }
        }
        run_n = 0;
      }
      if (MODE == 2) {
        // A single rune between two multi-rune pieces is a run of one: cutZh emits it as it is (T:246-249), no Viterbi.
        // Only runs of two or more are marked for k_runs -- and whatever touches the segment's edges, where the
        // neighbour is another lane's: the segment's first pieces while they are single (in_run starts true), and a
        // single rune still pending at its end.
        const uint32_t q = P0 + 3u * k;
        if (d == 1u) {
          if (in_run) {
            ma.set(q);
          } else if (pend != 0xFFFFFFFFu) {
            ma.set(pend);
            ma.set(q);
            pend = 0xFFFFFFFFu;
            in_run = true;
          } else {
            pend = q;
          }
        } else {
          if (pend != 0xFFFFFFFFu) {
            set_s(pend);
            set_e(pend + 2u);
            pend = 0xFFFFFFFFu;
          }
          in_run = false;
          set_s(q);
          set_e(q + 3u * d - 1u);
        }
      } else if (!single) {
        set_s(P0 + 3u * k);
        set_e(P0 + 3u * (k + d) - 1u);
      }
      k += d;
      if (k >= npos) {
        if (MODE == 2 && pend != 0xFFFFFFFFu) ma.set(pend);
        if (!HMM) {
          sa.flush();
          ea.flush();
          if (MODE == 2) ma.flush();
        }
        active = false;
      }
    }
    __syncwarp();
  }
}

// ==========================================================================================
// k_runs: cutZh's HMM half (T:228-253) with one lane per RUN of single-rune pieces instead of one lane per block.
// Used for the long blocks that k_emit cuts into segments (a run may cross a segment boundary).
// Runs are independent of each other (viterbi starts afresh for each, T:238,246) and k_emit<2> has marked their runes in
// m_bits, so they are found position-parallel: a marked rune starts a run iff the rune before it (3 bytes back, same
// document) is not marked.  A warp takes 32 words of the bitmap (1 KiB of text), collects the run starts in shared
// memory and hands them out to its lanes; a lane walks its run rune by rune -- emission row, one Viterbi step
// (stateTransitionRoute T:736-756: strict > from minFloat, list order), the four best paths carried as bit masks as in
// k_emit<1> -- until the next rune is not marked, then applies viterbi's tail (T:723-729) and cutHMM (T:273-285).
// (For short blocks k_emit<1> is faster -- 3.97 against 6.5 ms/GB on config 3 -- because few lanes of a warp are on a
// Viterbi step at the same time here; for 10k-rune blocks the order is reversed.)
// ==========================================================================================
constexpr int kRunThreads = 128;
constexpr int kRunWords = 8;                     // bitmap words per lane and chunk: a warp gathers the runs of 8 KiB of text
constexpr int kRunQueue = 32 * kRunWords * 11;   // run starts per warp and chunk: at most 11 three-byte runes start in a 32-byte word

__global__ void __launch_bounds__(kRunThreads) k_runs(const JbTables T, const EmitArgs A, uint32_t n, const uint32_t* __restrict__ ds_bits) {
  constexpr uint32_t kRegRun = 24;
  __shared__ uint16_t queue[kRunThreads / 32][kRunQueue];  // byte offsets inside the chunk
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((A.counters[C_FLAGS] & 1u) || A.counters[C_N_SEG] == 0) return;  // marks only come from the segments of long blocks
  const uintptr_t tbase = reinterpret_cast<uintptr_t>(A.text);
  const uint32_t nwords = (n + 31) / 32, nchunks = (nwords + 32 * kRunWords - 1) / (32 * kRunWords);
  uint16_t* const q = queue[warp];
  auto set_s = [&](uint32_t p) { atomicOr(&A.s_bits[p >> 5], 1u << (p & 31)); };
  auto set_e = [&](uint32_t p) { atomicOr(&A.e_bits[p >> 5], 1u << (p & 31)); };
  auto marked = [&](uint32_t p) -> bool {  // a single-rune piece starts at byte p, and p does not start a document
    if (p >= n) return false;
    const uint32_t w = p >> 5, b = 1u << (p & 31);
    return (__ldg(A.m_bits + w) & b) && !(__ldg(ds_bits + w) & b);
  };
  for (uint32_t chunk = blockIdx.x * (kRunThreads / 32) + warp; chunk < nchunks; chunk += gridDim.x * (kRunThreads / 32)) {
    // a lane looks at kRunWords consecutive words
    const uint32_t w0 = (chunk * 32 + lane) * kRunWords, cbase = chunk * 32 * kRunWords * 32;
    uint32_t starts[kRunWords], cnt = 0;
    {
      uint32_t prev = (w0 && w0 < nwords) ? __ldg(A.m_bits + w0 - 1) : 0u;
#pragma unroll
      for (int j = 0; j < kRunWords; j++) {
        const uint32_t w = w0 + j;
        const uint32_t mw = w < nwords ? __ldg(A.m_bits + w) : 0u, dw = w < nwords ? __ldg(ds_bits + w) : 0u;
        // the rune 3 bytes back is marked too and no document starts here: the run goes on; else it starts here
        starts[j] = mw & ~(((mw << 3) | (prev >> 29)) & ~dw);
        cnt += __popc(starts[j]);
        prev = mw;
      }
    }
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(FULL, inc, o);
      if (lane >= o) inc += v;
    }
    const uint32_t total = __shfl_sync(FULL, inc, 31);
    {
      uint32_t o = inc - cnt;
#pragma unroll
      for (int j = 0; j < kRunWords; j++) {
        uint32_t m = starts[j];
        while (m) {
          q[o++] = (uint16_t)((lane * kRunWords + j) * 32u + (uint32_t)__ffs(m) - 1u);
          m &= m - 1;
        }
      }
    }
    __syncwarp();
    // The lanes take runs from the queue as they finish (ballot order), so that in every iteration all of them do the
    // same thing: one Viterbi step (T:688-719).  A run that ends is flushed by its lane alone.
    const uint32_t lt_mask = (1u << lane) - 1u;
    uint32_t qnext = 0, p0 = 0, p = 0, run_n = 0;
    bool active = false;
    double V[4] = {0.0, 0.0, 0.0, 0.0};
    // viterbi's fullPath (T:715-716) in bit form, per state: bits 0..23 = which runes of its best path are E or S
    // (token ends), bits 24..31 = the path's length (a route with from == "" restarts it)
    uint32_t pm[4] = {0, 0, 0, 0};
    for (;;) {
      const uint32_t idle = __ballot_sync(FULL, !active);
      if (idle) {
        const uint32_t mine = qnext + __popc(idle & lt_mask);
        if (!active && mine < total) {
          p0 = p = cbase + q[mine];
          run_n = 0;
          active = true;
        }
        qnext = min(total, qnext + (uint32_t)__popc(idle));
        if (__all_sync(FULL, !active)) break;
      }
      if (active) {
        const uintptr_t ap = tbase + p, a4 = ap & ~(uintptr_t)3;
        const uint32_t xl = __ldg(reinterpret_cast<const uint32_t*>(a4));
        const uint32_t xh = (ap & 3) >= 2 ? __ldg(reinterpret_cast<const uint32_t*>(a4 + 4)) : 0u;  // (never a word past the text)
        const uint32_t x = __funnelshift_r(xl, xh, (uint32_t)(ap & 3) * 8u);
        const uint32_t cp = ((x & 0xFu) << 12) | ((x >> 2) & 0xFC0u) | ((x >> 16) & 0x3Fu);
        const double2* ep = reinterpret_cast<const double2*>(T.emit + (size_t)cp * 4);
        const double2 e0 = __ldg(ep), e1 = __ldg(ep + 1);
        const double em[4] = {e0.x, e0.y, e1.x, e1.y};
        if (run_n == 0) {
#pragma unroll
          for (int s = 0; s < 4; s++) V[s] = T.start[s] + em[s];
          pm[0] = pm[1] = 1u << 24;
          pm[2] = pm[3] = (1u << 24) | 1u;
        } else {
          double W[4];
          uint32_t code = 0, npm[4];
          const uint32_t step = (1u << 24) | (run_n < kRegRun ? (1u << run_n) : 0u);  // one more entry; E and S end a token
#pragma unroll
          for (int s = 0; s < 4; s++) {
            const int pa = (s == 0 || s == 3) ? 2 : 0, pb = (s == 0 || s == 3) ? 3 : 1;
            const double r0 = V[pa] + T.trans[s][0], r1 = V[pb] + T.trans[s][1];
            const bool t0 = r0 > JB_MINF;   // from = 1
            const double b0 = t0 ? r0 : JB_MINF;
            const bool t1 = r1 > b0;        // from = 2
            W[s] = (t1 ? r1 : b0) + em[s];
            code |= (t1 ? 2u : (t0 ? 1u : 0u)) << (2 * s);
            // fullPath[s] = fullPath[route.from] + [s]; fullPath[""] is nil (T:715-716)
            npm[s] = (t1 ? pm[pb] : (t0 ? pm[pa] : 0u)) + (s >= 2 ? step : (1u << 24));
          }
#pragma unroll
          for (int s = 0; s < 4; s++) {
            V[s] = W[s];
            pm[s] = npm[s];
          }
          A.bp[p / 3u] = (uint8_t)code;  // only read back for runs longer than the register window
        }
        run_n++;
        if (marked(p + 3u)) {
          p += 3u;
        } else {
          // ---- viterbi's tail (T:723-729) + cutHMM (T:273-285) ----
          if (run_n == 1) {
            set_s(p0);
            set_e(p0 + 2u);
          } else if (run_n <= kRegRun) {
            const uint32_t pf = V[2] > V[3] ? pm[2] : pm[3];  // T:723-729
            const uint32_t plen = pf >> 24;
            // path[j] applies to rune j (T:277-283): a short path drops the run's tail
            const uint32_t lm = (1u << plen) - 1u;
            const uint32_t es = ((pf & 0xFFFFFFu) >> (run_n - plen)) & lm;
            const uint32_t startm = ((es << 1) | 1u) & lm;
            or_span(A.s_bits, p0, spread3(startm), spread3(startm >> 16));
            or_span(A.e_bits, p0 + 2u, spread3(es), spread3(es >> 16));
          } else {
            const uint32_t i0 = p0 / 3u;
            int st2 = V[2] > V[3] ? 2 : 3;
            uint32_t kb = run_n - 1, plen = 0;
            for (;;) {  // back-trace; stops early where route.from == "" (T:715-716)
              const uint8_t code = A.bp[i0 + kb];
              A.bp[i0 + kb] = (uint8_t)(st2 >= 2 ? 0x80 : 0);  // the state of this path entry is E or S
              plen++;
              if (kb == 0) break;
              const int c = (code >> (2 * st2)) & 3;
              if (c == 0) break;
              st2 = (st2 == 0 || st2 == 3) ? (c == 1 ? 2 : 3) : (c == 1 ? 0 : 1);
              kb--;
            }
            const uint32_t shift = run_n - plen;
            bool prev_es = true;
            for (uint32_t j2 = 0; j2 < plen; j2++) {
              const bool es = A.bp[i0 + shift + j2] & 0x80;
              const uint32_t qq = p0 + 3u * j2;
              if (prev_es) set_s(qq);
              if (es) set_e(qq + 2u);
              prev_es = es;
            }
          }
          active = false;
        }
      }
      __syncwarp();
    }
    __syncwarp();
  }
}

// ==========================================================================================
// k_land: one lane per segment.  land(t) = where the walk k -> k + d(k) entered at rune t of the segment first steps
// at or past the segment's end, as an offset 0..15 into the next segment: land(t) = land(t + d(t)), with land(S + i) = i.
// Right to left with the sixteen values that follow t kept as nibbles of one 64-bit word W (nibble i = land(t + 1 + i)):
// land(t) is nibble d(t) - 1 of W, and W shifts by one nibble.  After rune 0, W holds land(0..15): all a segment's
// possible entries (a word is at most 16 runes on this path: PB == 4).
// ==========================================================================================
__global__ void __launch_bounds__(256) k_land(const EmitArgs A) {
  if (A.counters[C_FLAGS] & 1u) return;
  const uint32_t nseg = min(A.counters[C_N_SEG], A.segs_cap);
  for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < nseg; s += gridDim.x * blockDim.x) {
    const uint2 desc = A.segs[s];
    unsigned long long W = 0xFEDCBA9876543210ull;
    if (desc.y == kSegRunes) {  // (a shorter segment is its block's last: nothing to enter after it)
      const uint32_t i0 = desc.x / 3u;
      uint32_t i = i0 + kSegRunes - 1u, pw = __ldg(A.path + (i >> 3));
      for (;;) {
        const uint32_t d1 = (pw >> ((i & 7u) * 4u)) & 15u;
        W = (W << 4) | ((W >> (4u * d1)) & 15ull);
        if (i == i0) break;
        i--;
        if ((i & 7u) == 7u) pw = __ldg(A.path + (i >> 3));
      }
    }
    A.land[s] = W;
  }
}

// k_chain: one warp per long block threads the entry through its segments: entry(0) = 0, entry(j + 1) = nibble entry(j)
// of land[j].  32 segments' words are loaded at a time; every lane follows the chain over them with shuffles.  A
// segment that is entered at rune e > 0 begins there.
__global__ void __launch_bounds__(128) k_chain(const EmitArgs A) {
  if (A.counters[C_FLAGS] & 1u) return;
  const uint32_t nlong = min(A.counters[C_N_LONG], A.longs_cap);
  const uint32_t lane = threadIdx.x & 31u, nwarps = gridDim.x * (blockDim.x / 32u);
  for (uint32_t li = blockIdx.x * (blockDim.x / 32u) + (threadIdx.x >> 5); li < nlong; li += nwarps) {
    const uint2 lg = A.longs[li];
    uint32_t e = 0;
    for (uint32_t base = 0; base < lg.y; base += 32u) {
      const uint32_t idx = base + lane, cnt = min(32u, lg.y - base);
      const unsigned long long Wl = idx < lg.y ? A.land[lg.x + idx] : 0ull;
      uint32_t my = 0;
      for (uint32_t i = 0; i < cnt; i++) {
        const unsigned long long Wi = __shfl_sync(FULL, Wl, (int)i);
        if (lane == i) my = e;
        e = (uint32_t)(Wi >> (4u * e)) & 15u;
      }
      if (idx < lg.y && my) {
        uint2 d = A.segs[lg.x + idx];
        d.x += 3u * my;
        d.y = d.y > my ? d.y - my : 0u;
        A.segs[lg.x + idx] = d;
      }
    }
  }
}

int launch_emit(const JbTables& T, const EmitArgs& A0, bool hmm, int num_sms, cudaStream_t st, uint32_t n, const uint32_t* ds_bits) {
  const bool r16 = T.max_delta <= 16;
  EmitArgs A = A0;
  A.count_idx = C_N_BLK;
  A.cursor_idx = C_CUR_EMIT;
  // segments need d <= 16 (4-bit path entries) and a text that can hold a long block at all
  const bool split = r16 && A.segs && A.land && A.longs && n >= 3u * kLongRunes;
  if (!split) A.segs = nullptr;
  // one resident wave: CTAs per SM by register count
  static int occ[2] = {0, 0};
  if (!occ[hmm]) {
    int b = 0;
    if (hmm) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_emit<1, 4>, kEmThreads, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_emit<0, 4>, kEmThreads, 0);
    occ[hmm] = b > 0 ? b : 8;
  }
  const unsigned grid = (unsigned)num_sms * (unsigned)occ[hmm];
  if (r16) {
    if (hmm) k_emit<1, 4><<<grid, kEmThreads, 0, st>>>(T, A);
    else k_emit<0, 4><<<grid, kEmThreads, 0, st>>>(T, A);
  } else {
    if (hmm) k_emit<1, 8><<<grid, kEmThreads, 0, st>>>(T, A);
    else k_emit<0, 8><<<grid, kEmThreads, 0, st>>>(T, A);
  }
  int launches = 1;
  if (split) {  // each of these returns at once when k_emit found no long block
    k_land<<<(unsigned)num_sms * 8u, 256, 0, st>>>(A);
    k_chain<<<(unsigned)num_sms * 8u, 128, 0, st>>>(A);
    EmitArgs S = A;
    S.blocks = A.segs;
    S.blocks_cap = A.segs_cap;
    S.count_idx = C_N_SEG;
    S.cursor_idx = C_CUR_SEG;
    S.segs = nullptr;
    S.min_chunk = 32;
    if (hmm) {
      k_emit<2, 4><<<(unsigned)num_sms * 16u, kEmThreads, 0, st>>>(T, S);
      k_runs<<<(unsigned)num_sms * 8u, kRunThreads, 0, st>>>(T, S, n, ds_bits);
      launches += 4;
    } else {
      k_emit<0, 4><<<(unsigned)num_sms * 16u, kEmThreads, 0, st>>>(T, S);
      launches += 3;
    }
  }
  return cudaGetLastError() == cudaSuccess ? launches : -1;
}


// ==========================================================================================
// k_wide: Han blocks that contain a 4-byte rune (CJK extension B and beyond).  k_route steps over 3-byte runes
// only; it hands such a block over the moment it lands inside a 4-byte rune (k_scan does so for a block that
// ENDS with one).  They are rare (about one rune in 1e5 in real text), so one lane takes a whole block and
// restates the reference directly: buildDag (T:462-497), calcDagProba (T:502-548), maxIndexProba (T:565-578),
// findDagPath (T:552-562), cutZh / viterbi / cutHMM (T:221-285, 668-756), with per-rune scratch in HBM
// indexed by lead byte / 3.
// ==========================================================================================
struct WideCtx {
  const uint8_t* text;
  const uint32_t* ds_bits;
  uint32_t n;
  __device__ __forceinline__ bool ds_at(uint32_t p) const { return p >= n || ((ds_bits[p >> 5] >> (p & 31)) & 1); }
  __device__ __forceinline__ uint32_t len_at(uint32_t p) const { return text[p] >= 0xF0 ? 4u : 3u; }  // inside a block
  __device__ __forceinline__ uint32_t rune_at(uint32_t p) const {
    const uint8_t* b = text + p;
    if (b[0] >= 0xF0) return ((b[0] & 0x07u) << 18) | ((b[1] & 0x3Fu) << 12) | ((b[2] & 0x3Fu) << 6) | (b[3] & 0x3Fu);
    return ((b[0] & 0x0Fu) << 12) | ((b[1] & 0x3Fu) << 6) | (b[2] & 0x3Fu);
  }
  __device__ __forceinline__ uint32_t prev_lead(uint32_t p) const { return (text[p - 3] & 0xF0) == 0xE0 ? p - 3 : p - 4; }  // inside a block
};
// lead byte of a Han rune that ends right before p in the same document, or 0xFFFFFFFF
__device__ uint32_t w_han_before(const WideCtx& cx, const JbTables& T, uint32_t p) {
  const uint8_t* t = cx.text;
  if (p == 0 || cx.ds_at(p)) return 0xFFFFFFFFu;
  if (p >= 3 && (t[p - 3] & 0xF0) == 0xE0) {
    const uint32_t q = p - 3, L = t[q], c1 = t[q + 1], c2 = t[q + 2];
    if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80 || (L == 0xE0 && c1 < 0xA0) || (L == 0xED && c1 > 0x9F)) return 0xFFFFFFFFu;
    if (cx.ds_at(q + 1) || cx.ds_at(q + 2)) return 0xFFFFFFFFu;
    return s_is_han(((L & 0xFu) << 12) | ((c1 & 0x3Fu) << 6) | (c2 & 0x3Fu), T) ? q : 0xFFFFFFFFu;
  }
  if (p >= 4 && t[p - 4] >= 0xF0 && t[p - 4] <= 0xF4) {
    const uint32_t q = p - 4, L = t[q], c1 = t[q + 1], c2 = t[q + 2], c3 = t[q + 3];
    if ((c1 & 0xC0) != 0x80 || (c2 & 0xC0) != 0x80 || (c3 & 0xC0) != 0x80 || (L == 0xF0 && c1 < 0x90) || (L == 0xF4 && c1 > 0x8F)) return 0xFFFFFFFFu;
    if (cx.ds_at(q + 1) || cx.ds_at(q + 2) || cx.ds_at(q + 3)) return 0xFFFFFFFFu;
    return s_is_han(((L & 0x7u) << 18) | ((c1 & 0x3Fu) << 12) | ((c2 & 0x3Fu) << 6) | (c3 & 0x3Fu), T) ? q : 0xFFFFFFFFu;
  }
  return 0xFFFFFFFFu;
}
__device__ __forceinline__ void w_set(uint32_t* bits, uint32_t p) { atomicOr(&bits[p >> 5], 1u << (p & 31)); }
__device__ void w_load_emit(const JbTables& T, uint32_t cp, double e[4]) {
  if (cp < 0x10000) {
    const double2* p = reinterpret_cast<const double2*>(T.emit + (size_t)cp * 4);
    const double2 a = __ldg(p), b = __ldg(p + 1);
    e[0] = a.x, e[1] = a.y, e[2] = b.x, e[3] = b.y;
    return;
  }
  e[0] = e[1] = e[2] = e[3] = JB_MINF;  // missing emission (T:690-692)
  int lo = 0, hi = (int)T.n_emit_supp - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const uint32_t r = __ldg(T.emit_supp_rune + mid);
    if (r == cp) {
      for (int s = 0; s < 4; s++) e[s] = __ldg(T.emit_supp + (size_t)mid * 4 + s);
      return;
    }
    if (r < cp) lo = mid + 1;
    else hi = mid - 1;
  }
}

template <bool HMM>
__global__ void __launch_bounds__(128) k_wide(const JbTables T, const WideArgs A) {
  if (A.counters[C_FLAGS] & 1u) return;
  const uint32_t nw = min(A.counters[C_N_WIDE], A.wide_cap);
  const WideCtx cx{A.text, A.ds_bits, A.n};
  for (uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x; wi < nw; wi += gridDim.x * blockDim.x) {
    const uint32_t e = A.wide_list[wi];  // lead byte of the block's last rune
    const uint32_t blk_end = e + cx.len_at(e);
    uint32_t start = e;
    for (uint32_t q; (q = w_han_before(cx, T, start)) != 0xFFFFFFFFu;) start = q;
    // ---- route DP, right to left ----
    for (uint32_t q = e;; q = cx.prev_lead(q)) {
      const uint32_t r0 = cx.rune_at(q);
      uint32_t pos = q + cx.len_at(q);
      double w0;
      uint32_t info, child, parent, hs = r0 < 0x10000 ? JB_PARENT_FIRST(r0) : JB_PARENT_ROOT;
      if (r0 < 0x10000) {  // termFreq[string(iRune)] (T:468-472)
        const uint4 f = __ldg(reinterpret_cast<const uint4*>(T.first + r0));
        w0 = __longlong_as_double(((long long)f.y << 32) | (long long)f.x);
        info = f.z;
        child = f.w;
        parent = JB_PARENT_FIRST(r0);
      } else {
        double pw;
        uint32_t prb;
        const int ps = jb_probe_edge(T.entries, T.hash_mask, T.hash_shift, hs, JB_PARENT_ROOT, r0, &pw, &prb);
        if (ps >= 0 && jb_w_positive(pw)) {
          w0 = pw;
          info = (uint32_t)JB_MAX_DELTA << 8;
        } else {
          w0 = ps >= 0 ? pw : T.neg_log_total;  // freq 0: -Inf; missing: log(1) - total
          info = JB_FIRST_GATE;
        }
        child = ps >= 0 ? (prb >> 21) : 0u;
        parent = (uint32_t)ps;
      }
      double prev = JB_MINF, best_v = 0.0, v = w0 + (pos == blk_end ? 0.0 : A.R[pos / 3u]);
      uint32_t best_b = 0, last_b = pos - q;
      if (v >= prev) {  // maxIndexProba (T:565-578)
        best_b = last_b;
        best_v = v;
      }
      prev = v;
      if (!(info & JB_FIRST_GATE)) {
        const uint32_t maxlen = (info >> 8) & 0xFFu;
        for (uint32_t L = 1; L < maxlen && pos < blk_end;) {  // for j := range textRunes[i:] (T:473-482)
          const uint32_t rl = cx.rune_at(pos);
          const bool may = (L == 1 && r0 < 0x10000) ? ((child >> jb_bloom_bit(rl)) & 1) : ((child >> jb_bloom11(rl)) & 1);
          if (!may) break;
          double pw;
          uint32_t prb;
          const int ps = jb_probe_edge(T.entries, T.hash_mask, T.hash_shift, hs, parent, rl, &pw, &prb);
          if (ps < 0) break;
          L++;
          pos += cx.len_at(pos);
          if (jb_w_positive(pw)) {
            v = pw + (pos == blk_end ? 0.0 : A.R[pos / 3u]);
            last_b = pos - q;
            if (v >= prev) {
              best_b = last_b;
              best_v = v;
            }
            prev = v;
          }
          parent = (uint32_t)ps;
          child = prb >> 21;
        }
      }
      if (best_b == 0) {  // best.index == -1 -> return prev (T:574-576)
        best_b = last_b;
        best_v = prev;
      }
      A.R[q / 3u] = best_v;
      A.len8[q / 3u] = (uint8_t)best_b;
      if (q == start) break;
    }
    // ---- forward walk, HMM over runs of single-rune pieces ----
    uint32_t run_n = 0, run_s = 0, run_l = 0;  // runes, lead byte of the first / last rune of the run
    double V[4] = {0.0, 0.0, 0.0, 0.0};
    for (uint32_t p = start; p < blk_end;) {
      const uint32_t bl = A.len8[p / 3u], rlen = cx.len_at(p);
      const bool single = HMM && bl == rlen;
      if (single) {
        double em[4];
        w_load_emit(T, cx.rune_at(p), em);
        if (run_n == 0) {
          run_s = p;
          for (int s = 0; s < 4; s++) V[s] = T.start[s] + em[s];
        } else {
          double W[4];
          uint32_t code = 0;
          for (int s = 0; s < 4; s++) {  // stateTransitionRoute (T:736-756)
            const int pa = (s == 0 || s == 3) ? 2 : 0, pb = (s == 0 || s == 3) ? 3 : 1;
            const double r0v = V[pa] + T.trans[s][0], r1v = V[pb] + T.trans[s][1];
            double best = JB_MINF;
            uint32_t from = 0;
            if (r0v > best) {
              best = r0v;
              from = 1;
            }
            if (r1v > best) {
              best = r1v;
              from = 2;
            }
            W[s] = best + em[s];
            code |= from << (2 * s);
          }
          for (int s = 0; s < 4; s++) V[s] = W[s];
          A.code8[p / 3u] = (uint8_t)code;
        }
        run_l = p;
        run_n++;
      }
      if (HMM && run_n && (!single || p + bl >= blk_end)) {  // viterbi's tail (T:723-729) + cutHMM (T:273-285)
        if (run_n == 1) {
          w_set(A.s_bits, run_s);
          w_set(A.e_bits, run_s + cx.len_at(run_s) - 1u);
        } else {
          int st2 = V[2] > V[3] ? 2 : 3;
          uint32_t kb = run_l, plen = 0;
          for (;;) {  // back-trace; stops early where route.from == "" (T:715-716)
            const uint8_t code = A.code8[kb / 3u];
            A.code8[kb / 3u] = (uint8_t)(st2 >= 2 ? 0x80 : 0);
            plen++;
            if (kb == run_s) break;
            const int c = (code >> (2 * st2)) & 3;
            if (c == 0) break;
            st2 = (st2 == 0 || st2 == 3) ? (c == 1 ? 2 : 3) : (c == 1 ? 0 : 1);
            kb = cx.prev_lead(kb);
          }
          // path[j] applies to rune j (T:277-283): entry j sits at rune j + (run_n - plen); kb is entry 0
          uint32_t pr = run_s, pe = kb;
          bool prev_es = true;
          for (uint32_t j = 0; j < plen; j++) {
            const bool es = A.code8[pe / 3u] & 0x80;
            if (prev_es) w_set(A.s_bits, pr);
            if (es) w_set(A.e_bits, pr + cx.len_at(pr) - 1u);
            prev_es = es;
            pr += cx.len_at(pr);
            pe += cx.len_at(pe);
          }
        }
        run_n = 0;
      }
      if (!single) {
        w_set(A.s_bits, p);
        w_set(A.e_bits, p + bl - 1u);
      }
      p += bl;
    }
  }
}

int launch_wide(const JbTables& T, const WideArgs& A, bool hmm, int num_sms, cudaStream_t st) {
  if (hmm) k_wide<true><<<(unsigned)num_sms, 128, 0, st>>>(T, A);
  else k_wide<false><<<(unsigned)num_sms, 128, 0, st>>>(T, A);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace jb
