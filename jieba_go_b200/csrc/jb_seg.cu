// k_seg -- the segment kernel of the streaming path: buildDag + calcDagProba + maxIndexProba + findDagPath +
// cutZh / viterbi / cutHMM for a few dozen Han blocks at a time per CTA, entirely in shared memory.
//
// The reference (T = /root/reference/tokenizer.go) walks one block at a time: probes of termFreq for every
// prefix at every rune (buildDag T:462-497), then a right-to-left pass that adds log-probabilities and picks
// with maxIndexProba (calcDagProba T:502-548, T:565-578), then a forward walk (findDagPath T:552-562) and, with
// HMM, Viterbi over the runs of single runes (cutZh T:221-255, viterbi T:668-730, cutHMM T:273-285).
// Only the second and third steps are sequential, and only along one block.  So a CTA takes G blocks from k_scan's
// list (about 2,000 runes) and runs, separated by barriers:
//
//   decode   warp per block: UTF-8 -> u16 runes in shared memory, blocks laid out back to back with one SENTINEL
//            position after each (rune 0xFFFF: in no key, so every prefix chain ends there -- T:473's range over
//            textRunes[i:] -- and its route value is the 0.0 of T:522's {j, 0.0})
//   pass 1   thread per POSITION (all lanes busy, four positions in flight per thread): the first-rune table entry
//            (T:468-472: weight of edge (i,i+1), gate, Bloom of second runes), then -- gate and Bloom permitting -- the
//            2-rune key's hash entry.  A candidate's weight goes to a pool in shared memory, chained per position in
//            ascending length; a position whose prefix chain goes on (longer keys, or a displaced hash entry)
//            leaves a 16-byte task
//   pass 2   thread per task: the rest of the chain, one dependent probe per step (T:473-482)
//   route    lane per BLOCK, right to left, from shared memory only: pieceFreq + next.proba in the reference's
//            order (T:519, 529) folded into maxIndexProba's (prev, best) pair (T:565-578); the selected value
//            overwrites the position's single-rune weight (it is R[i] for the positions to its left)
//   emit     the same lane walks its block left to right (T:552-562) and sets the token bits; with HMM one Viterbi
//            step per single-rune piece with the four best paths carried as bit masks (see k_emit)
//
// Nothing per position ever reaches HBM: text is read once, token bits are OR-ed into the two bitmaps.
// Left to k_route / k_emit (the lane-per-block kernels, any length): blocks longer than kSgMaxRunes and the blocks of
// a round that ran out of pool / position space; to k_wide: blocks with a 4-byte rune.
#include "jb_stream.cuh"

#include "../../include/jieba_b200.h"

namespace jb {

#define FULL 0xFFFFFFFFu

constexpr int kSgThreads = 512;
constexpr int kSgWarps = kSgThreads / 32;
constexpr int kSgCap = 4096;    // positions per round (runes + one sentinel per block)
constexpr int kSgPool = 2560;   // candidates beyond the single rune per round
constexpr int kSgTasks = 1536;  // prefix chains that go on after pass 1
constexpr int kSgTPT = kSgTasks / kSgThreads;  // tasks per thread per level
constexpr int kSgNB = 128;      // blocks per round at most (one lane each in route / emit)
constexpr int kSgU = 2;         // positions in flight per thread in pass 1
constexpr uint32_t kSgSent = 0xFFFFu;
constexpr uint32_t kSgNone = 0x1FFFu;  // "no pool entry" (13 bits in a task, 16 in head / nx)
static_assert(kSgPool < (int)kSgNone && kSgCap <= 4096 && kSgTasks % kSgThreads == 0, "task encoding");

enum { SG_NB = 0, SG_P = 1, SG_DONE = 2, SG_NPOOL = 3, SG_OVF = 4, SG_NT0 = 5, SG_NT1 = 6 };

template <typename MT>
struct SegLayout {
  static constexpr size_t o_W1 = 0;
  static constexpr size_t o_WX = o_W1 + sizeof(double) * kSgCap;
  static constexpr size_t o_task = o_WX + sizeof(double) * kSgPool;
  static constexpr size_t o_mask = o_task + sizeof(uint4) * kSgTasks;
  static constexpr size_t o_rune = o_mask + sizeof(MT) * kSgCap;
  static constexpr size_t o_head = o_rune + 2 * (kSgCap + 32);
  static constexpr size_t o_nx = o_head + 2 * kSgCap;
  static constexpr size_t o_fb = (o_nx + 2 * kSgPool + 15) & ~(size_t)15;
  static constexpr size_t o_off = o_fb + 4 * kSgNB;
  static constexpr size_t o_nr = o_off + 2 * kSgNB;
  static constexpr size_t o_hist = o_nr + 2 * kSgNB;     // 256 x u32: counting sort of the blocks by length
  static constexpr size_t o_order = o_hist + 4 * 256;    // block of route lane t
  static constexpr size_t o_flag = o_order + kSgNB;
  static constexpr size_t o_ctl = (o_flag + kSgNB + 15) & ~(size_t)15;
  static constexpr size_t bytes = o_ctl + 64;
};

template <bool HMM, typename MT>
__global__ void __launch_bounds__(kSgThreads, 2) k_seg(const JbTables T, const SegArgs A) {
  extern __shared__ __align__(16) uint8_t sg_smem[];
  using LY = SegLayout<MT>;
  double* const W1 = reinterpret_cast<double*>(sg_smem + LY::o_W1);   // weight of edge (i,i+1), then R[i]
  double* const WX = reinterpret_cast<double*>(sg_smem + LY::o_WX);   // weights of the longer candidates
  uint4* const task = reinterpret_cast<uint4*>(sg_smem + LY::o_task);
  MT* const mask = reinterpret_cast<MT*>(sg_smem + LY::o_mask);       // bit L-1: candidate of L runes; then the chosen length
  uint16_t* const rune = reinterpret_cast<uint16_t*>(sg_smem + LY::o_rune);
  uint16_t* const head = reinterpret_cast<uint16_t*>(sg_smem + LY::o_head);  // first pool entry of the position; then Viterbi back-pointers
  uint16_t* const nx = reinterpret_cast<uint16_t*>(sg_smem + LY::o_nx);      // next pool entry of the same position
  uint32_t* const blk_fb = reinterpret_cast<uint32_t*>(sg_smem + LY::o_fb);  // first byte of the block
  uint16_t* const blk_off = reinterpret_cast<uint16_t*>(sg_smem + LY::o_off);
  uint16_t* const blk_nr = reinterpret_cast<uint16_t*>(sg_smem + LY::o_nr);
  uint32_t* const hist = reinterpret_cast<uint32_t*>(sg_smem + LY::o_hist);
  uint8_t* const order = sg_smem + LY::o_order;
  uint8_t* const blk_flag = sg_smem + LY::o_flag;
  uint32_t* const ctl = reinterpret_cast<uint32_t*>(sg_smem + LY::o_ctl);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lt_mask = (1u << lane) - 1u;
  if (A.counters[C_FLAGS] & 1u) return;  // the general pipeline redoes this batch
  const uint32_t nblocks = min(A.counters[C_N_BLK], A.blocks_cap);
  const uint4* __restrict__ first = reinterpret_cast<const uint4*>(T.first);
  const uint4* __restrict__ entries = reinterpret_cast<const uint4*>(T.entries);
  const uint32_t hmask = T.hash_mask;
  // blocks per round: fill about 5/6 of the position space going by the mean block length (an upper bound: all
  // text taken as Han); few blocks -> smaller rounds so that every CTA gets some
  uint32_t G;
  {
    const uint32_t avg = nblocks ? A.n / 3u / nblocks + 2u : 2u;
    G = (uint32_t)(kSgCap * 5 / 6) / avg;
    G = min(G, nblocks / gridDim.x + 1u);
    G = max(8u, min(G, (uint32_t)kSgNB));
  }

  auto push_wide = [&](uint32_t lastb) {
    const uint32_t wi = atomicAdd(&A.counters[C_N_WIDE], 1u);
    if (wi < A.wide_cap) A.wide_list[wi] = lastb;
    else atomicOr(&A.counters[C_FLAGS], 1u);
  };
  auto push_long = [&](uint32_t lastb, uint32_t nr) {
    const uint32_t li = atomicAdd(&A.counters[C_N_LONG], 1u);
    if (li < A.long_cap) A.long_blocks[li] = make_uint2(lastb, nr);
    else atomicOr(&A.counters[C_FLAGS], 1u);
  };
  // one more candidate (length L, weight w) of position pos, after pool entry prev of the same position
  auto add_cand = [&](uint32_t pos, uint32_t L, double w, uint32_t prev) -> uint32_t {
    const uint32_t idx = atomicAdd(&ctl[SG_NPOOL], 1u);
    if (idx >= (uint32_t)kSgPool) {
      ctl[SG_OVF] = 1u;
      return prev;
    }
    WX[idx] = w;
    nx[idx] = (uint16_t)kSgNone;
    mask[pos] = (MT)(mask[pos] | (MT)(1u << (L - 1u)));
    if (prev == kSgNone) head[pos] = (uint16_t)idx;
    else nx[prev] = (uint16_t)idx;
    return idx;
  };
  // One step of a position's prefix chain (T:473-482).  The task: `parent` (z) identifies the matched prefix of L
  // runes, `slot` (x) is the hash slot to look at for the key of L + 1 runes (hash state y), w = pos | prev << 12 |
  // L << 25 | home << 30 (home: x is that key's home slot).  e = entries[slot].  Returns true and rewrites the task when the chain goes on.
  auto chain_step = [&](uint4& tk, const uint4 e) -> bool {
    const uint32_t pos = tk.w & 0xFFFu, L = (tk.w >> 25) & 31u;
    uint32_t prev = (tk.w >> 12) & 0x1FFFu;
    const uint32_t rn = rune[pos + L];
    if (e.z == JB_PARENT_EMPTY) return false;  // _, found := termFreq[frag]; !found -> break (T:476-478)
    if (e.z == tk.z && JB_RB_RUNE(e.w) == rn) {
      const double w = __longlong_as_double(((long long)e.y << 32) | (long long)e.x);
      if (jb_w_positive(w)) prev = add_cand(pos, L + 1u, w, prev);  // val > 0 -> edge (T:479-481)
      const uint32_t rnext = rune[pos + L + 1u];
      if (!(((e.w >> 21) >> jb_bloom11(rnext)) & 1u)) return false;  // no key extends this one by rnext
      tk.y = jb_hash_next(tk.y, rnext);
      tk.z = tk.x;
      tk.x = tk.y & hmask;
      tk.w = pos | (prev << 12) | ((L + 1u) << 25) | (1u << 30);
      return true;
    }
    if (((tk.w >> 30) & 1u) && !(e.w & JB_RB_CONT)) return false;  // nothing was displaced from this home slot: the key is absent
    tk.x = (tk.x + 1u) & hmask;
    tk.w &= ~(1u << 30);
    return true;
  };
  auto run_chain = [&](uint4 tk) {
    while (chain_step(tk, __ldg(entries + tk.x))) {
    }
  };

  for (;;) {
    // ---- take G blocks from k_scan's list; lay them out back to back ------------------------------
    if (tid < 256) hist[tid] = 0;
    if (warp == 0) {
      uint32_t b0 = 0;
      if (lane == 0) b0 = atomicAdd(&A.counters[C_CUR_SEG], G);
      b0 = __shfl_sync(FULL, b0, 0);
      uint32_t run_off = 0, run_nb = 0;
      bool closed = false;  // a block did not fit: the rest of this grab goes to the long list too (keeps the layout dense)
      for (uint32_t c = 0; c < G && b0 + c < nblocks; c += 32) {
        const uint32_t bi = b0 + c + lane;
        bool valid = (c + lane < G) && bi < nblocks;
        uint32_t lastb = 0, nr = 0;
        if (valid) {
          const uint2 bd = A.blocks[bi];
          lastb = bd.x;
          nr = bd.y;
          if (nr == kWideBlock) {  // ends with a 4-byte rune
            push_wide(lastb);
            valid = false;
          } else {
            if (nr == 0) {  // the block began in an earlier k_scan tile: the nearest tile with a block start holds it
              uint32_t t = lastb / (uint32_t)kScTileBytes, sp;
              do sp = __ldg(A.tile_last_hs + --t);
              while (sp == 0xFFFFFFFFu);
              nr = (lastb - sp) / 3u + 1u;
            }
            if (nr > A.max_runes) {
              push_long(lastb, nr);
              valid = false;
            }
          }
        }
        const uint32_t sz = valid ? nr + 1u : 0u;
        uint32_t inc = sz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t v = __shfl_up_sync(FULL, inc, o);
          if (lane >= o) inc += v;
        }
        const uint32_t off = run_off + inc - sz;
        const uint32_t nofit = __ballot_sync(FULL, valid && off + sz > (uint32_t)kSgCap);
        const uint32_t firstbad = (closed ? 0u : (nofit ? (uint32_t)__ffs(nofit) - 1u : 32u));
        const bool take = valid && (uint32_t)lane < firstbad;
        if (valid && !take) push_long(lastb, nr);
        const uint32_t tm = __ballot_sync(FULL, take);
        if (take) {
          const uint32_t idx = run_nb + __popc(tm & lt_mask);
          blk_fb[idx] = lastb - 3u * (nr - 1u);
          blk_off[idx] = (uint16_t)off;
          blk_nr[idx] = (uint16_t)nr;
        }
        // totals of the blocks taken
        const uint32_t tot = __shfl_sync(FULL, inc, 31);
        uint32_t taken_sz = tot;
        if (firstbad < 32u) taken_sz = __shfl_sync(FULL, inc - sz, firstbad & 31u);
        if (closed) taken_sz = 0;
        run_off += taken_sz;
        run_nb += __popc(tm);
        if (nofit) closed = true;
      }
      if (lane == 0) {
        ctl[SG_NB] = run_nb;
        ctl[SG_P] = run_off;
        ctl[SG_DONE] = b0 >= nblocks ? 1u : 0u;
        ctl[SG_NPOOL] = 0;
        ctl[SG_OVF] = 0;
        ctl[SG_NT0] = 0;
        ctl[SG_NT1] = 0;
      }
    }
    __syncthreads();
    const uint32_t nb = ctl[SG_NB], P = ctl[SG_P];
    if (ctl[SG_DONE]) break;

    // ---- decode: UTF-8 -> runes (inside a Han block every rune has 3 bytes -- or 4: k_wide) ----------------
    // (and the first half of a counting sort of the blocks by length, longest first: route lanes of a warp then
    // finish at about the same time)
    uint32_t my_bucket = 0, my_rank = 0;
    if ((uint32_t)tid < nb) {
      my_bucket = 255u - min((uint32_t)blk_nr[tid] >> 2, 255u);
      my_rank = atomicAdd(&hist[my_bucket], 1u);
    }
    for (uint32_t b = warp; b < nb; b += kSgWarps) {
      const uint32_t fb = blk_fb[b], o = blk_off[b], nr = blk_nr[b];
      bool bad = false;
      for (uint32_t j = lane; j <= nr; j += 32) {
        if (j == nr) {
          rune[o + j] = (uint16_t)kSgSent;
          W1[o + j] = 0.0;  // {j, 0.0} at the end of the block (T:522)
        } else {
          const uint8_t* p = A.text + fb + 3u * j;
          const uint32_t c0 = __ldg(p), c1 = __ldg(p + 1), c2 = __ldg(p + 2);
          rune[o + j] = (uint16_t)(((c0 & 0xFu) << 12) | ((c1 & 0x3Fu) << 6) | (c2 & 0x3Fu));
          bad |= (c0 & 0xF0u) != 0xE0u;
        }
      }
      bad = __any_sync(FULL, bad);
      if (lane == 0) blk_flag[b] = bad ? 1 : 0;
    }
    if (tid < 32) rune[P + tid] = (uint16_t)kSgSent;
    __syncthreads();

    if (warp == 0) {  // second half of the counting sort: bucket counts -> bases
      uint32_t c[8], sum = 0;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        c[k] = hist[lane * 8 + k];
        sum += c[k];
      }
      uint32_t inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(FULL, inc, o);
        if (lane >= o) inc += v;
      }
      uint32_t run = inc - sum;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        hist[lane * 8 + k] = run;
        run += c[k];
      }
    }

    // ---- pass 1: first-rune entry and the 2-rune key of every position ---------------------------------
    for (uint32_t base = 0; base < P; base += kSgThreads * kSgU) {
      uint32_t r0[kSgU], r1[kSgU], h2[kSgU];
      uint4 f[kSgU], e2[kSgU];
      bool go2[kSgU];
#pragma unroll
      for (int u = 0; u < kSgU; u++) {
        const uint32_t i = base + u * kSgThreads + tid;
        r0[u] = i < P ? rune[i] : kSgSent;
        r1[u] = i < P ? rune[i + 1] : kSgSent;
        f[u] = make_uint4(0u, 0u, JB_FIRST_GATE, 0u);
        if (r0[u] != kSgSent) f[u] = ldg_keep(first + r0[u]);  // termFreq[string(iRune)] (T:468)
      }
#pragma unroll
      for (int u = 0; u < kSgU; u++) {
        // missing or freq 0 -> only edge (i,i+1) (T:469-472); the Bloom says whether any 2-rune key starts r0 r1
        go2[u] = r0[u] != kSgSent && !(f[u].z & JB_FIRST_GATE) && ((f[u].w >> jb_bloom_bit(r1[u])) & 1u);
        h2[u] = jb_hash_next(JB_PARENT_FIRST(r0[u]), r1[u]);
        e2[u] = make_uint4(0u, 0u, JB_PARENT_EMPTY, 0u);
        if (go2[u]) e2[u] = __ldg(entries + (h2[u] & hmask));
      }
#pragma unroll
      for (int u = 0; u < kSgU; u++) {
        const uint32_t i = base + u * kSgThreads + tid;
        const bool live = r0[u] != kSgSent;
        const bool m = go2[u] && e2[u].z == JB_PARENT_FIRST(r0[u]) && JB_RB_RUNE(e2[u].w) == r1[u];
        // a foreign entry in the home slot: linear probing goes on only if a key was displaced from it
        const bool x = go2[u] && !m && e2[u].z != JB_PARENT_EMPTY && (e2[u].w & JB_RB_CONT);
        const double w2 = __longlong_as_double(((long long)e2[u].y << 32) | (long long)e2[u].x);
        const bool cand = m && jb_w_positive(w2);  // val > 0 -> edge (T:479-481)
        uint32_t pe = kSgNone;
        {
          const uint32_t bm = __ballot_sync(FULL, cand);
          if (bm) {
            uint32_t pb = 0;
            if (lane == 0) pb = atomicAdd(&ctl[SG_NPOOL], (uint32_t)__popc(bm));
            pb = __shfl_sync(FULL, pb, 0);
            if (cand) {
              const uint32_t idx = pb + __popc(bm & lt_mask);
              if (idx < (uint32_t)kSgPool) {
                WX[idx] = w2;
                nx[idx] = (uint16_t)kSgNone;
                pe = idx;
              } else {
                ctl[SG_OVF] = 1u;
              }
            }
          }
        }
        if (live) {
          W1[i] = __longlong_as_double(((long long)f[u].y << 32) | (long long)f[u].x);
          mask[i] = (MT)(pe != kSgNone ? 3u : 1u);
          head[i] = (uint16_t)pe;
        }
        // does the chain go on?  matched and some key extends r0 r1 by the next rune -> the 3-rune key's home slot
        const uint32_t r2 = m ? rune[i + 2] : kSgSent;
        const bool cont = m && (((e2[u].w >> 21) >> jb_bloom11(r2)) & 1u);
        const bool need = cont || x;
        const uint32_t slot2 = h2[u] & hmask;
        const uint32_t hh = cont ? jb_hash_next(h2[u], r2) : h2[u];
        const uint32_t nslot = cont ? (hh & hmask) : ((slot2 + 1u) & hmask);
        const uint32_t par = cont ? slot2 : JB_PARENT_FIRST(r0[u]);
        const uint32_t tL = cont ? 2u : 1u;
        const uint32_t tm = __ballot_sync(FULL, need);
        if (tm) {
          uint32_t tb = 0;
          if (lane == 0) tb = atomicAdd(&ctl[SG_NT0], (uint32_t)__popc(tm));
          tb = __shfl_sync(FULL, tb, 0);
          if (need) {
            const uint32_t idx = tb + __popc(tm & lt_mask);
            const uint4 tk = make_uint4(nslot, hh, par, i | (pe << 12) | (tL << 25) | (cont ? 1u << 30 : 0u));
            if (idx < (uint32_t)kSgTasks) task[idx] = tk;
            else run_chain(tk);  // no room: finish the chain here
          }
        }
      }
    }
    // (the bucket bases: warp 0 wrote them before it began pass 1; every other warp passed a barrier... not yet: below)
    __syncthreads();
    if ((uint32_t)tid < nb) order[hist[my_bucket] + my_rank] = (uint8_t)tid;

    // ---- pass 2: the chains that go on, level by level: every live chain does one probe per level, the survivors
    // are compacted back into the list (all lanes busy whatever the chain lengths) -----------------------------
    {
      uint32_t which = 0;
      for (;;) {
        const uint32_t ncur = min(ctl[SG_NT0 + which], (uint32_t)kSgTasks);
        if (ncur == 0) break;
        uint4 tk[kSgTPT], e[kSgTPT];
        bool on[kSgTPT];
#pragma unroll
        for (int k = 0; k < kSgTPT; k++) {
          const uint32_t t = tid + k * kSgThreads;
          on[k] = t < ncur;
          if (on[k]) {
            tk[k] = task[t];
            e[k] = __ldg(entries + tk[k].x);
          }
        }
#pragma unroll
        for (int k = 0; k < kSgTPT; k++)
          if (on[k]) on[k] = chain_step(tk[k], e[k]);
        __syncthreads();  // every task of this level has been read
        if (tid == 0) ctl[SG_NT0 + which] = 0;  // (it counts the level after the next)
#pragma unroll
        for (int k = 0; k < kSgTPT; k++) {
          const uint32_t tm = __ballot_sync(FULL, on[k]);
          if (tm) {
            uint32_t tb = 0;
            if (lane == 0) tb = atomicAdd(&ctl[SG_NT0 + (which ^ 1u)], (uint32_t)__popc(tm));
            tb = __shfl_sync(FULL, tb, 0);
            if (on[k]) task[tb + __popc(tm & lt_mask)] = tk[k];
          }
        }
        __syncthreads();
        which ^= 1u;
      }
    }
    __syncthreads();  // (order[] and, when no chain went on, everything pass 1 wrote)

    // ---- route + emit: one lane per block, longest blocks first ---------------------------------------------
    const bool ovf = ctl[SG_OVF] != 0;
    if ((uint32_t)tid < nb) {
      const uint32_t b = order[tid];
      const uint32_t fb = blk_fb[b], o = blk_off[b], nr = blk_nr[b];
      if (blk_flag[b]) {
        push_wide(fb + 3u * (nr - 1u));
      } else if (ovf) {
        push_long(fb + 3u * (nr - 1u), nr);
      } else {
        // calcDagProba (T:502-548) right to left; candidates in ascending length into maxIndexProba's running
        // (prev, best) pair: each is compared with the previous one, the first with minFloat (T:565-578).
        // The next position's weight / mask / first longer candidate are fetched while this one is decided.
        double Rn = 0.0;
        uint32_t i = o + nr - 1u;
        double w1n = W1[i];
        uint32_t mn = mask[i], en = head[i];
        for (uint32_t j = nr; j-- > 0; i--) {
          double v = w1n + Rn;  // pieceFreq + nextBestPiece.proba (T:519-529)
          uint32_t mm = mn >> 1, e = en;
          if (j) {
            w1n = W1[i - 1u];
            mn = mask[i - 1u];
            en = head[i - 1u];
          }
          uint32_t best_d = v >= JB_MINF ? 1u : 0u, last_d = 1u;
          double best_v = v, prev_v = v;
          while (mm) {
            const uint32_t L = (uint32_t)__ffs(mm) + 1u;
            mm &= mm - 1u;
            v = WX[e] + W1[i + L];
            e = nx[e];
            if (v >= prev_v) {
              best_d = L;
              best_v = v;
            }
            prev_v = v;
            last_d = L;
          }
          if (best_d == 0u) {  // best.index == -1 -> return prev (T:574-576)
            best_d = last_d;
            best_v = prev_v;
          }
          W1[i] = best_v;
          mask[i] = (MT)best_d;
          Rn = best_v;
        }
        const uint32_t i0 = fb / 3u;
        if (A.dbg_R) {
          for (uint32_t j = 0; j < nr; j++) {
            A.dbg_R[i0 + j] = W1[o + j];
            A.dbg_D[i0 + j] = (uint8_t)mask[o + j];
          }
        }
        if (!HMM) {
          // findDagPath (T:552-562): one piece per step
          BitAcc2 sa, ea;
          sa.init(A.s_bits);
          ea.init(A.e_bits);
          for (uint32_t k = 0; k < nr;) {
            const uint32_t d = mask[o + k];
            sa.set(fb + 3u * k);
            ea.set(fb + 3u * (k + d) - 1u);
            k += d;
          }
          sa.flush();
          ea.flush();
        } else {
          constexpr uint32_t kRegRun = 24;  // the four best paths of runs up to this length are carried in registers
          auto set_s = [&](uint32_t q) { atomicOr(&A.s_bits[q >> 5], 1u << (q & 31)); };
          auto set_e = [&](uint32_t q) { atomicOr(&A.e_bits[q >> 5], 1u << (q & 31)); };
          uint32_t run_n = 0, run_s = 0;
          double V[4] = {0.0, 0.0, 0.0, 0.0};
          // viterbi's fullPath (T:715-716) in bit form, per state: bits 0..23 = which runes of its best path are E or S
          // (token ends), bits 24..31 = the path's length (a route with from == "" restarts it)
          uint32_t pm[4] = {0, 0, 0, 0};
          for (uint32_t k = 0; k < nr;) {
            const uint32_t d = mask[o + k];
            const bool single = d == 1u;
            if (single) {  // collect singletons (T:233-234): one Viterbi step per rune (T:688-719)
              const uint32_t cp = rune[o + k];
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
                  const double q0 = V[pa] + T.trans[s][0], q1 = V[pb] + T.trans[s][1];
                  double best = JB_MINF;
                  uint32_t from = 0;
                  if (q0 > best) {
                    best = q0;
                    from = 1;
                  }
                  if (q1 > best) {
                    best = q1;
                    from = 2;
                  }
                  W[s] = best + em[s];
                  code |= from << (2 * s);
                  // fullPath[s] = fullPath[route.from] + [s]; fullPath[""] is nil (T:715-716)
                  npm[s] = (from == 0 ? 0u : (from == 1 ? pm[pa] : pm[pb])) + (s >= 2 ? step : (1u << 24));
                }
#pragma unroll
                for (int s = 0; s < 4; s++) {
                  V[s] = W[s];
                  pm[s] = npm[s];
                }
                head[o + k] = (uint16_t)code;  // back-pointers: only read back for runs longer than the register window
              }
              run_n++;
            }
            if (run_n && (!single || k + 1u >= nr)) {  // flush the run: viterbi's tail (T:723-729) + cutHMM (T:273-285)
              const uint32_t q0 = fb + 3u * run_s;
              if (run_n == 1) {
                set_s(q0);
                set_e(q0 + 2u);
              } else if (run_n <= kRegRun) {
                const uint32_t pf = V[2] > V[3] ? pm[2] : pm[3];  // T:723-729
                const uint32_t plen = pf >> 24;
                // path[j] applies to rune j (T:277-283): a short path drops the run's tail
                const uint32_t lm = (1u << plen) - 1u;
                const uint32_t es = ((pf & 0xFFFFFFu) >> (run_n - plen)) & lm;
                const uint32_t starts = ((es << 1) | 1u) & lm;
                or_span(A.s_bits, q0, spread3(starts), spread3(starts >> 16));
                or_span(A.e_bits, q0 + 2u, spread3(es), spread3(es >> 16));
              } else {
                int st2 = V[2] > V[3] ? 2 : 3;
                uint32_t kb = run_s + run_n - 1u, plen = 0;
                for (;;) {  // back-trace; stops early where route.from == "" (T:715-716)
                  const uint32_t code = head[o + kb];
                  head[o + kb] = (uint16_t)(st2 >= 2 ? 0x100 : 0);  // the state of this path entry is E or S
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
                  const bool es = head[o + run_s + shift + j2] & 0x100;
                  const uint32_t qq = q0 + 3u * j2;
                  if (prev_es) set_s(qq);
                  if (es) set_e(qq + 2u);
                  prev_es = es;
                }
              }
              run_n = 0;
            }
            if (!single) {
              set_s(fb + 3u * k);
              set_e(fb + 3u * (k + d) - 1u);
            }
            k += d;
          }
        }
      }
    }
    if (ovf) G = max(8u, G / 2u);
    __syncthreads();
  }
}

template <bool HMM, typename MT>
static void launch_seg_t(const JbTables& T, const SegArgs& A, unsigned grid, cudaStream_t st) {
  const size_t sm = SegLayout<MT>::bytes;
  // (per device: the attribute belongs to the function in the current context)
  cudaFuncSetAttribute(k_seg<HMM, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k_seg<HMM, MT><<<grid, kSgThreads, sm, st>>>(T, A);
}

int launch_seg(const JbTables& T, const SegArgs& A, bool hmm, int num_sms, cudaStream_t st) {
  const unsigned grid = (unsigned)num_sms * 2u;  // persistent: two resident CTAs of 512 threads per SM (105 KB of shared memory each)
  if (T.max_delta <= 16) {
    if (hmm) launch_seg_t<true, uint16_t>(T, A, grid, st);
    else launch_seg_t<false, uint16_t>(T, A, grid, st);
  } else {
    if (hmm) launch_seg_t<true, uint32_t>(T, A, grid, st);
    else launch_seg_t<false, uint32_t>(T, A, grid, st);
  }
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace jb
