// bb_wide.cuh -- one CTA per environment: the episode runner for LONG polynomials (cyclic-n and other non-binomial
// ideals, BASELINE configs[4]), sm_100a.
//
// The warp-per-environment runner (bb_device.cuh) is built for binomial ideals, where a polynomial is two terms and the
// whole step is a few hundred instructions.  On cyclic-6 the picture is the opposite: a dividend of
// ~100 (up to ~550) terms is reduced ~83 times per step by polynomials of ~33 terms, each reduction scanning ~93 of
// ~444 reducer lead monomials, and an episode is a serial chain of ~80 000 such additions.  There the latency of ONE
// addition decides everything, so a whole CTA (BBW_WARPS warps) works on one environment:
//
//   * the dividend h lives in SHARED memory (two buffers of max_poly_terms terms, ping-pong); only remainder terms
//     and the reducers' term lists touch global memory;
//   * divisor search (buchberger.cpp:27-32): every thread tests a strided slice of the reducer lead monomials, the
//     first hit is a block-wide minimum;
//   * h <- h - (LT h / LT f) f (buchberger.cpp:33-35; Polynomial operator+ / Term*Polynomial, polynomials.cpp:148-202)
//     is a block-wide merge by RANK: every term of both lists binary-searches the other list, equal monomials are
//     paired (the h-side term carries the sum, the f-side term retires), every term is written at its merged rank and
//     one stream compaction drops retired / cancelled terms.  No windows, no serial dependence on the list length.
//   * pair selection, pair removal and update() (Gebauer-Moeller) are the warp routines of bb_device.cuh run by warp 0.
//
// Results are bit-identical to the warp runner (same arithmetic, same order of terms); tests run both.
#pragma once
#include "bb_device.cuh"

#ifndef BBW_WARPS
#define BBW_WARPS 8
#endif
#define BBW_THREADS (BBW_WARPS * 32)
static_assert(BBW_WARPS % 4 == 0, "per-warp partial results are read four at a time");
#ifndef BBW_MIN_CTAS
#define BBW_MIN_CTAS 2   // 128 registers; A/B on cyclic-6 (profiles/README.md v9): 4 warps x 7 -> 1923 ms, 4 x 4 -> 1834, 8 x 3 -> 1524, 8 x 2 -> 1437
#endif
#ifndef BBW_PER
#define BBW_PER 4        // merged positions per thread that block_merge keeps in registers
#endif
// merge variants (BBRunArgs::wide_flags).  Every variant is bit-identical; the flags exist so that the parity tests
// reach the fallback paths that the benchmark ideals rarely or never take.
#define BBW_FLAG_RANK_MERGE 1   // block_merge_rank for every addition (else: only when the reducer outgrows the staging buffer)
#define BBW_FLAG_TWO_WALKS 2    // merge path with the count + write walks (else: only for slices longer than BBW_PER)
#ifndef BBW_STAGE
#define BBW_STAGE 512    // terms of the staging buffer for the f-side operand of a merge (12 bytes each)
#endif

struct WideShared {
  // per-warp partial results, double-buffered (a slot is rewritten two barriers later, after every thread has read it)
  __align__(16) uint32_t wmin[2][BBW_WARPS];   // divisor search: (position in G_ << 16) | basis index of each warp's first hit
  __align__(16) int wcnt[2][BBW_WARPS];        // block_merge: per-warp survivor counts
  uint32_t pr; int row;     // the pair warp 0 took this step
  uint64_t gam;
  long long upd;            // result of warp_add_basis
  int status;
};

// Divisor search (buchberger.cpp:27-32), split in two so that it can share a barrier with other work:
// search_publish: every thread tests a strided slice of the reducer lead monomials, its first hit is its minimum, the
// warp's minimum goes to sh.wmin[slot][warp] packed as (position in G_ << 16) | basis index (both < 65536; the basis
// index ridx[r] is loaded next to the lead monomial so that the head record's address is known right after the
// barrier); search_collect (after a block barrier): the block-wide first divisor, or -1.
#define BBW_NONE 0xffffffffu
template <int NV>
__device__ __forceinline__ void search_publish(WideShared& sh, int slot, const uint64_t* __restrict__ rlm,
                                               const uint32_t* __restrict__ ridx, int nR, uint64_t lead, bool sorted) {
  typedef KL<NV> K;
  uint32_t best = BBW_NONE;
  // sorted: G_ ascends in lead monomial (keys descend), so a thread may stop at its first reducer whose lead monomial
  // exceeds `lead` (key below lead's): nothing after it can divide.  Same early exit as warp_reduce.
  const uint64_t stop = sorted ? lead : 0ull;
#pragma unroll 1
  for (int r = threadIdx.x; r < nR; r += BBW_THREADS) {
    const uint64_t l = rlm[r];
    const uint32_t ix = ridx[r];
    if (l < stop) break;
    if (K::divides(l, lead)) { best = ((uint32_t)r << 16) | ix; break; }
  }
  best = __reduce_min_sync(BB_FULL, best);
  if (bb_lane() == 0) sh.wmin[slot][threadIdx.x >> 5] = best;
}
__device__ __forceinline__ int search_collect(const WideShared& sh, int slot, uint32_t& gidx) {
  uint32_t m = BBW_NONE;
#pragma unroll
  for (int w = 0; w < BBW_WARPS; w += 4) {
    const uint4 v = *reinterpret_cast<const uint4*>(&sh.wmin[slot][w]);
    m = min(min(m, min(v.x, v.y)), min(v.z, v.w));
  }
  gidx = m & 0xffffu;
  return m == BBW_NONE ? -1 : (int)(m >> 16);
}
// sum of the warps' counts, and of the warps before `warp`
__device__ __forceinline__ void counts_collect(const WideShared& sh, int slot, int warp, int& before, int& all) {
  before = 0; all = 0;
#pragma unroll
  for (int w = 0; w < BBW_WARPS; w += 4) {
    const int4 v = *reinterpret_cast<const int4*>(&sh.wcnt[slot][w]);
    all += v.x + v.y + v.z + v.w;
    before += (w < warp ? v.x : 0) + (w + 1 < warp ? v.y : 0) + (w + 2 < warp ? v.z : 0) + (w + 3 < warp ? v.w : 0);
  }
}
template <int NV>
__device__ __forceinline__ int block_first_divisor(WideShared& sh, int& slot, const uint64_t* __restrict__ rlm,
                                                   const uint32_t* __restrict__ ridx, int nR, uint64_t lead, bool sorted,
                                                   uint32_t& gidx) {
  search_publish<NV>(sh, slot, rlm, ridx, nR, lead, sorted);
  __syncthreads();
  const int found = search_collect(sh, slot, gidx);
  slot ^= 1;
  return found;
}

// out = cA * mA * A + cB * mB * B for term lists in ascending key order (see warp_merge in bb_device.cuh for the
// conventions: adjX = key(mX) - bias, coefficients in [1,p), cancelled terms dropped).  A, B, O may be shared or global
// memory; O must not alias A or B and needs room for nA + nB terms (`cap`).  Returns the number of output terms, -1 if
// nA + nB > cap, -3 if a produced key has a guard bit set.  Ends with a block barrier: O is visible to every thread.
// This is the merge by RANK (every term binary-searches the other list, then one chunked stream compaction): the
// fallback of block_merge for a B operand longer than the staging buffer.
template <int NV>
__device__ __noinline__ int block_merge_rank(WideShared& sh, int& slot, BBField F, const uint64_t* Ak, const uint32_t* Ac,
                                           int nA, uint32_t cA, uint64_t adjA, const uint64_t* Bk, const uint32_t* Bc, int nB,
                                           uint32_t cB, uint64_t adjB, uint64_t* Ok, uint32_t* Oc, int cap) {
  typedef KL<NV> K;
  const int total = nA + nB;
  if (total > cap) return -1;
  uint64_t guard = 0;
  // phase 1: every term goes to its merged rank; a term that retires or cancels is written with coefficient 0
#pragma unroll 1
  for (int i = threadIdx.x; i < nA; i += BBW_THREADS) {
    const uint64_t a = Ak[i] + adjA;
    guard |= a;
    uint32_t c = (cA == 1u) ? Ac[i] : bbf_mulmod(F, Ac[i], cA);
    int lo = 0, hi = nB;              // lo = #{j : B_j < a}
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (Bk[mid] + adjB < a) lo = mid + 1; else hi = mid; }
    if (lo < nB && Bk[lo] + adjB == a) c = bbf_addmod(F, c, (cB == 1u) ? Bc[lo] : bbf_mulmod(F, Bc[lo], cB));
    Ok[i + lo] = a; Oc[i + lo] = c;
  }
#pragma unroll 1
  for (int j = threadIdx.x; j < nB; j += BBW_THREADS) {
    const uint64_t b = Bk[j] + adjB;
    guard |= b;
    int lo = 0, hi = nA;              // lo = #{i : A_i <= b}
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (Ak[mid] + adjA <= b) lo = mid + 1; else hi = mid; }
    const bool paired = lo > 0 && Ak[lo - 1] + adjA == b;   // its A partner (at rank j + lo - 1) carries the sum
    Ok[j + lo] = b; Oc[j + lo] = paired ? 0u : ((cB == 1u) ? Bc[j] : bbf_mulmod(F, Bc[j], cB));
  }
  const int bad = __syncthreads_or((guard & K::g_all) != 0ull);
  // phase 2: in-place stream compaction, one block-wide chunk at a time.  A chunk's survivors land at or below their
  // own positions, i.e. inside what this and earlier chunks have already read into registers.
  int no = 0;
  const int warp = threadIdx.x >> 5;
  const uint32_t ltm = bb_lt_mask();
#pragma unroll 1
  for (int c0 = 0; c0 < total; c0 += BBW_THREADS) {
    const int idx = c0 + threadIdx.x;
    uint64_t k = 0; uint32_t c = 0;
    if (idx < total) { k = Ok[idx]; c = Oc[idx]; }
    const uint32_t live = __ballot_sync(BB_FULL, c != 0u);
    if (bb_lane() == 0) sh.wcnt[slot][warp] = __popc(live);
    __syncthreads();
    int before, all;
    counts_collect(sh, slot, warp, before, all);
    if (c != 0u) { const int pos = no + before + __popc(live & ltm); Ok[pos] = k; Oc[pos] = c; }
    no += all;
    slot ^= 1;
  }
  __syncthreads();
  return bad ? -3 : no;
}

// out = cA * mA * A + cB * mB * B, same contract as block_merge_rank, by MERGE PATH.  sk / sc: BBW_STAGE terms of shared
// memory.  B (the reducer's tail: ~33 terms on cyclic-6) is staged there with its multiplier applied in one coalesced
// pass; thread t then owns the merged positions [t * per, (t + 1) * per): one binary search along its diagonal finds
// where its slice starts in A and in B (at most log2(nB) steps on shared memory), a sequential walk merges the slice
// (equal monomials: the A term carries the sum and the B term retires, a zero sum drops out), a block scan of the
// slices' survivor counts gives every slice its output offset, and a second walk writes the survivors in place.  Three
// barriers per addition whatever the length of the dividend; the rank merge needed log2(nB) dependent loads for EVERY
// term of A and one barrier per BBW_THREADS merged terms, which is what bounded the longest cyclic-6 episodes
// (profiles/README.md v6/v9).
// One merged position: consumes the smaller of the two heads (a = A[ia] + adjA, b = staged B[ib]; ~0 past the end) and
// returns its key k and coefficient c, c == 0 when nothing survives there (a cancelled sum, or a B term whose equal A
// term carried the sum).
__device__ __forceinline__ void merge_take(const BBField& F, const uint64_t* Ak, const uint32_t* Ac, int nA, uint32_t cA,
                                           uint64_t adjA, const uint64_t* sk, const uint32_t* sc, int nB, int& ia, int& ib,
                                           uint64_t& a, uint64_t& b, uint64_t& k, uint32_t& c) {
  if (a <= b) {   // A first on ties
    k = a;
    c = (cA == 1u) ? Ac[ia] : bbf_mulmod(F, Ac[ia], cA);
    if (a == b) c = bbf_addmod(F, c, sc[ib]);
    ia++;
    a = ia < nA ? Ak[ia] + adjA : ~0ull;
  } else {        // its A partner, if any, is the A term just before it
    k = b;
    c = (ia > 0 && Ak[ia - 1] + adjA == b) ? 0u : sc[ib];
    ib++;
    b = ib < nB ? sk[ib] : ~0ull;
  }
}
template <int NV, bool WRITE>
__device__ __forceinline__ int merge_walk(BBField F, const uint64_t* Ak, const uint32_t* Ac, int nA, uint32_t cA, uint64_t adjA,
                                          const uint64_t* sk, const uint32_t* sc, int nB, int ia, int ib, int count,
                                          uint64_t* Ok, uint32_t* Oc, uint64_t& guard) {
  int cnt = 0;
  uint64_t a = ia < nA ? Ak[ia] + adjA : ~0ull, b = ib < nB ? sk[ib] : ~0ull;
#pragma unroll 1
  for (int d = 0; d < count; d++) {
    uint64_t k; uint32_t c;
    merge_take(F, Ak, Ac, nA, cA, adjA, sk, sc, nB, ia, ib, a, b, k, c);
    if (!WRITE) guard |= k;
    if (c != 0u) { if (WRITE) { Ok[cnt] = k; Oc[cnt] = c; } cnt++; }
  }
  return cnt;
}

template <int NV>
__device__ __forceinline__ int block_merge(WideShared& sh, int& slot, BBField F, const uint64_t* Ak, const uint32_t* Ac,
                                           int nA, uint32_t cA, uint64_t adjA, const uint64_t* Bk, const uint32_t* Bc, int nB,
                                           uint32_t cB, uint64_t adjB, uint64_t* Ok, uint32_t* Oc, int cap, uint64_t* sk,
                                           uint32_t* sc, const uint64_t* __restrict__ rlm, const uint32_t* __restrict__ ridx,
                                           int nR, bool sorted, int flags, int& next_found, uint32_t& next_idx) {
  typedef KL<NV> K;
  next_found = -2;   // -2: the divisor of the result's lead monomial was not searched here
  const int total = nA + nB;
  if (total > cap) return -1;
  if (nB > BBW_STAGE || (flags & BBW_FLAG_RANK_MERGE)) return block_merge_rank<NV>(sh, slot, F, Ak, Ac, nA, cA, adjA, Bk, Bc, nB, cB, adjB, Ok, Oc, cap);
  uint64_t guard = 0;
#pragma unroll 1
  for (int j = threadIdx.x; j < nB; j += BBW_THREADS) {
    const uint64_t k = Bk[j] + adjB;
    guard |= k;
    sk[j] = k;
    sc[j] = (cB == 1u) ? Bc[j] : bbf_mulmod(F, Bc[j], cB);
  }
  __syncthreads();
  const int per = (total + BBW_THREADS - 1) / BBW_THREADS;
  const int d0 = min((int)threadIdx.x * per, total), count = min(per, total - d0);
  int lo = max(0, d0 - nB), hi = min(d0, nA);   // lo -> number of A terms among the first d0 merged terms
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (Ak[mid] + adjA <= sk[d0 - 1 - mid]) lo = mid + 1; else hi = mid;
  }
  const int ia = lo, ib = d0 - lo;
  // The result's lead monomial is known before the merge runs unless the two heads coincide (their sum may cancel):
  // it is the smaller head.  Its divisor search then shares the barrier of the survivor-count scan, which takes one
  // barrier and the scan of the reducer lead monomials off the serial chain of an addition.
  bool searched = false;
  if (rlm != nullptr && total > 0) {
    const uint64_t a0 = nA > 0 ? Ak[0] + adjA : ~0ull, b0 = nB > 0 ? sk[0] : ~0ull;
    if (a0 != b0) { search_publish<NV>(sh, slot, rlm, ridx, nR, a0 < b0 ? a0 : b0, sorted); searched = true; }
  }
  // Slices of at most BBW_PER positions (dividends up to BBW_PER * BBW_THREADS terms: nearly all) are walked once and
  // held in registers across the scan; longer ones are walked twice (count, then write).
  uint64_t rk[BBW_PER]; uint32_t rc[BBW_PER];
  int mine = 0;
  const bool small = per <= BBW_PER && !(flags & BBW_FLAG_TWO_WALKS);
  if (small) {
    int xa = ia, xb = ib;
    uint64_t a = xa < nA ? Ak[xa] + adjA : ~0ull, b = xb < nB ? sk[xb] : ~0ull;
#pragma unroll
    for (int d = 0; d < BBW_PER; d++) {
      rc[d] = 0u; rk[d] = 0ull;
      if (d < count) {
        merge_take(F, Ak, Ac, nA, cA, adjA, sk, sc, nB, xa, xb, a, b, rk[d], rc[d]);
        guard |= rk[d];
        mine += rc[d] != 0u;
      }
    }
  } else {
    mine = merge_walk<NV, false>(F, Ak, Ac, nA, cA, adjA, sk, sc, nB, ia, ib, count, Ok, Oc, guard);
  }
  int incl = mine;   // inclusive scan of the survivor counts inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(BB_FULL, incl, o);
    if (bb_lane() >= o) incl += v;
  }
  const int warp = threadIdx.x >> 5;
  if (bb_lane() == 31) sh.wcnt[slot][warp] = incl;
  const int bad = __syncthreads_or((guard & K::g_all) != 0ull);
  int before, all;
  counts_collect(sh, slot, warp, before, all);
  if (searched) next_found = search_collect(sh, slot, next_idx);
  slot ^= 1;
  const int base = before + incl - mine;
  if (small) {
    int o = base;
#pragma unroll
    for (int d = 0; d < BBW_PER; d++)
      if (rc[d] != 0u) { Ok[o] = rk[d]; Oc[o] = rc[d]; o++; }
  } else {
    merge_walk<NV, true>(F, Ak, Ac, nA, cA, adjA, sk, sc, nB, ia, ib, count, Ok + base, Oc + base, guard);
  }
  __syncthreads();
  return bad ? -3 : all;
}

// One environment step by the whole CTA (BuchbergerEnv::step, buchberger.cpp:318-329, with the pair chosen by
// `strategy`).  e and every scalar below are block-uniform.  hk / hc: the two shared dividend buffers of `cap` terms.
// Returns the number of polynomial additions; `pair` receives (j << 16) | i.
template <int NV>
__device__ __forceinline__ int block_step(const BBParams& P, Env& e, WideShared& sh, int& slot, uint64_t* hk, uint32_t* hc,
                                          int cap, uint64_t* sk, uint32_t* sc, int flags, int strategy, uint32_t* sel_rng,
                                          uint32_t& pair, Ctr& ct) {
  typedef KL<NV> K;
  const BBField F = P.F;
  const int tid = threadIdx.x;
  if (tid < 32) {   // warp 0: select the pair and take it out of P
    Env e0 = e;
    const int row = warp_select<NV>(P, e0, strategy, sel_rng);
    uint32_t pr; uint64_t gam;
    warp_take_pair(P, e0, row, pr, gam);
    if (tid == 0) { sh.pr = pr; sh.gam = gam; }
  }
  __syncthreads();
  e.nP--;
  const uint32_t pr = sh.pr;
  const uint64_t gam = sh.gam;
  pair = pr;
  const GHeadMem* gh = ENV_PTR(GHeadMem, e, P, o_ghead);
  const uint64_t* tk = ENV_PTR(uint64_t, e, P, o_tkey);
  const uint32_t* tc = ENV_PTR(uint32_t, e, P, o_tcoef);
  const GHead hf = load_head(gh + (pr & 0xffffu)), hg = load_head(gh + (pr >> 16));
  e.guard |= gam;
  ct.tread += hf.len + hg.len;
  int sug;
  {  // sugar of the S-polynomial: max(deg(gamma / LM f) + sug f, deg(gamma / LM g) + sug g)
    const int cg0 = (int)(uint32_t)(gam >> K::dshift);
    const int sf = (int)hf.sug + (int)(uint32_t)(hf.lm >> K::dshift) - cg0, sg = (int)hg.sug + (int)(uint32_t)(hg.lm >> K::dshift) - cg0;
    sug = sf > sg ? sf : sg;
  }
  // s = (gamma / LT f) tail(f) - (gamma / LT g) tail(g): the lead terms cancel exactly (buchberger.cpp:18-21)
  int cur = 0, pos = 0;
  int found; uint32_t fidx;   // first divisor of the current lead monomial (found == -2: not searched yet)
  int n = block_merge<NV>(sh, slot, F, tk + hf.off + 1, tc + hf.off + 1, (int)hf.len - 1, hf.invlc, gam - hf.lm,
                          tk + hg.off + 1, tc + hg.off + 1, (int)hg.len - 1, F.p - hg.invlc, gam - hg.lm, hk, hc, cap, sk, sc,
                          ENV_PTR(uint64_t, e, P, o_rlm), ENV_PTR(uint32_t, e, P, o_ridx), e.nG, P.sort_reducers != 0, flags, found, fidx);
  if (n < 0) { e.status = n == -1 ? BB_STATUS_OVERFLOW_SCRATCH : BB_STATUS_OVERFLOW_EXPONENT; return 1; }
  if (e.guard & K::g_all) { e.status = BB_STATUS_OVERFLOW_EXPONENT; return 1; }
  ct.twrite += (unsigned)n;
  // reduce(s, G_), buchberger.cpp:24-49
  const uint64_t* rlm = ENV_PTR(uint64_t, e, P, o_rlm);
  const uint32_t* ridx = ENV_PTR(uint32_t, e, P, o_ridx);
  uint64_t* rk = ENV_PTR(uint64_t, e, P, o_tkey) + e.nT;
  uint32_t* rc = ENV_PTR(uint32_t, e, P, o_tcoef) + e.nT;
  const int rcap = P.max_terms - e.nT, nR = e.nG;
  const bool sorted = P.sort_reducers != 0;
  int steps = 0, rlen = 0;
  while (n > 0) {
    const int hb = cur * cap + pos;
    const uint64_t* ck = hk + hb;
    const uint32_t* cc = hc + hb;
    const uint64_t lead = ck[0];
    if (found == -2) found = block_first_divisor<NV>(sh, slot, rlm, ridx, nR, lead, sorted, fidx);
    ct.lms += (found >= 0) ? (unsigned)(found + 1) : (unsigned)nR;
    if (found >= 0) {
      const GHead f = load_head(gh + fidx);
      const uint32_t c = bbf_mulmod(F, cc[0], f.invlc);
      const uint32_t nc = F.p - c;              // c != 0
      const uint64_t adj = lead - f.lm;         // key(LM h / LM f) - bias
      const int sf = (int)f.sug + (int)(uint32_t)(f.lm >> K::dshift) - (int)(uint32_t)(lead >> K::dshift);
      sug = sf > sug ? sf : sug;
      ct.tread += (unsigned)n + f.len;
      const int ob = cur ^ 1;
      const int n2 = block_merge<NV>(sh, slot, F, ck + 1, cc + 1, n - 1, 1u, 0ull, tk + f.off + 1, tc + f.off + 1,
                                     (int)f.len - 1, nc, adj, hk + ob * cap, hc + ob * cap, cap, sk, sc, rlm, ridx, nR, sorted, flags, found, fidx);
      if (n2 < 0) { e.status = n2 == -1 ? BB_STATUS_OVERFLOW_SCRATCH : BB_STATUS_OVERFLOW_EXPONENT; return 1 + steps; }
      n = n2; cur = ob; pos = 0;
      ct.twrite += (unsigned)n;
      steps++;
    } else {
      if (rlen >= rcap) { e.status = BB_STATUS_OVERFLOW_TERMS; return 1 + steps; }
      if (tid == 0) { rk[rlen] = lead; rc[rlen] = cc[0]; }
      rlen++; ct.moves++;
      pos++; n--;
      found = -2;
    }
  }
  if (rlen > 0) {
    ct.upb += (unsigned)e.nG; ct.upp += (unsigned)e.nP;
    __syncthreads();   // the remainder (written by thread 0) before warp 0 reads it
    if (tid < 32) {
      const long long r = warp_add_basis<NV>(P, e.base, e.nG, e.nP, e.nT, rlen, sug);
      if (tid == 0) sh.upd = r;
    }
    __syncthreads();
    const long long r = sh.upd;
    if (r < 0) {
      e.status = (r == -1) ? BB_STATUS_OVERFLOW_PAIRS : (r == -2 ? BB_STATUS_OVERFLOW_BASIS : BB_STATUS_OVERFLOW_EXPONENT);
      return 1 + steps;
    }
    e.nP = (int)(r & 0xffffffffll);
    ct.upp += (unsigned)(r >> 32);
    e.nG++; e.nT += rlen;
  }
  if (e.nP == 0) e.status = BB_STATUS_DONE;
  return 1 + steps;
}
