// bb_wide.cuh -- one CTA per environment: the episode runner for LONG polynomials (cyclic-n and other non-binomial
// ideals, BASELINE configs[4]), sm_100a.
//
// The warp-per-environment runner (bb_device.cuh) is built for binomial ideals, where a polynomial is two terms and the
// whole step is a few hundred instructions.  On cyclic-6 the picture is the opposite: a dividend of
// ~100 (up to ~550) terms is reduced ~83 times per step by polynomials of ~33 terms, each reduction scanning ~93 of
// ~444 reducer lead monomials, and an episode is a serial chain of ~80 000 such additions (697 000 in the longest of
// 1024).  There the latency of ONE addition decides everything, so a whole CTA (BBW_WARPS warps) works on one
// environment and an addition h <- h - (LT h / LT f) f (buchberger.cpp:33-35; Polynomial operator+ / Term*Polynomial,
// polynomials.cpp:148-202) is TWO block barriers long:
//
//   phase 1  the reducer's tail B (|B| <= BBW_BSTAGE terms, one thread each, loaded from the term arena one addition
//            AHEAD) is scaled and every B term binary-searches the dividend's tail A, which lives in shared memory.
//            Its rank gives its slot in the merged sequence; a bit per slot goes to a "B" bitmap and, when the monomial
//            coincides with an A term (whose coefficient then absorbs it), to a "dead" bitmap.  The warps that hold no
//            B term meanwhile run the divisor search (buchberger.cpp:27-32) for the NEXT lead monomial, which is known
//            before the merge happens: it is the smaller of the two heads (if they coincide, their sum -- and if that
//            is zero the search simply waits for the merge).
//   barrier
//   phase 2  every thread owns merged slots tid, tid + 256, ...: two prefix popcounts over the bitmaps say whether the
//            slot is an A or a B term, which one, and where it lands after the dead slots are squeezed out; the term is
//            written to the other dividend buffer.  The head record and the tail of the next reducer are fetched here.
//   barrier
//
// A coefficient sum that cancels to zero (probability ~ 1/p per coinciding pair) is written as it is, flagged, and
// removed by an in-place stream compaction before the next addition.  No windows, no scan of survivor counts, no
// serial dependence on the length of the dividend.  Reducers longer than BBW_BSTAGE terms take the rank merge
// (block_merge_rank: every term binary-searches the other list, one chunked compaction), which is also the test
// variant bb_set_wide(2).
//
// Pair selection and update() (Gebauer-Moeller) are the warp routines of bb_device.cuh run by warp 0; the pair is
// taken out of P by the whole block.  Results are bit-identical to the warp runner (same arithmetic, same order of
// terms, same counters); tests run both.
#pragma once
#include "bb_device.cuh"

#ifndef BBW_WARPS
#define BBW_WARPS 8
#endif
#define BBW_THREADS (BBW_WARPS * 32)
static_assert(BBW_WARPS % 4 == 0, "per-warp partial results are read four at a time");
#ifndef BBW_MIN_CTAS
#define BBW_MIN_CTAS 2
#endif
// variants (BBRunArgs::wide_flags).  Every variant is bit-identical; the flags exist so that the parity tests reach the
// fallback paths that the benchmark ideals rarely take.
#define BBW_FLAG_RANK_MERGE 1   // block_merge_rank for every addition (else: only when the reducer outgrows BBW_BSTAGE)
#define BBW_FLAG_COMPACT 2      // run the zero-coefficient compaction after every addition (else: only when a sum cancelled)
#ifndef BBW_BSTAGE
#define BBW_BSTAGE BBW_THREADS  // terms of a reducer's tail the bitmap merge takes: one per thread
#endif
static_assert(BBW_BSTAGE <= BBW_THREADS, "one B term per thread");
#define BBW_SENT 16                           // sentinel keys (all ones) kept behind every dividend in shared memory: see rank4
#define BBW_MAXCAP 4096                       // longest dividend (max_poly_terms) the bitmaps are sized for
#define BBW_BWORDS (BBW_MAXCAP / 32 + 1)

#ifdef BBW_TIMING   // diagnosis build only (make variant VFLAGS=-DBBW_TIMING): cycles per phase of an addition, warps 0 and 7
static __device__ unsigned long long bbw_dbg[64];
#define BBW_T(i) do { if (bb_lane() == 0 && ((threadIdx.x >> 5) == 0 || (threadIdx.x >> 5) == BBW_WARPS - 1)) { \
    const long long t_ = clock64(); atomicAdd(&bbw_dbg[((threadIdx.x >> 5) ? 16 : 0) + (i)], (unsigned long long)(t_ - pp.tlast)); pp.tlast = t_; } } while (0)
#else
#define BBW_T(i) do { } while (0)
#endif
struct WideBits {            // one merge's bitmaps over the merged (uncompacted) slots, double-buffered by parity
  uint32_t ub[BBW_BWORDS];   // slot holds a B term
  uint32_t db[BBW_BWORDS];   // slot holds a B term whose monomial coincides with the A term in the slot before it
};

struct WideShared {
  // per-warp partial results, double-buffered (a slot is rewritten two barriers later, after every thread has read it)
  __align__(16) uint32_t wmin[2][BBW_WARPS];   // divisor search: (position in G_ << 16) | basis index of each warp's first hit
  __align__(16) int wcnt[2][BBW_WARPS];        // compaction: per-warp survivor counts
  __align__(16) WideBits bits[2];
  uint32_t flags;           // slow path after a barrier that reported trouble: 1 exponent overflow, 2 a coefficient sum cancelled
  uint32_t pr; int row;     // the pair taken this step
  uint64_t gam;
  long long upd;            // result of warp_add_basis
};

// Divisor search (buchberger.cpp:27-32), split in two so that it can share a barrier with other work:
// search_publish: threads [first, BBW_THREADS) test a strided slice of the reducer lead monomials, a thread's first hit
// is its minimum, the warp's minimum goes to sh.wmin[slot][warp] packed as (position in G_ << 16) | basis index (both
// < 65536; the basis index ridx[r] is loaded next to the lead monomial so that the head record's address is known right
// after the barrier); search_collect (after a block barrier): the block-wide first divisor, or -1.
#define BBW_NONE 0xffffffffu
template <int NV>
__device__ __forceinline__ void search_publish(WideShared& sh, int slot, const uint64_t* __restrict__ rlm,
                                               const uint32_t* __restrict__ ridx, int nR, uint64_t lead, bool sorted,
                                               int first) {
  typedef KL<NV> K;
  uint32_t best = BBW_NONE;
  // sorted: G_ ascends in lead monomial (keys descend), so a thread may stop at its first reducer whose lead monomial
  // exceeds `lead` (key below lead's): nothing after it can divide.  Same early exit as warp_reduce.
  const uint64_t stop = sorted ? lead : 0ull;
  const int st = (int)threadIdx.x - first, stride = BBW_THREADS - first;
  if (st >= 0) {
#pragma unroll 1
    for (int r = st; r < nR; r += stride) {
      const uint64_t l = rlm[r];
      const uint32_t ix = ridx[r];
      if (l < stop) break;
      if (K::divides(l, lead)) { best = ((uint32_t)r << 16) | ix; break; }
    }
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
  search_publish<NV>(sh, slot, rlm, ridx, nR, lead, sorted, 0);
  __syncthreads();
  const int found = search_collect(sh, slot, gidx);
  slot ^= 1;
  return found;
}

// In-place stream compaction of (Ok, Oc)[0, total): terms with coefficient 0 drop out, order kept.  One block-wide chunk
// at a time: a chunk's survivors land at or below their own positions, i.e. inside what this and earlier chunks have
// already read into registers.  Ends with a block barrier.  Returns the number of survivors.  (The per-warp counts
// alternate between the two halves of sh.wcnt, which nothing else uses: the last read of a half lies before the
// barrier that ends the call, so every call may start with half 0.)
static __device__ __noinline__ int block_compact_zeros(WideShared& sh, uint64_t* Ok, uint32_t* Oc, int total) {
  int no = 0, slot = 0;
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
  if (threadIdx.x < BBW_SENT) Ok[no + threadIdx.x] = ~0ull;   // see rank4
  __syncthreads();
  return no;
}

// out = cA * mA * A + cB * mB * B for term lists in ascending key order (see warp_merge in bb_device.cuh for the
// conventions: adjX = key(mX) - bias, coefficients in [1,p), cancelled terms dropped).  A, B, O may be shared or global
// memory; O must not alias A or B and needs room for nA + nB terms (`cap`).  Returns the number of output terms, -1 if
// nA + nB > cap, -3 if a produced key has a guard bit set.  Ends with a block barrier: O is visible to every thread.
// This is the merge by RANK (every term binary-searches the other list, then one chunked stream compaction): the
// fallback of the bitmap merge for a reducer longer than BBW_BSTAGE terms.
template <int NV>
__device__ __noinline__ int block_merge_rank(WideShared& sh, BBField F, const uint64_t* Ak, const uint32_t* Ac,
                                           int nA, uint32_t cA, uint64_t adjA, const uint64_t* Bk, const uint32_t* Bc, int nB,
                                           uint32_t cB, uint64_t adjB, uint64_t* Ok, uint32_t* Oc, int cap) {
  typedef KL<NV> K;
  const int total = nA + nB;
  if (total > cap) return -1;
  uint64_t guard = 0;
  // every term goes to its merged rank; a term that retires or cancels is written with coefficient 0
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
  const int no = block_compact_zeros(sh, Ok, Oc, total);
  return bad ? -3 : no;
}

// What an addition hands to the next one.  Block-uniform unless noted; every member is a scalar so that the whole
// record lives in registers.
struct WidePipe {
  int found;            // position in G_ of the first divisor of the current lead monomial, -1: none, -2: not searched yet
  uint32_t fidx;        // its basis index
  uint64_t lead;        // found != -2: the current lead term (predicted before the merge that produced it ran)
  uint32_t lc;
  bool staged;          // the reducer (found, fidx) is loaded and staged for the next addition: everything below
  GHead f;
  uint32_t nc;          // -(LC h / LC f)
  uint64_t adj;         // key(LM h / LM f) - bias
  uint64_t kB;          // per thread: scaled term tid + 1 of f (valid when tid < f.len - 1 <= BBW_BSTAGE), also in sk / sc[spar]
  uint32_t cB;
  int spar;             // staging buffer that holds the staged reducer's tail
  uint32_t bad;         // per thread: a staged key overflowed its exponent fields (reported at the next barrier)
  int par;              // bitmap set of the next merge
  int used0, used1;     // bitmap words the last merge of each parity may have touched (cleared by the other parity's merge)
#ifdef BBW_TIMING
  long long tlast;
#endif
  __device__ __forceinline__ void init() {
#ifdef BBW_TIMING
    tlast = clock64();
#endif
    found = -2; fidx = 0u; lead = 0ull; lc = 0u; staged = false; nc = 0u; adj = 0ull; kB = 0ull; cB = 0u; spar = 0; bad = 0u;
    par = 0; used0 = used1 = 0;
    f.lm = f.k1 = 0ull; f.invlc = f.c1 = f.sug = f.off = f.len = 0u;
  }
};

// #{i : A_i < k} for an ascending list in shared memory that is followed by BBW_SENT sentinel keys (all ones), by 4-ary
// search: three independent probes per level.  The range [lo, lo + n] that still holds the answer shrinks to a quarter
// per level; a probe beyond its upper end reads an element that is >= k anyway (the list is sorted and the end of a
// range is a probed element >= k, or the sentinels), so no probe needs a bounds test.  A probe reaches at most 10
// positions past the end of the list (checked exhaustively for every length up to 4096): BBW_SENT = 16.
__device__ __forceinline__ int rank4(const uint64_t* A, int nA, uint64_t k) {
  int lo = 0, n = nA;
#pragma unroll 1
  while (n > 3) {
    const int q = (n + 3) >> 2;
    const uint64_t* p = A + lo + q - 1;
    const uint64_t v1 = p[0], v2 = p[q], v3 = p[2 * q];
    lo += ((v1 < k ? 1 : 0) + (v2 < k ? 1 : 0) + (v3 < k ? 1 : 0)) * q;
    n = q;
  }
  const uint64_t v1 = A[lo], v2 = A[lo + 1], v3 = A[lo + 2];
  return lo + (v1 < k ? 1 : 0) + (v2 < k ? 1 : 0) + (v3 < k ? 1 : 0);
}

// Loads nothing, computes everything the addition h <- h - (LT h / LT f) f needs from f's head record and the lead term
// (lead, lc), scales this thread's term of f's tail and stages it in sk / sc[pp.spar ^ 1]: pp.nc, pp.adj, pp.kB, pp.cB,
// pp.spar, pp.staged.  (kraw, craw): raw term tid + 1 of f, loaded by the caller when tid < f.len - 1 <= BBW_BSTAGE.
template <int NV>
__device__ __forceinline__ void stage_reducer(WidePipe& pp, const BBField F, uint64_t lead, uint32_t lc, uint64_t kraw,
                                              uint32_t craw, uint64_t* sk, uint32_t* sc) {
  typedef KL<NV> K;
  const int nB = (int)pp.f.len - 1;
  const uint32_t c = bbf_mulmod(F, lc, pp.f.invlc);
  pp.nc = F.p - c;                 // c != 0
  pp.adj = lead - pp.f.lm;
  pp.spar ^= 1;
  pp.staged = true;
  if (nB <= BBW_BSTAGE && (int)threadIdx.x < nB) {
    pp.kB = kraw + pp.adj; pp.cB = bbf_mulmod(F, craw, pp.nc);
    sk[pp.spar * BBW_BSTAGE + threadIdx.x] = pp.kB; sc[pp.spar * BBW_BSTAGE + threadIdx.x] = pp.cB;
    if (pp.kB & K::g_all) pp.bad = 1u;
  }
}

// One addition by the bitmap merge: O = A + (the staged reducer tail), A = nA terms in shared memory followed by
// sentinels, B = the nB = pp.f.len - 1 <= BBW_BSTAGE scaled tail terms of pp.f (this thread's in pp.kB / pp.cB, all of
// them in sk / sc[pp.spar]).  Also predicts O's lead term, searches its divisor (-> pp.lead / pp.lc / pp.found / pp.fidx;
// found = -2 when that has to wait for the merge) and loads, scales and stages that reducer (pp.staged).  Returns the
// number of terms of O (O visible to every thread, sentinels behind it), or -1 / -3 as block_merge_rank.
template <int NV>
__device__ __forceinline__ int block_merge_bits(WideShared& sh, int& slot, WidePipe& pp, const BBField F, const uint64_t* Ak,
                                                const uint32_t* Ac, const int nA, uint64_t* Ok, uint32_t* Oc, const int cap,
                                                uint64_t* sk, uint32_t* sc, const uint64_t* __restrict__ rlm,
                                                const uint32_t* __restrict__ ridx, const int nR, const bool sorted,
                                                const GHeadMem* gh, const uint64_t* __restrict__ tk,
                                                const uint32_t* __restrict__ tc, const int flags) {
  const int tid = threadIdx.x, lane = bb_lane(), warp = tid >> 5;
  BBW_T(0);   // since the last probe outside: loop overhead, synchronous staging
  const int nB = (int)pp.f.len - 1;
  const int total = nA + nB;
  const uint64_t* Bk = sk + pp.spar * BBW_BSTAGE;
  const uint32_t* Bc = sc + pp.spar * BBW_BSTAGE;
  const uint64_t b0 = nB > 0 ? pp.f.k1 + pp.adj : ~0ull;
  const uint32_t cb0 = nB > 0 ? bbf_mulmod(F, pp.f.c1, pp.nc) : 0u;
  pp.found = -2; pp.staged = false;
  if (total > cap) return -1;
  const int par = pp.par;
  WideBits& bm = sh.bits[par];
  const int nwords = (total + 31) >> 5;   // bitmap words this merge may set; word nwords stays clean
  const int oldw = par ? pp.used0 : pp.used1;
  if (par) { pp.used1 = nwords; pp.used0 = 0; } else { pp.used0 = nwords; pp.used1 = 0; }
  pp.par = par ^ 1;
  // ---- phase 1: B terms find their slots; the other warps predict the result's lead term and search its divisor
  const int nBw = (nB + 31) >> 5;
  if (tid < nB) {
    const uint64_t k = pp.kB;
    const int lo = rank4(Ak, nA, k);
    const bool coin = Ak[lo] == k;             // never a sentinel
    const int u = lo + tid + (coin ? 1 : 0);   // A first on ties
    atomicOr(&bm.ub[u >> 5], 1u << (u & 31));
    if (coin) atomicOr(&bm.db[u >> 5], 1u << (u & 31));
  }
  BBW_T(1);   // B terms ranked and marked
  // the lead term of the result: the smaller head; equal heads: their sum, unless it cancels
  bool predicted = total > 0;
  {
    const uint64_t a0 = Ak[0];                 // a sentinel when nA == 0
    const uint32_t ca = Ac[0];
    uint32_t lc = a0 < b0 ? ca : cb0;
    if (a0 == b0 && predicted) { lc = bbf_addmod(F, ca, cb0); predicted = lc != 0u; }
    pp.lead = a0 < b0 ? a0 : b0; pp.lc = lc;
  }
  if (predicted) search_publish<NV>(sh, slot, rlm, ridx, nR, pp.lead, sorted, nBw <= BBW_WARPS / 2 ? nBw * 32 : 0);
  BBW_T(2);   // prediction + divisor search
  __syncthreads();
  BBW_T(3);   // barrier 1
  // ---- phase 2: every merged slot is written to its place; the next reducer is fetched, scaled and staged
  if (predicted) {
    pp.found = search_collect(sh, slot, pp.fidx);
    slot ^= 1;
    if (pp.found >= 0) pp.f = load_head(gh + pp.fidx);   // in flight while the slots are written
  }
  {  // the other parity's bitmaps were last read before the barrier that ended the previous merge
    WideBits& o = sh.bits[par ^ 1];
#pragma unroll 1
    for (int w = tid; w < oldw; w += BBW_THREADS) { o.ub[w] = 0u; o.db[w] = 0u; }
  }
  BBW_T(4);   // collect, head load issued, bitmaps cleared
  int nout;
  uint32_t zero = 0u;
  if (nwords < 32) {   // up to 992 merged slots: lane w of every warp holds bitmap word w, prefix counts are one warp reduction
    const uint32_t myub = bm.ub[lane], mydb = bm.db[lane];
    const int pu = __popc(myub), pd = __popc(mydb);
    nout = total - (int)__reduce_add_sync(BB_FULL, pd);
#pragma unroll 1
    for (int u0 = warp * 32; u0 < total; u0 += BBW_THREADS) {
      const int w = u0 >> 5, u = u0 + lane;
      const int cbw = (int)__reduce_add_sync(BB_FULL, lane < w ? pu : 0), cdw = (int)__reduce_add_sync(BB_FULL, lane < w ? pd : 0);
      const uint32_t ubw = __shfl_sync(BB_FULL, myub, w), dbw = __shfl_sync(BB_FULL, mydb, w), dbn = __shfl_sync(BB_FULL, mydb, w + 1);
      const uint32_t below = (1u << lane) - 1u;
      const int j = cbw + __popc(ubw & below);          // B terms before this slot
      const int o = u - (cdw + __popc(dbw & below));    // its place once the dead slots are squeezed out
      if (u < total) {
        if (!((ubw >> lane) & 1u)) {                    // an A term; a dead slot right behind it is its coinciding B term
          const int ia = u - j;
          uint32_t c = Ac[ia];
          if (__funnelshift_rc(dbw, dbn, lane + 1) & 1u) { c = bbf_addmod(F, c, Bc[j]); zero |= c == 0u ? 2u : 0u; }
          Ok[o] = Ak[ia]; Oc[o] = c;
        } else if (!((dbw >> lane) & 1u)) {
          Ok[o] = Bk[j]; Oc[o] = Bc[j];
        }
      }
    }
  } else {             // longer: prefix counts by walking the bitmap words
    int wdone = 0, cb = 0, cd = 0;   // B / dead slots in bitmap words [0, wdone)
#pragma unroll 1
    for (int u = tid; u < total; u += BBW_THREADS) {
      const int w = u >> 5, b = u & 31;
      while (wdone < w) { cb += __popc(bm.ub[wdone]); cd += __popc(bm.db[wdone]); wdone++; }
      const uint32_t ubw = bm.ub[w], dbw = bm.db[w], dbn = bm.db[w + 1];
      const uint32_t below = (1u << b) - 1u;
      const int j = cb + __popc(ubw & below);
      const int o = u - (cd + __popc(dbw & below));
      if (!((ubw >> b) & 1u)) {
        const int ia = u - j;
        uint32_t c = Ac[ia];
        if (__funnelshift_rc(dbw, dbn, b + 1) & 1u) { c = bbf_addmod(F, c, Bc[j]); zero |= c == 0u ? 2u : 0u; }
        Ok[o] = Ak[ia]; Oc[o] = c;
      } else if (!((dbw >> b) & 1u)) {
        Ok[o] = Bk[j]; Oc[o] = Bc[j];
      }
    }
    int dead = 0;
#pragma unroll 1
    for (int w = 0; w < nwords; w++) dead += __popc(bm.db[w]);
    nout = total - dead;
  }
  if (tid < BBW_SENT) Ok[nout + tid] = ~0ull;
  BBW_T(5);   // slots written
  const uint32_t mybad = pp.bad | zero;
  pp.bad = 0u;
  if (pp.found >= 0) {   // the next reducer: its tail one term per thread, scaled and staged for the next addition
    const int nb2 = (int)pp.f.len - 1;
    uint64_t kraw = 0ull; uint32_t craw = 0u;
    if (nb2 <= BBW_BSTAGE && tid < nb2) { kraw = tk[pp.f.off + 1 + tid]; craw = tc[pp.f.off + 1 + tid]; }
    stage_reducer<NV>(pp, F, pp.lead, pp.lc, kraw, craw, sk, sc);
  }
  BBW_T(6);   // next reducer staged
  const int trouble = __syncthreads_or(mybad != 0u);
  BBW_T(7);   // barrier 2
  if (trouble || (flags & BBW_FLAG_COMPACT)) {   // rare: sort out what happened through shared memory
    if (tid == 0) sh.flags = 0u;
    __syncthreads();
    if (mybad) atomicOr(&sh.flags, mybad);
    __syncthreads();
    const uint32_t fl = sh.flags;
    if (fl & 1u) return -3;
    // a sum cancelled: squeeze the zero out; if it was the lead term the search starts over
    const uint32_t c0 = nout > 0 ? Oc[0] : 1u;
    const int n2 = block_compact_zeros(sh, Ok, Oc, nout);
    if (c0 == 0u) { pp.found = -2; pp.staged = false; }
    return n2;
  }
  return nout;
}

// Removes row `row` from the pair list keeping order (buchberger.cpp:319), by the whole block: chunk by chunk, read, barrier,
// write one place lower.  The pair and its cached lcm key go to (pr, gam).  e.nP is NOT changed here.
__device__ __forceinline__ void block_take_pair(const BBParams& P, const Env& e, int row, uint32_t& pr, uint64_t& gam) {
  uint32_t* pairs = ENV_PTR(uint32_t, e, P, o_pairs);
  uint64_t* plcm = ENV_PTR(uint64_t, e, P, o_plcm);
  pr = pairs[row];
  gam = plcm[row];
#pragma unroll 1
  for (int b0 = row; b0 < e.nP - 1; b0 += BBW_THREADS) {
    const int idx = b0 + threadIdx.x;
    const bool v = idx < e.nP - 1;
    uint32_t x = 0u; uint64_t y = 0ull;
    if (v) { x = pairs[idx + 1]; y = plcm[idx + 1]; }
    __syncthreads();   // every read of this chunk (and of (pr, gam) in the first) before any write
    if (v) { pairs[idx] = x; plcm[idx] = y; }
  }
}

// One environment step by the whole CTA (BuchbergerEnv::step, buchberger.cpp:318-329, with the pair chosen by
// `strategy`).  e and every scalar below are block-uniform.  hk / hc: the two shared dividend buffers of `cap` terms
// (+ BBW_SENT sentinels each, stride cap + BBW_SENT); sk / sc: two staging buffers of BBW_BSTAGE terms for a reducer's
// scaled tail.  Returns the number of polynomial additions; `pair` receives (j << 16) | i.
template <int NV>
__device__ __forceinline__ int block_step(const BBParams& P, Env& e, WideShared& sh, int& slot, WidePipe& pp, uint64_t* hk,
                                          uint32_t* hc, int cap, uint64_t* sk, uint32_t* sc, int flags, int strategy,
                                          uint32_t* sel_rng, uint32_t& pair, Ctr& ct) {
  typedef KL<NV> K;
  const BBField F = P.F;
  const int tid = threadIdx.x;
  const int bstride = cap + BBW_SENT;
  BBW_T(8);   // end of the previous step's bookkeeping
  if (tid < 32) {   // warp 0: select the pair
    const int row = warp_select<NV>(P, e, strategy, sel_rng);
    if (tid == 0) sh.row = row;
  }
  __syncthreads();
  uint32_t pr; uint64_t gam;
  block_take_pair(P, e, sh.row, pr, gam);
  e.nP--;
  pair = pr;
  const GHeadMem* gh = ENV_PTR(GHeadMem, e, P, o_ghead);
  const uint64_t* tk = ENV_PTR(uint64_t, e, P, o_tkey);
  const uint32_t* tc = ENV_PTR(uint32_t, e, P, o_tcoef);
  const uint64_t* rlm = ENV_PTR(uint64_t, e, P, o_rlm);
  const uint32_t* ridx = ENV_PTR(uint32_t, e, P, o_ridx);
  const int nR = e.nG;
  const bool sorted = P.sort_reducers != 0;
  const GHead hf = load_head(gh + (pr & 0xffffu)), hg = load_head(gh + (pr >> 16));
  e.guard |= gam;
  ct.tread += hf.len + hg.len;
  int sug;
  {  // sugar of the S-polynomial: max(deg(gamma / LM f) + sug f, deg(gamma / LM g) + sug g)
    const int cg0 = (int)(uint32_t)(gam >> K::dshift);
    const int sf = (int)hf.sug + (int)(uint32_t)(hf.lm >> K::dshift) - cg0, sg = (int)hg.sug + (int)(uint32_t)(hg.lm >> K::dshift) - cg0;
    sug = sf > sg ? sf : sg;
  }
  if (e.guard & K::g_all) { e.status = BB_STATUS_OVERFLOW_EXPONENT; return 1; }
  // s = (gamma / LT f) tail(f) - (gamma / LT g) tail(g): the lead terms cancel exactly (buchberger.cpp:18-21).
  // tail(f), scaled, is staged as the "dividend" A in buffer 1; the g side is then an ordinary addition into buffer 0.
  int cur = 0, pos = 0, n;
  const int nAf = (int)hf.len - 1, nBg = (int)hg.len - 1;
  if (nAf + nBg > cap) { e.status = BB_STATUS_OVERFLOW_SCRATCH; return 1; }
  pp.found = -2; pp.staged = false; pp.bad = 0u;
  if (nBg > BBW_BSTAGE || (flags & BBW_FLAG_RANK_MERGE)) {
    n = block_merge_rank<NV>(sh, F, tk + hf.off + 1, tc + hf.off + 1, nAf, hf.invlc, gam - hf.lm, tk + hg.off + 1,
                             tc + hg.off + 1, nBg, F.p - hg.invlc, gam - hg.lm, hk, hc, cap);
  } else {
    const uint64_t adjf = gam - hf.lm;
    uint64_t guard = 0;
#pragma unroll 1
    for (int t = tid; t < nAf; t += BBW_THREADS) {
      const uint64_t k = tk[hf.off + 1 + t] + adjf;
      guard |= k;
      hk[bstride + t] = k; hc[bstride + t] = bbf_mulmod(F, tc[hf.off + 1 + t], hf.invlc);
    }
    if (tid < BBW_SENT) hk[bstride + nAf + tid] = ~0ull;
    // the g side as a reducer of the "lead term" gamma with coefficient 1 / LC f ... i.e. nc = -(1 / LC g), adj = gamma - LM g
    uint64_t kraw = 0ull; uint32_t craw = 0u;
    if (tid < nBg) { kraw = tk[hg.off + 1 + tid]; craw = tc[hg.off + 1 + tid]; }
    pp.f = hg;
    stage_reducer<NV>(pp, F, gam, 1u, kraw, craw, sk, sc);   // lc = 1: c = 1 / LC g, nc = -(1 / LC g)
    if (__syncthreads_or((guard & K::g_all) != 0ull)) { e.status = BB_STATUS_OVERFLOW_EXPONENT; return 1; }
    BBW_T(9);   // select, take pair, S-polynomial staging
    n = block_merge_bits<NV>(sh, slot, pp, F, hk + bstride, hc + bstride, nAf, hk, hc, cap, sk, sc, rlm, ridx, nR, sorted, gh, tk,
                             tc, flags);
  }
  if (n < 0) { e.status = n == -1 ? BB_STATUS_OVERFLOW_SCRATCH : BB_STATUS_OVERFLOW_EXPONENT; return 1; }
  ct.twrite += (unsigned)n;
  // reduce(s, G_), buchberger.cpp:24-49
  uint64_t* rk = ENV_PTR(uint64_t, e, P, o_tkey) + e.nT;
  uint32_t* rc = ENV_PTR(uint32_t, e, P, o_tcoef) + e.nT;
  const int rcap = P.max_terms - e.nT;
  int steps = 0, rlen = 0;
#pragma unroll 1
  while (n > 0) {
    const int hb = cur * bstride + pos;
    uint64_t lead = pp.lead; uint32_t lc = pp.lc;
    if (pp.found == -2) {
      BBW_T(10);   // loop overhead before an unpredicted search
      lead = hk[hb]; lc = hc[hb];
      pp.found = block_first_divisor<NV>(sh, slot, rlm, ridx, nR, lead, sorted, pp.fidx); pp.staged = false;
      BBW_T(11);   // unpredicted divisor search (after a term move or a cancelled lead)
    }
    ct.lms += (pp.found >= 0) ? (unsigned)(pp.found + 1) : (unsigned)nR;
    if (pp.found >= 0) {
      if (!pp.staged) {   // not fetched ahead: load the reducer now
        pp.f = load_head(gh + pp.fidx);
        const int nb = (int)pp.f.len - 1;
        uint64_t kraw = 0ull; uint32_t craw = 0u;
        if (nb <= BBW_BSTAGE && tid < nb) { kraw = tk[pp.f.off + 1 + tid]; craw = tc[pp.f.off + 1 + tid]; }
        stage_reducer<NV>(pp, F, lead, lc, kraw, craw, sk, sc);
      }
      const int nB = (int)pp.f.len - 1;
      const int sf = (int)pp.f.sug + (int)(uint32_t)(pp.f.lm >> K::dshift) - (int)(uint32_t)(lead >> K::dshift);
      sug = sf > sug ? sf : sug;
      ct.tread += (unsigned)n + pp.f.len;
      const int ob = cur ^ 1;
      int n2;
      if (nB > BBW_BSTAGE || (flags & BBW_FLAG_RANK_MERGE)) {
        n2 = block_merge_rank<NV>(sh, F, hk + hb + 1, hc + hb + 1, n - 1, 1u, 0ull, tk + pp.f.off + 1, tc + pp.f.off + 1, nB, pp.nc,
                                  pp.adj, hk + ob * bstride, hc + ob * bstride, cap);
        pp.found = -2; pp.staged = false;
      } else {
        n2 = block_merge_bits<NV>(sh, slot, pp, F, hk + hb + 1, hc + hb + 1, n - 1, hk + ob * bstride, hc + ob * bstride, cap, sk,
                                  sc, rlm, ridx, nR, sorted, gh, tk, tc, flags);
      }
      if (n2 < 0) { e.status = n2 == -1 ? BB_STATUS_OVERFLOW_SCRATCH : BB_STATUS_OVERFLOW_EXPONENT; return 1 + steps; }
      n = n2; cur = ob; pos = 0;
      ct.twrite += (unsigned)n;
      steps++;
    } else {
      if (rlen >= rcap) { e.status = BB_STATUS_OVERFLOW_TERMS; return 1 + steps; }
      if (tid == 0) { rk[rlen] = lead; rc[rlen] = lc; }
      rlen++; ct.moves++;
      pos++; n--;
      pp.found = -2; pp.staged = false;
    }
  }
  pp.found = -2; pp.staged = false;
  BBW_T(12);   // loop exit
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
    BBW_T(13);   // update() by warp 0
  }
  if (e.nP == 0) e.status = BB_STATUS_DONE;
  return 1 + steps;
}
