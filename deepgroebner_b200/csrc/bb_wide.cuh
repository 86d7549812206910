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
#define BBW_WARPS 4
#endif
#define BBW_THREADS (BBW_WARPS * 32)
#ifndef BBW_MIN_CTAS
#define BBW_MIN_CTAS 7   // 1024 environments resident on 148 SMs
#endif

struct WideShared {
  int wmin[2][BBW_WARPS];   // block_scan: per-warp minima (double-buffered, see block_first_divisor)
  int wcnt[2][BBW_WARPS];   // block_merge: per-warp survivor counts (double-buffered)
  uint32_t pr; int row;     // the pair warp 0 took this step
  uint64_t gam;
  long long upd;            // result of warp_add_basis
  int status;
};

// First reducer (lowest position in G_) whose lead monomial divides `lead`, or -1.  `slot` alternates between calls:
// a slot is rewritten two calls later, after every thread has passed the barrier of the call in between.
template <int NV>
__device__ __forceinline__ int block_first_divisor(WideShared& sh, int& slot, const uint64_t* __restrict__ rlm, int nR,
                                                   uint64_t lead) {
  typedef KL<NV> K;
  int best = 0x7fffffff;
#pragma unroll 1
  for (int r = threadIdx.x; r < nR; r += BBW_THREADS)
    if (K::divides(rlm[r], lead)) { best = r; break; }   // ascending per thread: its first hit is its minimum
  best = (int)__reduce_min_sync(BB_FULL, (unsigned)best);
  if (bb_lane() == 0) sh.wmin[slot][threadIdx.x >> 5] = best;
  __syncthreads();
  int found = sh.wmin[slot][0];
#pragma unroll
  for (int w = 1; w < BBW_WARPS; w++) found = min(found, sh.wmin[slot][w]);
  slot ^= 1;
  return found == 0x7fffffff ? -1 : found;
}

// out = cA * mA * A + cB * mB * B for term lists in ascending key order (see warp_merge in bb_device.cuh for the
// conventions: adjX = key(mX) - bias, coefficients in [1,p), cancelled terms dropped).  A, B, O may be shared or global
// memory; O must not alias A or B and needs room for nA + nB terms (`cap`).  Returns the number of output terms, -1 if
// nA + nB > cap, -3 if a produced key has a guard bit set.  Ends with a block barrier: O is visible to every thread.
template <int NV>
__device__ __forceinline__ int block_merge(WideShared& sh, int& slot, BBField F, const uint64_t* Ak, const uint32_t* Ac,
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
    int before = 0, all = 0;
#pragma unroll
    for (int w = 0; w < BBW_WARPS; w++) { const int n = sh.wcnt[slot][w]; all += n; before += w < warp ? n : 0; }
    if (c != 0u) { const int pos = no + before + __popc(live & ltm); Ok[pos] = k; Oc[pos] = c; }
    no += all;
    slot ^= 1;
  }
  __syncthreads();
  return bad ? -3 : no;
}

// One environment step by the whole CTA (BuchbergerEnv::step, buchberger.cpp:318-329, with the pair chosen by
// `strategy`).  e and every scalar below are block-uniform.  hk / hc: the two shared dividend buffers of `cap` terms.
// Returns the number of polynomial additions; `pair` receives (j << 16) | i.
template <int NV>
__device__ __forceinline__ int block_step(const BBParams& P, Env& e, WideShared& sh, int& slot, uint64_t* hk, uint32_t* hc,
                                          int cap, int strategy, uint32_t* sel_rng, uint32_t& pair, Ctr& ct) {
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
  int n = block_merge<NV>(sh, slot, F, tk + hf.off + 1, tc + hf.off + 1, (int)hf.len - 1, hf.invlc, gam - hf.lm,
                          tk + hg.off + 1, tc + hg.off + 1, (int)hg.len - 1, F.p - hg.invlc, gam - hg.lm, hk, hc, cap);
  if (n < 0) { e.status = n == -1 ? BB_STATUS_OVERFLOW_SCRATCH : BB_STATUS_OVERFLOW_EXPONENT; return 1; }
  if (e.guard & K::g_all) { e.status = BB_STATUS_OVERFLOW_EXPONENT; return 1; }
  ct.twrite += (unsigned)n;
  // reduce(s, G_), buchberger.cpp:24-49
  const uint64_t* rlm = ENV_PTR(uint64_t, e, P, o_rlm);
  const uint32_t* ridx = ENV_PTR(uint32_t, e, P, o_ridx);
  uint64_t* rk = ENV_PTR(uint64_t, e, P, o_tkey) + e.nT;
  uint32_t* rc = ENV_PTR(uint32_t, e, P, o_tcoef) + e.nT;
  const int rcap = P.max_terms - e.nT, nR = e.nG;
  int steps = 0, rlen = 0;
  while (n > 0) {
    const uint64_t* ck = hk + (size_t)cur * cap + pos;
    const uint32_t* cc = hc + (size_t)cur * cap + pos;
    const uint64_t lead = ck[0];
    const int found = block_first_divisor<NV>(sh, slot, rlm, nR, lead);
    ct.lms += (found >= 0) ? (unsigned)(found + 1) : (unsigned)nR;
    if (found >= 0) {
      const GHead f = load_head(gh + ridx[found]);
      const uint32_t c = bbf_mulmod(F, cc[0], f.invlc);
      const uint32_t nc = F.p - c;              // c != 0
      const uint64_t adj = lead - f.lm;         // key(LM h / LM f) - bias
      const int sf = (int)f.sug + (int)(uint32_t)(f.lm >> K::dshift) - (int)(uint32_t)(lead >> K::dshift);
      sug = sf > sug ? sf : sug;
      ct.tread += (unsigned)n + f.len;
      const int ob = cur ^ 1;
      const int n2 = block_merge<NV>(sh, slot, F, ck + 1, cc + 1, n - 1, 1u, 0ull, tk + f.off + 1, tc + f.off + 1,
                                     (int)f.len - 1, nc, adj, hk + (size_t)ob * cap, hc + (size_t)ob * cap, cap);
      if (n2 < 0) { e.status = n2 == -1 ? BB_STATUS_OVERFLOW_SCRATCH : BB_STATUS_OVERFLOW_EXPONENT; return 1 + steps; }
      n = n2; cur = ob; pos = 0;
      ct.twrite += (unsigned)n;
      steps++;
    } else {
      if (rlen >= rcap) { e.status = BB_STATUS_OVERFLOW_TERMS; return 1 + steps; }
      if (tid == 0) { rk[rlen] = lead; rc[rlen] = cc[0]; }
      rlen++; ct.moves++;
      pos++; n--;
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
