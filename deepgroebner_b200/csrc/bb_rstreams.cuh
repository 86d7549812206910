// bb_rstreams.cuh -- the stream reducer of bb_streams.cuh with every stream in REGISTERS, one warp per environment, sm_100a.
//
// What bounds a cyclic-6 launch is the latency of one ROUND of the longest episode (one lead term of the dividend: divisor
// search, advance of the streams at it, next lead term; 4.6 rounds per addition, 697 000 dependent additions in the longest
// of 1024 seeded-Random episodes).  Measured on B200 (tools/ub/lat.cu, cycles per dependent operation): REDUX.MIN 22,
// REDUX.ADD 47, the 64-bit minimum + coefficient sum by three REDUX 108, the same by a shuffle butterfly 222, LDS 34,
// STS + BAR + LDS 58 (8 warps).  The CTA-per-environment round (bb_wide.cuh) folds twice with a barrier in between
// (108 + 58 + 108 cycles before anything else); a single warp folds once -- provided the per-lane part stays short.
// The first warp version (bb_streams.cuh) did not: streams were appended, exhausted ones stayed in the table until the
// next garbage collection, and everything beyond 32 streams was a loop over shared memory (~380 dependent instructions
// per round).  Here
//   * a stream lives in slot (row r, lane l) of BBR_ROWS x 32 register slots; a new stream takes the FIRST FREE slot
//     (one ballot), so the table is as small as the number of LIVE streams and rows above the highest live one are
//     skipped by a warp-uniform branch: no garbage collection, no shared memory, no barrier;
//   * the first 128 reducer lead monomials of G_ (and their basis indices) sit in registers for the whole reduction:
//     the divisor search of a round is four subtract-and-mask tests and one REDUX, off the fold's chain;
//   * one round per loop iteration with ONE call site: the loop starts from the pseudo lead term (all ones, 0), which
//     consumes nothing, and a cancelled monomial (S = 0) simply ignores its search result.
// If every slot holds a live stream, h is consolidated into the scratch list behind the slot's term arena and goes on as
// one stream (streams_consolidate's scheme; bb_set_wide(5) / (6) cap the live streams at 6 / 48 so that the tests reach
// it).  Results are those of the reference term for term: same lead-term sequence, same divisor per lead term (first in
// G_), same number of additions, same remainder; only the order in which coefficients at one monomial are summed differs.
#pragma once
#include "bb_streams.cuh"

#ifndef BBR_ROWS
#define BBR_ROWS 4               // register slots per lane: BBR_ROWS * 32 live streams before a consolidation
#endif
#define BBR_REG_REDUCERS 128     // reducer lead monomials of G_ held in registers (4 per lane)

struct RegStreams {
  // slot (r, lane): head key (all ones: free), key(multiplier) - bias, raw key of the term behind the head
  uint64_t k[BBR_ROWS], adj[BBR_ROWS], pk[BBR_ROWS];
  // head coefficient (scaled), multiplier coefficient, raw coefficient behind the head, index behind the head, end
  uint32_t c[BBR_ROWS], nc[BBR_ROWS], pc[BBR_ROWS], p[BBR_ROWS], e[BBR_ROWS];
  uint64_t rl[4];          // reducer lead monomials lane, lane + 32, lane + 64, lane + 96 of G_ (all ones: absent)
  uint32_t rc[4];          // (position in G_ << 16) | basis index; BBS_NONE: absent
  int nrows;               // warp-uniform: rows r >= nrows hold no live stream
  int kmax;                // live streams allowed (BBR_ROWS * 32, or less under bb_set_wide(5 / 6))
  int cz;                  // scratch half the next consolidation writes
  uint32_t bad;            // per lane: a produced key overflowed its exponent fields
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int r = 0; r < BBR_ROWS; r++) { k[r] = ~0ull; adj[r] = pk[r] = 0ull; c[r] = nc[r] = pc[r] = p[r] = e[r] = 0u; }
    nrows = 0; bad = 0u;
  }
};

// One round: the streams whose head is M advance (M = all ones: nothing to consume), and the next lead term of h comes
// back as (M2, S2): M2 all ones when h is exhausted, S2 in [0, p) (0: the monomial cancelled).  SEARCH: also M's first
// divisor in G_, best = (position << 16) | basis index or BBS_NONE (`want` false: the caller will not look at it, so the
// part of G_ beyond the registers is not scanned).
template <int NV, bool SEARCH>
__device__ __forceinline__ void rs_round(RegStreams& ws, const BBField F, const uint64_t M, const bool want, const uint64_t* rlm,
                                         const uint32_t* ridx, const int nR, const bool sorted, const uint64_t* tk,
                                         const uint32_t* tc, uint64_t& M2, uint32_t& S2, uint32_t& best) {
  typedef KL<NV> K;
  // (a) divisor search in the register-resident part of G_
  uint32_t cand = BBS_NONE;
  if (SEARCH) {
    const uint64_t mg = (M & K::ex_mask) | K::ge_mask;
#pragma unroll
    for (int q = 3; q >= 0; q--)
      if (((mg - (ws.rl[q] & K::ex_mask)) & K::ge_mask) == K::ge_mask) cand = ws.rc[q];   // absent: rc = BBS_NONE
  }
  // (b) this lane's slots: advance the heads at M, minimum head and the coefficient sum at it
  uint64_t mk = ~0ull;
  uint32_t ms = 0u;
#pragma unroll
  for (int r = 0; r < BBR_ROWS; r++) {
    if (r < ws.nrows) {
      uint64_t k = ws.k[r];
      if (k == M) {
        if (ws.p[r] < ws.e[r]) {
          k = ws.pk[r] + ws.adj[r];
          ws.c[r] = bbf_mulmod(F, ws.pc[r], ws.nc[r]);
          if (k & K::g_all) ws.bad = 1u;
          ws.p[r]++;
          if (ws.p[r] < ws.e[r]) { ws.pk[r] = tk[ws.p[r]]; ws.pc[r] = tc[ws.p[r]]; }   // needed a round later at the earliest
        } else {
          k = ~0ull;
        }
        ws.k[r] = k;
      }
      const uint32_t c = ws.c[r];
      if (k < mk) { mk = k; ms = c; } else if (k == mk) ms += c;   // at most BBR_ROWS values below 2^16
    }
  }
  // (c) 64-bit minimum through two 32-bit reductions, the coefficient sum at it
  const uint32_t hi = __reduce_min_sync(BB_FULL, (uint32_t)(mk >> 32));
  const uint32_t lo = __reduce_min_sync(BB_FULL, (uint32_t)(mk >> 32) == hi ? (uint32_t)mk : 0xffffffffu);
  const uint32_t wsum = __reduce_add_sync(BB_FULL, ((uint32_t)(mk >> 32) == hi && (uint32_t)mk == lo) ? ms : 0u);
  M2 = ((uint64_t)hi << 32) | lo; S2 = bbf_reduce(F, wsum);   // <= BBR_ROWS * 32 values below 2^16
  best = BBS_NONE;
  if (SEARCH) {
    best = __reduce_min_sync(BB_FULL, cand);
    if (best == BBS_NONE && nR > BBR_REG_REDUCERS && want) {   // the rest of G_ from memory, 128 per pass
      // sorted: G_ ascends in lead monomial (keys descend): a reducer whose key is below M's cannot divide, nor any after it
      const uint64_t stop = sorted ? M : 0ull;
      bool over = __any_sync(BB_FULL, ws.rl[0] < stop || ws.rl[1] < stop || ws.rl[2] < stop || ws.rl[3] < stop);
      const uint64_t mg = (M & K::ex_mask) | K::ge_mask;
      const int lane = bb_lane();
#pragma unroll 1
      for (int base = BBR_REG_REDUCERS; base < nR && !over; base += 128) {
        const int r0 = base + lane, r1 = r0 + 32, r2 = r0 + 64, r3 = r0 + 96;
        const bool v0 = r0 < nR, v1 = r1 < nR, v2 = r2 < nR, v3 = r3 < nR;
        const uint64_t l0 = v0 ? rlm[r0] : ~0ull, l1 = v1 ? rlm[r1] : ~0ull, l2 = v2 ? rlm[r2] : ~0ull, l3 = v3 ? rlm[r3] : ~0ull;
        const bool h0 = v0 && ((mg - (l0 & K::ex_mask)) & K::ge_mask) == K::ge_mask, h1 = v1 && ((mg - (l1 & K::ex_mask)) & K::ge_mask) == K::ge_mask;
        const bool h2 = v2 && ((mg - (l2 & K::ex_mask)) & K::ge_mask) == K::ge_mask, h3 = v3 && ((mg - (l3 & K::ex_mask)) & K::ge_mask) == K::ge_mask;
        const uint32_t c2 = h0 ? (uint32_t)r0 : (h1 ? (uint32_t)r1 : (h2 ? (uint32_t)r2 : (h3 ? (uint32_t)r3 : BBS_NONE)));
        const uint32_t b2 = __reduce_min_sync(BB_FULL, c2);
        if (b2 != BBS_NONE) { best = (b2 << 16) | ridx[b2]; break; }
        over = __any_sync(BB_FULL, l0 < stop || l1 < stop || l2 < stop || l3 < stop);
      }
    }
  }
}

// Opens a stream in the first free slot below kmax: head (hk, hc) already scaled, multiplier (adj, nc), the terms behind
// the head at [next, end).  False if every slot holds a live stream.
__device__ __forceinline__ bool rs_open(RegStreams& ws, uint64_t hk, uint32_t hc, uint64_t adj, uint32_t nc, uint32_t next,
                                        uint32_t end, const uint64_t* tk, const uint32_t* tc) {
  const int lane = bb_lane();
  // every access names its row by a constant: an index computed at run time would move the slots to local memory
  bool done = false;
#pragma unroll
  for (int r = 0; r < BBR_ROWS; r++) {
    if (!done && r * 32 < ws.kmax) {
      const uint32_t m = __ballot_sync(BB_FULL, ws.k[r] == ~0ull && r * 32 + lane < ws.kmax);
      if (m) {
        done = true;
        if (lane == __ffs((int)m) - 1) {
          ws.k[r] = hk; ws.c[r] = hc; ws.adj[r] = adj; ws.nc[r] = nc; ws.p[r] = next; ws.e[r] = end;
          if (next < end) { ws.pk[r] = tk[next]; ws.pc[r] = tc[next]; }
        }
        if (r >= ws.nrows) ws.nrows = r + 1;
      }
    }
  }
  if (!done) return false;
  // did the highest row drain meanwhile?
#pragma unroll
  for (int r = BBR_ROWS - 1; r >= 1; r--)
    if (ws.nrows == r + 1 && __ballot_sync(BB_FULL, ws.k[r] != ~0ull) == 0u) ws.nrows = r;
  return true;
}

// Consolidation: every pending term of h, from the lead term (M, S) on, is written in order to scratch half ws.cz and the
// streams are replaced by ONE stream over that list.  (M, S) becomes its head (the first term with a nonzero sum), or
// M = all ones if nothing is left.  sbase: term index of the scratch (= max_terms, see bb_streams.cuh).  Returns the number
// of terms, or -1 if they do not fit `cap`.
template <int NV>
__device__ __forceinline__ int rs_consolidate(RegStreams& ws, const BBField F, uint64_t& M, uint32_t& S, uint64_t* tk, uint32_t* tc,
                                           uint32_t sbase, int cap) {
  const uint32_t base = sbase + (uint32_t)(ws.cz * cap);
  int t = 0;
  uint64_t m = M, fm = ~0ull; uint32_t s = S, fs = 0u;
  while (m != ~0ull) {
    if (s != 0u) {
      if (t >= cap) return -1;
      if (t == 0) { fm = m; fs = s; }
      if (bb_lane() == 0) { tk[base + t] = m; tc[base + t] = s; }
      t++;
    }
    uint32_t best;
    rs_round<NV, false>(ws, F, m, false, nullptr, nullptr, 0, false, tk, tc, m, s, best);
  }
  ws.nrows = 0;     // every stream is exhausted: every slot is free
  ws.cz ^= 1;
  M = fm; S = fs;
  __syncwarp();     // lane 0's list before the new stream's owner reads it back
  if (t > 0) rs_open(ws, fm, fs, 0ull, 1u, base + 1u, base + (uint32_t)t, tk, tc);
  return t;
}

// reduce(spoly(G[i], G[j]), G_) (buchberger.cpp:18-49) for the pair heads (hf, hg) and gamma = the key of the pair's lcm,
// with the dividend as a set of register-resident streams.  The remainder goes to (rk, rc) [cap rcap]; returns its length
// or a negative BB_STATUS_* on a fault; `steps` = reductions, `sug` = the sugar of the result (polynomials.cpp:150, 198).
template <int NV>
__device__ __forceinline__ int warp_reduce_rstreams(const BBParams& P, const Env& e, RegStreams& ws, const GHead hf, const GHead hg,
                                                    const uint64_t gam, int& sug, int& steps, uint64_t* rk, uint32_t* rc,
                                                    int rcap, Ctr& ct) {
  typedef KL<NV> K;
  const BBField F = P.F;
  const GHeadMem* gh = ENV_PTR(GHeadMem, e, P, o_ghead);
  uint64_t* tk = ENV_PTR(uint64_t, e, P, o_tkey);      // the term arena and, from index max_terms on, the consolidation
  uint32_t* tc = ENV_PTR(uint32_t, e, P, o_tcoef);     // scratch (two halves of max_poly_terms terms)
  const uint64_t* rlm = ENV_PTR(uint64_t, e, P, o_rlm);
  const uint32_t* ridx = ENV_PTR(uint32_t, e, P, o_ridx);
  const int nR = e.nG;
  const bool sorted = P.sort_reducers != 0;
  const int lane = bb_lane();
  int rlen = 0;
  steps = 0;
  ws.clear();
#pragma unroll
  for (int q = 0; q < 4; q++) {   // G_ changes only between reductions
    const int r = lane + 32 * q;
    ws.rl[q] = r < nR ? rlm[r] : ~0ull;
    ws.rc[q] = r < nR ? (((uint32_t)r << 16) | ridx[r]) : BBS_NONE;
  }
  // s = (gamma / LT f) tail(f) - (gamma / LT g) tail(g): the lead terms cancel exactly (buchberger.cpp:18-21); two streams
  // whose heads come from the head records
  if (hf.len > 1u) {
    const uint64_t adj = gam - hf.lm, k = hf.k1 + adj;
    if (k & K::g_all) return -BB_STATUS_OVERFLOW_EXPONENT;
    rs_open(ws, k, bbf_mulmod(F, hf.c1, hf.invlc), adj, hf.invlc, hf.off + 2u, hf.off + hf.len, tk, tc);
  }
  if (hg.len > 1u) {
    const uint64_t adj = gam - hg.lm, k = hg.k1 + adj;
    const uint32_t nc = F.p - hg.invlc;
    if (k & K::g_all) return -BB_STATUS_OVERFLOW_EXPONENT;
    rs_open(ws, k, bbf_mulmod(F, hg.c1, nc), adj, nc, hg.off + 2u, hg.off + hg.len, tk, tc);
  }
  uint64_t M = ~0ull; uint32_t S = 0u;   // pseudo lead term: consumes nothing, is no term of h
#pragma unroll 1
  do {
    uint64_t M2; uint32_t S2, best;
    rs_round<NV, true>(ws, F, M, S != 0u, rlm, ridx, nR, sorted, tk, tc, M2, S2, best);
    if (S != 0u) {   // S == 0: the monomial cancelled, it is not a term of h
      ct.lms += (best != BBS_NONE) ? (best >> 16) + 1u : (unsigned)nR;
      if (best != BBS_NONE) {   // h <- h - (LT h / LT f) f: the lead terms cancel, f's tail becomes a stream
        const GHead f = load_head(gh + (best & 0xffffu));
        const uint32_t c = bbf_mulmod(F, S, f.invlc);
        const uint32_t nc = F.p - c;              // c != 0
        const uint64_t adj = M - f.lm;            // key(LM h / LM f) - bias
        const int sf = (int)f.sug + (int)(uint32_t)(f.lm >> K::dshift) - (int)(uint32_t)(M >> K::dshift);
        sug = sf > sug ? sf : sug;
        ct.tread += f.len;
        steps++;
        if (f.len > 1u) {
          const uint64_t k = f.k1 + adj;
          const uint32_t ck = bbf_mulmod(F, f.c1, nc);
          if (k & K::g_all) return -BB_STATUS_OVERFLOW_EXPONENT;
          if (!rs_open(ws, k, ck, adj, nc, f.off + 2u, f.off + f.len, tk, tc)) {
            // (inlined: a call would take ws's address and move every slot from registers to local memory)
            if (rs_consolidate<NV>(ws, F, M2, S2, tk, tc, (uint32_t)P.max_terms, P.max_poly_terms) < 0)
              return -BB_STATUS_OVERFLOW_SCRATCH;
            rs_open(ws, k, ck, adj, nc, f.off + 2u, f.off + f.len, tk, tc);
          }
          if (k < M2) { M2 = k; S2 = ck; } else if (k == M2) S2 = bbf_addmod(F, S2, ck);
        }
      } else {            // no divisor: the lead term moves to the remainder
        if (rlen >= rcap) return -BB_STATUS_OVERFLOW_TERMS;
        if (lane == 0) { rk[rlen] = M; rc[rlen] = S; }
        rlen++; ct.moves++;
      }
    }
    M = M2; S = S2;
  } while (M != ~0ull);
  if (__any_sync(BB_FULL, ws.bad != 0u)) return -BB_STATUS_OVERFLOW_EXPONENT;
  __syncwarp();   // lane 0's remainder before every lane reads it (update(), hashes)
  return rlen;
}

// BuchbergerEnv::step for the pair in row `row` of P, as warp_step (bb_device.cuh) with reduce() by register streams.
template <int NV>
__device__ __forceinline__ int warp_step_rstreams(const BBParams& P, Env& e, RegStreams& ws, int row, uint32_t& pair, Ctr& ct) {
  typedef KL<NV> K;
  if ((unsigned)row >= (unsigned)e.nP) { e.status = BB_STATUS_BAD_ACTION; pair = 0xffffffffu; return 0; }
  uint32_t pr; uint64_t gam;
  warp_take_pair(P, e, row, pr, gam);
  pair = pr;
  const GHeadMem* gh = ENV_PTR(GHeadMem, e, P, o_ghead);
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
  int steps = 0;
  const int rlen = warp_reduce_rstreams<NV>(P, e, ws, hf, hg, gam, sug, steps, ENV_PTR(uint64_t, e, P, o_tkey) + e.nT,
                                            ENV_PTR(uint32_t, e, P, o_tcoef) + e.nT, P.max_terms - e.nT, ct);
  if (rlen < 0) { e.status = -rlen; return 1 + steps; }
  if (rlen > 0) {
    ct.upb += (unsigned)e.nG; ct.upp += (unsigned)e.nP;
    const long long r = warp_add_basis<NV>(P, e.base, e.nG, e.nP, e.nT, rlen, sug);
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
