// bb_streams.cuh -- reduce() for LONG polynomials (cyclic-n and other non-binomial ideals, BASELINE configs[4]) without
// ever materialising the dividend, one warp per environment, sm_100a.
//
// warp_reduce / warp_merge (bb_device.cuh) are built for binomial ideals, where a polynomial is two terms.  On cyclic-6
//  a dividend of ~100 (up to ~550) terms is reduced ~83 times per step by polynomials of ~33 terms, and an
// episode is a serial chain of ~80 000 such additions (697 000 in the longest of 1024): materialising
// h <- h - (LT h / LT f) f (buchberger.cpp:33-35) costs O(|h| + |f|) work per addition and, whatever the parallel merge,
// several dependent passes over memory.  Round 2 built and measured three block-wide merges for it (merge path, a
// two-barrier bitmap merge, streams with one barrier per lead term: 1.85 / 1.5 / 1.5 us per addition, every one of them
// bound by ~250 dependent instructions per warp and round, with 7 of 8 warps waiting) before settling on this:
//
//   h = sum over STREAMS i of  nc_i * m_i * tail(f_i)        (the S-polynomial's two halves, then one stream per addition)
//
// A stream is a cursor into the term arena plus its multiplier; its current term (scaled) is its HEAD.  The lead term of
// h is the smallest head key (keys ascend as monomials descend) with the SUM of the coefficients of every head that
// carries it; a zero sum is a cancelled monomial and is skipped.  This is the division algorithm of buchberger.cpp:24-49
// with Polynomial operator+ (polynomials.cpp:148-177) evaluated lazily: the sequence of lead terms, the reducer chosen
// for each (first in G_ whose lead monomial divides), the number of additions (the reward) and the remainder are those
// of the reference, term for term and bit for bit -- only the order in which coefficients are summed differs, and
// addition mod p is associative.  An addition costs O(1) and each term of a reducer is touched once.
//
// One ROUND per lead term, no barrier: the warp (a) tests the reducer lead monomials against the current lead monomial M,
// four chunks of 32 with their loads in flight together (divisor search, buchberger.cpp:27-32), (b) every lane advances
// those of its streams whose head is M -- the next raw term was prefetched into registers a round earlier -- and takes
// the minimum over its heads with the coefficient sum at it, (c) three warp reductions give the next lead term.  A
// divisor f opens a new stream whose head comes straight from f's head record (no dependent load).
//
// The first 32 streams are one per lane in REGISTERS (head, multiplier, cursor and the prefetched term behind the head),
// the rest live in the warp's slice of shared memory (BBS_KMAX in all).  Exhausted streams are squeezed out whenever a new
// row of 32 would be started (streams_gc), so the table stays as short as the number of LIVE streams allows; if it is
// full of live streams, h is consolidated into a scratch list behind the slot's term arena and goes on as a single
// stream (bb_set_wide(2) / (3) cap the streams at 6 / 48 so that the tests reach both).
//
// Episode records are bit-identical to the materialising runner's (bb_set_wide(0): warp_merge / warp_reduce, an
// independent second implementation); of the traffic counters, terms_read / terms_written (|h| per addition) are not
// observable without materialising h and count the reducers' lengths only.
#pragma once
#include "bb_device.cuh"

#ifndef BBS_KMAX
#define BBS_KMAX 512            // streams of one step (44 bytes of shared memory each, per warp)
#endif
#define BBS_NONE 0xffffffffu

// The scratch list of a consolidation lies BEHIND the term arena of the slot (bbenv.cu lays hkey right after tkey and
// hcoef right after tcoef, max_terms a multiple of 8), so term index max_terms + i addresses scratch term i through the
// same two base pointers and a stream's cursor needs no flag.

struct WarpStreams {             // one per warp in dynamic shared memory; entry i belongs to lane i % 32; row 0 (i < 32)
  uint64_t key[BBS_KMAX];        // is only a spill area for the lanes' registers (streams_gc)
  uint64_t adj[BBS_KMAX];        // key(multiplier monomial) - bias
  uint64_t pkey[BBS_KMAX];       // raw key of the term behind the head (valid when ptr < end), unless still in the owner's registers
  uint32_t coef[BBS_KMAX];       // head: scaled coefficient
  uint32_t nc[BBS_KMAX];         // multiplier coefficient
  uint32_t pcoef[BBS_KMAX];
  uint32_t ptr[BBS_KMAX];        // index of the term behind the head
  uint32_t end[BBS_KMAX];        // one past the stream's last term
};

// Warp-uniform state of a step's reduction plus per-lane registers; every member is a scalar so that the whole record
// lives in registers.
struct StreamState {
  int K;                    // streams in use
  int kmax;                 // BBS_KMAX, or less under bb_set_wide(2 / 3)
  int cz;                   // scratch half the next consolidation writes
  // per lane: stream `lane` (row 0).  k0 is all ones when the lane has no stream or it is exhausted.
  uint64_t k0, adj0, pk0; uint32_t c0, nc0, pc0, p0, e0;
  // per lane: the raw term behind the head of stream pend_i >= 32, loaded but not yet stored to st.pkey / st.pcoef
  int pend_i; uint64_t pend_k; uint32_t pend_c;
  uint32_t bad;             // per lane: a produced key overflowed its exponent fields
  __device__ __forceinline__ void clear() {
    K = 0; k0 = ~0ull; adj0 = pk0 = 0ull; c0 = nc0 = pc0 = p0 = e0 = 0u; pend_i = -1; pend_k = 0ull; pend_c = 0u; bad = 0u;
  }
};

__device__ __forceinline__ uint32_t bbf_reduce(const BBField& F, uint32_t x) {   // x mod p for any x < 2^32
  const uint32_t q = __umulhi(x, F.mu);
  const uint32_t r = x - q * F.p;
  return r >= F.p ? r - F.p : r;
}

// One round (see the header): consumes the lead monomial M (streams whose head is M advance) when `consume`, searches
// M's first divisor in G_ when `search`, and returns the next lead term of h: (M2, S2), M2 all ones when h is exhausted,
// S2 in [0, p) (0: the monomial cancelled).  found / fidx: position in G_ and basis index of the divisor, found = -1 if
// none (or not searched).
template <int NV>
__device__ __forceinline__ void streams_round(StreamState& ws, WarpStreams& st, const BBField F, const uint64_t M,
                                              const bool consume, const bool search, const uint64_t* rlm, const uint32_t* ridx,
                                              const int nR, const bool sorted, const uint64_t* tk, const uint32_t* tc,
                                              uint64_t& M2, uint32_t& S2, int& found, uint32_t& fidx) {
  typedef KL<NV> K;
  const int lane = bb_lane();
  // (a) divisor search, 128 reducers per pass.  sorted: G_ ascends in lead monomial (keys descend), so the search may
  // stop at the first pass that holds a reducer whose lead monomial exceeds M (key below M's): nothing after it can divide.
  uint32_t best = BBS_NONE;
  if (search) {
    const uint64_t stop = sorted ? M : 0ull;
    const uint64_t mg = (M & K::ex_mask) | K::ge_mask;
#pragma unroll 1
    for (int base = 0; base < nR; base += 128) {
      const int r0 = base + lane, r1 = r0 + 32, r2 = r0 + 64, r3 = r0 + 96;
      const bool v0 = r0 < nR, v1 = r1 < nR, v2 = r2 < nR, v3 = r3 < nR;
      const uint64_t l0 = v0 ? rlm[r0] : ~0ull, l1 = v1 ? rlm[r1] : ~0ull, l2 = v2 ? rlm[r2] : ~0ull, l3 = v3 ? rlm[r3] : ~0ull;
      const bool h0 = v0 && ((mg - (l0 & K::ex_mask)) & K::ge_mask) == K::ge_mask, h1 = v1 && ((mg - (l1 & K::ex_mask)) & K::ge_mask) == K::ge_mask;
      const bool h2 = v2 && ((mg - (l2 & K::ex_mask)) & K::ge_mask) == K::ge_mask, h3 = v3 && ((mg - (l3 & K::ex_mask)) & K::ge_mask) == K::ge_mask;
      const uint32_t cand = h0 ? (uint32_t)r0 : (h1 ? (uint32_t)r1 : (h2 ? (uint32_t)r2 : (h3 ? (uint32_t)r3 : BBS_NONE)));
      best = __reduce_min_sync(BB_FULL, cand);
      if (best != BBS_NONE) break;
      if (__any_sync(BB_FULL, l0 < stop || l1 < stop || l2 < stop || l3 < stop)) break;   // absent entries are all ones
    }
  }
  // (b) this lane's streams: advance the ones at M, minimum head and the coefficient sum at it.  Row 0 in registers:
  uint64_t mk = ws.k0;
  if (consume && mk == M) {
    if (ws.p0 < ws.e0) {
      mk = ws.pk0 + ws.adj0;
      ws.c0 = bbf_mulmod(F, ws.pc0, ws.nc0);
      if (mk & K::g_all) ws.bad = 1u;
      ws.p0++;
      if (ws.p0 < ws.e0) { ws.pk0 = tk[ws.p0]; ws.pc0 = tc[ws.p0]; }   // the term behind the new head, needed a round later at the earliest
    } else {
      mk = ~0ull;
    }
    ws.k0 = mk;
  }
  uint32_t ms = ws.c0;
#pragma unroll 1
  for (int i = lane + 32; i < ws.K; i += 32) {
    uint64_t k = st.key[i];
    if (consume && k == M) {
      const uint32_t p = st.ptr[i];
      if (p < st.end[i]) {
        uint64_t kr; uint32_t cr;
        if (ws.pend_i == i) { kr = ws.pend_k; cr = ws.pend_c; ws.pend_i = -1; }
        else { kr = st.pkey[i]; cr = st.pcoef[i]; }
        k = kr + st.adj[i];
        const uint32_t c = bbf_mulmod(F, cr, st.nc[i]);
        if (k & K::g_all) ws.bad = 1u;
        st.key[i] = k; st.coef[i] = c; st.ptr[i] = p + 1u;
        if (p + 1u < st.end[i]) {   // fetch the term behind the new head; it stays in registers until it is needed
          if (ws.pend_i >= 0) { st.pkey[ws.pend_i] = ws.pend_k; st.pcoef[ws.pend_i] = ws.pend_c; }
          ws.pend_i = i; ws.pend_k = tk[p + 1u]; ws.pend_c = tc[p + 1u];
        }
      } else {
        k = ~0ull; st.key[i] = k;
      }
    }
    const uint32_t c = st.coef[i];
    if (k < mk) { mk = k; ms = c; } else if (k == mk) ms += c;   // at most BBS_KMAX / 32 values below 2^16: no overflow
  }
  // (c) 64-bit minimum through two 32-bit reductions, the coefficient sum at it
  const uint32_t hi = __reduce_min_sync(BB_FULL, (uint32_t)(mk >> 32));
  const uint32_t lo = __reduce_min_sync(BB_FULL, (uint32_t)(mk >> 32) == hi ? (uint32_t)mk : 0xffffffffu);
  const uint64_t wk = ((uint64_t)hi << 32) | lo;
  const uint32_t wsum = __reduce_add_sync(BB_FULL, mk == wk ? ms : 0u);   // <= BBS_KMAX values below 2^16
  M2 = wk; S2 = bbf_reduce(F, wsum);
  found = best == BBS_NONE ? -1 : (int)best;
  fidx = best == BBS_NONE ? 0u : ridx[best];
}

// Opens stream ws.K: head (hk, hc) already scaled, multiplier (adj, nc), the terms behind the head at [next, end).
// Only the owner lane touches the entry.
__device__ __forceinline__ void stream_open(StreamState& ws, WarpStreams& st, uint64_t hk, uint32_t hc, uint64_t adj, uint32_t nc,
                                            uint32_t next, uint32_t end, const uint64_t* tk, const uint32_t* tc) {
  const int i = ws.K;
  if (bb_lane() == (i & 31)) {
    if (i < 32) {
      ws.k0 = hk; ws.c0 = hc; ws.adj0 = adj; ws.nc0 = nc; ws.p0 = next; ws.e0 = end;
      if (next < end) { ws.pk0 = tk[next]; ws.pc0 = tc[next]; }
    } else {
      st.key[i] = hk; st.coef[i] = hc; st.adj[i] = adj; st.nc[i] = nc; st.ptr[i] = next; st.end[i] = end;
      if (next < end) {
        if (ws.pend_i >= 0) { st.pkey[ws.pend_i] = ws.pend_k; st.pcoef[ws.pend_i] = ws.pend_c; }
        ws.pend_i = i; ws.pend_k = tk[next]; ws.pend_c = tc[next];
      }
    }
  }
  ws.K = i + 1;
}

// Squeezes the exhausted streams out of the table (order kept).  Entries change owner lanes, so every register-held part
// goes to shared memory first and row 0 is read back afterwards.  Chunks of 32: a chunk's survivors land at or below their
// own positions, inside what has been read.
static __device__ __noinline__ void streams_gc(WarpStreams& st, StreamState& ws) {
  const int lane = bb_lane();
  const uint32_t ltm = bb_lt_mask();
  if (ws.pend_i >= 0) { st.pkey[ws.pend_i] = ws.pend_k; st.pcoef[ws.pend_i] = ws.pend_c; ws.pend_i = -1; }
  st.key[lane] = ws.k0; st.adj[lane] = ws.adj0; st.pkey[lane] = ws.pk0; st.coef[lane] = ws.c0; st.nc[lane] = ws.nc0;
  st.pcoef[lane] = ws.pc0; st.ptr[lane] = ws.p0; st.end[lane] = ws.e0;
  __syncwarp();
  const int K = ws.K;
  int w = 0;
#pragma unroll 1
  for (int b0 = 0; b0 < K; b0 += 32) {
    const int i = b0 + lane;
    const bool live = i < K && st.key[i] != ~0ull;
    uint64_t a = 0, b = 0, c = 0; uint32_t d = 0, e = 0, f = 0, g = 0, h = 0;
    if (live) { a = st.key[i]; b = st.adj[i]; c = st.pkey[i]; d = st.coef[i]; e = st.nc[i]; f = st.pcoef[i]; g = st.ptr[i]; h = st.end[i]; }
    const uint32_t m = __ballot_sync(BB_FULL, live);
    __syncwarp();   // every lane has read its entry before any lane writes
    if (live) {
      const int pos = w + __popc(m & ltm);
      st.key[pos] = a; st.adj[pos] = b; st.pkey[pos] = c; st.coef[pos] = d; st.nc[pos] = e; st.pcoef[pos] = f; st.ptr[pos] = g; st.end[pos] = h;
    }
    w += __popc(m);
    __syncwarp();
  }
  ws.K = w;
  ws.k0 = ~0ull;
  if (lane < w) {
    ws.k0 = st.key[lane]; ws.adj0 = st.adj[lane]; ws.pk0 = st.pkey[lane]; ws.c0 = st.coef[lane]; ws.nc0 = st.nc[lane];
    ws.pc0 = st.pcoef[lane]; ws.p0 = st.ptr[lane]; ws.e0 = st.end[lane];
  }
  __syncwarp();
}

// Consolidation: every pending term of h, from the lead term (M, S) on, is written in order to scratch half ws.cz and the
// streams are replaced by ONE stream over that list.  (M, S) becomes its head (the first term with a nonzero sum), or
// M = all ones if nothing is left.  sbase: term index of the scratch (= max_terms).  Returns the number of terms, or -1 if
// they do not fit `cap`.
template <int NV>
__device__ __noinline__ int streams_consolidate(StreamState& ws, WarpStreams& st, const BBField F, uint64_t& M, uint32_t& S,
                                                uint64_t* tk, uint32_t* tc, uint32_t sbase, int cap) {
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
    int found; uint32_t fidx;
    streams_round<NV>(ws, st, F, m, true, false, nullptr, nullptr, 0, false, tk, tc, m, s, found, fidx);
  }
  ws.pend_i = -1;   // every stream is exhausted
  ws.K = 0;
  ws.k0 = ~0ull;
  ws.cz ^= 1;
  M = fm; S = fs;
  __syncwarp();     // lane 0's list before lane 0 (the owner of stream 0) reads it back
  if (t > 0) stream_open(ws, st, fm, fs, 0ull, 1u, base + 1u, base + (uint32_t)t, tk, tc);
  return t;
}

// reduce(spoly(G[i], G[j]), G_) (buchberger.cpp:18-49) for the pair heads (hf, hg) and gamma = the key of the pair's lcm,
// with the dividend as a set of streams.  The remainder goes to (rk, rc) [cap rcap]; returns its length or a negative
// BB_STATUS_* on a fault; `steps` = reductions, `sug` = the sugar of the result (polynomials.cpp:150, 198).
template <int NV>
__device__ __forceinline__ int warp_reduce_streams(const BBParams& P, const Env& e, StreamState& ws, WarpStreams& st,
                                                   const GHead hf, const GHead hg, const uint64_t gam, int& sug, int& steps,
                                                   uint64_t* rk, uint32_t* rc, int rcap, Ctr& ct) {
  typedef KL<NV> K;
  const BBField F = P.F;
  const GHeadMem* gh = ENV_PTR(GHeadMem, e, P, o_ghead);
  uint64_t* tk = ENV_PTR(uint64_t, e, P, o_tkey);      // the term arena and, from index max_terms on, the consolidation
  uint32_t* tc = ENV_PTR(uint32_t, e, P, o_tcoef);     // scratch (two halves of max_poly_terms terms)
  const uint64_t* rlm = ENV_PTR(uint64_t, e, P, o_rlm);
  const uint32_t* ridx = ENV_PTR(uint32_t, e, P, o_ridx);
  const int nR = e.nG;
  const bool sorted = P.sort_reducers != 0;
  int rlen = 0;
  steps = 0;
  ws.clear();
  // s = (gamma / LT f) tail(f) - (gamma / LT g) tail(g): the lead terms cancel exactly (buchberger.cpp:18-21); two streams
  // whose heads come from the head records
  if (hf.len > 1u) {
    const uint64_t adj = gam - hf.lm, k = hf.k1 + adj;
    if (k & K::g_all) return -BB_STATUS_OVERFLOW_EXPONENT;
    stream_open(ws, st, k, bbf_mulmod(F, hf.c1, hf.invlc), adj, hf.invlc, hf.off + 2u, hf.off + hf.len, tk, tc);
  }
  if (hg.len > 1u) {
    const uint64_t adj = gam - hg.lm, k = hg.k1 + adj;
    const uint32_t nc = F.p - hg.invlc;
    if (k & K::g_all) return -BB_STATUS_OVERFLOW_EXPONENT;
    stream_open(ws, st, k, bbf_mulmod(F, hg.c1, nc), adj, nc, hg.off + 2u, hg.off + hg.len, tk, tc);
  }
  uint64_t M; uint32_t S, fidx; int found;
  streams_round<NV>(ws, st, F, ~0ull, false, false, rlm, ridx, nR, sorted, tk, tc, M, S, found, fidx);
#pragma unroll 1
  while (M != ~0ull) {
    uint64_t M2; uint32_t S2;
    if (S == 0u) {   // the monomial cancelled: it is not a term of h
      streams_round<NV>(ws, st, F, M, true, false, rlm, ridx, nR, sorted, tk, tc, M2, S2, found, fidx);
      M = M2; S = S2;
      continue;
    }
    streams_round<NV>(ws, st, F, M, true, true, rlm, ridx, nR, sorted, tk, tc, M2, S2, found, fidx);
    ct.lms += (found >= 0) ? (unsigned)(found + 1) : (unsigned)nR;
    if (found >= 0) {   // h <- h - (LT h / LT f) f: the lead terms cancel, f's tail becomes a stream
      const GHead f = load_head(gh + fidx);
      const uint32_t c = bbf_mulmod(F, S, f.invlc);
      const uint32_t nc = F.p - c;              // c != 0
      const uint64_t adj = M - f.lm;            // key(LM h / LM f) - bias
      const int sf = (int)f.sug + (int)(uint32_t)(f.lm >> K::dshift) - (int)(uint32_t)(M >> K::dshift);
      sug = sf > sug ? sf : sug;
      ct.tread += f.len;
      steps++;
      if (f.len > 1u) {
        if (ws.K >= ws.kmax || (ws.K >= 64 && (ws.K & 31) == 0)) {   // a new row, or no room: drop the exhausted streams first
          streams_gc(st, ws);
          if (ws.K >= ws.kmax && streams_consolidate<NV>(ws, st, F, M2, S2, tk, tc, (uint32_t)P.max_terms, P.max_poly_terms) < 0)
            return -BB_STATUS_OVERFLOW_SCRATCH;
        }
        const uint64_t k = f.k1 + adj;
        const uint32_t ck = bbf_mulmod(F, f.c1, nc);
        if (k & K::g_all) return -BB_STATUS_OVERFLOW_EXPONENT;
        stream_open(ws, st, k, ck, adj, nc, f.off + 2u, f.off + f.len, tk, tc);
        if (k < M2) { M2 = k; S2 = ck; } else if (k == M2) S2 = bbf_addmod(F, S2, ck);
      }
    } else {            // no divisor: the lead term moves to the remainder
      if (rlen >= rcap) return -BB_STATUS_OVERFLOW_TERMS;
      if (bb_lane() == 0) { rk[rlen] = M; rc[rlen] = S; }
      rlen++; ct.moves++;
    }
    M = M2; S = S2;
  }
  if (__any_sync(BB_FULL, ws.bad != 0u)) return -BB_STATUS_OVERFLOW_EXPONENT;
  __syncwarp();   // lane 0's remainder before every lane reads it (update(), hashes)
  return rlen;
}

// BuchbergerEnv::step for the pair in row `row` of P, as warp_step (bb_device.cuh) with reduce() by streams.
template <int NV>
__device__ __forceinline__ int warp_step_streams(const BBParams& P, Env& e, StreamState& ws, WarpStreams& st, int row,
                                                 uint32_t& pair, Ctr& ct) {
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
  const int rlen = warp_reduce_streams<NV>(P, e, ws, st, hf, hg, gam, sug, steps, ENV_PTR(uint64_t, e, P, o_tkey) + e.nT,
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
