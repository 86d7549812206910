// bb_wide.cuh -- the stream reducer of bb_streams.cuh run by a whole CTA (BBW_WARPS warps) on ONE environment, sm_100a.
//
// Why both: a cyclic-6 launch is as long as its longest episode (697 000 dependent additions in the longest of 1024
// seeded-Random episodes), so what decides it is the latency of one ROUND (one lead term of the dividend: divisor search
// + advance of the streams at it + next lead term, see bb_streams.cuh).  A single warp runs a round as ~380 dependent
// instructions; a CTA splits the divisor search and the streams over its warps -- thread t owns stream t in REGISTERS
// (streams beyond BBW_THREADS live in shared memory), one slice of the reducer lead monomials per thread -- and pays one
// block barrier per round: every warp reduces its part (three REDUX for the 64-bit minimum head key and the coefficient
// sum at it, one for the first divisor), writes a 16-byte record, and after the barrier every warp folds the records
// with the same four reductions (one record per lane).  Same algorithm, same results (the tests compare both with the
// materialising runner), shorter chain; the warp version is the one to use when there are many more episodes than SMs.
#pragma once
#include "bb_streams.cuh"

#ifndef BBW_WARPS
#define BBW_WARPS 8
#endif
#define BBW_THREADS (BBW_WARPS * 32)
#ifndef BBW_MIN_CTAS
#define BBW_MIN_CTAS 2
#endif
#ifndef BBW_KMAX
#define BBW_KMAX 1024           // streams of one step: BBW_THREADS in registers, the rest in shared memory
#endif
static_assert(BBW_KMAX % BBW_THREADS == 0 && BBW_KMAX > BBW_THREADS, "whole rows of streams");
static_assert(BBW_WARPS <= 32, "one record per lane in the fold");

struct WideStreams {             // dynamic shared memory: streams BBW_THREADS .. BBW_KMAX - 1 (entry i at index i - BBW_THREADS)
  uint64_t key[BBW_KMAX - BBW_THREADS];
  uint64_t adj[BBW_KMAX - BBW_THREADS];
  uint64_t pkey[BBW_KMAX - BBW_THREADS];
  uint32_t coef[BBW_KMAX - BBW_THREADS];
  uint32_t nc[BBW_KMAX - BBW_THREADS];
  uint32_t pcoef[BBW_KMAX - BBW_THREADS];
  uint32_t ptr[BBW_KMAX - BBW_THREADS];
  uint32_t end[BBW_KMAX - BBW_THREADS];
};

struct WideShared {
  __align__(16) uint4 wrec[2][BBW_WARPS];   // per-warp round results, double-buffered: (min head key lo, hi, coefficient sum, first divisor position)
  int row;                  // the pair row warp 0 selected
  long long upd;            // result of warp_add_basis
};

// One round by the block; contract as streams_round.  `half` = which half of sh.wrec this round writes.
template <int NV>
__device__ __forceinline__ void wide_round(WideShared& sh, int& half, StreamState& ws, WideStreams& st, const BBField F,
                                           const uint64_t M, const bool consume, const bool search, const uint64_t* rlm,
                                           const uint32_t* ridx, const int nR, const bool sorted, const uint64_t* tk,
                                           const uint32_t* tc, uint64_t& M2, uint32_t& S2, int& found, uint32_t& fidx) {
  typedef KL<NV> K;
  const int tid = threadIdx.x, lane = bb_lane();
  // (a) this thread's slice of the reducer lead monomials; sorted: stop at the first whose lead monomial exceeds M
  uint32_t best = BBS_NONE;
  if (search) {
    const uint64_t stop = sorted ? M : 0ull;
    const uint64_t mg = (M & K::ex_mask) | K::ge_mask;
#pragma unroll 1
    for (int r = tid; r < nR; r += BBW_THREADS) {
      const uint64_t l = rlm[r];
      if (l < stop) break;
      if (((mg - (l & K::ex_mask)) & K::ge_mask) == K::ge_mask) { best = (uint32_t)r; break; }
    }
  }
  // (b) this thread's streams: stream tid in registers, streams tid + BBW_THREADS, ... in shared memory
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
  for (int i = tid; i < ws.K - BBW_THREADS; i += BBW_THREADS) {
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
        if (p + 1u < st.end[i]) {
          if (ws.pend_i >= 0) { st.pkey[ws.pend_i] = ws.pend_k; st.pcoef[ws.pend_i] = ws.pend_c; }
          ws.pend_i = i; ws.pend_k = tk[p + 1u]; ws.pend_c = tc[p + 1u];
        }
      } else {
        k = ~0ull; st.key[i] = k;
      }
    }
    const uint32_t c = st.coef[i];
    if (k < mk) { mk = k; ms = c; } else if (k == mk) ms += c;
  }
  // (c) warp: the first divisor of the warp's slices, then the 64-bit minimum through two 32-bit reductions and the
  // coefficient sum at it.  The divisor's basis index is fetched here (one uniform load per warp, in flight during the
  // three reductions) and travels with its position in one word, position on top: the fold's minimum still picks the first
  // divisor and the block knows its head record's address one dependent load earlier.
  {
    best = __reduce_min_sync(BB_FULL, best);
    if (best != BBS_NONE) best = (best << 16) | ridx[best];   // positions and basis indices are below 2^16 (max_basis <= 65535)
    const uint32_t hi = __reduce_min_sync(BB_FULL, (uint32_t)(mk >> 32));
    const uint32_t lo = __reduce_min_sync(BB_FULL, (uint32_t)(mk >> 32) == hi ? (uint32_t)mk : 0xffffffffu);
    const uint32_t wsum = __reduce_add_sync(BB_FULL, ((uint32_t)(mk >> 32) == hi && (uint32_t)mk == lo) ? ms : 0u);
    if (lane == 0) sh.wrec[half][tid >> 5] = make_uint4(lo, hi, wsum, best);
  }
  __syncthreads();
  // (d) every warp folds the warps' records, one per lane, with the same reductions
  {
    uint4 v = make_uint4(0xffffffffu, 0xffffffffu, 0u, BBS_NONE);
    if (lane < BBW_WARPS) v = sh.wrec[half][lane];
    const uint32_t hi = __reduce_min_sync(BB_FULL, v.y);
    const uint32_t lo = __reduce_min_sync(BB_FULL, v.y == hi ? v.x : 0xffffffffu);
    const uint32_t gs = __reduce_add_sync(BB_FULL, (v.y == hi && v.x == lo) ? v.z : 0u);   // < BBW_KMAX values below 2^16
    best = __reduce_min_sync(BB_FULL, v.w);
    M2 = ((uint64_t)hi << 32) | lo; S2 = bbf_reduce(F, gs);
  }
  half ^= 1;
  found = best == BBS_NONE ? -1 : (int)(best >> 16);
  fidx = best == BBS_NONE ? 0u : (best & 0xffffu);
}

// Opens stream ws.K (owner: thread K % BBW_THREADS); as stream_open.
__device__ __forceinline__ void wide_open(StreamState& ws, WideStreams& st, uint64_t hk, uint32_t hc, uint64_t adj, uint32_t nc,
                                          uint32_t next, uint32_t end, const uint64_t* tk, const uint32_t* tc) {
  const int i = ws.K;
  if ((int)threadIdx.x == (i & (BBW_THREADS - 1))) {
    if (i < BBW_THREADS) {
      ws.k0 = hk; ws.c0 = hc; ws.adj0 = adj; ws.nc0 = nc; ws.p0 = next; ws.e0 = end;
      if (next < end) { ws.pk0 = tk[next]; ws.pc0 = tc[next]; }
    } else {
      const int j = i - BBW_THREADS;
      st.key[j] = hk; st.coef[j] = hc; st.adj[j] = adj; st.nc[j] = nc; st.ptr[j] = next; st.end[j] = end;
      if (next < end) {
        if (ws.pend_i >= 0) { st.pkey[ws.pend_i] = ws.pend_k; st.pcoef[ws.pend_i] = ws.pend_c; }
        ws.pend_i = j; ws.pend_k = tk[next]; ws.pend_c = tc[next];
      }
    }
  }
  ws.K = i + 1;
}

// Consolidation, as streams_consolidate: h from (M, S) on goes to scratch half ws.cz in order, one stream over it remains.
template <int NV>
__device__ __noinline__ int wide_consolidate(WideShared& sh, int& half, StreamState& ws, WideStreams& st, const BBField F,
                                             uint64_t& M, uint32_t& S, uint64_t* tk, uint32_t* tc, uint32_t sbase, int cap) {
  const uint32_t base = sbase + (uint32_t)(ws.cz * cap);
  int t = 0;
  uint64_t m = M, fm = ~0ull; uint32_t s = S, fs = 0u;
  while (m != ~0ull) {
    if (s != 0u) {
      if (t >= cap) return -1;
      if (t == 0) { fm = m; fs = s; }
      if (threadIdx.x == 0) { tk[base + t] = m; tc[base + t] = s; }
      t++;
    }
    int found; uint32_t fidx;
    wide_round<NV>(sh, half, ws, st, F, m, true, false, nullptr, nullptr, 0, false, tk, tc, m, s, found, fidx);
  }
  ws.pend_i = -1;
  ws.K = 0;
  ws.k0 = ~0ull;
  ws.cz ^= 1;
  M = fm; S = fs;
  __syncthreads();   // thread 0's list before thread 0 (the owner of stream 0) reads it back
  if (t > 0) wide_open(ws, st, fm, fs, 0ull, 1u, base + 1u, base + (uint32_t)t, tk, tc);
  return t;
}

// reduce(spoly(G[i], G[j]), G_) by the block; contract as warp_reduce_streams.
template <int NV>
__device__ __forceinline__ int block_reduce_streams(const BBParams& P, const Env& e, WideShared& sh, int& half, StreamState& ws,
                                                    WideStreams& st, const GHead hf, const GHead hg, const uint64_t gam, int& sug,
                                                    int& steps, uint64_t* rk, uint32_t* rc, int rcap, Ctr& ct) {
  typedef KL<NV> K;
  const BBField F = P.F;
  const GHeadMem* gh = ENV_PTR(GHeadMem, e, P, o_ghead);
  uint64_t* tk = ENV_PTR(uint64_t, e, P, o_tkey);
  uint32_t* tc = ENV_PTR(uint32_t, e, P, o_tcoef);
  const uint64_t* rlm = ENV_PTR(uint64_t, e, P, o_rlm);
  const uint32_t* ridx = ENV_PTR(uint32_t, e, P, o_ridx);
  const int nR = e.nG;
  const bool sorted = P.sort_reducers != 0;
  int rlen = 0;
  steps = 0;
  ws.clear();
  if (hf.len > 1u) {
    const uint64_t adj = gam - hf.lm, k = hf.k1 + adj;
    if (k & K::g_all) return -BB_STATUS_OVERFLOW_EXPONENT;
    wide_open(ws, st, k, bbf_mulmod(F, hf.c1, hf.invlc), adj, hf.invlc, hf.off + 2u, hf.off + hf.len, tk, tc);
  }
  if (hg.len > 1u) {
    const uint64_t adj = gam - hg.lm, k = hg.k1 + adj;
    const uint32_t nc = F.p - hg.invlc;
    if (k & K::g_all) return -BB_STATUS_OVERFLOW_EXPONENT;
    wide_open(ws, st, k, bbf_mulmod(F, hg.c1, nc), adj, nc, hg.off + 2u, hg.off + hg.len, tk, tc);
  }
  uint64_t M; uint32_t S, fidx; int found;
  wide_round<NV>(sh, half, ws, st, F, ~0ull, false, false, rlm, ridx, nR, sorted, tk, tc, M, S, found, fidx);
#pragma unroll 1
  while (M != ~0ull) {
    uint64_t M2; uint32_t S2;
    if (S == 0u) {   // the monomial cancelled: it is not a term of h
      wide_round<NV>(sh, half, ws, st, F, M, true, false, rlm, ridx, nR, sorted, tk, tc, M2, S2, found, fidx);
      M = M2; S = S2;
      continue;
    }
    wide_round<NV>(sh, half, ws, st, F, M, true, true, rlm, ridx, nR, sorted, tk, tc, M2, S2, found, fidx);
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
        if (ws.K >= ws.kmax && wide_consolidate<NV>(sh, half, ws, st, F, M2, S2, tk, tc, (uint32_t)P.max_terms, P.max_poly_terms) < 0)
          return -BB_STATUS_OVERFLOW_SCRATCH;
        const uint64_t k = f.k1 + adj;
        const uint32_t ck = bbf_mulmod(F, f.c1, nc);
        if (k & K::g_all) return -BB_STATUS_OVERFLOW_EXPONENT;
        wide_open(ws, st, k, ck, adj, nc, f.off + 2u, f.off + f.len, tk, tc);
        if (k < M2) { M2 = k; S2 = ck; } else if (k == M2) S2 = bbf_addmod(F, S2, ck);
      }
    } else {            // no divisor: the lead term moves to the remainder
      if (rlen >= rcap) return -BB_STATUS_OVERFLOW_TERMS;
      if (threadIdx.x == 0) { rk[rlen] = M; rc[rlen] = S; }
      rlen++; ct.moves++;
    }
    M = M2; S = S2;
  }
  if (__syncthreads_or(ws.bad != 0u)) return -BB_STATUS_OVERFLOW_EXPONENT;   // also: thread 0's remainder before warp 0 reads it
  return rlen;
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
// `strategy`).  e and every scalar below are block-uniform.  Pair selection and update() (Gebauer-Moeller) are the warp
// routines of bb_device.cuh run by warp 0.  Returns the number of polynomial additions; `pair` receives (j << 16) | i.
template <int NV>
__device__ __forceinline__ int block_step(const BBParams& P, Env& e, WideShared& sh, int& half, StreamState& ws, WideStreams& st,
                                          int strategy, uint32_t* sel_rng, uint32_t& pair, Ctr& ct) {
  typedef KL<NV> K;
  const int tid = threadIdx.x;
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
  const int rlen = block_reduce_streams<NV>(P, e, sh, half, ws, st, hf, hg, gam, sug, steps, ENV_PTR(uint64_t, e, P, o_tkey) + e.nT,
                                            ENV_PTR(uint32_t, e, P, o_tcoef) + e.nT, P.max_terms - e.nT, ct);
  if (rlen < 0) { e.status = -rlen; return 1 + steps; }
  if (rlen > 0) {
    ct.upb += (unsigned)e.nG; ct.upp += (unsigned)e.nP;
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
