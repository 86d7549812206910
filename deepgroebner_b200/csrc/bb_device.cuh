// bb_device.cuh -- warp-per-environment Buchberger step for sm_100a.
//
// One warp owns one environment slot.  All scalars of the slot (|G|, |P|, arena cursor, ...) are warp-uniform
// register values; all per-slot arrays live in the slot's HBM arena (L1/L2 resident in practice) and are
// accessed lane-strided, i.e. coalesced.  There is no block-level synchronisation on the step path.
//
// Reference semantics followed (deepgroebner/buchberger.cpp, polynomials.cpp) are cited at each function.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bbenv.h"
#include "bb_layout.cuh"

#define BB_FULL 0xffffffffu
#define BB_GOLD 0x9E3779B97F4A7C15ULL
#define BB_GOLD2 0xD1B54A32D192ED03ULL

// ---------------------------------------------------------------------------------------------------- parameters
struct __align__(16) BBEnvState {  // one per slot, 96 bytes
  int nG, nP, nT, status;
  int steps, adds, zero, nonzero;
  int rerolls, episode, truncated, pad1;
  unsigned long long trace_hash;
  unsigned long long rng;        // minstd_rand0 state of this environment's ideal stream
  double disc_return, discount;
  unsigned long long pad2[2];
};

struct BBDist {           // RandomBinomialIdealGenerator parameters (ideals.cpp:157-201)
  int enabled, d, s, homogeneous, pure, ncp;
  const double* cp;       // [d+1] cumulative degree probabilities (libstdc++ discrete_distribution::_M_cp)
  const uint64_t* basis;  // packed monomials of degree 0..d, lex-descending within a degree (ideals.cpp:39-64)
  const int* basis_off;   // [d+2]
};

struct BBParams {
  BBLayout L;
  int num_envs, k, cols, elimination, rewards, sort_input, sort_reducers;
  int max_basis, max_pairs, max_terms, max_poly_terms, max_gens, max_gen_terms;
  // per-slot arenas, each [num_envs][cap]
  uint64_t* tkey; uint32_t* tcoef;      // term arena: packed monomial, coefficient
  uint2* pmeta;                         // per basis polynomial: (offset, length) into the term arena
  uint64_t* lm; uint32_t* invlc;        // lead monomial key and 1/LC by basis index
  uint64_t* rlm; uint32_t* ridx;        // reducer list G_ in scan order: lead monomial key, basis index
  uint32_t* pairs;                      // pair list P in order: (j << 16) | i
  uint64_t* hkey; uint32_t* hcoef;      // dividend ping-pong scratch [num_envs][2][max_poly_terms]
  uint64_t* lscr;                       // [num_envs][max_basis] lcm scratch for update / order scratch
  BBEnvState* st;
  // staged input ideals, one per slot
  uint64_t* in_key; uint32_t* in_coef;  // [num_envs][max_gen_terms], each polynomial sorted descending
  int* in_off;                          // [num_envs][max_gens+1]; in_off[0] = 0, in_off[npoly] = nterms
  int* in_np;                           // [num_envs]
  // reduced-GB output arena, one per slot
  uint64_t* gkey; uint32_t* gcoef;      // [num_envs][max_terms]
  int* glen;                            // [num_envs][max_basis]
  int* gcount;                          // [num_envs][2] = (npolys, nterms)
  uint64_t* grlm; uint32_t* gridx; uint32_t* gflag;  // [num_envs][max_basis] scratch of warp_final_gb
  BBDist dist;
  unsigned long long* counters;         // bb_counters as 12 x u64
};

enum { CT_STEPS = 0, CT_ADDS, CT_TREAD, CT_TWRITE, CT_LMS, CT_MOVES, CT_UPB, CT_UPP, CT_OBS, CT_NONZERO, CT_ZERO,
       CT_EPISODES, CT_COUNT };

struct WarpCounters {
  unsigned long long v[CT_COUNT];
  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int i = 0; i < CT_COUNT; i++) v[i] = 0;
  }
};

// Warp-uniform view of one slot.
struct Env {
  uint64_t* tkey; uint32_t* tcoef; uint2* pmeta; uint64_t* lm; uint32_t* invlc; uint64_t* rlm; uint32_t* ridx;
  uint32_t* pairs; uint64_t* hkey; uint32_t* hcoef; uint64_t* lscr;
  int nG, nP, nT, status;
  uint64_t guard;  // OR of every produced monomial key: any guard bit set => exponent/degree overflow
};

__device__ __forceinline__ int bb_lane() { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t bb_lt_mask() { return (1u << bb_lane()) - 1u; }
__device__ __forceinline__ uint64_t bb_shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(BB_FULL, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(BB_FULL, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}
__device__ __forceinline__ uint64_t bb_mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
  z ^= z >> 27; z *= 0x94d049bb133111ebULL;
  z ^= z >> 31;
  return z;
}
__host__ __device__ __forceinline__ uint64_t bb_hash_item_impl(uint64_t x, uint64_t pos) {
  uint64_t z = x + BB_GOLD * (pos + 1);
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
  z ^= z >> 27; z *= 0x94d049bb133111ebULL;
  z ^= z >> 31;
  return z;
}

// e.guard is accumulated per lane; the overflow decision must be warp-uniform
__device__ __forceinline__ bool bb_guard_tripped(const BBLayout& L, uint64_t guard) {
  return __any_sync(BB_FULL, (guard & L.g_all) != 0ull);
}

__device__ __forceinline__ void env_bind(const BBParams& P, int slot, Env& e) {
  size_t s = (size_t)slot;
  e.tkey = P.tkey + s * P.max_terms;   e.tcoef = P.tcoef + s * P.max_terms;
  e.pmeta = P.pmeta + s * P.max_basis; e.lm = P.lm + s * P.max_basis; e.invlc = P.invlc + s * P.max_basis;
  e.rlm = P.rlm + s * P.max_basis;     e.ridx = P.ridx + s * P.max_basis;
  e.pairs = P.pairs + s * P.max_pairs;
  e.hkey = P.hkey + s * 2 * P.max_poly_terms; e.hcoef = P.hcoef + s * 2 * P.max_poly_terms;
  e.lscr = P.lscr + s * P.max_basis;
  e.guard = 0;
}
__device__ __forceinline__ void env_load(const BBParams& P, int slot, Env& e) {
  env_bind(P, slot, e);
  const BBEnvState& s = P.st[slot];
  e.nG = s.nG; e.nP = s.nP; e.nT = s.nT; e.status = s.status;
}
__device__ __forceinline__ void env_store(const BBParams& P, int slot, const Env& e) {
  if (bb_lane() == 0) {
    BBEnvState& s = P.st[slot];
    s.nG = e.nG; s.nP = e.nP; s.nT = e.nT; s.status = e.status;
  }
}

// ---------------------------------------------------------------------------------------------------- merge
// out = cA * mA * A  +  cB * mB * B     (Polynomial operator+ / Term*Polynomial, polynomials.cpp:148-202)
// A, B: term lists in ascending key order (descending grevlex), coefficients in [1,p).  adjX = key(mX) - bias.
// Cancelled terms are dropped.  Warp-cooperative merge path: every round each lane holds one element of each
// 32-wide window, finds its merged rank by a shuffle binary search in the other window, equal monomials are
// paired (A carries the sum, B retires), and the survivors are scattered in rank order.
// Returns the number of output terms, or -1 if `cap` would be exceeded.
__device__ __forceinline__ int warp_merge(const BBLayout& L, const uint64_t* __restrict__ Ak,
                                          const uint32_t* __restrict__ Ac, int nA, uint32_t cA, uint64_t adjA,
                                          const uint64_t* __restrict__ Bk, const uint32_t* __restrict__ Bc, int nB,
                                          uint32_t cB, uint64_t adjB, uint64_t* __restrict__ Ok,
                                          uint32_t* __restrict__ Oc, int cap, uint64_t& guard) {
  const int lane = bb_lane();
  int ia = 0, ib = 0, no = 0;
  if (nA + nB > 0 && nB == 0) {  // scaled copy
    if (nA > cap) return -1;
    for (int t = lane; t < nA; t += 32) {
      uint64_t k = Ak[t] + adjA; guard |= k;
      Ok[t] = k; Oc[t] = (cA == 1u) ? Ac[t] : bb_mulmod(L, Ac[t], cA);
    }
    return nA;
  }
  if (nA == 0) {
    if (nB > cap) return -1;
    for (int t = lane; t < nB; t += 32) {
      uint64_t k = Bk[t] + adjB; guard |= k;
      Ok[t] = k; Oc[t] = (cB == 1u) ? Bc[t] : bb_mulmod(L, Bc[t], cB);
    }
    return nB;
  }
  while (ia < nA || ib < nB) {
    uint64_t ak = ~0ull, bk = ~0ull;
    uint32_t ac = 0, bc = 0;
    const bool va = ia + lane < nA, vb = ib + lane < nB;
    if (va) { ak = Ak[ia + lane] + adjA; guard |= ak; ac = (cA == 1u) ? Ac[ia + lane] : bb_mulmod(L, Ac[ia + lane], cA); }
    if (vb) { bk = Bk[ib + lane] + adjB; guard |= bk; bc = (cB == 1u) ? Bc[ib + lane] : bb_mulmod(L, Bc[ib + lane], cB); }
    // rank of my A element: lane + #{b < a};  of my B element: lane + #{a <= b}  (A first on ties)
    int ca = 0, cb = 0;
#pragma unroll
    for (int step = 16; step; step >>= 1) {
      uint64_t vB = bb_shfl64(bk, ca + step - 1);
      uint64_t vA = bb_shfl64(ak, cb + step - 1);
      if (vB < ak) ca += step;
      if (vA <= bk) cb += step;
    }
    {
      uint64_t vB = bb_shfl64(bk, ca), vA = bb_shfl64(ak, cb);
      if (vB < ak) ca++;
      if (vA <= bk) cb++;
    }
    // partners: first b >= a sits at index ca; last a <= b sits at index cb-1
    uint64_t pB = bb_shfl64(bk, ca & 31);
    uint32_t pBc = __shfl_sync(BB_FULL, bc, ca & 31);
    uint64_t pA = bb_shfl64(ak, (cb - 1) & 31);
    const bool partA = va && ca < 32 && pB == ak;
    const bool partB = vb && cb > 0 && pA == bk;
    const int ra = lane + ca, rb = lane + cb;
    const bool emitA = va && ra < 32 && !(partA && ra == 31);
    const bool emitB = vb && rb < 32;
    uint32_t oc = ac;
    if (partA) oc = bb_addmod(L, ac, pBc);
    const bool liveA = emitA && oc != 0u;
    const bool liveB = emitB && !partB;
    uint32_t mine = (liveA ? (1u << ra) : 0u) | (liveB ? (1u << rb) : 0u);
    const uint32_t live = __reduce_or_sync(BB_FULL, mine);
    const int nlive = __popc(live);
    if (no + nlive > cap) return -1;
    if (liveA) { int pos = no + __popc(live & ((1u << ra) - 1u)); Ok[pos] = ak; Oc[pos] = oc; }
    if (liveB) { int pos = no + __popc(live & ((1u << rb) - 1u)); Ok[pos] = bk; Oc[pos] = bc; }
    no += nlive;
    const int da = __popc(__ballot_sync(BB_FULL, emitA)), db = __popc(__ballot_sync(BB_FULL, emitB));
    if (da + db == 0) return -1;  // only reachable with unsorted (overflowed) keys
    ia += da; ib += db;
  }
  return no;
}

// ---------------------------------------------------------------------------------------------------- reduce
// Division algorithm of buchberger.cpp:24-49 on the dividend h = (hk, hc, n) (ascending keys).
// Reducers are scanned IN ORDER (rlm[0..nR)), the first whose lead monomial divides LM(h) is used
// (h <- h - (LT h / LT f) f, steps++); otherwise LT(h) moves to the remainder.  The remainder is written
// to (rk, rc) [cap rcap]; returns its length or -1 (scratch overflow) / -2 (remainder overflow) /
// -3 (exponent overflow detected).
// hbuf_id: which half of the ping-pong scratch h currently lives in (0/1), or -1 if h lives elsewhere.
__device__ __forceinline__ int warp_reduce(const BBParams& P, Env& e, const uint64_t* hk, const uint32_t* hc, int n,
                                           int hbuf_id, const uint64_t* rlm, const uint32_t* ridx, int nR,
                                           uint64_t* rk, uint32_t* rc, int rcap, int& steps, WarpCounters& ct) {
  const BBLayout& L = P.L;
  const int lane = bb_lane();
  int rlen = 0;
  steps = 0;
  while (n > 0) {
    const uint64_t lead = hk[0];
    int found = -1;
    for (int base = 0; base < nR; base += 32) {
      int r = base + lane;
      bool ok = r < nR && bb_divides(L, rlm[r], lead);
      uint32_t b = __ballot_sync(BB_FULL, ok);
      if (b) { found = base + __ffs(b) - 1; break; }
    }
    ct.v[CT_LMS] += (found >= 0) ? (found + 1) : nR;
    if (found >= 0) {
      const int gi = ridx[found];
      const uint2 meta = e.pmeta[gi];
      const uint32_t c = bb_mulmod(L, hc[0], e.invlc[gi]);
      const uint64_t adj = lead - rlm[found];  // key(LM h / LM f) - bias
      const int ob = (hbuf_id == 0) ? 1 : 0;
      uint64_t* ok_ = e.hkey + (size_t)ob * P.max_poly_terms;
      uint32_t* oc_ = e.hcoef + (size_t)ob * P.max_poly_terms;
      int n2 = warp_merge(L, hk + 1, hc + 1, n - 1, 1u, 0ull, e.tkey + meta.x + 1, e.tcoef + meta.x + 1,
                          (int)meta.y - 1, bb_negmod(L, c), adj, ok_, oc_, P.max_poly_terms, e.guard);
      if (n2 < 0) return -1;
      if (bb_guard_tripped(L, e.guard)) return -3;  // garbage keys could otherwise keep the loop alive
      ct.v[CT_TREAD] += (unsigned)(n + (int)meta.y);
      ct.v[CT_TWRITE] += (unsigned)n2;
      __syncwarp();
      hk = ok_; hc = oc_; n = n2; hbuf_id = ob;
      steps++;
    } else {
      if (rlen >= rcap) return -2;
      if (lane == 0) { rk[rlen] = lead; rc[rlen] = hc[0]; }
      rlen++; hk++; hc++; n--;
      ct.v[CT_MOVES]++;
    }
  }
  __syncwarp();
  return rlen;
}

// ---------------------------------------------------------------------------------------------------- update
// update(G, P, f, elimination), buchberger.cpp:52-99, for the new basis element with index m = e.nG whose lead
// monomial key is fk (the element itself must already be in the arena; e.nG is NOT incremented here).
// GebauerMoeller: (1) old (i,j) dropped iff LM f | lcm_ij and lcm_ij != lcm_if and lcm_ij != lcm_jf  (:63-70);
// (2-4) with L_i = lcm(LM_i, LM f):  (i,m) is emitted iff no L_j strictly divides L_i, no j < i has L_j == L_i,
// and no j with L_j == L_i is coprime to f.  This is exactly what the reference's ascending std::map sweep with
// the "not divisible by a previously kept lcm" filter, v[0] representative and none_of(coprime) test produces
// (:72-85): a kept lcm is a divisibility-minimal distinct lcm, and divisors always precede in grevlex order.
// (5) new pairs in ascending i (:86), appended after the survivors (:91-92).
// Returns false on pair-list overflow.
__device__ __forceinline__ bool warp_update(const BBParams& P, Env& e, uint64_t fk, WarpCounters& ct) {
  const BBLayout& L = P.L;
  const int lane = bb_lane();
  const int m = e.nG;
  const uint32_t ltm = bb_lt_mask();
  ct.v[CT_UPB] += (unsigned)m;
  ct.v[CT_UPP] += (unsigned)e.nP;
  if (P.elimination == BB_ELIM_GEBAUERMOELLER) {
    const uint64_t fe = fk & L.ex_mask;
    int w = 0;
    for (int base = 0; base < e.nP; base += 32) {
      int idx = base + lane;
      bool valid = idx < e.nP;
      uint32_t pr = valid ? e.pairs[idx] : 0u;
      bool keep = false;
      if (valid) {
        uint64_t li = e.lm[pr & 0xffffu], lj = e.lm[pr >> 16];
        uint64_t l = bb_lcm_exps(L, li, lj);
        bool drop = bb_divides(L, fe, l) && l != bb_lcm_exps(L, li, fk) && l != bb_lcm_exps(L, lj, fk);
        keep = !drop;
      }
      uint32_t km = __ballot_sync(BB_FULL, keep);
      if (keep) e.pairs[w + __popc(km & ltm)] = pr;  // w + rank <= idx: never overtakes an unread entry of a later chunk
      w += __popc(km);
      __syncwarp();
    }
    e.nP = w;
    for (int i = lane; i < m; i += 32) {
      uint64_t li = e.lm[i];
      e.lscr[i] = bb_lcm_exps(L, li, fk) | (bb_coprime(L, li, fk) ? (1ull << 63) : 0ull);
    }
    __syncwarp();
    for (int base = 0; base < m; base += 32) {
      int i = base + lane;
      bool valid = i < m;
      uint64_t Li = valid ? (e.lscr[i] & L.ex_mask) : 0ull;
      bool bad = false;
      for (int j = 0; j < m; j++) {
        uint64_t Lj = e.lscr[j];
        uint64_t ej = Lj & L.ex_mask;
        if (ej == Li) bad |= (j < i) || (Lj >> 63);
        else bad |= bb_divides(L, ej, Li);
      }
      bool keep = valid && !bad;
      uint32_t km = __ballot_sync(BB_FULL, keep);
      int cnt = __popc(km);
      if (e.nP + cnt > P.max_pairs) return false;
      if (keep) e.pairs[e.nP + __popc(km & ltm)] = ((uint32_t)m << 16) | (uint32_t)i;
      e.nP += cnt;
      ct.v[CT_UPP] += (unsigned)cnt;
    }
  } else {
    for (int base = 0; base < m; base += 32) {
      int i = base + lane;
      bool keep = i < m;
      if (keep && P.elimination == BB_ELIM_LCM) keep = !bb_coprime(L, e.lm[i], fk);  // :58-62
      uint32_t km = __ballot_sync(BB_FULL, keep);
      int cnt = __popc(km);
      if (e.nP + cnt > P.max_pairs) return false;
      if (keep) e.pairs[e.nP + __popc(km & ltm)] = ((uint32_t)m << 16) | (uint32_t)i;
      e.nP += cnt;
      ct.v[CT_UPP] += (unsigned)cnt;
    }
  }
  __syncwarp();
  return true;
}

// Registers the polynomial stored at arena[off, off+len) as basis element m = e.nG: lead data, pair update,
// reducer-list insertion (upper_bound by lead monomial when sort_reducers: after every element whose lead
// monomial is <= the new one, buchberger.cpp:308-311 / 323-326), then nG++, nT += len.
__device__ __forceinline__ bool warp_add_basis(const BBParams& P, Env& e, int off, int len, WarpCounters& ct) {
  const BBLayout& L = P.L;
  const int lane = bb_lane();
  const int m = e.nG;
  if (m >= P.max_basis) { e.status = BB_STATUS_OVERFLOW_BASIS; return false; }
  const uint64_t fk = e.tkey[off];
  const uint32_t lc = e.tcoef[off];
  if (!warp_update(P, e, fk, ct)) { e.status = BB_STATUS_OVERFLOW_PAIRS; return false; }
  int pos = m;
  if (P.sort_reducers) {
    int cnt = 0;  // reducers with LM <= new LM  <=>  key >= new key
    for (int base = 0; base < m; base += 32) {
      int r = base + lane;
      cnt += __popc(__ballot_sync(BB_FULL, r < m && e.rlm[r] >= fk));
    }
    pos = cnt;
    for (int hi = m; hi > pos; hi -= 32) {
      int lo = hi - 32 > pos ? hi - 32 : pos;
      int idx = lo + lane;
      bool v = idx < hi;
      uint64_t k = 0; uint32_t ix = 0;
      if (v) { k = e.rlm[idx]; ix = e.ridx[idx]; }
      __syncwarp();
      if (v) { e.rlm[idx + 1] = k; e.ridx[idx + 1] = ix; }
      __syncwarp();
    }
  }
  if (lane == 0) {
    e.rlm[pos] = fk; e.ridx[pos] = (uint32_t)m;
    e.lm[m] = fk; e.invlc[m] = bb_invmod(L, lc);
    e.pmeta[m] = make_uint2((unsigned)off, (unsigned)len);
  }
  e.nG = m + 1;
  e.nT = off + len;
  __syncwarp();
  return true;
}

// ---------------------------------------------------------------------------------------------------- step
// BuchbergerEnv::step for the pair in row `row` of P (LeadMonomialsEnv::step(int), buchberger.cpp:398-408 ->
// :318-329): erase the pair, s = spoly(G[i], G[j]) (:18-21), (r, steps) = reduce(s, G_), if r != 0 update + sorted
// insert.  Returns the number of polynomial additions 1 + steps (reward = -(1+steps) under Additions, -1 under
// Reductions).  *pi, *pj receive the pair.
__device__ __forceinline__ int warp_step(const BBParams& P, Env& e, int row, int* pi, int* pj, WarpCounters& ct) {
  const BBLayout& L = P.L;
  const int lane = bb_lane();
  if (row < 0 || row >= e.nP) { e.status = BB_STATUS_BAD_ACTION; *pi = -1; *pj = -1; return 0; }
  const uint32_t pr = e.pairs[row];
  const int i = pr & 0xffffu, j = pr >> 16;
  *pi = i; *pj = j;
  // erase the pair, keeping order (:319)
  for (int base = row; base < e.nP - 1; base += 32) {
    int idx = base + lane;
    bool v = idx < e.nP - 1;
    uint32_t x = v ? e.pairs[idx + 1] : 0u;
    __syncwarp();
    if (v) e.pairs[idx] = x;
  }
  e.nP--;
  // S-polynomial: lead terms cancel exactly, so s = (gamma/LT f) tail(f) - (gamma/LT g) tail(g)
  const uint2 mf = e.pmeta[i], mg = e.pmeta[j];
  const uint64_t lf = e.lm[i], lg = e.lm[j];
  const uint64_t gam = bb_lcm(L, lf, lg);
  e.guard |= gam;
  int n = warp_merge(L, e.tkey + mf.x + 1, e.tcoef + mf.x + 1, (int)mf.y - 1, e.invlc[i], gam - lf,
                     e.tkey + mg.x + 1, e.tcoef + mg.x + 1, (int)mg.y - 1, bb_negmod(L, e.invlc[j]), gam - lg,
                     e.hkey, e.hcoef, P.max_poly_terms, e.guard);
  ct.v[CT_STEPS]++;
  if (n < 0) { e.status = BB_STATUS_OVERFLOW_SCRATCH; return 1; }
  if (bb_guard_tripped(L, e.guard)) { e.status = BB_STATUS_OVERFLOW_EXPONENT; return 1; }
  ct.v[CT_TREAD] += mf.y + mg.y;
  ct.v[CT_TWRITE] += (unsigned)n;
  __syncwarp();
  int steps = 0;
  int rlen = warp_reduce(P, e, e.hkey, e.hcoef, n, 0, e.rlm, e.ridx, e.nG, e.tkey + e.nT, e.tcoef + e.nT,
                         P.max_terms - e.nT, steps, ct);
  ct.v[CT_ADDS] += (unsigned)(1 + steps);
  if (rlen == -1) { e.status = BB_STATUS_OVERFLOW_SCRATCH; return 1 + steps; }
  if (rlen == -2) { e.status = BB_STATUS_OVERFLOW_TERMS; return 1 + steps; }
  if (rlen == -3 || bb_guard_tripped(L, e.guard)) { e.status = BB_STATUS_OVERFLOW_EXPONENT; return 1 + steps; }
  if (rlen > 0) {
    ct.v[CT_NONZERO]++;
    if (!warp_add_basis(P, e, e.nT, rlen, ct)) return 1 + steps;
  } else {
    ct.v[CT_ZERO]++;
  }
  if (e.nP == 0) e.status = BB_STATUS_DONE;
  return 1 + steps;
}

// ---------------------------------------------------------------------------------------------------- select
// First / Degree / Normal pair selection (buchberger.cpp:165-186); ties go to the first pair in P, which is
// what the (j,i) tie-break selects because P is always sorted by (j,i).
__device__ __forceinline__ int warp_select(const BBParams& P, const Env& e, int strategy) {
  const BBLayout& L = P.L;
  const int lane = bb_lane();
  if (strategy == BB_SELECT_FIRST || e.nP <= 1) return 0;
  if (strategy == BB_SELECT_DEGREE) {
    uint32_t best = 0xffffffffu;
    for (int idx = lane; idx < e.nP; idx += 32) {
      uint32_t pr = e.pairs[idx];
      uint32_t d = bb_sum_fields(L, bb_lcm_exps(L, e.lm[pr & 0xffffu], e.lm[pr >> 16]));
      uint32_t v = (d << 16) | (uint32_t)idx;
      best = v < best ? v : best;
    }
    best = __reduce_min_sync(BB_FULL, best);
    return (int)(best & 0xffffu);
  }
  // Normal: smallest lcm in grevlex == LARGEST key; lowest row on ties
  uint64_t bk = 0; int bi = 0x7fffffff;
  for (int idx = lane; idx < e.nP; idx += 32) {
    uint32_t pr = e.pairs[idx];
    uint64_t k = bb_lcm(L, e.lm[pr & 0xffffu], e.lm[pr >> 16]);
    if (k > bk) { bk = k; bi = idx; }  // strided ascending idx: first occurrence kept on ties
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    uint64_t ok = bb_shfl64(bk, lane ^ o);
    int oi = __shfl_xor_sync(BB_FULL, bi, o);
    if (ok > bk || (ok == bk && oi < bi)) { bk = ok; bi = oi; }
  }
  return bi;
}

// ---------------------------------------------------------------------------------------------------- observe
// Row r of the state matrix = first k exponent vectors of G[i] then of G[j], zero padded
// (lead_monomials_vector, buchberger.cpp:354-370; rows in P order, :402-406); rows [|P|, pmax) are -1.
__device__ __forceinline__ void warp_observe(const BBParams& P, const Env& e, int32_t* obs, int pmax,
                                             WarpCounters& ct) {
  const BBLayout& L = P.L;
  const int lane = bb_lane();
  const int cols = P.cols, half = P.L.n * P.k, n = P.L.n;
  const int rows = e.nP < pmax ? e.nP : pmax;
  const int live = rows * cols, total = pmax * cols;
  for (int x = lane; x < live; x += 32) {
    int row = x / cols, c = x - row * cols;
    uint32_t pr = e.pairs[row];
    int side = c >= half;
    int cc = c - side * half;
    int t = cc / n, v = cc - t * n;
    uint2 meta = e.pmeta[side ? (pr >> 16) : (pr & 0xffffu)];
    obs[x] = (t < (int)meta.y) ? (int32_t)bb_exp(L, e.tkey[meta.x + t], v) : 0;
  }
  for (int x = live + lane; x < total; x += 32) obs[x] = -1;
  ct.v[CT_OBS] += (unsigned)rows;
}

// ---------------------------------------------------------------------------------------------------- generator
// minstd_rand0 + libstdc++ distributions restated (ideals.h:177-179): the streams must match
// the reference generator bit for bit because "identical seeded inputs" is part of the parity contract.
__device__ __forceinline__ unsigned long long rng_seed(int seed) {
  unsigned long long s = (unsigned long long)(long long)seed % 2147483647ULL;  // int -> unsigned long, then mod m
  return s == 0 ? 1ULL : s;
}
__device__ __forceinline__ unsigned long long rng_next(unsigned long long& x) { x = (x * 16807ULL) % 2147483647ULL; return x; }
// uniform_int_distribution<int>(a,b): "fallback (2 divisions)" branch of bits/uniform_int_dist.h
__device__ __forceinline__ int rng_uniform(unsigned long long& x, int a, int b) {
  const unsigned long long urngrange = 2147483645ULL;
  unsigned long long uerange = (unsigned long long)(unsigned)(b - a) + 1ULL;
  unsigned long long scaling = urngrange / uerange, past = uerange * scaling, ret;
  do ret = rng_next(x) - 1ULL; while (ret >= past);
  return a + (int)(ret / scaling);
}
// generate_canonical<double,53>: two draws, (u1-1) + (u2-1)*R over R*R, all in round-to-nearest double ops
__device__ __forceinline__ double rng_canonical(unsigned long long& x) {
  const double R = 2147483646.0;
  double s = __dmul_rn((double)(rng_next(x) - 1ULL), 1.0);
  s = __dadd_rn(s, __dmul_rn((double)(rng_next(x) - 1ULL), R));
  double r = __ddiv_rn(s, __dmul_rn(R, R));
  if (r >= 1.0) r = __longlong_as_double(0x3FEFFFFFFFFFFFFFLL);  // nextafter(1,0)
  return r;
}
__device__ __forceinline__ int rng_degree(const BBDist& D, unsigned long long& x) {
  if (D.ncp < 2) return 0;
  double p = rng_canonical(x);
  int lo = 0, hi = D.ncp;
  while (lo < hi) { int mid = (lo + hi) >> 1; if (D.cp[mid] < p) lo = mid + 1; else hi = mid; }
  return lo;
}
// RandomBinomialIdealGenerator::next (ideals.cpp:168-201), executed by lane 0 into the slot's staging area.
// Returns false if 1000 trials fail (the reference throws).
__device__ __forceinline__ bool gen_binomial_ideal(const BBParams& P, int slot, unsigned long long& x) {
  const BBDist& D = P.dist;
  uint64_t* ik = P.in_key + (size_t)slot * P.max_gen_terms;
  uint32_t* ic = P.in_coef + (size_t)slot * P.max_gen_terms;
  int* io = P.in_off + (size_t)slot * (P.max_gens + 1);
  io[0] = 0;
  for (int i = 0; i < D.s; i++) {
    uint32_t c = D.pure ? (P.L.p - 1u) : (uint32_t)rng_uniform(x, 1, (int)P.L.p - 1);
    int d1, d2;
    if (D.homogeneous) d1 = d2 = rng_degree(D, x);
    else { d1 = rng_degree(D, x); d2 = rng_degree(D, x); }
    bool ok = false;
    for (int trials = 0; trials < 1000 && !ok; trials++) {
      int n1 = D.basis_off[d1 + 1] - D.basis_off[d1], n2 = D.basis_off[d2 + 1] - D.basis_off[d2];
      uint64_t m1 = D.basis[D.basis_off[d1] + rng_uniform(x, 0, n1 - 1)];
      uint64_t m2 = D.basis[D.basis_off[d2] + rng_uniform(x, 0, n2 - 1)];
      if (m1 != m2) {  // larger monomial (smaller key) leads with coefficient 1
        uint64_t hi = m1 < m2 ? m1 : m2, lo = m1 < m2 ? m2 : m1;
        ik[2 * i] = hi; ic[2 * i] = 1u; ik[2 * i + 1] = lo; ic[2 * i + 1] = c;
        ok = true;
      }
    }
    if (!ok) return false;
    io[i + 1] = 2 * (i + 1);
  }
  P.in_np[slot] = D.s;
  return true;
}

// ---------------------------------------------------------------------------------------------------- reset
// BuchbergerEnv::reset (buchberger.cpp:299-315) from the slot's staged ideal: generators are added one by one
// through update() and into the reducer list.  sort_input orders them by ascending lead monomial first (stable).
// `src` is the staging slot the ideal is read from (never written while fixed ideals are in use).
__device__ __forceinline__ void warp_load_ideal(const BBParams& P, int src_slot, Env& e, WarpCounters& ct) {
  const int lane = bb_lane();
  const uint64_t* ik = P.in_key + (size_t)src_slot * P.max_gen_terms;
  const uint32_t* ic = P.in_coef + (size_t)src_slot * P.max_gen_terms;
  const int* io = P.in_off + (size_t)src_slot * (P.max_gens + 1);
  const int np = P.in_np[src_slot];
  e.nG = 0; e.nP = 0; e.nT = 0; e.guard = 0;
  e.status = BB_STATUS_RUNNING;
  for (int q = 0; q < np; q++) {
    int src = q;
    if (P.sort_input) {  // the generator whose stable ascending-LM rank is q
      int mine = -1;
      for (int g = lane; g < np; g += 32) {
        uint64_t kg = ik[io[g]];
        int rank = 0;
        for (int o = 0; o < np; o++) { uint64_t ko = ik[io[o]]; rank += (ko > kg) || (ko == kg && o < g); }
        if (rank == q) mine = g;
      }
      uint32_t b = __ballot_sync(BB_FULL, mine >= 0);
      src = __shfl_sync(BB_FULL, mine, __ffs(b) - 1);
    }
    const int off = io[src], len = io[src + 1] - off;
    if (e.nT + len > P.max_terms) { e.status = BB_STATUS_OVERFLOW_TERMS; return; }
    for (int t = lane; t < len; t += 32) { e.tkey[e.nT + t] = ik[off + t]; e.tcoef[e.nT + t] = ic[off + t]; }
    __syncwarp();
    if (!warp_add_basis(P, e, e.nT, len, ct)) return;
  }
  if (e.nP == 0) e.status = BB_STATUS_DONE;
}

// Full reset of a slot: draws from the slot's stream when a distribution is set (re-rolling while P comes out
// empty, buchberger.cpp:313-314), else replays the staged ideal.
// rng: the slot's minstd_rand0 state (only lane 0's copy advances); rerolls counts skipped ideals.
// fixed_src: staging slot to replay when no distribution is set.
__device__ __forceinline__ void warp_reset(const BBParams& P, int slot, int fixed_src, Env& e, unsigned long long& rng,
                                           int& rerolls, WarpCounters& ct) {
  const int lane = bb_lane();
  rerolls = 0;
  if (P.dist.enabled) {
    for (;;) {
      int ok = 1;
      if (lane == 0) ok = gen_binomial_ideal(P, slot, rng) ? 1 : 0;
      ok = __shfl_sync(BB_FULL, ok, 0);
      __syncwarp();
      if (!ok) { e.nG = e.nP = e.nT = 0; e.status = BB_STATUS_EMPTY; return; }
      warp_load_ideal(P, slot, e, ct);
      if (e.status != BB_STATUS_DONE) return;
      rerolls++;
    }
  } else {
    warp_load_ideal(P, fixed_src, e, ct);
  }
}

// ---------------------------------------------------------------------------------------------------- hashes
// Position-salted additive checksums (order-sensitive, but computable in parallel on both sides).
__device__ __forceinline__ unsigned long long trace_hash_item(int i, int j, int adds, int t) {
  return bb_hash_item_impl((uint64_t)(uint32_t)i | ((uint64_t)(uint32_t)j << 16) | ((uint64_t)(uint32_t)adds << 32),
                           (uint64_t)t);
}
__device__ __forceinline__ unsigned long long warp_terms_hash(const BBLayout& L, const uint64_t* tk, const uint32_t* tc,
                                                              int nT, const int* lens_or_null, const uint2* meta_or_null,
                                                              int npoly) {
  const int lane = bb_lane();
  unsigned long long h = 0;
  for (int t = lane; t < nT; t += 32) {
    uint64_t k = tk[t];
    uint64_t elo = 0, ehi = 0;
    for (int v = 0; v < L.n; v++) {
      uint64_t x = bb_exp(L, k, v);
      if (v < 4) elo |= x << (16 * v); else ehi |= x << (16 * (v - 4));
    }
    h += bb_hash_item_impl((uint64_t)tc[t], 3ull * t) + bb_hash_item_impl(elo, 3ull * t + 1) +
         bb_hash_item_impl(ehi, 3ull * t + 2);
  }
  for (int p = lane; p < npoly; p += 32) {
    uint64_t len = lens_or_null ? (uint64_t)lens_or_null[p] : (uint64_t)meta_or_null[p].y;
    h += bb_mix64(len + BB_GOLD2 * (uint64_t)(p + 1));
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) h += bb_shfl64(h, lane ^ o);
  return h;
}

// ---------------------------------------------------------------------------------------------------- final GB
// interreduce(minimalize(G)), buchberger.cpp:102-122.  minimalize: G sorted ascending by lead monomial (stable),
// g kept iff no kept lead monomial divides LM g  <=>  no f with LM f strictly dividing LM g and no earlier f
// with the same lead monomial (every lead monomial is divisible by a kept one, by induction along the order).
// interreduce: g <- (1/LC g) * (LT g + reduce(g - LT g, Gmin)), Gmin scanned in ascending-LM order.
// Output goes to the slot's GB arena (gkey/gcoef/glen/gcount); returns false on overflow.  O(m^2/32) once per
// episode; does not modify the environment.
__device__ __forceinline__ bool warp_final_gb(const BBParams& P, int slot, Env& e, WarpCounters& ct) {
  const BBLayout& L = P.L;
  const int lane = bb_lane();
  const int m = e.nG;
  uint64_t* gk = P.gkey + (size_t)slot * P.max_terms;
  uint32_t* gc = P.gcoef + (size_t)slot * P.max_terms;
  int* gl = P.glen + (size_t)slot * P.max_basis;
  uint64_t* rlm2 = P.grlm + (size_t)slot * P.max_basis;    // lead monomials of Gmin, ascending
  uint32_t* ridx2 = P.gridx + (size_t)slot * P.max_basis;  // their basis indices
  uint32_t* flag = P.gflag + (size_t)slot * P.max_basis;   // kept flags by basis index
  for (int base = 0; base < m; base += 32) {
    int g = base + lane;
    if (g < m) {
      uint64_t kg = e.lm[g];
      bool kept = true;
      for (int o = 0; o < m; o++) {
        uint64_t ko = e.lm[o];
        if (ko == kg) kept &= !(o < g);
        else kept &= !bb_divides(L, ko, kg);
      }
      flag[g] = kept ? 1u : 0u;
    }
  }
  __syncwarp();
  int nmin = 0;
  for (int base = 0; base < m; base += 32) {
    int g = base + lane;
    bool kept = g < m && flag[g] != 0u;
    if (kept) {
      uint64_t kg = e.lm[g];
      int rank = 0;  // kept elements with a smaller lead monomial (larger key); kept lead monomials are distinct
      for (int o = 0; o < m; o++) rank += (flag[o] != 0u) && (e.lm[o] > kg);
      rlm2[rank] = kg; ridx2[rank] = (uint32_t)g;
    }
    nmin += __popc(__ballot_sync(BB_FULL, kept));
  }
  __syncwarp();
  int gT = 0;
  for (int q = 0; q < nmin; q++) {
    const int g = (int)ridx2[q];
    const uint2 meta = e.pmeta[g];
    if (gT + 1 > P.max_terms) return false;
    int steps;
    int rlen = warp_reduce(P, e, e.tkey + meta.x + 1, e.tcoef + meta.x + 1, (int)meta.y - 1, -1, rlm2, ridx2, nmin,
                           gk + gT + 1, gc + gT + 1, P.max_terms - gT - 1, steps, ct);
    if (rlen < 0) return false;
    const uint32_t inv = e.invlc[g];
    if (lane == 0) { gk[gT] = e.lm[g]; gc[gT] = 1u; gl[q] = 1 + rlen; }
    for (int t = lane; t < rlen; t += 32) gc[gT + 1 + t] = bb_mulmod(L, gc[gT + 1 + t], inv);
    gT += 1 + rlen;
    __syncwarp();
  }
  if (lane == 0) { P.gcount[2 * slot] = nmin; P.gcount[2 * slot + 1] = gT; }
  __syncwarp();
  return true;
}
