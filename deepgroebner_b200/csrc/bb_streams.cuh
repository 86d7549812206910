// bb_streams.cuh -- reduce() for LONG polynomials (cyclic-n and other non-binomial ideals, BASELINE configs[4]) without
// ever materialising the dividend, sm_100a.  What the two stream reducers (bb_wide.cuh: one CTA per environment,
// bb_rstreams.cuh: one warp per environment) share.
//
// warp_reduce / warp_merge (bb_device.cuh) are built for binomial ideals, where a polynomial is two terms.  On cyclic-6
//  a dividend of ~100 (up to ~550) terms is reduced ~83 times per step by polynomials of ~37 terms, and an
// episode is a serial chain of ~80 000 such additions (697 000 in the longest of 1024): materialising
// h <- h - (LT h / LT f) f (buchberger.cpp:33-35) costs O(|h| + |f|) work per addition and, whatever the parallel merge,
// several dependent passes over memory.  Round 2 built and measured three block-wide merges for it (merge path, a
// two-barrier bitmap merge, streams with one barrier per lead term: 1.85 / 1.5 / 1.5 us per addition) before settling on:
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
// One ROUND per lead term M: (a) M's first divisor among the reducer lead monomials (buchberger.cpp:27-32), (b) every
// stream whose head is M advances -- the next raw term was prefetched into registers a round earlier -- and each thread
// takes the minimum over its heads with the coefficient sum at it, (c) REDUX reductions give the next lead term.  A divisor
// f opens a new stream whose head comes straight from f's head record (no dependent load).  Measured on the longest
// cyclic-6 episode (697 000 additions): 955 000 rounds (73 % additions, 21 % remainder terms, 6 % cancelled monomials)
// for 26 M stream terms, i.e. ~27 streams advance per round and a few hundred are live: the monomials of h and of the
// reducers' multiples coincide massively, which is why h stays short while the streams are many.
//
// A stream slot is free when its head key is all ones; a new stream takes a free slot, so the tables stay as small as the
// number of LIVE streams.  If no slot is free, h is consolidated: every pending term is written in order to a scratch list
// behind the slot's term arena and h goes on as ONE stream over it (bb_set_wide(2 / 3 / 5 / 6) cap the slots so that the
// tests reach this).
//
// Episode records are bit-identical to the materialising runner's (bb_set_wide(0): warp_merge / warp_reduce, an
// independent second implementation); of the traffic counters, terms_read / terms_written (|h| per addition) are not
// observable without materialising h and count the reducers' lengths only.
#pragma once
#include "bb_device.cuh"

#ifndef BBS_KMAX
#define BBS_KMAX 1024           // BBRunArgs::stream_kmax when bb_set_wide does not cap the streams
#endif
#define BBS_NONE 0xffffffffu

// The scratch list of a consolidation lies BEHIND the term arena of the slot (bbenv.cu lays hkey right after tkey and
// hcoef right after tcoef, max_terms a multiple of 8), so term index max_terms + i addresses scratch term i through the
// same two base pointers and a stream's cursor needs no flag.  Two halves of max_poly_terms terms: a consolidation writes
// one while the stream over the other is still being read.

__device__ __forceinline__ uint32_t bbf_reduce(const BBField& F, uint32_t x) {   // x mod p for any x < 2^32
  const uint32_t q = __umulhi(x, F.mu);
  const uint32_t r = x - q * F.p;
  return r >= F.p ? r - F.p : r;
}
