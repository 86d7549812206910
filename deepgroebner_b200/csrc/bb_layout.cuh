// bb_layout.cuh -- grevlex-packed 64-bit monomials with guard bits, and GF(p) arithmetic.
//
// Replaces the reference's Monomial (8 ints + degree, polynomials.h:29-55) and Coefficient
// (polynomials.h:10-26) with register-friendly words.  Usable from host and device code.
//
// key(m) = [ cdeg : dw bits | x_{n-1} : w bits | ... | x_1 : w | x_0 : w ]      (n*w + dw <= 64, left-aligned)
//   cdeg = DMAX - deg(m), DMAX = 2^(dw-1) - 1.  The top bit of EVERY field is a guard bit that is zero in
//   a stored key.  Consequences:
//   * grevlex (polynomials.cpp:60-74: higher degree wins; then from the LAST variable the SMALLER exponent wins)
//     is one unsigned compare:   m1 > m2  <=>  key(m1) < key(m2).  Polynomials keep their terms in ascending
//     key order (= descending monomial order); UINT64_MAX is a natural merge sentinel.
//   * product/quotient (polynomials.cpp:41-57):  key(a*b) = key(a) + key(b) - BIAS,  key(a/b) = key(a) - key(b) + BIAS
//     with BIAS = DMAX << dshift; an exponent or degree that leaves its field sets a guard bit, which is
//     how overflow is DETECTED (status BB_STATUS_OVERFLOW_EXPONENT) instead of wrapping.
//   * divisibility (polynomials.cpp:93-98):  b | a  <=>  (((a & EX) | GE) - (b & EX)) & GE == GE.
//   * lcm (polynomials.cpp:110-118): field-wise max through the same borrow trick, degree re-summed.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BB_HD __host__ __device__ __forceinline__
#else
#define BB_HD inline
#endif

struct BBLayout {
  int n;            // variables
  int w;            // bits per exponent field (incl. guard)
  int dw;           // bits of the degree field (incl. guard)
  int dshift;       // bit position of the degree field  (= 64 - dw)
  int eshift;       // bit position of exponent field 0  (= dshift - n*w)
  uint64_t ex_mask; // all exponent fields (value + guard bits)
  uint64_t ge_mask; // guard bits of the exponent fields
  uint64_t g_all;   // guard bits of every field incl. the degree field (bit 63)
  uint64_t bias;    // DMAX << dshift  == key of the monomial 1
  uint32_t dmax;    // 2^(dw-1) - 1
  uint32_t emax;    // 2^(w-1) - 1, largest storable exponent
  uint32_t fmask;   // (1<<w)-1
  // GF(p)
  uint32_t p;
  uint32_t mu;      // floor(2^32 / p) for Barrett reduction of x < 2^32
};

// Field widths as a function of n alone, shared by the runtime layout (host side) and the compile-time one
// (device side, KL<NV> below) so that the two can never disagree.
constexpr int bb_field_width(int n) { return 64 / (n + 1) > 16 ? 16 : 64 / (n + 1); }
constexpr int bb_degree_width(int n) { return 64 - n * bb_field_width(n) > 16 ? 16 : 64 - n * bb_field_width(n); }
constexpr uint64_t bb_fields_mask(int n, uint64_t per_field) {
  uint64_t m = 0;
  const int w = bb_field_width(n), eshift = 64 - bb_degree_width(n) - n * w;
  for (int i = 0; i < n; i++) m |= per_field << (eshift + i * w);
  return m;
}

inline BBLayout bb_make_layout(int n, uint32_t p) {
  BBLayout L;
  L.n = n;
  L.w = bb_field_width(n);
  L.dw = bb_degree_width(n);
  L.dshift = 64 - L.dw;
  L.eshift = L.dshift - n * L.w;
  L.fmask = (1u << L.w) - 1u;
  L.ex_mask = bb_fields_mask(n, (uint64_t)L.fmask);
  L.ge_mask = bb_fields_mask(n, (uint64_t)1 << (L.w - 1));
  L.g_all = L.ge_mask | ((uint64_t)1 << 63);
  L.dmax = (1u << (L.dw - 1)) - 1u;
  L.emax = (1u << (L.w - 1)) - 1u;
  L.bias = (uint64_t)L.dmax << L.dshift;
  L.p = p;
  L.mu = (uint32_t)(((uint64_t)1 << 32) / p);
  return L;
}

// Compile-time layout for the device code: every mask and shift is an immediate, loops over variables unroll.
template <int NV>
struct KL {
  static constexpr int n = NV;
  static constexpr int w = bb_field_width(NV);
  static constexpr int dw = bb_degree_width(NV);
  static constexpr int dshift = 64 - dw;
  static constexpr int eshift = dshift - NV * w;
  static constexpr uint32_t fmask = (1u << w) - 1u;
  static constexpr uint64_t ex_mask = bb_fields_mask(NV, (uint64_t)fmask);
  static constexpr uint64_t ge_mask = bb_fields_mask(NV, (uint64_t)1 << (w - 1));
  static constexpr uint64_t g_all = ge_mask | ((uint64_t)1 << 63);
  static constexpr uint32_t dmax = (1u << (dw - 1)) - 1u;
  static constexpr uint32_t emax = (1u << (w - 1)) - 1u;
  static constexpr uint64_t bias = (uint64_t)dmax << dshift;

  static BB_HD uint32_t exp(uint64_t k, int i) { return (uint32_t)(k >> (eshift + i * w)) & fmask; }
  static BB_HD uint32_t deg(uint64_t k) { return dmax - (uint32_t)(k >> dshift); }
  // true iff b divides a (degree fields are ignored)
  static BB_HD bool divides(uint64_t b, uint64_t a) {
    return ((((a & ex_mask) | ge_mask) - (b & ex_mask)) & ge_mask) == ge_mask;
  }
  static BB_HD uint64_t lcm_exps(uint64_t a, uint64_t b) {
    uint64_t ea = a & ex_mask, eb = b & ex_mask;
    uint64_t t = ((ea | ge_mask) - eb) & ge_mask;  // guard set where a_i >= b_i
    uint64_t sel = t - (t >> (w - 1));             // value bits of those fields
    return (ea & sel) | (eb & ~sel);
  }
  static BB_HD bool coprime(uint64_t a, uint64_t b) {
    uint64_t ea = a & ex_mask, eb = b & ex_mask;
    uint64_t t = ((ea | ge_mask) - eb) & ge_mask;
    uint64_t sel = t - (t >> (w - 1));
    return ((eb & sel) | (ea & ~sel)) == 0;        // field-wise min is zero everywhere
  }
  static BB_HD uint32_t sum_fields(uint64_t exps) {
    uint32_t s = 0;
    uint64_t x = exps >> eshift;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < NV; i++) { s += (uint32_t)x & fmask; x >>= w; }
    return s;
  }
  static BB_HD uint64_t key_from_exps(uint64_t exps) { return exps | ((uint64_t)(dmax - sum_fields(exps)) << dshift); }
};

// GF(p) parameters as the kernels carry them
struct BBField {
  uint32_t p, mu;
};
BB_HD uint32_t bbf_mulmod(const BBField& F, uint32_t a, uint32_t b) {
  uint32_t x = a * b;  // < 2^32 because p < 2^16
#if defined(__CUDA_ARCH__)
  uint32_t q = __umulhi(x, F.mu);
#else
  uint32_t q = (uint32_t)(((uint64_t)x * F.mu) >> 32);
#endif
  uint32_t r = x - q * F.p;  // in [0, 2p)
  return r >= F.p ? r - F.p : r;
}
BB_HD uint32_t bbf_addmod(const BBField& F, uint32_t a, uint32_t b) { uint32_t r = a + b; return r >= F.p ? r - F.p : r; }
BB_HD uint32_t bbf_negmod(const BBField& F, uint32_t a) { return a ? F.p - a : 0u; }
// a^(p-2): the unique inverse, hence bit-identical to the extended-Euclid inverse of polynomials.cpp:11-23
BB_HD uint32_t bbf_invmod(const BBField& F, uint32_t a) {
  if (a == 1u) return 1u;
  uint32_t r = 1, b = a, e = F.p - 2;
  while (e) {
    if (e & 1u) r = bbf_mulmod(F, r, b);
    b = bbf_mulmod(F, b, b);
    e >>= 1;
  }
  return r;
}

// ---- monomials -----------------------------------------------------------------------------------------------
BB_HD uint64_t bb_pack(const BBLayout& L, const int* e) {
  uint64_t k = 0; uint32_t deg = 0;
  for (int i = 0; i < L.n; i++) { k |= (uint64_t)(uint32_t)e[i] << (L.eshift + i * L.w); deg += (uint32_t)e[i]; }
  return k | ((uint64_t)(L.dmax - deg) << L.dshift);
}
BB_HD uint32_t bb_exp(const BBLayout& L, uint64_t k, int i) { return (uint32_t)(k >> (L.eshift + i * L.w)) & L.fmask; }
BB_HD uint32_t bb_deg(const BBLayout& L, uint64_t k) { return L.dmax - (uint32_t)(k >> L.dshift); }
// true iff b divides a
BB_HD bool bb_divides(const BBLayout& L, uint64_t b, uint64_t a) {
  return ((((a & L.ex_mask) | L.ge_mask) - (b & L.ex_mask)) & L.ge_mask) == L.ge_mask;
}
// exponent fields of lcm(a,b) (no degree field)
BB_HD uint64_t bb_lcm_exps(const BBLayout& L, uint64_t a, uint64_t b) {
  uint64_t ea = a & L.ex_mask, eb = b & L.ex_mask;
  uint64_t t = ((ea | L.ge_mask) - eb) & L.ge_mask;  // guard set where a_i >= b_i
  uint64_t sel = t - (t >> (L.w - 1));               // value bits of those fields
  return (ea & sel) | (eb & ~sel);
}
BB_HD uint32_t bb_sum_fields(const BBLayout& L, uint64_t exps) {
  uint32_t s = 0;
  uint64_t x = exps >> L.eshift;
  for (int i = 0; i < L.n; i++) { s += (uint32_t)x & L.fmask; x >>= L.w; }
  return s;
}
BB_HD uint64_t bb_key_from_exps(const BBLayout& L, uint64_t exps) {
  return exps | ((uint64_t)(L.dmax - bb_sum_fields(L, exps)) << L.dshift);
}
BB_HD uint64_t bb_lcm(const BBLayout& L, uint64_t a, uint64_t b) { return bb_key_from_exps(L, bb_lcm_exps(L, a, b)); }
// gcd(a,b) == 1  <=>  lcm(a,b) == a*b  (the product criterion test of buchberger.cpp:60, 83)
BB_HD bool bb_coprime(const BBLayout& L, uint64_t a, uint64_t b) {
  uint64_t ea = a & L.ex_mask, eb = b & L.ex_mask;
  uint64_t t = ((ea | L.ge_mask) - eb) & L.ge_mask;
  uint64_t sel = t - (t >> (L.w - 1));
  uint64_t mn = (eb & sel) | (ea & ~sel);            // field-wise min
  return mn == 0;
}
// multiplier adjustment: key(a * m) = key(a) + bb_adj(m)   (two's complement wrap is intended)
BB_HD uint64_t bb_adj(const BBLayout& L, uint64_t m) { return m - L.bias; }
// key(a / b); caller guarantees b | a
BB_HD uint64_t bb_quot(const BBLayout& L, uint64_t a, uint64_t b) { return a - b + L.bias; }

// ---- GF(p), p < 2^16: operands canonical in [0,p) -------------------------------------------------------------
BB_HD uint32_t bb_mulmod(const BBLayout& L, uint32_t a, uint32_t b) {
  uint32_t x = a * b;  // < 2^32
#if defined(__CUDA_ARCH__)
  uint32_t q = __umulhi(x, L.mu);
#else
  uint32_t q = (uint32_t)(((uint64_t)x * L.mu) >> 32);
#endif
  uint32_t r = x - q * L.p;  // in [0, 2p)
  return r >= L.p ? r - L.p : r;
}
BB_HD uint32_t bb_addmod(const BBLayout& L, uint32_t a, uint32_t b) {
  uint32_t r = a + b;
  return r >= L.p ? r - L.p : r;
}
BB_HD uint32_t bb_negmod(const BBLayout& L, uint32_t a) { return a ? L.p - a : 0u; }
// a^(p-2): every nonzero residue has exactly one inverse, so this equals the extended-Euclid inverse of
// polynomials.cpp:11-23 bit for bit.
BB_HD uint32_t bb_invmod(const BBLayout& L, uint32_t a) {
  uint32_t r = 1, b = a, e = L.p - 2;
  while (e) {
    if (e & 1u) r = bb_mulmod(L, r, b);
    b = bb_mulmod(L, b, b);
    e >>= 1;
  }
  return r;
}
