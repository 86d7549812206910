/* TEST INFRASTRUCTURE ONLY.
 *
 * bb_oracle.c -- plain-C (C99) CPU restatement of the reference's Buchberger-environment hot path
 * (dylanpeifer/deepgroebner @94f3183e: deepgroebner/{polynomials,ideals,buchberger}.cpp).  It is the
 * CHECKER for the CUDA library, never the thing shipped or measured: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.  The product path (deepgroebner_b200/)
 * never imports, links or executes anything under oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_*.py check this file (a) against every known answer the
 * reference's own tests hold for this path (tests/test_polynomials.cpp, tests/test_buchberger.cpp,
 * tests/test_ideals.cpp, tests/test_buchberger.py -- restated one by one in tests/test_oracle_known_answers.py),
 * (b) against golden traces generated from the UNMODIFIED reference sources compiled here
 * (oracle/_ref/libdgref.so via oracle/ref_shim.cpp; generator script tests/golden/make_golden.py), and
 * (c) function-by-function against oracle/_ref on random inputs whenever that library is present.
 *
 * Every function cites the reference file:line it follows.  Written independently in C (arrays, explicit
 * merges); the reference is C++ with std::vector/std::map.  Two deliberate, documented deviations:
 *   - std::sort (unstable for >16 elements) is restated as a STABLE sort; results can differ from the
 *     reference only when two polynomials share a lead monomial AND more than 16 elements are sorted
 *     (buchberger.cpp:157-158 inside buchberger(); never in BuchbergerEnv::reset/step, which insert with
 *     upper_bound).
 *   - strategies that read std::random_device in the reference ("sample", unseeded "random") need a seed here.
 */
#define _POSIX_C_SOURCE 200809L
#include "bb_oracle.h"

#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define NV ORC_NV
#define W ORC_W

/* ------------------------------------------------------------------ field: polynomials.h:10-26 */
static int PRIME = 32003;

int orc_prime(void) { return PRIME; }
void orc_set_prime(int p) { PRIME = p; }
int orc_nslots(void) { return NV; }

/* Coefficient(int) ctor, polynomials.h:14 */
static int cnorm(long long i) {
  long long r = i % PRIME;
  return (int)((i < 0) ? r + PRIME : r);
}
static int cadd(int a, int b) { return cnorm((long long)a + b); }
static int csub(int a, int b) { return cnorm((long long)a - b); }
static int cmul(int a, int b) { return cnorm((long long)a * b); }
/* operator/ : inverse by extended Euclid, polynomials.cpp:11-23 */
static int cdiv(int c1, int c2) {
  long long a = 0, a_ = 1, b = PRIME, b_ = c2;
  while (b_ != 0) {
    long long q = b / b_, t;
    t = a - q * a_; a = a_; a_ = t;
    t = b - q * b_; b = b_; b_ = t;
  }
  return cnorm((long long)c1 * a);
}
int orc_coef_div(int a, int b) { return cdiv(cnorm(a), cnorm(b)); }
int orc_coef_mul(int a, int b) { return cmul(cnorm(a), cnorm(b)); }
int orc_coef_add(int a, int b) { return cadd(cnorm(a), cnorm(b)); }
int orc_coef_sub(int a, int b) { return csub(cnorm(a), cnorm(b)); }
int orc_coef_norm(int a) { return cnorm(a); }

/* ------------------------------------------------------------------ monomials: polynomials.h:29-55 */
typedef struct { int e[NV]; int deg; } Mono;
typedef struct { int c; Mono m; } Term;

static Mono mono_from(const int* e) {
  Mono m; m.deg = 0;
  for (int k = 0; k < NV; k++) { m.e[k] = e[k]; m.deg += e[k]; }
  return m;
}
static Mono mono_one(void) { Mono m; memset(&m, 0, sizeof m); return m; }
/* operator*, polynomials.cpp:41-47 */
static Mono mono_mul(const Mono* a, const Mono* b) {
  Mono m; m.deg = a->deg + b->deg;
  for (int k = 0; k < NV; k++) m.e[k] = a->e[k] + b->e[k];
  return m;
}
/* operator/, polynomials.cpp:50-57 */
static Mono mono_div(const Mono* a, const Mono* b) {
  Mono m; m.deg = a->deg - b->deg;
  for (int k = 0; k < NV; k++) m.e[k] = a->e[k] - b->e[k];
  return m;
}
/* grevlex operator>, polynomials.cpp:60-74: degree first, then from the LAST variable the SMALLER exponent wins */
static int mono_gt(const Mono* a, const Mono* b) {
  if (a->deg > b->deg) return 1;
  if (b->deg > a->deg) return 0;
  for (int k = NV - 1; k >= 0; k--) {
    if (b->e[k] > a->e[k]) return 1;
    if (a->e[k] > b->e[k]) return 0;
  }
  return 0;
}
static int mono_lt(const Mono* a, const Mono* b) { return mono_gt(b, a); }
/* operator==, polynomials.cpp:77-81 (ignores the cached degree) */
static int mono_eq(const Mono* a, const Mono* b) {
  for (int k = 0; k < NV; k++) if (a->e[k] != b->e[k]) return 0;
  return 1;
}
/* is_divisible(m1, m2): m2 divides m1, polynomials.cpp:93-98 */
static int mono_divisible(const Mono* a, const Mono* b) {
  for (int k = 0; k < NV; k++) if (a->e[k] < b->e[k]) return 0;
  return 1;
}
/* lcm, polynomials.cpp:110-118 */
static Mono mono_lcm(const Mono* a, const Mono* b) {
  Mono m; m.deg = 0;
  for (int k = 0; k < NV; k++) { m.e[k] = a->e[k] > b->e[k] ? a->e[k] : b->e[k]; m.deg += m.e[k]; }
  return m;
}
int orc_mono_cmp(const int* e1, const int* e2) {
  Mono a = mono_from(e1), b = mono_from(e2);
  return mono_gt(&a, &b) ? 1 : (mono_gt(&b, &a) ? -1 : 0);
}
int orc_mono_divisible(const int* e1, const int* e2) {
  Mono a = mono_from(e1), b = mono_from(e2);
  return mono_divisible(&a, &b);
}
void orc_mono_lcm(const int* e1, const int* e2, int* out) {
  Mono a = mono_from(e1), b = mono_from(e2), m = mono_lcm(&a, &b);
  for (int k = 0; k < NV; k++) out[k] = m.e[k];
}

/* ------------------------------------------------------------------ polynomials: polynomials.h:71-94 */
typedef struct { Term* t; int n, cap; int sug; } Poly;

static void poly_init(Poly* f) { f->t = NULL; f->n = 0; f->cap = 0; f->sug = 0; }
static void poly_free(Poly* f) { free(f->t); poly_init(f); }
static void poly_push(Poly* f, Term t) {
  if (f->n == f->cap) {
    f->cap = f->cap ? 2 * f->cap : 4;
    f->t = (Term*)realloc(f->t, sizeof(Term) * (size_t)f->cap);
  }
  f->t[f->n++] = t;
}
static Poly poly_copy(const Poly* f) {
  Poly g; poly_init(&g);
  if (f->n) {
    g.t = (Term*)malloc(sizeof(Term) * (size_t)f->n);
    memcpy(g.t, f->t, sizeof(Term) * (size_t)f->n);
  }
  g.n = g.cap = f->n; g.sug = f->sug;
  return g;
}
static void poly_move(Poly* dst, Poly* src) { poly_free(dst); *dst = *src; poly_init(src); }

/* Polynomial(vector<Term>) ctor, polynomials.cpp:139-145: sort descending, sugar = deg LM; no merging of
 * equal monomials, no dropping of zeros.  (stable insertion sort; std::sort is unspecified on ties) */
static Poly poly_make(const Term* t, int n) {
  Poly f; poly_init(&f);
  for (int i = 0; i < n; i++) poly_push(&f, t[i]);
  for (int i = 1; i < n; i++) {
    Term x = f.t[i]; int j = i - 1;
    while (j >= 0 && mono_gt(&x.m, &f.t[j].m)) { f.t[j + 1] = f.t[j]; j--; }
    f.t[j + 1] = x;
  }
  f.sug = n ? f.t[0].m.deg : 0;
  return f;
}
static Poly poly_from_term(Term t) { return poly_make(&t, 1); }

/* operator+, polynomials.cpp:148-177: two-pointer merge dropping cancelled terms, sugar = max */
static Poly poly_add(const Poly* f1, const Poly* f2) {
  Poly g; poly_init(&g);
  g.sug = f1->sug > f2->sug ? f1->sug : f2->sug;
  int i = 0, j = 0;
  while (i < f1->n && j < f2->n) {
    const Term* t1 = &f1->t[i]; const Term* t2 = &f2->t[j];
    if (mono_gt(&t1->m, &t2->m)) { poly_push(&g, *t1); i++; }
    else if (mono_gt(&t2->m, &t1->m)) { poly_push(&g, *t2); j++; }
    else {
      int c = cadd(t1->c, t2->c);
      if (c != 0) { Term t = *t1; t.c = c; poly_push(&g, t); }
      i++; j++;
    }
  }
  for (; i < f1->n; i++) poly_push(&g, f1->t[i]);
  for (; j < f2->n; j++) poly_push(&g, f2->t[j]);
  return g;
}
/* operator-, polynomials.cpp:180-185 */
static Poly poly_sub(const Poly* f1, const Poly* f2) {
  Poly f = poly_copy(f2);
  for (int i = 0; i < f.n; i++) f.t[i].c = cmul(cnorm(-1), f.t[i].c);
  Poly g = poly_add(f1, &f);
  poly_free(&f);
  return g;
}
/* Term * Polynomial, polynomials.cpp:196-202 */
static Poly term_mul_poly(const Term* t, const Poly* f) {
  Poly g; poly_init(&g);
  g.sug = t->m.deg + f->sug;
  for (int i = 0; i < f->n; i++) {
    Term x; x.c = cmul(t->c, f->t[i].c); x.m = mono_mul(&t->m, &f->t[i].m);
    poly_push(&g, x);
  }
  return g;
}
/* Polynomial * Polynomial, polynomials.cpp:205-210 */
static Poly poly_mul(const Poly* f1, const Poly* f2) {
  Poly g; poly_init(&g);
  for (int i = 0; i < f1->n; i++) {
    Poly p = term_mul_poly(&f1->t[i], f2);
    Poly s = poly_add(&g, &p);
    poly_free(&p); poly_move(&g, &s);
  }
  return g;
}
/* Term / Term, polynomials.h:61 */
static Term term_div(const Term* a, const Term* b) {
  Term t; t.c = cdiv(a->c, b->c); t.m = mono_div(&a->m, &b->m);
  return t;
}

/* ---- wire helpers */
static Poly poly_from_wire(const int* t, int n) {
  Term* tv = (Term*)malloc(sizeof(Term) * (size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; i++) { tv[i].c = cnorm(t[i * W]); tv[i].m = mono_from(t + i * W + 1); }
  Poly f = poly_make(tv, n);
  free(tv);
  return f;
}
static int poly_to_wire(const Poly* f, int* out, int cap) {
  if (f->n > cap) return -f->n;
  for (int i = 0; i < f->n; i++) {
    out[i * W] = f->t[i].c;
    for (int k = 0; k < NV; k++) out[i * W + 1 + k] = f->t[i].m.e[k];
  }
  return f->n;
}

typedef struct { Poly* p; int n, cap; } PolyVec;
static void pv_init(PolyVec* v) { v->p = NULL; v->n = 0; v->cap = 0; }
static void pv_clear(PolyVec* v) { for (int i = 0; i < v->n; i++) poly_free(&v->p[i]); v->n = 0; }
static void pv_free(PolyVec* v) { pv_clear(v); free(v->p); pv_init(v); }
static void pv_reserve(PolyVec* v) {
  if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 8; v->p = (Poly*)realloc(v->p, sizeof(Poly) * (size_t)v->cap); }
}
static void pv_push_copy(PolyVec* v, const Poly* f) { pv_reserve(v); v->p[v->n++] = poly_copy(f); }
static void pv_insert_copy(PolyVec* v, int at, const Poly* f) {
  pv_reserve(v);
  memmove(&v->p[at + 1], &v->p[at], sizeof(Poly) * (size_t)(v->n - at));
  v->p[at] = poly_copy(f); v->n++;
}
static PolyVec pv_copy(const PolyVec* v) {
  PolyVec o; pv_init(&o);
  for (int i = 0; i < v->n; i++) pv_push_copy(&o, &v->p[i]);
  return o;
}
static PolyVec pv_from_wire(const int* terms, const int* lens, int npoly) {
  PolyVec v; pv_init(&v);
  int off = 0;
  for (int p = 0; p < npoly; p++) {
    Poly f = poly_from_wire(terms + off * W, lens[p]);
    pv_reserve(&v); v.p[v.n++] = f;
    off += lens[p];
  }
  return v;
}
static int pv_to_wire(const PolyVec* v, int* terms, int cap_terms, int* lens, int cap_polys) {
  if (v->n > cap_polys) return -1;
  int off = 0;
  for (int p = 0; p < v->n; p++) {
    int n = poly_to_wire(&v->p[p], terms + off * W, cap_terms - off);
    if (n < 0) return -1;
    lens[p] = n; off += n;
  }
  return v->n;
}

int orc_poly_make(const int* t, int n, int* out, int cap) {
  Poly f = poly_from_wire(t, n); int r = poly_to_wire(&f, out, cap); poly_free(&f); return r;
}
int orc_poly_add(const int* f, int nf, const int* g, int ng, int* out, int cap) {
  Poly a = poly_from_wire(f, nf), b = poly_from_wire(g, ng), s = poly_add(&a, &b);
  int r = poly_to_wire(&s, out, cap); poly_free(&a); poly_free(&b); poly_free(&s); return r;
}
int orc_poly_sub(const int* f, int nf, const int* g, int ng, int* out, int cap) {
  Poly a = poly_from_wire(f, nf), b = poly_from_wire(g, ng), s = poly_sub(&a, &b);
  int r = poly_to_wire(&s, out, cap); poly_free(&a); poly_free(&b); poly_free(&s); return r;
}
int orc_poly_mul(const int* f, int nf, const int* g, int ng, int* out, int cap) {
  Poly a = poly_from_wire(f, nf), b = poly_from_wire(g, ng), s = poly_mul(&a, &b);
  int r = poly_to_wire(&s, out, cap); poly_free(&a); poly_free(&b); poly_free(&s); return r;
}
int orc_term_mul(const int* t, const int* f, int nf, int* out, int cap) {
  Term tt; tt.c = cnorm(t[0]); tt.m = mono_from(t + 1);
  Poly a = poly_from_wire(f, nf), s = term_mul_poly(&tt, &a);
  int r = poly_to_wire(&s, out, cap); poly_free(&a); poly_free(&s); return r;
}

/* parse_polynomial, polynomials.cpp:226-300: variables a..h, no spaces, no error checking.
 * Grammar restated as an iterative scanner: polynomial = sum of signed terms; each term = [int][*]mono. */
static Mono parse_mono(const char** s) {
  Mono m = mono_one();
  for (;;) {
    int c = **s;
    if (c < 'a' || c > 'a' + NV) break;
    int var = c - 'a'; (*s)++;
    int power = 1;
    if (**s == '^') { (*s)++; power = (int)strtol(*s, (char**)s, 10); }
    if (var < NV) { m.e[var] += power; m.deg += power; }
    if (**s == '*') (*s)++; else break;
  }
  return m;
}
int orc_parse_polynomial(const char* s, int* out, int cap) {
  Poly acc; poly_init(&acc);
  /* the reference builds Polynomial{t} + parse(rest) right-to-left; addition is commutative/associative on the
   * normal form, so a left fold yields the same polynomial */
  while (*s) {
    int sign = 1;
    while (*s == '+' || *s == '-') { if (*s == '-') sign = -sign; s++; }
    Term t; t.c = 1; t.m = mono_one();
    if (*s >= '0' && *s <= '9') {
      t.c = cnorm(strtol(s, (char**)&s, 10));
      if (*s == '*') { s++; t.m = parse_mono(&s); }
    } else {
      t.m = parse_mono(&s);
    }
    if (sign < 0) t.c = cmul(cnorm(-1), t.c);
    Poly p = poly_from_term(t), sum = poly_add(&acc, &p);
    poly_free(&p); poly_move(&acc, &sum);
  }
  int r = poly_to_wire(&acc, out, cap); poly_free(&acc); return r;
}

/* ------------------------------------------------------------------ spoly: buchberger.cpp:18-21 */
static Poly spoly(const Poly* f, const Poly* g) {
  Term gamma; gamma.c = cnorm(1); gamma.m = mono_lcm(&f->t[0].m, &g->t[0].m);
  Term tf = term_div(&gamma, &f->t[0]), tg = term_div(&gamma, &g->t[0]);
  Poly a = term_mul_poly(&tf, f), b = term_mul_poly(&tg, g), s = poly_sub(&a, &b);
  poly_free(&a); poly_free(&b);
  return s;
}
int orc_spoly(const int* f, int nf, const int* g, int ng, int* out, int cap) {
  Poly a = poly_from_wire(f, nf), b = poly_from_wire(g, ng), s = spoly(&a, &b);
  int r = poly_to_wire(&s, out, cap); poly_free(&a); poly_free(&b); poly_free(&s); return r;
}

/* ------------------------------------------------------------------ reduce: buchberger.cpp:24-49
 * lead-term-at-a-time division; divisor = FIRST f in F (in order) whose LM divides LM h; steps counts
 * reductions only, not term moves. */
static Poly reduce(const Poly* g, const PolyVec* F, int* steps_out) {
  int steps = 0;
  Poly r; poly_init(&r);
  Poly h = poly_copy(g);
  while (h.n != 0) {
    int found = 0;
    for (int k = 0; k < F->n; k++) {
      const Poly* f = &F->p[k];
      if (mono_divisible(&h.t[0].m, &f->t[0].m)) {
        Term q = term_div(&h.t[0], &f->t[0]);
        Poly qf = term_mul_poly(&q, f), nh = poly_sub(&h, &qf);
        poly_free(&qf); poly_move(&h, &nh);
        found = 1; steps++;
        break;
      }
    }
    if (!found) {
      Poly lt = poly_from_term(h.t[0]);
      Poly nr = poly_add(&r, &lt), nh = poly_sub(&h, &lt);
      poly_free(&lt); poly_move(&r, &nr); poly_move(&h, &nh);
    }
  }
  Poly out = poly_add(&r, &h);
  poly_free(&r); poly_free(&h);
  *steps_out = steps;
  return out;
}
int orc_reduce(const int* g, int ng, const int* Fterms, const int* Flens, int nF, int* out, int cap, int* steps) {
  Poly a = poly_from_wire(g, ng);
  PolyVec F = pv_from_wire(Fterms, Flens, nF);
  Poly r = reduce(&a, &F, steps);
  int n = poly_to_wire(&r, out, cap);
  poly_free(&a); poly_free(&r); pv_free(&F);
  return n;
}

/* ------------------------------------------------------------------ update: buchberger.cpp:52-99 */
typedef struct { int i, j; } SPair;
typedef struct { SPair* p; int n, cap; } PairVec;
static void pairs_init(PairVec* v) { v->p = NULL; v->n = 0; v->cap = 0; }
static void pairs_free(PairVec* v) { free(v->p); pairs_init(v); }
static void pairs_push(PairVec* v, SPair s) {
  if (v->n == v->cap) { v->cap = v->cap ? 2 * v->cap : 16; v->p = (SPair*)realloc(v->p, sizeof(SPair) * (size_t)v->cap); }
  v->p[v->n++] = s;
}
static PairVec pairs_copy(const PairVec* v) {
  PairVec o; pairs_init(&o);
  for (int i = 0; i < v->n; i++) pairs_push(&o, v->p[i]);
  return o;
}

/* Appends f to G and updates P.  GebauerMoeller (0): (1) drop old (i,j) whose lcm is divisible by LM f and
 * differs from both lcm(LM_i,LM f), lcm(LM_j,LM f) (:63-70); (2) group i by lcm(LM_i, LM f) in ascending
 * grevlex order, members ascending in i (the std::map of :72-75); (3) ascending over groups keep an lcm only if
 * no previously kept lcm divides it (:77-81); (4) emit (v[0], m) unless some member of the group is coprime to
 * f (:82-83); (5) sort new pairs by i (:86) and append (:91-92).  LCM (1): (i,m) unless coprime (:58-62).
 * None (2): every (i,m) (:53-57). */
static void update(PolyVec* G, PairVec* P, const Poly* f, int elimination) {
  int m = G->n;
  PairVec Pn; pairs_init(&Pn);
  const Mono* lf = &f->t[0].m;
  if (elimination == 2) {
    for (int i = 0; i < m; i++) { SPair s = {i, m}; pairs_push(&Pn, s); }
  } else if (elimination == 1) {
    for (int i = 0; i < m; i++) {
      Mono l = mono_lcm(&G->p[i].t[0].m, lf), pr = mono_mul(&G->p[i].t[0].m, lf);
      if (!mono_eq(&l, &pr)) { SPair s = {i, m}; pairs_push(&Pn, s); }
    }
  } else {
    int w = 0;
    for (int k = 0; k < P->n; k++) {
      SPair p = P->p[k];
      Mono l = mono_lcm(&G->p[p.i].t[0].m, &G->p[p.j].t[0].m);
      Mono li = mono_lcm(&G->p[p.i].t[0].m, lf), lj = mono_lcm(&G->p[p.j].t[0].m, lf);
      int drop = mono_divisible(&l, lf) && !mono_eq(&l, &li) && !mono_eq(&l, &lj);
      if (!drop) P->p[w++] = p;
    }
    P->n = w;

    Mono* L = (Mono*)malloc(sizeof(Mono) * (size_t)(m > 0 ? m : 1));
    int* idx = (int*)malloc(sizeof(int) * (size_t)(m > 0 ? m : 1));
    for (int i = 0; i < m; i++) { L[i] = mono_lcm(&G->p[i].t[0].m, lf); idx[i] = i; }
    for (int a = 1; a < m; a++) { /* stable insertion sort of indices by lcm ascending == std::map order */
      int x = idx[a], b = a - 1;
      while (b >= 0 && mono_lt(&L[x], &L[idx[b]])) { idx[b + 1] = idx[b]; b--; }
      idx[b + 1] = x;
    }
    Mono* kept = (Mono*)malloc(sizeof(Mono) * (size_t)(m > 0 ? m : 1));
    int nk = 0;
    for (int a = 0; a < m;) {
      int b = a;
      while (b < m && mono_eq(&L[idx[b]], &L[idx[a]])) b++;
      const Mono* mon = &L[idx[a]];
      int minimal = 1;
      for (int q = 0; q < nk; q++) if (mono_divisible(mon, &kept[q])) { minimal = 0; break; }
      if (minimal) {
        kept[nk++] = *mon;
        int coprime = 0;
        for (int q = a; q < b; q++) {
          Mono pr = mono_mul(&G->p[idx[q]].t[0].m, lf);
          if (mono_eq(&L[idx[q]], &pr)) { coprime = 1; break; }
        }
        if (!coprime) { SPair s = {idx[a], m}; pairs_push(&Pn, s); }
      }
      a = b;
    }
    for (int a = 1; a < Pn.n; a++) { /* sort new pairs by i (all distinct) */
      SPair x = Pn.p[a]; int b = a - 1;
      while (b >= 0 && x.i < Pn.p[b].i) { Pn.p[b + 1] = Pn.p[b]; b--; }
      Pn.p[b + 1] = x;
    }
    free(L); free(idx); free(kept);
  }
  pv_push_copy(G, f);
  for (int k = 0; k < Pn.n; k++) pairs_push(P, Pn.p[k]);
  pairs_free(&Pn);
}
int orc_update(const int* Gterms, const int* Glens, int nG, int* pairs, int nP, int cap_pairs, const int* f, int nf,
               int elimination) {
  PolyVec G = pv_from_wire(Gterms, Glens, nG);
  PairVec P; pairs_init(&P);
  for (int i = 0; i < nP; i++) { SPair s = {pairs[2 * i], pairs[2 * i + 1]}; pairs_push(&P, s); }
  Poly ff = poly_from_wire(f, nf);
  update(&G, &P, &ff, elimination);
  int n = P.n;
  if (n > cap_pairs) n = -1;
  else for (int i = 0; i < P.n; i++) { pairs[2 * i] = P.p[i].i; pairs[2 * i + 1] = P.p[i].j; }
  poly_free(&ff); pv_free(&G); pairs_free(&P);
  return n;
}

/* stable sort of a PolyVec ascending by LM (restates std::sort(.., LM <), see header note) */
static void pv_sort_by_lm(PolyVec* v) {
  for (int a = 1; a < v->n; a++) {
    Poly x = v->p[a]; int b = a - 1;
    while (b >= 0 && mono_lt(&x.t[0].m, &v->p[b].t[0].m)) { v->p[b + 1] = v->p[b]; b--; }
    v->p[b + 1] = x;
  }
}
/* upper_bound insert by LM: after every element whose LM is <= LM f (buchberger.cpp:308-311, 323-326) */
static void pv_insert_sorted(PolyVec* v, const Poly* f) {
  int lo = 0, hi = v->n;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (mono_lt(&f->t[0].m, &v->p[mid].t[0].m)) hi = mid; else lo = mid + 1;
  }
  pv_insert_copy(v, lo, f);
}

/* ------------------------------------------------------------------ minimalize / interreduce: buchberger.cpp:102-122 */
static PolyVec minimalize(const PolyVec* G) {
  PolyVec S = pv_copy(G), out; pv_init(&out);
  pv_sort_by_lm(&S);
  for (int a = 0; a < S.n; a++) {
    int div = 0;
    for (int b = 0; b < out.n; b++) if (mono_divisible(&S.p[a].t[0].m, &out.p[b].t[0].m)) { div = 1; break; }
    if (!div) pv_push_copy(&out, &S.p[a]);
  }
  pv_free(&S);
  return out;
}
static PolyVec interreduce(const PolyVec* G) {
  PolyVec out; pv_init(&out);
  for (int a = 0; a < G->n; a++) {
    const Poly* g = &G->p[a];
    Term inv; inv.c = cdiv(cnorm(1), g->t[0].c); inv.m = mono_one();
    Poly lt = poly_from_term(g->t[0]);
    Poly tail = poly_sub(g, &lt);
    int steps;
    Poly r = reduce(&tail, G, &steps);
    Poly s = poly_add(&r, &lt);
    Poly q = term_mul_poly(&inv, &s);
    pv_reserve(&out); out.p[out.n++] = q;
    poly_free(&lt); poly_free(&tail); poly_free(&r); poly_free(&s);
  }
  return out;
}
int orc_minimalize(const int* Gterms, const int* Glens, int nG, int* oterms, int cap_terms, int* olens, int cap_polys) {
  PolyVec G = pv_from_wire(Gterms, Glens, nG), M = minimalize(&G);
  int r = pv_to_wire(&M, oterms, cap_terms, olens, cap_polys);
  pv_free(&G); pv_free(&M); return r;
}
int orc_interreduce(const int* Gterms, const int* Glens, int nG, int* oterms, int cap_terms, int* olens, int cap_polys) {
  PolyVec G = pv_from_wire(Gterms, Glens, nG), M = interreduce(&G);
  int r = pv_to_wire(&M, oterms, cap_terms, olens, cap_polys);
  pv_free(&G); pv_free(&M); return r;
}

/* ------------------------------------------------------------------ libstdc++ <random> restated
 * std::default_random_engine = minstd_rand0: x <- 16807 x mod (2^31-1); seed(s): s mod m, 0 -> 1 */
typedef struct { uint64_t x; } Rng;
static void rng_seed(Rng* r, int seed) {
  uint64_t s = (uint64_t)(unsigned long)(long)seed; /* int -> unsigned long conversion as result_type */
  s %= 2147483647ULL;
  r->x = s == 0 ? 1 : s;
}
static uint64_t rng_next(Rng* r) { r->x = (r->x * 16807ULL) % 2147483647ULL; return r->x; }
/* uniform_int_distribution<int>(a,b), libstdc++ "fallback (2 divisions)" branch (bits/uniform_int_dist.h) */
static int rng_uniform_int(Rng* r, int a, int b) {
  const uint64_t urngrange = 2147483646ULL - 1ULL;
  uint64_t urange = (uint64_t)((unsigned)b - (unsigned)a);
  uint64_t ret;
  if (urngrange > urange) {
    uint64_t uerange = urange + 1, scaling = urngrange / uerange, past = uerange * scaling;
    do ret = rng_next(r) - 1ULL; while (ret >= past);
    ret /= scaling;
  } else if (urngrange == urange) {
    ret = rng_next(r) - 1ULL;
  } else {
    return a; /* upscaling never occurs on this path (ranges are tiny) */
  }
  return (int)(ret + (uint64_t)a);
}
/* generate_canonical<double,53>: two engine draws (bits/random.tcc:3349-3381) */
static double rng_canonical(Rng* r) {
  const double R = 2147483646.0;
  double sum = 0.0, tmp = 1.0;
  for (int k = 0; k < 2; k++) {
    sum += (double)(rng_next(r) - 1ULL) * tmp;
    tmp *= R;
  }
  double ret = sum / tmp;
  if (ret >= 1.0) ret = nextafter(1.0, 0.0);
  return ret;
}
/* discrete_distribution: normalise, partial sums, last = 1.0; sample = lower_bound(cp, canonical) (random.tcc:2657-2714) */
typedef struct { double* cp; double* prob; int n; } Discrete;
static void discrete_init(Discrete* d, const int* count, int n) {
  d->n = n; d->cp = NULL; d->prob = (double*)malloc(sizeof(double) * (size_t)n);
  for (int i = 0; i < n; i++) d->prob[i] = (double)count[i];
  if (n < 2) { d->n = 0; return; } /* degenerate: always 0 */
  double sum = 0.0;
  for (int i = 0; i < n; i++) sum += d->prob[i];
  for (int i = 0; i < n; i++) d->prob[i] /= sum;
  d->cp = (double*)malloc(sizeof(double) * (size_t)n);
  double acc = d->prob[0]; d->cp[0] = acc;
  for (int i = 1; i < n; i++) { acc = acc + d->prob[i]; d->cp[i] = acc; }
  d->cp[n - 1] = 1.0;
}
static void discrete_free(Discrete* d) { free(d->cp); free(d->prob); d->cp = d->prob = NULL; }
static int discrete_sample(const Discrete* d, Rng* r) {
  if (d->n == 0) return 0;
  double p = rng_canonical(r);
  int lo = 0, hi = d->n;
  while (lo < hi) { int mid = (lo + hi) / 2; if (d->cp[mid] < p) lo = mid + 1; else hi = mid; }
  return lo;
}
/* poisson_distribution for mean < 12 (random.tcc: multiply canonicals until below exp(-mean)) */
static int poisson_sample(double mean, Rng* r) {
  double thr = exp(-mean), prod = 1.0; int x = 0;
  do { prod *= rng_canonical(r); x += 1; } while (prod > thr);
  return x - 1;
}

/* ------------------------------------------------------------------ generators: ideals.cpp */
/* binomial, ideals.cpp:67-72 */
static int binomial(int n, int k) { return (k == 0 || k == n) ? 1 : binomial(n - 1, k - 1) + binomial(n - 1, k); }

/* basis(n,d), ideals.cpp:39-64: next_permutation over d stars then n-1 bars == exponent vectors in
 * lex-descending order.  Restated as a direct recursive enumeration of that order. */
typedef struct { Mono* m; int n, cap; } MonoVec;
static void basis_rec(MonoVec* B, int n, int var, int left, int* e) {
  if (var == n - 1) {
    e[var] = left;
    if (B->n == B->cap) { B->cap = B->cap ? 2 * B->cap : 16; B->m = (Mono*)realloc(B->m, sizeof(Mono) * (size_t)B->cap); }
    int full[NV] = {0};
    for (int k = 0; k < n; k++) full[k] = e[k];
    B->m[B->n++] = mono_from(full);
    return;
  }
  for (int x = left; x >= 0; x--) { e[var] = x; basis_rec(B, n, var + 1, left - x, e); }
}
static MonoVec basis(int n, int d) {
  MonoVec B; B.m = NULL; B.n = 0; B.cap = 0;
  int e[NV] = {0};
  basis_rec(&B, n, 0, d, e);
  return B;
}
int orc_basis(int n, int d, int* out, int cap) {
  MonoVec B = basis(n, d);
  int r = B.n;
  if (B.n > cap) r = -B.n;
  else for (int i = 0; i < B.n; i++) for (int k = 0; k < NV; k++) out[i * NV + k] = B.m[i].e[k];
  free(B.m);
  return r;
}
/* degree_distribution, ideals.cpp:75-100; dist: 0 Uniform 1 Weighted 2 Maximum */
static int degree_counts(int n, int d, int dist, int constants, int* count) {
  int c = 0;
  count[c++] = constants ? 1 : 0;
  if (dist == 0) for (int i = 1; i < d + 1; i++) count[c++] = binomial(n + i - 1, n - 1);
  else if (dist == 1) for (int i = 0; i < d; i++) count[c++] = 1;
  else { for (int i = 0; i < d - 1; i++) count[c++] = 0; count[c++] = 1; }
  return c;
}
int orc_degree_distribution(int n, int d, int dist, int constants, double* out, int cap) {
  int* count = (int*)malloc(sizeof(int) * (size_t)(d + 2));
  int c = degree_counts(n, d, dist, constants, count);
  Discrete dd; discrete_init(&dd, count, c);
  int r = c;
  if (c > cap) r = -1; else for (int i = 0; i < c; i++) out[i] = dd.prob[i];
  discrete_free(&dd); free(count);
  return r;
}
/* cyclic(n), ideals.cpp:16-36 */
static PolyVec cyclic(int n) {
  PolyVec F; pv_init(&F);
  for (int d = 1; d < n; d++) {
    Term* p = (Term*)malloc(sizeof(Term) * (size_t)n);
    for (int i = 0; i < n; i++) {
      int e[NV] = {0};
      for (int k = 0; k < d; k++) e[(i + k) % n] = 1;
      p[i].c = cnorm(1); p[i].m = mono_from(e);
    }
    Poly f = poly_make(p, n);
    pv_reserve(&F); F.p[F.n++] = f;
    free(p);
  }
  int e[NV] = {0};
  for (int i = 0; i < n; i++) e[i] = 1;
  Term t2[2];
  t2[0].c = cnorm(1); t2[0].m = mono_from(e);
  t2[1].c = cnorm(-1); t2[1].m = mono_one();
  Poly f = poly_make(t2, 2);
  pv_reserve(&F); F.p[F.n++] = f;
  return F;
}
int orc_cyclic(int n, int* oterms, int cap_terms, int* olens, int cap_polys) {
  PolyVec F = cyclic(n);
  int r = pv_to_wire(&F, oterms, cap_terms, olens, cap_polys);
  pv_free(&F); return r;
}

enum { GEN_FIXED = 0, GEN_BINOMIAL = 1, GEN_RANDOM = 2 };
typedef struct {
  int kind, n, s, homogeneous, pure, d;
  double lam;
  MonoVec* bases; /* bases[0..d] */
  Discrete degree_dist;
  Rng rng;
  PolyVec F; /* fixed */
} Gen;

/* FixedIdealGenerator ctor, ideals.cpp:146-154 -- quirk Q1: n = largest variable INDEX, not count */
static int fixed_nvars(const PolyVec* F) {
  int n = 0;
  for (int p = 0; p < F->n; p++)
    for (int t = 0; t < F->p[p].n; t++)
      for (int i = 0; i < NV; i++) if (F->p[p].t[t].m.e[i] != 0 && i > n) n = i;
  return n;
}
static Gen* gen_fixed(PolyVec F) {
  Gen* g = (Gen*)calloc(1, sizeof(Gen));
  g->kind = GEN_FIXED; g->F = F; g->n = fixed_nvars(&F);
  return g;
}
static Gen* gen_random(int kind, int n, int d, int s, double lam, int dist, int constants, int homogeneous, int pure) {
  Gen* g = (Gen*)calloc(1, sizeof(Gen));
  g->kind = kind; g->n = n; g->s = s; g->d = d; g->lam = lam; g->homogeneous = homogeneous; g->pure = pure;
  g->bases = (MonoVec*)malloc(sizeof(MonoVec) * (size_t)(d + 1));
  for (int i = 0; i < d + 1; i++) g->bases[i] = basis(n, i);
  int* count = (int*)malloc(sizeof(int) * (size_t)(d + 2));
  int c = degree_counts(n, d, dist, constants, count);
  discrete_init(&g->degree_dist, count, c);
  free(count);
  rng_seed(&g->rng, 1); /* reference seeds from std::random_device (ideals.cpp:163-164); callers always seed() */
  pv_init(&g->F);
  return g;
}
static void gen_free(Gen* g) {
  if (!g) return;
  if (g->bases) { for (int i = 0; i < g->d + 1; i++) free(g->bases[i].m); free(g->bases); }
  discrete_free(&g->degree_dist);
  pv_free(&g->F);
  free(g);
}
/* choice(begin,end,rng), ideals.h:68-73: a fresh uniform_int_distribution(0,size-1) per call */
static int choice_index(int size, Rng* r) { return rng_uniform_int(r, 0, size - 1); }

/* next(): RandomBinomialIdealGenerator ideals.cpp:168-201, RandomIdealGenerator :214-231, Fixed ideals.h:127.
 * returns 0 on success, -2 on the reference's 1000-trials throw */
static int gen_next(Gen* g, PolyVec* F) {
  pv_clear(F);
  if (g->kind == GEN_FIXED) {
    for (int i = 0; i < g->F.n; i++) pv_push_copy(F, &g->F.p[i]);
    return 0;
  }
  if (g->kind == GEN_BINOMIAL) {
    for (int i = 0; i < g->s; i++) {
      int c = g->pure ? cnorm(-1) : rng_uniform_int(&g->rng, 1, PRIME - 1);
      int d1, d2;
      if (g->homogeneous) d1 = d2 = discrete_sample(&g->degree_dist, &g->rng);
      else { d1 = discrete_sample(&g->degree_dist, &g->rng); d2 = discrete_sample(&g->degree_dist, &g->rng); }
      int ok = 0;
      for (int trials = 0; trials < 1000; trials++) {
        Mono m1 = g->bases[d1].m[choice_index(g->bases[d1].n, &g->rng)];
        Mono m2 = g->bases[d2].m[choice_index(g->bases[d2].n, &g->rng)];
        Term t[2];
        if (mono_lt(&m1, &m2)) { t[0].c = cnorm(1); t[0].m = m2; t[1].c = c; t[1].m = m1; ok = 1; }
        else if (mono_gt(&m1, &m2)) { t[0].c = cnorm(1); t[0].m = m1; t[1].c = c; t[1].m = m2; ok = 1; }
        if (ok) { Poly f = poly_make(t, 2); pv_reserve(F); F->p[F->n++] = f; break; }
      }
      if (!ok) return -2;
    }
    return 0;
  }
  for (int i = 0; i < g->s; i++) {
    Poly f; poly_init(&f);
    int terms = 2 + poisson_sample(g->lam, &g->rng);
    int d = discrete_sample(&g->degree_dist, &g->rng);
    for (int j = 0; j < terms; j++) {
      Term t; t.c = rng_uniform_int(&g->rng, 1, PRIME - 1);
      t.m = g->bases[d].m[choice_index(g->bases[d].n, &g->rng)];
      Poly p = poly_from_term(t), s = poly_add(&f, &p);
      poly_free(&p); poly_move(&f, &s);
      if (!g->homogeneous) d = discrete_sample(&g->degree_dist, &g->rng);
    }
    Term inv; inv.c = cdiv(cnorm(1), f.t[0].c); inv.m = mono_one();
    Poly q = term_mul_poly(&inv, &f);
    poly_free(&f);
    pv_reserve(F); F->p[F->n++] = q;
  }
  return 0;
}

/* parse_ideal_dist, ideals.cpp:103-143 */
static Gen* parse_ideal_dist(const char* s) {
  char buf[256]; char* args[16]; int na = 0;
  strncpy(buf, s, sizeof buf - 1); buf[sizeof buf - 1] = 0;
  for (char* p = buf; na < 16;) {
    args[na++] = p;
    char* q = strchr(p, '-');
    if (!q) break;
    *q = 0; p = q + 1;
  }
  if (na < 2) return NULL;
  if (strcmp(args[0], "cyclic") == 0) return gen_fixed(cyclic(atoi(args[1])));
  if (na < 4) return NULL;
  int constants = 0, homog = 0, pure = 0;
  for (int i = 0; i < na; i++) {
    if (!strcmp(args[i], "consts")) constants = 1;
    if (!strcmp(args[i], "homog")) homog = 1;
    if (!strcmp(args[i], "pure")) pure = 1;
  }
  const char* names[3] = {"uniform", "weighted", "maximum"};
  int n = atoi(args[0]), d = atoi(args[1]), sgen = atoi(args[2]);
  for (int t = 0; t < 3; t++)
    if (!strcmp(args[3], names[t])) return gen_random(GEN_BINOMIAL, n, d, sgen, 0.0, t, constants, homog, pure);
  if (na < 5) return NULL;
  double lam = atof(args[3]);
  for (int t = 0; t < 3; t++)
    if (!strcmp(args[4], names[t])) return gen_random(GEN_RANDOM, n, d, sgen, lam, t, constants, homog, 0);
  return NULL;
}

void* orc_gen_create(const char* dist) { return parse_ideal_dist(dist); }
void orc_gen_destroy(void* g) { gen_free((Gen*)g); }
void orc_gen_seed(void* g, int seed) { rng_seed(&((Gen*)g)->rng, seed); }
int orc_gen_nvars(void* g) { return ((Gen*)g)->n; }
int orc_gen_next(void* g, int* oterms, int cap_terms, int* olens, int cap_polys) {
  PolyVec F; pv_init(&F);
  int rc = gen_next((Gen*)g, &F);
  int r = rc < 0 ? rc : pv_to_wire(&F, oterms, cap_terms, olens, cap_polys);
  pv_free(&F);
  return r;
}

/* ------------------------------------------------------------------ selection: buchberger.cpp:160-241
 * std::min_element returns the FIRST minimal element; every key ends in (j,i) so ties cannot occur anyway. */
static int sugar_of_pair(const PolyVec* G, SPair p, const Mono* l) {
  Mono qi = mono_div(l, &G->p[p.i].t[0].m), qj = mono_div(l, &G->p[p.j].t[0].m);
  int a = G->p[p.i].sug + qi.deg, b = G->p[p.j].sug + qj.deg;
  return a > b ? a : b;
}
/* returns <0, 0, >0 comparing the selection keys of p1 and p2 (ascending variants) */
static int pair_key_cmp(const PolyVec* G, SPair p1, SPair p2, int base) {
  Mono m1 = mono_lcm(&G->p[p1.i].t[0].m, &G->p[p1.j].t[0].m);
  Mono m2 = mono_lcm(&G->p[p2.i].t[0].m, &G->p[p2.j].t[0].m);
  if (base == 1) { if (m1.deg != m2.deg) return m1.deg < m2.deg ? -1 : 1; }
  if (base == 3) {
    int s1 = sugar_of_pair(G, p1, &m1), s2 = sugar_of_pair(G, p2, &m2);
    if (s1 != s2) return s1 < s2 ? -1 : 1;
  }
  if (base == 2 || base == 3) { if (mono_lt(&m1, &m2)) return -1; if (mono_lt(&m2, &m1)) return 1; }
  if (p1.j != p2.j) return p1.j < p2.j ? -1 : 1;
  if (p1.i != p2.i) return p1.i < p2.i ? -1 : 1;
  return 0;
}
static int select_row(const PolyVec* G, const PairVec* P, int selection) {
  int base, rev;
  switch (selection) {
    case 0: base = 0; rev = 0; break;
    case 1: base = 1; rev = 0; break;
    case 2: base = 2; rev = 0; break;
    case 3: base = 3; rev = 0; break;
    case 5: base = 0; rev = 1; break;
    case 6: base = 1; rev = 1; break;
    case 7: base = 2; rev = 1; break;
    case 8: base = 3; rev = 1; break;
    default: return -1;
  }
  int best = 0;
  for (int k = 1; k < P->n; k++) {
    int c = pair_key_cmp(G, P->p[k], P->p[best], base);
    if (rev ? (c > 0) : (c < 0)) best = k;
  }
  return best;
}

/* ------------------------------------------------------------------ buchberger(): buchberger.cpp:125-266 */
typedef struct { int zero, nonzero, additions; double total_reward, discounted_return; } Stats;

static PolyVec buchberger_from(const PolyVec* F, const PairVec* S, int selection, int elimination, int rewards,
                               int sort_reducers, double gamma, int seed, Stats* st) {
  PolyVec G = pv_copy(F), G_ = pv_copy(F);
  PairVec P = pairs_copy(S);
  memset(st, 0, sizeof *st);
  double discount = 1.0;
  if (sort_reducers) pv_sort_by_lm(&G_);
  Rng rng; rng_seed(&rng, seed);
  while (P.n) {
    int row = (selection == 4) ? choice_index(P.n, &rng) : select_row(&G, &P, selection);
    SPair p = P.p[row];
    memmove(&P.p[row], &P.p[row + 1], sizeof(SPair) * (size_t)(P.n - row - 1)); P.n--;
    Poly s = spoly(&G.p[p.i], &G.p[p.j]);
    int steps;
    Poly r = reduce(&s, &G_, &steps);
    double reward = (rewards == 0) ? (-1.0 - steps) : -1.0;
    st->additions += steps + 1;
    st->total_reward += reward;
    st->discounted_return += discount * reward;
    discount *= gamma;
    if (r.n != 0) {
      update(&G, &P, &r, elimination);
      st->nonzero++;
      if (sort_reducers) pv_insert_sorted(&G_, &r); else pv_push_copy(&G_, &r);
    } else {
      st->zero++;
    }
    poly_free(&s); poly_free(&r);
  }
  PolyVec M = minimalize(&G), out = interreduce(&M);
  pv_free(&M); pv_free(&G); pv_free(&G_); pairs_free(&P);
  return out;
}
int orc_buchberger(const int* Fterms, const int* Flens, int nF, int selection, int elimination, int rewards,
                   int sort_input, int sort_reducers, double gamma, int seed, int* oterms, int cap_terms, int* olens,
                   int cap_polys, double* stats) {
  (void)sort_input; /* accepted and ignored by the reference too (buchberger.cpp:125-140) */
  PolyVec F = pv_from_wire(Fterms, Flens, nF), G; pv_init(&G);
  PairVec P; pairs_init(&P);
  for (int i = 0; i < F.n; i++) update(&G, &P, &F.p[i], elimination);
  Stats st;
  PolyVec out = buchberger_from(&G, &P, selection, elimination, rewards, sort_reducers, gamma, seed, &st);
  stats[0] = st.zero; stats[1] = st.nonzero; stats[2] = st.additions; stats[3] = st.total_reward;
  stats[4] = st.discounted_return;
  int r = pv_to_wire(&out, oterms, cap_terms, olens, cap_polys);
  pv_free(&F); pv_free(&G); pv_free(&out); pairs_free(&P);
  return r;
}

/* ------------------------------------------------------------------ BuchbergerEnv: buchberger.cpp:269-351 */
typedef struct {
  Gen* gen;
  int elimination, rewards, sort_input, sort_reducers;
  PolyVec G, G_;
  PairVec P;
} Env;

static Env* env_new(const char* dist, int elimination, int rewards, int sort_input, int sort_reducers) {
  Gen* g = parse_ideal_dist(dist);
  if (!g) return NULL;
  Env* e = (Env*)calloc(1, sizeof(Env));
  e->gen = g; e->elimination = elimination; e->rewards = rewards; e->sort_input = sort_input;
  e->sort_reducers = sort_reducers;
  pv_init(&e->G); pv_init(&e->G_); pairs_init(&e->P);
  return e;
}
static void env_free(Env* e) {
  if (!e) return;
  gen_free(e->gen); pv_free(&e->G); pv_free(&e->G_); pairs_free(&e->P); free(e);
}
/* reset, buchberger.cpp:299-315 (re-rolls while P is empty) */
static void env_reset(Env* e) {
  for (;;) {
    PolyVec F; pv_init(&F);
    if (gen_next(e->gen, &F) < 0) { pv_free(&F); return; }
    if (e->sort_input) pv_sort_by_lm(&F);
    pv_clear(&e->G); pv_clear(&e->G_); e->P.n = 0;
    for (int i = 0; i < F.n; i++) {
      update(&e->G, &e->P, &F.p[i], e->elimination);
      if (e->sort_reducers) pv_insert_sorted(&e->G_, &F.p[i]); else pv_push_copy(&e->G_, &F.p[i]);
    }
    pv_free(&F);
    if (e->P.n) return;
    if (e->gen->kind == GEN_FIXED) return; /* the reference would recurse forever here */
  }
}
/* step, buchberger.cpp:318-329 */
static double env_step(Env* e, SPair a) {
  int w = 0;
  for (int k = 0; k < e->P.n; k++) if (!(e->P.p[k].i == a.i && e->P.p[k].j == a.j)) e->P.p[w++] = e->P.p[k];
  e->P.n = w;
  Poly s = spoly(&e->G.p[a.i], &e->G.p[a.j]);
  int steps;
  Poly r = reduce(&s, &e->G_, &steps);
  if (r.n != 0) {
    update(&e->G, &e->P, &r, e->elimination);
    if (e->sort_reducers) pv_insert_sorted(&e->G_, &r); else pv_push_copy(&e->G_, &r);
  }
  poly_free(&s); poly_free(&r);
  return (e->rewards == 0) ? (-1.0 - steps) : -1.0;
}
static int strategy_code(const char* s) {
  if (!strcmp(s, "first")) return 0;
  if (!strcmp(s, "degree")) return 1;
  if (!strcmp(s, "normal")) return 2;
  if (!strcmp(s, "sugar")) return 3;
  return -1;
}
/* value, buchberger.cpp:332-351 (deterministic strategies only, see header) */
static double env_value(const Env* e, const char* strategy, double gamma) {
  int sel = strategy_code(strategy);
  if (sel < 0) return NAN;
  Stats st;
  PolyVec out = buchberger_from(&e->G, &e->P, sel, e->elimination, e->rewards, e->sort_reducers, gamma, 0, &st);
  pv_free(&out);
  return st.discounted_return;
}

/* same contract as ref_env_value_seeded (oracle/ref_shim.cpp): value() with explicit seeds for the random strategies */
static double env_value_one(const Env* e, int sel, double gamma, int seed) {
  Stats st;
  PolyVec out = buchberger_from(&e->G, &e->P, sel, e->elimination, e->rewards, e->sort_reducers, gamma, seed, &st);
  pv_free(&out);
  return st.discounted_return;
}
double orc_env_value_seeded(void* h, int selection, double gamma, int seed, int rollouts) {
  const Env* e = (const Env*)h;
  double best;
  if (selection == 100) {
    if (rollouts <= 0) rollouts = 101;
    best = env_value_one(e, 1, gamma, 0);
    for (int r = 1; r < rollouts; r++) { double v = env_value_one(e, 4, gamma, seed + r - 1); if (v > best) best = v; }
    return best;
  }
  if (selection != 4 || rollouts < 1) rollouts = 1;
  best = env_value_one(e, selection, gamma, seed);
  for (int r = 1; r < rollouts; r++) { double v = env_value_one(e, selection, gamma, seed + r); if (v > best) best = v; }
  return best;
}

void* orc_env_create(const char* dist, int elimination, int rewards, int sort_input, int sort_reducers) {
  return env_new(dist, elimination, rewards, sort_input, sort_reducers);
}
void orc_env_destroy(void* h) { env_free((Env*)h); }
void orc_env_seed(void* h, int seed) { rng_seed(&((Env*)h)->gen->rng, seed); }
void orc_env_set_ideal(void* h, const int* terms, const int* lens, int npoly) {
  Env* e = (Env*)h;
  gen_free(e->gen);
  e->gen = gen_fixed(pv_from_wire(terms, lens, npoly));
}
int orc_env_nvars(void* h) { return ((Env*)h)->gen->n; }
void orc_env_reset(void* h) { env_reset((Env*)h); }
double orc_env_step(void* h, int i, int j) { SPair a = {i, j}; return env_step((Env*)h, a); }
int orc_env_npairs(void* h) { return ((Env*)h)->P.n; }
int orc_env_nbasis(void* h) { return ((Env*)h)->G.n; }
int orc_env_nterms(void* h) {
  Env* e = (Env*)h; int n = 0;
  for (int i = 0; i < e->G.n; i++) n += e->G.p[i].n;
  return n;
}
int orc_env_pairs(void* h, int* pairs, int cap) {
  Env* e = (Env*)h;
  if (e->P.n > cap) return -1;
  for (int i = 0; i < e->P.n; i++) { pairs[2 * i] = e->P.p[i].i; pairs[2 * i + 1] = e->P.p[i].j; }
  return e->P.n;
}
int orc_env_basis(void* h, int* oterms, int cap_terms, int* olens, int cap_polys) {
  return pv_to_wire(&((Env*)h)->G, oterms, cap_terms, olens, cap_polys);
}
int orc_env_reducers(void* h, int* oterms, int cap_terms, int* olens, int cap_polys) {
  return pv_to_wire(&((Env*)h)->G_, oterms, cap_terms, olens, cap_polys);
}
double orc_env_value(void* h, const char* strategy, double gamma) { return env_value((Env*)h, strategy, gamma); }
int orc_env_select(void* h, int selection) { Env* e = (Env*)h; return select_row(&e->G, &e->P, selection); }
int orc_env_final_gb(void* h, int* oterms, int cap_terms, int* olens, int cap_polys) {
  Env* e = (Env*)h;
  PolyVec M = minimalize(&e->G), R = interreduce(&M);
  int r = pv_to_wire(&R, oterms, cap_terms, olens, cap_polys);
  pv_free(&M); pv_free(&R);
  return r;
}
/* same contract as ref_env_run (oracle/ref_shim.cpp) */
int orc_env_run(void* h, int selection, const int* actions, int nactions, int* trace, int cap_steps) {
  Env* e = (Env*)h;
  int t = 0;
  while (e->P.n) {
    int row;
    if (selection >= 0) row = select_row(&e->G, &e->P, selection);
    else { if (t >= nactions) break; row = actions[t]; }
    if (row < 0 || row >= e->P.n) return -2;
    SPair p = e->P.p[row];
    double reward = env_step(e, p);
    if (t >= cap_steps) return -1;
    trace[5 * t + 0] = p.i; trace[5 * t + 1] = p.j; trace[5 * t + 2] = (int)(-reward);
    trace[5 * t + 3] = e->P.n; trace[5 * t + 4] = e->G.n;
    t++;
  }
  return t;
}

/* ------------------------------------------------------------------ LeadMonomialsEnv: buchberger.cpp:354-408 */
typedef struct { Env* env; int n, k, cols; int* state; int nstate, capstate; } LmEnv;

static void lm_rebuild(LmEnv* e) {
  Env* b = e->env;
  int need = b->P.n * e->cols;
  if (need > e->capstate) { e->capstate = need * 2 + 64; e->state = (int*)realloc(e->state, sizeof(int) * (size_t)e->capstate); }
  int o = 0;
  for (int r = 0; r < b->P.n; r++) {
    int side[2] = {b->P.p[r].i, b->P.p[r].j};
    for (int sidx = 0; sidx < 2; sidx++) {
      const Poly* f = &b->G.p[side[sidx]];
      /* lead_monomials_vector(f, k, n), buchberger.cpp:354-370: first k exponent vectors, zero padded */
      for (int t = 0; t < e->k; t++)
        for (int j = 0; j < e->n; j++) e->state[o++] = (t < f->n) ? f->t[t].m.e[j] : 0;
    }
  }
  e->nstate = need;
}
void* orc_lm_create(const char* dist, int sort_input, int sort_reducers, int k) {
  Env* b = env_new(dist, 0, 0, sort_input, sort_reducers); /* hard-coded GM + Additions, buchberger.cpp:377 */
  if (!b) return NULL;
  LmEnv* e = (LmEnv*)calloc(1, sizeof(LmEnv));
  e->env = b; e->k = k; e->n = b->gen->n; e->cols = 2 * e->n * k;
  return e;
}
void orc_lm_destroy(void* h) { LmEnv* e = (LmEnv*)h; if (!e) return; env_free(e->env); free(e->state); free(e); }
void orc_lm_seed(void* h, int seed) { rng_seed(&((LmEnv*)h)->env->gen->rng, seed); }
void orc_lm_set_ideal(void* h, const int* terms, const int* lens, int npoly, int nvars) {
  LmEnv* e = (LmEnv*)h;
  orc_env_set_ideal(e->env, terms, lens, npoly);
  if (nvars > 0) { e->n = nvars; e->cols = 2 * nvars * e->k; }
}
void orc_lm_reset(void* h) { LmEnv* e = (LmEnv*)h; env_reset(e->env); lm_rebuild(e); }
double orc_lm_step(void* h, int action) {
  LmEnv* e = (LmEnv*)h;
  double r = env_step(e->env, e->env->P.p[action]);
  lm_rebuild(e);
  return r;
}
int orc_lm_cols(void* h) { return ((LmEnv*)h)->cols; }
int orc_lm_state(void* h, int* out, int cap) {
  LmEnv* e = (LmEnv*)h;
  if (e->nstate > cap) return -e->nstate;
  memcpy(out, e->state, sizeof(int) * (size_t)e->nstate);
  return e->nstate;
}
double orc_lm_value(void* h, const char* strategy, double gamma) { return env_value(((LmEnv*)h)->env, strategy, gamma); }

/* ------------------------------------------------------------------ CPU baseline driver (kind "port"), single thread */
void orc_bench_selection(const char* dist, int selection, int seed0, int count, int nthreads, int with_matrix, int k,
                         double* out) {
  (void)nthreads;
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  long long S = 0, A = 0;
  LmEnv* lm = with_matrix ? (LmEnv*)orc_lm_create(dist, 0, 1, k) : NULL;
  Env* env = with_matrix ? lm->env : env_new(dist, 0, 0, 0, 1);
  for (int e = seed0; e < seed0 + count; e++) {
    rng_seed(&env->gen->rng, e);
    env_reset(env);
    if (lm) lm_rebuild(lm);
    while (env->P.n) {
      int row = select_row(&env->G, &env->P, selection);
      double r = env_step(env, env->P.p[row]);
      if (lm) lm_rebuild(lm);
      S++; A += (long long)(-r);
    }
  }
  if (lm) orc_lm_destroy(lm); else env_free(env);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  out[0] = (double)S; out[1] = (double)A;
  out[2] = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
