// TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
//
// ref_shim.cpp: a flat extern "C" window onto the UNMODIFIED reference sources
// (/root/reference/deepgroebner/{polynomials,ideals,buchberger}.cpp), compiled where they
// lie by oracle/Makefile into oracle/_ref/libdgref.so.  Nothing from the reference is copied:
// this file only #includes its headers and calls its public functions/classes.
//
// Used (a) to pin oracle/bb_oracle.c (the plain-C restatement) and the CUDA path bit-exactly,
// (b) to generate tests/golden/*.json (tests/golden/make_golden.py), and
// (c) as the `--impl reference` / cpu_baseline arm of bench.py (kind "reference").
//
// Wire format shared with bb_oracle.c ("flat polys"): a polynomial list is
//   lens[npoly]            number of terms of each polynomial
//   terms[sum(lens) * 9]   per term: coefficient in [0,P) followed by the 8 exponents
// Pairs are int pairs (i, j) flattened.

#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <optional>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

// The C++ env cannot be handed an explicit ideal (ideal_gen is private, buchberger.h:200).
// Open the access specifiers for the reference headers only (std headers are already in).
#define private public
#define protected public
#include "polynomials.h"
#include "ideals.h"
#include "buchberger.h"
#undef private
#undef protected

namespace {

constexpr int W = 9;  // ints per term on the wire

Polynomial poly_from(const int* t, int n) {
  if (n == 0) return Polynomial{};
  std::vector<Term> tv;
  for (int i = 0; i < n; i++) {
    std::array<int, N> e{};
    for (int k = 0; k < N; k++) e[k] = t[i * W + 1 + k];
    tv.push_back(Term{Coefficient(t[i * W]), Monomial(e)});
  }
  return Polynomial(tv);
}

std::vector<Polynomial> polys_from(const int* terms, const int* lens, int npoly) {
  std::vector<Polynomial> F;
  int off = 0;
  for (int p = 0; p < npoly; p++) {
    F.push_back(poly_from(terms + off * W, lens[p]));
    off += lens[p];
  }
  return F;
}

int coef_int(Coefficient c) { return c.c; }

// returns number of terms, or -(needed) if cap too small
int poly_to(const Polynomial& f, int* out, int cap_terms) {
  int n = f.terms.size();
  if (n > cap_terms) return -n;
  for (int i = 0; i < n; i++) {
    out[i * W] = coef_int(f.terms[i].coeff);
    for (int k = 0; k < N; k++) out[i * W + 1 + k] = f.terms[i].monom[k];
  }
  return n;
}

// returns npoly or negative on overflow
int polys_to(const std::vector<Polynomial>& F, int* terms, int cap_terms, int* lens, int cap_polys) {
  if ((int)F.size() > cap_polys) return -1;
  int off = 0;
  for (size_t p = 0; p < F.size(); p++) {
    int n = poly_to(F[p], terms + off * W, cap_terms - off);
    if (n < 0) return -1;
    lens[p] = n;
    off += n;
  }
  return F.size();
}

EliminationType elim_of(int e) {
  return e == 0 ? EliminationType::GebauerMoeller : (e == 1 ? EliminationType::LCM : EliminationType::None);
}
RewardType rew_of(int r) { return r == 0 ? RewardType::Additions : RewardType::Reductions; }
SelectionType sel_of(int s) { return static_cast<SelectionType>(s); }  // First,Degree,Normal,Sugar,Random,Last,Codegree,Strange,Spice

// Row index the reference's comparator (buchberger.cpp:160-241) would pick via std::min_element.
int select_row(const std::vector<Polynomial>& G, const std::vector<SPair>& P, int selection) {
  std::function<bool(const SPair&, const SPair&)> select;
  switch (selection) {
    case 0:
      select = [](const SPair& p1, const SPair& p2) { return std::tie(p1.j, p1.i) < std::tie(p2.j, p2.i); };
      break;
    case 1:
      select = [&G](const SPair& p1, const SPair& p2) {
        int d1 = lcm(G[p1.i].LM(), G[p1.j].LM()).deg();
        int d2 = lcm(G[p2.i].LM(), G[p2.j].LM()).deg();
        return std::tie(d1, p1.j, p1.i) < std::tie(d2, p2.j, p2.i);
      };
      break;
    case 2:
      select = [&G](const SPair& p1, const SPair& p2) {
        Monomial m1 = lcm(G[p1.i].LM(), G[p1.j].LM());
        Monomial m2 = lcm(G[p2.i].LM(), G[p2.j].LM());
        return std::tie(m1, p1.j, p1.i) < std::tie(m2, p2.j, p2.i);
      };
      break;
    case 3:
      select = [&G](const SPair& p1, const SPair& p2) {
        Monomial m1 = lcm(G[p1.i].LM(), G[p1.j].LM());
        Monomial m2 = lcm(G[p2.i].LM(), G[p2.j].LM());
        int s1 = std::max(G[p1.i].sugar() + (m1 / G[p1.i].LM()).deg(), G[p1.j].sugar() + (m1 / G[p1.j].LM()).deg());
        int s2 = std::max(G[p2.i].sugar() + (m2 / G[p2.i].LM()).deg(), G[p2.j].sugar() + (m2 / G[p2.j].LM()).deg());
        return std::tie(s1, m1, p1.j, p1.i) < std::tie(s2, m2, p2.j, p2.i);
      };
      break;
    default:
      return -1;
  }
  return std::min_element(P.begin(), P.end(), select) - P.begin();
}

struct RefEnv {
  BuchbergerEnv env;
  RefEnv(const char* dist, int elim, int rew, bool si, bool sr) : env(dist, elim_of(elim), rew_of(rew), si, sr) {}
};

}  // namespace

extern "C" {

int ref_prime() { return P; }
int ref_nslots() { return N; }

// ---- L0: field + monomials (polynomials.cpp) ----
int ref_coef_div(int a, int b) { return coef_int(Coefficient(a) / Coefficient(b)); }
int ref_coef_mul(int a, int b) { return coef_int(Coefficient(a) * Coefficient(b)); }
int ref_coef_add(int a, int b) { return coef_int(Coefficient(a) + Coefficient(b)); }
int ref_coef_sub(int a, int b) { return coef_int(Coefficient(a) - Coefficient(b)); }
int ref_coef_norm(int a) { return coef_int(Coefficient(a)); }

int ref_mono_cmp(const int* e1, const int* e2) {
  std::array<int, N> a{}, b{};
  for (int k = 0; k < N; k++) { a[k] = e1[k]; b[k] = e2[k]; }
  Monomial m1(a), m2(b);
  return (m1 > m2) ? 1 : ((m2 > m1) ? -1 : 0);
}
int ref_mono_divisible(const int* e1, const int* e2) {
  std::array<int, N> a{}, b{};
  for (int k = 0; k < N; k++) { a[k] = e1[k]; b[k] = e2[k]; }
  return is_divisible(Monomial(a), Monomial(b)) ? 1 : 0;
}
void ref_mono_lcm(const int* e1, const int* e2, int* out) {
  std::array<int, N> a{}, b{};
  for (int k = 0; k < N; k++) { a[k] = e1[k]; b[k] = e2[k]; }
  Monomial m = lcm(Monomial(a), Monomial(b));
  for (int k = 0; k < N; k++) out[k] = m[k];
}

// normalises a term list the way the Polynomial ctor does (sort only; polynomials.cpp:139-145)
int ref_poly_make(const int* t, int n, int* out, int cap) { return poly_to(poly_from(t, n), out, cap); }
int ref_poly_add(const int* f, int nf, const int* g, int ng, int* out, int cap) {
  return poly_to(poly_from(f, nf) + poly_from(g, ng), out, cap);
}
int ref_poly_sub(const int* f, int nf, const int* g, int ng, int* out, int cap) {
  return poly_to(poly_from(f, nf) - poly_from(g, ng), out, cap);
}
int ref_poly_mul(const int* f, int nf, const int* g, int ng, int* out, int cap) {
  return poly_to(poly_from(f, nf) * poly_from(g, ng), out, cap);
}
int ref_term_mul(const int* t, const int* f, int nf, int* out, int cap) {
  std::array<int, N> e{};
  for (int k = 0; k < N; k++) e[k] = t[1 + k];
  Term tt{Coefficient(t[0]), Monomial(e)};
  return poly_to(tt * poly_from(f, nf), out, cap);
}
int ref_parse_polynomial(const char* s, int* out, int cap) { return poly_to(parse_polynomial(std::string(s)), out, cap); }

// ---- L1: spoly / reduce / update / minimalize / interreduce (buchberger.cpp:18-122) ----
int ref_spoly(const int* f, int nf, const int* g, int ng, int* out, int cap) {
  return poly_to(spoly(poly_from(f, nf), poly_from(g, ng)), out, cap);
}

int ref_reduce(const int* g, int ng, const int* Fterms, const int* Flens, int nF, int* out, int cap, int* steps) {
  auto [r, st] = reduce(poly_from(g, ng), polys_from(Fterms, Flens, nF));
  *steps = st.steps;
  return poly_to(r, out, cap);
}

// pairs: in/out, capacity cap_pairs; returns new number of pairs (G is NOT returned: it is G + [f])
int ref_update(const int* Gterms, const int* Glens, int nG, int* pairs, int nP, int cap_pairs, const int* f, int nf,
               int elimination) {
  std::vector<Polynomial> G = polys_from(Gterms, Glens, nG);
  std::vector<SPair> P;
  for (int i = 0; i < nP; i++) P.push_back(SPair{pairs[2 * i], pairs[2 * i + 1]});
  update(G, P, poly_from(f, nf), elim_of(elimination));
  if ((int)P.size() > cap_pairs) return -1;
  for (size_t i = 0; i < P.size(); i++) { pairs[2 * i] = P[i].i; pairs[2 * i + 1] = P[i].j; }
  return P.size();
}

int ref_minimalize(const int* Gterms, const int* Glens, int nG, int* oterms, int cap_terms, int* olens, int cap_polys) {
  return polys_to(minimalize(polys_from(Gterms, Glens, nG)), oterms, cap_terms, olens, cap_polys);
}
int ref_interreduce(const int* Gterms, const int* Glens, int nG, int* oterms, int cap_terms, int* olens, int cap_polys) {
  return polys_to(interreduce(polys_from(Gterms, Glens, nG)), oterms, cap_terms, olens, cap_polys);
}

// full buchberger(F, ...) (buchberger.cpp:125-140); stats[5] = zero, nonzero, additions, total_reward, discounted_return
int ref_buchberger(const int* Fterms, const int* Flens, int nF, int selection, int elimination, int rewards,
                   int sort_input, int sort_reducers, double gamma, int seed, int* oterms, int cap_terms, int* olens,
                   int cap_polys, double* stats) {
  auto [G, st] = buchberger(polys_from(Fterms, Flens, nF), sel_of(selection), elim_of(elimination), rew_of(rewards),
                            sort_input != 0, sort_reducers != 0, gamma, std::optional<int>(seed));
  stats[0] = st.zero_reductions; stats[1] = st.nonzero_reductions; stats[2] = st.polynomial_additions;
  stats[3] = st.total_reward; stats[4] = st.discounted_return;
  return polys_to(G, oterms, cap_terms, olens, cap_polys);
}

// ---- L0': generators (ideals.cpp) ----
void* ref_gen_create(const char* dist) {
  try { return parse_ideal_dist(std::string(dist)).release(); } catch (...) { return nullptr; }
}
void ref_gen_destroy(void* g) { delete static_cast<IdealGenerator*>(g); }
void ref_gen_seed(void* g, int seed) { static_cast<IdealGenerator*>(g)->seed(seed); }
int ref_gen_nvars(void* g) { return static_cast<IdealGenerator*>(g)->nvars(); }
int ref_gen_next(void* g, int* oterms, int cap_terms, int* olens, int cap_polys) {
  try {
    return polys_to(static_cast<IdealGenerator*>(g)->next(), oterms, cap_terms, olens, cap_polys);
  } catch (...) { return -2; }
}
int ref_basis(int n, int d, int* out, int cap) {
  auto B = basis(n, d);
  if ((int)B.size() > cap) return -(int)B.size();
  for (size_t i = 0; i < B.size(); i++) for (int k = 0; k < N; k++) out[i * N + k] = B[i][k];
  return B.size();
}
int ref_degree_distribution(int n, int d, int dist, int constants, double* out, int cap) {
  auto dd = degree_distribution(n, d, static_cast<DistributionType>(dist), constants != 0);
  auto p = dd.probabilities();
  if ((int)p.size() > cap) return -1;
  for (size_t i = 0; i < p.size(); i++) out[i] = p[i];
  return p.size();
}
int ref_cyclic(int n, int* oterms, int cap_terms, int* olens, int cap_polys) {
  return polys_to(cyclic(n), oterms, cap_terms, olens, cap_polys);
}

// ---- L2: BuchbergerEnv (buchberger.cpp:269-351) ----
void* ref_env_create(const char* dist, int elimination, int rewards, int sort_input, int sort_reducers) {
  try { return new RefEnv(dist, elimination, rewards, sort_input != 0, sort_reducers != 0); } catch (...) { return nullptr; }
}
void ref_env_destroy(void* h) { delete static_cast<RefEnv*>(h); }
void ref_env_seed(void* h, int seed) { static_cast<RefEnv*>(h)->env.seed(seed); }
// analogue of passing FixedIdealGenerator(F) to the Python env (buchberger.py:389-394)
void ref_env_set_ideal(void* h, const int* terms, const int* lens, int npoly) {
  static_cast<RefEnv*>(h)->env.ideal_gen = std::make_unique<FixedIdealGenerator>(polys_from(terms, lens, npoly));
}
int ref_env_nvars(void* h) { return static_cast<RefEnv*>(h)->env.nvars(); }
void ref_env_reset(void* h) { static_cast<RefEnv*>(h)->env.reset(); }
double ref_env_step(void* h, int i, int j) { return static_cast<RefEnv*>(h)->env.step(SPair{i, j}); }
int ref_env_npairs(void* h) { return static_cast<RefEnv*>(h)->env.P.size(); }
int ref_env_nbasis(void* h) { return static_cast<RefEnv*>(h)->env.G.size(); }
int ref_env_nterms(void* h) {
  int n = 0;
  for (const auto& g : static_cast<RefEnv*>(h)->env.G) n += g.size();
  return n;
}
int ref_env_pairs(void* h, int* pairs, int cap) {
  auto& P = static_cast<RefEnv*>(h)->env.P;
  if ((int)P.size() > cap) return -1;
  for (size_t i = 0; i < P.size(); i++) { pairs[2 * i] = P[i].i; pairs[2 * i + 1] = P[i].j; }
  return P.size();
}
int ref_env_basis(void* h, int* oterms, int cap_terms, int* olens, int cap_polys) {
  return polys_to(static_cast<RefEnv*>(h)->env.G, oterms, cap_terms, olens, cap_polys);
}
int ref_env_reducers(void* h, int* oterms, int cap_terms, int* olens, int cap_polys) {
  return polys_to(static_cast<RefEnv*>(h)->env.G_, oterms, cap_terms, olens, cap_polys);
}
double ref_env_value(void* h, const char* strategy, double gamma) {
  return static_cast<RefEnv*>(h)->env.value(std::string(strategy), gamma);
}
// value() with the random strategies made reproducible: BuchbergerEnv::value (buchberger.cpp:332-351) calls the
// pair-set overload of buchberger() on (G, P); "random"/"sample" seed it from std::random_device there.  Here the
// same overload is called with an explicit seed per rollout: selection 0..8 -> best of `rollouts` runs seeded
// seed + r (rollouts is 1 unless Random); selection 100 -> "sample": one Degree run, then rollouts-1 (default 100)
// Random runs seeded seed + 0.., best kept (:333-341).
double ref_env_value_seeded(void* h, int selection, double gamma, int seed, int rollouts) {
  auto& e = static_cast<RefEnv*>(h)->env;
  auto run = [&](SelectionType sel, int sd) {
    auto [G_, st] = buchberger(e.G, e.P, sel, e.elimination, e.rewards, e.sort_reducers, gamma, std::optional<int>(sd));
    return st.discounted_return;
  };
  if (selection == 100) {
    if (rollouts <= 0) rollouts = 101;
    double best = run(SelectionType::Degree, 0);
    for (int r = 1; r < rollouts; r++) best = std::max(best, run(SelectionType::Random, seed + r - 1));
    return best;
  }
  if (selection != 4 || rollouts < 1) rollouts = 1;
  double best = run(sel_of(selection), seed);
  for (int r = 1; r < rollouts; r++) best = std::max(best, run(sel_of(selection), seed + r));
  return best;
}
int ref_env_select(void* h, int selection) {
  auto& e = static_cast<RefEnv*>(h)->env;
  return select_row(e.G, e.P, selection);
}
int ref_env_final_gb(void* h, int* oterms, int cap_terms, int* olens, int cap_polys) {
  return polys_to(interreduce(minimalize(static_cast<RefEnv*>(h)->env.G)), oterms, cap_terms, olens, cap_polys);
}

// Run one episode from the env's CURRENT state (after reset) to completion.
//   selection >= 0 : on-the-fly First/Degree/Normal/Sugar row choice (comparators of buchberger.cpp:160-241)
//   selection <  0 : replay `actions` (row indices into P, as LeadMonomialsEnv::step(int) takes them)
// trace rows: (i, j, additions = -reward under Additions, |P| after, |G| after); returns number of steps or -1.
int ref_env_run(void* h, int selection, const int* actions, int nactions, int* trace, int cap_steps) {
  auto& e = static_cast<RefEnv*>(h)->env;
  int t = 0;
  while (!e.P.empty()) {
    int row;
    if (selection >= 0) row = select_row(e.G, e.P, selection);
    else { if (t >= nactions) break; row = actions[t]; }
    if (row < 0 || row >= (int)e.P.size()) return -2;
    SPair p = e.P[row];
    double reward = e.step(p);
    if (t >= cap_steps) return -1;
    trace[5 * t + 0] = p.i; trace[5 * t + 1] = p.j; trace[5 * t + 2] = (int)(-reward);
    trace[5 * t + 3] = e.P.size(); trace[5 * t + 4] = e.G.size();
    t++;
  }
  return t;
}

// ---- L2: LeadMonomialsEnv (buchberger.cpp:373-408) exactly as wrapped.pyx drives it ----
void* ref_lm_create(const char* dist, int sort_input, int sort_reducers, int k) {
  try { return new LeadMonomialsEnv(dist, sort_input != 0, sort_reducers != 0, k); } catch (...) { return nullptr; }
}
void ref_lm_destroy(void* h) { delete static_cast<LeadMonomialsEnv*>(h); }
void ref_lm_seed(void* h, int seed) { static_cast<LeadMonomialsEnv*>(h)->seed(seed); }
void ref_lm_set_ideal(void* h, const int* terms, const int* lens, int npoly, int nvars) {
  auto* e = static_cast<LeadMonomialsEnv*>(h);
  e->env.ideal_gen = std::make_unique<FixedIdealGenerator>(polys_from(terms, lens, npoly));
  if (nvars > 0) { e->n = nvars; e->cols = 2 * nvars * e->k; }  // quirk Q1: stock nvars() is off by one
}
void ref_lm_reset(void* h) { static_cast<LeadMonomialsEnv*>(h)->reset(); }
double ref_lm_step(void* h, int action) { return static_cast<LeadMonomialsEnv*>(h)->step(action); }
int ref_lm_cols(void* h) { return static_cast<LeadMonomialsEnv*>(h)->cols; }
int ref_lm_state(void* h, int* out, int cap) {
  auto& s = static_cast<LeadMonomialsEnv*>(h)->state;
  if ((int)s.size() > cap) return -(int)s.size();
  std::copy(s.begin(), s.end(), out);
  return s.size();
}
double ref_lm_value(void* h, const char* strategy, double gamma) {
  return static_cast<LeadMonomialsEnv*>(h)->value(std::string(strategy), gamma);
}

// ---- CPU baseline drivers (timed inside, multi-threaded: one independent env per thread) ----
//
// Workload "episodes to completion under a built-in selection": per episode e in [seed0, seed0+count):
//   env.seed(e); env.reset(); then select/step until P is empty (the env API the GPU path replaces).
// out[0]=env steps, out[1]=polynomial additions, out[2]=seconds (wall, max over threads is the caller's wall).
// with_matrix != 0 steps LeadMonomialsEnv(k) instead (state matrix rebuilt per step, as wrapped.pyx sees it)
// and picks the action from the matrix-free comparator on the inner env.
void ref_bench_selection(const char* dist, int selection, int seed0, int count, int nthreads, int with_matrix, int k,
                         double* out) {
  std::vector<long long> steps(nthreads, 0), adds(nthreads, 0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    th.emplace_back([&, t]() {
      if (with_matrix) {
        LeadMonomialsEnv env{dist, false, true, k};
        for (int e = seed0 + t; e < seed0 + count; e += nthreads) {
          env.seed(e);
          env.reset();
          while (!env.env.P.empty()) {
            int row = select_row(env.env.G, env.env.P, selection);
            double r = env.step(row);
            steps[t]++; adds[t] += (long long)(-r);
          }
        }
      } else {
        BuchbergerEnv env{dist};
        for (int e = seed0 + t; e < seed0 + count; e += nthreads) {
          env.seed(e);
          env.reset();
          while (!env.P.empty()) {
            int row = select_row(env.G, env.P, selection);
            double r = env.step(env.P[row]);
            steps[t]++; adds[t] += (long long)(-r);
          }
        }
      }
    });
  }
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  long long S = 0, A = 0;
  for (int t = 0; t < nthreads; t++) { S += steps[t]; A += adds[t]; }
  out[0] = (double)S; out[1] = (double)A; out[2] = std::chrono::duration<double>(t1 - t0).count();
}

// Mirror of scripts/random_episodes.cpp:13-29 (uniform-random actions on LeadMonomialsEnv(dist,false,true,2)),
// one env per thread, env seeded seed0+t, harness RNG minstd_rand0 default-seeded like the script.
void ref_bench_random(const char* dist, int seed0, int episodes, int nthreads, double* out) {
  std::vector<long long> steps(nthreads, 0);
  std::vector<double> ret(nthreads, 0.0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    th.emplace_back([&, t]() {
      LeadMonomialsEnv env{dist, false, true, 2};
      env.seed(seed0 + t);
      std::default_random_engine rng;
      for (int e = t; e < episodes; e += nthreads) {
        env.reset();
        while (env.state.size() != 0) {
          std::uniform_int_distribution<int> ud{0, static_cast<int>(env.state.size() / env.cols) - 1};
          int action = ud(rng);
          ret[t] += env.step(action);
          steps[t]++;
        }
      }
    });
  }
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  long long S = 0; double R = 0;
  for (int t = 0; t < nthreads; t++) { S += steps[t]; R += ret[t]; }
  out[0] = (double)S; out[1] = -R; out[2] = std::chrono::duration<double>(t1 - t0).count();
}

// Whole episodes through the reference's own buchberger() loop (buchberger.cpp:125-266), any SelectionType incl.
// seeded Random: episode e draws its ideal from the generator stream seed0 + e (first next() with a non-empty pair
// set, like BuchbergerEnv::reset) and runs buchberger(F, selection, GebauerMoeller, Additions, false, true, gamma,
// sel_seed0 + e).  One generator per thread.  out[0]=env steps (zero + nonzero reductions), out[1]=additions,
// out[2]=seconds.
void ref_bench_buchberger(const char* dist, int selection, int seed0, int count, int nthreads, int sel_seed0,
                          double gamma, double* out) {
  std::vector<long long> steps(nthreads, 0), adds(nthreads, 0);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    th.emplace_back([&, t]() {
      BuchbergerEnv env{dist};
      for (int e = t; e < count; e += nthreads) {
        env.seed(seed0 + e);
        env.reset();
        std::vector<Polynomial> F = env.G;  // after reset() G holds exactly the generators (sort_input = false)
        auto res = buchberger(F, sel_of(selection), EliminationType::GebauerMoeller, RewardType::Additions, false, true,
                              gamma, std::optional<int>(sel_seed0 + e));
        steps[t] += res.second.zero_reductions + res.second.nonzero_reductions;
        adds[t] += res.second.polynomial_additions;
      }
    });
  }
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  long long S = 0, A = 0;
  for (int t = 0; t < nthreads; t++) { S += steps[t]; A += adds[t]; }
  out[0] = (double)S; out[1] = (double)A; out[2] = std::chrono::duration<double>(t1 - t0).count();
}

// ---- whole-run parity records (bench.py's exhaustive gate and the full-size GPU tests) ----
//
// ref_run_records: for every episode e in [0, count) the record the CUDA path's bb_run writes (bb_episode_stats,
// include/bbenv.h), computed from the UNMODIFIED reference env driven exactly as BuchbergerEnv::reset/step are
// (buchberger.cpp:299-329) with the pair chosen by the comparators of buchberger.cpp:160-241 or, for Random, by
// choice() (ideals.h:68-73) on a std::default_random_engine seeded sel_seed0 + e * sel_stride (buchberger.cpp:190-197).
// Episode e draws its ideal from stream seed seeds[e] (or seed0 + e).  Checksums as include/bbenv.h "checksums".
// Threads pull episodes from an atomic counter.  Layout of one record = bb_episode_stats (72 bytes).
struct RefRecord {
  int32_t steps, additions, zero_reductions, nonzero_reductions, nbasis, nterms, status, rerolls;
  uint64_t trace_hash, basis_hash, gb_hash;
  int32_t gb_polys, gb_terms;
  double discounted_return;
};
static_assert(sizeof(RefRecord) == 72, "must match bb_episode_stats");

static inline uint64_t rr_mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL;
  z ^= z >> 27; z *= 0x94d049bb133111ebULL;
  z ^= z >> 31;
  return z;
}
static inline uint64_t rr_item(uint64_t x, uint64_t pos) { return rr_mix64(x + 0x9E3779B97F4A7C15ULL * (pos + 1)); }
static uint64_t rr_polys_hash(const std::vector<Polynomial>& F, int* nterms_out) {
  uint64_t h = 0, t = 0;
  for (size_t q = 0; q < F.size(); q++) {
    for (const Term& tm : F[q].terms) {
      uint64_t elo = 0, ehi = 0;
      for (int v = 0; v < N; v++) {
        const uint64_t x = (uint64_t)tm.monom[v];
        if (v < 4) elo |= x << (16 * v); else ehi |= x << (16 * (v - 4));
      }
      h += rr_item((uint64_t)coef_int(tm.coeff), 3 * t) + rr_item(elo, 3 * t + 1) + rr_item(ehi, 3 * t + 2);
      t++;
    }
    h += rr_mix64((uint64_t)F[q].terms.size() + 0xD1B54A32D192ED03ULL * (uint64_t)(q + 1));
  }
  if (nterms_out) *nterms_out = (int)t;
  return h;
}

int ref_run_records(const char* dist, int selection, int elimination, int rewards, int sort_input, int sort_reducers,
                    int seed0, const int* seeds, int count, int sel_seed0, int sel_stride, int max_steps, double gamma,
                    int compute_gb, int nthreads, void* out_records) {
  RefRecord* out = static_cast<RefRecord*>(out_records);
  if (rewards != 0) return -2;   // the record's addition count is read off the Additions reward
  if (selection < 0 || selection > 4) return -1;   // the reversed strategies are covered through ref_buchberger
  if (nthreads < 1) nthreads = 1;
  std::atomic<int> next{0};
  std::atomic<int> failed{0};
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    th.emplace_back([&]() {
      try {
        BuchbergerEnv env{dist, elim_of(elimination), rew_of(rewards), sort_input != 0, sort_reducers != 0};
        auto twin = parse_ideal_dist(std::string(dist));   // counts the re-rolls of reset() (buchberger.cpp:313-314)
        for (;;) {
          const int e = next.fetch_add(1);
          if (e >= count) break;
          const int sd = seeds ? seeds[e] : seed0 + e;
          env.seed(sd);
          env.reset();
          RefRecord r;
          std::memset(&r, 0, sizeof r);
          {
            twin->seed(sd);
            int rolls = 0;
            for (;; rolls++) {
              std::vector<Polynomial> F = twin->next();
              if (sort_input)
                std::sort(F.begin(), F.end(), [](const Polynomial& f, const Polynomial& g) { return f.LM() < g.LM(); });
              if (F == env.G || rolls > 1000) break;
            }
            r.rerolls = rolls;
          }
          std::default_random_engine rng(sel_seed0 + e * sel_stride);
          uint64_t th_ = 0;
          double ret = 0.0, disc = 1.0;
          const int g0 = (int)env.G.size();
          while (!env.P.empty() && (max_steps == 0 || r.steps < max_steps)) {
            int row;
            if (selection == 4) row = (int)(choice(env.P.begin(), env.P.end(), rng) - env.P.begin());
            else row = select_row(env.G, env.P, selection);
            const SPair p = env.P[row];
            const double reward = env.step(p);
            const int a = (int)(-reward);   // Additions rewards (checked on entry): reward = -(1 + reduction steps)
            th_ = th_ * 0x9E3779B97F4A7C15ULL + ((((uint64_t)(uint32_t)a) << 32) | ((uint64_t)p.j << 16) | (uint64_t)p.i) + 1ULL;
            ret += disc * reward;
            disc *= gamma;
            r.steps++;
            r.additions += a;
          }
          r.nbasis = (int)env.G.size();
          r.nonzero_reductions = r.nbasis - g0;
          r.zero_reductions = r.steps - r.nonzero_reductions;
          r.status = env.P.empty() ? 2 : 1;
          r.trace_hash = th_;
          r.basis_hash = rr_polys_hash(env.G, &r.nterms);
          if (compute_gb && env.P.empty()) {
            const std::vector<Polynomial> gb = interreduce(minimalize(env.G));
            r.gb_hash = rr_polys_hash(gb, &r.gb_terms);
            r.gb_polys = (int)gb.size();
          }
          r.discounted_return = ret;
          out[e] = r;
        }
      } catch (...) { failed = 3; }
    });
  }
  for (auto& x : th) x.join();
  return -failed.load();
}

}  // extern "C"
