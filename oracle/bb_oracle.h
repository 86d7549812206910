/* TEST INFRASTRUCTURE ONLY -- see bb_oracle.c.  Flat C API of the CPU oracle (prefix orc_).
 * oracle/ref_shim.cpp exports the same functions with prefix ref_ over the unmodified reference. */
#ifndef BB_ORACLE_H
#define BB_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NV 8 /* exponent slots per monomial, polynomials.h:29 */
#define ORC_W 9  /* ints per term on the wire: coefficient + 8 exponents */

/* elimination: 0 GebauerMoeller, 1 LCM, 2 None.  rewards: 0 Additions, 1 Reductions.
 * selection: 0 First 1 Degree 2 Normal 3 Sugar 4 Random 5 Last 6 Codegree 7 Strange 8 Spice (buchberger.h:111) */

int orc_prime(void);
void orc_set_prime(int p); /* reference is fixed at 32003 (polynomials.h:10); python tests also use FF(101) */
int orc_nslots(void);

int orc_coef_div(int a, int b);
int orc_coef_mul(int a, int b);
int orc_coef_add(int a, int b);
int orc_coef_sub(int a, int b);
int orc_coef_norm(int a);
int orc_mono_cmp(const int* e1, const int* e2);
int orc_mono_divisible(const int* e1, const int* e2);
void orc_mono_lcm(const int* e1, const int* e2, int* out);

int orc_poly_make(const int* t, int n, int* out, int cap);
int orc_poly_add(const int* f, int nf, const int* g, int ng, int* out, int cap);
int orc_poly_sub(const int* f, int nf, const int* g, int ng, int* out, int cap);
int orc_poly_mul(const int* f, int nf, const int* g, int ng, int* out, int cap);
int orc_term_mul(const int* t, const int* f, int nf, int* out, int cap);
int orc_parse_polynomial(const char* s, int* out, int cap);

int orc_spoly(const int* f, int nf, const int* g, int ng, int* out, int cap);
int orc_reduce(const int* g, int ng, const int* Fterms, const int* Flens, int nF, int* out, int cap, int* steps);
int orc_update(const int* Gterms, const int* Glens, int nG, int* pairs, int nP, int cap_pairs, const int* f, int nf,
               int elimination);
int orc_minimalize(const int* Gterms, const int* Glens, int nG, int* oterms, int cap_terms, int* olens, int cap_polys);
int orc_interreduce(const int* Gterms, const int* Glens, int nG, int* oterms, int cap_terms, int* olens, int cap_polys);
int orc_buchberger(const int* Fterms, const int* Flens, int nF, int selection, int elimination, int rewards,
                   int sort_input, int sort_reducers, double gamma, int seed, int* oterms, int cap_terms, int* olens,
                   int cap_polys, double* stats);

void* orc_gen_create(const char* dist);
void orc_gen_destroy(void* g);
void orc_gen_seed(void* g, int seed);
int orc_gen_nvars(void* g);
int orc_gen_next(void* g, int* oterms, int cap_terms, int* olens, int cap_polys);
int orc_basis(int n, int d, int* out, int cap);
int orc_degree_distribution(int n, int d, int dist, int constants, double* out, int cap);
int orc_cyclic(int n, int* oterms, int cap_terms, int* olens, int cap_polys);

void* orc_env_create(const char* dist, int elimination, int rewards, int sort_input, int sort_reducers);
void orc_env_destroy(void* h);
void orc_env_seed(void* h, int seed);
void orc_env_set_ideal(void* h, const int* terms, const int* lens, int npoly);
int orc_env_nvars(void* h);
void orc_env_reset(void* h);
double orc_env_step(void* h, int i, int j);
int orc_env_npairs(void* h);
int orc_env_nbasis(void* h);
int orc_env_nterms(void* h);
int orc_env_pairs(void* h, int* pairs, int cap);
int orc_env_basis(void* h, int* oterms, int cap_terms, int* olens, int cap_polys);
int orc_env_reducers(void* h, int* oterms, int cap_terms, int* olens, int cap_polys);
double orc_env_value(void* h, const char* strategy, double gamma);
double orc_env_value_seeded(void* h, int selection, double gamma, int seed, int rollouts);
int orc_env_select(void* h, int selection);
int orc_env_final_gb(void* h, int* oterms, int cap_terms, int* olens, int cap_polys);
int orc_env_run(void* h, int selection, const int* actions, int nactions, int* trace, int cap_steps);

void* orc_lm_create(const char* dist, int sort_input, int sort_reducers, int k);
void orc_lm_destroy(void* h);
void orc_lm_seed(void* h, int seed);
void orc_lm_set_ideal(void* h, const int* terms, const int* lens, int npoly, int nvars);
void orc_lm_reset(void* h);
double orc_lm_step(void* h, int action);
int orc_lm_cols(void* h);
int orc_lm_state(void* h, int* out, int cap);
double orc_lm_value(void* h, const char* strategy, double gamma);

void orc_bench_selection(const char* dist, int selection, int seed0, int count, int nthreads, int with_matrix, int k,
                         double* out);

#ifdef __cplusplus
}
#endif
#endif
