/*
 * oracle_api.h -- flat-array C interface shared by the two CPU checkers.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: it may be
 * imported / linked / executed only by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs, and there only as the checker.
 *
 * Two shared objects export exactly these symbols:
 *   oracle/_ref/libcrx_ref.so     the reference's OWN headers (from /root/reference) compiled
 *                                 behind this interface by oracle/ref_harness.cpp
 *                                 (built only where /root/reference exists);
 *   oracle/_build/libcrx_oracle.so  oracle/crx_oracle.cpp, an independent restatement of the
 *                                 same algorithms on flat arrays (travels everywhere).
 * tests/test_oracle_pin.py pins the restatement against the reference build and against the
 * golden vectors under tests/golden/ (generated from the reference build by
 * tests/golden/make_golden.py).
 *
 * Conventions: row-major double X[N][D]; metric 0 = "euclidean", 1 = "cosine";
 * ids are the decimal row index (unique, as the reference's string-keyed caches assume);
 * `seed` is the value the reference would have read from
 * std::chrono::system_clock::now().time_since_epoch().count() (lsh_cube.hpp:49,112;
 * initialization.hpp:42,75).
 */
#ifndef CRX_ORACLE_API_H
#define CRX_ORACLE_API_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* orc_kind(void); /* "reference" or "port" */

/* ---- known-answer helpers (utils.hpp:97-98, utils.cpp:22-50, crypto_rec.hpp:235-277) ---- */
int orc_mod_ii(int x, int n);
int orc_mod_li(long x, int n);
int orc_mod_iz(int x, size_t n);
int orc_mod_ui(unsigned int x, int n);
int orc_hamming(int num, int dist, int min_bit, int bits, int* out, int cap);
void orc_quicksort(double* sims, int* ids, int n);
void orc_rng_kat(uint64_t seed, double* normal_d3, float* normal_f2, float* uni_f1, int* uni_i1, int* uni_12_6);

/* ---- vector math (cust_vector.hpp:107-174) ---- */
double orc_inner_product(const double* a, const double* b, int d);
double orc_euclidean_distance(const double* a, const double* b, int d);
double orc_cosine_distance(const double* a, const double* b, int d);
double orc_cosine_similarity(const double* a, const double* b, int d);

/* ---- LSH tables / hypercube (lsh_cube.hpp:45-177, generators/, cust_hashtable.hpp) ---- */
/* bucket_ids[L][N]; det_hashes[L][N][k] (euclidean only, may be NULL) */
int orc_lsh_hash(const double* X, int64_t N, int D, int metric, int k, int L, int lsh_bucket_div, double w,
                 uint64_t seed, int32_t* bucket_ids, int32_t* det_hashes);
/* candidates of query row q_index of X (get_LSH_[filtered_]combined_buckets); returns count (may exceed cap) */
int64_t orc_lsh_candidates(const double* X, int64_t N, int D, int metric, int k, int L, int lsh_bucket_div, double w,
                           uint64_t seed, int64_t q_index, int filtered, int32_t* out, int64_t cap);
int orc_cube_hash(const double* X, int64_t N, int D, int metric, int k, double w, uint64_t seed, int32_t* vertex_ids);
int64_t orc_cube_candidates(const double* X, int64_t N, int D, int metric, int k, double w, uint64_t seed,
                            int64_t q_index, int probes, int32_t* out, int64_t cap);

/* ---- clustering phases ---- */
int orc_rand_selection(const double* X, int64_t N, int D, int K, uint64_t seed, int32_t* idx);
int orc_k_means_pp(const double* X, int64_t N, int D, int K, int metric, uint64_t seed, int32_t* idx);
/* C[K][D]; cidx[K] = row of X the centroid pointer aliases, or -1 (assignment.hpp:77-78 overwrite) */
int orc_lloyds_assignment(const double* X, int64_t N, int D, const double* C, int K, const int32_t* cidx, int metric,
                          int32_t* labels, double* dists);
int orc_lsh_range_assignment(const double* X, int64_t N, int D, const int32_t* cidx, int K, int metric, int k, int L,
                             int lsh_bucket_div, double w, uint64_t seed, int32_t* labels, double* dists,
                             int32_t* labels_before_lloyd);
/* lsh_range_assignment with centroids that are any vectors: C[K][D]; cidx[c] = the stored row centroid c IS (the pointer
 * aliases it and keeps that vector's id), or -1 = a heap vector called "k_means_center" (update.hpp:48) when shared_ids,
 * "center_<c>" otherwise */
int orc_lsh_range_assignment_vectors(const double* X, int64_t N, int D, const double* C, const int32_t* cidx, int K, int shared_ids,
                                     int metric, int k, int L, int lsh_bucket_div, double w, uint64_t seed, int32_t* labels,
                                     double* dists, int32_t* labels_before_lloyd);
int orc_cube_range_assignment(const double* X, int64_t N, int D, const int32_t* cidx, int K, int metric, int k, double w,
                              int probes, uint64_t seed, int32_t* labels, double* dists, int32_t* labels_before_lloyd);
/* returns the bool of k_means (update.hpp:38); newC receives the computed means either way */
int orc_k_means(const double* X, int64_t N, int D, const int32_t* labels, const double* C, int K, int metric,
                double min_dist, double* newC);
int orc_pam_lloyds(const double* X, int64_t N, int D, const int32_t* labels, const int32_t* cidx, int K, int metric,
                   int32_t* new_cidx);
int orc_silhouette(const double* X, int64_t N, int D, const int32_t* labels, const double* C, int K, int metric,
                   double* sils /* K+1 */);

/* ---- recommendation (crypto_rec.hpp:214-345, main.cpp:159-170, 205-216, 260-269, 353-373) ---- */
/* tables over base rows; queries = base rows themselves when Xq == NULL (rec A) else external rows (rec B).
 * unknown[N][D] (1 = coin unknown to that user), mean[N] = known_mean.
 * recs[Nq][Nrec]; nbr_idx[Nq][P] (-1 padded); nbr_sim[Nq][P]; ncand[Nq] (0 => reference prints nothing). */
int orc_recommend_lsh(const double* X, const uint8_t* unknown, const double* mean, int64_t N, int D,
                      const double* Xq, const uint8_t* unknown_q, const double* mean_q, int64_t Nq,
                      int metric, int k, int L, int lsh_bucket_div, double w, int P, int Nrec, uint64_t seed,
                      int32_t* recs, int32_t* nbr_idx, double* nbr_sim, int32_t* ncand);
/* neighbours = all base rows with labels[v] == qlabels[q], no top-P cut (get_top_N_recom :328) */
int orc_recommend_cluster(const double* X, const uint8_t* unknown, const double* mean, const int32_t* labels,
                          int64_t N, int D, int K,
                          const double* Xq, const uint8_t* unknown_q, const double* mean_q, const int32_t* qlabels,
                          int64_t Nq, int Nrec, int32_t* recs);

/* ---- persistent handles for the timed CPU baseline (bench.py): tables built once, then slices of the
 * rec-A query loop (main.cpp:159-170) and of lloyds_assignment run on them, possibly from several threads
 * (read-only after the build for cosine tables). ---- */
void* orc_rec_handle_create(const double* X, const uint8_t* unknown, const double* mean, int64_t N, int D, int metric,
                            int k, int L, int lsh_bucket_div, double w, uint64_t seed);
int orc_rec_handle_query(void* h, int64_t q_begin, int64_t q_end, int P, int Nrec, int32_t* recs, int32_t* ncand);
/* the same, also returning the neighbours get_P_closest kept: nbr_idx[nq][P] (-1 padded), nbr_sim[nq][P] */
int orc_rec_handle_query_nbr(void* h, int64_t q_begin, int64_t q_end, int P, int Nrec, int32_t* recs, int32_t* ncand,
                             int32_t* nbr_idx, double* nbr_sim);
void orc_rec_handle_destroy(void* h);

#ifdef __cplusplus
}
#endif
#endif
