/*
 * ref_harness.cpp -- the reference's OWN implementation behind oracle_api.h.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h).  This translation unit contains no algorithm of
 * its own: every arithmetic step is executed by the unmodified headers under /root/reference
 * (resolved through -I$(CRX_REF_DIR) by oracle/Makefile); the code below only marshals flat
 * arrays into CustVector objects and back.  Built into oracle/_ref/libcrx_ref.so, which is
 * git-ignored.  No reference source is copied into this repository.
 *
 * Three things are needed to run the reference headers at all:
 *  1. include order of main.cpp:10-20 (vector_reader.hpp brings utils.hpp/mod before
 *     cust_hashtable.hpp, which uses mod without including it);
 *  2. CustHashtable<T>::insertVector falls off the end of a non-void function
 *     (cust_hashtable.hpp:66-70; g++ 13 emits ud2) -- an explicit specialisation with the same
 *     two statements plus `return 0;` is declared before first use;
 *  3. every RNG seed is std::chrono::system_clock::now() (lsh_cube.hpp:49,112;
 *     initialization.hpp:42,75) -- the token `system_clock` is re-pointed at a settable fake
 *     clock for the reference headers only (all std headers are included first).
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <ctime>
#include <fstream>
#include <functional>
#include <iostream>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

static uint64_t g_seed = 1;

namespace std { namespace chrono {
struct crx_fake_clock {
    struct dur { unsigned long count() const { return (unsigned long)g_seed; } };
    struct tp { dur time_since_epoch() const { return dur(); } };
    static tp now() { return tp(); }
};
} }

#define system_clock crx_fake_clock
#define private public /* the harness reads CustHashtable::hashGenerator to dump detailed hashes */
#include "lib/in_out/vector_reader.hpp"
#include "lib/data_structures/cust_vector.hpp"
#include "lib/data_structures/cust_hashtable.hpp"

template <>
int CustHashtable<double>::insertVector(CustVector<double>* inVector) {
    unsigned int index = mod(hashGenerator->generate(inVector), buckets.size());
    buckets[index]->insertVector(inVector);
    return 0;
}

#include "lib/data_structures/tweet.h"
#include "lib/lsh_cube.hpp"
#include "lib/clustering_phases/initialization.hpp"
#include "lib/clustering_phases/assignment.hpp"
#include "lib/clustering_phases/update.hpp"
#include "lib/clustering_phases/silhouette.hpp"
#include "lib/crypto_rec.hpp"
#undef private
#undef system_clock

#include "oracle_api.h"

using std::string;
using std::vector;
typedef CustVector<double> CV;

static const char* metric_name(int m) { return m == 0 ? "euclidean" : "cosine"; }

static vector<CV> make_vectors(const double* X, int64_t N, int D, const char* prefix = "") {
    vector<CV> v;
    v.reserve(N);
    for (int64_t i = 0; i < N; i++)
        v.emplace_back(string(prefix) + std::to_string(i), vector<double>(X + i * D, X + (i + 1) * D));
    return v;
}

static vector<CV> make_users(const double* X, const uint8_t* unknown, const double* mean, int64_t N, int D,
                             const char* prefix = "") {
    vector<CV> v;
    v.reserve(N);
    for (int64_t i = 0; i < N; i++) {
        std::set<int> unk;
        for (int j = 0; j < D; j++)
            if (unknown[i * D + j]) unk.insert(j);
        v.emplace_back(string(prefix) + std::to_string(i), vector<double>(X + i * D, X + (i + 1) * D), unk, mean[i]);
    }
    return v;
}

/* centroid pointers: alias into vecs when cidx >= 0, else heap objects named like update.hpp:48 */
static vector<CV*> make_centroids(vector<CV>& vecs, const double* C, int K, int D, const int32_t* cidx) {
    vector<CV*> c(K);
    for (int i = 0; i < K; i++) {
        if (cidx && cidx[i] >= 0) c[i] = &vecs[cidx[i]];
        else c[i] = new CV("k_means_center", vector<double>(C + (size_t)i * D, C + (size_t)(i + 1) * D));
    }
    return c;
}
static void free_centroids(vector<CV*>& c) {
    for (auto p : c)
        if (p->getId() == "k_means_center") delete p;
}

extern "C" {

const char* orc_kind(void) { return "reference"; }

int orc_mod_ii(int x, int n) { return mod(x, n); }
int orc_mod_li(long x, int n) { return mod(x, n); }
int orc_mod_iz(int x, size_t n) { return mod(x, n); }
int orc_mod_ui(unsigned int x, int n) { return mod(x, n); }

int orc_hamming(int num, int dist, int min_bit, int bits, int* out, int cap) {
    vector<int> r = get_num_hamming_dist_from(num, dist, min_bit, bits);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = r[i];
    return (int)r.size();
}

void orc_quicksort(double* sims, int* ids, int n) {
    vector<double> s(sims, sims + n);
    vector<int> d(ids, ids + n);
    parallel_quickSort(s, d, 0, n - 1);
    std::copy(s.begin(), s.end(), sims);
    std::copy(d.begin(), d.end(), ids);
}

void orc_rng_kat(uint64_t seed, double* nd3, float* nf2, float* uf1, int* ui1, int* u12_6) {
    {
        std::default_random_engine e;
        e.seed((unsigned long)seed);
        std::normal_distribution<double> nd(0, 1);
        for (int i = 0; i < 3; i++) nd3[i] = nd(e);
    }
    std::default_random_engine e;
    e.seed((unsigned long)seed);
    std::normal_distribution<float> nf(0, 1);
    for (int i = 0; i < 2; i++) nf2[i] = nf(e);
    std::uniform_real_distribution<float> uf(0, 0.4f);
    uf1[0] = uf(e);
    std::uniform_int_distribution<int> ui(0, 100);
    ui1[0] = ui(e);
    for (int i = 0; i < 6; i++) {
        std::uniform_int_distribution<int> u12(1, 2);
        u12_6[i] = u12(e);
    }
}

double orc_inner_product(const double* a, const double* b, int d) {
    CV va("a", vector<double>(a, a + d)), vb("b", vector<double>(b, b + d));
    return (double)va.inner_product<double>(&vb, 0.0);
}
double orc_euclidean_distance(const double* a, const double* b, int d) {
    CV va("a", vector<double>(a, a + d)), vb("b", vector<double>(b, b + d));
    return va.euclideanDistance(&vb);
}
double orc_cosine_distance(const double* a, const double* b, int d) {
    CV va("a", vector<double>(a, a + d)), vb("b", vector<double>(b, b + d));
    return va.cosineDistance(&vb);
}
double orc_cosine_similarity(const double* a, const double* b, int d) {
    CV va("a", vector<double>(a, a + d)), vb("b", vector<double>(b, b + d));
    return va.cosineSimilarity(&vb);
}

int orc_lsh_hash(const double* X, int64_t N, int D, int metric, int k, int L, int div, double w, uint64_t seed,
                 int32_t* bucket_ids, int32_t* det_hashes) {
    g_seed = seed;
    vector<CV> vecs = make_vectors(X, N, D);
    vector<CustHashtable<double>*> tabs = create_LSH_hashtables<double>(vecs, metric_name(metric), k, L, div, w);
    for (int l = 0; l < L; l++) {
        for (int64_t i = 0; i < N; i++) bucket_ids[(size_t)l * N + i] = tabs[l]->getHash(&vecs[i]);
        if (det_hashes && tabs[l]->hashGenerator->hasDetailedHash()) {
            auto* m = tabs[l]->hashGenerator->getDetailedHashes();
            for (int64_t i = 0; i < N; i++) {
                const vector<int>& h = (*m)[vecs[i].getId()];
                for (int j = 0; j < k; j++) det_hashes[((size_t)l * N + i) * k + j] = h[j];
            }
        }
    }
    for (auto t : tabs) delete t;
    return 0;
}

int64_t orc_lsh_candidates(const double* X, int64_t N, int D, int metric, int k, int L, int div, double w,
                           uint64_t seed, int64_t q, int filtered, int32_t* out, int64_t cap) {
    g_seed = seed;
    vector<CV> vecs = make_vectors(X, N, D);
    vector<CustHashtable<double>*> tabs = create_LSH_hashtables<double>(vecs, metric_name(metric), k, L, div, w);
    vector<CV*> c = filtered ? get_LSH_filtered_combined_buckets(tabs, &vecs[q]) : get_LSH_combined_buckets(tabs, &vecs[q]);
    for (size_t i = 0; i < c.size() && (int64_t)i < cap; i++) out[i] = (int32_t)(c[i] - &vecs[0]);
    for (auto t : tabs) delete t;
    return (int64_t)c.size();
}

int orc_cube_hash(const double* X, int64_t N, int D, int metric, int k, double w, uint64_t seed, int32_t* ids) {
    g_seed = seed;
    vector<CV> vecs = make_vectors(X, N, D);
    /* NOTE: create_hypercube's engine is a stack local that EuclideanFGen keeps pointing at
     * (lsh_cube.hpp:113, euclidean_f_gen.hpp:55); getHash below only re-hashes rows seen at
     * build time, so no draw happens after the engine died. */
    CustHashtable<double>* cube = create_hypercube<double>(vecs, metric_name(metric), k, w);
    for (int64_t i = 0; i < N; i++) ids[i] = cube->getHash(&vecs[i]);
    delete cube;
    return 0;
}

int64_t orc_cube_candidates(const double* X, int64_t N, int D, int metric, int k, double w, uint64_t seed, int64_t q,
                            int probes, int32_t* out, int64_t cap) {
    g_seed = seed;
    vector<CV> vecs = make_vectors(X, N, D);
    CustHashtable<double>* cube = create_hypercube<double>(vecs, metric_name(metric), k, w);
    vector<CV*> c = get_hypercube_combined_buckets<double>(*cube, &vecs[q], probes, k);
    for (size_t i = 0; i < c.size() && (int64_t)i < cap; i++) out[i] = (int32_t)(c[i] - &vecs[0]);
    delete cube;
    return (int64_t)c.size();
}

int orc_rand_selection(const double* X, int64_t N, int D, int K, uint64_t seed, int32_t* idx) {
    g_seed = seed;
    vector<CV> vecs = make_vectors(X, N, D);
    vector<CV*> c = rand_selection(vecs, K);
    for (int i = 0; i < K; i++) idx[i] = (int32_t)(c[i] - &vecs[0]);
    return 0;
}

int orc_k_means_pp(const double* X, int64_t N, int D, int K, int metric, uint64_t seed, int32_t* idx) {
    g_seed = seed;
    vector<CV> vecs = make_vectors(X, N, D);
    vector<CV*> c = k_means_pp(vecs, K, metric_name(metric));
    for (int i = 0; i < K; i++) idx[i] = (int32_t)(c[i] - &vecs[0]);
    return 0;
}

int orc_lloyds_assignment(const double* X, int64_t N, int D, const double* C, int K, const int32_t* cidx, int metric,
                          int32_t* labels, double* dists) {
    vector<CV> vecs = make_vectors(X, N, D);
    vector<CV*> cent = make_centroids(vecs, C, K, D, cidx);
    lloyds_assignment(vecs, cent, metric_name(metric));
    for (int64_t i = 0; i < N; i++) {
        labels[i] = vecs[i].getCluster();
        dists[i] = vecs[i].getDistFromCentroid();
    }
    free_centroids(cent);
    return 0;
}

int orc_lsh_range_assignment(const double* X, int64_t N, int D, const int32_t* cidx, int K, int metric, int k, int L,
                             int div, double w, uint64_t seed, int32_t* labels, double* dists, int32_t* before) {
    g_seed = seed;
    vector<CV> vecs = make_vectors(X, N, D);
    vector<CustHashtable<double>*> tabs = create_LSH_hashtables<double>(vecs, metric_name(metric), k, L, div, w);
    vector<CV*> cent = make_centroids(vecs, nullptr, K, D, cidx);
    if (before) {
        /* the same three steps as lsh_range_assignment (assignment.hpp:109-129), stopped
         * before lloyds_for_remaining so the range-search labels can be inspected */
        remove_clustering(vecs);
        vector<vector<CV*>> comb(K);
        for (int c = 0; c < K; c++) comb[c] = get_LSH_combined_buckets<double>(tabs, cent[c]);
        range_assignment(comb, cent, metric_name(metric));
        for (int64_t i = 0; i < N; i++) before[i] = vecs[i].getCluster();
    }
    lsh_range_assignment(vecs, tabs, cent, metric_name(metric));
    for (int64_t i = 0; i < N; i++) {
        labels[i] = vecs[i].getCluster();
        dists[i] = vecs[i].getDistFromCentroid();
    }
    for (auto t : tabs) delete t;
    return 0;
}

int orc_lsh_range_assignment_vectors(const double* X, int64_t N, int D, const double* C, const int32_t* cidx, int K, int shared_ids,
                                     int metric, int k, int L, int div, double w, uint64_t seed, int32_t* labels, double* dists,
                                     int32_t* before) {
    g_seed = seed;
    vector<CV> vecs = make_vectors(X, N, D);
    vector<CustHashtable<double>*> tabs = create_LSH_hashtables<double>(vecs, metric_name(metric), k, L, div, w);
    vector<CV*> cent(K);
    for (int i = 0; i < K; i++) {
        if (cidx && cidx[i] >= 0) cent[i] = &vecs[cidx[i]];
        else cent[i] = new CV(shared_ids ? std::string("k_means_center") : "center_" + std::to_string(i),
                              vector<double>(C + (size_t)i * D, C + (size_t)(i + 1) * D));
    }
    if (before) {
        remove_clustering(vecs);
        vector<vector<CV*>> comb(K);
        for (int c = 0; c < K; c++) comb[c] = get_LSH_combined_buckets<double>(tabs, cent[c]);
        range_assignment(comb, cent, metric_name(metric));
        for (int64_t i = 0; i < N; i++) before[i] = vecs[i].getCluster();
    }
    lsh_range_assignment(vecs, tabs, cent, metric_name(metric));
    for (int64_t i = 0; i < N; i++) {
        labels[i] = vecs[i].getCluster();
        dists[i] = vecs[i].getDistFromCentroid();
    }
    for (int i = 0; i < K; i++)
        if (!(cidx && cidx[i] >= 0)) delete cent[i];
    for (auto t : tabs) delete t;
    return 0;
}

int orc_cube_range_assignment(const double* X, int64_t N, int D, const int32_t* cidx, int K, int metric, int k, double w,
                              int probes, uint64_t seed, int32_t* labels, double* dists, int32_t* before) {
    g_seed = seed;
    vector<CV> vecs = make_vectors(X, N, D);
    CustHashtable<double>* cube = create_hypercube<double>(vecs, metric_name(metric), k, w);
    vector<CV*> cent = make_centroids(vecs, nullptr, K, D, cidx);
    if (before) {
        remove_clustering(vecs);
        vector<vector<CV*>> comb(K);
        for (int c = 0; c < K; c++) comb[c] = get_hypercube_combined_buckets<double>(*cube, cent[c], probes, k);
        range_assignment(comb, cent, metric_name(metric));
        for (int64_t i = 0; i < N; i++) before[i] = vecs[i].getCluster();
    }
    cube_range_assignment(vecs, *cube, cent, metric_name(metric), probes, k);
    for (int64_t i = 0; i < N; i++) {
        labels[i] = vecs[i].getCluster();
        dists[i] = vecs[i].getDistFromCentroid();
    }
    delete cube;
    return 0;
}

int orc_k_means(const double* X, int64_t N, int D, const int32_t* labels, const double* C, int K, int metric,
                double min_dist, double* newC) {
    vector<CV> vecs = make_vectors(X, N, D);
    for (int64_t i = 0; i < N; i++) vecs[i].setCluster(labels[i], 0);
    vector<CV*> cent = make_centroids(vecs, C, K, D, nullptr);
    bool r = k_means(vecs, cent, metric_name(metric), min_dist);
    for (int c = 0; c < K; c++) {
        vector<double>* d = cent[c]->getDimensions();
        std::copy(d->begin(), d->end(), newC + (size_t)c * D);
    }
    free_centroids(cent);
    return r ? 1 : 0;
}

int orc_pam_lloyds(const double* X, int64_t N, int D, const int32_t* labels, const int32_t* cidx, int K, int metric,
                   int32_t* new_cidx) {
    vector<CV> vecs = make_vectors(X, N, D);
    for (int64_t i = 0; i < N; i++) vecs[i].setCluster(labels[i], 0);
    vector<CV*> cent = make_centroids(vecs, nullptr, K, D, cidx);
    bool r = pam_lloyds(vecs, cent, metric_name(metric));
    for (int c = 0; c < K; c++) new_cidx[c] = (int32_t)(cent[c] - &vecs[0]);
    return r ? 1 : 0;
}

int orc_silhouette(const double* X, int64_t N, int D, const int32_t* labels, const double* C, int K, int metric,
                   double* sils) {
    vector<CV> vecs = make_vectors(X, N, D);
    for (int64_t i = 0; i < N; i++) vecs[i].setCluster(labels[i], 0);
    vector<CV> cstore = make_vectors(C, K, D, "c");
    vector<CV*> cent(K);
    for (int c = 0; c < K; c++) cent[c] = &cstore[c];
    vector<vector<CV*>> clusters = separate_clusters_from_input(vecs, K);
    vector<double> s = silhouette_cluster(clusters, cent, metric_name(metric));
    std::copy(s.begin(), s.end(), sils);
    return 0;
}

int orc_recommend_lsh(const double* X, const uint8_t* unknown, const double* mean, int64_t N, int D,
                      const double* Xq, const uint8_t* unknown_q, const double* mean_q, int64_t Nq,
                      int metric, int k, int L, int div, double w, int P, int Nrec, uint64_t seed,
                      int32_t* recs, int32_t* nbr_idx, double* nbr_sim, int32_t* ncand) {
    g_seed = seed;
    vector<CV> base = make_users(X, unknown, mean, N, D);
    vector<CV> qstore;
    if (Xq) qstore = make_users(Xq, unknown_q, mean_q, Nq, D, "q");
    vector<CV>& queries = Xq ? qstore : base;
    if (!Xq) Nq = N;
    /* main.cpp:155-170 (rec A) / 201-216 (rec B) */
    vector<CustHashtable<double>*> tabs = create_LSH_hashtables<double>(base, metric_name(metric), k, L, div, w);
    for (int64_t u = 0; u < Nq; u++) {
        CV& user = queries[u];
        vector<CV*> neighbors = get_LSH_filtered_combined_buckets(tabs, &user);
        ncand[u] = (int32_t)neighbors.size();
        for (int j = 0; j < P; j++) { nbr_idx[u * P + j] = -1; nbr_sim[u * P + j] = 0; }
        for (int j = 0; j < Nrec; j++) recs[u * Nrec + j] = -1;
        if (!neighbors.empty()) {
            vector<double> sims = get_P_closest(neighbors, user, P);
            for (size_t j = 0; j < neighbors.size(); j++) {
                nbr_idx[u * P + j] = (int32_t)(neighbors[j] - &base[0]);
                nbr_sim[u * P + j] = sims[j];
            }
            vector<int> r = get_top_N_recom(neighbors, user, Nrec, sims);
            for (int j = 0; j < Nrec; j++) recs[u * Nrec + j] = r[j];
        }
    }
    for (auto t : tabs) delete t;
    return 0;
}

int orc_recommend_cluster(const double* X, const uint8_t* unknown, const double* mean, const int32_t* labels,
                          int64_t N, int D, int K,
                          const double* Xq, const uint8_t* unknown_q, const double* mean_q, const int32_t* qlabels,
                          int64_t Nq, int Nrec, int32_t* recs) {
    vector<CV> base = make_users(X, unknown, mean, N, D);
    for (int64_t i = 0; i < N; i++) base[i].setCluster(labels[i], 0);
    vector<CV> qstore;
    if (Xq) qstore = make_users(Xq, unknown_q, mean_q, Nq, D, "q");
    vector<CV>& queries = Xq ? qstore : base;
    if (!Xq) { Nq = N; qlabels = labels; }
    vector<vector<CV*>> clusters = separate_clusters_from_input(base, K);
    /* main.cpp:260-269 / 353-373 */
    for (int64_t u = 0; u < Nq; u++) {
        vector<CV*> neighbors = clusters[qlabels[u]];
        for (int j = 0; j < Nrec; j++) recs[u * Nrec + j] = -1;
        if (!neighbors.empty()) {
            vector<int> r = get_top_N_recom(neighbors, queries[u], Nrec);
            for (int j = 0; j < Nrec; j++) recs[u * Nrec + j] = r[j];
        }
    }
    return 0;
}

struct RefRecHandle {
    vector<CV> base;
    vector<CustHashtable<double>*> tabs;
};

void* orc_rec_handle_create(const double* X, const uint8_t* unknown, const double* mean, int64_t N, int D, int metric,
                            int k, int L, int div, double w, uint64_t seed) {
    g_seed = seed;
    RefRecHandle* h = new RefRecHandle();
    h->base = make_users(X, unknown, mean, N, D);
    h->tabs = create_LSH_hashtables<double>(h->base, metric_name(metric), k, L, div, w);
    return h;
}

int orc_rec_handle_query_nbr(void* hv, int64_t q_begin, int64_t q_end, int P, int Nrec, int32_t* recs, int32_t* ncand,
                             int32_t* nbr_idx, double* nbr_sim) {
    RefRecHandle* h = (RefRecHandle*)hv;
    for (int64_t u = q_begin; u < q_end; u++) {
        CV& user = h->base[u];
        vector<CV*> neighbors = get_LSH_filtered_combined_buckets(h->tabs, &user);
        ncand[u - q_begin] = (int32_t)neighbors.size();
        for (int j = 0; j < Nrec; j++) recs[(u - q_begin) * Nrec + j] = -1;
        if (nbr_idx) for (int j = 0; j < P; j++) { nbr_idx[(u - q_begin) * P + j] = -1; nbr_sim[(u - q_begin) * P + j] = 0; }
        if (!neighbors.empty()) {
            vector<double> sims = get_P_closest(neighbors, user, P);
            if (nbr_idx)
                for (size_t j = 0; j < neighbors.size(); j++) {
                    nbr_idx[(u - q_begin) * P + j] = (int32_t)(neighbors[j] - &h->base[0]);
                    nbr_sim[(u - q_begin) * P + j] = sims[j];
                }
            vector<int> r = get_top_N_recom(neighbors, user, Nrec, sims);
            for (int j = 0; j < Nrec; j++) recs[(u - q_begin) * Nrec + j] = r[j];
        }
    }
    return 0;
}

int orc_rec_handle_query(void* hv, int64_t q_begin, int64_t q_end, int P, int Nrec, int32_t* recs, int32_t* ncand) {
    return orc_rec_handle_query_nbr(hv, q_begin, q_end, P, Nrec, recs, ncand, nullptr, nullptr);
}

void orc_rec_handle_destroy(void* hv) {
    RefRecHandle* h = (RefRecHandle*)hv;
    for (auto t : h->tabs) delete t;
    delete h;
}

} /* extern "C" */
