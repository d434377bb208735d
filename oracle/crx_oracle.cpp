/*
 * crx_oracle.cpp -- CPU restatement ("port") of crypto-recommendation's hot path on flat arrays.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h): imported by tests/, smoke() and bench.py's
 * cpu_baseline / --impl reference legs as the checker, never by the product path.
 *
 * This is NOT the reference's code: it restates each algorithm on row-major arrays without the
 * CustVector / CustHashtable / string-keyed-cache machinery, citing the reference lines each
 * function follows.  It is pinned against the reference itself (oracle/_ref/libcrx_ref.so, built
 * from /root/reference by oracle/Makefile) by tests/test_oracle_pin.py and against the golden
 * vectors in tests/golden/ (made by tests/golden/make_golden.py from the reference build).
 *
 * Arithmetic contract (what parity hinges on; x86-64, GCC 13 libstdc++, -ffp-contract=off):
 *   - dot products: each product rounded to double, accumulated in x87 long double in index
 *     order (cust_vector.hpp:107-121);
 *   - squared norms / Euclidean sums: pow(x,2) == x*x rounded to double, accumulated in double
 *     in index order (cust_vector.hpp:126-174);
 *   - RNG: std::default_random_engine (minstd_rand0) + libstdc++ distributions, consumed in the
 *     order of SURVEY.md App. B.
 * Ids are assumed unique (decimal row index), which makes every string-keyed distance cache of
 * the reference value-transparent; the caches are therefore not restated.
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <random>
#include <unordered_map>
#include <vector>

#include "oracle_api.h"

namespace {

typedef std::default_random_engine Engine;

/* utils.hpp:97-98 -- (x % n + n) % n evaluated in the promoted type, result narrowed to int */
template <typename X, typename Nn>
inline int ref_mod(X x, Nn n) { return (int)((x % n + n) % n); }

/* cust_vector.hpp:107-121 */
inline long double dot_ld(const double* a, const double* b, int d) {
    long double acc = 0.0L;
    for (int i = 0; i < d; i++) {
        double p = a[i] * b[i];
        acc = acc + (long double)p;
    }
    return acc;
}
/* same with a float parameter vector: float*double promotes to double (euclidean_h_gen.hpp:80) */
inline long double dot_ld_f(const float* v, const double* x, int d) {
    long double acc = 0.0L;
    for (int i = 0; i < d; i++) {
        double p = (double)v[i] * x[i];
        acc = acc + (long double)p;
    }
    return acc;
}
/* Σ pow(x_i,2) in double, index order (cust_vector.hpp:148-151) */
inline double sqnorm(const double* a, int d) {
    double acc = 0;
    for (int i = 0; i < d; i++) acc = acc + a[i] * a[i];
    return acc;
}
/* cust_vector.hpp:126-136 */
inline double euclid(const double* a, const double* b, int d) {
    double acc = 0;
    for (int i = 0; i < d; i++) {
        double t = a[i] - b[i];
        acc = acc + t * t;
    }
    return std::sqrt(acc);
}
/* cust_vector.hpp:160-174; na/nb = Σ squares of a / b */
inline double cos_sim_n(const double* a, const double* b, int d, double na, double nb) {
    long double ip = dot_ld(a, b, d);
    double denom = std::sqrt(na) * std::sqrt(nb);
    return (double)(ip / denom);
}
inline double cos_sim(const double* a, const double* b, int d) { return cos_sim_n(a, b, d, sqnorm(a, d), sqnorm(b, d)); }
/* cust_vector.hpp:141-155 */
inline double cos_dist(const double* a, const double* b, int d) { return 1 - cos_sim(a, b, d); }
inline double metric_dist(int metric, const double* a, const double* b, int d) {
    return metric == 0 ? euclid(a, b, d) : cos_dist(a, b, d);
}

/* utils.cpp:22-50: all numbers at Hamming distance `dist`, flipping bit positions
 * p1 < p2 < ... < p_dist (>= min_bit) in lexicographic order of the position tuple. */
void hamming_ball(int num, int dist, int min_bit, int bits, std::vector<int>& out) {
    if (dist < 1) return;
    std::vector<int> pos(dist);
    for (int i = 0; i < dist; i++) pos[i] = min_bit + i;
    if (dist > bits - min_bit) return;
    for (;;) {
        int v = num;
        for (int i = 0; i < dist; i++) v ^= (1 << pos[i]);
        out.push_back(v);
        int i = dist - 1;
        while (i >= 0 && pos[i] == bits - dist + i) i--;
        if (i < 0) break;
        pos[i]++;
        for (int j = i + 1; j < dist; j++) pos[j] = pos[j - 1] + 1;
    }
}

/* crypto_rec.hpp:235-277: Lomuto partition, pivot = last element, `>=` goes left; two parallel
 * arrays.  Written with an explicit stack; left part is always processed before the right one,
 * which does not matter for the result (the sub-ranges are disjoint). */
template <typename V, typename T>
void lomuto_desc(V* key, T* val, int lo0, int hi0) {
    std::vector<std::pair<int, int>> st;
    st.emplace_back(lo0, hi0);
    while (!st.empty()) {
        int lo = st.back().first, hi = st.back().second;
        st.pop_back();
        if (lo >= hi) continue;
        V pivot = key[hi];
        int i = lo - 1;
        for (int j = lo; j < hi; j++) {
            if (key[j] >= pivot) {
                i++;
                std::swap(key[i], key[j]);
                std::swap(val[i], val[j]);
            }
        }
        std::swap(key[i + 1], key[hi]);
        std::swap(val[i + 1], val[hi]);
        st.emplace_back(lo, i);
        st.emplace_back(i + 2, hi);
    }
}

/* ---------------- hash families ---------------- */

struct EuclidH { /* euclidean_h_gen.hpp:58-82 */
    std::vector<float> v;
    float t, w;
    void init(int D, float w_in, Engine& e) {
        w = w_in;
        std::normal_distribution<float> nd(0, 1);
        v.resize(D);
        for (int i = 0; i < D; i++) v[i] = nd(e);
        std::uniform_real_distribution<float> ud(0, w);
        t = ud(e);
    }
    int h(const double* x, int D) const {
        long double ip = dot_ld_f(v.data(), x, D);
        return (int)std::floor((ip + t) / w);
    }
};

struct CosineH { /* cosine_h_gen.hpp:53-76 */
    std::vector<double> r;
    void init(int D, Engine& e) {
        std::normal_distribution<double> nd(0, 1);
        r.resize(D);
        for (int i = 0; i < D; i++) r[i] = nd(e);
    }
    int bit(const double* x, int D) const { return dot_ld(r.data(), x, D) >= 0 ? 1 : 0; }
};

const int PHI_M = 2147483647; /* int(pow(2,32)-5) as GCC constant-folds it (euclidean_phi_gen.hpp:70) */

struct LshTable {
    int metric, k;
    std::vector<EuclidH> eh;
    std::vector<int> rs;
    std::vector<CosineH> ch;
    size_t nbuckets;
    void init(int metric_, int k_, int D, float w, size_t nb, Engine& e) {
        metric = metric_; k = k_; nbuckets = nb;
        if (metric == 0) { /* euclidean_phi_gen.hpp:60-71 */
            std::uniform_int_distribution<int> ui(0, 100);
            eh.resize(k); rs.resize(k);
            for (int i = 0; i < k; i++) { eh[i].init(D, w, e); rs[i] = ui(e); }
        } else { /* cosine_g_gen.hpp:49-53 */
            ch.resize(k);
            for (int i = 0; i < k; i++) ch[i].init(D, e);
        }
    }
    /* generator value; det receives the k h-values for euclidean */
    int generate(const double* x, int D, int* det) const {
        if (metric == 0) { /* euclidean_phi_gen.hpp:82-97 */
            unsigned int acc = 0;
            for (int i = 0; i < k; i++) {
                int hi = eh[i].h(x, D);
                long temp = hi * rs[i];
                acc = acc + ref_mod(temp, PHI_M);
                if (det) det[i] = hi;
            }
            return ref_mod(acc, PHI_M);
        }
        int g = 0; /* cosine_g_gen.hpp:62-72 */
        for (int i = 0; i < k; i++) g = (g << 1) + ch[i].bit(x, D);
        return g;
    }
    /* cust_hashtable.hpp:68,123: mod in size_t arithmetic */
    int bucket_of(int gen) const { return ref_mod(gen, nbuckets); }
};

struct Lsh {
    int metric, k, L, D;
    int64_t N;
    std::vector<LshTable> tab;
    std::vector<int32_t> ids;  /* [L][N] bucket index */
    std::vector<int32_t> det;  /* [L][N][k] euclidean h tuples */
    std::vector<std::vector<std::vector<int32_t>>> buckets; /* [L][bucket] -> rows in insertion order */
    /* lsh_cube.hpp:45-74 */
    void build(const double* X, int64_t N_, int D_, int metric_, int k_, int L_, int div, double w, uint64_t seed) {
        metric = metric_; k = k_; L = L_; D = D_; N = N_;
        Engine e;
        e.seed((unsigned long)seed);
        tab.resize(L);
        ids.assign((size_t)L * N, 0);
        if (metric == 0) det.assign((size_t)L * N * k, 0);
        buckets.resize(L);
        for (int l = 0; l < L; l++) {
            size_t nb = metric == 0 ? (size_t)N / (size_t)div : (size_t)(int)std::pow(2, k);
            tab[l].init(metric, k, D, (float)w, nb, e);
            buckets[l].assign(nb, std::vector<int32_t>());
            for (int64_t i = 0; i < N; i++) {
                int g = tab[l].generate(X + i * D, D, metric == 0 ? &det[((size_t)l * N + i) * k] : nullptr);
                int b = tab[l].bucket_of(g);
                ids[(size_t)l * N + i] = b;
                buckets[l][b].push_back((int32_t)i);
            }
        }
    }
    /* lsh_cube.hpp:78-106 + cust_hashtable.hpp:74-113.  Candidates in ascending row order
     * (std::set of pointers into one contiguous vector).  `self` >= 0: the query is that base row. */
    void candidates(const double* q, int filtered, std::vector<int32_t>& out) const {
        std::vector<int32_t> all;
        std::vector<int> qd(k);
        for (int l = 0; l < L; l++) {
            int g = tab[l].generate(q, D, metric == 0 ? qd.data() : nullptr);
            int b = tab[l].bucket_of(g);
            const std::vector<int32_t>& bk = buckets[l][b];
            if (filtered && metric == 0) {
                for (int32_t v : bk) {
                    const int32_t* dv = &det[((size_t)l * N + v) * k];
                    bool same = true;
                    for (int j = 0; j < k; j++)
                        if (dv[j] != qd[j]) { same = false; break; }
                    if (same) all.push_back(v);
                }
            } else {
                all.insert(all.end(), bk.begin(), bk.end());
            }
        }
        std::sort(all.begin(), all.end());
        all.erase(std::unique(all.begin(), all.end()), all.end());
        out.swap(all);
    }
};

struct Cube {
    int metric, k, D;
    int64_t N;
    std::vector<EuclidH> eh;
    std::vector<std::unordered_map<int, int>> fmap; /* euclidean_f_gen.hpp:33 */
    std::vector<CosineH> ch;
    std::vector<int32_t> ids;
    std::vector<std::vector<int32_t>> buckets;
    Engine e;
    int vertex(const double* x) { /* hypercube_gen.hpp:63-73 */
        int g = 0;
        for (int i = 0; i < k; i++) {
            int bit;
            if (metric == 0) { /* euclidean_f_gen.hpp:65-79 */
                int h = eh[i].h(x, D);
                auto it = fmap[i].find(h);
                if (it != fmap[i].end()) bit = it->second;
                else {
                    std::uniform_int_distribution<int> u12(1, 2);
                    bit = ref_mod(h, u12(e));
                    fmap[i].emplace(h, bit);
                }
            } else bit = ch[i].bit(x, D);
            g = (g << 1) + bit;
        }
        return g;
    }
    /* lsh_cube.hpp:109-136 */
    void build(const double* X, int64_t N_, int D_, int metric_, int k_, double w, uint64_t seed) {
        metric = metric_; k = k_; D = D_; N = N_;
        e.seed((unsigned long)seed);
        if (metric == 0) {
            eh.resize(k); fmap.resize(k);
            for (int i = 0; i < k; i++) eh[i].init(D, (float)w, e);
        } else {
            ch.resize(k);
            for (int i = 0; i < k; i++) ch[i].init(D, e);
        }
        size_t nb = (size_t)(int)std::pow(2, k);
        buckets.assign(nb, std::vector<int32_t>());
        ids.resize(N);
        for (int64_t i = 0; i < N; i++) {
            int b = ref_mod(vertex(X + i * D), nb);
            ids[i] = b;
            buckets[b].push_back((int32_t)i);
        }
    }
    /* lsh_cube.hpp:140-177 */
    void probe(int home, int probes, std::vector<int32_t>& out) const {
        out = buckets[home];
        std::vector<int> neigh;
        size_t ni = 0;
        if (probes > 1) hamming_ball(home, 1, 0, k, neigh);
        int left = probes, dist = 1;
        while (left > 0) {
            if (ni < neigh.size()) {
                const std::vector<int32_t>& b = buckets[neigh[ni]];
                out.insert(out.end(), b.begin(), b.end());
                ni++; left--;
            } else {
                dist++;
                neigh.clear();
                hamming_ball(home, dist, 0, k, neigh);
                ni = 0;
                if (neigh.empty()) break;
            }
        }
    }
};

/* utils.hpp:161-178 */
double min_pair_distance(const double* X, int D, const int32_t* cidx, int K, int metric) {
    double m = -1;
    for (int a = 0; a < K; a++)
        for (int b = a + 1; b < K; b++) {
            double d = metric_dist(metric, X + (size_t)cidx[a] * D, X + (size_t)cidx[b] * D, D);
            if (m == -1 || d < m) m = d;
        }
    return m;
}

/* assignment.hpp:156-217 with data-point centroids; labels/dists hold the CustVector state */
void range_assign(const double* X, int D, const int32_t* cidx, int K, int metric,
                  const std::vector<std::vector<int32_t>>& comb, int32_t* labels, double* dists) {
    double radius = min_pair_distance(X, D, cidx, K, metric) / 2;
    double min_radius = 0;
    int assigned;
    do {
        assigned = 0;
        for (int c = 0; c < K; c++) {
            const double* cv = X + (size_t)cidx[c] * D;
            for (int32_t v : comb[c]) {
                if (labels[v] == -1 || dists[v] >= min_radius) {
                    double d = metric_dist(metric, cv, X + (size_t)v * D, D);
                    if (d >= min_radius && d < radius) {
                        if (labels[v] == -1) { labels[v] = c; dists[v] = d; assigned++; }
                        else if (dists[v] > d) { labels[v] = c; dists[v] = d; assigned++; }
                    }
                }
            }
            min_radius = radius;
            radius = radius * 2;
        }
    } while (assigned > 0);
}

/* assignment.hpp:156-217 as written, for centroids that are any vectors: the distance cache keyed by
 * "<centroid id>to<vector id>" (:183-194) is kept; idgroup[c] stands for the id string of centroid c (equal numbers =
 * equal strings, e.g. every k_means centre is "k_means_center", update.hpp:48) */
void range_assign_cached(const double* X, int64_t N, int D, const double* C, const int* idgroup, int K, int metric,
                         const std::vector<std::vector<int32_t>>& comb, int32_t* labels, double* dists) {
    std::vector<int32_t> iota(K);
    for (int c = 0; c < K; c++) iota[c] = c;
    double radius = min_pair_distance(C, D, iota.data(), K, metric) / 2;
    double min_radius = 0;
    std::unordered_map<uint64_t, double> cache;
    int assigned;
    do {
        assigned = 0;
        for (int c = 0; c < K; c++) {
            const double* cv = C + (size_t)c * D;
            for (int32_t v : comb[c]) {
                if (labels[v] == -1 || dists[v] >= min_radius) {
                    uint64_t key = (uint64_t)idgroup[c] * (uint64_t)N + (uint64_t)v;
                    double d;
                    auto it = cache.find(key);
                    if (it != cache.end()) d = it->second;
                    else { d = metric_dist(metric, cv, X + (size_t)v * D, D); cache[key] = d; }
                    if (d >= min_radius && d < radius) {
                        if (labels[v] == -1) { labels[v] = c; dists[v] = d; assigned++; }
                        else if (dists[v] > d) { labels[v] = c; dists[v] = d; assigned++; }
                    }
                }
            }
            min_radius = radius;
            radius = radius * 2;
        }
    } while (assigned > 0);
}

/* assignment.hpp:84-105 */
void lloyd_remaining(const double* X, int64_t N, int D, const double* C, int K, int metric, bool only_unassigned,
                     int32_t* labels, double* dists) {
    for (int64_t v = 0; v < N; v++) {
        if (only_unassigned && labels[v] != -1) continue;
        double mn = -1;
        int arg = 0;
        for (int c = 0; c < K; c++) {
            double d = metric_dist(metric, X + v * D, C + (size_t)c * D, D);
            if (mn == -1 || d < mn) { mn = d; arg = c; }
        }
        labels[v] = arg;
        dists[v] = mn;
    }
}

std::vector<double> gather_rows(const double* X, int D, const int32_t* idx, int K) {
    std::vector<double> C((size_t)K * D);
    for (int c = 0; c < K; c++) std::memcpy(&C[(size_t)c * D], X + (size_t)idx[c] * D, sizeof(double) * D);
    return C;
}

/* crypto_rec.hpp:281-345: predicted scores of the unknown coins from (neighbour, sim) lists,
 * Lomuto co-sort, resize(N) (pads with coin 0). */
void top_n_from_neighbours(const double* X, const double* mean, int D, const int32_t* nbr, const double* sim, int n_nbr,
                           const uint8_t* unk_q, double mean_q, int Nrec, int32_t* recs) {
    std::vector<int> uidx;
    std::vector<double> pred;
    for (int j = 0; j < D; j++) {
        if (!unk_q[j]) continue;
        double main_sum = 0, abs_sum = 0;
        for (int i = 0; i < n_nbr; i++) {
            double s = sim[i];
            abs_sum = abs_sum + std::fabs(s);
            main_sum = main_sum + (s * (X[(size_t)nbr[i] * D + j] - mean[nbr[i]]));
        }
        double p = main_sum / abs_sum;
        p = p + mean_q;
        uidx.push_back(j);
        pred.push_back(p);
    }
    lomuto_desc(pred.data(), uidx.data(), 0, (int)pred.size() - 1);
    uidx.resize(Nrec);
    for (int j = 0; j < Nrec; j++) recs[j] = uidx[j];
}

} // namespace

extern "C" {

const char* orc_kind(void) { return "port"; }

int orc_mod_ii(int x, int n) { return ref_mod(x, n); }
int orc_mod_li(long x, int n) { return ref_mod(x, n); }
int orc_mod_iz(int x, size_t n) { return ref_mod(x, n); }
int orc_mod_ui(unsigned int x, int n) { return ref_mod(x, n); }

int orc_hamming(int num, int dist, int min_bit, int bits, int* out, int cap) {
    std::vector<int> r;
    hamming_ball(num, dist, min_bit, bits, r);
    for (size_t i = 0; i < r.size() && (int)i < cap; i++) out[i] = r[i];
    return (int)r.size();
}

void orc_quicksort(double* sims, int* ids, int n) { lomuto_desc(sims, ids, 0, n - 1); }

void orc_rng_kat(uint64_t seed, double* nd3, float* nf2, float* uf1, int* ui1, int* u12_6) {
    {
        Engine e;
        e.seed((unsigned long)seed);
        std::normal_distribution<double> nd(0, 1);
        for (int i = 0; i < 3; i++) nd3[i] = nd(e);
    }
    Engine e;
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

double orc_inner_product(const double* a, const double* b, int d) { return (double)dot_ld(a, b, d); }
double orc_euclidean_distance(const double* a, const double* b, int d) { return euclid(a, b, d); }
double orc_cosine_distance(const double* a, const double* b, int d) { return cos_dist(a, b, d); }
double orc_cosine_similarity(const double* a, const double* b, int d) { return cos_sim(a, b, d); }

int orc_lsh_hash(const double* X, int64_t N, int D, int metric, int k, int L, int div, double w, uint64_t seed,
                 int32_t* bucket_ids, int32_t* det_hashes) {
    Lsh lsh;
    lsh.build(X, N, D, metric, k, L, div, w, seed);
    std::memcpy(bucket_ids, lsh.ids.data(), sizeof(int32_t) * lsh.ids.size());
    if (det_hashes && metric == 0) std::memcpy(det_hashes, lsh.det.data(), sizeof(int32_t) * lsh.det.size());
    return 0;
}

int64_t orc_lsh_candidates(const double* X, int64_t N, int D, int metric, int k, int L, int div, double w,
                           uint64_t seed, int64_t q, int filtered, int32_t* out, int64_t cap) {
    Lsh lsh;
    lsh.build(X, N, D, metric, k, L, div, w, seed);
    std::vector<int32_t> c;
    lsh.candidates(X + q * D, filtered, c);
    for (size_t i = 0; i < c.size() && (int64_t)i < cap; i++) out[i] = c[i];
    return (int64_t)c.size();
}

int orc_cube_hash(const double* X, int64_t N, int D, int metric, int k, double w, uint64_t seed, int32_t* ids) {
    Cube cube;
    cube.build(X, N, D, metric, k, w, seed);
    std::memcpy(ids, cube.ids.data(), sizeof(int32_t) * N);
    return 0;
}

int64_t orc_cube_candidates(const double* X, int64_t N, int D, int metric, int k, double w, uint64_t seed, int64_t q,
                            int probes, int32_t* out, int64_t cap) {
    Cube cube;
    cube.build(X, N, D, metric, k, w, seed);
    std::vector<int32_t> c;
    cube.probe(cube.ids[q], probes, c);
    for (size_t i = 0; i < c.size() && (int64_t)i < cap; i++) out[i] = c[i];
    return (int64_t)c.size();
}

/* initialization.hpp:40-68 */
int orc_rand_selection(const double* X, int64_t N, int D, int K, uint64_t seed, int32_t* idx) {
    (void)X; (void)D;
    Engine e;
    e.seed((unsigned long)seed);
    std::uniform_int_distribution<int> ui(0, (int)N - 1);
    idx[0] = ui(e);
    for (int i = 1; i < K; i++) {
        int r;
        bool clash;
        do {
            r = ui(e);
            clash = false;
            for (int j = 0; j < i; j++)
                if (idx[j] == r) { clash = true; break; }
        } while (clash);
        idx[i] = r;
    }
    return 0;
}

/* initialization.hpp:72-156 */
int orc_k_means_pp(const double* X, int64_t N, int D, int K, int metric, uint64_t seed, int32_t* idx) {
    Engine e;
    e.seed((unsigned long)seed);
    std::uniform_int_distribution<int> ui(0, (int)N - 1);
    idx[0] = ui(e);
    std::vector<double> mind(N), P(N);
    for (int i = 1; i < K; i++) {
        const double* cnew = X + (size_t)idx[i - 1] * D;
        double mx = 0;
        for (int64_t v = 0; v < N; v++) {
            double d = metric_dist(metric, X + v * D, cnew, D);
            if (i == 1 || d < mind[v]) mind[v] = d; /* running form of the `min == -1 || d < min` scan */
            if (mind[v] > mx) mx = mind[v];
        }
        for (int64_t v = 0; v < N; v++) {
            double t = mind[v] / mx;
            t = t * t;
            P[v] = v == 0 ? t : t + P[v - 1];
        }
        std::uniform_real_distribution<double> ur(0.0, P[N - 1]);
        double x = ur(e);
        int64_t lo = 0, hi = N - 1, chosen = 0;
        if (x > P[lo]) {
            while (hi - lo > 1) {
                int64_t m = lo + (hi - lo) / 2;
                if (x <= P[m]) hi = m; else lo = m;
            }
            chosen = hi;
        }
        idx[i] = (int32_t)chosen;
    }
    return 0;
}

/* assignment.hpp:55-80 */
int orc_lloyds_assignment(const double* X, int64_t N, int D, const double* C, int K, const int32_t* cidx, int metric,
                          int32_t* labels, double* dists) {
    std::vector<double> Cs;
    if (!C) { Cs = gather_rows(X, D, cidx, K); C = Cs.data(); }
    lloyd_remaining(X, N, D, C, K, metric, false, labels, dists);
    if (cidx)
        for (int c = 0; c < K; c++)
            if (cidx[c] >= 0) { labels[cidx[c]] = c; dists[cidx[c]] = 0; }
    return 0;
}

/* assignment.hpp:109-129 */
int orc_lsh_range_assignment(const double* X, int64_t N, int D, const int32_t* cidx, int K, int metric, int k, int L,
                             int div, double w, uint64_t seed, int32_t* labels, double* dists, int32_t* before) {
    Lsh lsh;
    lsh.build(X, N, D, metric, k, L, div, w, seed);
    for (int64_t i = 0; i < N; i++) { labels[i] = -1; dists[i] = 0; }
    std::vector<std::vector<int32_t>> comb(K);
    for (int c = 0; c < K; c++) lsh.candidates(X + (size_t)cidx[c] * D, 0, comb[c]);
    range_assign(X, D, cidx, K, metric, comb, labels, dists);
    if (before) std::memcpy(before, labels, sizeof(int32_t) * N);
    std::vector<double> C = gather_rows(X, D, cidx, K);
    lloyd_remaining(X, N, D, C.data(), K, metric, true, labels, dists);
    for (int c = 0; c < K; c++) { labels[cidx[c]] = c; dists[cidx[c]] = 0; }
    return 0;
}

/* assignment.hpp:109-129 with heap centroids (the second and later iterations of {range assignment, k_means}) */
int orc_lsh_range_assignment_vectors(const double* X, int64_t N, int D, const double* C, const int32_t* cidx, int K, int shared_ids,
                                     int metric, int k, int L, int div, double w, uint64_t seed, int32_t* labels, double* dists,
                                     int32_t* before) {
    Lsh lsh;
    lsh.build(X, N, D, metric, k, L, div, w, seed);
    for (int64_t i = 0; i < N; i++) { labels[i] = -1; dists[i] = 0; }
    std::vector<double> Cm((size_t)K * D);
    std::vector<int> group(K);
    for (int c = 0; c < K; c++) {
        bool stored = cidx && cidx[c] >= 0;
        std::memcpy(&Cm[(size_t)c * D], stored ? X + (size_t)cidx[c] * D : C + (size_t)c * D, sizeof(double) * D);
        group[c] = stored ? 1 + cidx[c] : (shared_ids ? 0 : (int)N + 1 + c);   /* stored rows keep their own unique ids */
    }
    std::vector<std::vector<int32_t>> comb(K);
    for (int c = 0; c < K; c++) lsh.candidates(&Cm[(size_t)c * D], 0, comb[c]);
    range_assign_cached(X, N + K + 2, D, Cm.data(), group.data(), K, metric, comb, labels, dists);
    if (before) std::memcpy(before, labels, sizeof(int32_t) * N);
    lloyd_remaining(X, N, D, Cm.data(), K, metric, true, labels, dists);
    for (int c = 0; c < K; c++)
        if (cidx && cidx[c] >= 0) { labels[cidx[c]] = c; dists[cidx[c]] = 0; }
    return 0;
}

/* assignment.hpp:132-152 */
int orc_cube_range_assignment(const double* X, int64_t N, int D, const int32_t* cidx, int K, int metric, int k, double w,
                              int probes, uint64_t seed, int32_t* labels, double* dists, int32_t* before) {
    Cube cube;
    cube.build(X, N, D, metric, k, w, seed);
    for (int64_t i = 0; i < N; i++) { labels[i] = -1; dists[i] = 0; }
    std::vector<std::vector<int32_t>> comb(K);
    for (int c = 0; c < K; c++) cube.probe(cube.ids[cidx[c]], probes, comb[c]);
    range_assign(X, D, cidx, K, metric, comb, labels, dists);
    if (before) std::memcpy(before, labels, sizeof(int32_t) * N);
    std::vector<double> C = gather_rows(X, D, cidx, K);
    lloyd_remaining(X, N, D, C.data(), K, metric, true, labels, dists);
    for (int c = 0; c < K; c++) { labels[cidx[c]] = c; dists[cidx[c]] = 0; }
    return 0;
}

/* update.hpp:38-86; newC = centres after the call (unchanged when it returns 0) */
int orc_k_means(const double* X, int64_t N, int D, const int32_t* labels, const double* C, int K, int metric,
                double min_dist, double* newC) {
    std::vector<double> S((size_t)K * D, 0.0);
    std::vector<int> cnt(K, 0);
    for (int64_t v = 0; v < N; v++) {
        int c = labels[v];
        cnt[c]++;
        double* s = &S[(size_t)c * D];
        for (int j = 0; j < D; j++) s[j] = s[j] + X[v * D + j];
    }
    for (int c = 0; c < K; c++) {
        double div = cnt[c];
        if (div != 0)
            for (int j = 0; j < D; j++) S[(size_t)c * D + j] = S[(size_t)c * D + j] / div;
    }
    for (int c = 0; c < K; c++) {
        double d = metric_dist(metric, &S[(size_t)c * D], C + (size_t)c * D, D);
        if (d > min_dist) {
            std::memcpy(newC, S.data(), sizeof(double) * S.size());
            return 1;
        }
    }
    std::memcpy(newC, C, sizeof(double) * (size_t)K * D);
    return 0;
}

/* update.hpp:90-142 */
int orc_pam_lloyds(const double* X, int64_t N, int D, const int32_t* labels, const int32_t* cidx, int K, int metric,
                   int32_t* new_cidx) {
    std::vector<std::vector<int32_t>> members(K);
    for (int64_t v = 0; v < N; v++) members[labels[v]].push_back((int32_t)v); /* utils.hpp:150-158 */
    int swapped = 0;
    for (int c = 0; c < K; c++) {
        const std::vector<int32_t>& m = members[c];
        new_cidx[c] = cidx[c];
        if (m.empty()) continue; /* the reference indexes clusters[c][0] here (UB); treated as "no change" */
        double best = -1;
        int arg = 0;
        for (size_t a = 0; a < m.size(); a++) {
            double s = 0;
            for (size_t b = 0; b < m.size(); b++) s = s + metric_dist(metric, X + (size_t)m[a] * D, X + (size_t)m[b] * D, D);
            if (best == -1 || s < best) { best = s; arg = (int)a; }
        }
        if (m[arg] != cidx[c]) { new_cidx[c] = m[arg]; swapped = 1; }
    }
    return swapped;
}

/* silhouette.hpp:32-144 */
int orc_silhouette(const double* X, int64_t N, int D, const int32_t* labels, const double* C, int K, int metric,
                   double* sils) {
    std::vector<std::vector<int32_t>> members(K);
    for (int64_t v = 0; v < N; v++) members[labels[v]].push_back((int32_t)v);
    std::vector<int> near(K, 0);
    for (int a = 0; a < K; a++) {
        double mn = -1;
        int arg = 0;
        for (int b = 0; b < K; b++) {
            if (b == a) continue;
            double d = metric_dist(metric, C + (size_t)a * D, C + (size_t)b * D, D);
            if (mn == -1 || d < mn) { mn = d; arg = b; }
        }
        near[a] = arg;
    }
    sils[K] = 0;
    int64_t total = 0;
    for (int c = 0; c < K; c++) {
        const std::vector<int32_t>& own = members[c];
        const std::vector<int32_t>& nb = members[near[c]];
        sils[c] = 0;
        for (size_t a = 0; a < own.size(); a++) {
            const double* xa = X + (size_t)own[a] * D;
            double ai = 0;
            for (size_t b = 0; b < own.size(); b++) ai = ai + metric_dist(metric, xa, X + (size_t)own[b] * D, D);
            if (own.size() != 1) ai = ai / (own.size() - 1);
            double bi = 0;
            for (size_t b = 0; b < nb.size(); b++) bi = bi + metric_dist(metric, xa, X + (size_t)nb[b] * D, D);
            bi = bi / nb.size();
            double mx = ai;
            if (bi > ai) mx = bi;
            sils[c] = sils[c] + (bi - ai) / mx;
        }
        sils[K] = sils[K] + sils[c];
        sils[c] = sils[c] / own.size();
        total += (int64_t)own.size();
    }
    sils[K] = sils[K] / total;
    return 0;
}

/* main.cpp:155-170 / 201-216 with crypto_rec.hpp:214-231, 281-324 */
int orc_recommend_lsh(const double* X, const uint8_t* unknown, const double* mean, int64_t N, int D,
                      const double* Xq, const uint8_t* unknown_q, const double* mean_q, int64_t Nq,
                      int metric, int k, int L, int div, double w, int P, int Nrec, uint64_t seed,
                      int32_t* recs, int32_t* nbr_idx, double* nbr_sim, int32_t* ncand) {
    Lsh lsh;
    lsh.build(X, N, D, metric, k, L, div, w, seed);
    if (!Xq) { Xq = X; unknown_q = unknown; mean_q = mean; Nq = N; }
    std::vector<double> nrm(N);
    for (int64_t i = 0; i < N; i++) nrm[i] = sqnorm(X + i * D, D); /* value-transparent cache of cust_vector.hpp:168-171 */
    std::vector<int32_t> cand;
    std::vector<double> sims;
    for (int64_t u = 0; u < Nq; u++) {
        const double* q = Xq + u * D;
        lsh.candidates(q, 1, cand);
        ncand[u] = (int32_t)cand.size();
        for (int j = 0; j < P; j++) { nbr_idx[u * P + j] = -1; nbr_sim[u * P + j] = 0; }
        for (int j = 0; j < Nrec; j++) recs[u * Nrec + j] = -1;
        if (cand.empty()) continue;
        double nq = sqnorm(q, D);
        sims.resize(cand.size());
        for (size_t i = 0; i < cand.size(); i++) sims[i] = cos_sim_n(X + (size_t)cand[i] * D, q, D, nrm[cand[i]], nq);
        lomuto_desc(sims.data(), cand.data(), 0, (int)cand.size() - 1);
        int keep = (int)std::min<size_t>(cand.size(), (size_t)P);
        for (int j = 0; j < keep; j++) { nbr_idx[u * P + j] = cand[j]; nbr_sim[u * P + j] = sims[j]; }
        top_n_from_neighbours(X, mean, D, cand.data(), sims.data(), keep, unknown_q + u * D, mean_q[u], Nrec,
                              recs + u * Nrec);
    }
    return 0;
}

/* main.cpp:260-269 / 353-373 with crypto_rec.hpp:328-345 */
int orc_recommend_cluster(const double* X, const uint8_t* unknown, const double* mean, const int32_t* labels,
                          int64_t N, int D, int K,
                          const double* Xq, const uint8_t* unknown_q, const double* mean_q, const int32_t* qlabels,
                          int64_t Nq, int Nrec, int32_t* recs) {
    if (!Xq) { Xq = X; unknown_q = unknown; mean_q = mean; qlabels = labels; Nq = N; }
    std::vector<std::vector<int32_t>> members(K);
    for (int64_t v = 0; v < N; v++) members[labels[v]].push_back((int32_t)v);
    std::vector<double> nrm(N);
    for (int64_t i = 0; i < N; i++) nrm[i] = sqnorm(X + i * D, D);
    std::vector<double> sims;
    for (int64_t u = 0; u < Nq; u++) {
        const std::vector<int32_t>& nb = members[qlabels[u]];
        for (int j = 0; j < Nrec; j++) recs[u * Nrec + j] = -1;
        if (nb.empty()) continue;
        const double* q = Xq + u * D;
        double nq = sqnorm(q, D);
        sims.resize(nb.size());
        for (size_t i = 0; i < nb.size(); i++) sims[i] = cos_sim_n(X + (size_t)nb[i] * D, q, D, nrm[nb[i]], nq);
        top_n_from_neighbours(X, mean, D, nb.data(), sims.data(), (int)nb.size(), unknown_q + u * D, mean_q[u], Nrec,
                              recs + u * Nrec);
    }
    return 0;
}

struct PortRecHandle {
    Lsh lsh;
    const double* X; const uint8_t* unknown; const double* mean;
    std::vector<double> nrm;
    int D;
};

void* orc_rec_handle_create(const double* X, const uint8_t* unknown, const double* mean, int64_t N, int D, int metric,
                            int k, int L, int div, double w, uint64_t seed) {
    PortRecHandle* h = new PortRecHandle();
    h->X = X; h->unknown = unknown; h->mean = mean; h->D = D;  /* caller keeps the arrays alive */
    h->lsh.build(X, N, D, metric, k, L, div, w, seed);
    h->nrm.resize(N);
    for (int64_t i = 0; i < N; i++) h->nrm[i] = sqnorm(X + i * D, D);
    return h;
}

int orc_rec_handle_query_nbr(void* hv, int64_t q_begin, int64_t q_end, int P, int Nrec, int32_t* recs, int32_t* ncand,
                             int32_t* nbr_idx, double* nbr_sim) {
    PortRecHandle* h = (PortRecHandle*)hv;
    int D = h->D;
    std::vector<int32_t> cand;
    std::vector<double> sims;
    for (int64_t u = q_begin; u < q_end; u++) {
        const double* q = h->X + u * D;
        h->lsh.candidates(q, 1, cand);
        ncand[u - q_begin] = (int32_t)cand.size();
        for (int j = 0; j < Nrec; j++) recs[(u - q_begin) * Nrec + j] = -1;
        if (nbr_idx) for (int j = 0; j < P; j++) { nbr_idx[(u - q_begin) * P + j] = -1; nbr_sim[(u - q_begin) * P + j] = 0; }
        if (cand.empty()) continue;
        sims.resize(cand.size());
        for (size_t i = 0; i < cand.size(); i++) sims[i] = cos_sim_n(h->X + (size_t)cand[i] * D, q, D, h->nrm[cand[i]], h->nrm[u]);
        lomuto_desc(sims.data(), cand.data(), 0, (int)cand.size() - 1);
        int keep = (int)std::min<size_t>(cand.size(), (size_t)P);
        if (nbr_idx) for (int j = 0; j < keep; j++) { nbr_idx[(u - q_begin) * P + j] = cand[j]; nbr_sim[(u - q_begin) * P + j] = sims[j]; }
        top_n_from_neighbours(h->X, h->mean, D, cand.data(), sims.data(), keep, h->unknown + u * D, h->mean[u], Nrec,
                              recs + (u - q_begin) * Nrec);
    }
    return 0;
}

int orc_rec_handle_query(void* hv, int64_t q_begin, int64_t q_end, int P, int Nrec, int32_t* recs, int32_t* ncand) {
    return orc_rec_handle_query_nbr(hv, q_begin, q_end, P, Nrec, recs, ncand, nullptr, nullptr);
}

void orc_rec_handle_destroy(void* hv) { delete (PortRecHandle*)hv; }

} /* extern "C" */
