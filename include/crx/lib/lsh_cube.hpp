// Drop-in for lib/lsh_cube.hpp (reference lsh_cube.hpp:28-177): same free functions, same signatures.
#ifndef LSH_CUBE_HPP
#define LSH_CUBE_HPP

#include <algorithm>
#include <set>
#include <string>
#include <vector>

#include "./data_structures/cust_hashtable.hpp"
#include "./data_structures/cust_vector.hpp"
#include "utils.hpp"

// lsh_cube.hpp:45-74.  The L returned objects share one GPU table set; delete each as the reference does.
template <typename vector_type>
std::vector<CustHashtable<vector_type>*> create_LSH_hashtables(std::vector<CustVector<vector_type> >& input_vectors, const std::string metric_type,
                                                               int k, int L, int lsh_bucket_div, double euclidean_h_w) {
    auto set = std::make_shared<crx::TableSet<vector_type> >();
    set->base = &input_vectors;
    set->metric = crx::metric_code(metric_type);
    set->k = k;
    set->L = L;
    set->pts.from_vector(input_vectors, true);   // with the rating metadata: the recommendation calls reuse these rows
    if (!input_vectors.empty()) crx::register_points(&input_vectors[0], input_vectors.size(), sizeof(CustVector<vector_type>), set->pts.pts, set.get());
    crx::check(crx_create_LSH_hashtables(crx::context(), set->pts.pts, set->metric, k, L, lsh_bucket_div, euclidean_h_w, crx::next_seed(), &set->lsh),
               "crx_create_LSH_hashtables");
    std::vector<CustHashtable<vector_type>*> tables;
    tables.reserve(L);
    for (int l = 0; l < L; l++) tables.emplace_back(new CustHashtable<vector_type>(set, l));
    return tables;
}

namespace crx {
template <typename T>
std::vector<CustVector<T>*> combined(std::vector<CustHashtable<T>*>& tables, CustVector<T>* q, bool filtered) {
    std::vector<CustVector<T>*> out;
    if (tables.empty()) return out;
    auto& set = tables[0]->set;
    int32_t row = row_of(*set->base, q);
    if ((int)tables.size() == set->L && set->lsh && set->base && !set->base->empty()) {
        // all tables of the set.  The bucket ids (and h tuples) of the stored rows were computed by the engine when the tables
        // were built and are mirrored on the host once (TableSet::load); a vector that is not stored is hashed for every
        // table in ONE engine call.  What is left is the reference's own bookkeeping: the union of L member lists.
        Timed timed(row >= 0 ? "combined buckets (stored row)" : "combined buckets (other vector)");
        set->load();
        int64_t N = set->n();
        const bool tuples = filtered && set->metric == CRX_EUCLIDEAN;
        std::vector<int32_t> b(set->L), d(tuples || row < 0 ? (size_t)set->L * set->k : 0);
        const Registered* ext_home = nullptr;
        int64_t ext_row = -1;
        if (row >= 0) {
            for (int l = 0; l < set->L; l++) {
                b[l] = set->ids[(size_t)l * N + row];
                if (tuples) std::copy(&set->det[((size_t)l * N + row) * set->k], &set->det[((size_t)l * N + row) * set->k] + set->k, d.begin() + (size_t)l * set->k);
            }
        } else {
            // a user of another vector: when that vector has a provably current device copy, all its rows are hashed in one call
            const Registered* home = home_of(q, sizeof(CustVector<T>));
            if (home && set->ext_hash_for(home)) {
                ext_row = (int64_t)(((const char*)q - home->begin) / (ptrdiff_t)home->stride);
                ext_home = home;
                for (int l = 0; l < set->L; l++) {
                    b[l] = set->ext_ids[(size_t)l * home->n + ext_row];
                    if (set->metric == CRX_EUCLIDEAN)
                        std::copy(&set->ext_det[((size_t)l * home->n + ext_row) * set->k], &set->ext_det[((size_t)l * home->n + ext_row) * set->k] + set->k, d.begin() + (size_t)l * set->k);
                }
            } else {
                std::vector<double> x(q->crxDimsRef().begin(), q->crxDimsRef().end());
                check(crx_lsh_hash_vector(set->lsh, x.data(), b.data(), set->metric == CRX_EUCLIDEAN ? d.data() : nullptr), "crx_lsh_hash_vector");
                if (home) { ext_row = (int64_t)(((const char*)q - home->begin) / (ptrdiff_t)home->stride); ext_home = home; }
            }
        }
        std::vector<int32_t> rows;
        set->union_rows(b.data(), tuples ? d.data() : nullptr, rows);
        CustVector<T>* base = &(*set->base)[0];
        out.reserve(rows.size());
        for (int32_t r : rows) out.push_back(base + r);
        if (filtered && row >= 0) { set->own.last_row = row; set->own.last_list.swap(rows); }
        else if (filtered && ext_home) {
            if (set->ext.queries != ext_home->pts || set->ext.epoch != ext_home->gen) set->ext.reset(ext_home->pts, ext_home->gen);
            set->ext.last_row = ext_row; set->ext.last_list.swap(rows);
        }
        return out;
    }
    std::set<CustVector<T>*> u;  // pointer order == row order (lsh_cube.hpp:96,104)
    for (auto t : tables) {
        std::vector<CustVector<T>*> b = filtered ? t->getFilteredBucketFor(q) : t->getBucketFor(q);
        u.insert(b.begin(), b.end());
    }
    return std::vector<CustVector<T>*>(u.begin(), u.end());
}
}  // namespace crx

template <typename vector_type>
std::vector<CustVector<vector_type>*> get_LSH_combined_buckets(std::vector<CustHashtable<vector_type>*>& lshHashtables, CustVector<vector_type>* queryVec) {
    return crx::combined(lshHashtables, queryVec, false);  // lsh_cube.hpp:78-90
}

template <typename vector_type>
std::vector<CustVector<vector_type>*> get_LSH_filtered_combined_buckets(std::vector<CustHashtable<vector_type>*>& lshHashtables, CustVector<vector_type>* queryVec) {
    return crx::combined(lshHashtables, queryVec, true);  // lsh_cube.hpp:94-106
}

// lsh_cube.hpp:109-136
template <typename vector_type>
CustHashtable<vector_type>* create_hypercube(std::vector<CustVector<vector_type> >& input_vectors, const std::string metric_type, int k, double euclidean_h_w) {
    auto set = std::make_shared<crx::TableSet<vector_type> >();
    set->base = &input_vectors;
    set->metric = crx::metric_code(metric_type);
    set->k = k;
    set->L = 1;
    set->pts.from_vector(input_vectors, false);
    crx::check(crx_create_hypercube(crx::context(), set->pts.pts, set->metric, k, euclidean_h_w, crx::next_seed(), &set->cube), "crx_create_hypercube");
    return new CustHashtable<vector_type>(set, 0);
}

// lsh_cube.hpp:140-177: home vertex, then Hamming-1 vertices, Hamming-2, ... one probe per extra vertex
template <typename vector_type>
std::vector<CustVector<vector_type>*> get_hypercube_combined_buckets(CustHashtable<vector_type>& hypercube, CustVector<vector_type>* queryVec, int probes, int k) {
    auto& set = hypercube.set;
    std::vector<CustVector<vector_type>*> out;
    int32_t row = crx::row_of(*set->base, queryVec);
    if (row < 0 || k != set->k) {
        std::fprintf(stderr, "crx: get_hypercube_combined_buckets needs a stored query row and the cube's own k\n");
        std::abort();
    }
    std::vector<int32_t> rows((size_t)set->n());
    int64_t count = 0;
    crx::check(crx_get_hypercube_combined_buckets(set->cube, row, probes, rows.data(), (int64_t)rows.size(), &count), "crx_get_hypercube_combined_buckets");
    for (int64_t i = 0; i < count && i < (int64_t)rows.size(); i++) out.push_back(&(*set->base)[rows[i]]);
    return out;
}

#endif  // LSH_CUBE_HPP
