// Drop-in for lib/clustering_phases/assignment.hpp (reference assignment.hpp:26-217).
#ifndef CLUSTER_ASSIGNMENT_H
#define CLUSTER_ASSIGNMENT_H

#include <set>
#include <string>
#include <vector>

#include "../data_structures/cust_hashtable.hpp"
#include "../data_structures/cust_vector.hpp"
#include "../lsh_cube.hpp"
#include "../utils.hpp"

namespace crx {
template <typename T>
void pack_centroids(std::vector<CustVector<T> >& vecs, std::vector<CustVector<T>*>& centroids, std::vector<double>& C, std::vector<int32_t>& rows) {
    int K = (int)centroids.size();
    int D = K ? (int)centroids[0]->getDimNumber() : 0;
    C.resize((size_t)K * D);
    rows.resize(K);
    for (int c = 0; c < K; c++) {
        const std::vector<T>& d = centroids[c]->crxDimsRef();
        for (int j = 0; j < D; j++) C[(size_t)c * D + j] = (double)d[j];
        rows[c] = row_of(vecs, centroids[c]);
    }
}
template <typename T>
void write_back(std::vector<CustVector<T> >& vecs, const std::vector<int32_t>& labels, const std::vector<double>& dists) {
    for (size_t i = 0; i < vecs.size(); i++) vecs[i].setCluster(labels[i], dists[i]);
}
}  // namespace crx

// assignment.hpp:55-80 (centroids that alias input vectors get setCluster(c, 0) through centroid_rows)
template <typename vector_type>
void lloyds_assignment(std::vector<CustVector<vector_type> >& input_vectors, std::vector<CustVector<vector_type>*>& centroids, std::string metric_type) {
    crx::Timed timed("lloyds_assignment");
    crx::Packed<vector_type> P;
    P.from_vector(input_vectors);
    std::vector<double> C;
    std::vector<int32_t> rows, labels(input_vectors.size());
    std::vector<double> dists(input_vectors.size());
    crx::pack_centroids(input_vectors, centroids, C, rows);
    crx::check(crx_lloyds_assignment(crx::context(), P.pts, C.data(), CRX_HOST, (int)centroids.size(), rows.data(), crx::metric_code(metric_type),
                                     labels.data(), dists.data(), CRX_HOST), "crx_lloyds_assignment");
    crx::write_back(input_vectors, labels, dists);
    for (size_t c = 0; c < centroids.size(); c++) centroids[c]->setCluster((int)c, 0);  // also covers non-aliasing centres
}

// assignment.hpp:84-105
template <typename vector_type>
void lloyds_for_remaining(std::vector<CustVector<vector_type> >& input_vectors, std::vector<CustVector<vector_type>*>& centroids, std::string metric_type) {
    crx::Packed<vector_type> P;
    P.from_vector(input_vectors);
    std::vector<double> C;
    std::vector<int32_t> rows, labels(input_vectors.size());
    std::vector<double> dists(input_vectors.size());
    for (size_t i = 0; i < input_vectors.size(); i++) { labels[i] = input_vectors[i].getCluster(); dists[i] = input_vectors[i].getDistFromCentroid(); }
    crx::pack_centroids(input_vectors, centroids, C, rows);
    crx::check(crx_lloyds_for_remaining(crx::context(), P.pts, C.data(), CRX_HOST, (int)centroids.size(), crx::metric_code(metric_type), labels.data(),
                                        dists.data(), CRX_HOST), "crx_lloyds_for_remaining");
    crx::write_back(input_vectors, labels, dists);
}

namespace crx {
template <typename T>
std::vector<int32_t> stored_rows(std::vector<CustVector<T> >& vecs, std::vector<CustVector<T>*>& centroids, const char* who) {
    std::vector<int32_t> rows(centroids.size());
    for (size_t c = 0; c < centroids.size(); c++) {
        rows[c] = row_of(vecs, centroids[c]);
        if (rows[c] < 0) {
            std::fprintf(stderr, "crx: %s needs centroids that are stored input vectors (after rand_selection / k_means_pp / pam_lloyds)\n", who);
            std::abort();
        }
    }
    return rows;
}
}  // namespace crx

// assignment.hpp:109-129.  Centroids that are stored input vectors (after rand_selection / k_means_pp / pam_lloyds) go by row
// number; any other centroid -- the heap centres k_means leaves behind -- is shipped as a vector and hashed like a query.
// The reference keys its distance cache by "<centroid id>to<vector id>" (assignment.hpp:183-194): with unique centroid ids
// the cache changes nothing; when all centroids carry the same id (k_means names every centre "k_means_center",
// update.hpp:48) they share one cached distance per vector, and that is reproduced (crx_lsh_range_assignment_vectors).
template <typename vector_type>
void lsh_range_assignment(std::vector<CustVector<vector_type> >& input_vectors, std::vector<CustHashtable<vector_type>*>& lsh_hashtables,
                          std::vector<CustVector<vector_type>*>& centroids, std::string metric_type) {
    auto& set = lsh_hashtables.at(0)->set;
    std::vector<int32_t> labels(input_vectors.size());
    std::vector<double> dists(input_vectors.size());
    std::vector<double> C;
    std::vector<int32_t> rows;
    crx::pack_centroids(input_vectors, centroids, C, rows);
    bool all_stored = true;
    for (int32_t r : rows) all_stored = all_stored && r >= 0;
    if (all_stored) {
        crx::check(crx_lsh_range_assignment(crx::context(), set->pts.pts, set->lsh, rows.data(), (int)rows.size(), crx::metric_code(metric_type), labels.data(),
                                            dists.data(), CRX_HOST, nullptr), "crx_lsh_range_assignment");
    } else {
        std::set<std::string> ids;
        for (auto c : centroids) ids.insert(c->getId());
        if (ids.size() != centroids.size() && ids.size() != 1) {
            std::fprintf(stderr, "crx: lsh_range_assignment: centroid ids must be all different or all equal (%zu distinct among %zu)\n", ids.size(), centroids.size());
            std::abort();
        }
        int shared = centroids.size() > 1 && ids.size() == 1;
        crx::check(crx_lsh_range_assignment_vectors(crx::context(), set->pts.pts, set->lsh, C.data(), rows.data(), (int)rows.size(), crx::metric_code(metric_type),
                                                    shared, labels.data(), dists.data(), CRX_HOST, nullptr), "crx_lsh_range_assignment_vectors");
    }
    crx::write_back(input_vectors, labels, dists);
    for (size_t c = 0; c < centroids.size(); c++) centroids[c]->setCluster((int)c, 0);  // assignment.hpp:126-128, heap centres included
}

// assignment.hpp:132-152
template <typename vector_type>
void cube_range_assignment(std::vector<CustVector<vector_type> >& input_vectors, CustHashtable<vector_type>& hypercube,
                           std::vector<CustVector<vector_type>*>& centroids, std::string metric_type, int probes, int k) {
    auto& set = hypercube.set;
    (void)k;
    std::vector<int32_t> rows = crx::stored_rows(input_vectors, centroids, "cube_range_assignment"), labels(input_vectors.size());
    std::vector<double> dists(input_vectors.size());
    crx::check(crx_cube_range_assignment(crx::context(), set->pts.pts, set->cube, rows.data(), (int)rows.size(), crx::metric_code(metric_type), probes,
                                         labels.data(), dists.data(), CRX_HOST, nullptr), "crx_cube_range_assignment");
    crx::write_back(input_vectors, labels, dists);
}

#endif  // CLUSTER_ASSIGNMENT_H
