// Drop-in for lib/clustering_phases/initialization.hpp (reference initialization.hpp:24-156).
#ifndef CLUSTER_INITIALIZATION_H
#define CLUSTER_INITIALIZATION_H

#include <string>
#include <vector>

#include "../data_structures/cust_vector.hpp"

// initialization.hpp:40-68: returned pointers alias elements of input_vectors
template <typename vector_type>
std::vector<CustVector<vector_type>*> rand_selection(std::vector<CustVector<vector_type> >& input_vectors, int cluster_num) {
    crx::Packed<vector_type> P;
    P.from_vector(input_vectors);
    std::vector<int32_t> rows(cluster_num);
    crx::check(crx_rand_selection(crx::context(), P.pts, cluster_num, crx::next_seed(), rows.data()), "crx_rand_selection");
    std::vector<CustVector<vector_type>*> centroids(cluster_num);
    for (int i = 0; i < cluster_num; i++) centroids[i] = &input_vectors[rows[i]];
    return centroids;
}

// initialization.hpp:72-156
template <typename vector_type>
std::vector<CustVector<vector_type>*> k_means_pp(std::vector<CustVector<vector_type> >& input_vectors, int cluster_num, std::string metric_type) {
    crx::Timed timed("k_means_pp");
    crx::Packed<vector_type> P;
    P.from_vector(input_vectors);
    std::vector<int32_t> rows(cluster_num);
    crx::check(crx_k_means_pp(crx::context(), P.pts, cluster_num, crx::metric_code(metric_type), crx::next_seed(), rows.data()), "crx_k_means_pp");
    std::vector<CustVector<vector_type>*> centroids(cluster_num);
    for (int i = 0; i < cluster_num; i++) centroids[i] = &input_vectors[rows[i]];
    return centroids;
}

#endif  // CLUSTER_INITIALIZATION_H
