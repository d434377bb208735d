// Drop-in for lib/clustering_phases/update.hpp (reference update.hpp:21-142).
#ifndef CLUSTER_UPDATE_H
#define CLUSTER_UPDATE_H

#include <string>
#include <vector>

#include "../data_structures/cust_vector.hpp"
#include "../utils.hpp"

// update.hpp:38-86: on `true` every centre is replaced by a heap CustVector named "k_means_center" (previous ones
// of that name are deleted); on `false` the centres are left untouched.
template <typename vector_type>
bool k_means(std::vector<CustVector<vector_type> >& input_vectors, std::vector<CustVector<vector_type>*>& centers, std::string metric_type, double min_dist) {
    crx::Timed timed("k_means");
    int K = (int)centers.size();
    int D = K ? (int)centers[0]->getDimNumber() : 0;
    crx::Packed<vector_type> P;
    P.from_vector(input_vectors);
    std::vector<int32_t> labels(input_vectors.size());
    for (size_t i = 0; i < input_vectors.size(); i++) labels[i] = input_vectors[i].getCluster();
    std::vector<double> oldc((size_t)K * D), newc((size_t)K * D);
    for (int c = 0; c < K; c++) {
        const std::vector<vector_type>& d = centers[c]->crxDimsRef();
        for (int j = 0; j < D; j++) oldc[(size_t)c * D + j] = (double)d[j];
    }
    int cont = 0;
    crx::check(crx_k_means(crx::context(), P.pts, labels.data(), CRX_HOST, oldc.data(), K, crx::metric_code(metric_type), min_dist, newc.data(), CRX_HOST, &cont),
               "crx_k_means");
    if (!cont) return false;
    for (int c = 0; c < K; c++) {
        if (centers[c]->getId() == "k_means_center") delete centers[c];
        centers[c] = new CustVector<vector_type>("k_means_center", std::vector<vector_type>(newc.begin() + (size_t)c * D, newc.begin() + (size_t)(c + 1) * D));
    }
    return true;
}

// update.hpp:90-142
template <typename vector_type>
bool pam_lloyds(std::vector<CustVector<vector_type> >& input_vectors, std::vector<CustVector<vector_type>*>& centroids, std::string metric_type) {
    int K = (int)centroids.size();
    crx::Packed<vector_type> P;
    P.from_vector(input_vectors);
    std::vector<int32_t> labels(input_vectors.size()), rows(K), out(K);
    for (size_t i = 0; i < input_vectors.size(); i++) labels[i] = input_vectors[i].getCluster();
    for (int c = 0; c < K; c++) rows[c] = crx::row_of(input_vectors, centroids[c]);
    int swapped = 0;
    crx::check(crx_pam_lloyds(crx::context(), P.pts, labels.data(), CRX_HOST, rows.data(), K, crx::metric_code(metric_type), out.data(), &swapped), "crx_pam_lloyds");
    for (int c = 0; c < K; c++)
        if (out[c] >= 0 && out[c] != rows[c]) centroids[c] = &input_vectors[out[c]];
    return swapped != 0;
}

#endif  // CLUSTER_UPDATE_H
