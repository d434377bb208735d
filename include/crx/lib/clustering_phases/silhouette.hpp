// Drop-in for lib/clustering_phases/silhouette.hpp (reference silhouette.hpp:19-144).
#ifndef CLUSTER_SILHOUETTE_H
#define CLUSTER_SILHOUETTE_H

#include <string>
#include <vector>

#include "../data_structures/cust_vector.hpp"

// silhouette.hpp:32-81: K per-cluster means followed by the overall mean
template <typename vector_type>
std::vector<double> silhouette_cluster(std::vector<std::vector<CustVector<vector_type>*> > clusters, std::vector<CustVector<vector_type>*>& centroids,
                                       std::string metric_type) {
    int K = (int)clusters.size();
    std::vector<CustVector<vector_type>*> all;
    std::vector<int32_t> labels;
    for (int c = 0; c < K; c++)
        for (auto v : clusters[c]) { all.push_back(v); labels.push_back(c); }
    crx::Packed<vector_type> P;
    P.from_pointers(all);
    int D = K ? (int)centroids[0]->getDimNumber() : 0;
    std::vector<double> C((size_t)K * D), sils((size_t)K + 1);
    for (int c = 0; c < K; c++) {
        const std::vector<vector_type>& d = centroids[c]->crxDimsRef();
        for (int j = 0; j < D; j++) C[(size_t)c * D + j] = (double)d[j];
    }
    crx::check(crx_silhouette_cluster(crx::context(), P.pts, labels.data(), CRX_HOST, C.data(), CRX_HOST, K, crx::metric_code(metric_type), sils.data()),
               "crx_silhouette_cluster");
    return sils;
}

#endif  // CLUSTER_SILHOUETTE_H
