// Drop-in for the hot-path helpers of lib/utils.hpp / utils.cpp (mod :97-98, get_num_hamming_dist_from
// utils.cpp:22-50, remove_clustering :143-147, separate_clusters_from_input :150-158,
// find_min_vector_distance :161-178).  String / file helpers of the same header are host I/O and out of scope.
#ifndef LIB_UTILS_H
#define LIB_UTILS_H

#include <string>
#include <vector>

#include "./data_structures/cust_vector.hpp"

// utils.hpp:97-98
template <typename x_type, typename n_type>
int mod(x_type x, n_type n) { return (x % n + n) % n; }

// utils.cpp:22-50
inline std::vector<int> get_num_hamming_dist_from(int num, int dist, int min_bit, int bits) {
    std::vector<int32_t> buf(1 << 16);
    int n = crx_get_num_hamming_dist_from(num, dist, min_bit, bits, buf.data(), (int)buf.size());
    if (n > (int)buf.size()) { buf.resize(n); crx_get_num_hamming_dist_from(num, dist, min_bit, bits, buf.data(), n); }
    return std::vector<int>(buf.begin(), buf.begin() + n);
}

template <typename dim_type>
void remove_clustering(std::vector<CustVector<dim_type> >& in_vectors) {
    for (size_t i = 0; i < in_vectors.size(); i++) in_vectors[i].resetCluster();
}

template <typename dim_type>
std::vector<std::vector<CustVector<dim_type>*> > separate_clusters_from_input(std::vector<CustVector<dim_type> >& in_vectors, int cluster_num) {
    std::vector<std::vector<CustVector<dim_type>*> > clusters(cluster_num);
    for (size_t i = 0; i < in_vectors.size(); i++) clusters[in_vectors[i].getCluster()].emplace_back(&in_vectors[i]);
    return clusters;
}

template <typename vector_type>
double find_min_vector_distance(std::vector<CustVector<vector_type>*>& vectors, std::string metric_type) {
    int K = (int)vectors.size();
    if (K < 2) return -1;
    crx::Packed<vector_type> P;
    P.from_pointers(vectors);
    std::vector<int32_t> a, b;
    for (int i = 0; i < K; i++)
        for (int j = i + 1; j < K; j++) { a.push_back(i); b.push_back(j); }
    std::vector<double> d(a.size());
    crx::check(crx_pair_op(crx::context(), P.pts, a.data(), P.pts, b.data(), (int64_t)a.size(), crx::metric_code(metric_type) == CRX_EUCLIDEAN ? 1 : 2, d.data()), "crx_pair_op");
    double m = -1;
    for (double v : d)
        if (m == -1 || v < m) m = v;
    return m;
}

#endif  // LIB_UTILS_H
