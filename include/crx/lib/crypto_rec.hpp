// Drop-in for the recommendation core of lib/crypto_rec.hpp (reference crypto_rec.hpp:214-345).  The data
// preparation (:79-210) and 10-fold validation helpers (:349-449) of the same header are out of scope.
#ifndef CRYPTO_REC_HPP
#define CRYPTO_REC_HPP

#include <string>
#include <vector>

#include "./data_structures/cust_vector.hpp"
#include "lsh_cube.hpp"

namespace crx {
// neighbours + user packed into one point set: rows 0..n-1 = neighbours, row n = the user
template <typename T>
void pack_users(std::vector<CustVector<T>*>& neighbors, CustVector<T>& user, Packed<T>& P) {
    std::vector<CustVector<T>*> all(neighbors.begin(), neighbors.end());
    all.push_back(&user);
    P.from_pointers(all, true);
}
}  // namespace crx

// crypto_rec.hpp:235-277 on (similarity, payload) pairs: literal Lomuto, descending, `>=` pivot
template <typename dim_type, typename type>
void parallel_quickSort(std::vector<dim_type>& sim, std::vector<type>& neighbors, int low, int high) {
    if (low >= high) return;
    int n = high - low + 1;
    std::vector<double> keys(sim.begin() + low, sim.begin() + high + 1);
    std::vector<int32_t> idx(n);
    for (int i = 0; i < n; i++) idx[i] = i;
    crx::check(crx_parallel_quickSort(crx::context(), keys.data(), idx.data(), n), "crx_parallel_quickSort");
    std::vector<dim_type> s2(n);
    std::vector<type> n2(n, neighbors[low]);
    for (int i = 0; i < n; i++) { s2[i] = sim[low + idx[i]]; n2[i] = neighbors[low + idx[i]]; }
    for (int i = 0; i < n; i++) { sim[low + i] = s2[i]; neighbors[low + i] = n2[i]; }
}

// crypto_rec.hpp:214-231: sorts and truncates `neighbors`, returns the parallel similarities
template <typename dim_type>
std::vector<double> get_P_closest(std::vector<CustVector<dim_type>*>& neighbors, CustVector<dim_type>& user, int P) {
    int64_t n = (int64_t)neighbors.size();
    std::vector<double> sims;
    if (n == 0) return sims;
    crx::Packed<dim_type> U;
    crx::pack_users(neighbors, user, U);
    std::vector<int32_t> rows(n);
    for (int64_t i = 0; i < n; i++) rows[i] = (int32_t)i;
    sims.resize((size_t)std::min<int64_t>(n, std::max(P, 0)));
    int64_t kept = 0;
    std::vector<double> buf((size_t)std::max<int64_t>(1, std::min<int64_t>(n, std::max(P, 0))));
    crx::check(crx_get_P_closest(crx::context(), U.pts, rows.data(), n, U.pts, n, P, buf.data(), &kept), "crx_get_P_closest");
    std::vector<CustVector<dim_type>*> sorted((size_t)n);
    for (int64_t i = 0; i < n; i++) sorted[i] = neighbors[rows[i]];
    if (n > P) {  // crypto_rec.hpp:225-228
        sorted.resize(P);
        neighbors = sorted;
        return std::vector<double>(buf.begin(), buf.begin() + kept);
    }
    // n <= P: the reference returns all n similarities; the ABI kept min(n, P) = n of them
    neighbors = sorted;
    return std::vector<double>(buf.begin(), buf.begin() + kept);
}

// crypto_rec.hpp:281-306
template <typename dim_type>
std::vector<dim_type> get_predicted_user_sim(std::vector<CustVector<dim_type>*>& neighbors, CustVector<dim_type>& user, std::vector<double> similarities) {
    crx::Packed<dim_type> U;
    crx::pack_users(neighbors, user, U);
    int64_t n = (int64_t)neighbors.size();
    std::vector<int32_t> rows(n);
    for (int64_t i = 0; i < n; i++) rows[i] = (int32_t)i;
    std::vector<double> pred(user.getDimNumber());
    crx::check(crx_get_top_N_recom(crx::context(), U.pts, rows.data(), similarities.data(), n, U.pts, n, 0, pred.data(), nullptr), "crx_get_top_N_recom");
    return std::vector<dim_type>(pred.begin(), pred.end());
}

// crypto_rec.hpp:310-324
template <typename dim_type>
std::vector<int> get_top_N_recom(std::vector<CustVector<dim_type>*>& neighbors, CustVector<dim_type>& user, int N, std::vector<double> similarities) {
    crx::Packed<dim_type> U;
    crx::pack_users(neighbors, user, U);
    int64_t n = (int64_t)neighbors.size();
    std::vector<int32_t> rows(n), recs(N);
    for (int64_t i = 0; i < n; i++) rows[i] = (int32_t)i;
    crx::check(crx_get_top_N_recom(crx::context(), U.pts, rows.data(), similarities.data(), n, U.pts, n, N, nullptr, recs.data()), "crx_get_top_N_recom");
    return std::vector<int>(recs.begin(), recs.end());
}

// crypto_rec.hpp:328-345: similarities to ALL neighbours are computed first, no top-P cut
template <typename dim_type>
std::vector<int> get_top_N_recom(std::vector<CustVector<dim_type>*>& neighbors, CustVector<dim_type>& user, int N) {
    crx::Packed<dim_type> U;
    crx::pack_users(neighbors, user, U);
    int64_t n = (int64_t)neighbors.size();
    std::vector<int32_t> rows(n), recs(N);
    for (int64_t i = 0; i < n; i++) rows[i] = (int32_t)i;
    crx::check(crx_get_top_N_recom(crx::context(), U.pts, rows.data(), nullptr, n, U.pts, n, N, nullptr, recs.data()), "crx_get_top_N_recom");
    return std::vector<int>(recs.begin(), recs.end());
}

#endif  // CRYPTO_REC_HPP
