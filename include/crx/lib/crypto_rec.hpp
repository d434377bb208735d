// Drop-in for lib/crypto_rec.hpp: the recommendation core (reference crypto_rec.hpp:214-345), the user-vector
// construction in front of it (:79-210, on the GPU through crx_user_vectors_build once the strings are resolved to
// indices; available when the reference's lib/data_structures/tweet.h is on the include path) and the 10-fold
// validation helpers behind it (:347-449, host-side container shuffling exactly as the reference does it).
#ifndef CRYPTO_REC_HPP
#define CRYPTO_REC_HPP

#include <cstdlib>
#include <cstring>
#include <ctime>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#if defined(__has_include)
#if __has_include("./data_structures/tweet.h")
#include "./data_structures/tweet.h"
#define CRX_HAVE_TWEET 1
#endif
#endif

#include "./data_structures/cust_vector.hpp"
#include "lsh_cube.hpp"

namespace crx {
// Device copies the per-user calls would otherwise upload again and again: the user as a one-row set with its rating
// metadata (get_P_closest and get_top_N_recom are called one after the other for the same user), and a neighbour list
// that is not part of a registered vector (the members of one cluster, main.cpp:367-372, passed once per user).  A copy
// is reused only for the same objects and while crx::content_epoch() has not moved since it was packed, i.e. no
// CustVector has been created, destroyed or written in between.
template <typename T>
struct DeviceCopies {
    struct Entry {
        std::vector<const void*> who;
        unsigned long epoch = 0, stamp = 0;
        Packed<T>* set = nullptr;
    };
    std::vector<Entry> entries;
    unsigned long clock_ = 0;
    explicit DeviceCopies(size_t slots) : entries(slots) {}
    crx_points* get(CustVector<T>* const* ptrs, size_t n) {
        Entry* victim = &entries[0];
        for (Entry& e : entries) {
            if (e.set && e.who.size() == n && std::memcmp(e.who.data(), ptrs, n * sizeof(void*)) == 0) {
                if (e.epoch == content_epoch()) { e.stamp = ++clock_; return e.set->pts; }
                victim = &e;   // same objects, possibly new content: packed again below
                break;
            }
            if (e.stamp < victim->stamp) victim = &e;
        }
        delete victim->set;
        victim->set = new Packed<T>();
        victim->set->build((int64_t)n, [&](int64_t i) { return ptrs[i]; }, true);
        victim->who.assign((const void* const*)ptrs, (const void* const*)ptrs + n);
        victim->epoch = content_epoch();
        victim->stamp = ++clock_;
        return victim->set->pts;
    }
};
template <typename T>
inline DeviceCopies<T>& user_copies() { static DeviceCopies<T> c(8); return c; }
template <typename T>
inline DeviceCopies<T>& list_copies() { static DeviceCopies<T> c(32); return c; }

// The neighbours and the user as rows of device point sets.  The neighbours point into a vector a table set has
// registered (create_LSH_hashtables) -- nothing but row numbers travels -- or are a list kept by list_copies(); the
// user is a row of the same registered set or a one-row set of its own.
template <typename T>
struct Resolved {
    crx_points* users = nullptr;
    crx_points* query_set = nullptr;
    int64_t query_row = 0;
    bool packed = false;      // rows[] are positions in `neighbors`, not rows of a registered vector
    std::vector<int32_t> rows;
    Resolved(std::vector<CustVector<T>*>& neighbors, CustVector<T>& user) {
        std::vector<CustVector<T>*> one(1, &user);
        users = registered_rows(neighbors, rows);
        if (users) {
            std::vector<int32_t> ur;
            if (registered_rows(one, ur) == users) { query_set = users; query_row = ur[0]; return; }
        } else {
            packed = true;
            if ((size_t)neighbors.size() * user.getDimNumber() <= ((size_t)1 << 22)) {
                users = list_copies<T>().get(neighbors.data(), neighbors.size());
            } else {   // too large to keep around: packed for this call only
                big.from_pointers(neighbors, true);
                users = big.pts;
            }
            rows.resize(neighbors.size());
            for (size_t i = 0; i < rows.size(); i++) rows[i] = (int32_t)i;
        }
        query_set = user_copies<T>().get(one.data(), 1);
        query_row = 0;
    }
private:
    Packed<T> big;
};
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

namespace crx {
// The batch (TableSet::Batch) that can answer a call about `user` and the row of the user in it: the user is a stored row
// of the tables the neighbours point into (own), or a row of another vector with a current device copy (ext).
template <typename T>
inline typename TableSet<T>::Batch* batch_of(std::vector<CustVector<T>*>& neighbors, CustVector<T>& user, TableSet<T>*& set, int64_t& row, int64_t& nq) {
    if (neighbors.empty()) return nullptr;
    const char* p0 = (const char*)neighbors[0];
    set = nullptr;
    for (const Registered& reg : registry())
        if (reg.table_set && !reg.dirty && p0 >= reg.begin && p0 < reg.end && reg.stride == sizeof(CustVector<T>)) { set = (TableSet<T>*)reg.table_set; break; }
    if (!set || !set->base || set->base->empty()) return nullptr;
    const char* q = (const char*)&user;
    const char* b0 = (const char*)&(*set->base)[0];
    if (q >= b0 && q < b0 + set->base->size() * sizeof(CustVector<T>)) {
        row = (int64_t)((q - b0) / (ptrdiff_t)sizeof(CustVector<T>));
        nq = set->n();
        return &set->own;
    }
    const Registered* home = home_of(&user, sizeof(CustVector<T>));
    if (home && set->ext.queries == home->pts && set->ext.epoch == home->gen) {
        row = (int64_t)((q - home->begin) / (ptrdiff_t)home->stride);
        nq = home->n;
        return &set->ext;
    }
    return nullptr;
}
}  // namespace crx

// crypto_rec.hpp:214-231: sorts and truncates `neighbors`, returns the parallel similarities
template <typename dim_type>
std::vector<double> get_P_closest(std::vector<CustVector<dim_type>*>& neighbors, CustVector<dim_type>& user, int P) {
    int64_t n = (int64_t)neighbors.size();
    if (n == 0) return std::vector<double>();
    crx::Timed timed("get_P_closest");
    {
        // the candidate list get_LSH_filtered_combined_buckets has just returned for this user: the batched call over all
        // users holds the answer (see TableSet::Batch)
        crx::TableSet<dim_type>* set = nullptr;
        int64_t row = -1, nq = 0;
        typename crx::TableSet<dim_type>::Batch* bt = crx::batch_of(neighbors, user, set, row, nq);
        if (bt && bt->last_row == row && (int64_t)bt->last_list.size() == n) {
            CustVector<dim_type>* base = &(*set->base)[0];
            bool same = true;
            for (int64_t i = 0; i < n && same; i++) same = neighbors[i] == base + bt->last_list[i];
            if (same && set->batch_for(*bt, P, nq) && bt->ncand[row] == n) {
                int64_t kept = std::min<int64_t>(n, P);
                std::vector<CustVector<dim_type>*> sorted((size_t)kept);
                for (int64_t i = 0; i < kept; i++) sorted[i] = base + bt->rows[(size_t)row * P + i];
                neighbors.swap(sorted);
                return std::vector<double>(bt->sims.begin() + (size_t)row * P, bt->sims.begin() + (size_t)row * P + kept);
            }
        }
    }
    crx::Resolved<dim_type> R(neighbors, user);
    // position -> neighbour, carried through the co-sort as the payload
    std::vector<int32_t> order(R.rows);
    int64_t kept = 0;
    std::vector<double> buf((size_t)std::max<int64_t>(1, std::min<int64_t>(n, std::max(P, 0))));
    crx::check(crx_get_P_closest(crx::context(), R.users, order.data(), n, R.query_set, R.query_row, P, buf.data(), &kept), "crx_get_P_closest");
    // rows -> pointers: rows are unique inside one call (a neighbour list holds every vector once)
    std::vector<CustVector<dim_type>*> sorted((size_t)kept);
    if (R.packed) {
        for (int64_t i = 0; i < kept; i++) sorted[i] = neighbors[order[i]];
    } else {
        const char* base = nullptr;
        for (const crx::Registered& reg : crx::registry()) if (reg.pts == R.users) base = reg.begin;
        for (int64_t i = 0; i < kept; i++) sorted[i] = (CustVector<dim_type>*)(const_cast<char*>(base) + (size_t)order[i] * sizeof(CustVector<dim_type>));
    }
    if (n > P) {  // crypto_rec.hpp:225-228
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
    crx::Resolved<dim_type> R(neighbors, user);
    int64_t n = (int64_t)neighbors.size();
    std::vector<double> pred(user.getDimNumber());
    crx::check(crx_get_top_N_recom(crx::context(), R.users, R.rows.data(), similarities.data(), n, R.query_set, R.query_row, 0, pred.data(), nullptr), "crx_get_top_N_recom");
    return std::vector<dim_type>(pred.begin(), pred.end());
}

// crypto_rec.hpp:310-324
template <typename dim_type>
std::vector<int> get_top_N_recom(std::vector<CustVector<dim_type>*>& neighbors, CustVector<dim_type>& user, int N, std::vector<double> similarities) {
    crx::Timed timed("get_top_N_recom (similarities)");
    {
        // neighbours and similarities are the ones the batched call produced for this user: its coin order is the literal sort
        // of the same predicted ratings, so the first N entries are this call's answer
        crx::TableSet<dim_type>* set = nullptr;
        int64_t row = -1, nq = 0;
        typename crx::TableSet<dim_type>::Batch* bt = crx::batch_of(neighbors, user, set, row, nq);
        int64_t n = (int64_t)neighbors.size();
        if (bt && bt->state == 1 && N >= 0 && N <= bt->nrec && n > 0 && n <= bt->P && (int64_t)similarities.size() == n &&
            n == std::min<int64_t>(bt->ncand[row], bt->P)) {
            int P = bt->P;
            CustVector<dim_type>* base = &(*set->base)[0];
            bool same = std::memcmp(similarities.data(), &bt->sims[(size_t)row * P], (size_t)n * sizeof(double)) == 0;
            for (int64_t i = 0; i < n && same; i++) same = neighbors[i] == base + bt->rows[(size_t)row * P + i];
            if (same) return std::vector<int>(bt->recs.begin() + (size_t)row * bt->nrec, bt->recs.begin() + (size_t)row * bt->nrec + N);
        }
    }
    crx::Resolved<dim_type> R(neighbors, user);
    int64_t n = (int64_t)neighbors.size();
    std::vector<int32_t> recs(N);
    crx::check(crx_get_top_N_recom(crx::context(), R.users, R.rows.data(), similarities.data(), n, R.query_set, R.query_row, N, nullptr, recs.data()), "crx_get_top_N_recom");
    return std::vector<int>(recs.begin(), recs.end());
}

namespace crx {
// The clustering recommendation (main.cpp:260-269) calls get_top_N_recom(members of the user's cluster, user, N) once per
// user: the same neighbour list again and again.  The first call for a list computes the recommendations of ALL its
// members in one engine call (crx_recommend_cluster); later calls for other members are answered from that result.
// A cached result is used only when the list's CONTENT (coordinates, unknown flags, means) is byte for byte what it was
// computed from -- a 64-bit hash finds the entry, a full comparison confirms it -- so it can never be stale; the
// function is pure, hence the answer is identical to recomputing it.
struct ClusterRecs {
    int64_t n = 0;
    int d = 0, N = 0;
    uint64_t hash = 0;
    std::vector<double> buf, mean;
    std::vector<uint8_t> unk;
    std::vector<int32_t> recs;   // [n][N]
    unsigned long stamp = 0;
    std::vector<const void*> who;   // the objects the image was last compared with ...
    unsigned long epoch = 0;        // ... and crx::content_epoch() at that moment
};
inline uint64_t image_hash(const void* p, size_t bytes, uint64_t h) {
    const unsigned char* b = (const unsigned char*)p;
    size_t i = 0;
    for (; i + 8 <= bytes; i += 8) { uint64_t w; std::memcpy(&w, b + i, 8); h = (h ^ w) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; }
    for (; i < bytes; i++) h = (h ^ b[i]) * 0x100000001B3ull;
    return h;
}
template <typename T>
inline bool cached_cluster_recs(std::vector<CustVector<T>*>& neighbors, CustVector<T>& user, int N, std::vector<int>& out) {
    int64_t n = (int64_t)neighbors.size();
    if (n < 2 || N <= 0 || (size_t)n * user.getDimNumber() > ((size_t)1 << 22)) return false;
    int64_t me = -1;
    for (int64_t i = 0; i < n; i++) if (neighbors[i] == &user) { me = i; break; }
    if (me < 0) return false;
    static std::vector<ClusterRecs> cache(64);
    static unsigned long clock_ = 0;
    // the same objects as at the last full comparison and no CustVector has been created, destroyed or written since
    // (crx::content_epoch): the image is what it was -- no need to pack and compare it again
    for (ClusterRecs& e : cache) {
        if (e.n != n || e.N != N || e.epoch != content_epoch() || e.who.size() != (size_t)n) continue;
        if (std::memcmp(e.who.data(), neighbors.data(), (size_t)n * sizeof(void*)) != 0) continue;
        e.stamp = ++clock_;
        out.assign(e.recs.begin() + (size_t)me * N, e.recs.begin() + (size_t)(me + 1) * N);
        return true;
    }
    ClusterRecs probe;
    probe.n = n; probe.N = N;
    Packed<T>::pack_host(n, [&](int64_t i) { return neighbors[i]; }, true, probe.d, probe.buf, probe.unk, probe.mean);
    probe.hash = image_hash(probe.buf.data(), probe.buf.size() * 8, 1469598103934665603ull);
    probe.hash = image_hash(probe.unk.data(), probe.unk.size(), probe.hash);
    probe.hash = image_hash(probe.mean.data(), probe.mean.size() * 8, probe.hash);
    ClusterRecs* victim = &cache[0];
    for (ClusterRecs& e : cache) {
        if (e.n == n && e.N == N && e.d == probe.d && e.hash == probe.hash && e.buf == probe.buf && e.unk == probe.unk && e.mean == probe.mean) {
            e.stamp = ++clock_;
            e.who.assign(neighbors.begin(), neighbors.end()); e.epoch = content_epoch();
            out.assign(e.recs.begin() + (size_t)me * N, e.recs.begin() + (size_t)(me + 1) * N);
            return true;
        }
        if (e.stamp < victim->stamp) victim = &e;
    }
    Packed<T> set;
    set.upload(n, probe.d, probe.buf, probe.unk, probe.mean);
    std::vector<int32_t> labels((size_t)n, 0);
    probe.recs.assign((size_t)n * N, 0);
    check(crx_recommend_cluster(context(), set.pts, labels.data(), CRX_HOST, 1, nullptr, nullptr, N, probe.recs.data(), CRX_HOST), "crx_recommend_cluster");
    probe.stamp = ++clock_;
    probe.who.assign(neighbors.begin(), neighbors.end()); probe.epoch = content_epoch();
    out.assign(probe.recs.begin() + (size_t)me * N, probe.recs.begin() + (size_t)(me + 1) * N);
    *victim = std::move(probe);
    return true;
}
}  // namespace crx

namespace crx {
// main.cpp:367-372: every user of a vector gets the recommendations of ONE of a few explicit neighbour lists (the members
// of its nearest cluster).  When the user's vector has a current device copy (home_of), the second call with a given list
// computes that list's recommendations for ALL rows of the vector in one engine call (crx_recommend_cluster with the
// vector as the query set); later calls with the same list objects are answered from it while content_epoch() stands.
template <typename T>
inline bool home_list_recs(std::vector<CustVector<T>*>& neighbors, CustVector<T>& user, int N, std::vector<int>& out) {
    size_t n = neighbors.size();
    if (n < 1 || N <= 0 || N > 128 || n * user.getDimNumber() > ((size_t)1 << 22)) return false;
    const Registered* home = home_of(&user, sizeof(CustVector<T>));
    if (!home || (double)home->n * N > 64e6 || crx_points_d(home->pts) != (int32_t)user.getDimNumber()) return false;
    int64_t hrow = (int64_t)(((const char*)&user - home->begin) / (ptrdiff_t)home->stride);
    struct Entry {
        std::vector<const void*> who;
        unsigned long epoch = 0, stamp = 0;
        const crx_points* home = nullptr;
        int N = 0, calls = 0;
        std::vector<int32_t> recs;   // [home rows][N]
    };
    static std::vector<Entry> cache(32);
    static unsigned long clock_ = 0;
    Entry* victim = &cache[0];
    Entry* hit = nullptr;
    for (Entry& e : cache) {
        if (e.home == home->pts && e.epoch == content_epoch() && e.N == N && e.who.size() == n &&
            std::memcmp(e.who.data(), neighbors.data(), n * sizeof(void*)) == 0) { hit = &e; break; }
        if (e.stamp < victim->stamp) victim = &e;
    }
    if (!hit) {
        *victim = Entry();
        victim->who.assign(neighbors.begin(), neighbors.end());
        victim->epoch = content_epoch(); victim->home = home->pts; victim->N = N; victim->calls = 1; victim->stamp = ++clock_;
        return false;
    }
    hit->stamp = ++clock_;
    if (hit->recs.empty()) {
        if (++hit->calls < 2) return false;
        Timed timed("crx_recommend_cluster, whole vector x list");
        crx_points* users = list_copies<T>().get(neighbors.data(), n);
        std::vector<int32_t> labels(n, 0), qlabels((size_t)home->n, 0);
        hit->recs.assign((size_t)home->n * N, 0);
        check(crx_recommend_cluster(context(), users, labels.data(), CRX_HOST, 1, home->pts, qlabels.data(), N, hit->recs.data(), CRX_HOST), "crx_recommend_cluster");
    }
    out.assign(hit->recs.begin() + (size_t)hrow * N, hit->recs.begin() + (size_t)(hrow + 1) * N);
    return true;
}
}  // namespace crx

// crypto_rec.hpp:328-345: similarities to ALL neighbours are computed first, no top-P cut
template <typename dim_type>
std::vector<int> get_top_N_recom(std::vector<CustVector<dim_type>*>& neighbors, CustVector<dim_type>& user, int N) {
    crx::Timed timed("get_top_N_recom (no similarities)");
    {
        std::vector<int> cached;
        if (crx::cached_cluster_recs(neighbors, user, N, cached)) return cached;
        if (crx::home_list_recs(neighbors, user, N, cached)) return cached;
    }
    crx::Resolved<dim_type> R(neighbors, user);
    int64_t n = (int64_t)neighbors.size();
    std::vector<int32_t> recs(N);
    crx::check(crx_get_top_N_recom(crx::context(), R.users, R.rows.data(), nullptr, n, R.query_set, R.query_row, N, nullptr, recs.data()), "crx_get_top_N_recom");
    return std::vector<int>(recs.begin(), recs.end());
}

// ------------------------------------------------------------------------------------------------
// user vectors (crypto_rec.hpp:79-210)
// ------------------------------------------------------------------------------------------------
namespace crx {
// mentions -> one CustVector per kept user; ids[u] names user u
template <typename dim_type>
std::vector<CustVector<dim_type> > build_user_vectors(const std::vector<int32_t>& m_user, const std::vector<int32_t>& m_coin,
                                                      const std::vector<double>& m_score, const std::vector<std::string>& ids,
                                                      int crypto_num) {
    int64_t n_users = (int64_t)ids.size();
    std::vector<double> X((size_t)n_users * crypto_num), mean(n_users);
    std::vector<uint8_t> unk((size_t)n_users * crypto_num), keep(n_users);
    std::vector<CustVector<dim_type> > out;
    if (n_users == 0 || crypto_num <= 0) return out;
    check(crx_user_vectors_build(context(), m_user.data(), m_coin.data(), m_score.data(), (int64_t)m_user.size(), n_users, crypto_num,
                                 X.data(), unk.data(), mean.data(), keep.data(), CRX_HOST), "crx_user_vectors_build");
    for (int64_t u = 0; u < n_users; u++) {
        if (!keep[u]) continue;  // "useless": nothing but zeros (:127 / :196)
        std::vector<dim_type> dims(X.begin() + (size_t)u * crypto_num, X.begin() + (size_t)(u + 1) * crypto_num);
        std::set<int> unknown;
        for (int j = 0; j < crypto_num; j++) if (unk[(size_t)u * crypto_num + j]) unknown.insert(unknown.end(), j);
        out.emplace_back(CustVector<dim_type>(ids[u], dims, unknown, mean[u]));
    }
    return out;
}
}  // namespace crx

#ifdef CRX_HAVE_TWEET
// crypto_rec.hpp:79-140.  Users come out in the iteration order of an unordered_map keyed by user id that saw
// the same insertions as the reference's user_map, i.e. in the reference's own order.
template <typename dim_type>
std::vector<CustVector<dim_type> > tweets_to_user_vectors(std::unordered_map<std::string, Tweet>& tweets, int crypto_num) {
    std::unordered_map<std::string, int32_t> first_seen;  // user id -> provisional index
    std::vector<int32_t> m_user, m_coin;
    std::vector<double> m_score;
    for (auto& kv : tweets) {
        std::string uid = kv.second.getUserId();
        auto it = first_seen.find(uid);
        if (it == first_seen.end()) it = first_seen.emplace(uid, (int32_t)first_seen.size()).first;
        double score = kv.second.getSentimentScore();
        for (int coin : kv.second.getCryptoIndexes()) { m_user.push_back(it->second); m_coin.push_back(coin); m_score.push_back(score); }
    }
    // renumber in map iteration order = output order
    std::vector<int32_t> order(first_seen.size());
    std::vector<std::string> ids;
    ids.reserve(first_seen.size());
    for (auto& kv : first_seen) { order[kv.second] = (int32_t)ids.size(); ids.push_back(kv.first); }
    for (auto& u : m_user) u = order[u];
    return crx::build_user_vectors<dim_type>(m_user, m_coin, m_score, ids, crypto_num);
}

// crypto_rec.hpp:143-210: one "user" per cluster of tweet vectors, id = the cluster number
template <typename dim_type>
std::vector<CustVector<dim_type> > clusters_to_user_vectors(std::unordered_map<std::string, Tweet>& tweets,
                                                            std::vector<CustVector<dim_type> >& vectors, int crypto_num, int user_num) {
    std::vector<int32_t> m_user, m_coin;
    std::vector<double> m_score;
    for (auto& vec : vectors) {
        auto it = tweets.find(vec.getId());
        if (it == tweets.end()) continue;
        double score = it->second.getSentimentScore();
        for (int coin : it->second.getCryptoIndexes()) { m_user.push_back(vec.getCluster()); m_coin.push_back(coin); m_score.push_back(score); }
    }
    std::vector<std::string> ids(user_num > 0 ? user_num : 0);
    for (int u = 0; u < user_num; u++) ids[u] = std::to_string(u);
    return crx::build_user_vectors<dim_type>(m_user, m_coin, m_score, ids, crypto_num);
}
#endif  // CRX_HAVE_TWEET

// ------------------------------------------------------------------------------------------------
// 10-fold validation helpers (crypto_rec.hpp:347-449): container shuffling driven by srand(time)/rand()
// ------------------------------------------------------------------------------------------------
template <typename dim_type>
std::vector<std::vector<CustVector<dim_type> > > split_to_10(std::vector<CustVector<dim_type> > input_vectors) {
    srand((int)time(0));
    size_t per_fold = input_vectors.size() / 10;
    std::vector<std::vector<CustVector<dim_type> > > folds(10);
    for (auto& fold : folds) {
        fold.reserve(per_fold);
        while (fold.size() < per_fold && !input_vectors.empty()) {
            size_t pick = (size_t)rand() % input_vectors.size();
            fold.emplace_back(input_vectors[pick]);
            input_vectors.erase(input_vectors.begin() + pick);
        }
    }
    return folds;
}

template <typename dim_type>
std::vector<CustVector<dim_type> > merge_except_for(std::vector<std::vector<CustVector<dim_type> > > vectors_to_merge, int not_merge_index) {
    std::vector<CustVector<dim_type> > merged;
    for (int i = 0; i < (int)vectors_to_merge.size(); i++)
        if (i != not_merge_index) merged.insert(merged.end(), vectors_to_merge[i].begin(), vectors_to_merge[i].end());
    return merged;
}

// crypto_rec.hpp:393-449.  Kept quirk: the random position inside the list of known coins is used directly as a
// coordinate index (:411-412), and every coordinate except that one enters the new mean (:421-430).
template <typename dim_type>
bool hide_one_score(CustVector<dim_type>& inVector, double* old_score) {
    std::vector<dim_type>& dims = *inVector.getDimensions();
    std::set<int> unknown = inVector.getUnknownIndexesSet();
    size_t known = dims.size() - unknown.size();
    if (known < 2) return false;
    srand((int)time(0));
    int hide = rand() % (int)known;
    *old_score = dims[hide];
    for (int i : unknown) dims[i] = 0;
    double total = 0;
    int counted = 0;
    bool all_zero = true;
    for (int i = 0; i < (int)dims.size(); i++) {
        if (i == hide) continue;
        total = total + dims[i];
        counted++;
        if (dims[i] != 0) all_zero = false;
    }
    if (all_zero) return false;
    double new_mean = total / counted;
    dims[hide] = new_mean;
    inVector.setKnownMean(new_mean);
    inVector.setUnknownIndexes(std::set<int>{hide});
    return true;
}

#endif  // CRYPTO_REC_HPP
