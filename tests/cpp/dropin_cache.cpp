// dropin_cache.cpp -- the drop-in headers remember results behind the reference's per-user signatures:
// the batched answer of the stored-user loop, the fan-out of one user's distances over recurring centroids, device copies
// of users and explicit neighbour lists.  Every remembered answer must be the one a cold call gives, and a change of an
// input (through a setter, getDimensions(), a new object at an old address) must be seen.  Prints "OK" or the first mismatch.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "lib/data_structures/cust_vector.hpp"
#include "lib/data_structures/cust_hashtable.hpp"
#include "lib/lsh_cube.hpp"
#include "lib/crypto_rec.hpp"

typedef CustVector<double> Vec;

static int g_bad = 0;
#define EXPECT(cond, ...) do { if (!(cond)) { if (!g_bad) { std::printf("MISMATCH %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); } g_bad++; } } while (0)

// rating-like users: a few known coins with small integer scores, the rest = mean (ties and equal similarities abound)
static std::vector<Vec> make_users(int n, int d, unsigned seed) {
    std::mt19937 g(seed);
    std::vector<Vec> out;
    for (int u = 0; u < n; u++) {
        std::vector<double> x(d, 0.0);
        std::set<int> unknown;
        int known = 1 + (int)(g() % 6);
        std::set<int> kn;
        while ((int)kn.size() < known) kn.insert((int)(g() % d));
        double total = 0;
        for (int c : kn) { x[c] = (double)((int)(g() % 9) - 3) * 0.5 + 0.25 * (double)(g() % 3); total += x[c]; }
        double mean = total / known;
        bool all_zero = true;
        for (int c : kn) if (x[c] != 0) all_zero = false;
        if (all_zero) { x[*kn.begin()] = 1.0; mean = 1.0 / known; }
        for (int c = 0; c < d; c++) if (!kn.count(c)) { x[c] = mean; unknown.insert(c); }
        out.emplace_back(Vec(std::to_string(u), x, unknown, mean));
    }
    return out;
}

// a copy of every argument at fresh addresses: nothing the headers remembered can apply to it
struct Cold {
    std::vector<Vec> store;
    std::vector<Vec*> ptrs;
    Vec user;
    Cold(const std::vector<Vec*>& neighbors, Vec& u) : user(u) {
        store.reserve(neighbors.size());
        for (Vec* v : neighbors) store.emplace_back(*v);
        for (Vec& v : store) ptrs.push_back(&v);
    }
};

int main() {
    crx::set_seed(12345);
    const int D = 40;

    // ---- 1. one user against recurring centroids (main.cpp:353-366) ----
    {
        std::vector<Vec> users = make_users(6, D, 1), cents = make_users(11, D, 2);
        std::vector<std::vector<double> > first(users.size(), std::vector<double>(cents.size()));
        for (size_t u = 0; u < users.size(); u++)
            for (size_t c = 0; c < cents.size(); c++) first[u][c] = users[u].euclideanDistance(&cents[c]);   // user 0: single calls; later users: fan-out
        for (size_t u = 0; u < users.size(); u++)
            for (size_t c = 0; c < cents.size(); c++) {
                Vec a(users[u]), b(cents[c]);   // fresh addresses: cold single-pair call
                double cold = a.euclideanDistance(&b);
                EXPECT(cold == first[u][c], "fan-out distance user %zu centroid %zu: %.17g vs cold %.17g", u, c, first[u][c], cold);
                EXPECT(users[u].cosineSimilarity(&cents[c]) == a.cosineSimilarity(&b), "cosine similarity user %zu centroid %zu", u, c);
            }
        // a centroid changes through getDimensions(): the next loop must see it
        (*cents[3].getDimensions())[5] += 2.5;
        for (size_t u = 0; u < users.size(); u++)
            for (size_t c = 0; c < cents.size(); c++) {
                double warm = users[u].euclideanDistance(&cents[c]);
                Vec a(users[u]), b(cents[c]);
                EXPECT(warm == a.euclideanDistance(&b), "after an edit: user %zu centroid %zu", u, c);
                if (c == 3) EXPECT(warm != first[u][c], "edit of centroid 3 not seen by user %zu", u);
            }
        // a user changes between two of its own calls
        double before = users[2].euclideanDistance(&cents[0]);
        (*users[2].getDimensions())[0] += 1.0;
        double after = users[2].euclideanDistance(&cents[0]);
        Vec a(users[2]), b(cents[0]);
        EXPECT(after == a.euclideanDistance(&b) && after != before, "edit of the left operand not seen");
    }

    // ---- 2. explicit neighbour lists (main.cpp:260-269, 367-372) ----
    {
        std::vector<Vec> members = make_users(60, D, 3), outsiders = make_users(5, D, 4);
        std::vector<Vec*> list;
        for (Vec& v : members) list.push_back(&v);
        for (int round = 0; round < 2; round++) {
            for (Vec& u : outsiders) {   // not a member: per-user call against the remembered device copy of the list
                std::vector<int> warm = get_top_N_recom(list, u, 5);
                Cold cold(list, u);
                EXPECT(warm == get_top_N_recom(cold.ptrs, cold.user, 5), "outsider %s round %d", u.getId().c_str(), round);
            }
            for (Vec& u : members) {     // a member: answered from one call for the whole list
                std::vector<int> warm = get_top_N_recom(list, u, 5);
                Cold cold(list, u);
                std::vector<Vec*> one_by_one(cold.ptrs);
                EXPECT(warm == get_top_N_recom(one_by_one, cold.user, 5), "member %s round %d", u.getId().c_str(), round);
            }
            // between the rounds one neighbour's mean and one coordinate change
            members[7].setKnownMean(members[7].getKnownMean() + 0.75);
            (*members[11].getDimensions())[3] -= 1.25;
        }
    }

    // ---- 3. the stored-user loop (main.cpp:159-170): batched answer == per-user answer ----
    {
        const int P = 20;
        std::vector<Vec> users = make_users(700, D, 5);
        std::vector<CustHashtable<double>*> tables = create_LSH_hashtables<double>(users, "cosine", 3, 4, 100, 0.4);
        struct Ans { std::vector<std::string> ids; std::vector<double> sims; std::vector<int> recs; };
        std::vector<Ans> per_user(users.size()), batched(users.size());
        // per-user path: the list is passed as a copy at the time another row was the last combined-bucket query
        for (size_t u = 0; u < users.size(); u++) {
            std::vector<Vec*> nb = get_LSH_filtered_combined_buckets(tables, &users[u]);
            get_LSH_filtered_combined_buckets(tables, &users[(u + 1) % users.size()]);
            if (nb.empty()) continue;
            per_user[u].sims = get_P_closest(nb, users[u], P);
            for (Vec* v : nb) per_user[u].ids.push_back(v->getId());
            per_user[u].recs = get_top_N_recom(nb, users[u], 2, per_user[u].sims);
        }
        for (size_t u = 0; u < users.size(); u++) {   // main.cpp's own order of calls: the second user triggers the batch
            std::vector<Vec*> nb = get_LSH_filtered_combined_buckets(tables, &users[u]);
            if (nb.empty()) continue;
            batched[u].sims = get_P_closest(nb, users[u], P);
            for (Vec* v : nb) batched[u].ids.push_back(v->getId());
            batched[u].recs = get_top_N_recom(nb, users[u], 2, batched[u].sims);
        }
        EXPECT(tables[0]->set->own.state == 1, "the batched call did not run (state %d)", tables[0]->set->own.state);
        for (size_t u = 0; u < users.size(); u++) {
            EXPECT(per_user[u].ids == batched[u].ids, "neighbours of user %zu", u);
            EXPECT(per_user[u].sims == batched[u].sims, "similarities of user %zu", u);
            EXPECT(per_user[u].recs == batched[u].recs, "coins of user %zu", u);
        }
        // a user that is not a table row, twice in a row and after an edit (main.cpp:205-216)
        std::vector<Vec> others = make_users(8, D, 6);
        for (int round = 0; round < 2; round++) {
            for (Vec& u : others) {
                std::vector<Vec*> nb = get_LSH_filtered_combined_buckets(tables, &u);
                if (nb.empty()) continue;
                std::vector<Vec*> nb2(nb);
                Vec copy(u);
                std::vector<double> s1 = get_P_closest(nb, u, P), s2 = get_P_closest(nb2, copy, P);
                EXPECT(s1 == s2 && nb == nb2, "external user %s round %d", u.getId().c_str(), round);
                EXPECT(get_top_N_recom(nb, u, 2, s1) == get_top_N_recom(nb2, copy, 2, s2), "coins of external user %s round %d", u.getId().c_str(), round);
            }
            others[1].setKnownMean(others[1].getKnownMean() - 0.5);
            (*others[2].getDimensions())[1] += 3.0;
        }
        // a stored vector is edited while its tables live: the bucket placement stays (as in the reference), the values are live
        users[3].setKnownMean(users[3].getKnownMean() + 1.0);
        (*users[10].getDimensions())[7] += 0.5;
        {
            std::vector<Vec> copies(users);   // same contents at fresh addresses; candidates by row number
            for (size_t u : {(size_t)3, (size_t)10, (size_t)11, (size_t)400}) {
                std::vector<Vec*> nb = get_LSH_filtered_combined_buckets(tables, &users[u]);
                if (nb.empty()) continue;
                std::vector<Vec*> nbc;
                for (Vec* v : nb) nbc.push_back(&copies[(size_t)(v - &users[0])]);
                std::vector<double> s1 = get_P_closest(nb, users[u], P), s2 = get_P_closest(nbc, copies[u], P);
                EXPECT(s1 == s2, "similarities of stored user %zu after an edit of stored vectors", u);
                bool same = nb.size() == nbc.size();
                for (size_t i = 0; same && i < nb.size(); i++) same = (nb[i] - &users[0]) == (nbc[i] - &copies[0]);
                EXPECT(same, "neighbours of stored user %zu after an edit of stored vectors", u);
                EXPECT(get_top_N_recom(nb, users[u], 2, s1) == get_top_N_recom(nbc, copies[u], 2, s2), "coins of stored user %zu after an edit", u);
            }
        }
        for (auto t : tables) delete t;
    }
    // ---- 4. users of ANOTHER vector against the tables (main.cpp:205-216), and against centroids and explicit lists
    // (main.cpp:353-373): the vector has a device copy from tables of its own that were deleted (as user_vectors has after
    // main.cpp:175-176); while no CustVector is created, destroyed or written, all its rows are answered from single calls
    {
        const int P = 20;
        std::vector<Vec> tableUsers = make_users(600, D, 7), queryUsers = make_users(300, D, 8), cents = make_users(12, D, 9), members = make_users(120, D, 10);
        auto give_home = [&]() { auto t0 = create_LSH_hashtables<double>(queryUsers, "cosine", 3, 4, 100, 0.4); for (auto t : t0) delete t; };
        std::vector<CustHashtable<double>*> tabs = create_LSH_hashtables<double>(tableUsers, "cosine", 3, 4, 100, 0.4);
        struct Ans { std::vector<std::string> ids; std::vector<double> sims; std::vector<int> recs; };
        auto lsh_loop = [&](std::vector<Vec>& qs, std::vector<Ans>& ans) {
            ans.assign(qs.size(), Ans());
            for (size_t u = 0; u < qs.size(); u++) {
                std::vector<Vec*> nb = get_LSH_filtered_combined_buckets(tabs, &qs[u]);
                if (nb.empty()) continue;
                ans[u].sims = get_P_closest(nb, qs[u], P);
                for (Vec* v : nb) ans[u].ids.push_back(v->getId());
                ans[u].recs = get_top_N_recom(nb, qs[u], 2, ans[u].sims);
            }
        };
        for (int round = 0; round < 2; round++) {
            give_home();
            std::vector<Ans> warm, cold;
            lsh_loop(queryUsers, warm);
            EXPECT(tabs[0]->set->ext.state == 1 && tabs[0]->set->ext_hashed != nullptr, "round %d: the whole-vector calls did not run (state %d)", round, tabs[0]->set->ext.state);
            std::vector<Vec> copies(queryUsers);       // fresh addresses (and the epoch moves: no home any more)
            lsh_loop(copies, cold);
            for (size_t u = 0; u < queryUsers.size(); u++)
                EXPECT(warm[u].ids == cold[u].ids && warm[u].sims == cold[u].sims && warm[u].recs == cold[u].recs, "round %d: external user %zu", round, u);
            // an edit between the rounds; then one more edit AFTER the home was given: the per-user path must see it
            queryUsers[5].setKnownMean(queryUsers[5].getKnownMean() + 0.5);
            (*queryUsers[9].getDimensions())[2] += 1.0;
        }
        give_home();
        (*queryUsers[11].getDimensions())[4] -= 2.0;
        {
            std::vector<Ans> warm, cold;
            lsh_loop(queryUsers, warm);
            std::vector<Vec> copies(queryUsers);
            lsh_loop(copies, cold);
            for (size_t u = 0; u < queryUsers.size(); u++)
                EXPECT(warm[u].ids == cold[u].ids && warm[u].sims == cold[u].sims && warm[u].recs == cold[u].recs, "after a late edit: external user %zu", u);
        }
        // nearest centroid + the recommendations of that cluster's list
        std::vector<std::vector<Vec*> > lists(3);
        for (size_t i = 0; i < members.size(); i++) lists[i % 3].push_back(&members[i]);
        auto centroid_loop = [&](std::vector<Vec>& qs, std::vector<std::vector<double> >& dist, std::vector<std::vector<int> >& recs) {
            dist.assign(qs.size(), std::vector<double>());
            recs.assign(qs.size(), std::vector<int>());
            for (size_t u = 0; u < qs.size(); u++) {
                for (size_t c = 0; c < cents.size(); c++) dist[u].push_back(qs[u].euclideanDistance(&cents[c]));
                recs[u] = get_top_N_recom(lists[u % 3], qs[u], 2);
            }
        };
        for (int round = 0; round < 2; round++) {
            give_home();
            std::vector<std::vector<double> > wd, cd;
            std::vector<std::vector<int> > wr, cr;
            centroid_loop(queryUsers, wd, wr);
            EXPECT(crx::pair_fan().m_home != nullptr, "round %d: the whole-vector pair call did not run", round);
            std::vector<Vec> copies(queryUsers);
            centroid_loop(copies, cd, cr);
            for (size_t u = 0; u < queryUsers.size(); u++) EXPECT(wd[u] == cd[u] && wr[u] == cr[u], "round %d: centroid loop, user %zu", round, u);
            (*cents[3].getDimensions())[1] += 0.75;
            members[4].setKnownMean(members[4].getKnownMean() - 0.25);
            (*queryUsers[7].getDimensions())[0] += 1.5;
        }
        for (auto t : tabs) delete t;
    }
    if (!g_bad) std::printf("OK\n");
    else std::printf("%d mismatches\n", g_bad);
    return g_bad ? 1 : 0;
}
