// dropin_driver.cpp -- one translation unit, two builds:
//   (a) against include/crx/lib/** + libcrx.so  (the product's drop-in headers), and
//   (b) -DCRX_REFERENCE_BUILD against the reference's own headers (in the build container only; used by
//       tests/golden/make_dropin_golden.sh to produce tests/golden/dropin_expected.txt).
// It calls the hot-path API with the reference's own names and signatures in the order main.cpp does
// (main.cpp:149-185, 240-275) plus the range-search / PAM / silhouette library calls.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <numeric>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#ifdef CRX_REFERENCE_BUILD
static uint64_t g_seed = 1;
namespace std { namespace chrono {
struct crx_fake_clock {
    struct dur { unsigned long count() const { return (unsigned long)g_seed; } };
    struct tp { dur time_since_epoch() const { return dur(); } };
    static tp now() { return tp(); }
};
} }
#define system_clock crx_fake_clock
#include "lib/in_out/vector_reader.hpp"
#include "lib/data_structures/cust_vector.hpp"
#include "lib/data_structures/cust_hashtable.hpp"
template <>
int CustHashtable<double>::insertVector(CustVector<double>* inVector) {
    unsigned int index = mod(hashGenerator->generate(inVector), buckets.size());
    buckets[index]->insertVector(inVector);
    return 0;
}
#include "lib/lsh_cube.hpp"
#include "lib/clustering_phases/initialization.hpp"
#include "lib/clustering_phases/assignment.hpp"
#include "lib/clustering_phases/update.hpp"
#include "lib/clustering_phases/silhouette.hpp"
#include "lib/crypto_rec.hpp"
#undef system_clock
static void set_seed_all(uint64_t s) { g_seed = s; }
#else
#include "lib/data_structures/cust_vector.hpp"
#include "lib/data_structures/cust_hashtable.hpp"
#include "lib/lsh_cube.hpp"
#include "lib/clustering_phases/initialization.hpp"
#include "lib/clustering_phases/assignment.hpp"
#include "lib/clustering_phases/update.hpp"
#include "lib/clustering_phases/silhouette.hpp"
#include "lib/crypto_rec.hpp"
static void set_seed_all(uint64_t s) { crx::set_seed(s); }
#endif

using namespace std;

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: %s input.bin output.txt\n", argv[0]); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 3;
    int64_t N; int32_t D;
    if (fread(&N, 8, 1, f) != 1 || fread(&D, 4, 1, f) != 1) return 4;
    vector<double> X((size_t)N * D), mean(N);
    vector<uint8_t> unk((size_t)N * D);
    if (fread(X.data(), 8, X.size(), f) != X.size() || fread(unk.data(), 1, unk.size(), f) != unk.size() || fread(mean.data(), 8, N, f) != (size_t)N) return 5;
    fclose(f);
    vector<CustVector<double>> users;
    users.reserve(N);
    for (int64_t i = 0; i < N; i++) {
        set<int> u;
        for (int j = 0; j < D; j++) if (unk[i * D + j]) u.insert(j);
        users.emplace_back(to_string(i), vector<double>(X.begin() + i * D, X.begin() + (i + 1) * D), u, mean[i]);
    }
    FILE* o = fopen(argv[2], "w");
    const int P = 20;

    // ---- Cosine LSH recommendation, part A (main.cpp:149-176)
    {
        string metric_type = "cosine";
        set_seed_all(1001);
        vector<CustHashtable<double>*> tabs = create_LSH_hashtables<double>(users, metric_type, 4, 5, 100, 0.4);
        for (auto& user : users) {
            vector<CustVector<double>*> neighbors = get_LSH_filtered_combined_buckets(tabs, &user);
            fprintf(o, "recA %s ncand %zu", user.getId().c_str(), neighbors.size());
            if (!neighbors.empty()) {
                vector<double> sims = get_P_closest(neighbors, user, P);
                vector<int> recs = get_top_N_recom(neighbors, user, 5, sims);
                fprintf(o, " nbr");
                for (auto nb : neighbors) fprintf(o, " %s", nb->getId().c_str());
                fprintf(o, " sim0 %.12g recs", sims[0]);
                for (int r : recs) fprintf(o, " %d", r);
            }
            fprintf(o, "\n");
        }
        for (auto t : tabs) delete t;
    }
    // ---- clustering recommendation (main.cpp:240-275) with k-means++ instead of rand_selection
    {
        string metric_type = "euclidean";
        set_seed_all(1002);
        vector<CustVector<double>*> centroids = k_means_pp(users, 8, metric_type);
        fprintf(o, "kpp");
        for (auto c : centroids) fprintf(o, " %s", c->getId().c_str());
        fprintf(o, "\n");
        bool cont = true;
        int it = 0;
        while (cont && it < 3) {
            lloyds_assignment(users, centroids, metric_type);
            fprintf(o, "labels%d", it);
            for (auto& u : users) fprintf(o, " %d", u.getCluster());
            fprintf(o, "\n");
            cont = k_means(users, centroids, metric_type, 0.05);
            fprintf(o, "cont%d %d c0 %.12g %.12g\n", it, (int)cont, (*centroids[0]->getDimensions())[0], (*centroids[7]->getDimensions())[D - 1]);
            it++;
        }
        vector<vector<CustVector<double>*>> clusters = separate_clusters_from_input(users, (int)centroids.size());
        for (auto& user : users) {
            vector<CustVector<double>*> neighbors = clusters[user.getCluster()];
            if (!neighbors.empty()) {
                vector<int> recs = get_top_N_recom(neighbors, user, 5);
                fprintf(o, "recC %s", user.getId().c_str());
                for (int r : recs) fprintf(o, " %d", r);
                fprintf(o, "\n");
            }
        }
        vector<double> sil = silhouette_cluster(clusters, centroids, metric_type);
        fprintf(o, "sil");
        for (double s : sil) fprintf(o, " %.10g", s);
        fprintf(o, "\n");
        for (auto c : centroids) if (c->getId() == "k_means_center") delete c;
    }
    // ---- range-search assignment + PAM (library API, assignment.hpp:109-152, update.hpp:90)
    for (int m = 0; m < 2; m++) {
        string metric_type = m == 0 ? "euclidean" : "cosine";
        set_seed_all(1003 + m);
        vector<CustVector<double>*> centroids = rand_selection(users, 6);
        fprintf(o, "rand");
        for (auto c : centroids) fprintf(o, " %s", c->getId().c_str());
        fprintf(o, "\n");
        vector<CustHashtable<double>*> tabs = create_LSH_hashtables<double>(users, metric_type, 3, 4, 10, 2.0);
        lsh_range_assignment(users, tabs, centroids, metric_type);
        fprintf(o, "lshrange%d", m);
        for (auto& u : users) fprintf(o, " %d", u.getCluster());
        fprintf(o, "\n");
        CustHashtable<double>* cube = create_hypercube<double>(users, metric_type, 5, 2.0);
        cube_range_assignment(users, *cube, centroids, metric_type, 7, 5);
        fprintf(o, "cuberange%d", m);
        for (auto& u : users) fprintf(o, " %d", u.getCluster());
        fprintf(o, "\n");
        fprintf(o, "dist%d %.12g %.12g\n", m, users[1].getDistFromCentroid(), users[N - 1].getDistFromCentroid());
        bool sw = pam_lloyds(users, centroids, metric_type);
        fprintf(o, "pam%d %d", m, (int)sw);
        for (auto c : centroids) fprintf(o, " %s", c->getId().c_str());
        fprintf(o, "\n");
        fprintf(o, "vec%d %.12g %.12g %.12g %.12g\n", m, (double)users[0].inner_product<double>(&users[1], 0.0), users[0].euclideanDistance(&users[1]),
                users[0].cosineDistance(&users[1]), users[0].cosineSimilarity(&users[1]));
        for (auto t : tabs) delete t;
        delete cube;
    }
    // ---- {lsh_range_assignment, k_means} iterated (assignment.hpp:109-129 + update.hpp:38-86): from the second iteration the
    // centroids are heap vectors that all carry the id "k_means_center"
    for (int m = 0; m < 2; m++) {
        string metric_type = m == 0 ? "euclidean" : "cosine";
        set_seed_all(1010 + m);
        vector<CustVector<double>*> centroids = rand_selection(users, 5);
        vector<CustHashtable<double>*> tabs = create_LSH_hashtables<double>(users, metric_type, 3, 4, 10, 2.0);
        bool cont = true;
        int it = 0;
        while (cont && it < 3) {
            lsh_range_assignment(users, tabs, centroids, metric_type);
            fprintf(o, "rangeloop%d_%d", m, it);
            for (auto& u : users) fprintf(o, " %d", u.getCluster());
            fprintf(o, "\n");
            cont = k_means(users, centroids, metric_type, 0.0);
            it++;
        }
        for (auto c : centroids) if (c->getId() == "k_means_center") delete c;
        for (auto t : tabs) delete t;
    }
    fclose(o);
    return 0;
}
