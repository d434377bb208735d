// tables.cuh -- shared declarations for LSH tables / hypercube / segment building.
#pragma once
#include <random>

#include "common.cuh"

// rows grouped by an int32 key: perm = rows stably sorted by key (insertion order inside a bucket is
// ascending row, as in VectorBucket::insertVector, vector_bucket.hpp:41), sorted = keys in that
// order, off[b] = first position of bucket b (off[nkeys] = n).
struct Segments {
    int32_t* perm = nullptr;
    int32_t* sorted = nullptr;
    int32_t* off = nullptr;  // [nkeys + 1]
    int64_t n = 0;
    int nkeys = 0;
    crx_ctx* owner = nullptr;
    void free_all() {
        if (owner) { crx_free(owner, perm); crx_free(owner, sorted); crx_free(owner, off); }
        perm = sorted = off = nullptr;
    }
};

// frees a function-local Segments on every exit path (early error returns included); the tables keep theirs by value
struct SegmentsGuard {
    Segments& s;
    explicit SegmentsGuard(Segments& seg) : s(seg) {}
    ~SegmentsGuard() { s.free_all(); }
    SegmentsGuard(const SegmentsGuard&) = delete;
    SegmentsGuard& operator=(const SegmentsGuard&) = delete;
};

// keys[n] on the device, values in [0, nkeys).  Allocates the three arrays with cudaMalloc.
int crx_build_segments(crx_ctx* c, const int32_t* keys, int64_t n, int nkeys, Segments* out);

struct crx_lsh {
    crx_ctx* ctx = nullptr;
    const crx_points* pts = nullptr;
    int metric = 0, k = 0, L = 0, D = 0;
    int64_t N = 0;
    int nbuckets = 0;
    float w = 0;
    // host copies of the drawn parameters
    std::vector<double> cos_r;  // [L][k][D]
    std::vector<float> euc_v;   // [L][k][D]
    std::vector<float> euc_t;   // [L][k]
    std::vector<int32_t> euc_r; // [L][k]
    // device parameters
    double* d_proj = nullptr;   // [L*k][ldp] projection vectors widened to double
    double* d_pnorm = nullptr;  // [L*k] Euclidean norm of each projection vector (error bound)
    float* d_t = nullptr;       // [L*k]
    int32_t* d_r = nullptr;     // [L*k]
    int ldp = 0;
    // per stored row
    int32_t* hvals = nullptr;   // [L][N][k] h values (euclidean only)
    int32_t* bucket = nullptr;  // [L][N] bucket index (CustHashtable::getHash)
    int32_t* gid = nullptr;     // [L][N] filtered-candidate group: bucket (cosine) / rank of the k-tuple (euclidean)
    std::vector<Segments> by_bucket;  // [L]
    std::vector<Segments> by_group;   // [L] (aliases by_bucket for cosine)
    std::vector<int> ngroups;         // [L]
};

struct crx_cube {
    crx_ctx* ctx = nullptr;
    const crx_points* pts = nullptr;
    int metric = 0, k = 0, D = 0;
    int64_t N = 0;
    float w = 0;
    std::vector<double> cos_r;  // [k][D]
    std::vector<float> euc_v;   // [k][D]
    std::vector<float> euc_t;   // [k]
    // euclidean_f_gen.hpp:33 maps, one per f, and the engine that keeps drawing for unseen h
    std::vector<std::vector<std::pair<int32_t, int32_t>>> fmap;  // sorted by h
    std::default_random_engine engine;
    int32_t* vertex = nullptr;  // [N]
    Segments by_vertex;
};

// hash every row of `pts` with the H = L*k projections in d_proj.
//   euclid: raw h -> hvals[L][N][k] (if non-NULL), phi bucket -> bucket[L][N] (if non-NULL)
//   cosine: g -> bucket[L][N]
int crx_hash_rows(crx_ctx* c, const crx_points* pts, int metric, int k, int L, const double* d_proj, int ldp,
                  const double* d_pnorm, const float* d_t, const int32_t* d_r, float w, int nbuckets, int32_t* hvals,
                  int32_t* bucket);
