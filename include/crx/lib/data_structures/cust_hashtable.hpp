// Drop-in for lib/data_structures/cust_hashtable.hpp (reference cust_hashtable.hpp:24-125).
// A CustHashtable is one table of a set built in one batch on the GPU (crx_create_LSH_hashtables /
// crx_create_hypercube); it keeps raw pointers into the caller's vector exactly like the reference
// (lsh_cube.hpp:70).  Hashes are computed by the engine; this class only does the bucket bookkeeping.
#ifndef LIB_CUST_HASHTABLE_H
#define LIB_CUST_HASHTABLE_H

#include <memory>
#include <vector>

#include "cust_vector.hpp"

namespace crx {

template <typename T>
struct TableSet {
    Packed<T> pts;
    crx_lsh* lsh = nullptr;
    crx_cube* cube = nullptr;
    std::vector<CustVector<T> >* base = nullptr;
    int metric = 0, k = 0, L = 1;
    std::vector<int32_t> ids;   // [L][N] bucket / vertex of every stored row (lazily downloaded)
    std::vector<int32_t> det;   // [L][N][k] euclidean h tuples
    std::vector<std::vector<std::vector<int32_t> > > members;  // [L][bucket] rows in insertion order
    ~TableSet() {
        if (pts.pts) unregister_points(pts.pts);
        if (lsh) crx_lsh_destroy(lsh);
        if (cube) crx_cube_destroy(cube);
    }
    int64_t n() const { return pts.n; }
    void load() {
        if (!ids.empty()) return;
        ids.resize((size_t)L * n());
        if (lsh) {
            check(crx_lsh_bucket_ids(lsh, ids.data(), CRX_HOST), "crx_lsh_bucket_ids");
            if (metric == CRX_EUCLIDEAN) {
                det.resize((size_t)L * n() * k);
                check(crx_lsh_detailed_hashes(lsh, det.data(), CRX_HOST), "crx_lsh_detailed_hashes");
            }
        } else {
            check(crx_cube_vertex_ids(cube, ids.data(), CRX_HOST), "crx_cube_vertex_ids");
        }
        members.resize(L);
        for (int l = 0; l < L; l++) {
            int32_t mx = 0;
            for (int64_t i = 0; i < n(); i++) mx = std::max(mx, ids[(size_t)l * n() + i]);
            members[l].resize((size_t)mx + 1);
            for (int64_t i = 0; i < n(); i++) members[l][ids[(size_t)l * n() + i]].push_back((int32_t)i);
        }
    }
};

}  // namespace crx

template <typename dim_type>
class CustHashtable {
public:
    std::shared_ptr<crx::TableSet<dim_type> > set;
    int table = 0;

    CustHashtable(std::shared_ptr<crx::TableSet<dim_type> > s, int table_index) : set(std::move(s)), table(table_index) {}

    // the tables are built in one batch; late inserts have no equivalent
    int insertVector(CustVector<dim_type>*) {
        std::fprintf(stderr, "crx: CustHashtable::insertVector after construction is not supported\n");
        std::abort();
    }

    // hash of a stored row, or of any other vector through the engine (cust_hashtable.hpp:123)
    void hashes_of(CustVector<dim_type>* q, int32_t* bucket, std::vector<int32_t>* tuple) {
        set->load();
        int32_t row = crx::row_of(*set->base, q);
        int64_t N = set->n();
        if (row >= 0) {
            *bucket = set->ids[(size_t)table * N + row];
            if (tuple && set->metric == CRX_EUCLIDEAN && set->lsh)
                tuple->assign(&set->det[((size_t)table * N + row) * set->k], &set->det[((size_t)table * N + row) * set->k] + set->k);
            return;
        }
        if (!set->lsh) {
            std::fprintf(stderr, "crx: hypercube queries must be stored rows (the reference's EuclideanFGen draws from a dead engine otherwise)\n");
            std::abort();
        }
        std::vector<double> x(q->getDimensions()->begin(), q->getDimensions()->end());
        std::vector<int32_t> b(set->L), d((size_t)set->L * set->k);
        crx::check(crx_lsh_hash_vector(set->lsh, x.data(), b.data(), set->metric == CRX_EUCLIDEAN ? d.data() : nullptr), "crx_lsh_hash_vector");
        *bucket = b[table];
        if (tuple && set->metric == CRX_EUCLIDEAN) tuple->assign(d.begin() + (size_t)table * set->k, d.begin() + (size_t)(table + 1) * set->k);
    }

    int getHash(CustVector<dim_type>* queryVector) {
        int32_t b;
        hashes_of(queryVector, &b, nullptr);
        return b;
    }

    std::vector<CustVector<dim_type>*> getBucketFromIndex(int index) {
        set->load();
        std::vector<CustVector<dim_type>*> out;
        if (index >= 0 && index < (int)set->members[table].size())
            for (int32_t r : set->members[table][index]) out.push_back(&(*set->base)[r]);
        return out;
    }

    std::vector<CustVector<dim_type>*> getBucketFor(CustVector<dim_type>* queryVector) { return getBucketFromIndex(getHash(queryVector)); }

    // members of bucket `index`; with a k-tuple of h values only those stored with the same tuple (cust_hashtable.hpp:81-97)
    std::vector<CustVector<dim_type>*> membersOf(int index, const int32_t* tuple) {
        if (!tuple) return getBucketFromIndex(index);
        set->load();
        std::vector<CustVector<dim_type>*> out;
        int64_t N = set->n();
        if (index >= 0 && index < (int)set->members[table].size())
            for (int32_t r : set->members[table][index]) {
                const int32_t* t = &set->det[((size_t)table * N + r) * set->k];
                bool same = true;
                for (int j = 0; j < set->k; j++) if (t[j] != tuple[j]) { same = false; break; }
                if (same) out.push_back(&(*set->base)[r]);
            }
        return out;
    }

    // cust_hashtable.hpp:74-103: euclidean tables keep only the bucket members whose k-tuple of h equals the query's
    std::vector<CustVector<dim_type>*> getFilteredBucketFor(CustVector<dim_type>* queryVector) {
        int32_t b;
        std::vector<int32_t> tq;
        hashes_of(queryVector, &b, &tq);
        if (!(set->lsh && set->metric == CRX_EUCLIDEAN)) return getBucketFromIndex(b);
        std::vector<CustVector<dim_type>*> out;
        int64_t N = set->n();
        if (b >= 0 && b < (int)set->members[table].size())
            for (int32_t r : set->members[table][b]) {
                const int32_t* t = &set->det[((size_t)table * N + r) * set->k];
                bool same = true;
                for (int j = 0; j < set->k; j++) if (t[j] != tq[j]) { same = false; break; }
                if (same) out.push_back(&(*set->base)[r]);
            }
        return out;
    }
};

#endif  // LIB_CUST_HASHTABLE_H
