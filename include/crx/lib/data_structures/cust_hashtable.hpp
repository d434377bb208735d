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

    // rows stored in bucket b[l] of every table l (euclidean tables with `tuples`: only those stored with the same k-tuple
    // of h values, cust_hashtable.hpp:81-97), each once, ascending: the std::set<CustVector*> order of lsh_cube.hpp:96,104
    std::vector<uint64_t> seen_bits;
    void union_rows(const int32_t* b, const int32_t* tuples, std::vector<int32_t>& rows) {
        load();
        int64_t N = n();
        seen_bits.assign((size_t)(N + 63) / 64, 0);
        for (int l = 0; l < L; l++) {
            if (b[l] < 0 || b[l] >= (int32_t)members[l].size()) continue;
            for (int32_t r : members[l][b[l]]) {
                if (tuples) {
                    const int32_t* t = &det[((size_t)l * N + r) * k];
                    bool same = true;
                    for (int j = 0; j < k; j++) if (t[j] != tuples[(size_t)l * k + j]) { same = false; break; }
                    if (!same) continue;
                }
                seen_bits[(size_t)r >> 6] |= 1ull << (r & 63);
            }
        }
        rows.clear();
        for (size_t w = 0; w < seen_bits.size(); w++)
            for (uint64_t m = seen_bits[w]; m; m &= m - 1) rows.push_back((int32_t)(w * 64 + (size_t)__builtin_ctzll(m)));
    }

    // ---- the recommendation loops (main.cpp:159-170 over the stored rows, :205-216 over another vector of users), answered
    // from ONE batched engine call ----
    // get_LSH_filtered_combined_buckets remembers the candidate list it returned for a user; when get_P_closest is then
    // called with exactly that list and that user, and it is the second such call, crx_recommend_lsh is run once for all
    // users (neighbour rows, similarities, the full literal order of the predicted coins) and this and the following calls
    // are answered from its output.  "All users" = the stored rows (the engine sees them as they were when the tables were
    // built -- the snapshot the per-user calls on registered rows use), or the rows of the vector the user lives in when
    // that vector has a provably current device copy (crx::home_of).
    struct Batch {
        const crx_points* queries = nullptr;  // NULL: the stored rows
        unsigned long epoch = 0;              // of the home registration (external queries)
        int P = -1, nrec = 0, state = 0;      // state: 0 not built, 1 ready, -1 not available for this P
        int wanted = 0;                       // qualifying get_P_closest calls seen so far
        int64_t last_row = -1;                // row of the last combined-bucket query (filtered, all tables) ...
        std::vector<int32_t> last_list;       // ... and the rows it returned
        std::vector<int32_t> rows, recs, ncand;   // [n][P], [n][nrec], [n]
        std::vector<double> sims;                  // [n][P]
        void reset(const crx_points* q, unsigned long e) { *this = Batch(); queries = q; epoch = e; }
    };
    Batch own, ext;
    // bucket ids (and h tuples) of every row of the external home vector, hashed in one engine call
    std::vector<int32_t> ext_ids, ext_det;
    const crx_points* ext_hashed = nullptr;
    unsigned long ext_hashed_epoch = 0;
    int ext_hash_wanted = 0;

    bool batch_for(Batch& b, int P, int64_t nq) {
        if (b.state == 1 && b.P == P) return true;
        if (b.state == -1 && b.P == P) return false;
        if (!lsh || ++b.wanted < 2) return false;
        Timed timed("batched crx_recommend_lsh");
        int nrec = std::min(pts.d, 128);
        // a few hundred MB of host results at most; larger tables keep the per-user calls
        if (P <= 0 || nq <= 0 || (double)nq * (P * 12.0 + nrec * 4.0) > 512e6) { b.state = -1; b.P = P; return false; }
        b.rows.assign((size_t)nq * P, -1); b.sims.assign((size_t)nq * P, 0.0); b.recs.assign((size_t)nq * nrec, 0); b.ncand.assign((size_t)nq, 0);
        int st = crx_recommend_lsh(context(), lsh, b.queries, 0, nq, P, nrec, b.recs.data(), b.rows.data(), b.sims.data(), b.ncand.data(), CRX_HOST);
        b.P = P; b.nrec = nrec;
        b.state = st == CRX_OK ? 1 : -1;   // outside the batched call's limits (P, D): the per-user calls answer
        return b.state == 1;
    }
    // all rows of the home vector hashed for every table (second qualifying call on)
    bool ext_hash_for(const Registered* home) {
        if (ext_hashed == home->pts && ext_hashed_epoch == home->gen) return true;
        if (!lsh || crx_points_d(home->pts) != pts.d || ++ext_hash_wanted < 2 || (double)home->n * L * (k + 1) * 4.0 > 512e6) return false;
        Timed timed("batched crx_lsh_hash_points");
        ext_ids.assign((size_t)L * home->n, 0);
        ext_det.assign(metric == CRX_EUCLIDEAN ? (size_t)L * home->n * k : 0, 0);
        check(crx_lsh_hash_points(lsh, home->pts, ext_ids.data(), metric == CRX_EUCLIDEAN ? ext_det.data() : nullptr, CRX_HOST), "crx_lsh_hash_points");
        ext_hashed = home->pts; ext_hashed_epoch = home->gen;
        return true;
    }
    ~TableSet() {
        if (pts.pts && retire_points(pts.pts)) pts.pts = nullptr;   // the registry keeps the device copy (crx_shim.hpp)
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
        std::vector<double> x(q->crxDimsRef().begin(), q->crxDimsRef().end());
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
