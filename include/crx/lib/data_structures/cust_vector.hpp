// Drop-in for lib/data_structures/cust_vector.hpp (reference cust_vector.hpp:23-72): same class name, same
// public interface.  The container part lives on the host; the four vector operations are evaluated by
// the engine (crx_pair_op) -- there is no host arithmetic fallback.
#ifndef LIB_CUST_VECTOR_H
#define LIB_CUST_VECTOR_H

#include <cmath>
#include <iostream>
#include <set>
#include <string>
#include <utility>
#include <vector>

#include "../../crx_shim.hpp"

namespace crx {
// One-row device copies of the vectors the per-pair operations are called on, kept in a small cache: loops like
// `for (c : centroids) user.euclideanDistance(c)` (main.cpp:356-363) then upload each vector once, not once per call.
// An entry is reused only when both the address and the coordinates are unchanged.
struct OneRow {
    const void* key = nullptr;
    std::vector<double> dims;
    crx_points* pts = nullptr;
    unsigned long stamp = 0;
};
inline crx_points* one_row_points(const void* key, const std::vector<double>& dims) {
    static std::vector<OneRow> cache(64);
    static unsigned long clock_ = 0;
    OneRow* victim = &cache[0];
    for (OneRow& e : cache) {
        if (e.pts && e.key == key && e.dims == dims) { e.stamp = ++clock_; return e.pts; }
        if (e.stamp < victim->stamp) victim = &e;
    }
    if (victim->pts) crx_points_destroy(victim->pts);
    victim->pts = nullptr;
    check(crx_points_create(context(), dims.data(), CRX_F64, 1, (int)dims.size(), CRX_HOST, &victim->pts), "crx_points_create");
    victim->key = key; victim->dims = dims; victim->stamp = ++clock_;
    return victim->pts;
}

// The same loops call ONE left operand against a recurring family of right operands (a user against every centroid).
// The right operands seen lately are kept as rows of one device set; when a new left operand meets one of them, the
// operation is evaluated against the whole family in one engine call (crx_pair_op with m pairs; every pair is computed
// exactly as a single-pair call computes it) and the following calls of that left operand are answered from the result.
// An answer is handed out only after the coordinates of both operands were compared with what it was computed from.
struct PairFan {
    static const int MAXB = 256;
    struct B { const void* key; std::vector<double> dims; unsigned long stamp; };
    std::vector<B> family;            // right operands, most of them recurring
    bool family_dirty = true;         // device set out of date
    crx_points* family_pts = nullptr;
    int family_d = 0;
    unsigned long clock_ = 0;
    // the last fan-out
    const void* a_key = nullptr;
    std::vector<double> a_dims;
    int op = -1;
    std::vector<double> out;          // [family.size()] at the time of the call
    bool valid = false;
    // all rows of the left operand's home vector (crx::home_of) against the family, in one call
    const crx_points* m_home = nullptr;
    unsigned long m_epoch = 0, family_version = 0, m_version = 0;
    int m_op = -1;
    std::vector<double> m_out;        // [home rows][family.size()]
    const crx_points* seen_home = nullptr;   // home and row of the last left operand that met the family
    int64_t seen_row = -1;

    int find(const void* key, const std::vector<double>& dims) {
        for (size_t i = 0; i < family.size(); i++) if (family[i].key == key && family[i].dims == dims) return (int)i;
        return -1;
    }
    void remember(const void* key, const std::vector<double>& dims) {
        family_version++;
        for (B& b : family) if (b.key == key) { b.dims = dims; b.stamp = ++clock_; family_dirty = true; valid = false; return; }
        if ((int)family.size() < MAXB) family.push_back(B{key, dims, ++clock_});
        else {
            B* victim = &family[0];
            for (B& b : family) if (b.stamp < victim->stamp) victim = &b;
            *victim = B{key, dims, ++clock_};
        }
        family_dirty = true; valid = false;
    }
    bool upload_family(int d) {
        for (size_t i = 0; i < family.size(); i++) if ((int)family[i].dims.size() != d) return false;
        if (family.size() < 2) return false;
        if (family_dirty || family_d != d) {
            if (family_pts) crx_points_destroy(family_pts);
            family_pts = nullptr;
            std::vector<double> buf(family.size() * (size_t)d);
            for (size_t i = 0; i < family.size(); i++) std::copy(family[i].dims.begin(), family[i].dims.end(), buf.begin() + i * (size_t)d);
            check(crx_points_create(context(), buf.data(), CRX_F64, (int64_t)family.size(), d, CRX_HOST, &family_pts), "crx_points_create");
            family_dirty = false; family_d = d;
        }
        return true;
    }
    // every row of `home` against every family member (the loop main.cpp:353-366 runs for all users, in one call)
    bool fan_home(const Registered* home, int d, int which_op) {
        size_t F = family.size();
        if (crx_points_d(home->pts) != d || (double)home->n * (double)F > 4e6 || !upload_family(d)) return false;
        Timed timed("pair operation, whole vector x family");
        std::vector<int32_t> ra((size_t)home->n * F), rb((size_t)home->n * F);
        for (int64_t i = 0; i < home->n; i++)
            for (size_t j = 0; j < F; j++) { ra[(size_t)i * F + j] = (int32_t)i; rb[(size_t)i * F + j] = (int32_t)j; }
        m_out.assign(ra.size(), 0.0);
        check(crx_pair_op(context(), home->pts, ra.data(), family_pts, rb.data(), (int64_t)ra.size(), which_op, m_out.data()), "crx_pair_op");
        m_home = home->pts; m_epoch = home->gen; m_op = which_op; m_version = family_version;
        return true;
    }
    // distances of `a` to every family member of its dimension; false when there is nothing to fan out to
    bool fan(const void* key, const std::vector<double>& a, int which_op) {
        int d = (int)a.size();
        std::vector<int32_t> rows_b;
        for (size_t i = 0; i < family.size(); i++) if ((int)family[i].dims.size() == d) rows_b.push_back((int32_t)i);
        if (rows_b.size() < 2 || rows_b.size() != family.size()) return false;
        if (family_dirty || family_d != d) {
            if (family_pts) crx_points_destroy(family_pts);
            family_pts = nullptr;
            std::vector<double> buf(family.size() * (size_t)d);
            for (size_t i = 0; i < family.size(); i++) std::copy(family[i].dims.begin(), family[i].dims.end(), buf.begin() + i * (size_t)d);
            check(crx_points_create(context(), buf.data(), CRX_F64, (int64_t)family.size(), d, CRX_HOST, &family_pts), "crx_points_create");
            family_dirty = false; family_d = d;
        }
        crx_points* pa = one_row_points(key, a);
        std::vector<int32_t> rows_a(rows_b.size(), 0);
        out.assign(rows_b.size(), 0.0);
        check(crx_pair_op(context(), pa, rows_a.data(), family_pts, rows_b.data(), (int64_t)rows_b.size(), which_op, out.data()), "crx_pair_op");
        a_key = key; a_dims = a; op = which_op; valid = true;
        return true;
    }
};
inline PairFan& pair_fan() { static PairFan f; return f; }
}  // namespace crx

template <typename dim_type>
class CustVector {
private:
    std::string id;
    std::vector<dim_type> dimensions;
    std::set<int> unknown_indexes;
    double known_mean;
    int cluster_i;
    double dist_from_centroid;

    template <typename in_dim_type>
    double pair_op(CustVector<in_dim_type>* other, int op) {
        const std::vector<in_dim_type>* od = &other->crxDimsRef();
        int d = (int)dimensions.size();
        std::vector<double> a(dimensions.begin(), dimensions.end()), b(od->begin(), od->end());
        (void)d;
        crx::Timed timed("CustVector pair operation");
        if (a.size() == b.size()) {
            crx::PairFan& f = crx::pair_fan();
            int at = f.find(other, b);
            if (at >= 0) {
                f.family[at].stamp = ++f.clock_;
                // the left operand is a row of a vector with a current device copy: from its second row on, all rows at once
                if (const crx::Registered* home = crx::home_of(this, sizeof(*this))) {
                    int64_t hrow = (int64_t)(((const char*)this - home->begin) / (ptrdiff_t)home->stride);
                    bool have = f.m_home == home->pts && f.m_epoch == home->gen && f.m_op == op && f.m_version == f.family_version;
                    if (!have && f.seen_home == home->pts && f.seen_row != hrow) have = f.fan_home(home, (int)a.size(), op);
                    f.seen_home = home->pts; f.seen_row = hrow;
                    if (have) return f.m_out[(size_t)hrow * f.family.size() + (size_t)at];
                }
                if (!(f.valid && f.a_key == (const void*)this && f.op == op && f.a_dims == a)) f.fan(this, a, op);
                if (f.valid && f.a_key == (const void*)this && f.op == op && f.a_dims == a) return f.out[at];
            } else {
                f.remember(other, b);
            }
        }
        crx_points* pa = crx::one_row_points(this, a);
        crx_points* pb = crx::one_row_points(other, b);   // (both stay valid: the cache holds 64 entries)
        int32_t zero = 0;
        double out = 0;
        crx::check(crx_pair_op(crx::context(), pa, &zero, pb, &zero, 1, op, &out), "crx_pair_op");
        return out;
    }

public:
    // every member that creates, destroys or can alter the coordinates / unknown set / mean reports it (crx::touched)
    CustVector(std::string in_id, std::vector<dim_type> dim_vector)
        : id(std::move(in_id)), dimensions(std::move(dim_vector)), known_mean(0), cluster_i(-1), dist_from_centroid(0) { crx::touched(this); }
    CustVector(std::string in_id, std::vector<dim_type> dim_vector, std::set<int> indexes, double mean)
        : id(std::move(in_id)), dimensions(std::move(dim_vector)), unknown_indexes(std::move(indexes)), known_mean(mean),
          cluster_i(-1), dist_from_centroid(0) { crx::touched(this); }
    CustVector(std::string in_id, std::vector<dim_type> dim_vector, int cluster, double distance)
        : id(std::move(in_id)), dimensions(std::move(dim_vector)), known_mean(0), cluster_i(cluster), dist_from_centroid(distance) { crx::touched(this); }
    CustVector(const CustVector& o)
        : id(o.id), dimensions(o.dimensions), unknown_indexes(o.unknown_indexes), known_mean(o.known_mean), cluster_i(o.cluster_i),
          dist_from_centroid(o.dist_from_centroid) { crx::touched(this); }
    CustVector(CustVector&& o) noexcept
        : id(std::move(o.id)), dimensions(std::move(o.dimensions)), unknown_indexes(std::move(o.unknown_indexes)), known_mean(o.known_mean),
          cluster_i(o.cluster_i), dist_from_centroid(o.dist_from_centroid) { crx::touched(this); crx::touched(&o); }
    CustVector& operator=(const CustVector& o) {
        id = o.id; dimensions = o.dimensions; unknown_indexes = o.unknown_indexes; known_mean = o.known_mean; cluster_i = o.cluster_i;
        dist_from_centroid = o.dist_from_centroid; crx::touched(this);
        return *this;
    }
    CustVector& operator=(CustVector&& o) noexcept {
        id = std::move(o.id); dimensions = std::move(o.dimensions); unknown_indexes = std::move(o.unknown_indexes); known_mean = o.known_mean;
        cluster_i = o.cluster_i; dist_from_centroid = o.dist_from_centroid; crx::touched(this); crx::touched(&o);
        return *this;
    }
    ~CustVector() { crx::touched(this); }

    // cust_vector.hpp:107-121.  Dimension mismatch prints and returns -1 like the reference.
    template <typename in_dim_type>
    long double inner_product(CustVector<in_dim_type>* inVector, long double strt) {
        if (dimensions.size() != inVector->crxDimsRef().size()) {
            std::cerr << id << " : Error in inner product with " << inVector->getId() << ". Different number of dimensions" << std::endl;
            return -1;
        }
        return strt + (long double)pair_op(inVector, 0);
    }
    template <typename in_dim_type>
    double euclideanDistance(CustVector<in_dim_type>* inVector) { return pair_op(inVector, 1); }  // :126-136
    template <typename in_dim_type>
    double cosineDistance(CustVector<in_dim_type>* inVector) { return pair_op(inVector, 2); }     // :141-155
    template <typename in_dim_type>
    double cosineSimilarity(CustVector<in_dim_type>* inVector) { return pair_op(inVector, 3); }   // :160-174

    // container updates (cust_vector.hpp:179-194); the engine computes cluster means itself (crx_k_means)
    template <typename in_dim_type>
    void addVectorToThis(CustVector<in_dim_type>* inVector) {
        crx::touched(this);
        const std::vector<in_dim_type>& in = inVector->crxDimsRef();
        for (size_t i = 0; i < dimensions.size(); i++) dimensions[i] = dimensions[i] + in[i];
    }
    void divDimensionsByD(double div_const) {
        crx::touched(this);
        if (div_const != 0)
            for (size_t i = 0; i < dimensions.size(); i++) dimensions[i] = dimensions[i] / div_const;
    }

    void setCluster(int index, double dist) { cluster_i = index; dist_from_centroid = dist; }
    void resetCluster() { cluster_i = -1; dist_from_centroid = 0; }
    void setKnownMean(double in_mean) { known_mean = in_mean; crx::touched(this); }
    void setUnknownIndexes(std::set<int> in_indexes) { unknown_indexes = std::move(in_indexes); crx::touched(this); }

    std::string getId() { return id; }
    std::vector<dim_type>* getDimensions() { crx::touched(this); return &dimensions; }   // hands out write access
    std::vector<int> getUnknownIndexes() { return std::vector<int>(unknown_indexes.begin(), unknown_indexes.end()); }
    std::set<int> getUnknownIndexesSet() { return unknown_indexes; }
    // not in the reference: read-only views for the packing code of the drop-in headers (the getters above copy, as the
    // reference's do: cust_vector.hpp:229-236)
    const std::set<int>& crxUnknownRef() const { return unknown_indexes; }
    const std::vector<dim_type>& crxDimsRef() const { return dimensions; }
    double getKnownMean() { return known_mean; }
    unsigned int getDimNumber() { return (unsigned int)dimensions.size(); }
    int getCluster() { return cluster_i; }
    double getDistFromCentroid() { return dist_from_centroid; }
};

#endif  // LIB_CUST_VECTOR_H
