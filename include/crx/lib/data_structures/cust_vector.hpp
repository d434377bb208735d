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
        std::vector<in_dim_type>* od = other->getDimensions();
        int d = (int)dimensions.size();
        std::vector<double> a(dimensions.begin(), dimensions.end()), b(od->begin(), od->end());
        (void)d;
        crx_points* pa = crx::one_row_points(this, a);
        crx_points* pb = crx::one_row_points(other, b);   // (both stay valid: the cache holds 64 entries)
        int32_t zero = 0;
        double out = 0;
        crx::check(crx_pair_op(crx::context(), pa, &zero, pb, &zero, 1, op, &out), "crx_pair_op");
        return out;
    }

public:
    CustVector(std::string in_id, std::vector<dim_type> dim_vector)
        : id(std::move(in_id)), dimensions(std::move(dim_vector)), known_mean(0), cluster_i(-1), dist_from_centroid(0) {}
    CustVector(std::string in_id, std::vector<dim_type> dim_vector, std::set<int> indexes, double mean)
        : id(std::move(in_id)), dimensions(std::move(dim_vector)), unknown_indexes(std::move(indexes)), known_mean(mean),
          cluster_i(-1), dist_from_centroid(0) {}
    CustVector(std::string in_id, std::vector<dim_type> dim_vector, int cluster, double distance)
        : id(std::move(in_id)), dimensions(std::move(dim_vector)), known_mean(0), cluster_i(cluster), dist_from_centroid(distance) {}
    CustVector(const CustVector& o) = default;
    CustVector& operator=(const CustVector& o) = default;

    // cust_vector.hpp:107-121.  Dimension mismatch prints and returns -1 like the reference.
    template <typename in_dim_type>
    long double inner_product(CustVector<in_dim_type>* inVector, long double strt) {
        if (dimensions.size() != inVector->getDimensions()->size()) {
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
        std::vector<in_dim_type>& in = *inVector->getDimensions();
        for (size_t i = 0; i < dimensions.size(); i++) dimensions[i] = dimensions[i] + in[i];
    }
    void divDimensionsByD(double div_const) {
        if (div_const != 0)
            for (size_t i = 0; i < dimensions.size(); i++) dimensions[i] = dimensions[i] / div_const;
    }

    void setCluster(int index, double dist) { cluster_i = index; dist_from_centroid = dist; }
    void resetCluster() { cluster_i = -1; dist_from_centroid = 0; }
    void setKnownMean(double in_mean) { known_mean = in_mean; }
    void setUnknownIndexes(std::set<int> in_indexes) { unknown_indexes = std::move(in_indexes); }

    std::string getId() { return id; }
    std::vector<dim_type>* getDimensions() { return &dimensions; }
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
