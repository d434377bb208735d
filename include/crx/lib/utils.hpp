// Drop-in for the hot-path helpers of lib/utils.hpp / utils.cpp (mod :97-98, get_num_hamming_dist_from
// utils.cpp:22-50, remove_clustering :143-147, separate_clusters_from_input :150-158,
// find_min_vector_distance :161-178) and of its string / file helpers (split utils.cpp:11-19, file_to_args :53-69,
// file_to_str_vectors :72-133, file_to_lexicon :136-152), which are plain host I/O.  Everything is inline: this header
// replaces utils.hpp AND utils.cpp (do not compile the reference's utils.cpp next to it).
#ifndef LIB_UTILS_H
#define LIB_UTILS_H

#include <cmath>
#include <fstream>
#include <functional>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "./data_structures/cust_vector.hpp"

// ---- strings and files ---------------------------------------------------------------------------
inline std::vector<std::string> split(const std::string& s, char delimiter) {
    std::vector<std::string> out;
    std::istringstream in(s);
    for (std::string piece; std::getline(in, piece, delimiter);) out.push_back(piece);
    return out;
}

template <typename conv_type>
std::vector<conv_type> split_convert(const std::string& s, char delimiter, std::function<conv_type(const std::string&)> conversion_f) {
    std::vector<conv_type> out;
    std::istringstream in(s);
    for (std::string piece; std::getline(in, piece, delimiter);) out.emplace_back(conversion_f(piece));
    return out;
}

template <typename a_type, typename b_type>
bool f_equals(a_type a, b_type b, double epsilon) { return std::abs(a - b) < epsilon; }

namespace crx {
// every line of a text file, split at `delimiter`; a trailing '\r' is dropped when strip_cr
inline bool read_split_lines(const std::string& filename, char delimiter, bool strip_cr, std::vector<std::vector<std::string> >& rows) {
    std::ifstream f(filename);
    if (!f.is_open()) return false;
    for (std::string line; std::getline(f, line);) {
        if (strip_cr && !line.empty() && line.back() == '\r') line.pop_back();
        rows.emplace_back(split(line, delimiter));
    }
    return true;
}
}  // namespace crx

inline std::vector<std::string> file_to_args(std::string filename, char delimiter) {
    std::vector<std::vector<std::string> > rows;
    std::vector<std::string> args;
    crx::read_split_lines(filename, delimiter, false, rows);
    for (auto& r : rows) args.insert(args.end(), r.begin(), r.end());
    return args;
}

inline std::vector<std::vector<std::string> > file_to_str_vectors(std::string filename, char delimiter) {
    std::vector<std::vector<std::string> > rows;
    crx::read_split_lines(filename, delimiter, true, rows);
    return rows;
}

// first line = "<label><delimiter><P>" (utils.cpp:105-113); the remaining lines are the records
inline std::vector<std::vector<std::string> > file_to_str_vectors(std::string filename, char delimiter, int* P) {
    std::vector<std::vector<std::string> > rows;
    if (!crx::read_split_lines(filename, delimiter, true, rows) || rows.empty()) return rows;
    if (rows[0].size() > 1) *P = std::stoi(rows[0][1]);
    rows.erase(rows.begin());
    return rows;
}

inline std::unordered_map<std::string, float> file_to_lexicon(std::string filename, char delimiter) {
    std::vector<std::vector<std::string> > rows;
    std::unordered_map<std::string, float> lexicon;
    crx::read_split_lines(filename, delimiter, false, rows);
    for (auto& r : rows) lexicon.emplace(r.at(0), std::stof(r.at(1)));
    return lexicon;
}

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

namespace crx {
template <typename q_type, typename dim_type>
double min_vector_dist(CustVector<q_type>* q, std::vector<CustVector<dim_type> >* vectors, int op) {
    if (vectors->empty()) return 0;
    Packed<dim_type> P;
    P.from_vector(*vectors);
    std::vector<CustVector<q_type>*> one(1, q);
    Packed<q_type> Q;
    Q.from_pointers(one);
    std::vector<int32_t> a(vectors->size(), 0), b(vectors->size());
    for (size_t i = 0; i < b.size(); i++) b[i] = (int32_t)i;
    std::vector<double> d(b.size());
    check(crx_pair_op(context(), Q.pts, a.data(), P.pts, b.data(), (int64_t)b.size(), op, d.data()), "crx_pair_op");
    double m = d[0];
    for (double v : d) if (v < m) m = v;
    return m;
}
}  // namespace crx

// utils.hpp:107-140
template <typename q_type, typename dim_type>
double min_vector_euclidean_dist(CustVector<q_type>* queryVector, std::vector<CustVector<dim_type> >* vectors) {
    return crx::min_vector_dist(queryVector, vectors, 1);
}
template <typename q_type, typename dim_type>
double min_vector_cosine_dist(CustVector<q_type>* queryVector, std::vector<CustVector<dim_type> >* vectors) {
    return crx::min_vector_dist(queryVector, vectors, 2);
}

#endif  // LIB_UTILS_H
