// ingest.cu -- binary columnar ingest: the step in front of the path.
//
// The reference reads its vector files line by line: getline, substr, one istringstream per line and stod per token
// (vector_reader.hpp:55-85, utils.hpp:85-94) -- 0.6 s for the 20k x 203 tweet vectors of its own run.  A file converted
// once (crx_columnar_write / tools/csv_to_columnar.py; the values are the strtod of every token, i.e. what stod returns)
// holds the ids and the coordinates as float64 COLUMNS; loading it is one mapped read, one upload and one transpose on the
// GPU into the row-major layout the engine works on.
//
//   "CRXCOL1\0" | int64 n | int32 d | int32 0 | int64 ids_bytes | n NUL-terminated ids | zero padding to 8 bytes |
//   float64 [d][n]   (coordinate j of all n vectors contiguous)
#include "common.cuh"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <string>
#include <vector>

struct crx_columnar {
    void* map = nullptr;
    size_t bytes = 0;
    int64_t n = 0;
    int d = 0;
    std::vector<const char*> ids;
    const double* cols = nullptr;   // [d][n] inside the mapping
};

namespace {

constexpr char MAGIC[8] = {'C', 'R', 'X', 'C', 'O', 'L', '1', '\0'};
constexpr size_t HEADER = 8 + 8 + 4 + 4 + 8;

// [d][n] -> [n][d]: 32 x 32 tiles through shared memory, both sides coalesced
__global__ void __launch_bounds__(256) columns_to_rows_kernel(const double* __restrict__ cols, int64_t n, int d, double* __restrict__ rows) {
    __shared__ double tile[32][33];
    const int64_t i0 = (int64_t)blockIdx.x * 32;   // vectors
    const int j0 = blockIdx.y * 32;                // coordinates
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        int j = j0 + r;
        int64_t i = i0 + tx;
        tile[r][tx] = (j < d && i < n) ? cols[(size_t)j * n + i] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        int64_t i = i0 + r;
        int j = j0 + tx;
        if (i < n && j < d) rows[(size_t)i * d + j] = tile[tx][r];
    }
}

// the columns on the device, transposed into a compact row-major [n][d] buffer (caller frees with crx_free)
int rows_on_device(crx_ctx* c, const crx_columnar* f, double** out) {
    *out = nullptr;
    size_t count = (size_t)f->n * f->d;
    double* d_cols = nullptr;
    double* d_rows = nullptr;
    CRX_TRY(crx_alloc(c, &d_cols, count));
    int st = crx_alloc(c, &d_rows, count);
    if (st != CRX_OK) { crx_free(c, d_cols); return st; }
    cudaError_t e = cudaMemcpyAsync(d_cols, f->cols, count * sizeof(double), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        CRX_KERNEL(c, "columns_to_rows");
        dim3 grid((unsigned)((f->n + 31) / 32), (unsigned)((f->d + 31) / 32));
        columns_to_rows_kernel<<<grid, 256, 0, c->stream>>>(d_cols, f->n, f->d, d_rows);
        e = cudaGetLastError();
    }
    crx_free(c, d_cols);
    if (e != cudaSuccess) { crx_free(c, d_rows); CRX_CUDA(e); }
    *out = d_rows;
    return CRX_OK;
}

}  // namespace

extern "C" {

int crx_columnar_write(const char* path, const char* const* ids, const double* rows, int64_t n, int d) {
    CRX_REQUIRE(path && ids && rows && n >= 0 && d >= 1, "argument");
    FILE* o = fopen(path, "wb");
    if (!o) { crx_set_error("cannot create %s", path); return CRX_ERR_INVALID; }
    int64_t ids_bytes = 0;
    for (int64_t i = 0; i < n; i++) ids_bytes += (int64_t)strlen(ids[i]) + 1;
    int32_t d32 = d, zero = 0;
    bool ok = fwrite(MAGIC, 1, 8, o) == 8 && fwrite(&n, 8, 1, o) == 1 && fwrite(&d32, 4, 1, o) == 1 && fwrite(&zero, 4, 1, o) == 1 &&
              fwrite(&ids_bytes, 8, 1, o) == 1;
    for (int64_t i = 0; ok && i < n; i++) { size_t len = strlen(ids[i]) + 1; ok = fwrite(ids[i], 1, len, o) == len; }
    const char pad[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t padding = (size_t)((8 - ids_bytes % 8) % 8);
    ok = ok && fwrite(pad, 1, padding, o) == padding;
    std::vector<double> col((size_t)n);
    for (int j = 0; ok && j < d; j++) {
        for (int64_t i = 0; i < n; i++) col[(size_t)i] = rows[(size_t)i * d + j];
        ok = n == 0 || fwrite(col.data(), sizeof(double), (size_t)n, o) == (size_t)n;
    }
    ok = (fclose(o) == 0) && ok;
    if (!ok) { crx_set_error("short write to %s", path); return CRX_ERR_INVALID; }
    return CRX_OK;
}

int crx_columnar_open(const char* path, crx_columnar** out) {
    CRX_REQUIRE(path && out, "NULL argument");
    *out = nullptr;
    int fd = open(path, O_RDONLY);
    if (fd < 0) { crx_set_error("cannot open %s", path); return CRX_ERR_INVALID; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || (size_t)sb.st_size < HEADER) { close(fd); crx_set_error("%s: not a columnar file", path); return CRX_ERR_INVALID; }
    void* map = mmap(nullptr, (size_t)sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) { crx_set_error("mmap of %s failed", path); return CRX_ERR_INVALID; }
    const char* p = (const char*)map;
    int64_t n, ids_bytes;
    int32_t d;
    memcpy(&n, p + 8, 8); memcpy(&d, p + 16, 4); memcpy(&ids_bytes, p + 24, 8);
    size_t ids_padded = ids_bytes >= 0 ? (size_t)ids_bytes + (size_t)((8 - ids_bytes % 8) % 8) : 0;
    bool ok = memcmp(p, MAGIC, 8) == 0 && n >= 0 && d >= 1 && ids_bytes >= n &&
              (size_t)sb.st_size == HEADER + ids_padded + (size_t)n * (size_t)d * sizeof(double);
    crx_columnar* f = nullptr;
    if (ok) {
        f = new crx_columnar();
        f->map = map; f->bytes = (size_t)sb.st_size; f->n = n; f->d = d;
        f->ids.reserve((size_t)n);
        const char* s = p + HEADER;
        const char* end = s + ids_bytes;
        for (int64_t i = 0; i < n && ok; i++) {
            const char* z = (const char*)memchr(s, 0, (size_t)(end - s));
            if (!z) { ok = false; break; }
            f->ids.push_back(s);
            s = z + 1;
        }
        ok = ok && s == end;
        f->cols = (const double*)(p + HEADER + ids_padded);
    }
    if (!ok) {
        delete f;
        munmap(map, (size_t)sb.st_size);
        crx_set_error("%s: not a columnar file (bad magic, sizes or id table)", path);
        return CRX_ERR_INVALID;
    }
    *out = f;
    return CRX_OK;
}

int64_t crx_columnar_n(const crx_columnar* f) { return f ? f->n : 0; }
int32_t crx_columnar_d(const crx_columnar* f) { return f ? f->d : 0; }
const char* crx_columnar_id(const crx_columnar* f, int64_t i) { return (f && i >= 0 && i < f->n) ? f->ids[(size_t)i] : nullptr; }

int crx_columnar_points(crx_ctx* c, const crx_columnar* f, crx_points** out) {
    CRX_REQUIRE(c && f && out, "NULL argument");
    CRX_REQUIRE(f->n >= 1, "empty file");
    CRX_CUDA(cudaSetDevice(c->device));
    double* d_rows = nullptr;
    CRX_TRY(rows_on_device(c, f, &d_rows));
    int st = crx_points_create(c, d_rows, CRX_F64, f->n, f->d, CRX_DEVICE, out);
    crx_free(c, d_rows);
    return st;
}

int crx_columnar_rows(crx_ctx* c, const crx_columnar* f, double* rows) {
    CRX_REQUIRE(c && f && rows, "NULL argument");
    if (f->n == 0) return CRX_OK;
    CRX_CUDA(cudaSetDevice(c->device));
    double* d_rows = nullptr;
    CRX_TRY(rows_on_device(c, f, &d_rows));
    cudaError_t e = cudaMemcpyAsync(rows, d_rows, (size_t)f->n * f->d * sizeof(double), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    crx_free(c, d_rows);
    CRX_CUDA(e);
    return CRX_OK;
}

int crx_columnar_close(crx_columnar* f) {
    if (!f) return CRX_OK;
    if (f->map) munmap(f->map, f->bytes);
    delete f;
    return CRX_OK;
}

}  // extern "C"
