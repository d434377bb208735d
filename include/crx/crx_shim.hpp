// crx_shim.hpp -- glue shared by the drop-in headers under include/crx/lib/: a process-wide context, the
// seed source, and packing of std::vector<CustVector<T>> into the contiguous buffers of the C ABI
// (include/crx.h).  Host side only marshals; every arithmetic step runs in libcrx.so on the GPU.
#ifndef CRX_SHIM_HPP
#define CRX_SHIM_HPP

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstddef>
#include <cstdlib>
#include <set>
#include <string>
#include <vector>

extern "C" {
#include "../crx.h"
}

template <typename dim_type>
class CustVector;

namespace crx {

inline void check(int status, const char* what) {
    if (status != CRX_OK) {  // the reference has no error channel: fail loudly instead of computing on the CPU
        std::fprintf(stderr, "crx: %s failed (%d): %s\n", what, status, crx_last_error());
        std::abort();
    }
}

inline crx_ctx* context() {
    static crx_ctx* ctx = nullptr;
    if (!ctx) {
        const char* dev = std::getenv("CRX_DEVICE");
        auto t0 = std::chrono::steady_clock::now();
        check(crx_ctx_create(dev ? std::atoi(dev) : 0, nullptr, &ctx), "crx_ctx_create");
        if (std::getenv("CRX_SHIM_PROFILE"))
            std::fprintf(stderr, "crx shim: %-34s %9.1f ms\n", "crx_ctx_create (CUDA start-up)",
                         std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    }
    return ctx;
}

// The reference seeds every engine from system_clock::now() (lsh_cube.hpp:49; initialization.hpp:42).
// set_seed() pins the value for reproducible runs; next_seed() is what each drop-in function consumes.
inline uint64_t& pinned_seed() { static uint64_t s = 0; return s; }
inline bool& seed_is_pinned() { static bool p = false; return p; }
inline void set_seed(uint64_t s) { pinned_seed() = s; seed_is_pinned() = true; }
inline void unset_seed() { seed_is_pinned() = false; }
inline uint64_t next_seed() {
    if (seed_is_pinned()) return pinned_seed();
    return (uint64_t)std::chrono::system_clock::now().time_since_epoch().count();
}

// CRX_SHIM_PROFILE=1: wall time and call count of each drop-in function, printed to stderr at exit (how
// tools/time_main.sh explains where an unchanged main.cpp spends its time).
struct ProfRow { const char* name; double ms; long calls; };
inline std::vector<ProfRow>& prof_rows() { static std::vector<ProfRow>* r = new std::vector<ProfRow>(); return *r; }   // read by an atexit handler: never destroyed
inline bool prof_on() {
    static int on = -1;
    if (on < 0) {
        const char* e = std::getenv("CRX_SHIM_PROFILE");
        on = e && e[0] == '1';
        if (on) std::atexit([] { for (const ProfRow& r : prof_rows()) std::fprintf(stderr, "crx shim: %-34s %9.1f ms %8ld calls\n", r.name, r.ms, r.calls); });
    }
    return on == 1;
}
struct Timed {
    const char* name;
    std::chrono::steady_clock::time_point t0;
    explicit Timed(const char* n) : name(prof_on() ? n : nullptr) { if (name) t0 = std::chrono::steady_clock::now(); }
    ~Timed() {
        if (!name) return;
        double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        for (ProfRow& r : prof_rows()) if (r.name == name) { r.ms += ms; r.calls++; return; }
        prof_rows().push_back(ProfRow{name, ms, 1});
    }
};

inline int metric_code(const std::string& m) {
    if (m == "euclidean") return CRX_EUCLIDEAN;
    if (m == "cosine") return CRX_COSINE;
    std::fprintf(stderr, "crx: unknown metric_type '%s'\n", m.c_str());
    std::abort();
}

// RAII handle over crx_points built from CustVector objects
template <typename T>
struct Packed {
    crx_points* pts = nullptr;
    int64_t n = 0;
    int d = 0;
    Packed() {}
    Packed(const Packed&) = delete;
    Packed& operator=(const Packed&) = delete;
    ~Packed() { if (pts) crx_points_destroy(pts); }

    // host image of `count` vectors: coordinates, unknown flags, known means
    template <typename Getter>
    static void pack_host(int64_t count, Getter get, bool ratings, int& d, std::vector<double>& buf, std::vector<uint8_t>& unk,
                          std::vector<double>& mean) {
        d = count ? (int)get(0)->getDimNumber() : 0;
        buf.assign((size_t)count * d, 0.0);
        unk.clear(); mean.clear();
        if (ratings) { unk.assign((size_t)count * d, 0); mean.resize(count); }
        for (int64_t i = 0; i < count; i++) {
            CustVector<T>* v = get(i);
            const std::vector<T>& dims = v->crxDimsRef();
            for (int j = 0; j < d; j++) buf[(size_t)i * d + j] = (double)dims[j];
            if (ratings) {
                for (int u : v->crxUnknownRef()) if (u >= 0 && u < d) unk[(size_t)i * d + u] = 1;
                mean[i] = v->getKnownMean();
            }
        }
    }
    void upload(int64_t count, int dims, const std::vector<double>& buf, const std::vector<uint8_t>& unk, const std::vector<double>& mean) {
        n = count; d = dims;
        check(crx_points_create(context(), buf.data(), CRX_F64, n, d, CRX_HOST, &pts), "crx_points_create");
        if (!unk.empty()) check(crx_points_set_ratings(pts, unk.data(), mean.data(), CRX_HOST), "crx_points_set_ratings");
    }
    template <typename Getter>
    void build(int64_t count, Getter get, bool ratings) {
        std::vector<double> buf, mean;
        std::vector<uint8_t> unk;
        int dims = 0;
        pack_host(count, get, ratings, dims, buf, unk, mean);
        upload(count, dims, buf, unk, mean);
    }
    void from_vector(std::vector<CustVector<T> >& vecs, bool ratings = false) {
        build((int64_t)vecs.size(), [&](int64_t i) { return &vecs[i]; }, ratings);
    }
    void from_pointers(const std::vector<CustVector<T>*>& vecs, bool ratings = false) {
        build((int64_t)vecs.size(), [&](int64_t i) { return vecs[i]; }, ratings);
    }
};

// Vectors that already live on the GPU: a table set registers the caller's vector (the one its tables point into, as
// the reference's tables do, lsh_cube.hpp:70) for exactly its own lifetime.  Functions that receive pointers into such
// a vector (the neighbour lists of the recommendation loop) then pass row numbers instead of packing and uploading the
// neighbours again on every call.
//
// When its tables are deleted a registration is RETIRED, not dropped: the device copy stays for as long as no CustVector
// inside the range has been created, destroyed or written since the copy was packed (crx::touched) -- every object of the
// range is then still alive and still equal to its row.  A later loop over the same vector
// (main.cpp:205-216 and :353-373 walk user_vectors again, against other tables and against centroids) finds its users
// there (home_of) and is answered from ONE engine call over all rows instead of one per user.
struct Registered {
    const char* begin;
    const char* end;
    size_t stride;
    crx_points* pts;       // built with the rating metadata
    void* table_set;       // the crx::TableSet<T> that owns the registration (NULL once retired)
    unsigned long gen;     // identifies this packing of the range (never reused)
    int64_t n;
    bool dirty;            // a CustVector inside the range was created, destroyed or written since pts was packed
};
inline std::vector<Registered>& registry() { static std::vector<Registered>* r = new std::vector<Registered>(); return *r; }
inline void register_points(const void* base, size_t n, size_t stride, crx_points* pts, void* table_set) {
    static unsigned long generation = 0;
    registry().push_back(Registered{(const char*)base, (const char*)base + n * stride, stride, pts, table_set, ++generation, (int64_t)n, false});
}
// Called by every CustVector member that creates, destroys or can alter (or hands out a way to alter) the image the engine
// sees -- coordinates, unknown set, known mean: constructors, assignment, destructor, setters, getDimensions().
// content_epoch() moves (remembered results that were verified against the content of objects outside any registered range
// stay valid only while it stands), and a registered range that holds the object is marked dirty: its device copy is no
// longer known to be current.
inline unsigned long& content_epoch() { static unsigned long e = 1; return e; }
inline void touched(const void* self) {
    content_epoch()++;
    const char* q = (const char*)self;
    for (Registered& reg : registry()) if (q >= reg.begin && q < reg.end) reg.dirty = true;
}
inline void* table_set_of(const crx_points* pts) {
    for (const Registered& reg : registry()) if (reg.pts == pts) return reg.table_set;
    return nullptr;
}
// the tables over `pts` are gone.  Returns true when the registry took the device copy over (the caller must not destroy it).
inline bool retire_points(crx_points* pts) {
    auto& r = registry();
    // retired copies that can never be valid again go; besides the new one at most one older retired copy stays
    int retired = 0;
    for (size_t i = r.size(); i-- > 0;) {
        if (r[i].table_set || r[i].pts == pts) continue;
        if (r[i].dirty || ++retired > 1) { crx_points_destroy(r[i].pts); r.erase(r.begin() + i); }
    }
    for (size_t i = 0; i < r.size(); i++) {
        if (r[i].pts != pts) continue;
        if (!r[i].dirty) { r[i].table_set = nullptr; return true; }
        r.erase(r.begin() + i);
        return false;
    }
    return false;
}
// the registration (live or retired) whose range holds `p` and whose device copy is provably current, or NULL
inline const Registered* home_of(const void* p, size_t stride) {
    const char* q = (const char*)p;
    for (const Registered& reg : registry())
        if (q >= reg.begin && q < reg.end && reg.stride == stride && !reg.dirty && (size_t)(q - reg.begin) % stride == 0) return &reg;
    return nullptr;
}
// the registered set that holds ALL of `ptrs` (rows[] receives their row numbers), or NULL
template <typename T>
inline crx_points* registered_rows(const std::vector<CustVector<T>*>& ptrs, std::vector<int32_t>& rows) {
    if (ptrs.empty()) return nullptr;
    for (const Registered& reg : registry()) {
        const char* p0 = (const char*)ptrs[0];
        // (a dirty registration -- some vector of the range was written after the device copy was packed -- is not used: the
        // caller packs the live values instead, as the reference's tables, which hold live pointers, would see them)
        if (!reg.table_set || reg.dirty || p0 < reg.begin || p0 >= reg.end || reg.stride != sizeof(CustVector<T>)) continue;
        rows.resize(ptrs.size());
        for (size_t i = 0; i < ptrs.size(); i++) {
            const char* q = (const char*)ptrs[i];
            if (q < reg.begin || q >= reg.end) return nullptr;
            rows[i] = (int32_t)((q - reg.begin) / (ptrdiff_t)reg.stride);
        }
        return reg.pts;
    }
    return nullptr;
}

// row of `p` inside `vecs`, or -1 when the pointer does not alias an element (e.g. k_means centres)
template <typename T>
inline int32_t row_of(std::vector<CustVector<T> >& vecs, const CustVector<T>* p) {
    if (vecs.empty()) return -1;
    const CustVector<T>* b = &vecs[0];
    if (p >= b && p < b + vecs.size()) return (int32_t)(p - b);
    return -1;
}

}  // namespace crx
#endif
