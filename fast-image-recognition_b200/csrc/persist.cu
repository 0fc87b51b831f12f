// persist.cu — on-disk form of a gallery and of a DirectedEnumeration build.
//
// The reference has no persisted index: every run re-parses the text features file (db_features.cpp:44-116) and
// DirectedEnumeration's constructor re-walks 0.015·N full gallery passes (ann.cpp:270-348).  One binary file holds what
// both produce — packed fp32 rows + labels, and optionally the pivot list, the pivot-distance rows and the threshold —
// so a process can start searching after one read and one upload.  Loading goes through the same entry points a caller
// would use (fir_gallery_create, fir_dem_from_state), so a loaded index answers bit-identically to the one saved.
//
// Layout (little endian): Header | labels int32[n] | rows float[n][d] | pivots int32[np] | P float[np][n] | uint64 FNV-1a of
// everything before it.
#include "handles.hpp"
#include <cstdio>
#include <cstring>
#include <exception>
#include <memory>
#include <vector>

namespace {

struct Header {
    char magic[8];            // "FIRB200\0"
    uint32_t version;         // 1
    uint32_t metric;
    uint64_t n;
    uint32_t d;
    uint32_t n_classes;
    int64_t index_offset;
    uint32_t n_pivots;        // 0: no directed-enumeration state in the file
    float threshold;
    uint32_t reserved[6];
};
static_assert(sizeof(Header) == 72, "index header layout");

const char kMagic[8] = {'F', 'I', 'R', 'B', '2', '0', '0', '\0'};

struct Fnv {
    uint64_t h = 1469598103934665603ull;
    void add(const void* p, size_t n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    }
};

struct File {
    FILE* f = nullptr;
    ~File() { if (f) std::fclose(f); }
};

bool put(FILE* f, Fnv& sum, const void* p, size_t n) { sum.add(p, n); return std::fwrite(p, 1, n, f) == n; }
bool get(FILE* f, Fnv& sum, void* p, size_t n) { if (std::fread(p, 1, n, f) != n) return false; sum.add(p, n); return true; }

}  // namespace

using namespace fir;

extern "C" {

int fir_index_save(const fir_gallery* g, const fir_dem* dem, const char* path) {
    if (!g || !path) return fail(FIR_ERR_BAD_ARG, "null argument");
    if (dem && dem_gallery(dem) != g) return fail(FIR_ERR_BAD_ARG, "the DEM handle was built over a different gallery");
    FIR_CUDA_TRY(cudaSetDevice(g->device));
    FIR_CUDA_TRY(cudaStreamSynchronize(g->stream));
    Header h{};
    std::memcpy(h.magic, kMagic, 8);
    h.version = 1; h.metric = (uint32_t)g->metric; h.n = (uint64_t)g->n; h.d = (uint32_t)g->d; h.n_classes = (uint32_t)g->n_classes;
    h.index_offset = g->index_offset;
    int32_t np = 0, chain = 0; float thr = 0.f;
    if (dem) FIR_TRY(fir_dem_info(dem, &np, &chain, &thr));
    h.n_pivots = (uint32_t)np; h.threshold = thr;
    std::vector<int32_t> labels, pivots;
    std::vector<float> rows, P;
    try {                                                          // no exception may cross the extern "C" boundary
        labels.resize((size_t)g->n); rows.resize((size_t)g->n * g->d); pivots.resize((size_t)np); P.resize((size_t)np * g->n);
    } catch (const std::exception&) { return fail(FIR_ERR_OOM, "fir_index_save: not enough host memory to stage the index"); }
    FIR_CUDA_TRY(cudaMemcpy(labels.data(), g->labels, sizeof(int32_t) * (size_t)g->n, cudaMemcpyDeviceToHost));
    FIR_CUDA_TRY(cudaMemcpy2D(rows.data(), sizeof(float) * g->d, g->rows, sizeof(float) * g->dp, sizeof(float) * g->d, (size_t)g->n,
                              cudaMemcpyDeviceToHost));
    if (np > 0) {
        FIR_TRY(fir_dem_get_pivots(dem, pivots.data()));
        FIR_TRY(fir_dem_get_pivot_matrix(dem, P.data()));
    }
    const std::string tmp = std::string(path) + ".tmp";
    {
        File out; out.f = std::fopen(tmp.c_str(), "wb");
        if (!out.f) return fail(FIR_ERR_BAD_ARG, std::string("cannot open for writing: ") + tmp);
        Fnv sum;
        bool ok = put(out.f, sum, &h, sizeof(h)) && put(out.f, sum, labels.data(), labels.size() * 4) && put(out.f, sum, rows.data(), rows.size() * 4) &&
                  put(out.f, sum, pivots.data(), pivots.size() * 4) && put(out.f, sum, P.data(), P.size() * 4);
        const uint64_t digest = sum.h;
        ok = ok && std::fwrite(&digest, 1, 8, out.f) == 8 && std::fflush(out.f) == 0;
        if (!ok) { std::remove(tmp.c_str()); return fail(FIR_ERR_INTERNAL, std::string("short write: ") + tmp); }
    }
    if (std::rename(tmp.c_str(), path) != 0) { std::remove(tmp.c_str()); return fail(FIR_ERR_INTERNAL, std::string("cannot rename onto ") + path); }
    return FIR_OK;
}

int fir_index_load(const char* path, fir_gallery** out_gallery, fir_dem** out_dem) {
    if (!path || !out_gallery) return fail(FIR_ERR_BAD_ARG, "null argument");
    *out_gallery = nullptr;
    if (out_dem) *out_dem = nullptr;
    File in; in.f = std::fopen(path, "rb");
    if (!in.f) return fail(FIR_ERR_BAD_ARG, std::string("cannot open: ") + path);
    Fnv sum;
    Header h{};
    if (!get(in.f, sum, &h, sizeof(h))) return fail(FIR_ERR_BAD_ARG, "index file: truncated header");
    if (std::memcmp(h.magic, kMagic, 8) != 0) return fail(FIR_ERR_BAD_ARG, "index file: bad magic");
    if (h.version != 1) return fail(FIR_ERR_UNSUPPORTED, "index file: unknown version");
    // every field is range-checked before any size is computed from it: n < 2^31, d <= 2^20, at most 32 search pivots
    // (what fir_dem_from_state accepts) and a non-negative offset keep all the products below far inside 64 bits
    if (h.n == 0 || h.d == 0 || h.metric > (uint32_t)FIR_KL || h.n > 0x7fffffffull || h.d > (1u << 20) || h.n_pivots > 32 || h.n_pivots > h.n ||
        h.index_offset < 0 || (uint64_t)h.index_offset + h.n > 0x7fffffffull || h.n_classes == 0 || h.n_classes > 0x7fffffffu)
        return fail(FIR_ERR_BAD_ARG, "index file: inconsistent header");
    // the sizes must add up to the file before anything is allocated from them
    const uint64_t body = h.n * 4 + h.n * (uint64_t)h.d * 4 + (uint64_t)h.n_pivots * 4 + (uint64_t)h.n_pivots * h.n * 4;
    if (std::fseek(in.f, 0, SEEK_END) != 0) return fail(FIR_ERR_INTERNAL, "index file: seek failed");
    const long long size = std::ftell(in.f);
    if (size < 0 || (uint64_t)size != sizeof(Header) + body + 8) return fail(FIR_ERR_BAD_ARG, "index file: size does not match its header (truncated?)");
    if (std::fseek(in.f, (long)sizeof(Header), SEEK_SET) != 0) return fail(FIR_ERR_INTERNAL, "index file: seek failed");
    std::vector<int32_t> labels, pivots;
    std::vector<float> rows, P;
    try {                                                          // no exception may cross the extern "C" boundary
        labels.resize((size_t)h.n); rows.resize((size_t)h.n * h.d); pivots.resize((size_t)h.n_pivots); P.resize((size_t)h.n_pivots * h.n);
    } catch (const std::exception&) { return fail(FIR_ERR_OOM, "index file: not enough host memory for its payload"); }
    uint64_t digest = 0;
    if (!get(in.f, sum, labels.data(), labels.size() * 4) || !get(in.f, sum, rows.data(), rows.size() * 4) ||
        !get(in.f, sum, pivots.data(), pivots.size() * 4) || !get(in.f, sum, P.data(), P.size() * 4) || std::fread(&digest, 1, 8, in.f) != 8)
        return fail(FIR_ERR_BAD_ARG, "index file: truncated");
    if (digest != sum.h) return fail(FIR_ERR_BAD_ARG, "index file: checksum mismatch (corrupted)");
    for (uint32_t i = 0; i < h.n_pivots; ++i)
        if (pivots[i] < 0 || (uint64_t)pivots[i] >= h.n) return fail(FIR_ERR_BAD_ARG, "index file: pivot out of range");
    fir_gallery* g = nullptr;
    FIR_TRY(fir_gallery_create(rows.data(), labels.data(), (int64_t)h.n, (int32_t)h.d, (int32_t)h.metric, FIR_HOST, h.index_offset, &g));
    int st = fir_gallery_set_num_classes(g, (int32_t)h.n_classes);
    fir_dem* dem = nullptr;
    if (st == FIR_OK && h.n_pivots > 0 && out_dem) st = fir_dem_from_state(g, pivots.data(), (int32_t)h.n_pivots, P.data(), h.threshold, &dem);
    if (st != FIR_OK) { fir_gallery_destroy(g); return st; }
    *out_gallery = g;
    if (out_dem) *out_dem = dem;
    return FIR_OK;
}

}  // extern "C"
