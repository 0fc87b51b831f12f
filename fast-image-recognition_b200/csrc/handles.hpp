// handles.hpp — the opaque handle types behind include/fir_b200.h.
#pragma once
#include "fir_common.cuh"

// replaces std::vector<ImageInfo> dbImages + the ImagesDatabase it borrows (qt_cpp/db_features.h:14-29):
// packed, zero-padded, device-resident, owned by the handle.
struct fir_gallery {
    int device = 0, n_sm = 148, cc_major = 0;
    cudaStream_t stream = 0;
    int64_t n = 0, index_offset = 0;
    int d = 0, dp = 0, metric = 0, n_classes = 1;
    float* rows = nullptr;        // [n][dp] fp32, zero padded
    int32_t* labels = nullptr;    // [n]
    std::vector<int32_t> h_labels;
    // tensor-core L2 path (built lazily on first use)
    bool tensor_ready = false;
    void* tensor_buf = nullptr;
    fir::TensorSide tside;
    CUtensorMap tmap_b;           // box 64 x 256 rows (single-CTA kernel)
    CUtensorMap tmap_b_half;      // box 64 x 128 rows (each CTA of a pair loads half a tile)
    const float* tensor_center = nullptr;          // optional [dp] (device, owned by the creator): the shadows are of x − center
    const unsigned char* tensor_exclude = nullptr; // optional [n] (device): rows that never become tensor-path candidates
    // prefix distances on the tensor path: row norms over the first prefix_d dimensions (shadow order), rebuilt when the prefix changes
    float* prefix_norm2 = nullptr; int prefix_d = 0; int tensor_d_eff = 0;
    // natural-order (class-major) shadow for the per-class reductions on tensor cores (built lazily by tensor_class_min)
    bool nat_ready = false; int nat_classes = 0;
    void* tensor_buf_nat = nullptr;
    fir::TensorSide tside_nat;
    CUtensorMap tmap_nat_half;
    int32_t* d_cls_begin = nullptr;   // [nat_classes + 1]
    float* d_stats = nullptr;     // [2] max ||x||, max ||x - fp16(x)||
    float* d_l1max = nullptr;     // chi2/KL approximate path: [0] max ||x||_1, [4] flagged count, [5] max bound
    double* kl_ent = nullptr;     // KL entropy form: Σ x·ln x + ln2·Σ x per gallery row (set with d_l1max)
    bool has_negative = false;    // set with d_l1max: some gallery element is < 0 (KL's approximate error model then does not apply)
    fir::Workspace ws;
    // diagnostics: where the last tensor-path call left its candidate lists (valid until the next call)
    const float* dbg_cand_val = nullptr; const float* dbg_cand_exact = nullptr; const int32_t* dbg_cand_idx = nullptr;
    int64_t dbg_nq = 0; int dbg_slots = 0, dbg_R = 0; uint64_t dbg_generation = 0;
    // optional per-kernel timing: CUDA event pairs recorded on the launching stream around the dominant kernel
    bool profiling = false;
    struct EvPair { cudaEvent_t a, b; int kind; };
    std::vector<EvPair> ev_pool; size_t ev_used = 0;
    EvPair* prof_begin(int kind) {
        if (!profiling) return nullptr;
        if (ev_used == ev_pool.size()) { EvPair e{}; cudaEventCreate(&e.a); cudaEventCreate(&e.b); ev_pool.push_back(e); }
        EvPair* e = &ev_pool[ev_used++]; e->kind = kind; cudaEventRecord(e->a, stream); return e;
    }
    void prof_end(EvPair* e) { if (e) cudaEventRecord(e->b, stream); }
    fir_search_stats stats{};
};

namespace fir {
int exact_topk_device(fir_gallery* g, const float* dq, int64_t nq, int k, int d_end, const int32_t* qmap,
                      const int32_t* n_active, float* part_d, int32_t* part_i, int nsplit, float* od, int32_t* oi,
                      int64_t active_offset = 0, int64_t active_cap = 0);
int tensor_search_topk(fir_gallery* g, const float* queries, int64_t nq, int k, int memspace, int32_t* out_idx, float* out_dist, int d_end = 0);
constexpr int kApproxDeclined = -100;   // approx_search_topk: the error model does not cover this gallery — the caller takes the exact path
int approx_search_topk(fir_gallery* g, const float* queries, int64_t nq, int k, int memspace, int32_t* out_idx, float* out_dist);
// per-class nearest neighbour (L2) through the tensor-core candidate kernel; kApproxDeclined when the gallery is not class-major
int tensor_class_min(fir_gallery* g, const float* queries, int64_t nq, int memspace, float* out_min, int32_t* out_arg);
const fir_gallery* dem_gallery(const fir_dem* dem);      // the gallery a DEM handle was built over
size_t twd_workspace_bytes(const fir_gallery* g, int64_t nq, int64_t* mq_out);
int twd_run(fir_gallery* g, const float* dq, int64_t nq, int64_t mq, int kind, int type, double threshold, int feat_count, int last_feature,
            int32_t* d_index, int32_t* d_label, unsigned char* d_unrel);
}
