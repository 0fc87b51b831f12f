// sharded.cu — gallery row-sharding across the GPUs of one B200 box, behind the C-ABI.
//
// The reference is single-threaded and single-device (qt_cpp/ann.cpp:94-126 walks one std::vector<ImageInfo>); here the
// class-major gallery is cut into contiguous row shards, one per GPU, queries are replicated, every GPU runs the complete
// single-GPU pipeline on its shard (so its k results are already reference-exact, global indices through index_offset) and
// ONE exchange step follows:
//   top-k      : (dist, idx) pairs packed into 64-bit keys (ordered distance bits << 32 | global index), one ncclAllGather of
//                Q·k·8 bytes per rank, k-way merge by key = the lexicographic (dist, idx) order of the reference's scan;
//   class min  : ncclAllReduce(min) on the same packed keys, Q·C·8 bytes;
//   PNN scores : ncclAllReduce(sum) on the fp64 partial Parzen sums, Q·C·8 bytes.
// A replicated HOST query batch is uploaded in 1/G slices (each GPU over its own PCIe link) and assembled on every GPU by
// an all-gather over NVLink instead of G identical uploads.
//
// Two ways in:  fir_comm_* + fir_shard_*   one process per GPU (torchrun / MPI-style launchers; bench.py --gpus N)
//               fir_sharded_*              ONE process driving all GPUs (ncclCommInitAll) — what the C++ adapters of
//                                          include/fir_b200_compat.hpp use, so a testANN-shaped program scales to 8 GPUs.
// NCCL is bound at run time (dlopen of libnccl.so.2): libfir_b200.so has no link-time NCCL dependency and single-GPU users
// never load it.
#include "fir_common.cuh"
#include "handles.hpp"
#include "sharded.hpp"
#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <mutex>
#include <thread>

namespace fir {

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
        if (!api.handle) { api.error = std::string("NCCL library not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : "?"); return; }
        bool ok = true;
        auto sym = [&](const char* s) { void* p = dlsym(api.handle, s); if (!p) { ok = false; api.error = std::string("NCCL symbol missing: ") + s; } return p; };
        api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
        api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
        api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
        if (!ok) { dlclose(api.handle); api.handle = nullptr; }
    });
    return api.handle ? &api : nullptr;
}

}  // namespace fir

namespace fir {

int comm_take(fir_comm* c, int slot, size_t bytes, cudaStream_t s, void** out) {
    if (bytes > c->cap[slot]) {
        if (c->buf[slot]) { FIR_CUDA_TRY(cudaStreamSynchronize(s)); FIR_CUDA_TRY(cudaFree(c->buf[slot])); c->buf[slot] = nullptr; c->cap[slot] = 0; }
        const size_t want = bytes + bytes / 8 + 4096;
        FIR_CUDA_TRY(cudaMalloc(&c->buf[slot], want));
        c->cap[slot] = want;
    }
    *out = c->buf[slot];
    return FIR_OK;
}

// ---- packed keys -----------------------------------------------------------------------------------
__global__ void pack_topk_keys_kernel(const float* __restrict__ d, const int32_t* __restrict__ i, int64_t cells, unsigned long long* __restrict__ keys) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cells) return;
    const int32_t idx = i[t];
    keys[t] = idx < 0 ? ~0ull : (((unsigned long long)ordered_bits(d[t]) << 32) | (uint32_t)idx);
}
// per-class minima: (100000, -1) = "no match in this class" (fir_class_min) ↔ ~0
__global__ void unpack_keys_kernel(const unsigned long long* __restrict__ keys, int64_t cells, float empty_dist, float* __restrict__ d, int32_t* __restrict__ i) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= cells) return;
    const unsigned long long key = keys[t];
    if (key == ~0ull) { d[t] = empty_dist; i[t] = -1; }
    else { d[t] = from_ordered_bits((uint32_t)(key >> 32)); i[t] = (int32_t)(uint32_t)key; }
}

// k-way merge of `world` sorted key lists per query (gathered: [world][nq][k]); one warp per query, lane l owns list l.
__global__ void __launch_bounds__(128) merge_keys_kernel(const unsigned long long* __restrict__ gathered, int world, int64_t nq, int k,
                                                         float* __restrict__ od, int32_t* __restrict__ oi) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= nq) return;
    const unsigned long long* mine = lane < world ? gathered + ((int64_t)lane * nq + q) * k : nullptr;
    int head = 0;
    unsigned long long cur = mine ? mine[0] : ~0ull;
    for (int r = 0; r < k; ++r) {
        unsigned long long best = cur; int bl = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
            if (ov < best || (ov == best && ol < bl)) { best = ov; bl = ol; }
        }
        if (lane == 0) {
            if (best == ~0ull) { od[q * k + r] = 0.f; oi[q * k + r] = -1; }
            else { od[q * k + r] = from_ordered_bits((uint32_t)(best >> 32)); oi[q * k + r] = (int32_t)(uint32_t)best; }
        }
        if (lane == bl && best != ~0ull) { ++head; cur = head < k ? mine[head] : ~0ull; }
    }
}

// score argmax with strict '<' from -DBL_MAX ⇒ lowest class on ties (classification.cpp:217-225)
__global__ void pnn_argmax_kernel(const double* __restrict__ scores, int64_t nq, int n_classes, int32_t* __restrict__ label) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    double mx = -1.7976931348623157e308; int best = -1;
    for (int c = 0; c < n_classes; ++c) { const double v = scores[q * n_classes + c]; if (mx < v) { mx = v; best = c; } }
    label[q] = best;
}

// ---- one sharded call, cut into phases so that ONE host thread can drive every GPU of the box: phases marked [collective]
// are issued for all ranks inside ncclGroupStart/End, the others rank by rank (everything is asynchronous on the rank's
// stream until `finish`) -----------------------------------------------------------------------------
enum { OP_TOPK = 0, OP_CLASSMIN = 1, OP_PNN = 2 };
struct ShardCall {
    int op = OP_TOPK;
    fir_gallery* g = nullptr; fir_comm* c = nullptr;
    const float* q_in = nullptr; int memspace = FIR_HOST; int64_t nq = 0;
    int k = 1, path = FIR_PATH_AUTO; double var = 0; int64_t n_total = 0;
    bool write_out = true;
    void* out_a = nullptr; void* out_b = nullptr;        // topk: idx, dist   classmin: min, arg   pnn: scores, label
    // staging
    int64_t per = 0, cells = 0;
    float* qslice = nullptr; float* qfull = nullptr; const float* dq = nullptr;
    float* lf = nullptr; int32_t* li = nullptr; double* ls = nullptr;
    unsigned long long* keys = nullptr; unsigned long long* gathered = nullptr;
    float* mf = nullptr; int32_t* mi = nullptr;
};

static int shard_prepare(ShardCall& a) {
    fir_gallery* g = a.g; fir_comm* c = a.c;
    cudaStream_t s = g->stream;
    const int W = c->world, d = g->d;
    a.per = ceil_div(a.nq, W);
    a.dq = a.q_in;
    if (a.memspace == FIR_HOST) {
        FIR_TRY(comm_take(c, 0, sizeof(float) * (size_t)a.per * d, s, (void**)&a.qslice));
        FIR_TRY(comm_take(c, 1, sizeof(float) * (size_t)a.per * W * d, s, (void**)&a.qfull));
        const int64_t lo = std::min<int64_t>(a.nq, (int64_t)c->rank * a.per), hi = std::min<int64_t>(a.nq, lo + a.per);
        float* dst = W > 1 ? a.qslice : a.qfull;
        if (hi - lo < a.per) FIR_CUDA_TRY(cudaMemsetAsync(dst, 0, sizeof(float) * (size_t)a.per * d, s));
        if (hi > lo) FIR_CUDA_TRY(cudaMemcpyAsync(dst, a.q_in + lo * d, sizeof(float) * (size_t)(hi - lo) * d, cudaMemcpyHostToDevice, s));
        a.dq = a.qfull;
    }
    const int C = g->n_classes;
    a.cells = a.op == OP_TOPK ? a.nq * a.k : a.nq * C;
    if (a.op == OP_PNN) {
        FIR_TRY(comm_take(c, 2, 8 * (size_t)a.cells, s, (void**)&a.ls));
        FIR_TRY(comm_take(c, 3, 4 * (size_t)a.nq, s, (void**)&a.li));
    } else {
        FIR_TRY(comm_take(c, 2, 4 * (size_t)a.cells, s, (void**)&a.lf));
        FIR_TRY(comm_take(c, 3, 4 * (size_t)a.cells, s, (void**)&a.li));
        FIR_TRY(comm_take(c, 4, 8 * (size_t)a.cells, s, (void**)&a.keys));
        if (a.op == OP_TOPK) FIR_TRY(comm_take(c, 5, 8 * (size_t)a.cells * W, s, (void**)&a.gathered));
        FIR_TRY(comm_take(c, 6, 4 * (size_t)a.cells, s, (void**)&a.mf));
        FIR_TRY(comm_take(c, 7, 4 * (size_t)a.cells, s, (void**)&a.mi));
    }
    return FIR_OK;
}
static int shard_gather_queries(ShardCall& a) {          // [collective]
    if (a.memspace != FIR_HOST || a.c->world == 1) return FIR_OK;
    FIR_NCCL_TRY(nccl_api()->AllGather(a.qslice, a.qfull, (size_t)a.per * a.g->d, ncclFloat32, a.c->comm, a.g->stream));
    return FIR_OK;
}
static int shard_local(ShardCall& a) {
    fir_gallery* g = a.g;
    cudaStream_t s = g->stream;
    const unsigned blocks = (unsigned)ceil_div(a.cells, 256);
    if (a.op == OP_TOPK) {
        FIR_TRY(fir_search_topk(g, a.dq, a.nq, a.k, 0, a.path, FIR_DEVICE, a.li, a.lf));
        pack_topk_keys_kernel<<<blocks, 256, 0, s>>>(a.lf, a.li, a.cells, a.keys);
    } else if (a.op == OP_CLASSMIN) {
        FIR_TRY(fir_class_min(g, a.dq, a.nq, FIR_DEVICE, a.lf, a.li));
        pack_topk_keys_kernel<<<blocks, 256, 0, s>>>(a.lf, a.li, a.cells, a.keys);      // arg = -1 (no match) packs to ~0
    } else {
        FIR_TRY(fir_pnn_scores(g, a.dq, a.nq, a.var, a.n_total, FIR_DEVICE, a.ls, a.li));
    }
    FIR_CUDA_TRY(cudaGetLastError());
    return FIR_OK;
}
static int shard_exchange(ShardCall& a) {                // [collective]
    NcclApi* N = nccl_api();
    fir_comm* c = a.c;
    if (c->world == 1) { if (a.op == OP_TOPK) a.gathered = a.keys; return FIR_OK; }
    if (a.op == OP_TOPK) FIR_NCCL_TRY(N->AllGather(a.keys, a.gathered, (size_t)a.cells, ncclUint64, c->comm, a.g->stream));
    else if (a.op == OP_CLASSMIN) FIR_NCCL_TRY(N->AllReduce(a.keys, a.keys, (size_t)a.cells, ncclUint64, ncclMin, c->comm, a.g->stream));
    else FIR_NCCL_TRY(N->AllReduce(a.ls, a.ls, (size_t)a.cells, ncclFloat64, ncclSum, c->comm, a.g->stream));
    return FIR_OK;
}
static int shard_finish(ShardCall& a) {
    if (!a.write_out) return FIR_OK;
    fir_gallery* g = a.g;
    cudaStream_t s = g->stream;
    const bool host = a.memspace == FIR_HOST;
    if (a.op == OP_PNN) {
        pnn_argmax_kernel<<<(unsigned)ceil_div(a.nq, 128), 128, 0, s>>>(a.ls, a.nq, g->n_classes, host ? a.li : (a.out_b ? (int32_t*)a.out_b : a.li));
        FIR_CUDA_TRY(cudaGetLastError());
        if (host) {
            if (a.out_a) FIR_CUDA_TRY(cudaMemcpyAsync(a.out_a, a.ls, 8 * (size_t)a.cells, cudaMemcpyDeviceToHost, s));
            if (a.out_b) FIR_CUDA_TRY(cudaMemcpyAsync(a.out_b, a.li, 4 * (size_t)a.nq, cudaMemcpyDeviceToHost, s));
        } else if (a.out_a) FIR_CUDA_TRY(cudaMemcpyAsync(a.out_a, a.ls, 8 * (size_t)a.cells, cudaMemcpyDeviceToDevice, s));
        return FIR_OK;
    }
    // topk: out_a = idx, out_b = dist (may be NULL);  classmin: out_a = min, out_b = arg
    float* df = host ? a.mf : (a.op == OP_TOPK ? (a.out_b ? (float*)a.out_b : a.mf) : (float*)a.out_a);
    int32_t* di = host ? a.mi : (a.op == OP_TOPK ? (int32_t*)a.out_a : (int32_t*)a.out_b);
    if (a.op == OP_TOPK) merge_keys_kernel<<<(unsigned)ceil_div(a.nq, 4), 128, 0, s>>>(a.gathered, a.c->world, a.nq, a.k, df, di);
    else unpack_keys_kernel<<<(unsigned)ceil_div(a.cells, 256), 256, 0, s>>>(a.keys, a.cells, 100000.0f, df, di);
    FIR_CUDA_TRY(cudaGetLastError());
    if (host) {
        void* hf = a.op == OP_TOPK ? a.out_b : a.out_a; void* hi = a.op == OP_TOPK ? a.out_a : a.out_b;
        if (hf) FIR_CUDA_TRY(cudaMemcpyAsync(hf, df, 4 * (size_t)a.cells, cudaMemcpyDeviceToHost, s));
        if (hi) FIR_CUDA_TRY(cudaMemcpyAsync(hi, di, 4 * (size_t)a.cells, cudaMemcpyDeviceToHost, s));
    }
    return FIR_OK;
}

static int check_call(fir_gallery* g, fir_comm* c, const float* q, int64_t nq, int memspace) {
    if (!g || !c) return fail(FIR_ERR_BAD_ARG, "gallery / communicator is null");
    if (nq < 0 || (nq > 0 && !q)) return fail(FIR_ERR_BAD_ARG, "bad query pointer");
    if (memspace != FIR_HOST && memspace != FIR_DEVICE) return fail(FIR_ERR_BAD_ARG, "bad memspace");
    if (c->world > 32) return fail(FIR_ERR_UNSUPPORTED, "more than 32 shards");
    if (c->world > 1 && !nccl_api()) return fail(FIR_ERR_NCCL, "NCCL unavailable");
    return FIR_OK;
}

// one rank, all phases (one process per GPU)
static int run_rank(ShardCall& a) {
    FIR_CUDA_TRY(cudaSetDevice(a.g->device));
    FIR_TRY(shard_prepare(a));
    FIR_TRY(shard_gather_queries(a));
    FIR_TRY(shard_local(a));
    FIR_TRY(shard_exchange(a));
    FIR_TRY(shard_finish(a));
    if (a.memspace == FIR_HOST) FIR_CUDA_TRY(cudaStreamSynchronize(a.g->stream));
    return FIR_OK;
}

}  // namespace fir

using namespace fir;

// the whole sharded gallery in one process
struct fir_sharded {
    int n_gpus = 0;
    int64_t n = 0; int d = 0, metric = 0, n_classes = 1;
    std::vector<int> devices;
    std::vector<fir_gallery*> shards;
    std::vector<fir_comm*> comms;
    std::vector<cudaStream_t> streams;
    std::vector<int64_t> lo;
};

static int run_all(fir_sharded* s, std::vector<ShardCall>& calls) {
    NcclApi* N = s->n_gpus > 1 ? nccl_api() : nullptr;
    if (s->n_gpus > 1 && !N) return fail(FIR_ERR_NCCL, "NCCL unavailable");
    auto each = [&](int (*phase)(ShardCall&), bool collective) -> int {
        int st = FIR_OK;
        if (collective && N) FIR_NCCL_TRY(N->GroupStart());
        for (int r = 0; r < s->n_gpus && st == FIR_OK; ++r) {
            if (cudaSetDevice(s->devices[r]) != cudaSuccess) { st = fail(FIR_ERR_CUDA, "cudaSetDevice failed"); break; }
            st = phase(calls[r]);
        }
        if (collective && N) { ncclResult_t e = N->GroupEnd(); if (st == FIR_OK && e != ncclSuccess) st = fail(FIR_ERR_NCCL, std::string("ncclGroupEnd: ") + N->GetErrorString(e)); }
        return st;
    };
    FIR_TRY(each(shard_prepare, false));
    FIR_TRY(each(shard_gather_queries, true));
    FIR_TRY(each(shard_local, false));
    FIR_TRY(each(shard_exchange, true));
    FIR_TRY(each(shard_finish, false));
    for (int r = 0; r < s->n_gpus; ++r) {                 // every rank's stream is drained: staging buffers are reusable afterwards
        FIR_CUDA_TRY(cudaSetDevice(s->devices[r]));
        FIR_CUDA_TRY(cudaStreamSynchronize(s->streams[r]));
    }
    return FIR_OK;
}

// ---- directed enumeration over the shards of ONE process: one host thread per GPU runs the rank-level collective calls
// (fir_shard_dem_build / fir_dem_search on the shard's handle); the communicators of ncclCommInitAll are used one per thread.
struct fir_sharded_dem {
    fir_sharded* s = nullptr;
    std::vector<fir_dem*> dems;
};

template <class F>
static int on_every_shard(fir_sharded* s, F f) {
    std::vector<int> st((size_t)s->n_gpus, FIR_OK);
    std::vector<std::string> msg((size_t)s->n_gpus);
    std::vector<std::thread> th;
    for (int r = 0; r < s->n_gpus; ++r)
        th.emplace_back([&, r] {
            if (cudaSetDevice(s->devices[r]) != cudaSuccess) { st[r] = FIR_ERR_CUDA; msg[r] = "cudaSetDevice failed"; return; }
            st[r] = f(r);
            if (st[r] != FIR_OK) msg[r] = fir_last_error_string();          // the message is thread-local: carry it to the caller's thread
        });
    for (auto& t : th) t.join();
    for (int r = 0; r < s->n_gpus; ++r)
        if (st[r] != FIR_OK) return fail(st[r], "shard " + std::to_string(r) + ": " + msg[r]);
    return FIR_OK;
}

extern "C" {

int fir_comm_unique_id(void* id_out) {
    if (!id_out) return fail(FIR_ERR_BAD_ARG, "id_out is null");
    NcclApi* N = nccl_api();
    if (!N) return fail(FIR_ERR_NCCL, "NCCL unavailable");
    static_assert(sizeof(ncclUniqueId) == FIR_COMM_ID_BYTES, "FIR_COMM_ID_BYTES must match ncclUniqueId");
    ncclUniqueId id;
    FIR_NCCL_TRY(N->GetUniqueId(&id));
    std::memcpy(id_out, &id, sizeof(id));
    return FIR_OK;
}

int fir_comm_init_rank(const void* id, int32_t rank, int32_t world, fir_comm** out) {
    if (!out) return fail(FIR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(FIR_ERR_BAD_ARG, "bad rank / world");
    fir_comm* c = new fir_comm();
    c->rank = rank; c->world = world;
    if (cudaGetDevice(&c->device) != cudaSuccess) { delete c; return fail(FIR_ERR_CUDA, "no CUDA device"); }
    if (world > 1) {
        NcclApi* N = nccl_api();
        if (!N) { delete c; return fail(FIR_ERR_NCCL, "NCCL unavailable"); }
        if (!id) { delete c; return fail(FIR_ERR_BAD_ARG, "id is null"); }
        ncclUniqueId uid;
        std::memcpy(&uid, id, sizeof(uid));
        ncclResult_t e = N->CommInitRank(&c->comm, world, uid, rank);
        if (e != ncclSuccess) { delete c; return fail(FIR_ERR_NCCL, std::string("ncclCommInitRank: ") + N->GetErrorString(e)); }
    }
    *out = c;
    return FIR_OK;
}

int fir_comm_destroy(fir_comm* c) {
    if (!c) return FIR_OK;
    cudaSetDevice(c->device);
    for (int i = 0; i < fir_comm::SLOTS; ++i) if (c->buf[i]) cudaFree(c->buf[i]);
    if (c->comm && c->owns_comm && nccl_api()) nccl_api()->CommDestroy(c->comm);
    delete c;
    return FIR_OK;
}

int fir_comm_info(const fir_comm* c, int32_t* rank, int32_t* world, int32_t* nccl_version) {
    if (!c) return fail(FIR_ERR_BAD_ARG, "communicator is null");
    if (rank) *rank = c->rank;
    if (world) *world = c->world;
    if (nccl_version) { int v = 0; if (nccl_api()) nccl_api()->GetVersion(&v); *nccl_version = v; }
    return FIR_OK;
}

int fir_shard_search_topk(fir_gallery* g, fir_comm* c, const float* queries, int64_t nq, int32_t k, int32_t path, int32_t memspace,
                          int32_t* out_idx, float* out_dist) {
    FIR_TRY(check_call(g, c, queries, nq, memspace));
    if (nq > 0 && !out_idx) return fail(FIR_ERR_BAD_ARG, "out_idx is null");
    if (k < 1 || k > 1024) return fail(FIR_ERR_BAD_ARG, "k must be in [1,1024]");
    if (nq == 0) return FIR_OK;
    ShardCall a; a.op = OP_TOPK; a.g = g; a.c = c; a.q_in = queries; a.memspace = memspace; a.nq = nq; a.k = k; a.path = path;
    a.out_a = out_idx; a.out_b = out_dist;
    return run_rank(a);
}

int fir_shard_class_min(fir_gallery* g, fir_comm* c, const float* queries, int64_t nq, int32_t memspace, float* out_min, int32_t* out_arg) {
    FIR_TRY(check_call(g, c, queries, nq, memspace));
    if (nq > 0 && (!out_min || !out_arg)) return fail(FIR_ERR_BAD_ARG, "null outputs");
    if (nq == 0) return FIR_OK;
    ShardCall a; a.op = OP_CLASSMIN; a.g = g; a.c = c; a.q_in = queries; a.memspace = memspace; a.nq = nq; a.out_a = out_min; a.out_b = out_arg;
    return run_rank(a);
}

int fir_shard_pnn_scores(fir_gallery* g, fir_comm* c, const float* queries, int64_t nq, double var, int64_t n_total, int32_t memspace,
                         double* out_scores, int32_t* out_label) {
    FIR_TRY(check_call(g, c, queries, nq, memspace));
    if (!(var > 0)) return fail(FIR_ERR_BAD_ARG, "var must be > 0");
    if (n_total <= 0) return fail(FIR_ERR_BAD_ARG, "n_total (global gallery rows) must be given for a sharded gallery");
    if (nq == 0) return FIR_OK;
    ShardCall a; a.op = OP_PNN; a.g = g; a.c = c; a.q_in = queries; a.memspace = memspace; a.nq = nq; a.var = var; a.n_total = n_total;
    a.out_a = out_scores; a.out_b = out_label;
    return run_rank(a);
}

// ---- single process, all GPUs ------------------------------------------------------------------------
int fir_sharded_destroy(fir_sharded* s) {
    if (!s) return FIR_OK;
    for (size_t r = 0; r < s->shards.size(); ++r) {
        cudaSetDevice(s->devices[r]);
        if (s->shards[r]) fir_gallery_destroy(s->shards[r]);
        if (r < s->comms.size() && s->comms[r]) fir_comm_destroy(s->comms[r]);
        if (r < s->streams.size() && s->streams[r]) cudaStreamDestroy(s->streams[r]);
    }
    delete s;
    return FIR_OK;
}

int fir_sharded_create(const float* rows, const int32_t* labels, int64_t n, int32_t d, int32_t metric, int32_t n_gpus, fir_sharded** out) {
    if (!out) return fail(FIR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (!rows || n <= 0 || d <= 0) return fail(FIR_ERR_BAD_ARG, "empty gallery");
    int avail = 0;
    FIR_CUDA_TRY(cudaGetDeviceCount(&avail));
    if (n_gpus <= 0) n_gpus = avail;
    if (n_gpus > avail) return fail(FIR_ERR_BAD_ARG, "n_gpus exceeds the visible devices");
    n_gpus = (int)std::min<int64_t>(n_gpus, n);
    int prev = 0;
    cudaGetDevice(&prev);
    fir_sharded* s = new fir_sharded();
    s->n_gpus = n_gpus; s->n = n; s->d = d; s->metric = metric;
    auto bail = [&](int code) { fir_sharded_destroy(s); cudaSetDevice(prev); return code; };
    int32_t n_classes = 1;
    if (labels) for (int64_t i = 0; i < n; ++i) n_classes = std::max(n_classes, labels[i] + 1);
    s->n_classes = n_classes;
    std::vector<ncclComm_t> comms((size_t)n_gpus, nullptr);
    for (int r = 0; r < n_gpus; ++r) s->devices.push_back(r);
    if (n_gpus > 1) {
        NcclApi* N = nccl_api();
        if (!N) return bail(fail(FIR_ERR_NCCL, "NCCL unavailable"));
        ncclResult_t e = N->CommInitAll(comms.data(), n_gpus, s->devices.data());
        if (e != ncclSuccess) return bail(fail(FIR_ERR_NCCL, std::string("ncclCommInitAll: ") + N->GetErrorString(e)));
    }
    for (int r = 0; r < n_gpus; ++r) {
        if (cudaSetDevice(r) != cudaSuccess) return bail(fail(FIR_ERR_CUDA, "cudaSetDevice failed"));
        const int64_t lo = n * r / n_gpus, hi = n * (r + 1) / n_gpus;
        fir_gallery* g = nullptr;
        int st = fir_gallery_create(rows + lo * d, labels ? labels + lo : nullptr, hi - lo, d, metric, FIR_HOST, lo, &g);
        if (st != FIR_OK) return bail(st);
        s->shards.push_back(g);
        s->lo.push_back(lo);
        cudaStream_t stream = nullptr;
        if (cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) return bail(fail(FIR_ERR_CUDA, "cudaStreamCreate failed"));
        s->streams.push_back(stream);
        fir_gallery_set_stream(g, stream);
        if (n_classes > g->n_classes) fir_gallery_set_num_classes(g, n_classes);
        fir_comm* c = new fir_comm();
        c->rank = r; c->world = n_gpus; c->device = r; c->comm = comms[r];
        s->comms.push_back(c);
    }
    cudaSetDevice(prev);
    *out = s;
    return FIR_OK;
}

int fir_sharded_info(const fir_sharded* s, int32_t* n_gpus, int64_t* n, int32_t* d, int32_t* n_classes) {
    if (!s) return fail(FIR_ERR_BAD_ARG, "sharded gallery is null");
    if (n_gpus) *n_gpus = s->n_gpus;
    if (n) *n = s->n;
    if (d) *d = s->d;
    if (n_classes) *n_classes = s->n_classes;
    return FIR_OK;
}

int fir_sharded_shard(fir_sharded* s, int32_t rank, fir_gallery** shard, fir_comm** comm) {
    if (!s || rank < 0 || rank >= s->n_gpus) return fail(FIR_ERR_BAD_ARG, "bad shard index");
    if (shard) *shard = s->shards[rank];
    if (comm) *comm = s->comms[rank];
    return FIR_OK;
}

static int sharded_run(fir_sharded* s, ShardCall proto) {
    if (!s) return fail(FIR_ERR_BAD_ARG, "sharded gallery is null");
    if (proto.nq < 0 || (proto.nq > 0 && !proto.q_in)) return fail(FIR_ERR_BAD_ARG, "bad query pointer");
    if (proto.nq == 0) return FIR_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    std::vector<ShardCall> calls((size_t)s->n_gpus, proto);
    for (int r = 0; r < s->n_gpus; ++r) { calls[r].g = s->shards[r]; calls[r].c = s->comms[r]; calls[r].write_out = r == 0; }
    const int st = run_all(s, calls);
    cudaSetDevice(prev);
    return st;
}

int fir_sharded_search_topk(fir_sharded* s, const float* queries, int64_t nq, int32_t k, int32_t path, int32_t* out_idx, float* out_dist) {
    if (nq > 0 && !out_idx) return fail(FIR_ERR_BAD_ARG, "out_idx is null");
    if (k < 1 || k > 1024) return fail(FIR_ERR_BAD_ARG, "k must be in [1,1024]");
    ShardCall a; a.op = OP_TOPK; a.q_in = queries; a.memspace = FIR_HOST; a.nq = nq; a.k = k; a.path = path; a.out_a = out_idx; a.out_b = out_dist;
    return sharded_run(s, a);
}

int fir_sharded_class_min(fir_sharded* s, const float* queries, int64_t nq, float* out_min, int32_t* out_arg) {
    if (nq > 0 && (!out_min || !out_arg)) return fail(FIR_ERR_BAD_ARG, "null outputs");
    ShardCall a; a.op = OP_CLASSMIN; a.q_in = queries; a.memspace = FIR_HOST; a.nq = nq; a.out_a = out_min; a.out_b = out_arg;
    return sharded_run(s, a);
}

int fir_sharded_pnn_scores(fir_sharded* s, const float* queries, int64_t nq, double var, double* out_scores, int32_t* out_label) {
    if (!(var > 0)) return fail(FIR_ERR_BAD_ARG, "var must be > 0");
    ShardCall a; a.op = OP_PNN; a.q_in = queries; a.memspace = FIR_HOST; a.nq = nq; a.var = var; a.n_total = s ? s->n : 0; a.out_a = out_scores; a.out_b = out_label;
    return sharded_run(s, a);
}

int fir_sharded_dem_destroy(fir_sharded_dem* d) {
    if (!d) return FIR_OK;
    for (size_t r = 0; r < d->dems.size(); ++r) { cudaSetDevice(d->s->devices[r]); fir_dem_destroy(d->dems[r]); }
    delete d;
    return FIR_OK;
}

int fir_sharded_dem_build(fir_sharded* s, const fir_dem_params* params, fir_sharded_dem** out) {
    if (!out) return fail(FIR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    if (!s || !params) return fail(FIR_ERR_BAD_ARG, "null argument");
    int prev = 0;
    cudaGetDevice(&prev);
    fir_sharded_dem* d = new fir_sharded_dem();
    d->s = s;
    d->dems.assign((size_t)s->n_gpus, nullptr);
    const int st = on_every_shard(s, [&](int r) { return fir_shard_dem_build(s->shards[r], s->comms[r], s->n, params, &d->dems[r]); });
    cudaSetDevice(prev);
    if (st != FIR_OK) { fir_sharded_dem_destroy(d); return st; }
    *out = d;
    return FIR_OK;
}

int fir_sharded_dem_info(const fir_sharded_dem* d, int32_t* n_pivots, int32_t* chain_rows, float* threshold) {
    if (!d || d->dems.empty()) return fail(FIR_ERR_BAD_ARG, "sharded dem is null");
    return fir_dem_info(d->dems[0], n_pivots, chain_rows, threshold);
}

int fir_sharded_dem_get_pivots(const fir_sharded_dem* d, int32_t* out_pivots) {
    if (!d || d->dems.empty()) return fail(FIR_ERR_BAD_ARG, "sharded dem is null");
    return fir_dem_get_pivots(d->dems[0], out_pivots);
}

int fir_sharded_dem_search(fir_sharded_dem* d, const float* queries, int64_t nq, int32_t count_to_check, int32_t* out_idx, float* out_dist,
                           uint8_t* out_below, int32_t* out_evals) {
    if (!d) return fail(FIR_ERR_BAD_ARG, "sharded dem is null");
    if (nq < 0 || (nq > 0 && (!queries || !out_idx))) return fail(FIR_ERR_BAD_ARG, "bad arguments");
    if (nq == 0) return FIR_OK;
    int prev = 0;
    cudaGetDevice(&prev);
    fir_sharded* s = d->s;
    std::vector<std::vector<int32_t> > scratch((size_t)s->n_gpus);           // every rank receives the answer; rank 0 writes the caller's buffers
    const int st = on_every_shard(s, [&](int r) {
        if (r == 0) return fir_dem_search(d->dems[0], queries, nq, count_to_check, FIR_HOST, out_idx, out_dist, out_below, out_evals);
        scratch[r].resize((size_t)nq);
        return fir_dem_search(d->dems[r], queries, nq, count_to_check, FIR_HOST, scratch[r].data(), nullptr, nullptr, nullptr);
    });
    cudaSetDevice(prev);
    return st;
}

}  // extern "C"
