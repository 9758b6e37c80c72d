// rs_dist_api.inl -- host side of the multi-GPU sort behind rs_dist_* (include/rsb200.h); included by rs_api.cu.
//
// One NVSwitch box, keys only.  Per GPU: local onesweep into a buffer pair every peer can read (peer access inside one
// process, CUDA IPC across processes) -> exact splitters (rs_dist_plan.h / rs_dist.cuh: refinement rounds on the device,
// one ncclAllReduce per round on the sort's stream) -> cut table all-gathered -> the final local step is a merge tree
// whose first level reads the peers' sorted runs straight over NVLink (exchange and merge are one kernel,
// rs_merge.cuh) -> a one-word all-reduce so that nobody overwrites a shard a peer is still reading.
// The only host synchronisation of a sort is the read-back of the cut table (the merge jobs are kernel arguments).
// NCCL is loaded at run time (dlopen) so that single-GPU users of librsb200.so need no NCCL at all.
// The reference has no counterpart (one OpenCL device, Common/ComputeState.cpp:10-12).

#include <dlfcn.h>
#include <nccl.h>

#include "rs_dist.cuh"

namespace {

struct NcclApi {
    void* lib = nullptr;
    decltype(&ncclGetUniqueId)       GetUniqueId       = nullptr;
    decltype(&ncclCommInitRank)      CommInitRank      = nullptr;
    decltype(&ncclCommInitAll)       CommInitAll       = nullptr;
    decltype(&ncclCommDestroy)       CommDestroy       = nullptr;
    decltype(&ncclAllReduce)         AllReduce         = nullptr;
    decltype(&ncclAllGather)         AllGather         = nullptr;
    decltype(&ncclGroupStart)        GroupStart        = nullptr;
    decltype(&ncclGroupEnd)          GroupEnd          = nullptr;
    decltype(&ncclCommGetAsyncError) CommGetAsyncError = nullptr;
    decltype(&ncclGetErrorString)    GetErrorString    = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {getenv("RSB200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
        for (const char* name : names) {
            if (!name) continue;
            api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (api.lib) break;
        }
        if (!api.lib) { api.error = "libnccl.so.2 not found (set RSB200_NCCL_LIB)"; return; }
        auto sym = [&](const char* n) {
            void* p = dlsym(api.lib, n);
            if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n;
            return p;
        };
        api.GetUniqueId       = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank      = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll       = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy       = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllReduce         = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.AllGather         = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.GroupStart        = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
        api.GroupEnd          = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
        api.CommGetAsyncError = reinterpret_cast<decltype(api.CommGetAsyncError)>(sym("ncclCommGetAsyncError"));
        api.GetErrorString    = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    });
    return api.error.empty() ? &api : nullptr;
}

enum DistEvent { EV_START = 0, EV_SORTED, EV_PLANNED, EV_LEVEL1, EV_MERGED, EV_CLOSED, EV_COUNT };

struct DistRank {
    rs_context   ctx        = nullptr;
    bool         own_ctx    = false;
    rs_stream    stream     = nullptr;
    bool         own_stream = false;
    rs_sorter    sorter     = nullptr;
    ncclComm_t   comm       = nullptr;
    int          rank       = 0;
    void*        pair[2]    = {nullptr, nullptr};          // local-sort ping-pong, readable by every peer
    void*        peer[DIST_MAX_WORLD][2] = {};             // every rank's pair as mapped on this device
    bool         ipc_mapped = false;
    void*        merge_buf[2] = {nullptr, nullptr};
    uint64_t*    splits     = nullptr;
    uint64_t     splits_cap = 0;
    DistPlanState*      state = nullptr;                  // device-resident refinement state
    unsigned long long* pos = nullptr, *counts = nullptr, *eq = nullptr, *eq_all = nullptr, *mine = nullptr, *table = nullptr;
    unsigned long long* h_table = nullptr;                 // pinned copy of the cut table
    int*         done_flag  = nullptr;
    cudaEvent_t  ev[EV_COUNT] = {};
    bool         ev_valid   = false;
    int          sel        = 0;
    uint64_t     n_local    = 0;
};

}  // namespace

struct rs_dist_s {
    rs_key_type type;
    KeyInfo     ki;
    int         world          = 0;
    uint64_t    cap            = 0;   // keys per GPU the buffers were sized for
    bool        single_process = false;
    std::vector<DistRank> ranks;      // the GPUs this process drives (all of them, or one)
};

namespace {

int nccl_fail(NcclApi* api, const char* what, ncclResult_t r) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, api ? api->GetErrorString(r) : "NCCL unavailable");
    return fail(RS_CALCULATION_FAILED, buf);
}
#define RS_NCCL(api, expr)                                          \
    do {                                                            \
        ncclResult_t _r = (expr);                                   \
        if (_r != ncclSuccess) return nccl_fail((api), #expr, _r);  \
    } while (0)

// One collective on every GPU this process drives (grouped, as NCCL requires for several communicators per thread).
template <typename F>
int dist_collective(rs_dist d, NcclApi* api, F&& f) {
    if (d->ranks.size() > 1) RS_NCCL(api, api->GroupStart());
    for (auto& R : d->ranks) {
        DeviceGuard  guard(R.ctx->device);
        ncclResult_t r = f(R);
        if (r != ncclSuccess) {
            if (d->ranks.size() > 1) api->GroupEnd();
            return nccl_fail(api, "NCCL collective", r);
        }
    }
    if (d->ranks.size() > 1) RS_NCCL(api, api->GroupEnd());
    return RS_OK;
}

int dist_alloc_rank(rs_dist d, DistRank& R) {
    DeviceGuard guard(R.ctx->device);
    const size_t kb    = d->ki.bytes;
    const size_t slack = 32 * static_cast<size_t>(d->world) + 64;   // intermediate runs of the merge tree start on 128-byte lines
    const int    W     = d->world;
    // Sizes are rounded to whole 2 MiB pages.  With allocations of cap * 4 + 256 bytes (a partial last large page) the
    // merge-path partition kernel -- random 32-byte probes of a peer's shard over NVLink -- took 5.32 ms at 2 GPUs; with
    // rounded sizes 2.66 ms, same kernel, same inputs (8 GPUs: 2.86 -> see profiles/README.md).  Streaming reads of the
    // same mappings did not care.  Presumably the peer mapping falls back to small pages and the probes miss the TLB.
    auto whole_pages = [](size_t bytes) { return (bytes + (size_t{2} << 20) - 1) & ~((size_t{2} << 20) - 1); };
    for (int i = 0; i < 2; ++i) {
        RS_CUDA(cudaMalloc(&R.pair[i], whole_pages(std::max<size_t>(d->cap, 1) * kb + 256)), RS_INITIALIZATION_FAILED);
        RS_CUDA(cudaMalloc(&R.merge_buf[i], whole_pages((d->cap + slack) * kb + 256)), RS_INITIALIZATION_FAILED);
    }
    R.splits_cap = (d->cap + slack) / 1024 + 4ull * W + 64;
    RS_CUDA(cudaMalloc(reinterpret_cast<void**>(&R.splits), R.splits_cap * sizeof(uint64_t)), RS_INITIALIZATION_FAILED);
    const size_t fan = 1u << distplan::LEVEL_BITS;
    RS_CUDA(cudaMalloc(reinterpret_cast<void**>(&R.state), sizeof(DistPlanState)), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaMalloc(reinterpret_cast<void**>(&R.pos), sizeof(unsigned long long) * W * (fan + 1)), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaMalloc(reinterpret_cast<void**>(&R.counts), sizeof(unsigned long long) * W * fan), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaMalloc(reinterpret_cast<void**>(&R.eq), sizeof(unsigned long long) * W), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaMalloc(reinterpret_cast<void**>(&R.eq_all), sizeof(unsigned long long) * W * W), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaMalloc(reinterpret_cast<void**>(&R.mine), sizeof(unsigned long long) * (W + 2)), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaMalloc(reinterpret_cast<void**>(&R.table), sizeof(unsigned long long) * W * (W + 2)), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaMalloc(reinterpret_cast<void**>(&R.done_flag), 256), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaMemset(R.done_flag, 0, 256), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaMemset(R.state, 0, sizeof(DistPlanState)), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&R.h_table), sizeof(unsigned long long) * W * (W + 2), cudaHostAllocDefault),
            RS_HOST_BUFFERS_FAILED);
    for (auto& e : R.ev) RS_CUDA(cudaEventCreate(&e), RS_INITIALIZATION_FAILED);
    int rc = rs_sorter_create(R.ctx, d->type, std::max<uint64_t>(d->cap, 1), RS_NO_KEY_BUFFERS | RS_NO_PROFILE, &R.sorter);
    return rc;
}

void dist_free_rank(rs_dist d, DistRank& R) {
    if (!R.ctx) return;
    DeviceGuard guard(R.ctx->device);
    if (R.stream) cudaStreamSynchronize(R.stream->st);
    if (R.ipc_mapped)
        for (int r = 0; r < d->world; ++r)
            if (r != R.rank)
                for (int i = 0; i < 2; ++i)
                    if (R.peer[r][i]) cudaIpcCloseMemHandle(R.peer[r][i]);
    if (R.comm) { if (NcclApi* api = nccl_api()) api->CommDestroy(R.comm); }
    if (R.sorter) rs_sorter_destroy(R.sorter);
    for (int i = 0; i < 2; ++i) { cudaFree(R.pair[i]); cudaFree(R.merge_buf[i]); }
    cudaFree(R.splits); cudaFree(R.state); cudaFree(R.pos); cudaFree(R.counts); cudaFree(R.eq); cudaFree(R.eq_all);
    cudaFree(R.mine); cudaFree(R.table); cudaFree(R.done_flag);
    if (R.h_table) cudaFreeHost(R.h_table);
    for (auto& e : R.ev) if (e) cudaEventDestroy(e);
    if (R.own_stream && R.stream) rs_stream_destroy(R.stream);
    if (R.own_ctx && R.ctx) rs_context_destroy(R.ctx);
    R = DistRank{};
}

// pairwise merge tree over the P runs a rank owns (radix-sort_b200/dist.py:_merge_tree was the first version of this)
template <typename KeyT>
int dist_merge_tree(rs_dist d, DistRank& R, std::vector<std::pair<const void*, uint64_t>> runs, void** out) {
    const size_t kb = sizeof(KeyT);
    int level = 0;
    if (runs.size() == 1) runs.push_back({runs[0].first, 0});  // one rank: bring the run into the output buffer
    while (runs.size() > 1) {
        char* const dst = static_cast<char*>(R.merge_buf[level & 1]);
        std::vector<std::pair<const void*, uint64_t>> next;
        std::vector<const void*> ap, bp;
        std::vector<void*>       op;
        std::vector<uint64_t>    an, bn;
        uint64_t off = 0;
        for (size_t i = 0; i + 1 < runs.size(); i += 2) {
            off = (off + 31) & ~31ull;  // runs start on a 128-byte line: 16-byte stores now, aligned loads on the next level
            void* o = dst + off * kb;
            ap.push_back(runs[i].first); an.push_back(runs[i].second);
            bp.push_back(runs[i + 1].first); bn.push_back(runs[i + 1].second);
            op.push_back(o);
            next.push_back({o, runs[i].second + runs[i + 1].second});
            off += runs[i].second + runs[i + 1].second;
        }
        if (runs.size() & 1) next.push_back(runs.back());  // odd run out: joins the next level (it may still be peer memory)
        for (size_t lo = 0; lo < ap.size(); lo += MERGE_MAX_JOBS) {
            const int k  = static_cast<int>(std::min<size_t>(MERGE_MAX_JOBS, ap.size() - lo));
            const int rc = merge_runs_typed<KeyT>(R.ctx, R.stream->st, d->ki.is_signed, k, ap.data() + lo, an.data() + lo, bp.data() + lo,
                                                  bn.data() + lo, op.data() + lo, R.splits, R.splits_cap);
            if (rc != RS_OK) return rc;
        }
        if (level == 0) RS_CUDA(cudaEventRecord(R.ev[EV_LEVEL1], R.stream->st), RS_CALCULATION_FAILED);
        runs = next;
        ++level;
    }
    *out = const_cast<void*>(runs[0].first);
    return RS_OK;
}

template <typename KeyT>
int dist_sort_typed(rs_dist d, const void* const* d_shards, const uint64_t* n_in, void** d_out, uint64_t* n_out) {
    NcclApi* api = nccl_api();
    if (!api) return fail(RS_CALCULATION_FAILED, "NCCL unavailable");
    const int  W    = d->world;
    const int  bits = static_cast<int>(sizeof(KeyT)) * 8;
    const KeyT sign = d->ki.is_signed ? static_cast<KeyT>(static_cast<KeyT>(1) << (bits - 1)) : static_cast<KeyT>(0);

    // 1. local sorts into the peer-visible pair
    for (size_t i = 0; i < d->ranks.size(); ++i) {
        DistRank& R = d->ranks[i];
        DeviceGuard guard(R.ctx->device);
        if (n_in[i] > d->cap) return fail(RS_CALCULATION_FAILED, "shard larger than max_keys_per_gpu");
        if (n_in[i] && !d_shards[i]) return fail(RS_CALCULATION_FAILED, "null shard");
        R.n_local = n_in[i];
        R.sel     = 0;
        RS_CUDA(cudaEventRecord(R.ev[EV_START], R.stream->st), RS_CALCULATION_FAILED);
        if (R.n_local) {
            int idx = 0;
            int rc  = rs_sorter_sort_device(R.sorter, R.stream, d_shards[i], R.pair[0], R.pair[1], nullptr, nullptr, nullptr, R.n_local, 0,
                                            bits, &idx);
            if (rc != RS_OK) return rc;
            R.sel = idx;
        }
        RS_CUDA(cudaEventRecord(R.ev[EV_SORTED], R.stream->st), RS_CALCULATION_FAILED);
        R.ev_valid = true;
    }
    if (W == 1) {
        DistRank& R = d->ranks[0];
        d_out[0] = R.pair[R.sel];
        n_out[0] = R.n_local;
        for (int e : {EV_PLANNED, EV_LEVEL1, EV_MERGED, EV_CLOSED}) cudaEventRecord(R.ev[e], R.stream->st);
        return RS_OK;
    }

    // 2. exact splitters: refinement rounds on the device, one all-reduce per round.  The first all-reduce also orders
    //    "every peer's local sort is complete" before any peer read below.
    const int levels = distplan::num_levels(bits);
    for (int level = 0; level < levels; ++level) {
        int width = 0, shift = 0;
        distplan::level_geometry(bits, level, &width, &shift);
        const int fan  = 1 << width;
        const int rows = level == 0 ? 1 : W - 1;
        for (auto& R : d->ranks) {
            DeviceGuard guard(R.ctx->device);
            const KeyT* keys = static_cast<const KeyT*>(R.pair[R.sel]);
            const int   nq   = rows * (fan + 1);
            rs_split_positions_kernel<KeyT><<<(nq + 127) / 128, 128, 0, R.stream->st>>>(keys, R.n_local, R.state, rows, level != 0, fan, shift,
                                                                                       bits, sign, R.pos);
            rs_split_counts_kernel<<<(rows * fan + 255) / 256, 256, 0, R.stream->st>>>(R.pos, rows, fan, R.counts);
            RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
        }
        int rc = dist_collective(d, api, [&](DistRank& R) {
            return api->AllReduce(R.counts, R.counts, static_cast<size_t>(rows) * fan, ncclUint64, ncclSum, R.comm, R.stream->st);
        });
        if (rc != RS_OK) return rc;
        for (auto& R : d->ranks) {
            DeviceGuard guard(R.ctx->device);
            rs_split_select_kernel<<<1, 32, 0, R.stream->st>>>(R.counts, level, fan, shift, W, R.state);
            RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
        }
    }
    // copies of every exact splitter value, split by count in rank order
    for (auto& R : d->ranks) {
        DeviceGuard guard(R.ctx->device);
        const KeyT* keys = static_cast<const KeyT*>(R.pair[R.sel]);
        rs_split_positions_kernel<KeyT><<<1, 64, 0, R.stream->st>>>(keys, R.n_local, R.state, W - 1, 1, 1, 0, bits, sign, R.pos);
        rs_split_equal_kernel<<<1, 32, 0, R.stream->st>>>(R.pos, W, R.eq);
        RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
    }
    int rc = dist_collective(d, api, [&](DistRank& R) {
        return api->AllGather(R.eq, R.eq_all, static_cast<size_t>(W - 1), ncclUint64, R.comm, R.stream->st);
    });
    if (rc != RS_OK) return rc;
    for (auto& R : d->ranks) {
        DeviceGuard guard(R.ctx->device);
        rs_split_cuts_kernel<<<1, 32, 0, R.stream->st>>>(R.pos, R.eq_all, R.state, W, R.rank, R.n_local, static_cast<unsigned long long>(R.sel),
                                                        R.mine);
        RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
    }
    // 3. everybody learns everybody's cut positions and buffer selector; the one host read-back of the sort
    rc = dist_collective(d, api, [&](DistRank& R) {
        return api->AllGather(R.mine, R.table, static_cast<size_t>(W + 2), ncclUint64, R.comm, R.stream->st);
    });
    if (rc != RS_OK) return rc;
    for (auto& R : d->ranks) {
        DeviceGuard guard(R.ctx->device);
        RS_CUDA(cudaMemcpyAsync(R.h_table, R.table, sizeof(unsigned long long) * W * (W + 2), cudaMemcpyDeviceToHost, R.stream->st),
                RS_CALCULATION_FAILED);
        RS_CUDA(cudaEventRecord(R.ev[EV_PLANNED], R.stream->st), RS_CALCULATION_FAILED);
    }
    for (auto& R : d->ranks) {
        DeviceGuard guard(R.ctx->device);
        RS_CUDA(cudaStreamSynchronize(R.stream->st), RS_CALCULATION_FAILED);
        ncclResult_t async = ncclSuccess;
        RS_NCCL(api, api->CommGetAsyncError(R.comm, &async));
        if (async != ncclSuccess) return nccl_fail(api, "asynchronous NCCL error", async);
    }
    // 4. final local step: merge tree over my P runs; level 1 reads the peers' memory (exchange + merge in one kernel)
    for (size_t i = 0; i < d->ranks.size(); ++i) {
        DistRank& R = d->ranks[i];
        DeviceGuard guard(R.ctx->device);
        std::vector<std::pair<const void*, uint64_t>> runs;
        uint64_t total = 0;
        for (int src = 0; src < W; ++src) {
            const unsigned long long* row = R.h_table + static_cast<size_t>(src) * (W + 2);
            const uint64_t lo = row[R.rank], hi = row[R.rank + 1];
            const int      sl = static_cast<int>(row[W + 1]) & 1;
            runs.push_back({static_cast<const char*>(R.peer[src][sl]) + lo * sizeof(KeyT), hi - lo});
            total += hi - lo;
        }
        if (total > d->cap) return fail(RS_CALCULATION_FAILED, "received more keys than max_keys_per_gpu");
        void* out = R.merge_buf[0];
        if (total) {
            rc = dist_merge_tree<KeyT>(d, R, runs, &out);
            if (rc != RS_OK) return rc;
        } else {
            RS_CUDA(cudaEventRecord(R.ev[EV_LEVEL1], R.stream->st), RS_CALCULATION_FAILED);
        }
        RS_CUDA(cudaEventRecord(R.ev[EV_MERGED], R.stream->st), RS_CALCULATION_FAILED);
        d_out[i] = out;
        n_out[i] = total;
    }
    // 5. nobody may overwrite its sorted shard (next local sort) before every rank has pulled from it
    rc = dist_collective(d, api, [&](DistRank& R) {
        return api->AllReduce(R.done_flag, R.done_flag, 1, ncclInt32, ncclSum, R.comm, R.stream->st);
    });
    if (rc != RS_OK) return rc;
    for (auto& R : d->ranks) {
        DeviceGuard guard(R.ctx->device);
        RS_CUDA(cudaEventRecord(R.ev[EV_CLOSED], R.stream->st), RS_CALCULATION_FAILED);
    }
    return RS_OK;
}

}  // namespace

extern "C" {

int rs_dist_unique_id(unsigned char id[RS_DIST_ID_BYTES]) {
    if (!id) return fail(RS_INITIALIZATION_FAILED, "null argument");
    NcclApi* api = nccl_api();
    if (!api) return fail(RS_INITIALIZATION_FAILED, "NCCL unavailable");
    static_assert(sizeof(ncclUniqueId) == RS_DIST_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    RS_NCCL(api, api->GetUniqueId(&u));
    memcpy(id, &u, sizeof u);
    return RS_OK;
}

static int dist_check_args(rs_key_type type, int world, KeyInfo* ki) {
    if (!key_info(type, ki) || ki->finf || ki->bytes < 4) return fail(RS_INITIALIZATION_FAILED, "the multi-GPU sort serves 32- and 64-bit integer keys");
    if (world < 1 || world > DIST_MAX_WORLD) return fail(RS_INITIALIZATION_FAILED, "1..16 GPUs");
    return RS_OK;
}

int rs_dist_create_rank(rs_context ctx, rs_stream st, rs_key_type type, uint64_t max_keys_per_gpu, int rank, int world,
                        const unsigned char id[RS_DIST_ID_BYTES], rs_dist* out) {
    if (!ctx || !st || !id || !out) return fail(RS_INITIALIZATION_FAILED, "null argument");
    *out = nullptr;
    KeyInfo ki{};
    if (int rc = dist_check_args(type, world, &ki); rc != RS_OK) return rc;
    if (rank < 0 || rank >= world) return fail(RS_INITIALIZATION_FAILED, "bad rank");
    NcclApi* api = nccl_api();
    if (!api) return fail(RS_INITIALIZATION_FAILED, "NCCL unavailable");
    rs_dist d = new (std::nothrow) rs_dist_s{};
    if (!d) return fail(RS_INITIALIZATION_FAILED, "out of host memory");
    d->type = type; d->ki = ki; d->world = world; d->cap = max_keys_per_gpu; d->single_process = false;
    d->ranks.resize(1);
    DistRank& R = d->ranks[0];
    R.ctx = ctx; R.stream = st; R.rank = rank;
    DeviceGuard guard(ctx->device);
    int rc = dist_alloc_rank(d, R);
    if (rc != RS_OK) { rs_dist_destroy(d); return rc; }
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    if (ncclResult_t r = api->CommInitRank(&R.comm, world, u, rank); r != ncclSuccess) { rs_dist_destroy(d); return nccl_fail(api, "ncclCommInitRank", r); }
    // map every peer's pair: CUDA IPC handles travel through an NCCL all-gather of device bytes
    struct Handles { cudaIpcMemHandle_t h[2]; };
    Handles mine{};
    Handles* d_all = nullptr;
    std::vector<Handles> all(world);
    cudaError_t e = cudaIpcGetMemHandle(&mine.h[0], R.pair[0]);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&mine.h[1], R.pair[1]);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&d_all), sizeof(Handles) * (world + 1));
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_all + world, &mine, sizeof mine, cudaMemcpyHostToDevice, st->st);
    if (e != cudaSuccess) { cudaFree(d_all); rs_dist_destroy(d); return fail(RS_INITIALIZATION_FAILED, "exporting the exchange buffers (CUDA IPC)", e); }
    ncclResult_t nr = api->AllGather(d_all + world, d_all, sizeof(Handles), ncclChar, R.comm, st->st);
    if (nr == ncclSuccess) {
        e = cudaMemcpyAsync(all.data(), d_all, sizeof(Handles) * world, cudaMemcpyDeviceToHost, st->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st->st);
    }
    cudaFree(d_all);
    if (nr != ncclSuccess) { rs_dist_destroy(d); return nccl_fail(api, "all-gather of the IPC handles", nr); }
    if (e != cudaSuccess) { rs_dist_destroy(d); return fail(RS_INITIALIZATION_FAILED, "all-gather of the IPC handles", e); }
    R.ipc_mapped = true;
    for (int r = 0; r < world; ++r)
        for (int i = 0; i < 2; ++i) {
            if (r == rank) { R.peer[r][i] = R.pair[i]; continue; }
            e = cudaIpcOpenMemHandle(&R.peer[r][i], all[r].h[i], cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) { R.peer[r][i] = nullptr; rs_dist_destroy(d); return fail(RS_INITIALIZATION_FAILED, "mapping a peer's exchange buffer (CUDA IPC)", e); }
        }
    *out = d;
    return RS_OK;
}

int rs_dist_create(int ngpus, const int* devices, rs_key_type type, uint64_t max_keys_per_gpu, rs_dist* out) {
    if (!out) return fail(RS_INITIALIZATION_FAILED, "null argument");
    *out = nullptr;
    KeyInfo ki{};
    if (int rc = dist_check_args(type, ngpus, &ki); rc != RS_OK) return rc;
    NcclApi* api = nccl_api();
    if (!api) return fail(RS_INITIALIZATION_FAILED, "NCCL unavailable");
    rs_dist d = new (std::nothrow) rs_dist_s{};
    if (!d) return fail(RS_INITIALIZATION_FAILED, "out of host memory");
    d->type = type; d->ki = ki; d->world = ngpus; d->cap = max_keys_per_gpu; d->single_process = true;
    d->ranks.resize(ngpus);
    std::vector<int> devs(ngpus);
    for (int i = 0; i < ngpus; ++i) devs[i] = devices ? devices[i] : i;
    for (int i = 0; i < ngpus; ++i) {
        DistRank& R = d->ranks[i];
        R.rank = i;
        int rc = rs_context_create(devs[i], &R.ctx);
        if (rc == RS_OK) { R.own_ctx = true; rc = rs_stream_create(R.ctx, &R.stream); }
        if (rc == RS_OK) { R.own_stream = true; rc = dist_alloc_rank(d, R); }
        if (rc != RS_OK) { rs_dist_destroy(d); return rc; }
    }
    // one process: peer access instead of IPC
    for (int i = 0; i < ngpus; ++i) {
        DeviceGuard guard(devs[i]);
        for (int j = 0; j < ngpus; ++j) {
            if (i != j) {
                cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                if (e != cudaSuccess) { rs_dist_destroy(d); return fail(RS_INITIALIZATION_FAILED, "cudaDeviceEnablePeerAccess", e); }
            }
            for (int k = 0; k < 2; ++k) d->ranks[i].peer[j][k] = d->ranks[j].pair[k];
        }
    }
    std::vector<ncclComm_t> comms(ngpus);
    if (ncclResult_t r = api->CommInitAll(comms.data(), ngpus, devs.data()); r != ncclSuccess) { rs_dist_destroy(d); return nccl_fail(api, "ncclCommInitAll", r); }
    for (int i = 0; i < ngpus; ++i) d->ranks[i].comm = comms[i];
    *out = d;
    return RS_OK;
}

int rs_dist_destroy(rs_dist d) {
    if (!d) return RS_OK;
    for (auto& R : d->ranks) dist_free_rank(d, R);
    delete d;
    return RS_OK;
}

int rs_dist_num_local(rs_dist d, int* n) {
    if (!d || !n) return fail(RS_CALCULATION_FAILED, "null argument");
    *n = static_cast<int>(d->ranks.size());
    return RS_OK;
}

int rs_dist_sort(rs_dist d, const void* const* d_shards, const uint64_t* n_in, void** d_out, uint64_t* n_out) {
    if (!d || !d_shards || !n_in || !d_out || !n_out) return fail(RS_CALCULATION_FAILED, "null argument");
    if (d->ki.bytes == 4) return dist_sort_typed<uint32_t>(d, d_shards, n_in, d_out, n_out);
    return dist_sort_typed<uint64_t>(d, d_shards, n_in, d_out, n_out);
}

int rs_dist_sync(rs_dist d) {
    if (!d) return fail(RS_CALCULATION_FAILED, "null argument");
    NcclApi* api = nccl_api();
    for (auto& R : d->ranks) {
        DeviceGuard guard(R.ctx->device);
        RS_CUDA(cudaStreamSynchronize(R.stream->st), RS_CALCULATION_FAILED);
        if (api && R.comm) {
            ncclResult_t async = ncclSuccess;
            RS_NCCL(api, api->CommGetAsyncError(R.comm, &async));
            if (async != ncclSuccess) return nccl_fail(api, "asynchronous NCCL error", async);
        }
    }
    return RS_OK;
}

int rs_dist_stage_ms(rs_dist d, int local_index, double out[RS_DIST_STAGES]) {
    if (!d || !out) return fail(RS_CALCULATION_FAILED, "null argument");
    if (local_index < 0 || local_index >= static_cast<int>(d->ranks.size())) return fail(RS_CALCULATION_FAILED, "bad local index");
    DistRank& R = d->ranks[local_index];
    for (int i = 0; i < RS_DIST_STAGES; ++i) out[i] = 0.0;
    if (!R.ev_valid) return RS_OK;
    DeviceGuard guard(R.ctx->device);
    RS_CUDA(cudaEventSynchronize(R.ev[EV_CLOSED]), RS_CALCULATION_FAILED);
    for (int i = 0; i + 1 < EV_COUNT; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, R.ev[i], R.ev[i + 1]) == cudaSuccess) out[i] = ms;
    }
    float total = 0.f;
    if (cudaEventElapsedTime(&total, R.ev[EV_START], R.ev[EV_CLOSED]) == cudaSuccess) out[RS_DIST_STAGES - 1] = total;
    cudaGetLastError();
    return RS_OK;
}

}  // extern "C"
