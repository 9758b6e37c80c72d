// dist_runner -- the multi-GPU sort driven from C++ through the C ABI alone (rs_dist_create: one process, all GPUs of
// the box): what a C++ host in the reference's language would write.  For uint32/int32/uint64/int64 x the five
// reference datasets (device-generated per shard) it sorts 2^log2n keys over N GPUs and checks, on the host, that the
// concatenation of the GPUs' outputs equals std::sort of the concatenated inputs (memcmp, as the reference's
// ValidateResults does, src/CRadixSortTask.cpp:225-252) and that the shards are exactly balanced.   usage: dist_runner [ngpus=all] [log2n_total=22]
#include "rsb200.h"

#include <cuda_runtime_api.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CHECK(expr)                                                                                       \
    do {                                                                                                  \
        int _rc = (expr);                                                                                 \
        if (_rc != RS_OK) {                                                                               \
            std::fprintf(stderr, "%s failed: %s (%s)\n", #expr, rs_status_string(_rc), rs_last_error_string()); \
            return 1;                                                                                     \
        }                                                                                                 \
    } while (0)

template <typename T>
static int run_type(rs_key_type type, const char* name, int ngpus, uint64_t total, int* failures) {
    const uint64_t per = total / ngpus;
    rs_dist d = nullptr;
    CHECK(rs_dist_create(ngpus, nullptr, type, per + 1024, &d));
    std::vector<rs_context> ctx(ngpus);
    std::vector<rs_stream>  st(ngpus);
    std::vector<void*>      shard(ngpus);
    for (int g = 0; g < ngpus; ++g) {
        CHECK(rs_context_create(g, &ctx[g]));
        CHECK(rs_stream_create(ctx[g], &st[g]));
        CHECK(rs_device_alloc(ctx[g], (per + 64) * sizeof(T), &shard[g]));
    }
    const char* kinds[] = {"zeros", "range", "inverted", "uniform", "random"};
    for (int kind = 0; kind < 5; ++kind) {
        std::vector<uint64_t> n_in(ngpus), n_out(ngpus);
        std::vector<void*>    out(ngpus);
        std::vector<const void*> in(ngpus);
        std::vector<T> all_in;
        for (int g = 0; g < ngpus; ++g) {
            n_in[g] = per - (kind == 3 ? static_cast<uint64_t>(g) * 7 : 0);   // ragged shards for one dataset
            CHECK(rs_dataset_fill_device(ctx[g], st[g], type, static_cast<rs_dataset_kind>(kind), 0x5EED, g * per, total, shard[g], n_in[g]));
            CHECK(rs_stream_sync(st[g]));
            in[g] = shard[g];
        }
        CHECK(rs_dist_sort(d, in.data(), n_in.data(), out.data(), n_out.data()));
        CHECK(rs_dist_sync(d));
        // exact check on the host: concatenation of the outputs == std::sort of the concatenated inputs, exact balance
        std::vector<T> all_out;
        uint64_t n_all = 0;
        for (int g = 0; g < ngpus; ++g) n_all += n_in[g];
        all_in.resize(n_all);
        all_out.resize(n_all);
        uint64_t off_in = 0, off_out = 0;
        bool balanced = true;
        for (int g = 0; g < ngpus; ++g) {
            balanced = balanced && n_out[g] == ((g + 1) * n_all) / ngpus - (g * n_all) / ngpus;
            if (off_out + n_out[g] > n_all) { balanced = false; break; }
            cudaSetDevice(g);
            cudaMemcpy(all_in.data() + off_in, shard[g], n_in[g] * sizeof(T), cudaMemcpyDeviceToHost);
            cudaMemcpy(all_out.data() + off_out, out[g], n_out[g] * sizeof(T), cudaMemcpyDeviceToHost);
            off_in += n_in[g];
            off_out += n_out[g];
        }
        std::sort(all_in.begin(), all_in.end());
        const bool ok = balanced && off_out == n_all && std::memcmp(all_in.data(), all_out.data(), n_all * sizeof(T)) == 0;
        std::printf("[rs_dist_create x%d] %s %-9s total=%llu: %s\n", ngpus, name, kinds[kind], static_cast<unsigned long long>(n_all),
                    ok ? "OK" : "FAIL");
        if (!ok) ++*failures;
    }
    double ms[RS_DIST_STAGES];
    CHECK(rs_dist_stage_ms(d, 0, ms));
    std::printf("  stages of the last sort on GPU 0 (ms): local sort %.3f, splitters %.3f, merge level 1 %.3f, local levels %.3f, closing %.3f, total %.3f\n",
                ms[0], ms[1], ms[2], ms[3], ms[4], ms[5]);
    for (int g = 0; g < ngpus; ++g) {
        CHECK(rs_device_free(ctx[g], shard[g]));
        rs_stream_destroy(st[g]);
        rs_context_destroy(ctx[g]);
    }
    CHECK(rs_dist_destroy(d));
    return 0;
}

int main(int argc, char** argv) {
    int ngpus = 0;
    CHECK(rs_device_count(&ngpus));
    if (argc > 1) ngpus = std::min(ngpus, std::atoi(argv[1]));
    const int log2n = argc > 2 ? std::atoi(argv[2]) : 22;
    if (ngpus < 1) { std::fprintf(stderr, "no CUDA device: there is no CPU fallback\n"); return 2; }
    int failures = 0;
    if (run_type<uint32_t>(RS_KEY_U32, "u32", ngpus, 1ull << log2n, &failures)) return 1;
    if (run_type<int32_t>(RS_KEY_I32, "i32", ngpus, 1ull << log2n, &failures)) return 1;
    if (run_type<uint64_t>(RS_KEY_U64, "u64", ngpus, 1ull << log2n, &failures)) return 1;
    if (run_type<int64_t>(RS_KEY_I64, "i64", ngpus, 1ull << log2n, &failures)) return 1;
    std::printf("dist_runner %s\n", failures ? "FAILED" : "PASSED");
    return failures ? 1 : 0;
}
