/*
 * rsb200.h -- C ABI of the B200-native radix sort (librsb200.so).
 *
 * This is the drop-in boundary for the GPU path of gyatskov/radix-sort.  Every entry
 * point names the reference interface it replaces (paths relative to the reference
 * checkout).  The reference drives an OpenCL context/queue/buffer layer from
 * RadixSortGPU<T>; here the same steps are plain C calls over CUDA:
 *
 *   reference                                              this ABI
 *   ---------------------------------------------------    --------------------------------
 *   ComputeState::init           Common/ComputeState.cpp:14  rs_context_create / rs_stream_create
 *   RadixSortGPU::initialize     src/RadixSortGPU.cpp:451    rs_sorter_create
 *   ComputeDeviceData ctor       src/ComputeDeviceData.cpp:10  (device buffers owned by rs_sorter)
 *   RadixSortGPU::Resize         src/RadixSortGPU.cpp:287    rs_round_up_keys
 *   RadixSortGPU::padGPUData     src/RadixSortGPU.cpp:269    rs_sorter_pad
 *   RadixSortGPU::uploadData     src/RadixSortGPU.cpp:299    rs_sorter_upload
 *   RadixSortGPU::calculate      src/RadixSortGPU.cpp:310    rs_sorter_sort
 *   RadixSortGPU::downloadData   src/RadixSortGPU.cpp:348    rs_sorter_download (+ _download_histograms)
 *   RadixSortGPU::getRuntimes    src/RadixSortGPU.cpp:591    rs_sorter_runtimes
 *   RadixSortGPU::release        src/RadixSortGPU.cpp:444    rs_sorter_destroy
 *   OperationStatus              src/OperationStatus.h:4     rs_status (same ordinals 0..11)
 *   TypeNameString<T>            Common/CLTypeInformation.h:8  rs_key_type
 *
 * Conventions: plain pointers and sizes only; no exception crosses the boundary; every
 * function returns an rs_status ordinal (RS_OK == 0) unless documented otherwise.  All
 * device work is stream-ordered on the rs_stream given; the only blocking calls are
 * rs_stream_sync, rs_sorter_download, rs_sorter_runtimes and the *_result readers.
 * A sorter may be used from one host thread at a time; distinct sorters are independent.
 *
 * There is no CPU fallback anywhere behind this ABI: without a CUDA device every call
 * that needs one returns RS_INITIALIZATION_FAILED.
 */
#ifndef RSB200_H
#define RSB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSB200_VERSION_MAJOR 0
#define RSB200_VERSION_MINOR 1

/* Same ordinals as the reference's enum class OperationStatus (src/OperationStatus.h:4-17). */
typedef enum rs_status {
    RS_OK = 0,
    RS_HOST_BUFFERS_FAILED = 1,
    RS_INITIALIZATION_FAILED = 2,
    RS_DATA_UPLOAD_FAILED = 3,
    RS_CALCULATION_FAILED = 4,
    RS_DATA_DOWNLOAD_FAILED = 5,
    RS_CLEANUP_FAILED = 6,
    RS_RESIZE_FAILED = 7,
    RS_KERNEL_CREATION_FAILED = 8,
    RS_PROGRAM_CREATION_FAILED = 9,
    RS_NO_SOURCE_FOUND = 10,
    RS_LOADING_SOURCE_FAILED = 11
} rs_status;

/* Key types.  0..3 are the ones the reference instantiates (src/RadixSortGPU.cpp:598-601); 4 and 5 are the other two
 * rows of its type table (Common/CLTypeInformation.h:12-22: int16_t / uint16_t), sorted in two passes; 6..9 are
 * floating-point keys (additive: the reference excludes floats, doc/doc.tex:64).  Floats are ordered by value through
 * an order-preserving view of their bit patterns applied inside the digit extraction -- keys are stored and moved
 * unmodified, no extra pass: -inf < ... < -0.0 < +0.0 < ... < +inf < every NaN (NaNs of either sign and any payload
 * sort last, stable among themselves, as numpy and torch place them).
 * Integer-only entries (rs_merge_runs_device, rs_lower_bound_device, rs_sorter_sort_segments_device and the multi-GPU
 * path) reject types >= RS_KEY_F32 and 16-bit keys where stated. */
typedef enum rs_key_type {
    RS_KEY_U32 = 0,
    RS_KEY_I32 = 1,
    RS_KEY_U64 = 2,
    RS_KEY_I64 = 3,
    RS_KEY_U16 = 4,
    RS_KEY_I16 = 5,
    RS_KEY_F32 = 6,
    RS_KEY_F64 = 7,
    RS_KEY_F16 = 8,  /* IEEE binary16 */
    RS_KEY_BF16 = 9  /* bfloat16 */
} rs_key_type;

/* Dataset generators of the reference (src/Dataset.h:84-137). */
typedef enum rs_dataset_kind {
    RS_DATASET_ZEROS = 0,
    RS_DATASET_RANGE = 1,
    RS_DATASET_INVERTED_RANGE = 2,
    RS_DATASET_UNIFORM = 3, /* "Random Uniform": counter-based generator on the device */
    RS_DATASET_RANDOM = 4   /* "Random Random":  32-bit values, counter-based on the device */
} rs_dataset_kind;

/* rs_sorter_create flags */
#define RS_WITH_PERMUTATION 0x1u /* carry a uint32 index payload (stable permutation)          */
#define RS_NO_TMA           0x2u /* force the plain-load tile path (debug / A-B measurements)    */
#define RS_NO_PROFILE       0x4u /* do not record per-phase CUDA events in rs_sorter_sort        */
#define RS_NO_KEY_BUFFERS   0x8u /* allocate only scratch; use rs_sorter_sort_device exclusively */
#define RS_NO_PASS_SKIP     0x10u /* run every pass even when its digit is the same for all keys       */
#define RS_NO_SMALL_PATH    0x40u /* always take the multi-kernel path, even for inputs one CTA could sort alone    */
#define RS_NO_ATOMIC_RANK   0x20u /* rank with the shared-memory lane bitmap even if the atomic mode is available */

typedef struct rs_context_s* rs_context;
typedef struct rs_stream_s*  rs_stream;
typedef struct rs_sorter_s*  rs_sorter;

/* Mirrors struct Statistics (src/Statistics.h:4-31), including its update rule. */
typedef struct rs_statistics {
    double min, max, avg, sum;
    uint64_t n;
} rs_statistics;

/* Mirrors struct RuntimesGPU (src/RadixSortGPU.h:18-24). Milliseconds, CUDA-event timed.
 * time_histo  = the single upfront multi-digit histogram kernel,
 * time_scan   = the digit-offset scan kernel,
 * time_reorder= sum of the onesweep pass kernels of one sort,
 * time_paste  = always 0 (the paste step is subsumed by the decoupled look-back),
 * time_total  = first launch to last kernel end of one sort. */
typedef struct rs_runtimes {
    rs_statistics time_histo, time_scan, time_reorder, time_paste, time_total;
} rs_runtimes;

/* ---- library / errors ------------------------------------------------------------------ */
int         rs_version(void);                 /* major*1000 + minor */
const char* rs_status_string(int status);     /* "OK", "HOST_BUFFERS_FAILED", ... */
const char* rs_last_error_string(void);       /* thread-local detail of the last failure */

/* ---- device / context / stream  (replaces Common/ComputeState.cpp:14-104) ---------------- */
int rs_device_count(int* n);
int rs_context_create(int device, rs_context* out);
int rs_context_destroy(rs_context ctx);
int rs_context_sm_count(rs_context ctx, int* n);
/* 1 if sorters of this context rank with one shared-memory atomic per key (the fast default).  That mode is stable
 * -- and every LSD pass must be -- only because the hardware serves the lanes of one warp instruction that hit the same
 * shared word in ascending lane order.  NVIDIA does not document this, so it is not assumed: rs_context_create tests
 * the property in isolation AND sorts 2^20 keys + index payload of heavily colliding digits through the production
 * kernels, checking sortedness and the stable permutation on the device; any violation (a future driver or chip)
 * leaves the context in the lane-bitmap mode, which relies only on the PTX memory model and is about a third slower.
 * RS_NO_ATOMIC_RANK per sorter, or env RSB200_NO_ATOMIC_RANK=1 globally, select the bitmap mode by hand. */
int rs_context_atomic_rank(rs_context ctx, int* enabled);
int rs_stream_create(rs_context ctx, rs_stream* out);
/* Wrap an existing cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); not owned. */
int rs_stream_wrap(rs_context ctx, void* cuda_stream, rs_stream* out);
int rs_stream_sync(rs_stream s);
int rs_stream_destroy(rs_stream s);

/* Pinned host memory for the upload / download stage (src/RadixSortGPU.cpp:365-429). */
int rs_host_alloc(size_t bytes, void** out);
int rs_host_free(void* p);

/* ---- sorter (replaces RadixSortGPU<T> + ComputeDeviceData<T>) ---------------------------- */
/* Next multiple of 1024 (= _NUM_GROUPS*_NUM_ITEMS_PER_GROUP), src/RadixSortGPU.cpp:287-297.
 * Kept for callers that size host buffers like the reference; the kernels need no padding. */
uint64_t rs_round_up_keys(uint64_t n);

int rs_sorter_create(rs_context ctx, rs_key_type type, uint64_t max_keys, uint32_t flags, rs_sorter* out);
int rs_sorter_destroy(rs_sorter s);

/* Fill internal input keys [first_key, rounded) with max-1 (src/RadixSortGPU.cpp:269-285). */
int rs_sorter_pad(rs_sorter s, rs_stream st, uint64_t first_key, uint64_t rounded);

/* H2D of n keys and (if the sorter carries a permutation and h_perm != NULL) n indices.
 * With h_perm == NULL and RS_WITH_PERMUTATION the index payload is initialised to 0..n-1 on
 * the device (what src/HostData.cpp:20 uploads). Asynchronous on st. */
int rs_sorter_upload(rs_sorter s, rs_stream st, const void* h_keys, const uint32_t* h_perm, uint64_t n);

/* Sort the first n uploaded keys (stable LSD over all key bits, 8-bit digits). Asynchronous. */
int rs_sorter_sort(rs_sorter s, rs_stream st, uint64_t n);

/* D2H of the sorted keys / permutation; blocks until done. Either pointer may be NULL. */
int rs_sorter_download(rs_sorter s, rs_stream st, void* h_keys_out, uint32_t* h_perm_out, uint64_t n);

/* Same copies, enqueued only: the host buffers (pinned, rs_host_alloc) are valid after rs_stream_sync(st).
 * Lets a caller overlap the download of one batch with the upload of the next on a second stream. */
int rs_sorter_download_async(rs_sorter s, rs_stream st, void* h_keys_out, uint32_t* h_perm_out, uint64_t n);

/* D2H of the digit histogram table of the last sort: passes x 256 uint64 counts (pass-major).
 * Counterpart of the reference's `histograms`/`globsum` read-backs (src/RadixSortGPU.cpp:411-428),
 * whose contents the reference never asserts. Blocks. Returns number of passes in *passes. */
int rs_sorter_download_histograms(rs_sorter s, rs_stream st, uint64_t* h_hist, size_t capacity, int* passes);

/* Device-pointer entry: sort n keys (bits [begin_bit,end_bit)) from d_keys_in.
 * Pass p reads (p==0 ? in : buf[(p-1)&1]) and writes buf[p&1]; the result is in
 * buf[(passes-1)&1] and *result_index (may be NULL) receives that index. d_keys_in may alias
 * d_keys_buf1 (classic double buffer). Value pointers may all be NULL (keys only). n may
 * exceed the sorter's max_keys only if scratch was sized for it (max_keys >= n required). */
int rs_sorter_sort_device(rs_sorter s, rs_stream st,
                          const void* d_keys_in, void* d_keys_buf0, void* d_keys_buf1,
                          const uint32_t* d_vals_in, uint32_t* d_vals_buf0, uint32_t* d_vals_buf1,
                          uint64_t n, int begin_bit, int end_bit, int* result_index);

/* Number of passes the last sort skipped on the device because their digit was constant over all keys
 * (a stable pass over a constant digit is the identity: Zeros, the upper half of 64-bit keys holding 32-bit
 * values).  Throughput of such a sort must be reported against the unskipped byte count and flagged. Blocks. */
int rs_sorter_last_skipped_passes(rs_sorter s, rs_stream st, int* skipped);

int rs_sorter_runtimes(rs_sorter s, rs_runtimes* out);       /* blocks on pending events */
int rs_sorter_reset_runtimes(rs_sorter s);
/* number of kernels launched by this sorter so far (for the bench's gpu_launches claim) */
int rs_sorter_launch_count(rs_sorter s, uint64_t* out);

/* ---- multi-GPU building blocks (what rs_dist_* is made of; for callers that bring their own collectives) ----------
 * The reference has no counterpart (single OpenCL device, Common/ComputeState.cpp:10-12). */
int rs_device_alloc(rs_context ctx, size_t bytes, void** out);      /* plain cudaMalloc: exportable through CUDA IPC */
int rs_device_free(rs_context ctx, void* p);
#define RS_IPC_HANDLE_BYTES 64
int rs_ipc_export(rs_context ctx, const void* d_ptr, unsigned char handle[RS_IPC_HANDLE_BYTES]);  /* d_ptr: base of an rs_device_alloc */
int rs_ipc_open(rs_context ctx, const unsigned char handle[RS_IPC_HANDLE_BYTES], void** out);     /* map a peer GPU's allocation */
int rs_ipc_close(rs_context ctx, void* p);
/* Fused exchange + sort: the input is the concatenation of num_segments (<= 16) device segments that may live in
 * peer memory (rs_ipc_open).  ONE kernel pulls them over NVLink into d_gathered while it builds the digit histograms
 * of the sort (visiting the segments round-robin from first_segment); the passes then run as in
 * rs_sorter_sort_device with d_gathered as the input (it may alias d_keys_buf1).  The caller guarantees that the
 * peers' segments are complete (a collective after the peers' local sorts) and stay untouched until this call's
 * gather kernel has run on every rank (a collective after it).  Keys only. */
int rs_sorter_sort_segments_device(rs_sorter s, rs_stream st, int num_segments, const void* const* d_seg_ptrs,
                                   const uint64_t* seg_counts, int first_segment, void* d_gathered, void* d_keys_buf0,
                                   void* d_keys_buf1, int begin_bit, int end_bit, int* result_index);

/* One level of pairwise merges of sorted key runs (the multi-GPU final step, SURVEY.md 8e.5: "or a P-way merge"):
 * job j merges a[j] (a_counts[j] keys) and b[j] into out[j]; inputs may be peer memory (rs_ipc_open), so the first
 * level is exchange and merge in one kernel.  d_splits: device scratch of splits_capacity uint64
 * (>= sum over jobs of ceil((na+nb)/4096) + num_jobs).  Keys only; order is the key type's.
 * Runs may start at any element.  The kernel reads WHOLE 128-byte lines (every line that holds at least one key of
 * a run, nothing else), so a run must not end in the last partial line of a mapping -- cudaMalloc, rs_device_alloc
 * and torch's allocator all hand out multiples of 256 bytes or more.  Outputs that start on a 16-byte boundary take
 * the vector store path (8-byte keys). */
int rs_merge_runs_device(rs_context ctx, rs_stream st, rs_key_type type, int num_jobs, const void* const* d_a_ptrs,
                         const uint64_t* a_counts, const void* const* d_b_ptrs, const uint64_t* b_counts,
                         void* const* d_out_ptrs, uint64_t* d_splits, uint64_t splits_capacity);

/* ---- multi-GPU sort (BASELINE.json configs[4]) -------------------------------------------------------
 * One NVSwitch box, keys only, 32- and 64-bit integer keys.  Every GPU sorts its shard (onesweep), exact splitters are
 * found by refinement rounds on the devices (one ncclAllReduce per round, on the sort's stream), and the final local
 * step is a merge tree whose first level reads the peers' sorted runs straight over NVLink (exchange and merge are
 * one kernel).  Afterwards GPU r holds exactly floor((r+1)N/P) - floor(rN/P) keys, keys(r) <= keys(r+1), and the
 * concatenation over the GPUs is the sorted input.  NCCL (libnccl.so.2) is loaded at run time; RSB200_NCCL_LIB
 * overrides the path.  Asynchronous NCCL errors (ncclCommGetAsyncError) surface as RS_CALCULATION_FAILED.
 * The reference has no counterpart (one OpenCL device, Common/ComputeState.cpp:10-12).
 *
 * Two ways to drive it:
 *   rs_dist_create       one process drives all `ngpus` devices (peer access between them); rs_dist_sort takes and
 *                        returns arrays of `ngpus` entries, in the order of `devices` (NULL = 0 .. ngpus-1)
 *   rs_dist_create_rank  one process per GPU (rank of world): `id` comes from rs_dist_unique_id on one rank and is
 *                        handed to the others by the caller (MPI, torch.distributed, a file ...); the buffers the peers
 *                        read are exported through CUDA IPC; work runs on the caller's stream `st`; arrays have 1 entry */
typedef struct rs_dist_s* rs_dist;
#define RS_DIST_ID_BYTES 128
#define RS_DIST_STAGES 6
int rs_dist_unique_id(unsigned char id[RS_DIST_ID_BYTES]);
int rs_dist_create(int ngpus, const int* devices, rs_key_type type, uint64_t max_keys_per_gpu, rs_dist* out);
int rs_dist_create_rank(rs_context ctx, rs_stream st, rs_key_type type, uint64_t max_keys_per_gpu, int rank, int world,
                        const unsigned char id[RS_DIST_ID_BYTES], rs_dist* out);
int rs_dist_destroy(rs_dist d);
int rs_dist_num_local(rs_dist d, int* n);   /* GPUs this handle drives: ngpus, or 1 */
/* Sort the union of the shards (d_shards[i]: n_in[i] <= max_keys_per_gpu keys on local GPU i; the shards are not
 * modified).  d_out[i] / n_out[i] receive local GPU i's slice of the global order: a pointer into buffers owned by the
 * handle, valid until the next rs_dist_sort.  Synchronises the streams once in the middle (the cut table is read back);
 * the tail of the work (merge tree, closing all-reduce) is only enqueued: rs_dist_sync, or the caller's stream. */
int rs_dist_sort(rs_dist d, const void* const* d_shards, const uint64_t* n_in, void** d_out, uint64_t* n_out);
int rs_dist_sync(rs_dist d);
/* Device time (ms, CUDA events) of the stages of the last sort on local GPU `local_index`: [0] local sort,
 * [1] splitters + cut table, [2] merge level 1 (NVLink), [3] local merge levels, [4] closing all-reduce, [5] total. Blocks. */
int rs_dist_stage_ms(rs_dist d, int local_index, double out[RS_DIST_STAGES]);

/* ---- device-side helpers used by the harness and the multi-GPU path ----------------------- */
/* Fill d_out[0..n) with elements [global_offset, global_offset+n) of a dataset of total_n
 * elements (src/Dataset.h:84-137; the random kinds use a documented counter-based hash). */
int rs_dataset_fill_device(rs_context ctx, rs_stream st, rs_key_type type, rs_dataset_kind kind,
                           uint64_t seed, uint64_t global_offset, uint64_t total_n,
                           void* d_out, uint64_t n);

/* Order check + order-independent multiset fingerprint of d_keys[0..n), all on the device.
 * result[0] = number of adjacent pairs out of order (type-aware), result[1] = sum of keys
 * mod 2^64, result[2] = sum of mix64(key) mod 2^64, result[3] = xor of mix64(key).
 * Blocks. The same fingerprint over the input must equal the one over the output. */
int rs_check_device(rs_context ctx, rs_stream st, rs_key_type type, const void* d_keys, uint64_t n,
                    uint64_t result[4]);

/* Check that d_perm is a stable sorting permutation: keys_in[perm[i]] == keys_out[i] and, for equal
 * neighbours, perm[i] < perm[i+1]. result[0] = gather mismatches, result[1] = stability violations. */
int rs_check_permutation_device(rs_context ctx, rs_stream st, rs_key_type type, const void* d_keys_in,
                                const void* d_keys_out, const uint32_t* d_perm, uint64_t n,
                                uint64_t result[2]);

/* out[i] = first index j in sorted d_keys[0..n) with key[j] >= query[i] (type-aware order). */
int rs_lower_bound_device(rs_context ctx, rs_stream st, rs_key_type type, const void* d_keys, uint64_t n,
                          const void* d_queries, uint64_t num_queries, uint64_t* d_out);

#ifdef __cplusplus
}
#endif
#endif /* RSB200_H */
