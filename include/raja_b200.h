/* raja_b200 -- C ABI of the B200-native execution back-end for RAJA's data-parallel core.
 *
 * RAJA (the reference, /root/reference) has no FFI: its back-end boundary is C++
 * overload resolution on the execution-policy type, resolved in the user's TU
 * (SURVEY.md section 8b).  That boundary is re-created in include/RAJA_b200/RAJA.hpp.
 * This header is the C ABI *below* it: every non-templated piece of the back-end --
 * pools, stream bookkeeping, the type-erased scan / sort entry points the policy
 * adapter forwards to, and the named forall/reduce/atomic workloads of the
 * reference's examples and exercises, instantiated from the same forall engine
 * with fixed loop bodies -- behind plain pointers and sizes.
 *
 * Conventions
 *   - all data pointers are DEVICE-accessible pointers (device memory, or pinned
 *     mapped host memory) unless a parameter says "host";
 *   - `stream` is a cudaStream_t passed as void*; NULL is the legacy default stream.
 *     Work is enqueued on it and the call returns without synchronising
 *     (camp::resources::Cuda semantics of an Async policy; the policy adapter and
 *     the Python mirror add the wait for synchronous policies);
 *   - lengths are 64-bit (the reference truncates scan/sort lengths to int:
 *     policy/cuda/scan.hpp:65, sort.hpp:115,407);
 *   - every function returns 0 on success, a cudaError_t value for CUDA failures,
 *     or a negative RB200_E* code for argument errors; rb200_last_error() returns
 *     a thread-local message.  There is no CPU fallback anywhere in this library.
 */
#ifndef RAJA_B200_H
#define RAJA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RB200_VERSION 100 /* 0.1.0 */

typedef void* rb200_stream_t; /* cudaStream_t */

typedef enum rb200_dtype {
  RB200_I32 = 0,
  RB200_U32 = 1,
  RB200_I64 = 2,
  RB200_U64 = 3,
  RB200_F32 = 4,
  RB200_F64 = 5
} rb200_dtype;

/* binary operators: RAJA::operators::plus / minimum / maximum
 * (include/RAJA/util/Operators.hpp:247-375) */
typedef enum rb200_op { RB200_PLUS = 0, RB200_MINIMUM = 1, RB200_MAXIMUM = 2 } rb200_op;

/* reducer kinds: RAJA::ReduceSum/Min/Max/MinLoc/MaxLoc
 * (include/RAJA/policy/cuda/reduce.hpp:1174-1378) */
typedef enum rb200_reduce_kind {
  RB200_REDUCE_SUM = 0,
  RB200_REDUCE_MIN = 1,
  RB200_REDUCE_MAX = 2,
  RB200_REDUCE_MINLOC = 3,
  RB200_REDUCE_MAXLOC = 4
} rb200_reduce_kind;

#define RB200_EINVAL (-1)  /* bad enum / NULL pointer / negative length */
#define RB200_ENOTSUP (-2) /* dtype/op combination not provided */
#define RB200_ENOMEM (-3)

/* ---- library ---------------------------------------------------------------------- */
int rb200_version(void);
const char* rb200_last_error(void);
int rb200_sm_count(void);
/* cudaStreamSynchronize / cudaDeviceSynchronize: RAJA::synchronize<cuda_synchronize>
 * (policy/cuda/synchronize.hpp:39-42) and camp::resources::Cuda::wait(). */
int rb200_stream_synchronize(rb200_stream_t stream);
int rb200_device_synchronize(void);
/* Stream bookkeeping of RAJA::cuda::launch() / RAJA::cuda::synchronize()
 * (policy/cuda/MemUtils_CUDA.hpp:207-284, g_stream_info_map): every kernel the back-end enqueues marks its
 * stream as "has work in flight"; rb200_synchronize_streams() waits for exactly the marked streams of the
 * current device and clears the marks (streams nothing was launched on since the last call are skipped);
 * rb200_stream_mark_synced() clears one mark after the caller has waited for that stream itself.
 * Returns the number of streams waited for, or a negative error. */
void rb200_stream_mark_launch(rb200_stream_t stream);
void rb200_stream_mark_synced(rb200_stream_t stream);
int rb200_synchronize_streams(void);
/* Launch accounting for kernels instantiated OUTSIDE the library from the same header-only engine (the RAJA policy layer
 * in the user's translation unit): bracket each launch with these two calls and it is counted, marked on its stream and --
 * with profiling enabled -- timed like the library's own launches. */
void rb200_launch_begin(const char* kernel, rb200_stream_t stream);
void rb200_launch_end(const char* kernel, rb200_stream_t stream);
/* NVTX range around a named launch: RAJA::Name (policy/cuda/params/kernel_name.hpp:28-71) */
void rb200_range_push(const char* name);
void rb200_range_pop(void);

/* ---- launch accounting / profiling (no reference counterpart; bench.py uses it) -------
 * Every kernel the library launches is counted.  With profiling enabled each launch is
 * also bracketed by CUDA events recorded on the launching stream. */
long long rb200_launch_count(void);                 /* launches since the last reset */
int rb200_profile_enable(int on);                   /* 1: record events per launch */
int rb200_profile_reset(void);                      /* zero the counter, drop recorded events */
/* Sums the recorded launches whose kernel name equals `kernel` (NULL: all).  Synchronises
 * on the recorded events.  Names: forall_kernel, forall_reduce_kernel, scan_kernel,
 * histogram_kernel, scan_bins_kernel, onesweep_kernel, store_value_kernel, split_valloc_kernel. */
int rb200_profile_query(const char* kernel, double* total_ms, long long* launches);

/* ---- pools: replace RAJA::cuda::{device,device_zeroed,pinned}_mempool_type
 *      (policy/cuda/MemUtils_CUDA.hpp:69-167, util/basic_mempool.hpp:302-430).
 *      Stream-ordered: a freed block is only handed back to allocations on the same
 *      stream, which makes the reference's "free right after enqueue"
 *      (policy/cuda/scan.hpp:81, sort.hpp:142) actually safe. ----------------------- */
void* rb200_device_malloc(size_t bytes, rb200_stream_t stream);
void* rb200_device_zeroed_malloc(size_t bytes, rb200_stream_t stream); /* caller returns it zeroed */
void* rb200_pinned_malloc(size_t bytes);                                /* cudaHostAlloc, mapped */
int rb200_device_free(void* p, rb200_stream_t stream);
int rb200_device_zeroed_free(void* p, rb200_stream_t stream);
int rb200_pinned_free(void* p);
int rb200_pool_trim(void);                 /* release every cached block */
size_t rb200_pool_bytes_cached(void);      /* bytes currently held by the pools */

/* ---- forall, no reducer: policy/cuda/forall.hpp:383-655 ---------------------------- */
/* a[i] = b[i] + s * c[i]   (stream triad) */
int rb200_forall_triad(rb200_dtype t, void* a, const void* b, const void* c, double s, int64_t n, rb200_stream_t stream);
/* a[i] += s * b[i]         (examples/tut_daxpy.cpp:157-160) */
int rb200_forall_daxpy(rb200_dtype t, void* a, const void* b, double s, int64_t n, rb200_stream_t stream);

/* ---- forall + reducers: policy/cuda/reduce.hpp, params/reduce.hpp ------------------- */
/* Folds x[0..n) with the reducer `kind`.  *result_val receives the value (type t);
 * for MINLOC/MAXLOC *result_loc receives the index of the FIRST extremum (the
 * sequential answer), -1 if n == 0.  n == 0 yields the operator identity.
 * deterministic != 0: last-block-done combine in fixed slot order (default);
 * deterministic == 0: one atomic per block on the result (SUM/MIN/MAX of types with
 * hardware atomics only). */
int rb200_reduce(rb200_reduce_kind kind, rb200_dtype t, const void* x, int64_t n, void* result_val, int64_t* result_loc,
                 int deterministic, rb200_stream_t stream);
/* result = sum_i x[i] * y[i]   (exercises/dot-product_solution.cpp:135-138) */
int rb200_reduce_dot(rb200_dtype t, const void* x, const void* y, int64_t n, void* result, rb200_stream_t stream);
/* sum, min, max, minloc, maxloc of x in ONE pass (examples/forall-param-reductions.cpp:137-145).
 * Result layout for element type T:  T sum, T min, T max, T minloc_val, T maxloc_val,
 * then (8-byte aligned) int64 minloc, int64 maxloc. */
int rb200_reduce_multi(rb200_dtype t, const void* x, int64_t n, void* result, rb200_stream_t stream);
size_t rb200_reduce_multi_result_bytes(rb200_dtype t);
size_t rb200_reduce_multi_loc_offset(rb200_dtype t);

/* ---- scans: policy/cuda/scan.hpp:52-235 ---------------------------------------------- */
/* inclusive != 0: out[i] = x0 (+) ... (+) xi;  else out[i] = init (+) x0 ... x(i-1).
 * `init_host` points to one host value of type t (exclusive only; NULL = identity).
 * in == out selects the in-place variants. */
int rb200_scan(int inclusive, rb200_op op, rb200_dtype t, const void* in, void* out, int64_t n, const void* init_host,
               rb200_stream_t stream);
/* Same, seeded with a carry that lives in DEVICE memory (one value of type t; NULL = none):
 * out[i] = carry (+) [init (+)] x0 (+) ...   The sharded-range policy uses it to continue a scan
 * across GPUs without a host round trip (carry = fold of the lower ranks' totals). */
int rb200_scan_carry(int inclusive, rb200_op op, rb200_dtype t, const void* in, void* out, int64_t n, const void* init_host,
                     const void* carry_dev, rb200_stream_t stream);

/* ---- sorts: policy/cuda/sort.hpp:101-653 ----------------------------------------------- */
/* Stable; ascending (RAJA::operators::less) unless descending != 0 (operators::greater).
 * Sorted in place (temporaries come from the device pool). */
int rb200_sort(rb200_dtype key_t, void* keys, int64_t n, int descending, rb200_stream_t stream);
/* val_bytes in {4, 8}: values are opaque and permuted with their keys. */
int rb200_sort_pairs(rb200_dtype key_t, void* keys, void* vals, int val_bytes, int64_t n, int descending,
                     rb200_stream_t stream);
/* Same with caller-provided ping-pong partners (keys_alt / vals_alt, same sizes):
 * no pool traffic; the result still lands in keys / vals. */
int rb200_sort_pairs_buffered(rb200_dtype key_t, void* keys, void* keys_alt, void* vals, void* vals_alt, int val_bytes,
                              int64_t n, int descending, rb200_stream_t stream);

/* Stable multi-way split by nsplit <= 7 splitters (the exchange step of the multi-GPU sample sort,
 * SURVEY.md section 8e; no counterpart in the reference, which is single-device): element i goes
 * to bucket b = number of splitters that sort at or before keys_in[i], in the order the sorts
 * above use (ascending / descending, signed and floating keys by value, -0.0 == +0.0); inside a
 * bucket the input order is kept.  Buckets are written back to back into keys_out / vals_out
 * (must not alias the inputs).  `splitters_host`: nsplit HOST values of type key_t, already in
 * sort order.  `counts_dev`: 8 uint64 in DEVICE memory, receives the bucket sizes (unused ones 0).
 * val_bytes in {0, 4, 8}. */
int rb200_partition_pairs(rb200_dtype key_t, const void* keys_in, void* keys_out, const void* vals_in, void* vals_out,
                          int val_bytes, int64_t n, const void* splitters_host, int nsplit, int descending,
                          uint64_t* counts_dev, rb200_stream_t stream);

/* ---- peer-memory communicator (multi-GPU sharded range) -----------------
 * No counterpart in the reference (one device per process, MPI left to the application).  Each rank
 * owns one "symmetric workspace" in device memory that every process of the node has mapped (CUDA
 * IPC / fabric handles; the caller does the handle exchange, e.g. torch symmetric memory) and
 * passes the addresses of all of them, as seen by THIS process, in rank order.  The first
 * rb200_comm_header_bytes() bytes of each workspace belong to the library; the rest is the
 * caller's (the receive buffers of rb200_comm_exchange_pairs live there).  Calls are collective:
 * every rank issues the same sequence on its stream. */
typedef struct rb200_comm* rb200_comm_t;
int rb200_comm_create(int rank, int world, void* const* ws_ptrs, size_t ws_bytes, rb200_comm_t* out);
int rb200_comm_destroy(rb200_comm_t c);
size_t rb200_comm_header_bytes(void);
/* dst_dev[r * bytes ...] = rank r's src_dev[0 .. bytes); bytes a multiple of 8, <= 504.  One 1-CTA
 * kernel per rank: stores into the peers' workspaces over NVLink, then waits for their flags. */
int rb200_comm_allgather_small(rb200_comm_t c, const void* src_dev, int bytes, void* dst_dev, rb200_stream_t stream);
int rb200_comm_barrier(rb200_comm_t c, rb200_stream_t stream);
/* Fused stable partition + all-to-all of the sample sort: like rb200_partition_pairs with
 * nsplit = world - 1, but bucket d is written straight into rank d's receive buffers (at byte
 * offsets recv_keys_offset / recv_vals_offset of its workspace, recv_capacity elements each), in
 * source-rank order and, inside one source, in source order.  Synchronises the stream once (the
 * count matrix decides the write positions); *n_received = elements now in this rank's buffers.
 * Returns RB200_ENOMEM on every rank if any rank would overflow its receive buffers. */
int rb200_comm_exchange_pairs(rb200_comm_t c, rb200_dtype key_t, const void* keys_in, const void* vals_in, int val_bytes,
                              int64_t n, const void* splitters_host, int nsplit, int descending, size_t recv_keys_offset,
                              size_t recv_vals_offset, int64_t recv_capacity, int64_t* n_received, rb200_stream_t stream);
/* The whole exchange step of the sample sort without host-side plumbing: a regular sample of the UNSORTED shard
 * (samples_per_rank keys, <= 4096) is all-gathered through peer memory, every rank sorts the same gathered samples on its
 * device and picks the world - 1 splitters there (no splitter ever visits the host), then the fused partition + all-to-all
 * of rb200_comm_exchange_pairs runs with them.  One host synchronisation remains (the count matrix: the caller needs
 * *n_received).  RB200_ENOMEM -- returned by EVERY rank, they all see the same count matrix -- when some rank would receive
 * more than recv_capacity elements (heavily duplicated keys): the caller falls back to a collective all-to-all. */
int rb200_comm_sample_exchange_pairs(rb200_comm_t c, rb200_dtype key_t, const void* keys_in, const void* vals_in, int val_bytes, int64_t n,
                                     int samples_per_rank, int descending, size_t recv_keys_offset, size_t recv_vals_offset,
                                     int64_t recv_capacity, int64_t* n_received, rb200_stream_t stream);
/* count elements of dtype t (count * size <= 504 bytes; src_dev 8-byte aligned) combined over the ranks IN RANK ORDER (run-to-run
 * deterministic for floating point): exclusive_prefix == 0: every rank receives the fold over all ranks (all-reduce);
 * exclusive_prefix != 0: rank r receives the fold over the ranks below it, the identity on rank 0 (the carry of a sharded scan).
 * One 1-CTA kernel over peer memory, stream ordered, no host synchronisation. */
int rb200_comm_allreduce_small(rb200_comm_t c, rb200_op op, rb200_dtype t, const void* src_dev, void* dst_dev, int count, int exclusive_prefix,
                               rb200_stream_t stream);

/* ---- atomics: policy/cuda/atomic.hpp:665-860 (atomicAdd) ----------------------------------- */
/* hist[keys[i]] += 1 for i in [0,n)  (exercises/atomic-histogram_solution.cpp:199-202).
 * hist_t in {I32, U32, U64, F32, F64}.  aggregate != 0 uses the warp-aggregated add. */
int rb200_atomic_histogram(rb200_dtype hist_t, void* hist, const int32_t* keys, int64_t n, int aggregate,
                           rb200_stream_t stream);
/* *pi += sum_i dx / (1 + x_i^2), x_i = (i + 0.5) dx, one atomicAdd per i
 * (examples/pi-reduce_vs_atomic.cpp:183-186; caller multiplies by 4). */
int rb200_atomic_pi(double* pi, int64_t num_bins, int aggregate, rb200_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* RAJA_B200_H */
