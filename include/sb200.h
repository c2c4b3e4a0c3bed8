/* sb200.h -- C ABI of libsb200.so: the B200-native (sm_100a) implementation of SparseBase's
 * data-parallel preprocessing hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces the arithmetic of one function
 * of the reference (sparcityeu/SparseBase v0.3.1; paths relative to src/sparsebase/) and is
 * what the reference-side shim (sparsebase_b200/host/) binds:
 *
 *   sb200_coo_sort            format/coo.cc:96-157          COO ctor: sortedness check + (row,col) sort
 *   sb200_coo_to_csr          converter/converter_order_two.cc:162-212 (CooCsrFunctionConditional)
 *   sb200_csr_to_coo          converter/converter_order_two.cc:71-118  (CsrCooFunctionConditional)
 *   sb200_coo_to_csc          converter/converter_order_two.cc:20-70   (CooCscFunctionConditional)
 *   sb200_csr_to_csc          converter/converter_order_two.cc:119-128 (CsrCscFunctionConditional)
 *   sb200_compressed_sort     format/csr.cc:99-157, format/csc.cc:99-157  CSR/CSC ctor check + segment sort
 *   sb200_degree_reorder      reorder/degree_reorder.cc:22-62  (DegreeReorder::CalculateReorderCSR)
 *   sb200_rcm_reorder         reorder/rcm_reorder.cc:22-166    (RCMReorder::peripheral + GetReorderCSR)
 *   sb200_permute2d           permute/permute_order_two.cc:21-79 (PermuteOrderTwo::PermuteOrderTwoCSR)
 *   sb200_permute1d           permute/permute_order_one.cc:17-37 (PermuteOrderOne::PermuteArray)
 *   sb200_inverse_permutation bases/reorder_base.h:662-671     (ReorderBase::InversePermutation)
 *   sb200_degrees             feature/degrees.cc:93-105        (Degrees::GetDegreesCSR)
 *   sb200_degree_distribution feature/degree_distribution.cc:146-162 (GetDegreeDistributionCSR)
 *   sb200_degree_features     feature/degrees_degree_distribution.cc:147-166 (fused degrees +
 *                             distribution), feature/min_max_avg_degree.cc:168-191,
 *                             feature/bandwidth.cc:92-111, feature/profile.cc:92-106
 *   sb200_edges_to_coo        io/edge_list_reader.cc:28-151 (EdgeListReader::ReadCOO: self-edge
 *                             removal, undirected expansion, (row,col) sort, unique)
 *   sb200_malloc/free/memcpy_*  converter/converter_order_two_cuda.cu:11-105,
 *                             converter/converter_order_one_cuda.cu:10-43, utils/utils_cuda.cuh:6-9
 *   sb200_can_access_peer     converter/converter_cuda.cu:12-21 (CUDAPeerToPeer)
 *   sb200_device_count        context/cuda_context_cuda.cu:9-15 (CUDAContext ctor validation)
 *   sb200_partition_rows, sb200_*_block, sb200_exclusive_scan, sb200_rank_keys,
 *   sb200_max_degree, sb200_degree_histogram, sb200_degree_rank_combine
 *                             (new) per-GPU pieces of the row-block sharded multi-GPU path; the
 *                             reference has no distributed code
 *
 * Conventions
 *  - Plain C: pointers and sizes only, no C++/torch types.  Every function returns 0 on
 *    success or an SB200_ERR_* code and never throws; sb200_last_error() returns a
 *    thread-local message for the last failure.
 *  - Unless a parameter is prefixed h_, array pointers are DEVICE pointers on `device`.
 *    Outputs are caller-allocated (sb200_malloc or any cudaMalloc-compatible allocator, so
 *    they can be owned by a CUDACSR / CUDAArray with the reference's CUDADeleter = cudaFree).
 *  - Element types are given as SB200_* dtype codes: id_type (IDType), nnz_type (NNZType),
 *    val_type (ValueType; SB200_VOID or a NULL vals pointer means "no values").
 *    Supported: id_type in {I32,U32,I64,U64}, nnz_type in {I32,U32,I64,U64} with
 *    sizeof(nnz) >= sizeof(id), val_type any 4- or 8-byte type or VOID.  Values are moved
 *    bit-exactly and never interpreted (except as a tie-break key, see sb200_compressed_sort).
 *    Index values must be non-negative and below 2^31 (4-byte) / 2^62 (8-byte).
 *  - `stream` is a cudaStream_t (NULL = the legacy default stream).  Calls are asynchronous
 *    with respect to the host unless stated otherwise; scratch memory comes from a private
 *    stream-ordered memory pool of `device` (see sb200_trim).  A call never changes the
 *    calling thread's current device.
 *  - There is NO CPU fallback anywhere in this library: without a CUDA device every compute
 *    entry point fails with SB200_ERR_CUDA.
 */
#ifndef SB200_H_
#define SB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_ABI_VERSION 1

enum sb200_dtype {
  SB200_VOID = 0,
  SB200_I32 = 1,
  SB200_I64 = 2,
  SB200_U32 = 3,
  SB200_U64 = 4,
  SB200_F32 = 5,
  SB200_F64 = 6
};

enum sb200_error {
  SB200_OK = 0,
  SB200_ERR_CUDA = 1,        /* a CUDA runtime call failed (message has the cudaError string) */
  SB200_ERR_BAD_DTYPE = 2,   /* unsupported dtype combination */
  SB200_ERR_BAD_ARG = 3,     /* null pointer / negative size / size limit exceeded */
  SB200_ERR_BAD_DEVICE = 4,  /* device id out of range (reference: utils::CUDADeviceException) */
  SB200_ERR_ALLOC = 5,       /* allocation failed (reference: utils::AllocationException) */
  SB200_ERR_INTERNAL = 6
};

int sb200_abi_version(void);
const char *sb200_last_error(void);

/* ---- contexts, memory and transfer (reference: CUDAContext, CUDACSR/CUDAArray transfer) ---- */
int sb200_device_count(int *h_count);
int sb200_can_access_peer(int device, int peer_device, int *h_can);
int sb200_enable_peer_access(int device, int peer_device);
int sb200_malloc(int device, size_t bytes, void **h_out_ptr);
int sb200_free(int device, void *ptr);
int sb200_malloc_host(size_t bytes, void **h_out_ptr); /* pinned host memory */
int sb200_free_host(void *h_ptr);
int sb200_memcpy_h2d(int device, void *dst, const void *h_src, size_t bytes, void *stream);
int sb200_memcpy_d2h(int device, void *h_dst, const void *src, size_t bytes, void *stream);
int sb200_memcpy_d2d(int dst_device, void *dst, int src_device, const void *src, size_t bytes,
                     void *stream);
int sb200_stream_synchronize(int device, void *stream);
/* Scratch memory comes from a private stream-ordered pool per device that keeps freed blocks
 * for the next call.  sb200_trim synchronises the device and returns those blocks to the
 * driver (call it before another allocator needs the memory, e.g. after a one-off large
 * conversion).  The device's default memory pool is never touched by this library. */
int sb200_trim(int device);

/* ---- format constructors ---- */

/* COO constructor semantics (ignore_sort=false): if the (row,col) sequence has an inversion,
 * sort row/col/vals IN PLACE by (row, col).  *h_was_sorted (optional, host) receives 1 when
 * the input was already sorted.  Synchronises the stream (the check result steers the host). */
int sb200_coo_sort(int device, int64_t n, int64_t m, int64_t nnz, void *row, void *col,
                   void *vals, int id_type, int val_type, int *h_was_sorted, void *stream);

/* CSR / CSC constructor semantics (ignore_sort=false) on a compressed layout ptr[n_seg+1],
 * idx[nnz], vals[nnz]: if ANY segment is not non-decreasing, sort EVERY segment by
 * (idx, value) in place.  *h_was_sorted as above.  Synchronises the stream. */
int sb200_compressed_sort(int device, int64_t n_seg, int64_t n_idx, int64_t nnz, const void *ptr,
                          void *idx, void *vals, int id_type, int nnz_type, int val_type,
                          int *h_was_sorted, void *stream);

/* ---- order-two conversions ---- */

/* COO (constructed, i.e. (row,col)-sorted unless the caller used ignore_sort) -> CSR.
 * out_row_ptr[n+1], out_col[nnz], out_vals[nnz].  Followed by the CSR-constructor check/sort. */
int sb200_coo_to_csr(int device, int64_t n, int64_t m, int64_t nnz, const void *row,
                     const void *col, const void *vals, void *out_row_ptr, void *out_col,
                     void *out_vals, int id_type, int nnz_type, int val_type, void *stream);

/* CSR -> COO: out_row[nnz] (row_ptr expanded), out_col, out_vals copied. */
int sb200_csr_to_coo(int device, int64_t n, int64_t m, int64_t nnz, const void *row_ptr,
                     const void *col, const void *vals, void *out_row, void *out_col,
                     void *out_vals, int id_type, int nnz_type, int val_type, void *stream);

/* COO -> CSC: out_col_ptr[n+1] -- n = dims[0] entries, exactly like the reference
 * (converter_order_two.cc:32 sizes col_ptr with the ROW count; m > n is undefined behaviour
 * there and SB200_ERR_BAD_ARG here), out_row[nnz] ascending within each column,
 * out_vals[nnz]. */
int sb200_coo_to_csc(int device, int64_t n, int64_t m, int64_t nnz, const void *row,
                     const void *col, const void *vals, void *out_col_ptr, void *out_row,
                     void *out_vals, int id_type, int nnz_type, int val_type, void *stream);

/* CSR -> CSC (transpose of the layout): = csr_to_coo followed by coo_to_csc. */
int sb200_csr_to_csc(int device, int64_t n, int64_t m, int64_t nnz, const void *row_ptr,
                     const void *col, const void *vals, void *out_col_ptr, void *out_row,
                     void *out_vals, int id_type, int nnz_type, int val_type, void *stream);

/* ---- reorderings: out_inv[n] is the reference's result convention inv[old] = new ---- */
int sb200_degree_reorder(int device, int64_t n, const void *row_ptr, int ascending,
                         void *out_inv, int id_type, int nnz_type, void *stream);

int sb200_rcm_reorder(int device, int64_t n, int64_t nnz, const void *row_ptr, const void *col,
                      void *out_inv, int id_type, int nnz_type, void *stream);

/* Diagnostics of the last sb200_rcm_reorder call made by the calling thread:
 * h_out4 = {levels walked by the persistent cluster kernel, levels done with grid-wide
 * kernels, number of BFS traversals, number of non-trivial connected components}. */
int sb200_rcm_last_stats(int64_t *h_out4);
/* h_out2 = {times the cluster re-split the frontier evenly over its CTAs, times the kernel was
 * relaunched with another cluster size}. */
int sb200_rcm_last_resplits(int64_t *h_out2);
/* The BFS traversals of peripheral() after the first are run as the Cuthill-McKee traversal
 * itself (same level sets).  h_out3 = {confirmed: eccentricity did not grow, one BFS saved;
 * continued: it grew and the new root was unique by degree, no replay needed; replayed: it grew
 * and the new root depended on the FIFO order, the BFS was repeated literally}. */
int sb200_rcm_last_speculation(int64_t *h_out3);
/* SM-cycle counters of the narrow regime's per-level phases {claims, cluster barrier, recheck,
 * compaction, sibling sort + state update, 0, 0, 0} accumulated by CTA 0 (profiling aid). */
int sb200_rcm_last_cycles(int64_t *h_out8);

/* ---- applying a permutation ---- */

/* row_order[n] / col_order[m] are inverse permutations (inv[old] = new); NULL = identity.
 * out_row_ptr[n+1], out_col[nnz] (ascending within each row), out_vals[nnz]. */
int sb200_permute2d(int device, int64_t n, int64_t m, int64_t nnz, const void *row_ptr,
                    const void *col, const void *vals, const void *row_order,
                    const void *col_order, void *out_row_ptr, void *out_col, void *out_vals,
                    int id_type, int nnz_type, int val_type, void *stream);

/* out[order[i]] = vals[i]  (== out[i] = vals[inverse(order)[i]]) */
int sb200_permute1d(int device, int64_t len, const void *vals, const void *order, void *out,
                    int id_type, int val_type, void *stream);

/* out[perm[i]] = i */
int sb200_inverse_permutation(int device, int64_t len, const void *perm, void *out, int id_type,
                              void *stream);

/* ---- degree features ---- */
int sb200_degrees(int device, int64_t n, const void *row_ptr, void *out_degrees, int id_type,
                  int nnz_type, void *stream);

/* out_dist[i] = (row_ptr[i+1]-row_ptr[i]) / (FeatureType) nnz ; feature_type in {F32,F64} */
int sb200_degree_distribution(int device, int64_t n, int64_t nnz, const void *row_ptr,
                              void *out_dist, int nnz_type, int feature_type, void *stream);

/* Fused degree features and the quality metrics of a reordering: one pass over row_ptr
 * (degrees, distribution, min / max degree) and one over col (bandwidth, profile).
 * out_degrees (IDType[n]) and out_dist (FeatureType[n]) may be NULL; col may be NULL (then
 * bandwidth = profile = 0).  h_out_scalars[4] (host) = {min_degree, max_degree,
 * bandwidth = max |i-j| + 1 over the nonzeros, profile = sum_i (i - min(i, smallest column of
 * row i))}; *h_out_avg (host, optional) = (row_ptr[n] - row_ptr[0]) / (FeatureType) n rounded
 * in FeatureType.  Synchronises the stream. */
int sb200_degree_features(int device, int64_t n, int64_t nnz, const void *row_ptr,
                          const void *col, void *out_degrees, void *out_dist,
                          int64_t *h_out_scalars, double *h_out_avg, int id_type, int nnz_type,
                          int feature_type, void *stream);

/* ---- edge list -> COO (the step before the COO constructor) ---- */

/* EdgeListReader::ReadCOO on a device edge list u[n_edges], v[n_edges], w[n_edges] (w may be
 * NULL): edges with u == v are dropped when remove_self_edges; with read_undirected the reverse
 * edge follows every kept edge; n = max u + 1, m = max v + 1 over the kept edges, both set to
 * the larger when square or read_undirected; the entries are sorted by (row, col) and, with
 * remove_duplicates, the first entry of every run of equal (row, col) is kept -- the first in
 * INPUT order (stable sort; the reference's std::sort leaves the survivor unspecified when
 * duplicate edges carry different weights).  out_row / out_col / out_vals must hold
 * n_edges * (read_undirected ? 2 : 1) entries; h_out3 (host) = {n, m, nnz}.
 * Synchronises the stream. */
int sb200_edges_to_coo(int device, int64_t n_edges, const void *u, const void *v, const void *w,
                       int remove_duplicates, int remove_self_edges, int read_undirected,
                       int square, void *out_row, void *out_col, void *out_vals,
                       int64_t *h_out3, int id_type, int val_type, void *stream);

/* ---- multi-GPU sharding helper ---- */

/* Split rows [0,n) into `parts` contiguous blocks with (nearly) equal nnz: h_bounds[parts+1]
 * (host, int64) with bounds[0]=0, bounds[parts]=n and bounds[k] = the first row whose
 * row_ptr >= k*nnz/parts.  Synchronises the stream. */
int sb200_partition_rows(int device, int64_t n, int64_t nnz, const void *row_ptr, int nnz_type,
                         int parts, int64_t *h_bounds, void *stream);

/* Row-block variants: the caller owns rows [row_lo, row_lo + n_local) of an n x m matrix.
 * coo_to_csr_block: row[] holds GLOBAL row ids of that block; out_row_ptr[n_local+1] holds
 * block-local offsets.  csr_to_csc_block: row_ptr[n_local+1] is block-local; out_col_ptr has
 * m+1 entries (one per global column) and out_row holds GLOBAL row ids. */
int sb200_coo_to_csr_block(int device, int64_t row_lo, int64_t n_local, int64_t m, int64_t nnz,
                           const void *row, const void *col, const void *vals,
                           void *out_row_ptr, void *out_col, void *out_vals, int id_type,
                           int nnz_type, int val_type, void *stream);
int sb200_csr_to_csc_block(int device, int64_t row_lo, int64_t n_local, int64_t m, int64_t nnz,
                           const void *row_ptr, const void *col, const void *vals,
                           void *out_col_ptr, void *out_row, void *out_vals, int id_type,
                           int nnz_type, int val_type, void *stream);

/* out[0..n] (n+1 entries) = exclusive prefix sums of in[0..n-1]; dtype in {I32,U32,I64,U64}. */
int sb200_exclusive_scan(int device, int64_t n, const void *in, void *out, int dtype,
                         void *stream);

/* out_rank[i] = position of keys[i] in the stable ascending order of the keys (equal keys are
 * ranked by index); keys < key_bound. */
int sb200_rank_keys(int device, int64_t n, const void *keys, int64_t key_bound, void *out_rank,
                    int id_type, void *stream);

/* *h_out = max_i (row_ptr[i+1] - row_ptr[i]).  Synchronises the stream. */
int sb200_max_degree(int device, int64_t n, const void *row_ptr, int nnz_type, int64_t *h_out,
                     void *stream);

/* out_hist[d] (uint64, nbins entries, zeroed here) = number of rows with degree d. */
int sb200_degree_histogram(int device, int64_t n, const void *row_ptr, int nnz_type,
                           int64_t nbins, void *out_hist, void *stream);

/* out[i] = local_rank[i] + offset[degree(i)] (offset: int64 table indexed by degree); with
 * flip_from >= 0 the result is flip_from - that (DegreeReorder descending = reversed order). */
int sb200_degree_rank_combine(int device, int64_t n, const void *row_ptr, const void *local_rank,
                              const void *offset, int64_t flip_from, void *out, int id_type,
                              int nnz_type, void *stream);


/* reorder::ReorderHeatmap::ReorderHeatmapCSRArrayArray (reorder/reorder_heatmap.cc:43-120):
 * out_heat[num_parts * num_parts] (device, feature_type) = share of the nonzeros that fall into
 * every cell of a num_parts x num_parts grid once rows / columns are renumbered by order_r[n] /
 * order_c[m] (order[i] = new position of i; NULL = identity).  Blocks hold n / num_parts rows and
 * columns, the last one takes the remainder; the division is a float division as in the
 * reference.  SB200_ERR_BAD_ARG where the reference throws ReorderException (num_parts larger
 * than a dimension). */
int sb200_reorder_heatmap(int device, int64_t n, int64_t m, int64_t nnz, const void *row_ptr,
                          const void *col, const void *order_r, const void *order_c,
                          int num_parts, void *out_heat, int id_type, int nnz_type,
                          int feature_type, void *stream);

/* reorder::BOBAReorder::GetReorderCOO (reorder/boba_reorder.cc:35-137): out_inv[max(n, m)],
 * inv[v] = new position of v.  The COO must be (row, col)-sorted, as format::COO's constructor
 * leaves it.  Vertices are placed by their first appearance in the row array of the list sorted
 * by (col, row), then by their first appearance in its column array, then by id; the sequential
 * and the (single-threaded) parallel variant of the reference give this same order.
 * SB200_ERR_BAD_ARG when 2 * nnz does not fit in IDType (the reference computes it in IDType). */
int sb200_boba_reorder(int device, int64_t n, int64_t m, int64_t nnz, const void *row,
                       const void *col, void *out_inv, int id_type, void *stream);

/* ---- multi-GPU: row-block sharded operators over peer memory ----
 *
 * The reference has no distributed code (its "multi-GPU" is the peer copy of
 * converter/converter_order_two_cuda.cu:41-76).  These are the sharded forms of the operators
 * above: rank r of `world` owns the contiguous row block
 * [bounds[r], bounds[r+1]) with a block-local row_ptr and global column ids; results are
 * bit-identical to the single-GPU operators.  SPMD: every rank makes the same calls with its own
 * shard.  Every rank owns a WINDOW in its HBM that all peers address directly (CUDA IPC between
 * processes, peer access inside one process; NVLink / NVSwitch underneath); an exchange is a
 * kernel that stores straight into the destination rank's window followed by a flag barrier --
 * counts and offsets never leave the device.  The window must hold the largest exchange of the
 * operators used: about 2 x (shard bytes) + (n + 1) * sizeof(NNZType) is enough for all of them;
 * an operator that does not fit fails with SB200_ERR_BAD_ARG and a message giving the size.
 *
 *   one process per GPU     sb200_mg_comm_create on every rank, all-gather the 64-byte handles
 *                           with whatever the application uses (MPI, torch.distributed ...),
 *                           sb200_mg_comm_connect
 *   one process, ndev GPUs  sb200_mg_comm_create_local: ndev connected communicators; call the
 *                           operators from one host thread per GPU (sb200_mg_run_ranks), every
 *                           rank on its own stream.  Several ranks may share a device (tests):
 *                           that needs non-blocking streams, no cudaMalloc / cudaFree between
 *                           collective calls, and CUDA_MODULE_LOADING=EAGER (a lazily loaded
 *                           kernel synchronises the context on its first launch and would wait
 *                           for the other rank's barrier kernel)
 */
typedef struct sb200_mg_comm sb200_mg_comm_t;
int sb200_mg_comm_create(int device, int rank, int world, size_t window_bytes,
                         sb200_mg_comm_t **out, void *h_out_handle64);
int sb200_mg_comm_connect(sb200_mg_comm_t *comm, const void *h_all_handles /* world x 64 B */);
int sb200_mg_comm_create_local(int ndev, const int *devices, size_t window_bytes,
                               sb200_mg_comm_t **out_comms /* [ndev] */);
int sb200_mg_comm_destroy(sb200_mg_comm_t *comm);
int sb200_mg_comm_info(const sb200_mg_comm_t *comm, int *h_rank, int *h_world,
                       size_t *h_window_bytes);
/* One process, ndev GPUs: runs fn(rank, user) on one host thread per rank and returns the first
 * non-zero code (the operators below are collective: every rank must be inside the same call). */
int sb200_mg_run_ranks(int ndev, int (*fn)(int rank, void *user), void *user);
/* every rank has reached this point of `stream` and all earlier peer stores are visible */
int sb200_mg_barrier(sb200_mg_comm_t *comm, void *stream);
/* h_out[world] = every rank's value.  Synchronises the stream. */
int sb200_mg_allgather_i64(sb200_mg_comm_t *comm, int64_t value, int64_t *h_out, void *stream);

/* COO -> CSR of this rank's block (all nonzeros of rows [row_lo, row_lo + n_local), (row,col)-
 * sorted as the COO constructor leaves them): out_row_ptr[n_local+1] block-local, out_col /
 * out_vals[nnz_local]; h_out2 = {global nnz, nnz of the blocks before this one}.
 * Synchronises the stream. */
int sb200_mg_coo_to_csr(sb200_mg_comm_t *comm, int64_t row_lo, int64_t n_local, int64_t m,
                        int64_t nnz_local, const void *row, const void *col, const void *vals,
                        void *out_row_ptr, void *out_col, void *out_vals, int64_t *h_out2,
                        int id_type, int nnz_type, int val_type, void *stream);

/* DegreeReorder of the whole matrix from the block-local row_ptr's: out_inv[n] (the FULL
 * permutation, inv[old] = new, with the reference's tie rule) on every rank.  h_bounds[world+1]
 * (host) = the row blocks.  Synchronises the stream. */
int sb200_mg_degree_reorder(sb200_mg_comm_t *comm, int64_t n, const int64_t *h_bounds,
                            const void *row_ptr, int ascending, void *out_inv, int id_type,
                            int nnz_type, void *stream);

/* Permute2D: row_order[n] / col_order[m] are the FULL inverse permutations on every rank (NULL =
 * identity).  The result is sharded by nnz-balanced blocks of the NEW rows:
 *   _run    does the work and leaves this rank's new block in its window; h_out_bounds[world+1] =
 *           the new row blocks, h_out2[3] = {rows, nnz, nnz of the blocks before this one} of
 *           this rank's block.  Synchronises.
 *   _fetch  copies the block into out_row_ptr[rows+1] (block-local), out_col / out_vals[nnz] and
 *           releases the window (val_type / out_vals as in _run). */
int sb200_mg_permute2d_run(sb200_mg_comm_t *comm, int64_t n, int64_t m, int64_t nnz_total,
                           const int64_t *h_bounds, const void *row_ptr, const void *col,
                           const void *vals, const void *row_order, const void *col_order,
                           int64_t *h_out_bounds, int64_t *h_out2, int id_type, int nnz_type,
                           int val_type, void *stream);
int sb200_mg_permute2d_fetch(sb200_mg_comm_t *comm, int64_t n, int64_t new_rows, int64_t new_nnz,
                             void *out_row_ptr, void *out_col, void *out_vals, int id_type,
                             int nnz_type, int val_type, void *stream);

/* CSR -> CSC: the result is sharded by nnz-balanced blocks of COLUMNS (col_ptr block-local, row
 * ids global and ascending inside a column).  _run / _fetch as for Permute2D (h_out2[3] =
 * {columns, nnz, nnz of the blocks before}); _fetch takes the rank's column block
 * [col_lo, col_lo + n_cols) = h_out_bounds[rank], h_out2[0] of _run. */
int sb200_mg_csr_to_csc_run(sb200_mg_comm_t *comm, int64_t n, int64_t m, int64_t nnz_total,
                            const int64_t *h_bounds, const void *row_ptr, const void *col,
                            const void *vals, int64_t *h_out_bounds, int64_t *h_out2,
                            int id_type, int nnz_type, int val_type, void *stream);
int sb200_mg_csr_to_csc_fetch(sb200_mg_comm_t *comm, int64_t m, int64_t col_lo, int64_t n_cols,
                              void *out_col_ptr, void *out_row, void *out_vals, int id_type,
                              int nnz_type, int val_type, void *stream);

/* Permute1D: vals / order / out are this rank's block [h_bounds[rank], h_bounds[rank+1]) of the
 * arrays; out[order[i]] = vals[i] over the whole array (permute/permute_order_one.cc:17-37). */
int sb200_mg_permute1d(sb200_mg_comm_t *comm, const int64_t *h_bounds, const void *vals,
                       const void *order, void *out, int id_type, int val_type, void *stream);

/* Number of kernels launched by this library on the calling thread since the last reset
 * (bench.py reports it as gpu_launches). */
int64_t sb200_launch_count(void);
void sb200_reset_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SB200_H_ */
