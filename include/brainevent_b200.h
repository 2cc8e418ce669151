/*
 * brainevent_b200.h -- C ABI of the B200-native event-driven synaptic propagation library.
 *
 * Drop-in boundary for ONE path of chaobrain/brainevent: BinaryArray / BitPackedBinary /
 * CompactBinary events against CSR/CSC, FixedNumPerPre/PerPost, JITC{Scalar,Normal,Uniform} and
 * dense matrices.  Every entry point is `extern "C"`, takes plain device pointers, sizes and a
 * `cudaStream_t` (as void*), enqueues on that stream, never synchronises, keeps no global
 * mutable state, and returns 0 on success or a BE_B200_E* code (message via
 * be_b200_last_error()).  Outputs are treated as uninitialised on entry (the library zeroes what
 * it accumulates into), like the reference's cuda_raw kernels (_fcn/binary_fcnmv.cu:203).
 *
 * Citations `ref:` are paths under /root/reference/brainevent/ of the interface replaced.
 * The XLA-FFI handlers with the reference's own target names are declared in
 * brainevent_b200_xla.h; they are thin decoders over the functions below.
 */
#ifndef BRAINEVENT_B200_H_
#define BRAINEVENT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BE_B200_ABI_VERSION 1

/* ---- error codes (ref: include/brainevent/check.h:35-47 throws; here: return codes) -------- */
#define BE_B200_OK 0
#define BE_B200_EINVAL 3       /* XLA_FFI_Error_Code INVALID_ARGUMENT */
#define BE_B200_EUNIMPL 12     /* UNIMPLEMENTED */
#define BE_B200_EINTERNAL 13   /* INTERNAL (CUDA launch failure) */

/* ---- weight / output dtypes (ref: include/brainevent/tensor.h:42-59) ----------------------- */
#define BE_B200_F32 0
#define BE_B200_F64 1
#define BE_B200_F16 2
#define BE_B200_BF16 3

/* ---- event (spike) encodings ---------------------------------------------------------------
 * BOOL    : uint8/int8 per element, active when non-zero      (ref: BinaryArray bool, cuda_common.h:120-165)
 * F32     : float per element, active when > 0                (ref: BinaryArray float)
 * PACKED  : uint32 words, bit b of word w = element 32w+b     (ref: _event/bitpack_binary.py:32-76)
 *           1-D: ceil(n/32) words.  2-D [n, batch]: packed along batch, ceil(batch/32) words/row.
 * COMPACT : `spikes` = packed words as above, `aux` = int32 active_ids[n] followed by
 *           int32 n_active at aux[n]  (ref: _event/compact_binary.py:99-107)
 */
#define BE_B200_SPK_BOOL 0
#define BE_B200_SPK_F32 1
#define BE_B200_SPK_PACKED 2
#define BE_B200_SPK_COMPACT 3
/* COMPACT whose producer has already cleared the consumer's output vector (the fused
 * exchange + compaction of be_b200_wait_exchange_compact): 1-D scatter consumers skip their own clear */
#define BE_B200_SPK_COMPACT_CLEARED 4

/* ---- JITC weight families (ref: _jit_scalar/, _jit_normal/, _jit_uniform/) ------------------ */
#define BE_B200_JIT_SCALAR 0   /* w = p0 */
#define BE_B200_JIT_NORMAL 1   /* w = p0 + N01(seed,row,col) * p1 */
#define BE_B200_JIT_UNIFORM 2  /* w = p0 + U01(seed,row,col) * (p1 - p0) */

const char* be_b200_last_error(void);
int be_b200_abi_version(void);
/* Device capability check: returns 0 when the current device is sm_100 (B200). */
int be_b200_check_device(void);

/* Kernel-duration instrumentation for benchmarks: between begin/end every operator records a
 * CUDA event pair around its dominant kernel on the launching stream; end() synchronises those
 * events and returns the per-launch durations in milliseconds (launch order). */
int be_b200_profile_begin(int max_records);
int be_b200_profile_end(float* ms_out, int capacity, int* n_out);

/* Scratch space.  Every operator takes a caller-owned device workspace (the reference aliases
 * task_begin/task_end/status buffers the same way, ref: _csr/binary.py:55-124).  `n_src` is the
 * number of source (event) rows, `n_batch` the event batch width (1 for matvec).  The buffer
 * must be 256-byte aligned and ZERO on first use; the library leaves its control words zero
 * again after every call (self-resetting scan state), so no per-call memset is needed -- but a
 * buffer must not be re-used with a different (n_src, n_batch) without zeroing it again, and
 * concurrent calls (different streams) need separate buffers. */
size_t be_b200_workspace_bytes(int64_t n_src, int64_t n_batch);

/* ---- event encoders (ref: _event/compact.py:258,406,589; _event/compact.cu:70-430;
 *                          _jit_scalar/binary_jitsmv.cu:107-125 pack_bool) ---------------------- */
/* packed[ceil(n/32)] <- bits of spikes[n].  nonzero_pred=1 uses `!= 0` for floats (compaction
 * convention, ref: _event/compact.py:327), 0 uses `> 0` (operator convention). */
int be_b200_bitpack_1d(const void* spikes, int spk_kind, int64_t n, int nonzero_pred,
                       uint32_t* packed, void* stream);
/* Ordered stream compaction: active_ids[n] ascending with zero tail, n_active[1], and (when
 * packed != NULL) the bit-packed words in the same pass.  ref: binary_1d_array_index_p. */
int be_b200_compact_1d(const void* spikes, int spk_kind, int64_t n, int nonzero_pred,
                       int32_t* active_ids, int32_t* n_active, uint32_t* packed,
                       void* workspace, size_t workspace_bytes, void* stream);
/* 2-D [n, batch] row-major: packed[n, ceil(batch/32)] (may be NULL: compact-only primitive),
 * rows with any active column in ascending order.  ref: binary_2d_array_index_p /
 * binary_2d_compact_only_p. */
int be_b200_compact_2d(const void* spikes, int spk_kind, int64_t n, int64_t batch,
                       int nonzero_pred, uint32_t* packed, int32_t* active_ids,
                       int32_t* n_active, void* workspace, size_t workspace_bytes, void* stream);

/* ---- CSR / CSC (ref: _csr/binary.py:128 binary_csrmv, :264 binary_csrmm;
 *                      _csr/binary_indexed.py:70,615 *_indexed) --------------------------------
 * CSR structure with m rows, k columns, nse stored synapses.  weights: [nse] or [1] (homo=1).  indptr int32 or int64
 * (indptr_is64).  perm (nullable): indexed variant reads weights[perm[j]] (perm_is64 selects
 * int64).  transpose=1: events[m] -> out[k] (atomic scatter over active rows);
 * transpose=0: events[k] -> out[m] (deterministic gather over all rows).
 * mm: events [m or k, n] row-major, out [k or m, n] row-major (batch innermost).  For the
 * scatter matmat (transpose=1) `out_batch_major=1` writes the physical layout [n, k] instead --
 * the reference's own physical layout for this op (ref: _csr/binary.py:1263-1286) and the fast
 * one here: each batch column is a contiguous output vector, columns are processed in L2-sized
 * groups. */
int be_b200_binary_csrmv(const void* weights, int w_dtype, int homo, const int32_t* indices,
                         const void* indptr, int indptr_is64, const void* perm, int perm_is64,
                         const void* spikes, const void* aux, int spk_kind, int64_t m, int64_t k,
                         int64_t nse, int transpose, void* out, void* workspace,
                         size_t workspace_bytes, void* stream);
/* spk @ csr (events[m] -> out[k]) with the variant picked ON THE DEVICE from the spike rate: the
 * active-row atomic scatter over the CSR when n_active <= rate_threshold * m, else the
 * deterministic post-major gather over the CSC mirror (csc_indptr/csc_indices/perm from
 * be_b200_csr_to_csc; offsets share indptr's width).  No host synchronisation: both kernels are
 * enqueued and the one not selected exits at once.  ref: _csr/main.py:1641-1655 (mirror route). */
int be_b200_binary_csrmv_auto(const void* weights, int w_dtype, int homo, const int32_t* indices,
                              const void* indptr, int indptr_is64, const void* csc_indptr,
                              const int32_t* csc_indices, const void* perm, const void* spikes,
                              const void* aux, int spk_kind, int64_t m, int64_t k, int64_t nse,
                              float rate_threshold, void* out, void* workspace,
                              size_t workspace_bytes, void* stream);
int be_b200_binary_csrmm(const void* weights, int w_dtype, int homo, const int32_t* indices,
                         const void* indptr, int indptr_is64, const void* perm, int perm_is64,
                         const void* spikes, const void* aux, int spk_kind, int64_t m, int64_t k,
                         int64_t nse, int64_t n, int transpose, int out_batch_major, void* out,
                         void* workspace, size_t workspace_bytes, void* stream);

/* ---- fixed-number connectivity / ELL (ref: _fcn/binary.py:43 binary_fcnmv, :512 binary_fcnmm;
 *      _fcn/binary_fcnmv.cu:188-252) -----------------------------------------------------------
 * indices int32 [n_rows, n_conn]; weights [n_rows, n_conn] or [1].  transpose=1: events[n_rows]
 * -> out[n_out] scatter; transpose=0: events[n_src] -> out[n_rows] gather (n_out == n_rows). */
int be_b200_binary_fcnmv(const void* weights, int w_dtype, int homo, const int32_t* indices,
                         int64_t n_rows, int64_t n_conn, const void* spikes, const void* aux,
                         int spk_kind, int64_t n_src, int64_t n_out, int transpose, void* out,
                         void* workspace, size_t workspace_bytes, void* stream);
int be_b200_binary_fcnmm(const void* weights, int w_dtype, int homo, const int32_t* indices,
                         int64_t n_rows, int64_t n_conn, const void* spikes, const void* aux,
                         int spk_kind, int64_t n_src, int64_t n_out, int64_t n, int transpose,
                         int out_batch_major, void* out, void* workspace, size_t workspace_bytes,
                         void* stream);

/* ---- JITC (ref: _jit_scalar/binary.py:44,171; _jit_normal/binary.py:44,177;
 *      _jit_uniform/binary.py:44,177; CUDA _jit_scalar/binary_jitsmv.cu:127-214) ---------------
 * p0, p1, clen, seed are DEVICE pointers to one element each (XLA buffers in the reference):
 * p0/p1 in w_dtype, clen/seed int32.  corder=1 ("notrans"): one generated row per output,
 * gathers events over the walk dimension vec_len.  corder=0 ("trans"): one generated row per
 * event, scatters over out_len.  chunk_size = max(1, ceil(shape[1]/4)) of the *logical* shape
 * (ref: _misc.py:74-122); the caller passes it, exactly like the reference's FFI attribute.
 * row_offset / row_count restrict the generated rows (gather: output rows; scatter: event rows)
 * and col_begin/col_end restrict the walked columns written (scatter) -- used by the post-sharded
 * multi-GPU path; pass 0, -1, 0, -1 for the whole matrix.
 * mm: events [vec_len, n] row-major, out [out_len, n]; stride-4 walk (a different matrix than
 * mv, ref: _misc.py:33-38). */
int be_b200_binary_jitmv(int kind, const void* p0, const void* p1, int w_dtype,
                         const int32_t* clen, const int32_t* seed, const void* spikes,
                         const void* aux, int spk_kind, int64_t vec_len, int64_t out_len,
                         int64_t chunk_size, int corder, int64_t row_offset, int64_t row_count,
                         int64_t col_begin, int64_t col_end, void* out, void* workspace,
                         size_t workspace_bytes, void* stream);
int be_b200_binary_jitmm(int kind, const void* p0, const void* p1, int w_dtype,
                         const int32_t* clen, const int32_t* seed, const void* spikes,
                         const void* aux, int spk_kind, int64_t vec_len, int64_t out_len,
                         int64_t n, int64_t chunk_size, int corder, void* out, void* workspace,
                         size_t workspace_bytes, void* stream);
/* Connectivity dump of generated rows [row_begin, row_begin+n_rows): counts[n_rows] and, when
 * cols != NULL, the first `cap` columns of each row in walk order at cols[r*cap ..].  Parity
 * instrument (ref: _jit_scalar/csr.py jits_to_csr count/fill). */
int be_b200_jitc_row_connections(const int32_t* clen, const int32_t* seed, int64_t row_begin,
                                 int64_t n_rows, int64_t k, int64_t chunk_size, int stride,
                                 int32_t* counts, int32_t* cols, int64_t cap, void* stream);

/* ---- float (non-event) products through the same connectivity: the callers either side of the
 *      event path.  x / X hold VALUES in the weight dtype; X is [rows, n] row-major
 *      and out is [rows_out, n] row-major.  No workspace.
 *      ref: _csr/float.py:49 csrmv, :559 csrmm (_csr/float_csrmv.cu, float_csrmm.cu);
 *           _fcn/float.py:33 fcnmv, :136 fcnmm (_fcn/float_fcnmv.cu, float_fcnmm.cu);
 *           _jit_scalar/float.py:159 jitsmv, :1024 jitsmm and the normal/uniform analogues
 *           (float_jitsmv.cu, float_jitsmm.cu, ... next to each family's float.py). */
int be_b200_csrmv(const void* weights, int w_dtype, int homo, const int32_t* indices,
                  const void* indptr, int indptr_is64, const void* x, int64_t m, int64_t k,
                  int64_t nse, int transpose, void* out, void* stream);
int be_b200_csrmm(const void* weights, int w_dtype, int homo, const int32_t* indices,
                  const void* indptr, int indptr_is64, const void* X, int64_t m, int64_t k,
                  int64_t nse, int64_t n, int transpose, void* out, void* stream);
/* indices [n_rows, n_conn]; transpose=1: x[n_rows] -> out[n_other]; transpose=0: x[n_other] -> out[n_rows] */
int be_b200_fcnmv(const void* weights, int w_dtype, int homo, const int32_t* indices,
                  int64_t n_rows, int64_t n_conn, const void* x, int64_t n_other, int transpose,
                  void* out, void* stream);
int be_b200_fcnmm(const void* weights, int w_dtype, int homo, const int32_t* indices,
                  int64_t n_rows, int64_t n_conn, const void* X, int64_t n_other, int64_t n,
                  int transpose, void* out, void* stream);
/* same parameter meaning as be_b200_binary_jitmv / _jitmm without windows */
int be_b200_jitmv(int kind, const void* p0, const void* p1, int w_dtype, const int32_t* clen,
                  const int32_t* seed, const void* x, int64_t vec_len, int64_t out_len,
                  int64_t chunk_size, int corder, void* out, void* stream);
int be_b200_jitmm(int kind, const void* p0, const void* p1, int w_dtype, const int32_t* clen,
                  const int32_t* seed, const void* X, int64_t vec_len, int64_t out_len, int64_t n,
                  int64_t chunk_size, int corder, void* out, void* stream);

/* ---- JITC -> CSR materialisation (ref: _jit_scalar/csr.py:488 jits_to_csr with its count (:112)
 *      and fill (:310) passes, _jit_normal/csr.py:465, _jit_uniform/csr.py; csr.cu next to each).
 *      The result is the canonical CSR (columns ascending inside every row) of the [n_rows, n_cols]
 *      matrix that jit?mv (stride 32) or jit?mm (stride 4) applies with transpose=0 and the given
 *      corder.  count writes int64 indptr[n_rows+1] (indptr[n_rows] = nnz); the host reads nnz,
 *      allocates indices[nnz] / data[nnz] and calls fill with the same parameters.  One workspace
 *      (256-byte aligned, be_b200_jitc_csr_workspace_bytes(n_rows, nnz); nnz = 0 for the count
 *      pass) serves both. ------------------------------------------------------------------------ */
size_t be_b200_jitc_csr_workspace_bytes(int64_t n_rows, int64_t nnz);
int be_b200_jitc_csr_count(const int32_t* clen, const int32_t* seed, int64_t n_rows, int64_t n_cols,
                           int64_t chunk_size, int stride, int corder, int64_t* indptr,
                           void* workspace, size_t workspace_bytes, void* stream);
int be_b200_jitc_csr_fill(int kind, const void* p0, const void* p1, int w_dtype, const int32_t* clen,
                          const int32_t* seed, int64_t n_rows, int64_t n_cols, int64_t chunk_size,
                          int stride, int corder, const int64_t* indptr, int64_t nnz,
                          int32_t* indices, void* data, void* workspace, size_t workspace_bytes,
                          void* stream);

/* ---- dense 2-D spike matrix -> sparse event layouts (ref: _event/compact.py
 *      binary_2d_pair_stream_encode :706, binary_2d_row_sparse_encode :875, binary_2d_csr_encode :1180,
 *      binary_2d_csc_encode :1265 and _event/compact.cu).  spikes [n_src, n_batch] row-major, kind
 *      BOOL or F32 (active when != 0); n_src * n_batch < 2^31.  Outputs have the reference's static
 *      capacity and zero-filled tails; order is row-major (column-major for CSC) like the numba
 *      kernels.  workspace: be_b200_workspace_bytes(n_src * n_batch, 1), zero on first use.
 *        pair_stream int32 [n_src*n_batch, 2] (row, col), n_pairs int32[1]
 *        csr: indices int32 [n_src*n_batch] (batch column ids), indptr int32 [n_src + 1]
 *        csc: indices int32 [n_src*n_batch] (source row ids), indptr int32 [n_batch + 1];
 *             scratch = n_src*n_batch bytes
 *        row_sparse: spike_indices int32 [n_src, row_size] of 1-based column ids, zero padded;
 *             max_row_nnz int32[1] receives the largest row count (the host raises if > row_size,
 *             as the reference validates before launching) --------------------------------------- */
int be_b200_binary_2d_pair_stream_encode(const void* spikes, int spk_kind, int64_t n_src, int64_t n_batch,
                                         int32_t* pair_stream, int32_t* n_pairs, void* workspace,
                                         size_t workspace_bytes, void* stream);
int be_b200_binary_2d_csr_encode(const void* spikes, int spk_kind, int64_t n_src, int64_t n_batch,
                                 int32_t* indices, int32_t* indptr, void* workspace,
                                 size_t workspace_bytes, void* stream);
int be_b200_binary_2d_csc_encode(const void* spikes, int spk_kind, int64_t n_src, int64_t n_batch,
                                 int32_t* indices, int32_t* indptr, void* scratch, void* workspace,
                                 size_t workspace_bytes, void* stream);
int be_b200_binary_2d_row_sparse_encode(const void* spikes, int spk_kind, int64_t n_src, int64_t n_batch,
                                        int64_t row_size, int32_t* spike_indices, int32_t* max_row_nnz,
                                        void* stream);

/* ---- spike-triggered additive plasticity.  In place on `weights`:
 *      for every ACTIVE row r and stored slot j of r:
 *          w[slot(j)] = clamp(w[slot(j)] + trace[indices[j]], w_min, w_max),  slot(j) = perm ? perm[j] : j
 *      Pass -INFINITY / +INFINITY for an absent bound.  n_conn >= 0 selects fixed-number (ELL) rows
 *      and ignores indptr.  Events follow the be_b200_binary_* convention (length n_rows).
 *      ref: _csr/plasticity_binary.py:45 update_csr_on_binary_pre (rows = pre, perm = NULL),
 *           :477 update_csr_on_binary_post (rows = post over the CSC mirror, perm = weight_indices),
 *           _fcn/plasticity_binary.py:207/:269 (ELL rows), plasticity_binary_on_{pre,post}.cu,
 *           plasticity_row_driven.cu.  The reference clips the whole array afterwards: be_b200_clip.
 *      dense (ref: _dense/plasticity_binary.py:42, :360): W [rows, cols] row-major; on_post=0:
 *      W[i,:] += trace[cols] for active i (events length rows); on_post=1: W[:,j] += trace[rows]
 *      for active j (events length cols). ------------------------------------------------------- */
int be_b200_plasticity_rows(void* weights, int w_dtype, const int32_t* indices, const void* indptr,
                            int indptr_is64, int64_t n_conn, const void* perm, int perm_is64,
                            const void* spikes, const void* aux, int spk_kind, const void* trace,
                            int64_t n_rows, int64_t n_trace, int64_t nse, double w_min, double w_max,
                            void* workspace, size_t workspace_bytes, void* stream);
int be_b200_plasticity_dense(void* W, int w_dtype, int64_t rows, int64_t cols, const void* spikes,
                             const void* aux, int spk_kind, const void* trace, int on_post,
                             double w_min, double w_max, void* workspace, size_t workspace_bytes,
                             void* stream);
int be_b200_clip(void* w, int w_dtype, int64_t n, double w_min, double w_max, void* stream);

/* ---- dense event ops (ref: _dense/binary.py:79 binary_densemv, :487 binary_densemm;
 *      _dense/binary_densemv.cu, binary_densemm.cu) --------------------------------------------
 * W row-major [rows, cols].  transpose=1: events[rows] -> out[cols]; transpose=0: events[cols]
 * -> out[rows].  mm: events [k, n] -> out [m, n]. */
int be_b200_binary_densemv(const void* W, int w_dtype, int64_t rows, int64_t cols,
                           const void* spikes, const void* aux, int spk_kind, int transpose,
                           void* out, void* workspace, size_t workspace_bytes, void* stream);
int be_b200_binary_densemm(const void* W, int w_dtype, int64_t rows, int64_t cols,
                           const void* spikes, const void* aux, int spk_kind, int64_t n,
                           int transpose, void* out, void* workspace, size_t workspace_bytes,
                           void* stream);

/* ---- CSC mirror (ref: _misc.py:1341-1384 _csr_to_csc_index_numpy; _csr/csr_to_csc.cu) --------
 * Stable by-column reordering: csc_indptr[n_post+1], csc_indices[nse] (= row ids), perm[nse].
 * Offsets are int64 when is64, else int32.  Setup-time; scratch comes from `workspace`
 * (>= be_b200_csc_workspace_bytes). */
size_t be_b200_csc_workspace_bytes(int64_t m, int64_t n_post, int64_t nse);
int be_b200_csr_to_csc(const void* indptr, int is64, const int32_t* indices, int64_t m,
                       int64_t n_post, int64_t nse, void* csc_indptr, int32_t* csc_indices,
                       void* perm, void* workspace, size_t workspace_bytes, void* stream);

/* ---- multi-GPU spike exchange over NVLink peer memory (net-new: the reference is single-device) --
 * Fused "bit-pack my block + all-gather the words": each rank packs into its OWN symmetric
 * buffer and publishes a step number in a flag slot on each peer; be_b200_wait_exchange then
 * holds the stream (acquire loads, bounded spin) until all `world` peers have published and
 * pulls every peer's words over NVLink into `gathered_out` = uint32[world * words_per_rank].  peer_bufs[p] / peer_flags[p] are DEVICE arrays of peer-mapped pointers (e.g.
 * from a torch symmetric-memory rendezvous): buffer = uint32[2 * world * words_per_rank]
 * (double-buffered by the parity of the device-side step counter), flags = uint32[world], both
 * zero before first use.  `ctl` is a local zero-initialised control block of
 * be_b200_exchange_ctl_bytes() (ctl[2] != 0 afterwards means a wait timed out). */
size_t be_b200_exchange_ctl_bytes(void);
int be_b200_pack_publish(const void* spikes_local, int spk_kind, int64_t n_local,
                         void* const* peer_bufs, void* const* peer_flags, int world, int rank,
                         int64_t words_per_rank, void* ctl, void* stream);
int be_b200_wait_exchange(const void* my_flags, void* const* peer_bufs, int world, int64_t words_per_rank,
                          uint32_t* gathered_out, void* ctl, void* stream);
/* Scatter consumers: the wait kernel also compacts the gathered words into ids_and_count =
 * int32[n_bits + 1] (active ids in any order, count at [n_bits]) and clears clear_ptr[0..clear_bytes)
 * (the consumer's output) while it waits; consume with spk_kind BE_B200_SPK_COMPACT_CLEARED,
 * spikes = gathered_out, aux = ids_and_count.  The pack of the same step must be issued through
 * be_b200_pack_publish_reset with reset_counter = ids_and_count + n_bits. */
int be_b200_pack_publish_reset(const void* spikes_local, int spk_kind, int64_t n_local,
                               void* const* peer_bufs, void* const* peer_flags, int world, int rank,
                               int64_t words_per_rank, void* ctl, int32_t* reset_counter, void* stream);
int be_b200_wait_exchange_compact(const void* my_flags, void* const* peer_bufs, int world,
                                  int64_t words_per_rank, int64_t n_bits, uint32_t* gathered_out,
                                  int32_t* ids_and_count, void* clear_ptr, size_t clear_bytes,
                                  void* ctl, void* stream);
/* pack + publish that also resets the compaction counter (nullable) and clears clear_ptr[0..clear_bytes)
 * (nullable) after its own words are out -- the first kernel of the fused two-kernel step of
 * be_b200_sharded_step. */
int be_b200_pack_publish_clear(const void* spikes_local, int spk_kind, int64_t n_local,
                               void* const* peer_bufs, void* const* peer_flags, int world, int rank,
                               int64_t words_per_rank, void* ctl, int32_t* reset_counter,
                               void* clear_ptr, size_t clear_bytes, void* stream);
/* Flag waits are bounded by wall time: ctl word [3] (uint32 milliseconds, 0 = 20 s default).  A wait that
 * gives up latches ctl word [2]; from then on every wait kernel of that exchange skips its pull (scatter
 * consumers see n_active = 0) and the host entry points (be_b200_sharded_wait, dist.HostStepPipeline)
 * fail: after a timeout a rank may have run ahead of a peer's double buffer, so nothing later is trusted. */

/* ---- post-sharded projection: one rank's share as an object (net-new) ------------------------
 * Rank g of `world` owns the post block [g*per, (g+1)*per), per = 32 * words_per_rank, and holds only the
 * synapses that end there: rows = ALL n_neurons pre neurons, column ids LOCAL to the block.  The spikes of a
 * timestep arrive as each rank's own block (n_local values) and are exchanged as bit-packed words over
 * peer memory.  peer_bufs_host[p] = rank p's symmetric buffer mapped into this process
 * (be_b200_sharded_symm_bytes() bytes, zeroed before the rendezvous; world == 1: any device buffer).
 *   stored connectivity (set_fcn / set_csr): two kernels per timestep, the all-gather fused into the
 *     scatter (own rows first, each peer's words as its flag arrives) -- csrc/sharded.cu;
 *   set_jitc: the GLOBAL shape (n_neurons x n_post_global) keys the RNG; corder=1 (gather order, rows =
 *     post) takes a row window, corder=0 (scatter order) a column window;
 *   set_dense: W [n_neurons, n_out] row-major = this rank's post columns of the dense matrix.
 * All connectivity pointers are DEVICE pointers that stay owned by the caller. */
#define BE_B200_CONN_CSR 0
#define BE_B200_CONN_FCN 1
#define BE_B200_CONN_JITC 2
#define BE_B200_CONN_DENSE 3
typedef struct be_b200_sharded be_b200_sharded;
size_t be_b200_sharded_symm_bytes(int world, int64_t words_per_rank);
/* ctl: NULL, or a zeroed device block of be_b200_exchange_ctl_bytes() shared with direct
 * be_b200_pack_publish / be_b200_wait_exchange calls on the same buffers (one step counter) */
be_b200_sharded* be_b200_sharded_create(int world, int rank, int64_t n_neurons, int64_t n_out,
                                        int64_t words_per_rank, void* const* peer_bufs_host, void* ctl);
int be_b200_sharded_set_fcn(be_b200_sharded* h, const void* weights, int w_dtype, int homo,
                            const int32_t* indices, int64_t n_rows, int64_t n_conn);
int be_b200_sharded_set_csr(be_b200_sharded* h, const void* weights, int w_dtype, int homo,
                            const int32_t* indices, const void* indptr, int indptr_is64,
                            int64_t n_rows, int64_t nse);
int be_b200_sharded_set_jitc(be_b200_sharded* h, int jit_kind, const void* p0, const void* p1, int w_dtype,
                             const int32_t* clen, const int32_t* seed, int64_t n_post_global,
                             int64_t chunk_size, int corder);
int be_b200_sharded_set_dense(be_b200_sharded* h, const void* W, int w_dtype);
/* options: "fused" (stored connectivity: 1 = two-kernel fused step, 0 = pack / wait+compact / scatter),
 * "lanes_per_row" (0 auto, 8, 16, 32), "chunk_words" (0 auto, 16, 32, 64: spike words per work item of the fused
 * scatter), "spin_ms" (flag-wait bound) */
int be_b200_sharded_set_option(be_b200_sharded* h, const char* name, int64_t value);
/* One timestep, device resident: spikes_local [n_local] (BOOL or F32) -> out [n_out] in the weight dtype.
 * Enqueues on `stream`, never synchronises; collective (every rank calls it equally often). */
int be_b200_sharded_step(be_b200_sharded* h, const void* spikes_local, int spk_kind, void* out, void* stream);
/* The two phases of a step, for callers that publish early and consume later: publish = bit-pack + flag the
 * peers (+ clear `out`); consume = pull the peers' words and propagate them into `out`.  step = both. */
int be_b200_sharded_publish(be_b200_sharded* h, const void* spikes_local, int spk_kind, void* out, void* stream);
int be_b200_sharded_consume(be_b200_sharded* h, void* out, void* stream);
int be_b200_sharded_timed_out(be_b200_sharded* h);
/* Host-buffer pipeline (float32): `slots` timesteps in flight, copy-in / graph of the step / copy-out on three
 * streams.  Host buffers must be pinned and stay untouched until wait() or the slots-th next submit. */
int be_b200_sharded_host_init(be_b200_sharded* h, int slots, int use_graph);
int be_b200_sharded_submit_host(be_b200_sharded* h, const uint8_t* spikes_host, float* out_host);
int be_b200_sharded_wait(be_b200_sharded* h);
int be_b200_sharded_run_host(be_b200_sharded* h, int64_t n_steps, const uint8_t* const* spikes_host, int n_in,
                             float* const* out_host, int n_out_bufs);
void be_b200_sharded_destroy(be_b200_sharded* h);

/* ---- host-buffer session (end-to-end entry: host spikes in, host result out) -----------------
 * A projection keeps its connectivity resident in HBM (uploaded once at creation, like the
 * reference's jax arrays) and per call copies the host spike vector in, runs the path on its
 * own stream and copies the result back, synchronising before it returns. */
typedef struct be_b200_projection be_b200_projection;
/* kind: 0 = CSR (weights[nse]|[1], indices[nse], indptr int32[m+1]), 1 = FCN (indices [m, n_conn]),
 * 2 = JITC (p0, p1, prob -> clen, seed; jit_kind in n_conn's place is NOT used: see jit_kind). */
be_b200_projection* be_b200_projection_create_csr(const float* weights, int homo,
                                                  const int32_t* indices, const int32_t* indptr,
                                                  int64_t m, int64_t k);
be_b200_projection* be_b200_projection_create_fcn(const float* weights, int homo,
                                                  const int32_t* indices, int64_t n_rows,
                                                  int64_t n_conn, int64_t n_post);
/* Same as create_fcn but the connectivity already lives in device memory and stays owned by the
 * caller (no copy); used when the matrix is generated on the device. */
be_b200_projection* be_b200_projection_adopt_fcn(const float* d_weights, int homo,
                                                 const int32_t* d_indices, int64_t n_rows,
                                                 int64_t n_conn, int64_t n_post);
be_b200_projection* be_b200_projection_create_jitc(int jit_kind, float p0, float p1, int32_t clen,
                                                   int32_t seed, int64_t n_pre, int64_t n_post,
                                                   int64_t chunk_size, int corder);
/* events: host uint8[n_pre] (pinned for best speed); out: host float[n_post].  spk @ W. */
int be_b200_projection_apply_host(be_b200_projection* p, const uint8_t* events_host,
                                  float* out_host);
/* Pipelined form of apply_host: submit() enqueues copy-in, the hot path and copy-out on three
 * streams and returns; at most four calls are in flight (the fifth submit first waits for the
 * oldest).  Host buffers must be pinned and stay untouched until wait() (or the fourth-next
 * submit) returns.  PCIe transfers of neighbouring timesteps then overlap the kernels. */
int be_b200_projection_submit_host(be_b200_projection* p, const uint8_t* events_host, float* out_host);
int be_b200_projection_wait(be_b200_projection* p);
void be_b200_projection_destroy(be_b200_projection* p);

#ifdef __cplusplus
}
#endif
#endif /* BRAINEVENT_B200_H_ */
