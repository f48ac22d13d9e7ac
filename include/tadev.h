/*
 * tadev.h — C ABI of libtadev, the B200 (sm_100a) contraction engine that sits behind
 * TiledArray's tile plug-in and evaluator plug-in interfaces.
 *
 * Conventions (all entry points):
 *   - return int status: 0 = TADEV_OK, otherwise a TADEV_E* code; never throw.
 *     tadev_last_error() returns a thread-local message for the last failure.
 *   - every device operation takes an explicit cudaStream_t (passed as void*), is
 *     asynchronous with respect to the host, and is thread-safe for concurrent calls on
 *     distinct streams (the contract TiledArray's MADWorld task threads rely on,
 *     reference: src/TiledArray/external/device.h:847-907).
 *   - tiles are dense row-major FP64 blocks living in device memory (NOT unified memory).
 *   - pointers named d_* are device pointers, h_* are host pointers.
 *   - there is no CPU fallback: without a CUDA device every compute entry returns
 *     TADEV_ENODEVICE. Pure host-logic entries (ProcGrid, pmap, permutation planning,
 *     SUMMA schedule) work without a device; they are marked [host].
 *
 * Each entry cites the reference interface (file:line under ValeevGroup/tiledarray) it replaces.
 */
#ifndef TADEV_H_INCLUDED
#define TADEV_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TADEV_OK 0
#define TADEV_EINVAL 1    /* bad argument (the analogue of TA_ASSERT -> TiledArray::Exception) */
#define TADEV_ECUDA 2     /* CUDA runtime error (reference: DeviceSafeCall, external/device.h:71-97) */
#define TADEV_ENODEVICE 3 /* no CUDA device: there is deliberately no CPU path */
#define TADEV_ENCCL 4     /* NCCL error */
#define TADEV_ENOMEM 5

#define TADEV_OP_N 0 /* blas::Op::NoTrans  (math/blas.h) */
#define TADEV_OP_T 1 /* blas::Op::Trans */

typedef struct tadev_ctx tadev_ctx;
typedef void* tadev_stream; /* cudaStream_t */

const char* tadev_last_error(void);
const char* tadev_version(void);

/* ---- context, streams, memory ------------------------------------------------------
 * replaces device::Env (external/device.h:536-605: rank->device map, N non-blocking
 * streams, Umpire pools) and BLASQueuePool (device/blas.h:66-76). */
int tadev_init(int device, size_t pool_bytes, tadev_ctx** out);
int tadev_finalize(tadev_ctx* ctx);
int tadev_device_count(int* n);
/* number of compute streams owned by the ctx (TA_DEVICE_NUM_STREAMS analogue, default 3) */
int tadev_num_streams(tadev_ctx* ctx, int* n);
int tadev_get_stream(tadev_ctx* ctx, int i, tadev_stream* out);
/* stream_for(range) analogue (external/device.h:899-907): stream = ordinal % nstreams */
int tadev_stream_for(tadev_ctx* ctx, uint64_t ordinal, tadev_stream* out);
int tadev_stream_sync(tadev_ctx* ctx, tadev_stream s);
int tadev_alloc(tadev_ctx* ctx, size_t bytes, void** d_ptr, tadev_stream s);
int tadev_free(tadev_ctx* ctx, void* d_ptr, tadev_stream s);
int tadev_memcpy_h2d(tadev_ctx* ctx, void* d_dst, const void* h_src, size_t bytes, tadev_stream s);
int tadev_memcpy_d2h(tadev_ctx* ctx, void* h_dst, const void* d_src, size_t bytes, tadev_stream s);
int tadev_memset(tadev_ctx* ctx, void* d_dst, int byte, size_t bytes, tadev_stream s);
/* CUDA-event timing on a caller stream (bench.py times on the launching stream) */
int tadev_event_create(tadev_ctx* ctx, void** ev);
int tadev_event_record(tadev_ctx* ctx, void* ev, tadev_stream s);
int tadev_event_elapsed_ms(tadev_ctx* ctx, void* ev_start, void* ev_stop, float* ms); /* syncs on ev_stop */
int tadev_event_destroy(tadev_ctx* ctx, void* ev);
/* Asynchronous completion and cross-stream ordering (the tile plug-in's contract: a tile op enqueues work and
 * returns; the caller's runtime completes its task from a host function on the stream - external/device.h:847-875
 * sync_madness_task_with + madness::add_device_task, reduce_task.h:460-486 - and orders dependent work on other
 * streams with events instead of host synchronisation). Sync events have timing disabled (cheap to record). */
int tadev_sync_event_create(tadev_ctx* ctx, void** ev);
int tadev_stream_wait_event(tadev_ctx* ctx, tadev_stream s, void* ev); /* work enqueued on s later waits for ev */
int tadev_event_query(tadev_ctx* ctx, void* ev, int* done);           /* non-blocking */
int tadev_event_sync(tadev_ctx* ctx, void* ev);                       /* blocks the calling host thread */
typedef void (*tadev_host_fn)(void* user);
/* fn(user) runs on a runtime thread once everything enqueued on s so far has finished (cudaLaunchHostFunc); it
 * must not call CUDA / tadev functions. This is the hook a MADNESS device task completes from. */
int tadev_stream_add_callback(tadev_ctx* ctx, tadev_stream s, tadev_host_fn fn, void* user);
int tadev_memcpy_d2d(tadev_ctx* ctx, void* d_dst, const void* d_src, size_t bytes, tadev_stream s);
int tadev_host_alloc(size_t bytes, void** h_ptr); /* pinned */
int tadev_host_free(void* h_ptr);

/* ---- tile GEMM -------------------------------------------------------------------------
 * replaces Tensor::gemm (tensor/tensor.h:3132-3219) -> detail::gemm (tensor/kernels.h:92-231)
 * -> math::blas::gemm (math/blas.h:171-177) and device::btas::gemm (device/btas.h:52-222),
 * and folds ContractReduce's pair-accumulate + add_to merge (tile_op/contract_reduce.h:386-453)
 * into one launch.
 *
 * A task is one (left tile, right tile) product contributing to a result tile:
 *   C[m x n] (+)= alpha * op(A) * op(B), all row-major, natural leading dimensions
 *   (lda = k|m, ldb = n|k, ldc = n, exactly tensor/kernels.h:146-158).
 * Tasks are grouped by result tile: group g owns tasks [task_begin[g], task_begin[g+1]) which
 * are accumulated IN REGISTERS in list order and written to C once; if accumulate[g] != 0 the
 * previous contents of C are added (beta = 1), otherwise C is overwritten (beta = 0).       */
typedef struct {
  const double* A; /* device */
  const double* B; /* device */
  int32_t k;       /* contracted (fused inner) extent of this pair */
  int32_t reserved;
} tadev_gemm_task;

typedef struct {
  double* C;          /* device, m x n row-major */
  int32_t m, n;       /* fused outer extents */
  int32_t task_begin; /* first task of this group */
  int32_t task_end;   /* one past the last task */
  int32_t accumulate; /* 0: C = alpha*sum ; 1: C += alpha*sum */
  /* optional L2 rasterisation hint: 0 = none, else ((row0+1) << 16) | (col0+1) where (row0, col0)
   * is the origin of this result tile in a global grid of 128x128 blocks of the whole result
   * matrix. When every group of a launch carries a hint, CTAs walk the result in square
   * super-blocks so that operand panels are shared through L2 (see gemm_f64_ws.cu). */
  int32_t raster;
} tadev_gemm_group;

/* h_groups/h_tasks are HOST arrays; the call copies them to the device on `s` (staging is
 * owned by the ctx) and launches one grouped DMMA kernel. */
int tadev_gemm_grouped_f64(tadev_ctx* ctx, tadev_stream s, int opA, int opB, double alpha,
                           const tadev_gemm_group* h_groups, int ngroups,
                           const tadev_gemm_task* h_tasks, int ntasks);
/* Same, but descriptors already live in device memory (generic cp.async kernel; d_tile_prefix[g] = first
 * 128x128 work item of group g, ngroups + 1 entries). The SUMMA driver feeds the GEMM kernels with lists its
 * own device kernels build per window (csrc/tilelist.cu, see tadev_build_tile_lists). */
int tadev_gemm_grouped_f64_dev(tadev_ctx* ctx, tadev_stream s, int opA, int opB, double alpha,
                               const tadev_gemm_group* d_groups, int ngroups,
                               const tadev_gemm_task* d_tasks, const int32_t* d_tile_prefix,
                               int total_cta_tiles);
/* Single-pair convenience with the exact reference signature semantics:
 *   beta == 0: result allocated/overwritten ("gemm(left,right,factor,helper)", tile_interface.h:803)
 *   beta == 1: accumulate ("gemm(result,left,right,factor,helper)", tile_interface.h:825). */
int tadev_gemm_f64(tadev_ctx* ctx, tadev_stream s, int opA, int opB, int m, int n, int k,
                   double alpha, const double* d_A, const double* d_B, double beta, double* d_C);

/* ---- tile permutation --------------------------------------------------------------------
 * replaces detail::permute (tensor/permute.h:118-209) and librett_permute
 * (external/librett.h:81-111). perm is in TiledArray image form (permutation.h:69-79):
 *   out.extent[perm[i]] = extent[i];  out[perm applied to idx] = in[idx].
 * elem_bytes in {4, 8, 16}. */
int tadev_permute(tadev_ctx* ctx, tadev_stream s, int rank, const int64_t* extent,
                  const int32_t* perm, int elem_bytes, const void* d_in, void* d_out);
/* Same permutation applied to `ntiles` tiles of identical extents in ONE launch (every tile of a
 * contraction operand gets the same permutation: dist_eval/array_eval.h:42,170). h_in / h_out are
 * HOST arrays of device pointers. */
int tadev_permute_batched(tadev_ctx* ctx, tadev_stream s, int rank, const int64_t* extent,
                          const int32_t* perm, int elem_bytes, int ntiles, const void* const* h_in,
                          void* const* h_out);
/* result(+)= arg: ContractReduce partial-result merge (contract_reduce.h:397-398; GPU ref
 * device/btas_um_tensor.h:377-384 axpy). */
int tadev_add_to_f64(tadev_ctx* ctx, tadev_stream s, size_t n, double* d_result, const double* d_arg);
int tadev_scale_f64(tadev_ctx* ctx, tadev_stream s, size_t n, double* d_x, double factor);
/* squared Frobenius norms of `ntiles` (<= 65535) tiles (for truncate / true result shapes, dist_array.h:1553;
 * Tensor::squared_norm, tile_interface.h). d_ptrs/d_sizes are device arrays; out is a device array of doubles;
 * max_elems = host-known upper bound of the sizes (launch geometry). Deterministic: a tile's result depends only
 * on its own data and size (fixed 16384-element chunks summed in ascending order). */
int tadev_tile_sqnorms_f64(tadev_ctx* ctx, tadev_stream s, int ntiles, const double* const* d_ptrs,
                           const int64_t* d_sizes, int64_t max_elems, double* d_out);
/* squared Frobenius norm of ONE tile handed to the host (Tensor::squared_norm / norm of the tile interface):
 * blocks until the reduction on s has finished. */
int tadev_sqnorm_f64(tadev_ctx* ctx, tadev_stream s, size_t n, const double* d_x, double* h_out);
/* Batched element-wise tile operations (tile_op/add.h, subt.h, scal.h, mult.h; GPU reference
 * device/btas_um_tensor.h:377-470 and device/kernel/thrust/mult_kernel.h): for every tile t
 *   TADEV_EW_AXPBY: out[t] = alpha * x[t] + beta * y[t]      (add, subt, scale, copy)
 *   TADEV_EW_MULT:  out[t] = alpha * x[t] .* y[t]             (Hadamard product)
 * h_* are HOST arrays of device pointers / element counts; a NULL x or y entry is a zero tile (block-
 * sparse operands whose tile is absent); out may alias x or y. One launch for all tiles. */
#define TADEV_EW_AXPBY 0
#define TADEV_EW_MULT 1
int tadev_tiles_binary_f64(tadev_ctx* ctx, tadev_stream s, int op, int ntiles, double* const* h_out,
                           const double* const* h_x, const double* const* h_y, const int64_t* h_elems,
                           double alpha, double beta);
/* counter-based uniform(-1,1) fill keyed by (seed, global element offset): synthetic inputs
 *  and on-the-fly tile generation for tensors that do not fit in HBM. */
int tadev_fill_uniform_f64(tadev_ctx* ctx, tadev_stream s, double* d_x, size_t n, uint64_t seed,
                           uint64_t offset);

/* ---- SparseShape screening (FP32, bit-exact vs oracle) ----------------------------------
 * replaces SparseShape<float>::scale_tile_norms (sparse_shape.h:149-217), gemm (:1589-1681),
 * perm (:1222), mask (:653-676), is_zero (:495-498). All arrays are device arrays. */
/* norms[i] *= left[i / nright] * right[i % nright]; hard-zero below threshold; count zeros.
 * rank-1 shapes pass nright = 0: norms[i] /= left[i] (the reference's divide branch).     */
int tadev_shape_scale_f32(tadev_ctx* ctx, tadev_stream s, float* d_norms, const float* d_left,
                          int64_t nleft, const float* d_right, int64_t nright, float threshold,
                          uint64_t* d_nzero);
/* out[m,n] = |factor| * sum_k (a[m,k]*ksz[k]) * (b[k,n]*ksz[k]), sequential k, every fp32 multiply and add
 * rounded separately (no FMA contraction: the oracle's fixed order, oracle/ta_oracle.py shape_gemm_kernel);
 * values < threshold are hard-zeroed and counted. Kt == 0: outer product a[m]*b[n]*|factor|. */
int tadev_shape_gemm_f32(tadev_ctx* ctx, tadev_stream s, int Mt, int Nt, int Kt, const float* d_a,
                         const float* d_b, const float* d_ksz, float abs_factor, float threshold,
                         float* d_out, uint64_t* d_nzero);
int tadev_shape_mask_f32(tadev_ctx* ctx, tadev_stream s, int64_t n, float* d_norms,
                         const float* d_mask, float thr_this, float thr_mask, uint64_t* d_nzero);

/* ---- tile-pair list of one SUMMA step ---------------------------------------------------
 * replaces Summa::contract (dist_eval/contraction_eval.h:1311-1384) + get_col/get_row
 * (:655-676): for step k on grid position (r,c) of a Pr x Pc grid, emit, in (i,j) row-major
 * order, the pairs {(i,j): i%Pr==r, j%Pc==c, a[i,k]>=thr, b[k,j]>=thr, c[i,j]>=thr}.
 * NULL norm arrays mean dense. Output is a device int2 list + device count. */
int tadev_build_pairlist(tadev_ctx* ctx, tadev_stream s, int k, int Pr, int Pc, int r, int c, int Mt,
                         int Nt, int Kt, const float* d_a, const float* d_b, const float* d_c,
                         float threshold, int32_t* d_pair_i, int32_t* d_pair_j, int32_t* d_npairs);

/* The list kernel the SUMMA driver runs per window of steps (csrc/tilelist.cu), as a stand-alone entry: for the
 * steps d_ksteps[0..nsteps) and every LOCAL result tile g = li * ncl + lj (i = r + li*Pr, j = c + lj*Pc; nrl x ncl
 * local tiles, row-major) the chained contributions {k : a[i,k] >= thr, b[k,j] >= thr, c[i,j] >= thr} in step
 * order: d_task_k[d_group_begin[g] .. d_group_begin[g+1]). Same predicate as Summa::contract
 * (contraction_eval.h:1311-1384), regrouped per result tile (the ContractReduce chains the grouped GEMM executes).
 * Multi-CTA; all arrays are device arrays; NULL norms = dense. *d_ntasks = total (also d_group_begin[nrl*ncl]). */
int tadev_build_tile_lists(tadev_ctx* ctx, tadev_stream s, int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt,
                           const int32_t* d_ksteps, int nsteps, const float* d_a, const float* d_b, const float* d_c,
                           float threshold, int32_t* d_group_begin, int32_t* d_task_k, int64_t capacity,
                           int32_t* d_ntasks);

/* ---- [host] process grid / pmap / permutation planning ------------------------------------ */
/* ProcGrid (proc_grid.h:97-260): grid dims + this rank's coordinates and local counts.
 * rank >= proc_rows*proc_cols  =>  rank_row = rank_col = -1 and local counts 0. */
typedef struct {
  int32_t proc_rows, proc_cols, proc_size;
  int32_t rank_row, rank_col;
  int64_t local_rows, local_cols, local_size;
} tadev_proc_grid;
int tadev_proc_grid_make(int rank, int nprocs, int64_t rows, int64_t cols, int64_t row_size,
                         int64_t col_size, tadev_proc_grid* out);
/* CyclicPmap::owner (pmap/cyclic_pmap.h:123-134) */
int tadev_cyclic_owner(int64_t tile, int64_t cols, int proc_rows, int proc_cols, int* owner);

/* GEMMPermutationOptimizer (expressions/permopt.h:254-376) + the result permutation of
 * ContEngine::init_struct (expressions/cont_engine.h:354-529). Index lists are comma-separated
 * strings ("i,k,a,c"). Output permutations are in image form; perm_*[0] == -1 means identity. */
typedef struct {
  int32_t left_rank, right_rank, result_rank, inner_rank;
  int32_t opA, opB;          /* TADEV_OP_* handed to the tile GEMM */
  int32_t left_permtype;     /* 1 identity, 2 matrix_transpose, 3 general (permopt.h:39) */
  int32_t right_permtype;
  int32_t perm_left[16];     /* explicit argument-tile permutation (general case) */
  int32_t perm_right[16];
  int32_t perm_result[16];   /* GEMM result order -> target order */
  char left_target[256];     /* target index lists, comma separated */
  char right_target[256];
  char result_gemm[256];
} tadev_contraction_plan;
int tadev_plan_contraction(const char* target, const char* left, const char* right,
                           tadev_contraction_plan* out);
/* Same, but also plans the operand-exchanged product (C^T = B^T A^T) and returns it with
 * *swapped = 1 when it needs fewer explicit tile permutations; the caller then passes the arrays
 * in exchanged order. The reference has no such step (it permutes result tiles instead). */
int tadev_plan_contraction_opt(const char* target, const char* left, const char* right,
                               tadev_contraction_plan* out, int32_t* swapped);

/* [host] GeneralPermutationOptimizer: plan of a general product (fused + contracted + free indices).
 * *nfused = number of fused indices; 0 = not a general product (plan untouched). left_target /
 * right_target / result_gemm receive the canonical index lists (fused..., external..., contracted...). */
int tadev_plan_general_product(const char* target, const char* left, const char* right,
                               tadev_contraction_plan* out, int32_t* nfused);

/* ---- multi-GPU: communicators + SUMMA driver ------------------------------------------------
 * replaces detail::Summa (dist_eval/contraction_eval.h:55-2027) and its world.gop.bcast
 * row/column tile broadcasts (:712,:849,:903) by NCCL broadcasts of packed panels on
 * row/column communicators of a Pr x Pc grid (one process per GPU). */
int tadev_comm_unique_id(void* out128);
/* Creates the world communicator and, by splitting it, a THIN and a WIDE pair of row / column communicators for the
 * Pr x Pc grid (rank r*Pc + c sits at grid position (r, c); ranks >= Pr*Pc stay outside). NCCL is configured without
 * thread-block clusters (cgaClusterSize = 1) and capped to the SMs the persistent GEMM leaves free: 4 for the thin pair
 * (TADEV_SM_RESERVE), 12 for the wide pair (TADEV_SM_RESERVE_WIDE) that tadev_summa_f64 picks for contractions whose
 * panel traffic is large next to their GEMM work (block-sparse; TADEV_WIDE_COMM=0/1 forces the choice). */
int tadev_comm_init(tadev_ctx* ctx, const void* unique_id128, int rank, int nranks, int Pr, int Pc);
int tadev_comm_destroy(tadev_ctx* ctx);
/* Point-to-point redistribution of tiles over the world communicator: this rank sends nsend tiles and
 * receives nrecv. Both lists are derived by every rank from replicated metadata and must be in the same
 * relative order on the two sides of every (sender, receiver) pair (e.g. ascending tile ordinal).
 * h_* are host arrays; pointers are device pointers. One NCCL group. */
int tadev_exchange_tiles(tadev_ctx* ctx, tadev_stream s, int nsend, const void* const* h_src, const size_t* h_sbytes,
                         const int32_t* h_dst_rank, int nrecv, void* const* h_dst, const size_t* h_rbytes,
                         const int32_t* h_src_rank);
/* Shape replication: element-wise max of a device norm array over all ranks (the SparseShape constructor's
 * world.gop.max, sparse_shape.h:416); every rank passes the norms of its own tiles and zeros elsewhere. */
int tadev_shape_allreduce_max_f32(tadev_ctx* ctx, tadev_stream s, float* d_norms, int64_t n);
/* which=0: row communicator (A panels), which=1: column communicator (B panels). */
int tadev_bcast_panel(tadev_ctx* ctx, tadev_stream s, int which, int root, void* d_buf, size_t bytes);

/* Lazy operands (the analogue of TiledArray's lazy tiles: dist_eval/array_eval.h:42 LazyArrayTile,
 * evaluated when the contraction asks for them): the tile table of a lazy operand holds opaque
 * non-zero TOKENS instead of pointers, and the driver asks the provider to materialise the `ntiles`
 * tiles tokens[0..ntiles) at d_dst[0..ntiles) (elems[n] doubles each, row-major in the operand's GEMM
 * layout) by enqueueing work on stream s. Used for argument permutations performed just in time per
 * SUMMA window (no permuted copy of the operand) and for tensors generated on the fly. */
typedef int (*tadev_tile_provider)(void* user, tadev_stream s, int ntiles, const uint64_t* tokens,
                                   double* const* d_dst, const size_t* elems);

typedef struct {
  int32_t Mt, Nt, Kt;       /* fused tile-grid extents of C (Mt x Nt) and the contraction (Kt) */
  const int64_t* m_ext;     /* [Mt] fused element extent of each tile row      (host) */
  const int64_t* n_ext;     /* [Nt]                                              (host) */
  const int64_t* k_ext;     /* [Kt]                                              (host) */
  int32_t opA, opB;         /* storage of argument tiles: N: A(i,k) is m x k; T: k x m */
  double alpha;
  /* replicated scaled norms (host, row-major [Mt,Kt], [Kt,Nt], [Mt,Nt]); NULL = dense */
  const float* a_norms;
  const float* b_norms;
  const float* c_norms;
  float threshold;
  /* device pointers (host arrays of device pointers, full tile grids, row-major ordinals).
   * Only entries owned by this rank under the cyclic maps (proc_grid.h:566-597) are read:
   *   A(i,k): (i%Pr, k%Pc)   B(k,j): (k%Pr, j%Pc)   C(i,j): (i%Pr, j%Pc). */
  const double* const* a_tiles;
  const double* const* b_tiles;
  double* const* c_tiles;
  int32_t accumulate;       /* 0: C = alpha*A*B ; 1: C += */
  int32_t depth;            /* SUMMA pipeline depth (0 = reference default, :1925-1977) */
  int32_t steps_per_launch; /* K steps fused into one grouped-GEMM launch (0 = auto) */
  /* TADEV_SUMMA_*_ON_HOST: the tile table of that array holds PINNED HOST pointers (TiledArray's
   * arrays live in host memory; the reference GPU path migrates tiles through unified memory,
   * device/um_storage.h:52-79). Host operands are streamed to the device panel by panel on a
   * copy stream, overlapped with the GEMM of the previous window; a host result is produced in
   * row blocks, each copied back while the next block computes. Also the out-of-core path. */
  int32_t flags;
  int32_t row_blocks;       /* result row blocks (0 = auto: 1 for device-resident operands) */
  int32_t reserved;
  /* TADEV_SUMMA_*_LAZY: tile providers of lazy operands (see tadev_tile_provider) */
  tadev_tile_provider a_provider;
  void* a_user;
  tadev_tile_provider b_provider;
  void* b_user;
} tadev_summa_plan;
#ifndef TADEV_MEM_DEVICE
#define TADEV_MEM_DEVICE 0
#define TADEV_MEM_HOST 1 /* pinned host memory; streamed through the GPU by the SUMMA driver */
#define TADEV_MEM_LAZY 2 /* never stored: tiles generated on the device when needed (tadev_uniform_source) */
#endif
#define TADEV_SUMMA_A_ON_HOST 1
#define TADEV_SUMMA_B_ON_HOST 2
#define TADEV_SUMMA_C_ON_HOST 4
#define TADEV_SUMMA_A_LAZY 8
#define TADEV_SUMMA_B_LAZY 16

/* built-in providers. token = index + 1 into the source's tables. */
typedef struct {
  tadev_ctx* ctx;
  uint64_t seed;            /* tile `index` is filled by tadev_fill_uniform_f64(seed, offset = index << 32) */
} tadev_uniform_source;
int tadev_provider_uniform(void* uniform_source, tadev_stream s, int ntiles, const uint64_t* tokens,
                           double* const* d_dst, const size_t* elems);
typedef struct {
  tadev_ctx* ctx;
  int32_t rank;
  int32_t perm[16];         /* image form, as tadev_permute */
  const int64_t* extents;   /* [ntable][rank] extents of each source tile (host) */
  const void* const* src;   /* [ntable] source tiles (host array): device pointers, or pinned host pointers */
  /* where the source tiles live: TADEV_MEM_DEVICE (permute straight out of them), TADEV_MEM_HOST (each tile is
   * uploaded into a stream-ordered scratch buffer first) or TADEV_MEM_LAZY (generated into the scratch buffer by
   * tadev_fill_uniform_f64(lazy_seed, offset = ordinals[index] << 32) first; src may be NULL). This is how a
   * host-resident or lazy operand that needs an explicit argument permutation is evaluated tile by tile
   * (the reference's ArrayEvalImpl does the same per tile, dist_eval/array_eval.h:170,330). */
  int32_t src_memory;
  int32_t reserved;
  uint64_t lazy_seed;
  const int64_t* ordinals;  /* [ntable] tile ordinal in the ORIGINAL tiling (lazy sources) */
} tadev_permute_source;
int tadev_provider_permute(void* permute_source, tadev_stream s, int ntiles, const uint64_t* tokens,
                           double* const* d_dst, const size_t* elems);

typedef struct {
  int64_t nsteps, nsteps_skipped, npairs, nlaunches;
  double flops;       /* 2*m*n*k summed over executed pairs on this rank */
  int64_t bcast_bytes; /* bytes this rank sent or received in panel broadcasts */
  float device_ms;    /* CUDA-event time of the whole contraction on this rank */
  int32_t row_blocks; /* result row blocks used */
  int64_t h2d_bytes;  /* host->device bytes moved inside the call (host-resident operands) */
  int64_t d2h_bytes;  /* device->host bytes moved inside the call (host-resident result) */
  int64_t lazy_tiles; /* tiles materialised by providers */
  float gemm_ms;      /* sum of the grouped-GEMM launches' own durations (CUDA events around each launch) */
  float list_ms;      /* sum of the device tile-list builder launches' durations */
} tadev_summa_stats;

int tadev_summa_f64(tadev_ctx* ctx, const tadev_summa_plan* plan, tadev_summa_stats* stats);

/* ---- the contraction engine ------------------------------------------------------------------
 * replaces the expression-layer half of the path: ContEngine::perm_indices / init_struct /
 * init_distribution / make_dist_eval (expressions/cont_engine.h:176-677) and the argument / result
 * tile permutations around Summa (dist_eval/array_eval.h:42,170; contract_reduce.h:370-378).
 * The caller describes the two argument arrays; `create` plans the contraction (optionally
 * exchanging the operands), derives the result tiling, screens the result shape on the device and
 * lays out the local result tiles; `eval` runs it into a caller-provided arena. Indices shared by both
 * arguments AND kept in the target make a general (fused-index, batched) product (cont_engine.h:679-1100,
 * tile_op/batched_contract_reduce.h, SparseShape::gemm_batched): one grouped-GEMM launch, single rank. This is what a
 * TA::DistArray-level binding calls for `c("i,j") = a("i,k") * b("k,j")` (include/tiledarray.hpp,
 * tiledarray_b200/tiledarray.py). */
typedef struct {
  int32_t rank;
  int32_t memory;            /* TADEV_MEM_* */
  const int64_t* bounds;     /* tile boundaries, dimension after dimension: ntiles[d] + 1 entries each */
  const int32_t* ntiles;     /* [rank] */
  const float* norms;        /* SparseShape data (scaled norms, row-major over the tile grid); NULL = dense */
  const void* const* tiles;  /* [prod ntiles] tile pointers; NULL entry = zero or not local. Lazy arrays: NULL
                                table = every non-zero tile is local, else non-NULL entries mark local tiles */
  uint64_t lazy_seed;
  const int32_t* owners;     /* [prod ntiles] rank that holds each tile (the array's process map), or NULL. With
                                owners, a multi-rank contraction redistributes device tiles that are not where SUMMA's
                                cyclic maps need them (what ArrayEvalImpl does tile by tile, dist_eval/array_eval.h:170);
                                without, every operand tile must already be local where it is needed. */
} tadev_array_desc;

typedef struct {
  int32_t exchange_operands;    /* 1: may evaluate C^T = B^T A^T (tadev_plan_contraction_opt) */
  int32_t stream_permutes;      /* 1 / 0 / -1 (auto): permute argument tiles per SUMMA window */
  int64_t stream_permute_bytes; /* auto: stream when the permuted copy would exceed this size */
  int32_t depth, steps_per_launch, row_blocks; /* forwarded to the SUMMA plan (0 = auto) */
  float threshold;              /* SparseShape threshold */
  /* optional result mask: `(a("m,k") * b("k,n")).set_shape(mask)` (expressions/expr.h:116, applied after
   * make_shape by SparseShape::mask, cont_engine.h:526-528, sparse_shape.h:653-676; the reference's own
   * block-sparse example uses it, examples/gemm/ta_sparse.cpp:190). Scaled norms over the TARGET tile grid
   * (host, row-major), read during tadev_contraction_create only; NULL = none. */
  const float* mask_norms;
  float mask_threshold;
} tadev_contract_options;
int tadev_contract_options_default(tadev_contract_options* o);

typedef struct tadev_contraction tadev_contraction;
typedef struct {
  int32_t rank, swapped;          /* result rank; 1 if the operands were exchanged */
  const int64_t* bounds;          /* result tiling in TARGET order (layout as tadev_array_desc) */
  const int32_t* ntiles;
  const float* norms;             /* result shape in target order; NULL = dense */
  uint64_t nzero;                 /* zero tiles in the result shape */
  int64_t nlocal;                 /* local non-zero result tiles */
  const int64_t* ordinals;        /* [nlocal] ordinal in the target tile grid */
  const int64_t* elems;           /* [nlocal] elements */
  const int64_t* offsets;         /* [nlocal] element offset of the tile in the result arena (16-byte aligned) */
  int64_t arena_elems;            /* doubles the result arena must hold */
  int32_t Pr, Pc, Mt, Nt, Kt, opA, opB, needs_result_permute;
} tadev_contraction_info;          /* pointers stay valid until tadev_contraction_destroy */

typedef struct {
  tadev_summa_stats summa;
  float permute_ms;               /* up-front argument permutations + result permutation */
} tadev_contract_stats;

/* [host] the distribution SUMMA wants for the operands of an expression (no device, no tile tables needed):
 * the ProcGrid for `nranks` ranks and, per tile of each operand (original tiling, row-major ordinal), its
 * position (frow, fcol) in the fused GEMM-side tile grid; owner = (frow % Pr) * Pc + fcol % Pc
 * (proc_grid.h:566-597). *_role: 0 = the operand is the GEMM's left operand A(i,k), 1 = right B(k,j) (after
 * the optional operand exchange). Arrays created with these maps are contracted without redistribution, and
 * an arena ordered by (fcol, frow) [role 0] / (frow, fcol) [role 1] makes SUMMA panels contiguous. */
typedef struct {
  int32_t swapped, Pr, Pc, Mt, Nt, Kt, opA, opB, left_role, right_role;
} tadev_contraction_layout_info;
int tadev_contraction_layout(const char* target, const char* left_idx, const char* right_idx,
                             const tadev_array_desc* left, const tadev_array_desc* right,
                             const tadev_contract_options* options /* NULL = defaults */, int nranks,
                             tadev_contraction_layout_info* info, int32_t* left_frow, int32_t* left_fcol,
                             int32_t* right_frow, int32_t* right_fcol);

int tadev_contraction_create(tadev_ctx* ctx, const char* target, const char* left_idx, const char* right_idx,
                             const tadev_array_desc* left, const tadev_array_desc* right, double factor,
                             const tadev_contract_options* options /* NULL = defaults */, tadev_contraction** out);
int tadev_contraction_info_get(const tadev_contraction* c, tadev_contraction_info* info);
int tadev_contraction_owner(const tadev_contraction* c, int64_t target_ordinal, int* owner);
/* result_arena: device (TADEV_MEM_DEVICE) or pinned host (TADEV_MEM_HOST) memory of arena_elems doubles.
 * accumulate != 0: C += (the arena holds the previous result with the same layout). */
int tadev_contraction_eval(tadev_contraction* c, void* result_arena, int result_memory, int accumulate,
                           tadev_contract_stats* stats);
/* Same, into caller-owned tiles: result_tiles[t] = storage of local result tile t (info.ordinals[t]). This is
 * the entry `c("m,n") += a("m,k") * b("k,n")` uses: the existing array's own tile pointers, whatever its
 * arena layout (a result permutation, if any, is applied to the product before it is added). */
int tadev_contraction_eval_tiles(tadev_contraction* c, void* const* result_tiles, int result_memory, int accumulate,
                                 tadev_contract_stats* stats);
int tadev_contraction_destroy(tadev_contraction* c);

/* ---- the element-wise engine -----------------------------------------------
 * c(target) = alpha * a(a_idx) [ + beta * b(b_idx)   (TADEV_EW_AXPBY; b == NULL: scale / copy / permute)
 *                              | .* b(b_idx)          (TADEV_EW_MULT: Hadamard product) ]
 * replaces AddEngine / SubtEngine / ScalEngine / Hadamard MultEngine (expressions/add_engine.h,
 * subt_engine.h, scal_engine.h, mult_engine.h). Result shape = SparseShape::scale / add / mult
 * (sparse_shape.h:1243-1563). Same create / info / eval / destroy protocol as the contraction
 * engine (the info struct is shared; grid and op fields are unused). Device-resident arrays. */
typedef struct tadev_elementwise tadev_elementwise;
int tadev_elementwise_create(tadev_ctx* ctx, int op, const char* target, double alpha, const char* a_idx,
                             const tadev_array_desc* a, double beta, const char* b_idx /* or NULL */,
                             const tadev_array_desc* b /* or NULL */, float threshold, tadev_elementwise** out);
int tadev_elementwise_info_get(const tadev_elementwise* e, tadev_contraction_info* info);
int tadev_elementwise_eval(tadev_elementwise* e, void* result_arena, float* ms /* device time, may be NULL */);
int tadev_elementwise_destroy(tadev_elementwise* e);

/* [host] the schedule the driver will execute, for inspection/tests: per step the root grid
 * column/row and the pair list of this rank. Arrays are caller-allocated with capacities.
 * pairs are (i,j) global tile coordinates in row-major order. */
int tadev_summa_schedule(int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt, const float* a_norms,
                         const float* b_norms, const float* c_norms, float threshold,
                         int32_t* step_k, int32_t* step_pair_begin, int32_t* nsteps_out,
                         int32_t* pair_i, int32_t* pair_j, int64_t pair_capacity, int64_t* npairs_out);

/* [host] the communication side of the same schedule: every step in which this rank takes part
 * in a broadcast or contracts. flags: bit0 contract, bit1 A panel broadcast along my grid row
 * (root column k % Pc), bit2 B panel broadcast along my grid column (root row k % Pr).
 * a_rows / b_cols are the concatenated panel contents (global tile rows of A(:,k) with i%Pr==r,
 * global tile cols of B(k,:) with j%Pc==c, non-zero only); *_begin have nsteps+1 entries. */
int tadev_summa_steps(int Pr, int Pc, int r, int c, int Mt, int Nt, int Kt, const float* a_norms,
                      const float* b_norms, const float* c_norms, float threshold, int32_t* step_k,
                      int32_t* step_flags, int32_t* a_begin, int32_t* a_rows, int32_t* b_begin,
                      int32_t* b_cols, int32_t* nsteps_out);

/* [host] the panel broadcasts tadev_summa_f64 issues for `plan` at grid position (r, c), in issue
 * order: comm[n] = 0 row / 1 column communicator, group[n] = ordinal of the NCCL group the call is
 * part of, k[n] = SUMMA step, root[n], bytes[n]. Only the host-side fields of the plan are read
 * (extents, norms, flags, steps_per_launch, row_blocks). NCCL may reorder calls inside a group, so
 * all ranks of a communicator must produce the same calls in the same groups — checked on the CPU
 * by tests/test_host_logic.py for every grid position. Pass NULL arrays to query the count. */
int tadev_summa_comm_trace(int Pr, int Pc, int r, int c, const tadev_summa_plan* plan, int32_t* comm,
                           int32_t* group, int32_t* k, int32_t* root, int64_t* bytes, int64_t capacity,
                           int64_t* n_out);

/* ---- measurement helpers -------------------------------------------------------------------
 * Register-resident DMMA / DFMA issue-rate probes: the FP64 roofline denominator is not in
 * MEASURED_PEAKS.json and must be measured on the box. Returns TFLOP/s. */
int tadev_probe_fp64_peak(tadev_ctx* ctx, int kind /*0 dmma, 1 dfma, 2 both*/, int iters,
                          double* tflops, float* ms);
int tadev_probe_copy_gbs(tadev_ctx* ctx, size_t bytes, int iters, double* gbs);
/* pinned-host <-> device copy rates (GB/s): each direction alone and both at once */
int tadev_probe_pcie_gbs(tadev_ctx* ctx, size_t bytes, double* h2d, double* d2h, double* h2d_bidir,
                         double* d2h_bidir);
/* number of kernels this library launched since init (bench.py's gpu_launches) */
int tadev_launch_count(tadev_ctx* ctx, int64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* TADEV_H_INCLUDED */
