/*
 * rayuela_b200.h -- C ABI of librayuela_b200.so: the B200 (sm_100a) implementation of Rayuela.jl's two
 * data-parallel hot paths.  Plain pointers and sizes only; every entry point cites the reference
 * interface it replaces (paths relative to the Rayuela.jl tree).
 *
 * Array layouts are the memory images Julia's ccall passes (column-major):
 *   X   d-by-n  float32  -> X[l*d + t]
 *   C   d-by-(m*h) float32 (hcat(C...)) -> C[(j*h + c)*d + t]
 *   B   m-by-n  uint8, 0-based -> B[l*m + k]
 *   dists / idx  k-by-nq -> out[q*k + r]
 * There is no CPU fallback anywhere in this library: every call needs a CUDA device and fails with
 * RAYUELA_ERR_CUDA (or aborts, for the void compat symbols) when there is none.
 */
#ifndef RAYUELA_B200_H_
#define RAYUELA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status ------------------------------------------------------------------------------------- */
#define RAYUELA_OK 0
#define RAYUELA_ERR_ARG (-1)   /* bad shape / unsupported parameter (e.g. h > 256: codes are bytes here) */
#define RAYUELA_ERR_CUDA (-2)  /* CUDA runtime error (incl. no device) */
#define RAYUELA_ERR_OOM (-3)

/* flags */
#define RAYUELA_DEVICE_PTRS 1u /* all array arguments are device pointers on the current device; X and C must be 16-byte
                                * aligned when d % 4 == 0 (any cudaMalloc / CuArray / torch allocation is) */
#define RAYUELA_FAST_UNARIES 2u /* rayuela_encode_icm / rayuela_get_unaries: OPT-IN tensor-core unaries.  The one dense
                                 * contraction of the path, -2*C'X (CUBLAS sgemm in the reference, src/LSQ_GPU.jl:74), runs as
                                 * a tcgen05 bf16x3 GEMM (hi*hi + hi*lo + lo*hi, fp32 accumulation in TMEM) instead of the exact
                                 * fp32 kernel: unaries within 2^-15 * sum|x_t c_t| of the exact ones, so codes are no longer
                                 * bit-identical to the oracle (near-ties may flip; qerror agrees to ~1e-5).  Needs d <= 128;
                                 * otherwise, and by default, the exact kernel runs.  Also: RAYUELA_B200_FAST_UNARIES=1. */

#define RAYUELA_FAST_LUT 4u     /* rayuela_index_search, LSQ scan: OPT-IN tensor-core lookup tables.  The per-query m*256 table
                                 * -2<q,c> (the dense Q x C' contraction, pairwise_byte.cpp:42-49) is built by the same tcgen05
                                 * bf16x3 GEMM; the byte-code scan, the norm add and the top-k stay exact given that table.
                                 * Distances within ~1e-5 relative of the exact ones; ids may differ on near-ties.  Needs
                                 * d <= 128.  Also: RAYUELA_B200_FAST_LUT=1.  Off by default. */

/* Last error message of the calling thread ("" if none). The reference has no error channel at all
 * (void symbols, deps/src/*.cpp extern blocks); Julia-side checks are error() strings. */
const char* rayuela_last_error(void);

/* Select the CUDA device used by subsequent calls from this thread (default: current device).
 * The reference hard-codes device 0 (src/LSQ_GPU.jl:41,45). */
int rayuela_set_device(int device);
/* Number of kernels this library has launched since load (bench.py's gpu_launches claim). */
uint64_t rayuela_launch_count(void);

/* Multi-GPU inside ONE process -- how a single Julia process drives the whole base (encode_icm_cuda is called once
 * for all of it, src/LSQ_GPU.jl:218-264, with `nsplits` serial chunks through device 0, :41,45,236-255).
 * rayuela_init(devices, n) configures a device set (a device may be listed twice: two shards on one GPU); without it
 * the list is read once from the environment variable RAYUELA_B200_DEVICES ("0,1,2,3").  With more than one slot,
 * every HOST-pointer call of rayuela_encode_icm splits the base in contiguous splitarray slices (src/utils.jl:179-203),
 * one per slot, encoded concurrently (codes bit-identical to the single-device call: vectors are independent and the
 * RNG is keyed on the global index), and rayuela_index_create builds a base-sharded index whose rayuela_index_search
 * scans every shard for all queries, copies the per-shard top-k peer-to-peer to the first slot and merges them by the
 * (dist, id) total order (ids and distances identical to the single-device search).  Device-pointer calls always run
 * on the pointers' device.  n_devices <= 1 (or never calling it, with the variable unset) = single-device mode.
 * rayuela_shutdown releases the per-slot streams; free multi-device indexes first. */
int rayuela_init(const int* devices, int n_devices);
int rayuela_shutdown(void);
int rayuela_device_count(void); /* slots of the configured set (1 in single-device mode) */

/* ---- path (1): LSQ / LSQ++ ICM-ILS encoding ------------------------------------------------------ */

/* Replaces encode_icm_fully! (src/LSQ.jl:152-252) together with its callee loop
 * iterated_conditional_modes_cpp! (src/LSQ.jl:42-80) / `condition` (deps/src/encode_icm.cpp:3-61), i.e.
 * everything encoding_icm (src/LSQ.jl:272-294) and encode_icm_cuda_single (src/LSQ_GPU.jl:4-216) do:
 * unaries, pairwise tables, ilsiter x {perturb, icmiter x m conditioning steps, cost, strict-< accept}.
 *   B          in/out: initial codes, overwritten with the result (oldB semantics, src/LSQ.jl:248)
 *   g0         global index of the first vector (RNG is keyed on the global index, so any sharding of
 *              the base set across calls / GPUs gives identical codes)
 *   orders     ilsiter-by-m visiting orders (0-based, host memory) or NULL: randord ? Philox randperm(seed)
 *              : identity  (src/LSQ.jl:218-221)
 *   snap_iters n_snap ILS iteration counts (1-based, host memory) at which B is copied to B_snap[s] and
 *              qerror to objs[s] (host memory) -- encode_icm_cuda's `ilsiters` (src/LSQ_GPU.jl:193-204)
 *   cost_out   n floats or NULL: final veccost per vector
 *   stats      2*ilsiter ints (host memory) or NULL: (#equal, #better) per ILS iteration, the figures
 *              the reference prints at src/LSQ.jl:243-245
 *   h          256 takes the tuned kernels (the reference's cpp=true path, src/LSQ.jl:42-80, is 256-only, :173-175);
 *              1..255 takes plain exact kernels with the same arithmetic contract -- the reference's cpp=false path
 *              iterated_conditional_modes! (src/LSQ.jl:83-149), which works for any h (single device, C is d-by-(m*h))
 */
int rayuela_encode_icm(const float* X, const float* C, uint8_t* B, int64_t n, int d, int m, int h,
                       int ilsiter, int icmiter, int npert, int randord, uint64_t seed, int64_t g0,
                       const int* orders, const int* snap_iters, int n_snap, uint8_t* B_snap, float* objs,
                       float* cost_out, int* stats, unsigned flags, void* stream);

/* Conditioning steps (n * ilsiter * icmiter * m in the reference, src/LSQ.jl:64-78) actually executed by the last
 * rayuela_encode_icm call of this thread that requested `stats`, and that total: a step whose conditioning codes
 * did not change since its last evaluation is skipped (its result is provably unchanged). Measurement aid. */
int rayuela_encode_icm_steps(uint64_t* executed, uint64_t* total);

/* Of the executed steps of that call, how many took the exact fp32 rows: a step is first evaluated from a 14-bit
 * quantised copy of the tables with a rigorous error window; steps with more than one candidate inside the window
 * (near-ties) evaluate the exact fp32 chains of those candidates, and only with more than 4 (m > 8: 8) of them, or
 * when the integer fields cannot hold the vector's unaries, re-read the whole fp32 rows -- the result is bit-identical
 * either way. Measurement aid. */
int rayuela_encode_icm_exact_steps(uint64_t* exact);

/* CUDA-event timings (ms) of the last rayuela_encode_icm call of this thread that requested `stats`:
 * {tables + setup, unaries (K1), ICM / ILS kernel (K3), whole call on the device}.  The reference times the same
 * phases with time_ns() around host calls (src/LSQ_GPU.jl:50-55,58,213).  With several chunks on alternating streams the
 * phases overlap, so unaries + ICM may exceed the whole-call figure.  Single-device calls only. */
int rayuela_encode_icm_timings(float* ms4);

/* Replaces get_unaries (src/utils.jl:121-149): U[l][j][c] = -2<C_j[:,c], x_l> + ||C_j[:,c]||^2, written vector-major
 * as n-by-(m*h) (the reference keeps m separate h-by-n matrices; U[l*m*h + j*h + c] is its unaries[j][c, l]).  The
 * encoder computes these internally; the entry point exists for callers that want the table (e.g. ChainQ-style
 * encoders) and for validating the RAYUELA_FAST_UNARIES mode against the exact kernel. */
int rayuela_get_unaries(const float* X, const float* C, int64_t n, int d, int m, int h, float* U, unsigned flags,
                        void* stream);

/* Replaces veccost (src/qerrors.jl:36-66).  mean_out (host double, may be NULL) receives qerror
 * (src/qerrors.jl:69-74). cost may be NULL when only the mean is wanted. */
int rayuela_veccost(const float* X, const uint8_t* B, const float* C, int64_t n, int d, int m, int h,
                    float* cost, double* mean_out, unsigned flags, void* stream);

/* Exact-signature replacement of the reference symbol `condition` (deps/src/encode_icm.cpp:157-168), called
 * at src/LSQ.jl:71-75.  Host pointers.  Kept for drop-in completeness; it is the wrong granularity for a
 * GPU (one call per codebook step with an n*256 float host scratch) -- use rayuela_encode_icm. */
void condition(unsigned char* B, float* ub, float* binaries, float* binaries_t, int* cbpair2binaryidx,
               int* to_condition, int j, int n, int m);

/* ---- path (2): asymmetric-distance linear scan --------------------------------------------------- */

/* Exact-signature replacements of the reference symbols (host pointers, synchronous):
 *   linscan_aqd_query                deps/src/linscan_aqd.cpp:107-113, called at src/Linscan.jl:19-23 (PQ/OPQ; 0-based ids)
 *   linscan_aqd_query_extra_byte     deps/src/linscan_aqd_pairwise_byte.cpp:181-188, src/Linscan.jl:135-141 (LSQ; 1-based ids)
 *   linscan_aqd_cq_query_extra_byte  deps/src/linscan_aqd_pairwise_byte.cpp:190-197, src/Linscan.jl:173-179 (CQ; 1-based ids) */
void linscan_aqd_query(float* dists, unsigned int* res, unsigned char* codes, float* centers, float* queries,
                       int N, unsigned int NQ, int B, int K, int dim1codes, int dim1queries, int subdim);
void linscan_aqd_query_extra_byte(float* dists, int* idx, unsigned char* codes, float* queries,
                                  float* codebooks, float* dbnorms, int nqueries, int ncodes, int m, int h,
                                  int d, int nn);
void linscan_aqd_cq_query_extra_byte(float* dists, int* idx, unsigned char* codes, float* queries,
                                     float* codebooks, int nqueries, int ncodes, int m, int h, int d, int nn);

/* Handle API: the encoded base set is uploaded / re-laid-out once and scanned many times. */
#define RAYUELA_SCAN_LSQ 0 /* lut = -2<q,c>,  + dbnorms[i], ids 1-based (pairwise_byte.cpp:14-94) */
#define RAYUELA_SCAN_CQ 1  /* lut = ||q-c||^2, ids 1-based            (pairwise_byte.cpp:97-176) */
#define RAYUELA_SCAN_PQ 2  /* lut = per-subspace ||c-q||^2, ids 0-based (linscan_aqd.cpp:37-102) */
typedef struct rayuela_index rayuela_index;

/* id_offset is added to every returned id (global ids for a base shard of a multi-GPU index).
 * h: entries per codebook, 1..256 like the reference's scans (pairwise_byte.cpp takes h; codes are bytes). */
int rayuela_index_create(rayuela_index** out, int kind, const uint8_t* codes, const float* dbnorms, int64_t n,
                         int m, int h, int64_t id_offset, unsigned flags, void* stream);
/* codebooks: d-by-(m*h) (LSQ/CQ) or sub-by-h-by-m (PQ, d = m*sub).  dists/idx: k-by-nq.
 * Lookup-table entries must be finite and below 1e37 in magnitude (inf / NaN in the queries or codebooks, or
 * overflowing products): otherwise host-pointer calls fail with RAYUELA_ERR_ARG and device-pointer calls (which
 * cannot report without synchronising) return NaN distances and id -1 for every query. */
int rayuela_index_search(rayuela_index* ix, const float* queries, const float* codebooks, int nq, int d, int k,
                         float* dists, int32_t* idx, unsigned flags, void* stream);
int rayuela_index_free(rayuela_index* ix);

/* Merge S per-shard result lists (each k-by-nq, sorted by (dist, id)) into the global top-k by the same
 * total order std::partial_sort uses on pair<float,int> (pairwise_byte.cpp:82).  Used after the all-gather
 * of per-GPU results; in/out layouts [S][nq][k] and [nq][k].  Any S and k: up to 16384 keys per query are sorted
 * in shared memory, larger merges (2 shards at the reference's default k = 10000) run as a tree of pairwise
 * rank merges in global memory (needs S*nq*k*12 bytes of scratch). */
int rayuela_topk_merge(const float* dists_in, const int32_t* idx_in, int S, int nq, int k, float* dists_out,
                       int32_t* idx_out, unsigned flags, void* stream);

/* ---- PQ / OPQ encode ------------------------------------------------------------------------------ */
/* Replaces quantize_pq (src/PQ.jl:18-48); quantize_opq (src/OPQ.jl:19-27) is this on R'X.
 * Cpq: sub-by-h-by-m (cat(C..., dims=3)); B out m-by-n uint8 0-based. */
int rayuela_quantize_pq(const float* X, const float* Cpq, int64_t n, int d, int m, int h, uint8_t* B,
                        unsigned flags, void* stream);

/* ---- ChainQ Viterbi encode ("next" row 3) ----------------------------------------------------------------- */
/* Replaces quantize_chainq (src/ChainQ.jl:287-348): unaries, chain tables 2*C[i]'*C[i+1], exact min-sum Viterbi
 * over the m-chain (first-minimum ties as src/ChainQ.jl:97-110 / deps/src/encode_icm.cpp:108-118), back-trace.
 * B out: m-by-n uint8, 0-based. */
int rayuela_quantize_chainq(const float* X, const float* C, int64_t n, int d, int m, int h, uint8_t* B,
                            unsigned flags, void* stream);
/* Exact-signature replacement of the reference symbol (deps/src/encode_icm.cpp:170-178; src/ChainQ.jl:26-28):
 * unaries = vcat(unaries...) ((m*256)-by-n), binaries = hcat(binaries...) (256-by-256*(m-1)); host pointers. */
void viterbi_encoding(unsigned char* B, float* unaries, float* binaries, int n, int m);

/* ---- norm quantization ("next" row 2) ------------------------------------------------------------------- */
/* Replaces quantize_norms (src/utils.jl:29-59): norms_out[i] = ||sum_k C_k[:, b_k]||^2 (reconstruct + sequential
 * sum of squares) and norm_codes[i] = first-minimum argmin_c (norm - cbnorms[c])^2 over the 256 norm centroids
 * (0-based).  cbnorms / norm_codes may be NULL to get the norms only (what get_norms_codebook, :4-26, clusters). */
int rayuela_quantize_norms(const uint8_t* B, const float* C, const float* cbnorms, int64_t n, int d, int m, int h,
                           uint8_t* norm_codes, float* norms_out, unsigned flags, void* stream);

/* ---- codebook update, data-parallel half ("next" row 1) ------------------------------------------------- */
/* Replaces fast_bin_matmul (src/codebook_update.jl:96-171), the O(n) part of update_codebooks_fast_bin
 * (:175-204): A = B'B + rho*I ((m*h)-by-(m*h) double, symmetric) and b = B'X' ((m*h)-by-d double, column-major).
 * A holds exact counts; b is accumulated in Float64 in ascending vector order like the reference, so both are
 * bit-identical to it.  The dense solve (LAPACK getrf!/getrs!, :193-196) stays with the caller. */
int rayuela_fast_bin_matmul(const float* X, const uint8_t* B, int64_t n, int d, int m, int h, double rho,
                            double* A, double* b, unsigned flags, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* RAYUELA_B200_H_ */
