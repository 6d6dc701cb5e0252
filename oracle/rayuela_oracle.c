/*
 * rayuela_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, gcc) of the two Rayuela.jl hot paths, used as the
 * parity checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under rayuela.jl_b200/ may link, import or call
 * this file: the product path is CUDA-only and fails loudly without its extension.
 *
 * Pinning status (see DESIGN.md "Oracle"):
 *   - linscan_*      : pinned bit-for-bit against the reference's own C++
 *                      (deps/src/linscan_aqd*.cpp compiled unmodified into oracle/_ref/).
 *   - condition      : pinned bit-for-bit against the reference's own C++
 *                      (deps/src/encode_icm.cpp compiled unmodified into oracle/_ref/).
 *   - Julia host half (get_unaries/get_binaries/veccost/perturb_codes!/encode_icm_fully!,
 *     quantize_pq)   : PARITY UNPINNED -- the reference holds no golden vectors or tests
 *                      for these (test/runtests.jl runs only xvecs + chainq), Julia is not
 *                      installed, the arithmetic lives in OpenBLAS / @simd loops whose
 *                      association order is unspecified, and the RNG is Julia's global
 *                      MersenneTwister.  This file fixes ONE order (sequential fmaf chains
 *                      for GEMM-like sums, sequential unfused sums for veccost) and a
 *                      counter-based Philox4x32-10 RNG with the reference's distributions.
 *
 * All matrices follow the layouts Julia's ccall would pass (column-major):
 *   X  d-by-n   -> X[l*d + t]            C  d-by-(m*h) -> C[(j*h + c)*d + t]
 *   B  m-by-n   -> B[l*m + k] (uint8, 0-based)
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC (see oracle/Makefile).
 * -ffp-contract=off matters: every fused operation below is an explicit fmaf().
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define H256 256

/* ------------------------------------------------------------------------------------------
 * Philox4x32-10 (Salmon et al., SC'11).  Replaces Julia's MersenneTwister draws at
 * src/LSQ.jl:18,30,219 -- same distributions, reproducible stream that is a pure function of
 * (seed, ILS iteration, GLOBAL vector index), so results do not depend on sharding.
 * ---------------------------------------------------------------------------------------- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; r++) {
    if (r > 0) { k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  }
}

static inline uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }

void orc_philox(uint32_t* ctr4, uint32_t k0, uint32_t k1) { philox4x32_10(ctr4, k0, k1); }

/* randperm(m) of src/LSQ.jl:219 -- Fisher-Yates driven by Philox words.
 * counter = {it, 0, 0, 0xFFFFFFFF - block}.  perm is 0-based. */
void orc_randperm(uint64_t seed, int it, int m, int* perm) {
  uint32_t w[4]; int have = 0, blk = 0;
  for (int i = 0; i < m; i++) perm[i] = i;
  for (int i = m - 1; i >= 1; i--) {
    if (have == 0) {
      w[0] = (uint32_t)it; w[1] = 0; w[2] = 0; w[3] = 0xFFFFFFFFu - (uint32_t)blk;
      philox4x32_10(w, (uint32_t)seed, (uint32_t)(seed >> 32));
      have = 4; blk++;
    }
    uint32_t r = w[4 - have]; have--;
    int j = (int)mulhi32(r, (uint32_t)(i + 1));
    int t = perm[i]; perm[i] = perm[j]; perm[j] = t;
  }
}

/* perturb_codes! (src/LSQ.jl:5-39, replace=true): npert positions drawn uniformly from the m
 * codebooks WITH replacement, npert values drawn uniformly from the h entries, applied in
 * order j = 1..npert so a later draw on the same position wins (src/LSQ.jl:32-36).
 * g0 = global index of B's first vector. */
void orc_perturb_codes(uint8_t* B, int64_t n, int m, int h, int npert, uint64_t seed, int it, int64_t g0) {
  for (int64_t l = 0; l < n; l++) {
    uint64_t g = (uint64_t)(g0 + l);
    for (int jb = 0; jb * 4 < npert; jb++) {
      uint32_t pw[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)it, (uint32_t)jb};
      uint32_t vw[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)it, 0x80000000u | (uint32_t)jb};
      philox4x32_10(pw, (uint32_t)seed, (uint32_t)(seed >> 32));
      philox4x32_10(vw, (uint32_t)seed, (uint32_t)(seed >> 32));
      for (int t = 0; t < 4 && jb * 4 + t < npert; t++) {
        int pos = (int)mulhi32(pw[t], (uint32_t)m);
        int val = (int)mulhi32(vw[t], (uint32_t)h);
        B[l * m + pos] = (uint8_t)val;
      }
    }
  }
}

/* ------------------------------------------------------------------------------------------
 * get_unaries (src/utils.jl:121-149): unaries[j] = -2*C[j]'*X .+ diag(C[j]'*C[j]).
 * The reference uses OpenBLAS sgemm (order unpinned); fixed here to a sequential-t fmaf chain
 * starting from +0, then one rounding for (-2*dot + ||c||^2)  (-2*dot is exact).
 * Output U[j][l][c]  (m matrices, each h-by-n column-major, as the reference holds them).
 * ---------------------------------------------------------------------------------------- */
static inline float dot_seq(const float* a, const float* b, int d) {
  float s = 0.0f;
  for (int t = 0; t < d; t++) s = fmaf(a[t], b[t], s);
  return s;
}

void orc_codebook_sqnorms(const float* C, int d, int mh, float* nrm) {
  for (int e = 0; e < mh; e++) nrm[e] = dot_seq(C + (size_t)e * d, C + (size_t)e * d, d);
}

void orc_get_unaries(const float* X, const float* C, int64_t n, int d, int m, int h, float* U) {
  int mh = m * h;
  float* nrm = (float*)malloc(sizeof(float) * mh);
  orc_codebook_sqnorms(C, d, mh, nrm);
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < n; l++) {
    const float* x = X + l * d;
    for (int j = 0; j < m; j++) {
      float* u = U + ((size_t)j * n + l) * h;
      for (int c = 0; c < h; c++) {
        float dt = dot_seq(C + ((size_t)j * h + c) * d, x, d);
        u[c] = -2.0f * dt + nrm[j * h + c];
      }
    }
  }
  free(nrm);
}

/* get_binaries (src/utils.jl:152-171): binaries[idx] = 2*C[i]'*C[j] for i<j, idx enumerating
 * (i,j) row-major over i<j; each h-by-h column-major: bin[idx][b*h + a] = 2*<C_i[:,a], C_j[:,b]>.
 * binaries_t[idx] = transpose (src/LSQ.jl:180-183).  cbi (2-by-ncbi, 1-based in Julia) is
 * returned 0-based as cbi[2*idx+0]=i, cbi[2*idx+1]=j. */
void orc_get_binaries(const float* C, int d, int m, int h, float* bin, float* bin_t, int* cbi) {
  int idx = 0;
  for (int i = 0; i < m; i++)
    for (int j = i + 1; j < m; j++, idx++) {
      cbi[2 * idx] = i; cbi[2 * idx + 1] = j;
      float* t = bin + (size_t)idx * h * h;
      float* tt = bin_t + (size_t)idx * h * h;
#pragma omp parallel for schedule(static)
      for (int b = 0; b < h; b++)
        for (int a = 0; a < h; a++) {
          float v = 2.0f * dot_seq(C + ((size_t)i * h + a) * d, C + ((size_t)j * h + b) * d, d);
          t[(size_t)b * h + a] = v;
          tt[(size_t)a * h + b] = v;
        }
    }
}

/* condition (deps/src/encode_icm.cpp:3-61, C symbol at :157-168): one ICM step for codebook j.
 * Same signature as the reference symbol so the two are interchangeable in tests.
 * Loop variables are declared inside the parallel loop (the reference shares k/binariidx/bb/ubi
 * across threads, a latent race that -O3 hides). */
void orc_condition(unsigned char* B, float* ub, float* binaries, float* binaries_t,
                   int* cbpair2binaryidx, int* to_condition, int j, int n, int m) {
  const int H = H256;
#pragma omp parallel for schedule(static)
  for (int l = 0; l < n; l++) {
    float* u = ub + (size_t)l * H;
    for (int kidx = 0; kidx < m - 1; kidx++) {
      int k = to_condition[kidx];
      int bidx = cbpair2binaryidx[j * m + k];
      const float* bb = (j < k ? binaries : binaries_t) + (size_t)H * H * bidx;
      const float* row = bb + (size_t)B[(size_t)l * m + k] * H;
      for (int ll = 0; ll < H; ll++) u[ll] += row[ll];
    }
    float minv = u[0]; int mini = 0;
    for (int k = 1; k < H; k++) { float v = u[k]; if (v < minv) { minv = v; mini = k; } }
    B[(size_t)l * m + j] = (unsigned char)mini;
  }
}

/* The conditioning loop of iterated_conditional_modes! (src/LSQ.jl:99-142), the pure-Julia twin of `condition` that
 * the reference runs with cpp=false (e.g. experiment_lsq, src/LSQ.jl:431) and that works for ANY h: same absorb order
 * (k ascending as listed in to_condition), same first-minimum scan. */
void orc_condition_h(unsigned char* B, float* ub, float* binaries, float* binaries_t,
                     int* cbpair2binaryidx, int* to_condition, int j, int n, int m, int h) {
#pragma omp parallel for schedule(static)
  for (int l = 0; l < n; l++) {
    float* u = ub + (size_t)l * h;
    for (int kidx = 0; kidx < m - 1; kidx++) {
      int k = to_condition[kidx];
      int bidx = cbpair2binaryidx[j * m + k];
      const float* bb = (j < k ? binaries : binaries_t) + (size_t)h * h * bidx;
      const float* row = bb + (size_t)B[(size_t)l * m + k] * h;
      for (int ll = 0; ll < h; ll++) u[ll] += row[ll];
    }
    float minv = u[0]; int mini = 0;
    for (int k = 1; k < h; k++) { float v = u[k]; if (v < minv) { minv = v; mini = k; } }
    B[(size_t)l * m + j] = (unsigned char)mini;
  }
}

/* veccost (src/qerrors.jl:36-66): CB = sum_k C_k[:,b_k] accumulated k = 1..m from +0, then
 * cost = sum_t (CB[t]-x[t])^2.  The reference's @simd lets LLVM reassociate; fixed here to the
 * sequential, unfused order of the source text. */
void orc_veccost(const float* X, const uint8_t* B, const float* C, int64_t n, int d, int m, int h, float* cost) {
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < n; l++) {
    float acc = 0.0f;
    for (int t = 0; t < d; t++) {
      float cb = 0.0f;
      for (int k = 0; k < m; k++) cb += C[((size_t)k * h + B[l * m + k]) * d + t];
      float df = cb - X[l * d + t];
      float sq = df * df;
      acc += sq;
    }
    cost[l] = acc;
  }
}

/* qerror (src/qerrors.jl:69-74) = mean(veccost).  Julia's mean over Float32 uses pairwise
 * summation in Float32; the exact association is unpinned, so the oracle accumulates in double
 * and rounds once (tests compare qerror within 1e-4 relative, north_star). */
double orc_qerror(const float* X, const uint8_t* B, const float* C, int64_t n, int d, int m, int h) {
  float* cost = (float*)malloc(sizeof(float) * n);
  orc_veccost(X, B, C, n, d, m, h, cost);
  double s = 0; for (int64_t l = 0; l < n; l++) s += cost[l];
  free(cost);
  return s / (double)n;
}

typedef void (*condition_fn)(unsigned char*, float*, float*, float*, int*, int*, int, int, int);

/* encode_icm_fully! (src/LSQ.jl:152-252) + iterated_conditional_modes_cpp! (src/LSQ.jl:42-80).
 *   B        m-by-n uint8 0-based, in/out (plays oldB: mutated in place, src/LSQ.jl:248)
 *   orders   ilsiter-by-m visiting orders (0-based) or NULL -> randord ? orc_randperm : identity
 *   step     the ICM step function: orc_condition, or the reference's own `condition` symbol
 *   snap_iters/n_snap/B_snap/objs : encode_icm_cuda's ilsiters snapshots (src/LSQ_GPU.jl:193-204)
 *   cost_out n floats or NULL: final per-vector cost
 *   stats    2*ilsiter ints or NULL: (#equal, #better) per ILS iteration (src/LSQ.jl:239-245)
 * The per-step copyto!(ub, unaries[j]) of src/LSQ.jl:69 is kept, so timing this function is
 * timing the reference's CPU algorithm, not a flattered variant. */
/* _ex: the same loop with two TIMING-ONLY extras for bench.py's CPU arm (parity tests never pass them):
 *   U_pre / bin_pre / bint_pre  unaries [m][n][h] and tables built by the caller with a BLAS sgemm, as the
 *                               reference does (src/utils.jl:135-136,164) -- the fixed-order fmaf chains above
 *                               pin the arithmetic for parity but would handicap a timed baseline;
 *   phases[5]                   seconds spent in {tables, unaries, ICM steps incl. the per-step unary copy,
 *                               veccost, perturb + accept + snapshots}. */
int orc_encode_icm_fully_ex(const float* X, const float* C, uint8_t* B, int64_t n, int d, int m, int h,
                            int ilsiter, int icmiter, int npert, int randord, uint64_t seed, int64_t g0,
                            const int* orders, condition_fn step,
                            const int* snap_iters, int n_snap, uint8_t* B_snap, float* objs,
                            float* cost_out, int* stats,
                            const float* U_pre, const float* bin_pre, const float* bint_pre, double* phases) {
  if (h != H256 && step) return -1; /* the C++ step is 256-only, src/LSQ.jl:173-175; cpp=false takes any h */
  if (h < 1 || h > 256) return -1;  /* codes are bytes below the boundary */
  if (!step && h == H256) step = orc_condition;
  int ncbi = m * (m - 1) / 2;
  size_t hh = (size_t)h * h;
  float* U = U_pre ? (float*)U_pre : (float*)malloc(sizeof(float) * (size_t)m * n * h);
  float* bin = bin_pre ? (float*)bin_pre : (float*)malloc(sizeof(float) * hh * (ncbi > 0 ? ncbi : 1));
  float* bin_t = bint_pre ? (float*)bint_pre : (float*)malloc(sizeof(float) * hh * (ncbi > 0 ? ncbi : 1));
  double ph[5] = {0, 0, 0, 0, 0}, t0;
  int* cbi = (int*)malloc(sizeof(int) * 2 * (ncbi > 0 ? ncbi : 1));
  float* ub = (float*)malloc(sizeof(float) * (size_t)n * h);
  uint8_t* newB = (uint8_t*)malloc((size_t)n * m);
  float* prevcost = (float*)malloc(sizeof(float) * n);
  float* newcost = (float*)malloc(sizeof(float) * n);
  int* pair2idx = (int*)calloc((size_t)m * m, sizeof(int));
  int* to_cond = (int*)malloc(sizeof(int) * (m > 1 ? m - 1 : 1));
  int* order = (int*)malloc(sizeof(int) * m);
  if (!U || !bin || !bin_t || !ub || !newB || !prevcost || !newcost) return -2;

  t0 = omp_get_wtime();
  if (bin_pre && bint_pre) {
    int idx = 0;
    for (int i = 0; i < m; i++) for (int j = i + 1; j < m; j++, idx++) { cbi[2 * idx] = i; cbi[2 * idx + 1] = j; }
  } else {
    orc_get_binaries(C, d, m, h, bin, bin_t, cbi);           /* src/LSQ.jl:288 */
  }
  ph[0] = omp_get_wtime() - t0; t0 = omp_get_wtime();
  if (!U_pre) orc_get_unaries(X, C, n, d, m, h, U);          /* src/LSQ.jl:168 */
  ph[1] = omp_get_wtime() - t0;
  for (int i = 0; i < ncbi; i++) {                           /* src/LSQ.jl:186-190 */
    pair2idx[cbi[2 * i] * m + cbi[2 * i + 1]] = i;
    pair2idx[cbi[2 * i + 1] * m + cbi[2 * i]] = i;
  }

  for (int it = 0; it < ilsiter; it++) {
    t0 = omp_get_wtime();
    orc_veccost(X, B, C, n, d, m, h, prevcost);              /* src/LSQ.jl:201 */
    ph[3] += omp_get_wtime() - t0; t0 = omp_get_wtime();
    memcpy(newB, B, (size_t)n * m);                          /* src/LSQ.jl:207 */
    if (orders) memcpy(order, orders + (size_t)it * m, sizeof(int) * m);
    else if (randord) orc_randperm(seed, it, m, order);      /* src/LSQ.jl:218-221 */
    else for (int i = 0; i < m; i++) order[i] = i;
    orc_perturb_codes(newB, n, m, h, npert, seed, it, g0);   /* src/LSQ.jl:225 */
    ph[4] += omp_get_wtime() - t0; t0 = omp_get_wtime();
    for (int i = 0; i < icmiter; i++)                        /* src/LSQ.jl:64-78 */
      for (int s = 0; s < m; s++) {
        int j = order[s];
        int q = 0;
        for (int k = 0; k < m; k++) if (k != j) to_cond[q++] = k;  /* ascending, src/LSQ.jl:211-216 */
        memcpy(ub, U + (size_t)j * n * h, sizeof(float) * (size_t)n * h);  /* src/LSQ.jl:69 */
        if (step) step(newB, ub, bin, bin_t, pair2idx, to_cond, j, (int)n, m);
        else orc_condition_h(newB, ub, bin, bin_t, pair2idx, to_cond, j, (int)n, m, h);   /* src/LSQ.jl:83-149 */
      }
    ph[2] += omp_get_wtime() - t0; t0 = omp_get_wtime();
    orc_veccost(X, newB, C, n, d, m, h, newcost);            /* src/LSQ.jl:237 */
    ph[3] += omp_get_wtime() - t0; t0 = omp_get_wtime();
    int neq = 0, nbet = 0;
    for (int64_t l = 0; l < n; l++) {
      if (newcost[l] == prevcost[l]) neq++;
      if (newcost[l] < prevcost[l]) { nbet++; memcpy(B + l * m, newB + l * m, m); }  /* strict <, src/LSQ.jl:242-247 */
    }
    if (stats) { stats[2 * it] = neq; stats[2 * it + 1] = nbet; }
    for (int s = 0; s < n_snap; s++)
      if (snap_iters[s] == it + 1) {
        if (B_snap) memcpy(B_snap + (size_t)s * n * m, B, (size_t)n * m);
        if (objs) objs[s] = (float)orc_qerror(X, B, C, n, d, m, h);
      }
    ph[4] += omp_get_wtime() - t0;
  }
  if (cost_out) orc_veccost(X, B, C, n, d, m, h, cost_out);
  if (phases) memcpy(phases, ph, sizeof ph);
  if (!U_pre) free(U);
  if (!bin_pre) free(bin);
  if (!bint_pre) free(bin_t);
  free(cbi); free(ub); free(newB); free(prevcost); free(newcost);
  free(pair2idx); free(to_cond); free(order);
  return 0;
}

int orc_encode_icm_fully(const float* X, const float* C, uint8_t* B, int64_t n, int d, int m, int h,
                         int ilsiter, int icmiter, int npert, int randord, uint64_t seed, int64_t g0,
                         const int* orders, condition_fn step,
                         const int* snap_iters, int n_snap, uint8_t* B_snap, float* objs,
                         float* cost_out, int* stats) {
  return orc_encode_icm_fully_ex(X, C, B, n, d, m, h, ilsiter, icmiter, npert, randord, seed, g0, orders, step,
                                 snap_iters, n_snap, B_snap, objs, cost_out, stats, NULL, NULL, NULL, NULL);
}

/* OpenMP team size of this process (liboracle and oracle/_ref share one libgomp): bench.py's CPU arm forces it to
 * the host core count -- torchrun exports OMP_NUM_THREADS=1 -- and reports what it got. */
void orc_set_num_threads(int t) { if (t > 0) omp_set_num_threads(t); }
int orc_get_max_threads(void) { return omp_get_max_threads(); }

/* ------------------------------------------------------------------------------------------
 * Linear scan.  (dist, idx) pairs ordered lexicographically, as std::partial_sort over
 * std::pair<float,int> does (deps/src/linscan_aqd_pairwise_byte.cpp:82, linscan_aqd.cpp:91).
 * A bounded max-heap of the k best replaces the reference's 1e7-pair chunk buffer; the result is
 * identical because the order is total.
 * ---------------------------------------------------------------------------------------- */
typedef struct { float d; int64_t i; } pair_t;
static inline int pair_less(pair_t a, pair_t b) { return a.d < b.d || (a.d == b.d && a.i < b.i); }
static void heap_sift_down(pair_t* hp, int k, int p) {
  for (;;) {
    int l = 2 * p + 1, r = l + 1, g = p;
    if (l < k && pair_less(hp[g], hp[l])) g = l;
    if (r < k && pair_less(hp[g], hp[r])) g = r;
    if (g == p) return;
    pair_t t = hp[p]; hp[p] = hp[g]; hp[g] = t; p = g;
  }
}
static int pair_cmp(const void* a, const void* b) {
  pair_t x = *(const pair_t*)a, y = *(const pair_t*)b;
  return pair_less(x, y) ? -1 : (pair_less(y, x) ? 1 : 0);
}
static inline void topk_push(pair_t* hp, int* cnt, int k, pair_t p) {
  if (*cnt < k) {
    hp[(*cnt)++] = p;
    if (*cnt == k) for (int s = k / 2 - 1; s >= 0; s--) heap_sift_down(hp, k, s);
  } else if (pair_less(p, hp[0])) { hp[0] = p; heap_sift_down(hp, k, 0); }
}

/* kind: 0 = linscan_aqd_query_extra_byte   (LSQ:  lut = -2<q,c>, + dbnorms, ids 1-based)  pairwise_byte.cpp:14-94
 *       1 = linscan_aqd_cq_query_extra_byte (CQ:   lut = ||q-c||^2, no norms, ids 1-based) pairwise_byte.cpp:97-176
 *       2 = linscan_aqd_query               (PQ:   per-subspace ||c-q||^2, ids 0-based)    linscan_aqd.cpp:37-102
 * For kind 2 `codebooks` is the sub-by-h-by-m array (centers[t*subdim+s]) and d = m*subdim.
 * id_offset is added to every id (global ids for a base shard). */
int orc_linscan(int kind, float* dists, int32_t* idx, const uint8_t* codes, const float* queries,
                const float* codebooks, const float* dbnorms, int nq, int64_t n, int m, int h, int d,
                int nn, int64_t id_offset) {
  if (nn > n) return -1;
  int mh = m * h;
  int subdim = (kind == 2) ? d / m : d;
#pragma omp parallel
  {
    float* lut = (float*)malloc(sizeof(float) * mh);
    pair_t* hp = (pair_t*)malloc(sizeof(pair_t) * nn);
#pragma omp for schedule(dynamic, 1)
    for (int qi = 0; qi < nq; qi++) {
      const float* q = queries + (size_t)qi * d;
      for (int j = 0; j < mh; j++) {
        float t = 0.0f;
        if (kind == 0) {
          const float* c = codebooks + (size_t)j * d;
          for (int k = 0; k < d; k++) t -= 2 * q[k] * c[k];               /* pairwise_byte.cpp:45-47 */
        } else if (kind == 1) {
          const float* c = codebooks + (size_t)j * d;
          for (int k = 0; k < d; k++) t += (q[k] - c[k]) * (q[k] - c[k]); /* pairwise_byte.cpp:127-130 */
        } else {
          int kk = j / h;
          const float* c = codebooks + (size_t)j * subdim;
          for (int s = 0; s < subdim; s++) { float df = c[s] - q[kk * subdim + s]; t += df * df; } /* linscan_aqd.cpp:66-74 */
        }
        lut[j] = t;
      }
      int cnt = 0;
      const uint8_t* code = codes;
      for (int64_t i = 0; i < n; i++, code += m) {
        float s = 0.0f;
        for (int k = 0; k < m; k++) s += lut[h * k + code[k]];             /* pairwise_byte.cpp:70-73 */
        if (kind == 0) s += dbnorms[i];                                    /* pairwise_byte.cpp:74 */
        pair_t p = {s, i + (kind == 2 ? 0 : 1) + id_offset};               /* pairwise_byte.cpp:76 / linscan_aqd.cpp:88 */
        topk_push(hp, &cnt, nn, p);
      }
      qsort(hp, nn, sizeof(pair_t), pair_cmp);
      for (int j = 0; j < nn; j++) { dists[(size_t)qi * nn + j] = hp[j].d; idx[(size_t)qi * nn + j] = (int32_t)hp[j].i; }
    }
    free(lut); free(hp);
  }
  return 0;
}

/* ------------------------------------------------------------------------------------------
 * quantize_pq (src/PQ.jl:18-48).  PARITY UNPINNED: arithmetic is in Distances.jl 0.8.0
 * (Manifest.toml:135-139) pairwise(SqEuclidean): max(sa2[i] + sb2[j] - 2*r[i,j], 0) with r from
 * BLAS gemm, and Clustering.jl 0.12.2 (Manifest.toml:79-83) update_assignments!: first minimum,
 * strict <.  Fixed here to sequential fmaf dots / norms.
 * C is the sub-by-h-by-m array (Cpq[(k*h + c)*sub + s]); codes out m-by-n uint8 0-based.
 * ---------------------------------------------------------------------------------------- */
void orc_quantize_pq(const float* X, const float* Cpq, int64_t n, int d, int m, int h, uint8_t* B) {
  int sub = d / m;
  float* cn = (float*)malloc(sizeof(float) * m * h);
  for (int e = 0; e < m * h; e++) cn[e] = dot_seq(Cpq + (size_t)e * sub, Cpq + (size_t)e * sub, sub);
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < n; l++)
    for (int k = 0; k < m; k++) {
      const float* x = X + l * d + k * sub;
      float xn = dot_seq(x, x, sub);
      float best = 0; int bi = 0;
      for (int c = 0; c < h; c++) {
        float r = dot_seq(Cpq + ((size_t)k * h + c) * sub, x, sub);
        float v = (cn[k * h + c] + xn) - 2 * r;
        v = v > 0.0f ? v : 0.0f;
        if (c == 0 || v < best) { best = v; bi = c; }
      }
      B[l * m + k] = (uint8_t)bi;
    }
  free(cn);
}

/* ------------------------------------------------------------------------------------------
 * fast_bin_matmul (src/codebook_update.jl:96-171): A = B'B + rho*I, b = B'X' exploiting that each
 * column of the indicator matrix has exactly one 1 per codebook.
 *   A  (m*h)-by-(m*h) double: A[(i,ci),(j,cj)] = #{l : B[i,l]=ci and B[j,l]=cj}  (+rho on the diagonal);
 *      the reference counts in Float32 (exact below 2^24) and widens when adding rho*I (:166).
 *   b  (m*h)-by-d double, column-major (b[t*mh + i*h + c]) = sum over l with B[i,l]=c of X[t,l], accumulated
 *      in Float64 in ascending l (:151-161; the @simd runs over t, so the order over l is as written).
 * PINNED by construction: every quantity is either an exact integer or a sequential double sum.
 * ---------------------------------------------------------------------------------------- */
void orc_fast_bin_matmul(const float* X, const uint8_t* B, int64_t n, int d, int m, int h, double rho,
                         double* A, double* b) {
  size_t mh = (size_t)m * h;
  memset(A, 0, sizeof(double) * mh * mh);
  memset(b, 0, sizeof(double) * mh * d);
  for (int64_t l = 0; l < n; l++)
    for (int i = 0; i < m; i++) {
      size_t ri = (size_t)i * h + B[l * m + i];
      for (int j = 0; j < m; j++) {
        size_t rj = (size_t)j * h + B[l * m + j];
        A[rj * mh + ri] += 1.0;
      }
      for (int t = 0; t < d; t++) b[(size_t)t * mh + ri] += (double)X[l * d + t];
    }
  for (size_t r = 0; r < mh; r++) A[r * mh + r] += rho;
}

/* quantize_norms (src/utils.jl:29-59): reconstruct (src/qerrors.jl:6-33) accumulates the codebook entries from
 * +0 in codebook order; the norm is the sum of squares (the reference's @simd leaves the order open; fixed to
 * sequential, unfused); code = first minimum of (norm - cbnorms[c])^2 (findmin). cbnorms/codes may be NULL. */
void orc_quantize_norms(const uint8_t* B, const float* C, const float* cbnorms, int64_t n, int d, int m, int h,
                        uint8_t* codes, float* norms) {
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < n; l++) {
    float nrm = 0.0f;
    for (int t = 0; t < d; t++) {
      float cb = 0.0f;
      for (int k = 0; k < m; k++) cb += C[((size_t)k * h + B[l * m + k]) * d + t];
      float sq = cb * cb;
      nrm += sq;
    }
    if (norms) norms[l] = nrm;
    if (cbnorms && codes) {
      float best = 0; int bi = 0;
      for (int c = 0; c < h; c++) {
        float df = nrm - cbnorms[c];
        float v = df * df;
        if (c == 0 || v < best) { best = v; bi = c; }
      }
      codes[l] = (uint8_t)bi;
    }
  }
}

/* viterbi_encoding (deps/src/encode_icm.cpp:63-152; Julia twin src/ChainQ.jl:36-128): exact min-sum Viterbi on
 * the chain.  unaries n-by-(m*H) (vector-major), binaries (m-1) tables bb[j*H + k] = cost of k -> j.
 * Same signature as the reference symbol.  PINNED against oracle/_ref/encode_icm.so (tests/test_oracle.py). */
void orc_viterbi_encoding(unsigned char* B, float* unaries, float* binaries, int n, int m) {
  const int H = H256;
#pragma omp parallel
  {
    float* U = (float*)malloc(sizeof(float) * m * H);
    float* mincost = (float*)calloc(H, sizeof(float));
    int* minidx = (int*)calloc((size_t)m * H, sizeof(int));
#pragma omp for schedule(static)
    for (int idx = 0; idx < n; idx++) {
      for (int i = 0; i < m * H; i++) U[i] = unaries[(size_t)idx * H * m + i];
      for (int i = 0; i < m - 1; i++) {
        if (i > 0) for (int j = 0; j < H; j++) U[i * H + j] += mincost[j];
        const float* bb = binaries + (size_t)H * H * i;
        float newcost[H256];
        for (int j = 0; j < H; j++) {
          float minv = U[i * H] + bb[j * H];
          int mini = 0;
          for (int k = 1; k < H; k++) {
            float c = U[i * H + k] + bb[j * H + k];
            if (c < minv) { minv = c; mini = k; }
          }
          newcost[j] = minv;
          minidx[i * H + j] = mini;
        }
        memcpy(mincost, newcost, sizeof(newcost));
      }
      if (m > 1) for (int j = 0; j < H; j++) U[(m - 1) * H + j] += mincost[j];
      float minv = U[(m - 1) * H];
      int state = 0;
      for (int j = 1; j < H; j++) if (U[(m - 1) * H + j] < minv) { minv = U[(m - 1) * H + j]; state = j; }
      B[(size_t)idx * m + (m - 1)] = (unsigned char)state;
      for (int i = m - 2; i >= 0; i--) { state = minidx[i * H + state]; B[(size_t)idx * m + i] = (unsigned char)state; }
    }
    free(U); free(mincost); free(minidx);
  }
}

/* quantize_chainq (src/ChainQ.jl:287-348): unaries (get_unaries), binaries[i] = 2*C[i]'*C[i+1] (:303-306), Viterbi.
 * vit: orc_viterbi_encoding or the reference's own symbol. */
typedef void (*viterbi_fn)(unsigned char*, float*, float*, int, int);
void orc_quantize_chainq(const float* X, const float* C, int64_t n, int d, int m, int h, uint8_t* B, viterbi_fn vit) {
  size_t hh = (size_t)h * h;
  float* U = (float*)malloc(sizeof(float) * (size_t)m * n * h);       /* [j][l][c] as the reference holds them */
  float* U2 = (float*)malloc(sizeof(float) * (size_t)m * n * h);      /* vcat(unaries...): [l][j][c] */
  float* bin = (float*)malloc(sizeof(float) * hh * (m > 1 ? m - 1 : 1));
  orc_get_unaries(X, C, n, d, m, h, U);
  for (int64_t l = 0; l < n; l++)
    for (int j = 0; j < m; j++)
      memcpy(U2 + ((size_t)l * m + j) * h, U + ((size_t)j * n + l) * h, sizeof(float) * h);
  for (int i = 0; i + 1 < m; i++)
#pragma omp parallel for schedule(static)
    for (int b = 0; b < h; b++)
      for (int a = 0; a < h; a++)
        bin[(size_t)i * hh + (size_t)b * h + a] = 2.0f * dot_seq(C + ((size_t)i * h + a) * d, C + ((size_t)(i + 1) * h + b) * d, d);
  (vit ? vit : orc_viterbi_encoding)(B, U2, bin, (int)n, m);
  free(U); free(U2); free(bin);
}
