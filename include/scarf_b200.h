/* scarf_b200 -- C-ABI of the B200-native make_graph hot path.
 *
 * The reference (parashardhapola/scarf 0.32.3) has no FFI: its seam is Python
 * (GraphDataStore.make_graph, scarf/datastore/graph_datastore.py:513).  Each entry point below
 * replaces the arithmetic of one reference call site (cited per function, paths relative to the
 * reference checkout).  INTEGRATION.md shows the ctypes stub a Scarf maintainer would add.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller unless marked "host";
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), no hidden
 *     allocation, no global mutable state; scratch is passed in (see *_workspace_bytes);
 *   - return value: 0 = ok, >0 = argument error, <0 = -(cudaError_t); text via scf_last_error()
 *     (thread local);
 *   - "fx" arrays are int64 fixed-point accumulators (value * 2^shift): integer adds are
 *     associative, so results are bit-identical for any launch geometry and any number of GPUs
 *     (an int64 sum all-reduce between ranks keeps that property).
 */
#ifndef SCARF_B200_H
#define SCARF_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCF_VERSION 100
#define SCF_COLSTAT_SHIFT 34 /* sum x, sum x^2 : x <= log1p(sf) */
#define SCF_GRAM_SHIFT 36    /* |sum z_j z_k| <= n_rows (unit-variance columns) */

int32_t scf_version(void);
const char* scf_last_error(void);

/* ---- K0a: per-row sum and count of stored values over a column subset -----------------------
 * Replaces  nCounts / nFeatures  (scarf/datastore/base_datastore.py:345-366) when col_map==NULL
 * and the renormalisation scalar  counts.sum(axis=1)  of RNAassay.normed (scarf/assay.py:814-823)
 * when col_map selects the HVG columns (col_map[g] < 0 = not selected).
 * row_ids (nullable): CSR rows to process, output row r <-> CSR row row_ids[r]. */
int32_t scf_csr_row_sums(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                         const int64_t* row_ids, int64_t n_sel, const int32_t* col_map,
                         double* out_sum, int32_t* out_nnz, void* stream);

/* ---- K0b: per-gene nnz / sum / sum of squares of v = sf*c/row_div[r] -------------------------
 * Replaces the three dask passes of RNAassay.set_feature_stats (scarf/assay.py:848-858) over
 * norm_lib_size (scarf/assay.py:41-51): row_div[r] = nCounts of output row r.  row_div==NULL ->
 * v = c (gives per-feature nCells, scarf/assay.py:201-225).  Outputs are ACCUMULATED (caller
 * zeroes); gene_sum / gene_sumsq may be NULL. */
int32_t scf_csr_gene_stats(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                           const int64_t* row_ids, int64_t n_sel, int32_t n_genes,
                           const double* row_div, double sf, unsigned long long* gene_nnz,
                           double* gene_sum, double* gene_sumsq, void* stream);

/* Same statistics without a reduction per stored value: a CTA owns (block of rows) x (window of 1024 genes), adds
 * the rows one after the other into shared-memory accumulators with plain read-modify-writes (a gene occurs once per
 * row) and flushes one reduction per gene.  Needs ascending column ids inside every row (the CSR contract of this
 * library).  Results are ACCUMULATED into gene_nnz / gene_sum / gene_sumsq exactly like scf_csr_gene_stats
 * (gene_sum == NULL: nnz only).  workspace (device): scf_csr_gene_stats_workspace_bytes(n_sel, n_genes). */
int64_t scf_csr_gene_stats_workspace_bytes(int64_t n_sel, int32_t n_genes);
int32_t scf_csr_gene_stats_windowed(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                                    const int64_t* row_ids, int64_t n_sel, int32_t n_genes,
                                    const double* row_div, double sf, unsigned long long* gene_nnz,
                                    double* gene_sum, double* gene_sumsq, void* workspace,
                                    int64_t workspace_bytes, void* stream);

/* ---- K1a: column sums of the normalised HVG matrix -------------------------------------------
 * x = log1p(sf*c/row_sum[r]) (log_transform) or sf*c/row_sum[r]   (scarf/assay.py:54-64,826)
 * accumulates sum x and sum x^2 per selected column as int64 fixed point (<< SCF_COLSTAT_SHIFT):
 * the mu / sigma block of make_graph (scarf/datastore/graph_datastore.py:767-796).
 * sum_fx / sumsq_fx are [n_rep][n_cols]: CTAs spread their atomics over n_rep copies (same-address
 * atomics serialise in L2); the caller adds the copies (integer adds: any order, same result). */
int32_t scf_csr_hvg_colstats(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                             const int64_t* row_ids, int64_t n_sel, const int32_t* col_map,
                             const double* row_sum, double sf, int32_t log_transform, int32_t n_cols,
                             int32_t n_rep, int64_t* sum_fx, int64_t* sumsq_fx, void* stream);

/* ---- K1a' / K1b': the same two steps through a compact matrix ---------------------------------
 * scf_csr_hvg_compact does K1a's pass and, in the same scan, writes the selected non-zero entries of
 * every row as a small CSR of normalised values (out_col int32, out_x float64; row r occupies
 * [row_off[r], row_off[r+1]), row_off = exclusive prefix sum of the out_nnz of scf_csr_row_sums).
 * scf_hvg_dense_scale then produces exactly the Z (and z_lo) of scf_csr_norm_scale from that matrix
 * without re-reading the raw counts or re-evaluating log1p: HBM sees 12 B per selected entry
 * instead of 8 B per stored value.  make_graph uses this pair; scf_csr_norm_scale stays for callers
 * that already know mu / sigma (run_mapping). */
int32_t scf_csr_hvg_compact(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                            const int64_t* row_ids, int64_t n_sel, const int32_t* col_map,
                            const double* row_sum, double sf, int32_t log_transform,
                            const int64_t* row_off, int32_t* out_col, double* out_x, int32_t n_cols,
                            int32_t n_rep, int64_t* sum_fx, int64_t* sumsq_fx, void* stream);
int32_t scf_hvg_dense_scale(const int64_t* row_off, const int32_t* cols, const double* xs,
                            int64_t n_sel, int32_t n_cols, const double* mu, const double* sigma,
                            float* z, float* z_lo, int64_t ldz, void* stream);

/* ---- K1b: fused lib-size normalise -> log1p -> HVG gather -> z-scale --------------------------
 * Z[r, col_map[g]] = (x - mu)/sigma   (AnnStream.transform_z, scarf/ann.py:191-192; the dense
 * row also gets (0 - mu)/sigma where the cell has no count).  Z is float32 row-major with row
 * stride ldz >= n_cols; columns [n_cols, ldz) are written as 0.  mu==NULL -> 0, sigma==NULL -> 1
 * (plain normalised values, used by the parity tests of the normalisation itself).
 * missing_fill (nullable, float64[n_cols]): value of x for columns that do not exist in this
 * matrix at all (align_features fills them with 1.0, scarf/mapping_utils.py:208-211); a column
 * with missing_fill[j] == missing_fill[j] (not NaN) ignores the CSR and uses that x.
 * z_lo (nullable, same shape / stride as z): Z - tf32_trunc(Z), the low plane of the 3xTF32 Gram. */
int32_t scf_csr_norm_scale(const int64_t* indptr, const int32_t* indices, const uint32_t* data,
                           const int64_t* row_ids, int64_t n_sel, const int32_t* col_map,
                           int32_t n_cols, const double* row_sum, double sf, int32_t log_transform,
                           const double* mu, const double* sigma, const double* missing_fill,
                           float* z, float* z_lo, int64_t ldz, void* stream);

/* ---- K2: Gram accumulation  G += Z^T Z  (upper triangle m <= n of the n_cols x n_cols matrix) ---
 * Replaces the per-block SVD updates of sklearn IncrementalPCA.partial_fit driven by
 * AnnStream._fit_pca (scarf/ann.py:207-256) by the exact covariance route.
 * Rows are consumed in fixed slabs of SCF_GRAM_SLAB rows (one reference block, scarf/ann.py:187-189):
 * float32 partial sums inside a slab, int64 fixed point << SCF_GRAM_SHIFT across slabs, so G is
 * bit-identical however the slabs are spread over CTAs or GPUs (shards aligned to the slab size).
 * Rows past n_rows are not read.  ldg = row stride of g_fx.  After the last accumulate call (and
 * after the all-reduce between ranks) scf_gram_symmetrize copies the upper triangle into the lower.
 * mode: 0 = FP32 SIMT, 1 = TF32 tcgen05, 3 = 3xTF32 tcgen05 (needs z_lo = Z - tf32_trunc(Z), the
 * second plane scf_csr_norm_scale writes; same shape / stride as z; NULL otherwise).
 * The tcgen05 modes need ldz % 32 == 0, an even ldg and 16-byte aligned z / z_lo / g_fx. */
#define SCF_GRAM_SLAB 1000
int32_t scf_gram_accumulate(const float* z, const float* z_lo, int64_t ldz, int64_t n_rows,
                            int32_t n_cols, int64_t* g_fx, int64_t ldg, int32_t mode, void* stream);
int32_t scf_gram_symmetrize(int64_t* g_fx, int32_t n_cols, int64_t ldg, void* stream);

/* ---- K4: projection  Y[:, :dims] = Z @ V ;  Y[:, dims:ldy] = 0 ----------------------------------
 * AnnStream.reducer (scarf/ann.py:138): V = loadings (n_cols x dims, float32 row-major, stride ldv) */
int32_t scf_project(const float* z, int64_t ldz, int64_t n_rows, int32_t n_cols, const float* v,
                    int64_t ldv, int32_t dims, float* y, int64_t ldy, void* stream);

/* Same product on the tensor cores (3xTF32, FP32 accumulate): needs the second plane z_lo = Z - tf32_trunc(Z) that
 * scf_csr_norm_scale / scf_hvg_dense_scale write for the Gram kernel (same shape / stride as z), ldz % 32 == 0,
 * dims <= 128, 16-byte aligned z / z_lo / y / workspace.  Relative error of an entry <= ~2e-5 (truncating tensor-core
 * accumulation over four interleaved chains).  workspace (device): scf_project_tc_workspace_bytes(ldz, dims). */
int64_t scf_project_tc_workspace_bytes(int64_t ldz, int32_t dims);
int32_t scf_project_tc(const float* z, const float* z_lo, int64_t ldz, int64_t n_rows, int32_t n_cols,
                       const float* v, int64_t ldv, int32_t dims, float* y, int64_t ldy, void* workspace,
                       int64_t workspace_bytes, void* stream);

/* ---- K3 helper: all eigenpairs of a small symmetric positive semi-definite matrix -------------------
 * The Rayleigh-Ritz matrices of the PCA eigensolve (scarf_b200/graph.py: eig_topk, which replaces the SVD updates of
 * IncrementalPCA.partial_fit, scarf/ann.py:207-256).  One CTA, one-sided Jacobi with the matrix in shared memory,
 * n <= scf_sym_eig_max_n() (168; worthwhile for n <~ 100).  a: n x n float64, row stride lda (symmetrised on load);
 * evals[n] ascending; evecs n x n row-major (row stride ldv), column k = eigenvector of evals[k]; info (device int32,
 * nullable): sweeps used, or -1 if the iteration did not converge. */
int32_t scf_sym_eig_max_n(void);
int32_t scf_sym_eig_jacobi(const double* a, int32_t n, int64_t lda, double* evals, double* evecs, int64_t ldv,
                           int32_t* info, void* stream);

/* The fast path of the same step: Householder tridiagonalisation, multisection on the Sturm count, inverse iteration
 * in a warp per eigenpair, back-transformation (one warp per eigenvector) and an orthogonality check -- four launches, the
 * tridiagonalisation in one CTA and the rest over the whole GPU, n in [3, 160].  Same outputs as scf_sym_eig_jacobi; work: 2 * n * n + 3 * n doubles
 * of device scratch; ok (device int32): 1 when the eigenvectors it produced are orthonormal to 1e-9, 0 when the
 * eigenvalues are too tightly clustered for inverse iteration without re-orthogonalisation (then call
 * scf_sym_eig_jacobi). */
int32_t scf_sym_eig_tridiag(const double* a, int32_t n, int64_t lda, double* evals, double* evecs, int64_t ldv,
                            double* work, int32_t* ok, void* stream);

/* ---- K3: top eigenpairs of the PCA covariance -------------------------------------------------------
 * Replaces the fit of sklearn's IncrementalPCA (AnnStream._fit_pca, scarf/ann.py:207-256) on the exact-covariance route:
 *   cov = gram_fx * scale - mean_weight * col_mean col_mean^T        (col_mean nullable)
 * with gram_fx the mirrored int64 fixed-point Gram of scf_gram_accumulate / scf_gram_symmetrize (row stride ldg), e.g.
 * scale = 2^-36 / (n - 1), mean_weight = n / (n - 1) for a PCA fitted on n rows whose column mean is col_mean.
 * Chebyshev-filtered subspace iteration on dims + 32 columns in FP64, every kernel in this library (no cuSOLVER /
 * cuBLAS): evals[dims] descending, evecs [h, dims] row major with sklearn's sign rule (svd_flip, v-based: the entry of
 * largest magnitude of every component is positive), residual |C v - lambda v| <= tol * lambda_1 for every pair;
 * evecs_f32 (nullable): the same loadings rounded to float32, [h, ld32] row major with zeroed pad columns -- the
 * operand of scf_project / scf_project_tc.
 * Deterministic (ranks that hold the same Gram get bit-identical loadings).  dims + 8 <= 160 columns.
 * The call synchronises the stream once per round; host_report (host memory, >= 176 doubles) receives
 * [0] residual, [1] trace, [2] norm bound, [3] Cholesky break-down flag, [4] smallest Cholesky pivot, [5] rounds
 * (negative: not converged), [6] restarts, [7] 1 if the eigenvalue-based orthonormalisation was used,
 * [8 ...) the Ritz values of the last round, [168] kernels launched.  Returns 2 for a zero / non-finite covariance, 3 without convergence
 * within max_rounds.  workspace (device): scf_eig_topk_workspace_bytes(h, dims) (-1: unsupported sizes). */
int64_t scf_eig_topk_workspace_bytes(int32_t h, int32_t dims);
int32_t scf_eig_topk(const int64_t* gram_fx, int64_t ldg, int32_t h, double scale, const double* col_mean,
                     double mean_weight, int32_t dims, double tol, int32_t max_rounds, double* evals, double* evecs,
                     float* evecs_f32, int64_t ld32, double* host_report, void* workspace, int64_t workspace_bytes,
                     void* stream);

/* ---- K5: exact k nearest neighbours, squared L2 --------------------------------------------------
 * Replaces hnswlib Index(space='l2').knn_query + fix_knn_query (scarf/ann.py:14-52,194-205):
 *   d(a,b) = (float) sum_t ((double)a_t - (double)b_t)^2   (t ascending), order by (d, index),
 *   query i is reference i + self_offset and is excluded when self_offset >= 0 (-1: keep all).
 * q: nq x ld float32, ref: nref x ld float32 (row stride ld >= dim).  out_idx int64 [nq,k],
 * out_dist float32 [nq,k].  method: 0 = FP64 SIMT brute force, 1 = tcgen05 FP16 candidates (operands
 * scaled by an exact power of two and rounded to FP16, FP32 accumulate) + exact FP64 re-rank with a
 * proven guard band; rows that fail the guard are repaired inside the same call (tensor-core threshold
 * collect, FP64 scan for the rest).  k <= 24 and dim <= 189 run on the tensor cores, other shapes take
 * method 0.  workspace: scf_knn_workspace_bytes(nq, nref, dim, k, method). */
int64_t scf_knn_workspace_bytes(int64_t nq, int64_t nref, int32_t dim, int32_t k, int32_t method);
/* diagnostics: byte offset inside the workspace of an int32 that, after scf_knn_l2(method 1), holds the
 * number of query rows whose guard band could not be proven (recomputed by method 0); -1 if n/a. */
int64_t scf_knn_fail_count_offset(int64_t nq, int64_t nref, int32_t dim, int32_t k, int32_t method);
/* diagnostics (host only, no GPU work): how method 1 would run a shape.  out16 (host memory, 16 x int32): [0] k' per
 * candidate list, [1] 64-wide K chunks, [2] 1 = CTA-pair kernel, [3] units (CTAs, or CTA pairs) launched, [4] rounds of
 * whole query tiles per unit, [5] units that share the remaining query tiles, [6] lists a cut query tile keeps per column
 * group, [7] candidate lists per query row, [8] column groups per row, [9] reference tiles, [10] query tiles, [11] TMA
 * stages, [12] MMA-issuing warps, [13] dynamic shared memory in KiB, [14] candidates per row the re-rank can take,
 * [15] 0.  Returns 1 when the shape takes method 0 (k > 24 or dim > 189). */
int32_t scf_knn_plan(int64_t nq, int64_t nref, int32_t dim, int32_t k, int32_t* out16);
/* measurement hook: the next scf_knn_l2(method 1) issued by the calling thread records the two cudaEvent_t (created
 * by the caller with timing enabled) on its stream around the tcgen05 distance / top-k' kernel alone, so that the
 * dominant kernel can be timed live without a profiler (bench.py `roofline`).  One shot; NULL, NULL cancels. */
int32_t scf_knn_time_next_call(void* event_start, void* event_stop);
int32_t scf_knn_l2(const float* q, int64_t nq, const float* ref, int64_t nref, int32_t dim,
                   int64_t ld, int32_t k, int64_t self_offset, int64_t* out_idx, float* out_dist,
                   int32_t method, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- K6: smooth_knn_dist + compute_membership_strengths + COO assembly ---------------------------
 * umap-learn functions called per chunk by smoothen_dists (scarf/knn_utils.py:89-159).
 * dist float32 [n,k] ascending; idx int64 [n,k] GLOBAL ids.  Row r belongs to chunk
 * (row_offset + r) / chunk_size; its chunk-local id (used by the "neighbour == i" rule, SURVEY
 * fact 6) is (row_offset + r) % chunk_size.  chunk_mean float32 [n_chunks_total] = mean distance
 * of each chunk, indexed by the GLOBAL chunk id: scf_chunk_sums gives this shard's float64 sum
 * per chunk (written only for the chunks the shard touches; sum all-reduce completes a chunk that
 * straddles shards), the caller divides by rows_in_chunk * k and rounds to float32.
 * Outputs: sigma,rho float32 [n]; edges int64 [n*k,2] (row_offset + r, idx); weights float32 [n*k];
 * chunk_min float32 [n_chunks_total] = min non-zero weight per chunk (caller presets +inf),
 * chunk_has_zero int32 [n_chunks_total] = 1 where the chunk produced a zero weight (caller presets 0):
 * the floor of scarf/knn_utils.py:121,147-158 is min(1, min over chunks with a zero of chunk_min). */
int32_t scf_chunk_sums(const float* dist, int64_t n, int32_t k, int64_t row_offset,
                       int64_t chunk_size, double* chunk_sum, void* stream);
int32_t scf_smooth_knn(const float* dist, int64_t n, int32_t k, float local_connectivity,
                       float bandwidth, int64_t row_offset, int64_t chunk_size,
                       const float* chunk_mean, float* sigma, float* rho, void* stream);
int32_t scf_membership_coo(const int64_t* idx, const float* dist, const float* sigma,
                           const float* rho, int64_t n, int32_t k, int64_t row_offset,
                           int64_t chunk_size, int64_t* edges, float* weights, float* chunk_min,
                           int32_t* chunk_has_zero, void* stream);
/* load_graph's symmetrisation (scarf/datastore/graph_datastore.py:1052-1075: g + g.T - g.multiply(g.T), optionally
 * scipy.sparse.triu) of the stored graph: idx int64 [n, k] neighbour ids (the second column of `edges`), weights
 * float64 [n, k]; only the first use_k neighbours of a row take part (_store_to_sparse, :474-511).  Writes 2 n use_k
 * COO slots (out_row / out_col int64, out_val float64); slots with out_row < 0 are unused. */
int32_t scf_graph_symmetrize(const int64_t* idx, const double* weights, int64_t n, int32_t k, int32_t use_k,
                             int32_t upper_only, int64_t* out_row, int64_t* out_col, double* out_val, void* stream);
/* zero weights := floor (scarf/knn_utils.py:154-158); floor is a host value */
int32_t scf_fill_zero_weights(float* weights, int64_t n, float floor_value, void* stream);

/* ---- host helper (no GPU work): robust LOWESS --------------------------------------------------
 * statsmodels lowess(endog, exog, frac, it, delta=0, return_sorted=False) as called by fit_lowess
 * (scarf/feat_utils.py:22,38-40) on the <= n_bins (200) binned (log mean, log variance) points of
 * mark_hvgs' trend removal.  All pointers are HOST pointers; out[i] = fitted value at exog[i]. */
int32_t scf_host_lowess(const double* endog, const double* exog, int64_t n, double frac, int32_t it,
                        double* out);

/* ---- the same fit on the device (one CTA, no host round trip) -------------------------------------
 * endog / exog / out: DEVICE float64 [n], n <= 512; valid (nullable, uint8 [n]): points with valid[i] == 0 are
 * left out of the fit (mark_hvgs' empty bins) and get out[i] = NaN.  Same arithmetic as scf_host_lowess with the
 * window sums split over four lanes (agreement ~1e-13 relative).  If fewer than 2 points are usable or frac * n_usable is outside [2, n_usable] (the host routine's
 * argument error) every out[i] is NaN. */
int32_t scf_lowess(const double* endog, const double* exog, const uint8_t* valid, int32_t n, double frac,
                   int32_t it, double* out, void* stream);

/* ---- mark_hvgs after the statistics pass, fused into one CTA ----------------------------------------
 * Trend removal (MetaData.remove_trend -> fit_lowess, scarf/metadata.py:586-617, scarf/feat_utils.py:11-45) and the
 * HVG choice (RNAassay.mark_hvgs, scarf/assay.py:1014-1063 + MetaData.multi_sift, scarf/metadata.py:483-533) on the
 * per-gene vectors of scf_csr_gene_stats*: avg = sum / n_cells_total, var = sumsq / m_cells - (sum / m_cells)^2,
 * n_bins equal-width bins on log(avg) over the genes of feat_i with avg > 0 (np.histogram edges, last edge + 0.1), the
 * minimum-log(var) gene of every bin, LOWESS(frac, it = 100) through them, c_var = exp(log var - fit(bin)); eligible =
 * min_cells < nnz < max_cells & min_mean < sum/nnz < max_mean (strict; pass -inf / +inf for open bounds) & feat_i &
 * keep; selected = eligible & c_var > (top_n + 1)-th largest eligible c_var.  All pointers DEVICE; feat_i / keep / hv:
 * one byte per gene (keep nullable); col_map[g] = rank of gene g among the selected ones or -1; *n_sel = their number.
 * n_bins <= 512.  workspace: scf_hvg_select_workspace_bytes(n_genes). */
int64_t scf_hvg_select_workspace_bytes(int32_t n_genes);
int32_t scf_hvg_select(const unsigned long long* gene_nnz, const double* gene_sum, const double* gene_sumsq,
                       const uint8_t* feat_i, const uint8_t* keep, int32_t n_genes, double m_cells,
                       double n_cells_total, int32_t n_bins, double lowess_frac, int32_t top_n,
                       double min_cells, double max_cells, double min_mean, double max_mean, uint8_t* hv,
                       int32_t* col_map, int32_t* n_sel, void* workspace, int64_t workspace_bytes,
                       void* stream);

/* ---- reading stores the reference wrote: chunk codec (host) and dense -> CSR (device) -------------------
 * scf_host_blosc_*: decoder of the Blosc-1 frames numcodecs.Blosc(cname='lz4', shuffle=SHUFFLE | BITSHUFFLE) writes
 * for every Zarr chunk of a Scarf store (create_zarr_dataset, scarf/writers.py:58-89; read back through zarr by
 * Assay.rawData, scarf/assay.py:134).  HOST pointers; pure functions (safe from several threads).  _info reports the
 * decoded size / element size / flag byte of a frame; _decode needs dst_bytes == that size.  Stored (memcpy) frames
 * and LZ4 frames are decoded, any other inner codec is an argument error. */
int32_t scf_host_blosc_info(const void* frame, int64_t frame_bytes, int64_t* nbytes, int32_t* typesize,
                            int32_t* flags);
int32_t scf_host_blosc_decode(const void* frame, int64_t frame_bytes, void* dst, int64_t dst_bytes);

/* scf_dense_row_nnz / scf_dense_to_csr: a dense row-major uint32 block [n_rows, ld] of raw counts (the decoded
 * chunks of `<assay>/counts`, scarf/writers.py:164-204) to the CSR rows the path computes on -- the GPU form of
 * Assay.to_raw_sparse (scarf/assay.py:175-199).  Pass 1 writes the number of non-zero values of the first n_cols
 * columns of every row; the caller builds row_ptr (absolute int64 start of every row in indices / data); pass 2
 * writes column ids (ascending) and counts.  ld % 4 == 0, block 16-byte aligned; columns >= n_cols are ignored. */
int32_t scf_dense_row_nnz(const uint32_t* dense, int64_t n_rows, int32_t n_cols, int64_t ld, int64_t* row_nnz,
                          void* stream);
int32_t scf_dense_to_csr(const uint32_t* dense, int64_t n_rows, int32_t n_cols, int64_t ld, const int64_t* row_ptr,
                         int32_t* indices, uint32_t* data, void* stream);

#ifdef __cplusplus
}
#endif
#endif
