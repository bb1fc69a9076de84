/* lyssa_b200.h — C-ABI of the B200-native Batch-OMP / K-SVD / ODL engine.
 *
 * Drop-in boundary for ONE hot path of ektormak/Lyssandra (paths relative to the reference
 * tree): lyssa.sparse_coding.sparse_encoder(algorithm='bomp').encode() and the
 * lyssa.dict_learning approximate-K-SVD / online learners that consume its codes.  The
 * reference has no FFI of its own (pure Python over NumPy/OpenBLAS); each entry point below
 * names the reference call site it replaces.  INTEGRATION.md shows the ctypes stub a
 * maintainer would add at those call sites.
 *
 * Conventions
 *  - every function returns 0 on success, a negative LYS_E* code on failure;
 *    lys_last_error() returns a thread-local message for the last failure.
 *  - unless the name ends in _host, all data pointers are DEVICE pointers of the current
 *    device; `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Calls are asynchronous with respect to the host unless stated otherwise.
 *  - the library never allocates caller-visible memory: ask *_workspace_bytes() and pass a
 *    scratch buffer (256-byte aligned).
 *  - arithmetic is float32, indices int32.  "signal" = one datapoint (a column of the
 *    reference's X); "atom" = one dictionary column.
 *  - strides are in ELEMENTS.  The reference keeps signals in columns, X(n,N) C-order:
 *    x_feat_stride = N, x_sig_stride = 1.  Signal-major storage (N,n): x_feat_stride = 1,
 *    x_sig_stride = n.  D is the reference's (n,K) C-order array: D[f*ldd + atom].
 *  - sparse codes are (idx,val)[N][k]: the atoms of signal i in SELECTION order, unused
 *    slots padded with idx = -1, val = 0 (a signal may stop early, see lys_bomp_encode).
 */
#ifndef LYSSA_B200_H
#define LYSSA_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define LYS_API __attribute__((visibility("default")))
#else
#define LYS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LYS_OK            0
#define LYS_EINVAL       -1   /* bad argument (shape, stride, null pointer, k out of range) */
#define LYS_ECUDA        -2   /* a CUDA runtime call or kernel launch failed */
#define LYS_EWORKSPACE   -3   /* workspace too small */
#define LYS_EUNSUPPORTED -4   /* shape outside what the kernels are built for */

#define LYS_MAX_NONZERO   32  /* k  <= 32 */
#define LYS_OMP_MAX_NONZERO 64  /* atoms per signal of the `omp` coder */
#define LYS_MAX_ATOMS   4096  /* K  <= 4096 */
#define LYS_MAX_FEATURES 256  /* n  <= 256 */

/* ---- library ------------------------------------------------------------------------ */
LYS_API int         lys_version(void);                 /* 10000*major + 100*minor + patch */
LYS_API const char* lys_last_error(void);
LYS_API const char* lys_build_fingerprint(void);   /* hash of the sources the library was built from (loader's staleness check) */
/* SM count / compute capability of `device`; fails unless it is an sm_100 part. */
LYS_API int         lys_device_info(int device, int* sm_count, int* cc_major, int* cc_minor);

/* ---- measurement hooks (bench.py) ------------------------------------------------------
 * lys_bomp_launch_count: number of kernel launches ONE lys_bomp_encode call with these
 * shapes performs (bench.py's gpu_launches claim).
 * lys_profile_enable(1): every launch of the encode path's DOMINANT kernel is bracketed by
 * CUDA events on its own stream; lys_profile_fetch synchronises those events and returns
 * the summed device time (ms), the number of launches and the kernel's name. */
LYS_API int lys_bomp_launch_count(int n, int K, int64_t N, int k);
LYS_API int lys_profile_enable(int on);
LYS_API int lys_profile_fetch(double* kernel_ms, int64_t* launches, const char** kernel_name, int reset);

/* ---- K1: Gram = D^T D --------------------------------------------------------------
 * replaces `Gram = fast_dot(D.T, D)`, lyssa/sparse_coding.py:630.  G is (K,K) row-major. */
LYS_API int lys_gram(const float* D, int64_t ldd, int n, int K, float* G, void* stream);
/* the same with a scratch buffer, which lets it run on the tcgen05 correlation GEMM (n = 64 or 128, K % 256 == 0:
 * G = D^T D is the correlation of the atoms with the dictionary; 8.6e-8 |d||d| from float64) — what the learners call
 * once per minibatch / iteration */
LYS_API size_t lys_gram_workspace_bytes(int n, int K);
LYS_API int lys_gram_ws(const float* D, int64_t ldd, int n, int K, float* G, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K2+K3+K4: Batch-OMP encode -------------------------------------------------------
 * replaces `Alpha = fast_dot(D.T, X)` (sparse_coding.py:631) and
 * run_parallel(batch_omp, ...) (sparse_coding.py:718 -> :302-367), including the dense
 * zero-fill + scatter `Z[Dx, i] = z` (:308,:365; lyssa/utils/__init__.py:69,89).
 *
 * Semantics mirrored from batch_omp: first-maximum argmax of |alpha| (:322); stop when the
 * picked atom is already selected (:323-325) or the Cholesky pivot 1 - w.w drops below
 * machine epsilon of the compute type (:335,:345); the atom self-product is the literal 1
 * (:334,:337,:344) so D must have unit-norm columns; `tol` does not exist (the reference
 * ignores it, :302,:536).
 *
 * Outputs: idx,val (N,k) as described above; nsel (N) number of atoms selected (may be
 * NULL); Z (optional, may be NULL) the dense code matrix the reference returns, written in
 * full (zeros included): element (atom c, signal i) at Z[c*z_atom_stride + i*z_sig_stride].
 * z_atom_stride = 1, z_sig_stride >= K (signal-major, returned to Python as a transposed
 * view) is the fast layout. */
LYS_API size_t lys_bomp_workspace_bytes(int n, int K, int64_t N, int k);
LYS_API int lys_bomp_encode(const float* X, int64_t x_feat_stride, int64_t x_sig_stride,
                    const float* D, int64_t ldd, const float* G,
                    int n, int K, int64_t N, int k,
                    int32_t* idx, float* val, int32_t* nsel,
                    float* Z, int64_t z_atom_stride, int64_t z_sig_stride,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Same call with option flags.  On the fused tcgen05 path (n <= 64, K in {256,...,1024}, k <= 10) the correlations
 * D^T r of every greedy step are three fp16 split products (fp32-faithful values).  LYS_BOMP_SCREEN computes ONE fp16
 * product that only ranks: a winner is accepted when a measured error bound certifies it (96.6 % of the decisions at
 * cfg2) and is otherwise recomputed exactly in fp32 by the warp (bomp_fused.cu), so the selection is the exact
 * first-maximum argmax of :322 either way and both modes return the same codes.  A third of the tensor work — but
 * measured SLOWER (2.21 vs 1.41 ms per 1M patches): the kernel is bound by the per-signal
 * scan/update chain, not by the tensor pipe, and the exact recomputations lengthen that chain.  Kept as an option
 * for A/B measurements; other shapes ignore the flag. */
#define LYS_BOMP_SCREEN 2
LYS_API int lys_bomp_encode_ex(const float* X, int64_t x_feat_stride, int64_t x_sig_stride,
                       const float* D, int64_t ldd, const float* G,
                       int n, int K, int64_t N, int k,
                       int32_t* idx, float* val, int32_t* nsel,
                       float* Z, int64_t z_atom_stride, int64_t z_sig_stride,
                       void* workspace, size_t workspace_bytes, int flags, void* stream);

/* ---- the reference's plain OMP coder on the same Gram / correlation front end ------------------------------
 * replaces algorithm 'omp' (the reference's DEFAULT algorithm; lyssa/sparse_coding.py:618-625 dispatch, :19-66 omp/_omp):
 * per signal, while the continue criterion holds: first-maximum argmax of |D^T r| (:40), stop if already selected
 * (:41-42), z = inv(G[I,I]) (D^T x)[I] with the TRUE Gram (no unit-norm assumption, unlike batch_omp; :45-53),
 * r = x - D_I z (:54).  Criterion (:27-34): `strict` != 0 — the n_nonzero_coefs form — continue while fewer than
 * k_max atoms are selected and ||r|| > tol (the reference forces tol = 1e-10 there); `strict` == 0 — the
 * tolerance-only form — continue while ||r|| >= tol, up to k_max <= LYS_OMP_MAX_NONZERO atoms; *truncated (device
 * int32, may be NULL, must be zeroed by the caller) counts the signals that still satisfied the criterion at k_max
 * atoms (the reference would have gone on).  ||r|| comes from ||x||^2 - y.y in float32, i.e. it is resolved to about
 * 3e-4 ||x||; a singular G[I,I] (the reference's LinAlgError, :48-51) stops the signal with the previous z.
 * idx/val are (N, k_max); Z as in lys_bomp_encode. */
LYS_API size_t lys_omp_workspace_bytes(int n, int K, int64_t N, int k_max);
LYS_API int lys_omp_encode(const float* X, int64_t x_feat_stride, int64_t x_sig_stride,
                   const float* D, int64_t ldd, const float* G,
                   int n, int K, int64_t N, int k_max, float tol, int strict,
                   int32_t* idx, float* val, int32_t* nsel,
                   float* Z, int64_t z_atom_stride, int64_t z_sig_stride, int32_t* truncated,
                   void* workspace, size_t workspace_bytes, void* stream);

/* The correlation GEMM alone (test/bring-up hook): alpha (C,K) row-major = X^T D.
 * impl: 0 = what lys_bomp_encode uses, 1 = fp32 SIMT, 2 = tcgen05 bf16x3 (n = 64 or 128, K % 256 == 0). */
LYS_API int lys_corr_gemm(const float* X, int64_t x_feat_stride, int64_t x_sig_stride,
                          const float* D, int64_t ldd, int n, int K, int64_t C, float* alpha,
                          int impl, void* stream);

/* ---- sibling coders on the same correlation front end -----------------------------------
 * lys_thresh_encode replaces algorithm 'thresh' (lyssa/sparse_coding.py:636-641 dispatch,
 * :416-425 thresholding): Alpha = D^T X, every signal keeps its k LARGEST SIGNED correlations,
 * Z = Alpha there.  k in [1, K] (nonzero_percentage is resolved by the host: floor(p*K), :420).
 * lys_iht_encode replaces algorithm 'iht' (:671-690, :433-446): Z0 = thresh, then n_iter times
 * Z <- Z - eta * D^T (D Z - X) followed by keeping the k largest |Z| per signal; k <= 32.
 * Outputs and Z addressing as lys_bomp_encode; codes come out in descending order of the
 * selection key, ties to the lower atom index (the reference's argsort order on exact ties is
 * unspecified).  No Gram matrix is needed.  The values are fp32-faithful correlations: an fp32 GEMM on the two-kernel
 * path, the three-product fp16 split accumulated in fp32 in tensor memory on the fused path (n <= 64, K in
 * {256,...,1024}, k <= 10) — the error of an fp32 FMA dot product either way. */
LYS_API size_t lys_thresh_workspace_bytes(int n, int K, int64_t N);
LYS_API int lys_thresh_encode(const float* X, int64_t x_feat_stride, int64_t x_sig_stride,
                      const float* D, int64_t ldd, int n, int K, int64_t N, int k,
                      int32_t* idx, float* val, int32_t* nsel,
                      float* Z, int64_t z_atom_stride, int64_t z_sig_stride,
                      void* workspace, size_t workspace_bytes, void* stream);
/* The selection alone, on correlations the caller already holds: alpha (N,K) row-major (one signal per row).  Replaces
 * the bare functions `thresholding(Alpha, ...)` (lyssa/sparse_coding.py:416-425) and `soft_thresholding(Alpha, ...)`
 * (lyssa/feature_encoding.py:26-37 — the same computation), which take Alpha = D^T X as their argument. */
LYS_API int lys_topk_select(const float* alpha, int K, int64_t N, int k, int32_t* idx, float* val, int32_t* nsel,
                    float* Z, int64_t z_atom_stride, int64_t z_sig_stride, void* stream);
LYS_API size_t lys_iht_workspace_bytes(int n, int K, int64_t N);
LYS_API int lys_iht_encode(const float* X, int64_t x_feat_stride, int64_t x_sig_stride,
                   const float* D, int64_t ldd, int n, int K, int64_t N, int k, float eta, int n_iter,
                   int32_t* idx, float* val, int32_t* nsel,
                   float* Z, int64_t z_atom_stride, int64_t z_sig_stride,
                   void* workspace, size_t workspace_bytes, void* stream);

/* Same call for HOST buffers (what `sparse_encoder.encode(X, D)` is for NumPy arrays):
 * uploads D, forms the Gram matrix, streams X through the device in chunks with copies
 * overlapped with compute, writes idx/val/nsel and/or dense Z back to host memory.
 * Synchronous.  Pinned host buffers give full PCIe bandwidth; pageable ones work.
 * Any of idx/val/nsel/Z may be NULL.  `device` < 0 means the current device. */
LYS_API int lys_bomp_encode_host(const float* X, int64_t x_feat_stride, int64_t x_sig_stride,
                         const float* D, int64_t ldd,
                         int n, int K, int64_t N, int k,
                         int32_t* idx, float* val, int32_t* nsel,
                         float* Z, int64_t z_atom_stride, int64_t z_sig_stride,
                         int device);

/* Dense Z from sparse codes (zero-fill + scatter), same Z addressing as above.
 * replaces `Z = np.zeros(...)` + `Z[Dx, i] = z`, sparse_coding.py:308,:365. */
LYS_API int lys_codes_to_dense(const int32_t* idx, const float* val, int64_t N, int k, int K,
                       float* Z, int64_t z_atom_stride, int64_t z_sig_stride, void* stream);

/* ---- K5+K10: residual and approximation error from sparse codes ------------------------
 * replaces `R = Y - fast_dot(D, X)` (lyssa/dict_learning/ksvd.py:103) and
 * approx_error = ||X - D Z||_F^2 (lyssa/dict_learning/utils.py:14-19).
 * R (optional) is signal-major (N,n), row stride n.  err (optional) is ONE device double.
 * workspace: lys_residual_workspace_bytes. */
LYS_API size_t lys_residual_workspace_bytes(int n, int K, int64_t N);
LYS_API int lys_residual(const float* X, int64_t x_feat_stride, int64_t x_sig_stride,
                 const float* D, int64_t ldd, const int32_t* idx, const float* val,
                 int n, int K, int64_t N, int k, float* R, double* err,
                 void* workspace, size_t workspace_bytes, void* stream);

/* ||A||_F^2 of a contiguous, 16-byte aligned float array into ONE device double (deterministic).  After
 * lys_approx_ksvd_sweep the residual R it maintains IS X - D Z, so this replaces the reference's second dense
 * recomputation `error = approx_error(D, Z, X)` (ksvd.py:220) with one pass over R.  workspace: 32 KB. */
LYS_API int lys_frobenius2(const float* A, int64_t count, double* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K6: users-of-atom index (CSR by atom) ---------------------------------------------
 * replaces the per-atom scan `omega_k = X[k, :] != 0`, ksvd.py:111.
 * rowptr (K+1) int32; entries (N*k) int32, entry = i*k + slot, ascending inside each atom
 * (deterministic).  Slots with idx < 0 or val == 0 are not users (ksvd.py:111). */
LYS_API size_t lys_atom_csr_workspace_bytes(int K, int64_t N, int k);
LYS_API int lys_build_atom_csr(const int32_t* idx, const float* val, int64_t N, int k, int K,
                       int32_t* rowptr, int32_t* entries,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- K7-K9: approximate K-SVD sweep -----------------------------------------------------
 * replaces the atom loop of approx_ksvd, ksvd.py:105-124: sequentially for every atom,
 * d <- normalize(R_k x_k^T) (:116-119), x_k <- R_k^T d (:121), residual refresh (:123),
 * with R_k never materialised (R_k = R[:,users] + d x).  D (n,K) row-major and val are
 * updated IN PLACE as in the reference; R (N,n) is the residual from lys_residual and is
 * kept current.  unused (K) int32 receives 1 for atoms without users (:112-115).
 * `comm` is NULL for one GPU, or a handle from lys_comm_create when the signals are
 * sharded over ranks: the per-atom sums (2n+3 fixed-point words: the look-ahead form of
 * R_k x, x.x and the user count) are then all-reduced inside the kernel through peer-mapped
 * mailboxes (see below).  The reduction is integer arithmetic, so the sweep is bitwise
 * reproducible and every rank ends with bit-identical D.  `idx` is the code index array the
 * CSR was built from (N,k). */
LYS_API size_t lys_ksvd_sweep_workspace_bytes(int n, int K, int64_t N, int k);
LYS_API int lys_approx_ksvd_sweep(float* R, float* D, int64_t ldd,
                          const int32_t* idx, float* val,
                          const int32_t* rowptr, const int32_t* entries,
                          int n, int K, int64_t N, int k, int n_cycles,
                          int32_t* unused, void* comm,
                          void* workspace, size_t workspace_bytes, void* stream);

/* ---- exact K-SVD sweep ------------------------------------------
 * replaces the atom loop of ksvd(), lyssa/dict_learning/ksvd.py:19-43: per atom, (d, x) <- top singular triplet of
 * R_k = R[:,users] + d x (:36-39; the reference calls scikit-learn's randomized_svd(R_k, 1, n_iter=10)), then
 * R[:,users] = R_k - d x (:41).  Same in-place contract, CSR and residual as lys_approx_ksvd_sweep.  The triplet comes
 * from the n x n Gram matrix R_k R_k^T (one integer all-reduce per atom, then a replicated eigen-solve, ksvd_exact.cu):
 * d is the top left singular vector to ~1e-6 wherever the two largest singular values differ by more than 0.4 %,
 * with the sign that keeps d.d_old >= 0 (the reference's sign is arbitrary).  n <= 64; single GPU (comm is not taken:
 * the per-atom payload is n(n+1)/2 words); other n return LYS_EUNSUPPORTED. */
LYS_API size_t lys_ksvd_exact_workspace_bytes(int n, int K);
LYS_API int lys_ksvd_exact_sweep(float* R, float* D, int64_t ldd, float* val,
                         const int32_t* rowptr, const int32_t* entries,
                         int n, int K, int64_t N, int k, int n_cycles,
                         int32_t* unused, void* workspace, size_t workspace_bytes, void* stream);

/* ---- K8 / K14 helpers -------------------------------------------------------------------
 * norm_cols: D[:,c] /= (||D[:,c]||_2 + eps), lyssa/utils/math.py:65-71 (eps = 2^-52).
 * gather_cols: D[:,j] = X[:, cols[j]], the device half of init_dictionary
 * (lyssa/dict_learning/utils.py:64) and of unused-atom replacement (ksvd.py:205). */
LYS_API int lys_norm_cols(float* D, int64_t ldd, int n, int K, void* stream);
LYS_API int lys_gather_cols(const float* X, int64_t x_feat_stride, int64_t x_sig_stride, int n,
                    const int64_t* cols, int n_cols, float* D, int64_t ldd, const int32_t* dst_cols,
                    void* stream);

/* ---- K12/K13: online dictionary learning ------------------------------------------------
 * accumulate: A = beta*A + Z_b Z_b^T, B = beta*B + X_b Z_b^T from sparse codes
 * (lyssa/dict_learning/online_dict_learn.py:84-85).  A (K,K), B (n,K) row-major.  Every sum runs in a fixed order
 * (users of an atom ascending), so A and B are bitwise reproducible and identical on every rank that is given the
 * same minibatch — the multi-GPU learner all-gathers the minibatch's codes and signals and accumulates redundantly
 * instead of all-reducing the K x K statistics.  beta == 0 writes the sums alone.
 * update: D <- norm_cols(clamp(D + (B - D A) diag(1/(A_kk + eps)))), :91-98 (Jacobi, stale
 * D A; clamp only if non_neg). */
LYS_API size_t lys_odl_accumulate_workspace_bytes(int K, int64_t b, int k);
LYS_API int lys_odl_accumulate(const float* Xb, int64_t x_feat_stride, int64_t x_sig_stride,
                       const int32_t* idx, const float* val, int n, int K, int64_t b, int k,
                       float beta, float* A, float* B, void* workspace, size_t workspace_bytes, void* stream);
LYS_API size_t lys_odl_update_workspace_bytes(int n, int K);
LYS_API int lys_odl_update_dict(float* D, int64_t ldd, const float* A, const float* B, int n, int K,
                        int non_neg, void* workspace, size_t workspace_bytes, void* stream);

/* ---- ScSPM spatial-pyramid pooling from sparse codes (first "next" row) -----
 * replaces the per-image loop of sc_spm_extractor.encode, lyssa/feature_extract/spatial_pyramid.py:45-97,
 * with the pooling operators of lyssa/feature_extract/pooling.py:4-26, for a batch of images:
 * patch p belongs to image patch_img[p], its top-left pixel is (patch_pos[2p], patch_pos[2p+1]) = (row, col)
 * (:60-63), img_hw[2i], img_hw[2i+1] = height, width of image i, `levels` is a HOST array (e.g. {1,2,4}).
 * F (n_imgs, total_cells*K) row-major receives, per image, the (cell, atom) features in the reference's order
 * (level-major, cell-major, atom-minor, :94-96).  pooling: 0 = max |z| (sc_max_pooling), 1 = sum, 2 = average
 * over the patches of the cell; l2_normalize != 0 applies x/(||x||+eps) per non-empty cell
 * (feature_extract/preproc.py:8-15).  cell_count (n_imgs*total_cells int32) is scratch / returns the patches per cell. */
LYS_API int lys_spm_total_cells(const int32_t* levels, int n_levels);
LYS_API int lys_spm_pool(const int32_t* idx, const float* val, int64_t N, int k, int K,
                         const int32_t* patch_img, const float* patch_pos, float patch_size,
                         const int32_t* img_hw, int n_imgs, const int32_t* levels, int n_levels,
                         int pooling, int l2_normalize, float* F, int32_t* cell_count, void* stream);

/* ---- dense SIFT descriptors (second "next" row) -------------------------------------
 * replaces DsiftExtractor.process_image / extract_sift_patches / normalize_sift,
 * lyssa/feature_extract/dsift.py:75-162, for one grayscale image (H, W) float32 on the device.
 * lys_dsift_grid: the sampling grid of :101-109 (n_h x n_w patches, offsets of the first patch).
 * gh25 / gw25: the 5x5 Gaussian-derivative kernels of gen_dgauss (:23-35) and bin_weights (4, patch_size):
 * the separable factor of the bilinear weight matrix (:57-73) — HOST arrays (a few dozen floats the host side
 * computes with the reference's own formulas).  desc (n_h*n_w, 128) row-major, descriptor element
 * angle*16 + bin as in :139-141; pos (n_h*n_w, 2) = top-left (row, col) of every patch, patch order of :107-109. */
LYS_API size_t lys_dsift_workspace_bytes(int H, int W);
LYS_API int lys_dsift_grid(int H, int W, int grid_spacing, int patch_size, int* n_h, int* n_w, int* off_h, int* off_w);
LYS_API int lys_dsift(const float* img, int64_t row_stride, int H, int W, int grid_spacing, int patch_size,
                      float nrml_thres, float sift_thres, const float* gh25, const float* gw25, const float* bin_weights,
                      float* desc, float* pos, void* workspace, size_t workspace_bytes, void* stream);
/* Same for n_imgs images of ONE size: `imgs` is a HOST array of device pointers; descriptors / positions of image i
 * land at desc + i*P*128, pos + i*P*2 (P = n_h*n_w).  One pair of launches per 128 images; the workspace holds the
 * orientation maps of as many images as fit (>= lys_dsift_workspace_bytes(H, W), ideally n_imgs times that). */
LYS_API int lys_dsift_batch(const float* const* imgs, int n_imgs, int64_t row_stride, int H, int W, int grid_spacing, int patch_size,
                    float nrml_thres, float sift_thres, const float* gh25, const float* gw25, const float* bin_weights,
                    float* desc, float* pos, void* workspace, size_t workspace_bytes, void* stream);

/* ---- multi-GPU: peer-mapped exchange buffers for the sweep's per-atom all-reduce ---------
 * One process per GPU.  Each rank creates a comm (allocates its exchange buffer), exports
 * a 64-byte handle, the host (torch.distributed) all-gathers the handles, and every rank
 * opens its peers'.  No reference counterpart: the reference is single-host
 * multiprocessing (lyssa/utils/__init__.py:92-146). */
#define LYS_COMM_HANDLE_BYTES 64
LYS_API int lys_comm_create(int rank, int world, void** comm);
LYS_API int lys_comm_export(void* comm, unsigned char handle[LYS_COMM_HANDLE_BYTES]);
LYS_API int lys_comm_connect(void* comm, const unsigned char* all_handles /* world*64 bytes */);
LYS_API int lys_comm_destroy(void* comm);

#ifdef __cplusplus
}
#endif
#endif /* LYSSA_B200_H */
