/* tda_b200.h — C-ABI of the B200-native windowed-TDA engine (libtda_b200.so).
 *
 * The reference (Ignaciagothe/tda-eeg-audio) has no FFI of its own: its hot path is a chain of
 * Python calls into scipy / numpy / ripser / persim.  Each entry point below replaces one of
 * those call sites (cited per function) and is what a ctypes binding on the reference side
 * would load.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  Unless the name ends in _host, every pointer is a
 *     DEVICE pointer owned by the caller; the library allocates nothing persistent.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant, and
 *     performs no hidden synchronisation.  *_host variants take host pointers, stage through
 *     their own pinned/device buffers and return after the results are in host memory.
 *   - return value: 0 = ok, <0 = argument error (TDA_E_*), >0 = cudaError_t from the runtime.
 *   - data-dependent conditions are reported per item in a `status` array, never by aborting
 *     the batch (the reference's convention: a failing unit yields NaN / is skipped,
 *     /root/reference/scripts/utils.py:188-191,
 *     /root/reference/scripts/tda_eeg_classification_v2.py:565-567).
 */
#ifndef TDA_B200_H
#define TDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDA_E_ARG (-1)      /* bad argument (null pointer, negative size, ...)          */
#define TDA_E_SIZE (-2)     /* size outside what the kernel family supports              */
#define TDA_E_WORKSPACE (-3) /* workspace too small (see the matching *_workspace_bytes)  */

/* per-item status bits */
#define TDA_ST_OK 0
#define TDA_ST_H1_TRUNCATED 1 /* more H1 bars than cap1; counts[.,1] holds the true count   */
#define TDA_ST_NAN_INPUT 2    /* NaN in the upper triangle; those edges were left out       */
#define TDA_ST_INTERNAL 4     /* internal capacity exhausted on the last tier (never for N<=64) */

int tda_version(void);
/* number of kernels this library has launched since load (all entry points; thread-safe) */
unsigned long long tda_launch_count(void);
/* Per-kernel timing with CUDA events recorded on the launching stream around every launch the
 * library makes while enabled (enable(1) clears earlier samples, enable(0) stops and clears).
 * query() synchronises the recorded events of kernel `kernel` ("rips_small_w2", "rips_small_w4",
 * "rips_small_w64", "pers_features", "aggregate_windows", ...) and returns their summed
 * duration and count. */
int tda_profile_enable(int on);
int tda_profile_query(const char* kernel, double* total_ms, int* launches);

/* ------------------------------------------------------------------------------------------
 * Vietoris–Rips persistent homology, H0 and H1, Z/2, on a batch of small distance matrices.
 * Replaces: ripser.ripser(dm, maxdim=1, thresh=t, distance_matrix=True)["dgms"]
 *   /root/reference/scripts/utils.py:140 (compute_eeg_persistence)
 *   /root/reference/scripts/utils.py:131 (compute_audio_persistence, after the distance stage)
 *   /root/reference/scripts/tda_eeg_classification_v2.py:170-175 (compute_persistence_diagram)
 *
 * D       (B, N, ld) float32, item stride `strideB` elements (0 => N*ld); only D[b][i][j], i<j
 *         is read (ripser reads the upper triangle of its float32 copy).  2 <= N <= 64.
 *         ld == 0: D is the CONDENSED form, (B, N(N-1)/2) float32 (strideB 0 => N(N-1)/2): the upper
 *         triangle in row-major order (0,1),(0,2),...,(0,N-1),(1,2),... -- exactly the vector
 *         `DParam` that ripser.py builds from the dense matrix and hands its C++ core
 *         (`rips_dm(DParam, N, coeff, maxdim, thresh, ...)`), i.e. the
 *         reference's own FFI format for this call; half the bytes of the dense form.
 * thresh  edges longer than thresh are absent; +inf => no threshold (ripser would substitute the
 *         enclosing radius, which yields the same diagrams).
 * bd0     (B, N, 2) float32  (birth, death) of H0, finite bars in ascending death order
 *         (zero-length bars omitted), then one (0, +inf) per surviving component.
 * pr0     (B, N, 2) int64    (birth vertex, death edge index C(i,2)+j), -1 = none.
 * bd1     (B, cap1, 2) float32 (birth, death) of H1 in ripser's emission order (descending
 *         birth edge in the filtration order); zero-persistence pairs omitted; (b, +inf) for
 *         cycles still alive at thresh.
 * pr1     (B, cap1, 2) int64  (birth edge index, death triangle index C(a,3)+C(b,2)+c), -1.
 * counts  (B, 2) int32  number of rows written for H0 / H1 (true H1 count even if > cap1).
 * status  (B) int32  TDA_ST_* bits.
 * ws      workspace of at least tda_rips_h01_workspace_bytes(B, N) bytes (device).
 * Any of pr0 / pr1 may be NULL (indices not wanted).
 */
size_t tda_rips_h01_workspace_bytes(int B, int N);
int tda_rips_h01_batched(const float* D, int B, int N, int ld, long long strideB, float thresh,
                         float* bd0, long long* pr0, float* bd1, long long* pr1, int* counts,
                         int cap1, int* status, void* ws, size_t ws_bytes, void* stream);

/* Rips H0+H1 for big clouds, 2 <= N <= 2048 (the 1,000-2,000 point Takens clouds of the scaling
 * stress, BASELINE.json configs[4]; the 97-248 point Takens clouds of the audio path; any batch
 * above the 64-point engine).  Same outputs and conventions as tda_rips_h01_batched; differences:
 * `npts` (may be NULL) gives the point count of every item (<= N, the leading npts x npts block of
 * the ld x ld matrix is used), and the H0 output has an explicit row capacity cap0 (bd0 (B, cap0, 2)).
 * Phases over a chunk of clouds (edge keys + stable radix sort + rank matrices: per cloud in shared
 * memory up to 256 points, grid-wide over L2-resident groups of clouds above; Kruskal; first-cofacet /
 * apparent-pair classification of every edge) followed by one CTA per cloud for the serial part
 * (cocycle sweep over the edges a live class can see).  The chunk size follows from `ws_bytes`;
 * tda_rips_h01_large_workspace_bytes returns a size that holds min(B, what fits 48 GB) clouds.
 * Replaces ripser(...) inside compute_audio_persistence, /root/reference/scripts/utils.py:131. */
size_t tda_rips_h01_large_workspace_bytes(int B, int N);
int tda_rips_h01_large(const float* D, const int* npts, int B, int N, int ld, long long strideB,
                       float thresh, float* bd0, long long* pr0, int cap0, float* bd1, long long* pr1,
                       int cap1, int* counts, int* status, void* ws, size_t ws_bytes, void* stream);

/* Same contract with HOST pointers: chunks the batch, overlaps H2D / kernels / D2H on internal
 * streams, returns when all outputs are in host memory.  `device` is the CUDA ordinal. */
int tda_rips_h01_host(const float* D, int B, int N, float thresh, float* bd0, long long* pr0,
                      float* bd1, long long* pr1, int* counts, int cap1, int* status, int device);
/* The same with the condensed input of ripser's C++ entry (see `ld == 0` above): Dc (B, N(N-1)/2)
 * float32 HOST.  The batched counterpart of ripser.py's `DRFDM(DParam, maxdim, thresh, coeff)`. */
int tda_rips_h01_condensed_host(const float* Dc, int B, int N, float thresh, float* bd0, long long* pr0,
                                float* bd1, long long* pr1, int* counts, int cap1, int* status,
                                int device);

/* ------------------------------------------------------------------------------------------
 * Persistence statistics / entropy features of a batch of diagrams.
 * Replaces: extract_features  /root/reference/scripts/utils.py:144-177
 *         ≡ extract_persistence_features  /root/reference/scripts/tda_eeg_classification_v2.py:179-250
 * bd     (B, cap, 2) float32 rows (birth, death), padded; row count of item b is
 *        counts[b*count_stride] (clamped to cap) — pass counts+0 / counts+1 with stride 2 for the
 *        H0 / H1 outputs of tda_rips_h01_batched.
 * feats  item b gets 11 float64 at feats[b*feat_stride ...]: n_features, n_essential, mean_birth,
 *        std_birth, mean_death, std_death, mean_persistence, std_persistence, max_persistence,
 *        total_persistence, persistence_entropy (np.std ddof=0; std := 0 when <= 1 finite row;
 *        all zero except n_essential when no finite row).
 */
int tda_pers_features(const float* bd, int cap, const int* counts, int count_stride, int B,
                      double* feats, int feat_stride, void* stream);

/* Mean / std (ddof=0) over the windows of every (recording, band, H0|H1, feature).
 * Replaces: /root/reference/scripts/tda_eeg_classification_v2.py:429-436.
 * feats (R, Bd, Wn, 2, 11) float64 -> table (R, Bd*44) float64, column order of
 * /root/reference/features/feature_names.txt: band*44 + feat*4 + {h0_mean,h0_std,h1_mean,h1_std}. */
int tda_aggregate_windows(const double* feats, int R, int Bd, int Wn, double* table, void* stream);

/* ------------------------------------------------------------------------------------------
 * Audio front end (FP64).
 * tda_resample_poly_f64 replaces sig_proc.resample_poly(audio, fs_target, fs_audio) in resample_audio,
 *   /root/reference/scripts/utils.py:77-79.  Computes y = upfirdn(h_padded, x, up, down)[n_pre_remove :
 *   n_pre_remove + n_out] for every row, zero padding at both ends (scipy's padtype='constant').
 *   x (n_seq, n_in) rows of stride x_stride; up <= 8 and down coprime (scipy reduces them by their
 *   gcd); hpoly (up, qmax) DEVICE, hpoly[p][q] = h_padded[p + q*up] with zeros beyond the filter —
 *   the Kaiser FIR scipy designs with firwin, a host-side constant like the Butterworth
 *   coefficients; y (n_seq, n_out) rows of stride y_stride.  n_seq <= 65535 per call.
 * tda_hilbert_envelope_f64 replaces np.abs(sig_proc.hilbert(s)) in compute_envelope,
 *   /root/reference/scripts/utils.py:58-59 (the butter(4)+filtfilt that follows is tda_filtfilt_f64, form 1).
 *   cuFFT Z2Z plans are cached per (device, T, n_seq) for the life of the library.
 *   ws: tda_hilbert_envelope_workspace_bytes(n_seq, T) bytes. */
int tda_resample_poly_f64(const double* x, long long n_seq, long long n_in, long long x_stride, int up, int down,
                          const double* hpoly, int qmax, long long n_pre_remove, long long n_out, double* y,
                          long long y_stride, void* stream);
size_t tda_hilbert_envelope_workspace_bytes(long long n_seq, long long T);
int tda_hilbert_envelope_f64(const double* x, long long n_seq, long long T, long long x_stride, double* env,
                             long long env_stride, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Zero-phase IIR filtering of many sequences at once, FP64, scipy's exact recursion
 * (odd extension by padlen, zi scaled by the first sample, forward pass, reverse, second pass).
 * Replaces: signal.sosfiltfilt(sos, x)  /root/reference/notebooks/1_preprocesamiento.ipynb:262-263  (form 0)
 *           sig_proc.filtfilt(b, a, x)  /root/reference/scripts/utils.py:63,74                    (form 1)
 * x      n_seq rows of T float64 samples, row stride x_stride (0 => T).
 * form   0: coef = n_bands x n x 6 (sos rows b0 b1 b2 a0 a1 a2), zi = n_bands x n x 2 (sosfilt_zi), n <= 4
 *        1: coef = n_bands x 2 x n (b then a), zi = n_bands x (n-1) (lfilter_zi), n <= 9 taps
 *        (coefficients / zi are HOST pointers: a few dozen doubles designed once per band.)
 * padlen 3*ntaps as scipy computes it (27 for the order-4 band-pass, 15 for the order-4 low-pass).
 * y      (n_bands, n_seq, T) float64: every band's filter applied to every sequence.
 * ws     tda_filtfilt_workspace_bytes(...) bytes (the padded forward-pass intermediate).
 */
size_t tda_filtfilt_workspace_bytes(long long n_seq, int n_bands, long long T, int padlen);
int tda_filtfilt_f64(const double* x, long long n_seq, long long T, long long x_stride, int form,
                     int n_bands, int n, const double* coef, const double* zi, int padlen, double* y,
                     void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Sliding windows -> Pearson correlation -> correlation distance, fused (windows never stored).
 * Replaces: create_sliding_windows      /root/reference/notebooks/1_preprocesamiento.ipynb:314-364
 *           compute_correlation_matrix  /root/reference/notebooks/2_graph_construction.ipynb:86-97
 *           correlation_to_distance     /root/reference/notebooks/2_graph_construction.ipynb:100-122
 * x       (R, C, T) float64, recording stride strideR (0 => C*T); windows start at w*step,
 *         W = (T-win)/step + 1 of them (none if T < win).
 * method  0 "euclidean" sqrt(2(1-r)) [the pipeline's], 1 "abs" 1-|r|, 2 "standard" 1-r, 3 "sqrt" sqrt(1-r^2).
 * D       float32 distances, corr float64 correlations (either may be NULL); window (rec, w) is
 *         written at rec*strideO + w*C*C (strideO 0 => W*C*C), both triangles, zero diagonal in D.
 */
int tda_corrdist_windows(const double* x, int R, int C, long long T, long long strideR, int win, int step,
                         int method, float* D, double* corr, long long strideO, void* stream);

/* correlation_to_distance on one n x n float64 correlation matrix (float64 out), methods as above.
 * Replaces: /root/reference/notebooks/2_graph_construction.ipynb:100-122. */
int tda_corr_to_dist_f64(const double* corr, int n, int method, double* dist, void* stream);

/* (D + D^T)/2, zero diagonal, max(.,0), cast to float32 for B matrices: the float64 preamble of
 * compute_eeg_persistence (/root/reference/scripts/utils.py:137-139) ≡ compute_persistence_diagram
 * (/root/reference/scripts/tda_eeg_classification_v2.py:166-168) before ripser's float32 cast. */
int tda_symmetrize_f64_to_f32(const double* D, long long B, int n, float* out, void* stream);

/* Input checker of the feature path, batched: validate_distance_matrix
 * (/root/reference/scripts/tda_eeg_classification_v2.py:110-140) for B float64 n x n matrices at once.
 * flags[b] = TDA_DM_* bits; stats (B, 3) float64 = { max |D - D^T|, min D, max |diag D| } (the numbers
 * the reference prints in its messages; NaN as soon as one NaN is involved, like np.max / np.min). */
#define TDA_DM_ASYMMETRIC 1 /* not np.allclose(D, D.T, rtol=1e-5, atol=1e-8)  */
#define TDA_DM_NEGATIVE 2   /* an entry below -1e-10                          */
#define TDA_DM_DIAGONAL 4   /* not np.allclose(diag, 0, atol=1e-10)           */
#define TDA_DM_NAN 8
#define TDA_DM_INF 16
int tda_validate_distance_f64(const double* D, long long B, int n, int* flags, double* stats, void* stream);

/* ------------------------------------------------------------------------------------------
 * Audio side: delay, Takens embedding, pairwise distances.
 * tda_compute_tau     replaces compute_tau  /root/reference/scripts/utils.py:92-104
 *   wins (B rows of L float64, row stride `stride`, 0 => L); max_lag < 0 => L/4 (the default);
 *   tau[b] = first i in [1, min(max_lag, L-1)) with autocorrelation <= 0, else max(max_lag/10, 1).
 * tda_takens_cloud    replaces takens_embedding (utils.py:107-116) + the min-max normalisation of
 *   compute_audio_persistence (utils.py:127-130, skipped when normalise == 0)
 *   tau[b*tau_stride] is the delay of item b (tau_stride 0 => one shared delay);
 *   pts (B, ldp, dim) float64, npts[b] = ceil((L-(dim-1)tau)/subsample) clamped to [0, ldp].
 * tda_pairwise_dist_f32  replaces sklearn pairwise_distances inside ripser(point_cloud)
 *   (float64 Gram trick, max(.,0), zero diagonal, sqrt, cast to float32); npts may be NULL (= ldp);
 *   D (B, ld, ld) float32, only the leading npts[b] x npts[b] block is written. */
int tda_compute_tau(const double* wins, long long B, int L, long long stride, int max_lag, int* tau,
                    void* stream);
int tda_takens_cloud(const double* wins, long long B, int L, long long stride, const int* tau,
                     int tau_stride, int dim, int subsample, int normalise, int ldp, double* pts,
                     int* npts, void* stream);
int tda_pairwise_dist_f32(const double* pts, const int* npts, long long B, int ldp, int dim, int ld,
                          float* D, void* stream);

/* ------------------------------------------------------------------------------------------
 * Exact 1-Wasserstein distance between persistence diagrams, batched (L2 ground metric,
 * Euclidean distance to the diagonal, sum of matched costs), float64.
 * Replaces: safe_wasserstein -> persim.wasserstein  /root/reference/scripts/utils.py:180-191
 *           (per-window calls at /root/reference/scripts/tda_eeg_audio_comparison.py:95-96,
 *            /root/reference/scripts/matched_vs_mismatched.py:87-95).
 * bdA (., capA, 2) / bdB (., capB, 2) float32 padded diagrams; row counts nA[i*nA_stride],
 * nB[i*nB_stride] (clamped to the caps).  Rows with a non-finite coordinate are ignored; an empty
 * diagram counts as the single point (0,0), as safe_wasserstein does.  Pair k compares diagram
 * idxA[k] of A with diagram idxB[k] of B (NULL => k), so matched and mismatched pairings need no
 * copies.  out[k] float64.  limA / limB (0 => the cap) are upper bounds on the row counts actually
 * present (rows beyond them are ignored); shared memory is sized by them and linear in limA + limB
 * (about 57 bytes per point: up to ~4,000 points per pair, then TDA_E_SIZE). */
int tda_wasserstein_batched(const float* bdA, const int* nA, int nA_stride, int capA, int limA,
                            const float* bdB, const int* nB, int nB_stride, int capB, int limB,
                            const int* idxA, const int* idxB, long long B, double* out, void* stream);
/* The same for float64 diagrams (persim.wasserstein works in float64; the drop-in call
 * wasserstein(dgm1, dgm2) of arbitrary diagrams goes through this entry, so nothing is rounded). */
int tda_wasserstein_batched_f64(const double* bdA, const int* nA, int nA_stride, int capA, int limA,
                                const double* bdB, const int* nB, int nB_stride, int capB, int limB,
                                const int* idxA, const int* idxB, long long B, double* out, void* stream);

/* End-to-end host entry for the EEG feature path: host distance matrices in, host feature table
 * out (process_file_features, /root/reference/scripts/tda_eeg_classification_v2.py:338-442, for
 * R recordings x Bd bands x Wn windows at once).  D (R,Bd,Wn,N,N) float32 HOST.  Optional host
 * outputs (NULL to skip): bd0 (B,N,2), bd1 (B,cap1,2), counts (B,2), status (B), feats (B,2,11);
 * table (R, Bd*44) float64 is mandatory.  Chunked over three streams; returns when done.  The
 * staging buffers and streams are cached per `device` ordinal, so one process may call with several. */
int tda_eeg_features_host(const float* D, int R, int Bd, int Wn, int N, float thresh, int cap1,
                          float* bd0, float* bd1, int* counts, int* status, double* feats,
                          double* table, int device);
/* The same with condensed windows: Dc (R,Bd,Wn,N(N-1)/2) float32 HOST (half the PCIe bytes). */
int tda_eeg_features_condensed_host(const float* Dc, int R, int Bd, int Wn, int N, float thresh,
                                    int cap1, float* bd0, float* bd1, int* counts, int* status,
                                    double* feats, double* table, int device);
/* The same with the argument type the reference's own per-window call receives: D64 (R,Bd,Wn,N,N)
 * float64 HOST, as compute_eeg_persistence(dm) / compute_persistence_diagram(distance_matrix) take it
 * (/root/reference/scripts/utils.py:135-141, tda_eeg_classification_v2.py:143-176).  Every chunk is
 * symmetrised ((D + D^T)/2), its diagonal zeroed, clamped at 0 and cast to float32 on the device
 * (tda_symmetrize_f64_to_f32) -- the preparation those functions do on the host -- before the Rips
 * kernels.  Twice the PCIe bytes of the float32 entry. */
int tda_eeg_features_f64_host(const double* D64, int R, int Bd, int Wn, int N, float thresh, int cap1,
                              float* bd0, float* bd1, int* counts, int* status, double* feats,
                              double* table, int device);

#ifdef __cplusplus
}
#endif
#endif /* TDA_B200_H */
