/* wst2d.h — C ABI of libwst_b200.so: batched 2-D wavelet scattering features on one B200.
 *
 * The reference (a pure-Python repo) has no FFI for this path; the interface it exposes is the
 * Python call surface below, and these entry points are what a ctypes binding for that surface
 * binds (INTEGRATION.md shows the stub):
 *
 *   reference interface replaced                                   entry point
 *   -------------------------------------------------------------  ---------------------------
 *   kymatio `Scattering2D(J, shape, L, max_order)` constructed at   wst2d_plan_create
 *     src/training/train_and_save_model.py:359,
 *     src/inference/inference.py:242,
 *     src/visualization/visualize_features.py:210,
 *     src/visualization/compare_wst_coefficients.py:37
 *   `scattering(channel)` + np.mean/np.std over (-2,-1) at          wst2d_forward (features),
 *     src/training/train_and_save_model.py:368-372,                 wst2d_forward (maps != NULL)
 *     src/inference/inference.py:254-266,                             for the coefficient maps of
 *     src/visualization/visualize_features.py:213-217                 visualize_features.py:222
 *   `load_rgb_image` uint8 HWC -> float32/255 CHW at                wst2d_forward_u8
 *     src/training/train_and_save_model.py:51-56
 *
 * All pointers are plain device or host pointers; no torch types.  Every function returns 0 on
 * success or a negative code, and wst2d_last_error() (thread-local) describes the failure.
 * There is no CPU fallback: creating a plan without a usable CUDA device fails with
 * WST2D_ERR_CUDA.
 */
#ifndef WST2D_H
#define WST2D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct wst2d_plan wst2d_plan;

#define WST2D_OK            0
#define WST2D_ERR_ARG      (-1)   /* bad argument (NULL, non-positive size, 2^J > min(H,W), ...) */
#define WST2D_ERR_UNSUPPORTED (-2) /* configuration outside what the library supports (see wst2d_plan_create_ex) */
#define WST2D_ERR_CUDA     (-3)   /* CUDA runtime error, or no device */

/* Build the filter bank and tables for `Scattering2D(J, shape=(H, W), L, max_order)` on `device`.
 * Geometry follows kymatio: padded side ((M + 2^J) / 2^J + 1) * 2^J, reflect padding,
 * K = 1 + L*J + L^2*J*(J-1)/2 coefficients (1 + L*J when max_order == 1),
 * output maps (H_p / 2^J - 2) x (W_p / 2^J - 2). */
int wst2d_plan_create(wst2d_plan** out, int device, int H, int W, int J, int L, int max_order);
int wst2d_plan_destroy(wst2d_plan* plan);

/* Same with an explicit engine.  The reference builds its transform from whatever image it loads
 * (train_and_save_model.py:355-359), so every (H, W) with 2^J <= min(H, W), rectangular included, and every L is
 * accepted:
 *   WST2D_ENGINE_AUTO        the fused FFT cascade when one is compiled for the padded size (square images of side 32,
 *                            48, 64, 96, 128, 224, 256, 512 at the J of csrc/wst_configs.inc, L <= 8), otherwise
 *                            WST2D_ENGINE_GEMM_SIMT
 *   WST2D_ENGINE_FFT         the fused cascade or WST2D_ERR_UNSUPPORTED
 *   WST2D_ENGINE_GEMM_SIMT   every DFT of the cascade as a dense DFT-matrix product on the fp32 pipe (any size)
 *   WST2D_ENGINE_GEMM_TF32X3 the same products on the tensor cores as 3xTF32 (fp32-level accuracy)
 * The environment variable WST_ENGINE=fft|gemm|gemm_tf32x3 overrides AUTO (A/B measurements). */
enum { WST2D_ENGINE_AUTO = 0, WST2D_ENGINE_FFT = 1, WST2D_ENGINE_GEMM_SIMT = 2, WST2D_ENGINE_GEMM_TF32X3 = 3 };
int wst2d_plan_create_ex(wst2d_plan** out, int device, int H, int W, int J, int L, int max_order, int engine);
/* The engine a plan runs on (WST2D_ENGINE_FFT / _GEMM_SIMT / _GEMM_TF32X3). */
int wst2d_plan_engine(const wst2d_plan* plan);
/* Signals in flight of the fused cascade: CTAs of its persistent grid (SMs x resident CTAs per SM); 0 for the GEMM engines. */
int wst2d_plan_grid(const wst2d_plan* plan);

/* Geometry of a plan; any output pointer may be NULL. */
int wst2d_query(const wst2d_plan* plan, int* K, int* h, int* w, int* Hp, int* Wp);

/* x_dev: [B][C][H][W] float32, contiguous, on the plan's device.
 * feats_dev: [B][C][2][K] float32 — per (patch, channel): mean(K) then population std(K) of each
 *            coefficient map over its h x w pixels (the training layout of
 *            train_and_save_model.py:371-376 once flattened per patch).  May be NULL.
 * maps_dev:  [B][C][K][h][w] float32 coefficient maps, or NULL.
 * Work is enqueued on `cuda_stream` (a cudaStream_t; NULL = default stream) without host sync. */
int wst2d_forward(const wst2d_plan* plan, const float* x_dev, int64_t B, int C,
                  float* feats_dev, float* maps_dev, void* cuda_stream);

/* Same, from uint8 pixels: x_dev is [B][H][W][C] uint8 (PIL / load_rgb_image order); the /255
 * scaling and the HWC -> CHW transpose happen on the device. */
int wst2d_forward_u8(const wst2d_plan* plan, const uint8_t* x_dev, int64_t B, int C,
                     float* feats_dev, float* maps_dev, void* cuda_stream);

/* Whole-scene tiling (BASELINE configs[4]; the reference describes "patch extraction" from large rasters,
 * docs/README.md:327-330, and pins rasterio/tifffile, requirements.txt:32,41, but ships no tiler): the plan's
 * H x W window slides over raster_dev [C][Himg][Wimg] float32 with steps (stride_y, stride_x); tiles are
 * numbered row-major over the ((Himg-H)/stride_y+1) x ((Wimg-W)/stride_x+1) grid and the tiles
 * [tile_begin, tile_begin + tile_count) are processed straight from the raster (no tile copies), so ranks of
 * a multi-GPU job simply take disjoint tile ranges.  feats_dev [tile_count][C][2][K], maps_dev
 * [tile_count][C][K][h][w] or NULL. */
int wst2d_forward_scene(const wst2d_plan* plan, const float* raster_dev, int C, int Himg, int Wimg,
                        int stride_y, int stride_x, int64_t tile_begin, int64_t tile_count,
                        float* feats_dev, float* maps_dev, void* cuda_stream);

/* The reference's "advanced statistics" extractor (extract_advanced_features,
 * src/training/train_and_save_model.py:58-112 = src/inference/inference.py:181-235), 18 statistics per channel
 * in the reference's order: mean std var min max range skew kurt cv p10 p25 p50 p75 p90 iqr mad grad_mean
 * edge_density.  x_dev: float32 [B][C][H][W], or uint8 [B][H][W][C] when is_u8 != 0; out_dev: float32
 * [B][C][18].  Needs no plan.  The image and its |laplace| map (2*H*W floats) must fit in shared memory
 * (128x128 does); inputs are assumed finite (the reference drops non-finite pixels).  Errors are reported by
 * wst2d_advanced_stats_last_error(). */
int wst2d_advanced_stats(int device, const void* x_dev, int is_u8, int64_t B, int C, int H, int W,
                         float* out_dev, void* cuda_stream);
const char* wst2d_advanced_stats_last_error(void);

/* The five noise models of the reference's robustness sweep (src/preprocessing/add_noise.py:14-72:
 * add_gaussian_noise, add_salt_and_pepper_noise, add_speckle_noise, add_poisson_noise, add_uniform_noise) for a
 * batch of uint8 images [B][H][W][C] resident on the device; out_dev has the same shape (out_dev == img_dev is
 * allowed) and feeds wst2d_forward_u8 directly.  intensity is the reference's 0..100 percentage.
 *   wst2d_add_noise        draws from a counter-based generator (Philox4x32-10, key = seed, counter = element
 *                          index): same distributions as the reference's numpy calls, different stream.
 *   wst2d_add_noise_draws  the caller supplies the draws numpy produced for the reference, and the output is then
 *                          the reference's bit for bit.  draws_dev: float64 [B][H][W][C] for gaussian (already
 *                          scaled by sigma), speckle (standard normal) and uniform; int64 [B][H][W][C] Poisson
 *                          counts; for salt_and_pepper int64 [B][2 (salt, pepper)][2 (row, col)][n_coords]
 *                          (n_coords is ignored by the other models).
 * Errors: wst2d_noise_last_error(). */
enum {
    WST2D_NOISE_GAUSSIAN = 0,
    WST2D_NOISE_SALT_AND_PEPPER = 1,
    WST2D_NOISE_SPECKLE = 2,
    WST2D_NOISE_POISSON = 3,
    WST2D_NOISE_UNIFORM = 4
};
int wst2d_add_noise(int device, int kind, double intensity, const uint8_t* img_dev, int64_t B, int H, int W, int C,
                    uint64_t seed, uint8_t* out_dev, void* cuda_stream);
int wst2d_add_noise_draws(int device, int kind, double intensity, const uint8_t* img_dev, int64_t B, int H, int W,
                          int C, const void* draws_dev, int64_t n_coords, uint8_t* out_dev, void* cuda_stream);
const char* wst2d_noise_last_error(void);

/* Host-buffer convenience path (what a drop-in extractor calls): x_host [B][C][H][W] float32 and
 * feats_host [B][C][2][K] live in host memory (pinned for full overlap); copies are chunked and
 * double-buffered against compute on two internal streams.  Synchronous on return. */
int wst2d_forward_host(const wst2d_plan* plan, const float* x_host, int64_t B, int C, float* feats_host);
/* Same from uint8 pixels as PIL delivers them, x_host [B][H][W][C] (load_rgb_image's input,
 * src/training/train_and_save_model.py:51-56): a quarter of the host-to-device bytes. */
int wst2d_forward_host_u8(const wst2d_plan* plan, const uint8_t* x_host, int64_t B, int C, float* feats_host);

/* Debug/test export of the plan's full-resolution Fourier-domain filters, host pointers:
 * psi_hat [J*L][Hp][Wp], phi_hat [Hp][Wp]; either may be NULL. */
int wst2d_plan_filters(const wst2d_plan* plan, float* psi_hat, float* phi_hat);

/* Number of kernels wst2d_forward launches for a batch of B*C signals (for launch accounting): the fused cascade
 * pools in-kernel, so this is 1 — 2 when the batch is a few persistent-grid waves with a ragged last one, whose
 * signals then run as a second, split launch (wst2d_forward_host launches that per chunk); the GEMM engines launch one
 * kernel per matrix product of the cascade and chunk of signals.
 * Environment knobs read per call, for A/B measurements only (results are bit-identical either way):
 *   WST_NO_SPLIT=1        never share a signal among CTAs (small batches, ragged last wave)
 *   WST_NO_TAIL_SPLIT=1   keep small batches split, run a ragged last wave as a whole extra wave
 *   WST_STATIC_SCHED=1    fixed-stride assignment of signals to CTAs instead of the ticket counter
 *   WST_HOST_NO_RAMP=1    wst2d_forward_host: equal chunks from the first one on
 *   WST_HOST_CHUNK_SIGNALS=n   wst2d_forward_host: signals per chunk (default 6 waves) */
int wst2d_launch_count(const wst2d_plan* plan, int64_t B, int C);

/* Optional per-kernel timing for bench.py's roofline: when enabled, wst2d_forward records CUDA events
 * around each launch on the caller's stream; wst2d_profile_read synchronises the device, returns the
 * summed durations (ms) since the last read and the number of cascade launches, and resets. */
int wst2d_profile(wst2d_plan* plan, int enable);
int wst2d_profile_read(wst2d_plan* plan, double* cascade_ms, double* pool_ms, int* cascade_launches);

/* Debug: run the cascade over nsig signals with a cycle-counting executor and return, per phase tag
 * (csrc/wst_cascade.h PhaseKind * 8 + level; ntags must equal wst2d_debug_num_phase_tags()), the SM
 * cycles CTA 0 spent in that phase.  Synchronous; fused cascades only.  Used by tools/phase_profile.py. */
int wst2d_debug_phase_cycles(const wst2d_plan* plan, const float* x_dev, int64_t nsig, int64_t* cycles_host,
                             int ntags);
int wst2d_debug_num_phase_tags(void);

/* Measured fp32 FMA rate of `device` in TFLOP/s (dependent-chain-free FMA loop on all SMs): the
 * denominator of the compute roofline, which MEASURED_PEAKS.json does not carry. */
int wst2d_fma_peak(int device, double* tflops);

const char* wst2d_last_error(void);
const char* wst2d_version(void);

#ifdef __cplusplus
}
#endif
#endif /* WST2D_H */
