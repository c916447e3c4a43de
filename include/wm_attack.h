/* libwmattack — C ABI of the B200-native "tailored attacking layer".
 *
 * Drop-in scope: the differentiable distortion pipeline of
 * yingqichao/video-watermarking-forgery-detection (noise_layers/ and utils/JPEG.py::DiffJPEG).
 * The reference has no FFI: its boundary is Python nn.Modules.  Each entry point below is
 * the device-side replacement of the torch op graph behind ONE reference forward (cited as
 * reference file:line); the Python host layer (wmattack/) re-creates the nn.Module surface
 * on top of these calls through ctypes.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in `_host`.
 *   - Images are float32 [B, C, H, W]; the innermost (W) stride is 1.  Inputs carry element
 *     strides (sb, sc, sh) so that slices such as x[:, :, t] of a [B,3,T,H,W] clip are read
 *     in place; outputs are written densely (NCHW contiguous).
 *   - `stream` is a cudaStream_t passed as void*.  Every call is asynchronous on that
 *     stream, allocates nothing, and never synchronises.
 *   - Return value: 0 = ok; < 0 = WM_E_* invalid-argument code; > 0 = cudaError_t.
 *     wm_last_error() returns a thread-local human-readable message for the last failure.
 *   - There is no CPU path.
 */
#ifndef WM_ATTACK_H
#define WM_ATTACK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WM_ABI_VERSION 4

#define WM_OK 0
#define WM_E_NULL (-1)      /* required pointer is NULL */
#define WM_E_SHAPE (-2)     /* unsupported shape (e.g. H,W not multiples of 16 for DiffJPEG) */
#define WM_E_ALIGN (-3)     /* pointer / stride alignment not met (see each call) */
#define WM_E_ARG (-4)       /* enum / scalar argument out of range */

/* rounding surrogates of the JPEG quantiser */
#define WM_ROUND_ONLY_AT_0 0 /* utils/JPEG.py:482 round_only_at_0 (== JpegSS.round_ss, noise_layers/jpeg.py:255) */
#define WM_ROUND_CUBIC 1     /* utils/JPEG.py:472 diff_round */
#define WM_ROUND_HARD 2      /* torch.round (half to even) */
#define WM_ROUND_FOURIER 3   /* utils/JPEG_utils.py:36 diff_round (9-term Fourier series) */

/* element types of the tensors that cross the boundary in reduced precision (autocast, models/IRNcrop_model.py:340):
 * the entry points that take a `*_dtype` argument read that type directly / store the input gradient in it (cast fused
 * into the kernel's load / store, round-to-nearest-even as torch's .to()); all arithmetic stays float32 */
#define WM_DT_F32 0
#define WM_DT_F16 1
#define WM_DT_BF16 2

int wm_version(void);
const char* wm_last_error(void);

/* Store epilogue (models/IRNp_model.py:674-680): an OPTIONAL argument of the forward entry
 * points marked "ep" below.  When ep != NULL and ep->x != NULL the kernel writes
 *     Quantization( x + (clamp(v, 0, 1) - x) )       [clamp only if clamp01, Quantization only if quantize]
 * instead of its own value v, reading x (dense [B,C,H,W] in the layout of the output, 32-byte aligned) at the output
 * position - no separate pass over the attacked batch, and `y` may be a slice of a K-way batch.  The struct lives in
 * HOST memory and is copied into the launch.  A code path that cannot apply it fails with WM_E_ARG instead of
 * silently ignoring it: wm_jpeg8_fwd (needs W % 8 == 0, aligned, subsample 0), wm_gaussblur (zero border,
 * k in {3,5,7}, W % 4 == 0), wm_median_fwd (W % 4 == 0), wm_resize_fwd, wm_diffjpeg_fwd, wm_gaussnoise_fwd. */
typedef struct wm_store_epilogue {
    const float* x;
    int clamp01;
    int quantize;
} wm_store_epilogue;

/* ------------------------------------------------------------------------------------------
 * DiffJPEG  —  utils/JPEG.py:501-540 (compress_jpeg :256-291, decompress_jpeg :431-469)
 *   x: [B,3,H,W] in [0,1] of x_dtype elements (WM_DT_*), strides (x_sb, x_sc, x_sh) in elements, multiples of
 *      8, base pointer aligned to 8 elements (32 bytes for float32); H, W multiples of 16.  y is float32; the
 *      backward entry points store gx as gx_dtype elements (dense NCHW).
 *   factor: quality_to_factor(quality) (utils/JPEG.py:487); if factor_per_sample != NULL it
 *      is a device array [B] that overrides `factor` per image (quality sweep extension).
 * ------------------------------------------------------------------------------------------ */
int wm_diffjpeg_fwd(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                    float* y, int B, int H, int W,
                    float factor, const float* factor_per_sample, int rounding,
                    const wm_store_epilogue* ep, void* stream);

/* gx = d<gy, DiffJPEG(x)>/dx, recomputed from x (nothing saved by the forward).
 * Replaces autograd over the ~30 saved activations of the reference graph. */
int wm_diffjpeg_bwd(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                    const float* gy, int64_t g_sb, int64_t g_sc, int64_t g_sh,
                    void* gx, int gx_dtype, int B, int H, int W,
                    float factor, const float* factor_per_sample, int rounding, void* stream);

/* Training pair: the forward additionally saves 7 B/px of state — round'(q) of every luminance
 * (dY [B,H,W]) and chroma (dC [B,2,H/2,W/2]) coefficient and the clamp code (0 outside / 1 inside /
 * 2 on a bound) of every output value (clamp_codes [B,H,W/8], 48 bits per 8-pixel row) — and the
 * backward runs from gy + that state alone (no x, no forward recomputation; 31 B/px each way,
 * every rounding mode).  The layouts are private to this pair. */
int wm_diffjpeg_fwd_save(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh, float* y,
                         float* dY, float* dC, uint64_t* clamp_codes, int B, int H, int W,
                         float factor, const float* factor_per_sample, int rounding, void* stream);
int wm_diffjpeg_bwd_saved(const float* gy, int64_t g_sb, int64_t g_sc, int64_t g_sh,
                          const float* dY, const float* dC, const uint64_t* clamp_codes, void* gx, int gx_dtype,
                          int B, int H, int W, void* stream);

/* compress_jpeg.forward (utils/JPEG.py:279-291): rounded quantised coefficients,
 * coef_y [B, H*W/64, 8, 8], coef_cb / coef_cr [B, H*W/256, 8, 8] (block raster order). */
int wm_diffjpeg_compress(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                         float* coef_y, float* coef_cb, float* coef_cr, int B, int H, int W,
                         float factor, const float* factor_per_sample, int rounding, void* stream);

/* decompress_jpeg.forward (utils/JPEG.py:452-469) */
int wm_diffjpeg_decompress(const float* coef_y, const float* coef_cb, const float* coef_cr,
                           float* y, int B, int H, int W,
                           float factor, const float* factor_per_sample, void* stream);

/* ------------------------------------------------------------------------------------------
 * 8x8-unit 4:4:4 JPEG simulators — Jpeg / JpegSS / JpegMask (noise_layers/jpeg.py:214-306)
 * and HiDDeN JpegCompression (noise_layers/jpeg_compression.py:65-159) share one kernel:
 *   yuv = fwd_color * rgb ; zero-pad to x8 ; [subsample] ; C = D X D^T ;
 *   HARD: rint(C/T)*T   SS: ss(C/T)*T   MASK: C*T (T in {0,1}) ; X' = D^T C' D ;
 *   rgb' = inv_color * yuv' ; un-pad.   No clamp.
 * params_host: host pointer to wm_jpeg8_params (copied into the launch).
 * Any H, W >= 1 (the reference is correct for square inputs only, jpeg.py:123-127).
 * ------------------------------------------------------------------------------------------ */
#define WM_JPEG8_HARD 0
#define WM_JPEG8_SS 1
#define WM_JPEG8_MASK 2

typedef struct wm_jpeg8_params {
    float fwd_color[9];   /* row-major 3x3, includes the x255 of jpeg.py:168 where applicable */
    float inv_color[9];   /* row-major 3x3, includes the /255 of jpeg.py:200 where applicable */
    float table[3][64];   /* per channel, [u*8+v], u = vertical frequency: quant steps or keep mask */
    int variant;          /* WM_JPEG8_* */
    int subsample;        /* 0, or 2 = in-block chroma decimation of jpeg.py:202-211 */
} wm_jpeg8_params;

/* x_dtype / gx_dtype (WM_DT_*): float16 / bfloat16 images and gradients are read / stored directly on the vector
 * path (W % 8 == 0, subsample == 0, pointers aligned to 8 elements, strides multiples of 8); any other geometry
 * takes float32 only (WM_E_ARG otherwise).  wm_jpeg8_bwd reads x (float32) only for JpegSS without saved state. */
int wm_jpeg8_fwd(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                 float* y, int B, int H, int W, const wm_jpeg8_params* params_host,
                 const wm_store_epilogue* ep, void* stream);
int wm_jpeg8_bwd(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                 const float* gy, int64_t g_sb, int64_t g_sc, int64_t g_sh,
                 void* gx, int gx_dtype, int B, int H, int W, const wm_jpeg8_params* params_host, void* stream);
/* JpegSS training pair: the forward also saves ss'(q) of every coefficient (d: [B,3,ceil8(H),W],
 * 12 B/px) and the backward runs from gy + d alone (no x, no recompute).  Fast-path geometry only:
 * W % 8 == 0, subsample == 0, 32-byte aligned pointers, strides multiples of 8. */
int wm_jpeg8_fwd_save(const void* x, int x_dtype, int64_t x_sb, int64_t x_sc, int64_t x_sh, float* y, float* d,
                      int B, int H, int W, const wm_jpeg8_params* params_host, void* stream);
int wm_jpeg8_bwd_saved(const float* gy, int64_t g_sb, int64_t g_sc, int64_t g_sh, const float* d, void* gx, int gx_dtype,
                       int B, int H, int W, const wm_jpeg8_params* params_host, void* stream);
/* std_quantization output (noise_layers/jpeg.py:52-82) as a [B,3,Hp,Wp] coefficient image,
 * Hp/Wp = H/W rounded up to x8: the integer-exact parity target for Jpeg. */
int wm_jpeg8_quantised(const float* x, int64_t x_sb, int64_t x_sc, int64_t x_sh,
                       float* coef, int B, int H, int W, const wm_jpeg8_params* params_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * Separable Gaussian blur, depth-wise over N = B*C planes.
 *   border 0 = zero padding  (GaussianBlur, noise_layers/gaussian_blur.py:44-56)
 *   border 1 = reflect       (GF -> kornia GaussianBlur2d, noise_layers/gaussian_filter.py:9)
 * taps_host: k normalised 1-D taps (k odd, k <= 31).  x plane stride = x_sp elements, row
 * stride x_sh.  `adjoint` != 0 applies the transpose operator (== backward); for border 0
 * it is the same filter.  Zero border, k in {3,5,7}: rows on 16-byte boundaries are staged by TMA, any other
 * geometry (W % 4 != 0) by the same kernel fed with cp.async.
 * ------------------------------------------------------------------------------------------ */
int wm_gaussblur(const float* x, int64_t x_sp, int64_t x_sh, float* y, int N, int H, int W,
                 const float* taps_host, int k, int border, int adjoint,
                 const wm_store_epilogue* ep, void* stream);
/* The same filter at the autocast boundary (models/IRNcrop_model.py:340: the attacks run under torch.cuda.amp.autocast):
 * x_dtype / y_dtype are WM_DT_*; ONE side may be float16 / bfloat16 - a typed SOURCE is staged by TMA as it is and widened
 * on the way to registers (the forward of a half image), a typed RESULT is rounded to nearest even in the store (the
 * gradient of a half image) - instead of a separate .float() / .to(dtype) pass over the tensor.  Zero border, k in {3,5,7},
 * rows on 16-byte boundaries (W % 8 == 0 for a 2-byte side, aligned strides); WM_E_ALIGN otherwise (convert first). */
int wm_gaussblur_typed(const void* x, int x_dtype, int64_t x_sp, int64_t x_sh, void* y, int y_dtype, int N, int H, int W,
                       const float* taps_host, int k, void* stream);

/* ------------------------------------------------------------------------------------------
 * k x k median, zero padding, k in {3,5}  (MiddleBlur, noise_layers/middle_filter.py:5-13 ->
 * kornia MedianBlur).  idx (optional, uint8 [N,H] rows of idx_sh >= W bytes) receives the raster position
 * inside the window of the FIRST element equal to the median; wm_median_bwd routes gy through it.
 * Rows on 16-byte boundaries (W % 4 == 0, aligned strides) are staged by TMA, any other geometry by the same
 * kernels fed with cp.async; give idx a 16-byte aligned base and an idx_sh that is a multiple of 16 (the bytes past W
 * are padding) and the backward keeps its idx ring on TMA for every W.
 * ------------------------------------------------------------------------------------------ */
int wm_median_fwd(const float* x, int64_t x_sp, int64_t x_sh, float* y, uint8_t* idx, int64_t idx_sh,
                  int N, int H, int W, int k, const wm_store_epilogue* ep, void* stream);
int wm_median_bwd(const float* gy, const uint8_t* idx, int64_t idx_sh, float* gx, int N, int H, int W, int k, void* stream);
/* Autocast boundary: float16 / bfloat16 source planes through the same TMA ring (widening is exact: the median and its
 * position are those of the float32 image), and gx stored in that type.  Rows on 16-byte boundaries only (W % 8 == 0 for
 * the 2-byte planes; WM_E_ALIGN otherwise); WM_DT_F32 forwards to the entry points above. */
int wm_median_fwd_typed(const void* x, int x_dtype, int64_t x_sp, int64_t x_sh, float* y, uint8_t* idx, int64_t idx_sh,
                        int N, int H, int W, int k, void* stream);
int wm_median_bwd_typed(const float* gy, const uint8_t* idx, int64_t idx_sh, void* gx, int gx_dtype, int N, int H, int W, int k,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * Elementwise attacks over n contiguous floats.  Randomness: Philox4x32-10 keyed by `seed`,
 * counter = element index / 4 + `offset`; if `inject` != NULL that device array [n] is used
 * instead (parity tests inject the reference's random tensor).
 * ------------------------------------------------------------------------------------------ */
/* Device-resident randomness (CUDA-graph capture of the stochastic layers).  Every entry point below that
 * takes (seed, offset) also accepts seed == WM_RNG_FROM_DEVICE with `offset` = a DEVICE pointer (cast to
 * uint64_t) to two uint64 {seed, offset}.  wm_rng_reserve copies a device-resident generator state
 * {seed, next offset} into such a slot and advances the state by `count` Philox counters (one counter =
 * four values) in the same stream: a captured forward then draws fresh numbers at every replay and its
 * backward, given the same slot, regenerates exactly those numbers. */
#define WM_RNG_FROM_DEVICE 0xFFFFFFFFFFFFFFFFull
int wm_rng_reserve(uint64_t* state, uint64_t* slot, uint64_t count, void* stream);

/* Gaussian (noise_layers/gaussian.py:10-17, clamp=1) and GN (noise_layers/gaussian_noise.py:13-16, clamp=0).
 * x_dtype / gx_dtype (WM_DT_*): the image may be float16 / bfloat16, the masked gradient is stored in that type. */
int wm_gaussnoise_fwd(const void* x, int x_dtype, float* y, int64_t n, float mean, float std, int clamp,
                      uint64_t seed, uint64_t offset, const float* inject,
                      const wm_store_epilogue* ep, void* stream);
int wm_gaussnoise_bwd(const float* x, const float* gy, float* gx, int64_t n, float mean, float std,
                      int clamp, uint64_t seed, uint64_t offset, const float* inject, void* stream);
/* Training pair of the clamped layer (Gaussian, noise_layers/gaussian.py:10-17): the forward also writes one
 * bit per value — the pass mask of torch.clamp's backward, 0 <= x + noise <= 1 — into maskbits
 * (4 * ceil(n / 128) words, 16-byte aligned: word 4*(i/128) + j holds value i + 4*l + j at bit l), and the
 * backward is gx = bit ? gy : 0: it reads neither x nor regenerates the noise. */
int wm_gaussnoise_fwd_mask(const void* x, int x_dtype, float* y, uint32_t* maskbits, int64_t n, float mean, float std,
                           uint64_t seed, uint64_t offset, const float* inject, void* stream);
int wm_gaussnoise_bwd_mask(const float* gy, const uint32_t* maskbits, void* gx, int gx_dtype, int64_t n, void* stream);
/* SaltPepper (noise_layers/salt_pepper_noise.py:11-19) */
int wm_saltpepper_fwd(const float* x, float* y, int64_t n, float prob,
                      uint64_t seed, uint64_t offset, const float* inject, void* stream);
int wm_saltpepper_bwd(const float* gy, float* gx, int64_t n, float prob,
                      uint64_t seed, uint64_t offset, const float* inject, void* stream);
/* crop.Dropout (noise_layers/crop.py:142-147): y = rdn > prob ? cover : image */
int wm_dropout_elem_fwd(const float* image, const float* cover, float* y, int64_t n, float prob,
                        uint64_t seed, uint64_t offset, const float* inject, void* stream);
int wm_dropout_elem_bwd(const float* gy, float* g_image, float* g_cover, int64_t n, float prob,
                        uint64_t seed, uint64_t offset, const float* inject, void* stream);
/* dropout.Dropout (noise_layers/dropout.py:14-27): mask [H*W] in {0,1} shared by all planes */
int wm_dropout_mask_fwd(const float* noised, const float* cover, const float* mask_hw, float* y,
                        int64_t planes, int64_t hw, void* stream);
int wm_dropout_mask_bwd(const float* gy, const float* mask_hw, float* g_noised, float* g_cover,
                        int64_t planes, int64_t hw, void* stream);
/* Fill mask_hw[hw] with Bernoulli(keep) in {0,1} from Philox (fast path of dropout.py:21) */
int wm_bernoulli_mask(float* mask_hw, int64_t hw, float keep, uint64_t seed, uint64_t offset, void* stream);
/* Quantization (models/modules/Quantization.py:7-10): rint(x*255)/255; clamp01 != 0 applies
 * the utils/JPEG_utils.py:48 variant (clamp to [0,1] first). */
int wm_quantize8_fwd(const float* x, float* y, int64_t n, int clamp01, void* stream);
/* Cropout (noise_layers/crop.py:128-134): y = cover with image pasted inside the box */
int wm_cropout_fwd(const float* image, const float* cover, float* y, int64_t planes, int H, int W,
                   int h0, int h1, int w0, int w1, void* stream);

/* ------------------------------------------------------------------------------------------
 * Interpolation (F.interpolate(size=..., align_corners=False) semantics, ATen
 * upsample_bilinear2d / upsample_bicubic2d A=-0.75) over N = B*C planes.
 *   mode 0 = bilinear, 1 = bicubic.  The source window (h0, w0, Hin, Win) addresses a crop of
 *   a [N, Hsrc, Wsrc] plane stack (plane stride x_sp, row stride x_sh) so that
 *   Crop (noise_layers/crop.py:48-53) needs no copy; the window must lie inside the plane
 *   (0 <= h0, h0 + Hin <= Hsrc, likewise for w: WM_E_SHAPE otherwise - it is read in place).  clamp01: clamp the result to [0,1]
 *   (Resize, noise_layers/resize.py:53); maskbits (optional, uint32 [N, Hout, ceil(Wout/32)])
 *   receives bit (ox & 31) of word (ox >> 5) = 1 where 0 <= pre-clamp value <= 1.
 * wm_interp_bwd is the exact transpose as a deterministic gather (no atomics): gx is the dense
 *   [N, Hsrc, Wsrc] gradient (zero outside the source window).  gy is first masked by
 *   `maskbits` (from the forward) or, if that is NULL and `pre` != NULL, by 0 <= pre <= 1
 *   (torch.clamp backward).  workspace: device scratch of N*Hin*Wout floats, only needed when
 *   wm_interp_is_tiled(...) == 0 (scale factors outside [0.4, 2.2]).
 * ------------------------------------------------------------------------------------------ */
int wm_interp_fwd(const float* x, int64_t x_sp, int64_t x_sh, int Hsrc, int Wsrc, int h0, int w0, int Hin, int Win,
                  float* y, int N, int Hout, int Wout, int mode, int clamp01,
                  uint32_t* maskbits, void* stream);
int wm_interp_bwd(const float* gy, const float* pre, const uint32_t* maskbits, int N, int Hout, int Wout,
                  float* gx, int Hsrc, int Wsrc, int h0, int w0, int Hin, int Win,
                  int mode, float* workspace, void* stream);
int wm_interp_is_tiled(int Hin, int Win, int Hout, int Wout, int N);

/* ------------------------------------------------------------------------------------------
 * Fused Resize round trip (Resize.forward, noise_layers/resize.py:38-53):
 *   y = clamp( interpolate( interpolate(x, (Hm, Wm)), (H, W) ), 0, 1 )   in ONE kernel,
 * x: N planes [H, W] (plane stride x_sp, row stride x_sh), y dense [N, H, W].
 * Supported when wm_resize_is_fused(...) == 1 (both ratios Hm/H, Wm/W within [0.45, 2.2]); otherwise compose two
 * wm_interp_fwd calls.  Rows on 16-byte boundaries (W % 4 == 0, aligned x / y) are staged by TMA, any other geometry
 * by the same kernel fed with cp.async (no store epilogue there).
 * tables: device workspace of wm_resize_table_floats(H, W, Hm, Wm) floats filled once per
 *   geometry by wm_resize_tables (band starts + weights of the per-axis operators U*D and
 *   their transposes); it may be cached and shared by any number of fwd/bwd calls.
 *   wm_resize_tables also PROVES the geometry: its last 4-byte word (index table_floats - 4, as int32) is 0
 *   iff every band fits the kernel's register / shared-memory windows for both directions; a caller must
 *   read it once per new geometry (after the stream reaches that point) and use wm_interp_fwd / wm_interp_bwd
 *   twice when it is non-zero.
 * maskbits (optional, uint32 [N, H, ceil(W/128), 4], 16-byte aligned): word k of a (row, 128-column
 *   tile) holds in bit l the flag 0 <= pre-clamp value <= 1 of column 128*tile + 4*l + k.
 * wm_resize_bwd is the exact adjoint gx = D^T U^T (gy .* mask) (gy, gx dense), deterministic.
 * ------------------------------------------------------------------------------------------ */
int wm_resize_is_fused(int H, int W, int Hm, int Wm, int N);
int64_t wm_resize_table_floats(int H, int W, int Hm, int Wm);
int wm_resize_tables(float* tables, int H, int W, int Hm, int Wm, int mode, void* stream);
int wm_resize_fwd(const float* x, int64_t x_sp, int64_t x_sh, float* y, int N, int H, int W, int Hm, int Wm,
                  int mode, uint32_t* maskbits, const float* tables, const wm_store_epilogue* ep, void* stream);
int wm_resize_bwd(const float* gy, const uint32_t* maskbits, float* gx, int N, int H, int W, int Hm, int Wm,
                  int mode, const float* tables, void* stream);
/* Autocast boundary: float16 / bfloat16 source planes staged as they are (widened in the row pass), and the adjoint's
 * result stored in that type.  Rows on 16-byte boundaries only (WM_E_ALIGN otherwise), fused geometries only, no store
 * epilogue; WM_DT_F32 forwards to the entry points above. */
int wm_resize_fwd_typed(const void* x, int x_dtype, int64_t x_sp, int64_t x_sh, float* y, int N, int H, int W, int Hm, int Wm,
                        int mode, uint32_t* maskbits, const float* tables, void* stream);
int wm_resize_bwd_typed(const float* gy, const uint32_t* maskbits, void* gx, int gx_dtype, int N, int H, int W, int Hm, int Wm,
                        int mode, const float* tables, void* stream);

/* ------------------------------------------------------------------------------------------
 * Neighbours of the attack layer in the trainers' step (SURVEY 8f "next" rows).
 *
 * Post-attack epilogue (models/IRNp_model.py:674-680):
 *   out = Quantization( x + (clamp(sim, 0, 1) - x).detach() )      one pass instead of five;
 *   `out` may be a slice of the K-way batch, which removes the torch.cat.  Bit-identical values
 *   (same fp32 operation order).  The backward is the identity on x (straight-through), so
 *   the K slices of a bank reduce with wm_slice_sum: out[i] = sum_k g[k*n + i].
 * Tamper / splice (models/IRNcrop_model.py:348, models/IRNp_model.py:600):
 *   out = a * (1 - mask) + b * mask,  a, b: [B, C, H, W], mask: [B, 1, H, W] (any H*W; vector kernel when H*W % 4 == 0).
 *   bwd: ga = gy * (1 - mask), gb = gy * mask (either may be NULL).
 * ------------------------------------------------------------------------------------------ */
/* 8-bit frames -> [0,1] float32, dst[i] = src[i] / 255 (data format on the host side of the path:
 * upload bytes, convert on the device; the correctly rounded quotient, i.e. the values numpy / CPU torch
 * give for u8.astype(float32) / 255). */
int wm_u8_to_unit_float(const uint8_t* src, float* dst, int64_t n, void* stream);
/* [0,1] float32 -> 8-bit frames, dst[i] = rint(clamp(src[i], 0, 1) * 255) (round-half-even, torch.round): the
 * integer k of the Quantization layer's value k/255 (models/modules/Quantization.py:9), so that attacked
 * frames leave the device as bytes (4x less D2H traffic).  NaN -> 0. */
int wm_unit_float_to_u8(const float* src, uint8_t* dst, int64_t n, void* stream);
/* Shared-read bank ("read the input tile once, emit K attacked variants"): the 3x3-neighbourhood
 * members of a K-way attack bank computed from ONE staged tile of x, each finished by the epilogue above (with x = the
 * input itself) and written into its own slice.  12 + 12 K' B/px instead of 24 K' for K' members.  Every slice is
 * bit-identical to the member's own forward entry point called with the same store epilogue.
 *   x: N planes [H, W] (plane stride x_sp, row stride x_sh; 16-byte aligned, strides multiples of 4), W % 4 == 0;
 *   a NULL output pointer switches that member off; outputs are dense [N, H, W], 16-byte aligned.
 *   y_blur     GaussianBlur(kernel_size=3) (noise_layers/gaussian_blur.py:53-56), zero padding, taps blur_taps
 *   y_median   MiddleBlur(3) (noise_layers/middle_filter.py:11-13)
 *   y_noise    Gaussian (noise_layers/gaussian.py:10-17): clamp01(x + mean + std * N(0,1)) if noise_clamp, Philox
 *              (seed, offset) with the counter convention of wm_gaussnoise_fwd on the DENSE [N,H,W] element index
 *   y_identity Identity (noise_layers/identity.py): the epilogue of x itself                                      */
typedef struct wm_bank3_desc {
    float* y_blur; float blur_taps[3];
    float* y_median;
    float* y_noise; float noise_mean, noise_std; int noise_clamp; uint64_t seed, offset;
    float* y_identity;
    int clamp01, quantize;
} wm_bank3_desc;
int wm_bank3_ok(int N, int H, int W);
int wm_bank3_fwd(const float* x, int64_t x_sp, int64_t x_sh, int N, int H, int W,
                 const wm_bank3_desc* desc_host, void* stream);
int wm_attack_epilogue_fwd(const float* x, const float* sim, float* out, int64_t n, int clamp01, int quantize,
                           void* stream);
/* Hybrid attack (models/IRNcrop_model.py:357-373, the softmax mix as intended): the convex mix of K attacked versions of a
 * batch, followed by the trainer's clamp_with_grad and Quantization, in one pass:
 *     mixed = sum_k alpha[b, k] * t[k];   out = Quantization( mixed + (clamp(mixed, 0, 1) - mixed).detach() )
 * t[k]: K dense [B, chw] tensors (device pointers in the HOST struct), alpha: device [B, K]; products are added left to
 * right in fp32 (bit-identical to the torch expression).  wm_mix_bwd writes g[k] = alpha[b, k] * gy into t[k] (a NULL
 * member is skipped); clamp_with_grad and Quantization are straight-through. */
#define WM_MIX_MAX 8
typedef struct wm_mix_desc {
    float* t[WM_MIX_MAX];
    int K;
    int clamp01, quantize;
} wm_mix_desc;
int wm_mix_fwd(const wm_mix_desc* desc_host, const float* alpha, float* out, int64_t B, int64_t chw, void* stream);
int wm_mix_bwd(const float* gy, const float* alpha, const wm_mix_desc* grads_host, int64_t B, int64_t chw, void* stream);
int wm_slice_sum(const float* g, float* out, int64_t n, int K, void* stream);
int wm_splice_fwd(const float* a, const float* b, const float* mask, float* out, int64_t B, int C, int64_t hw,
                  void* stream);
int wm_splice_bwd(const float* gy, const float* mask, float* ga, float* gb, int64_t B, int C, int64_t hw,
                  void* stream);

/* ------------------------------------------------------------------------------------------
 * Crop + resize back, fast path.  Replaces Crop.forward, noise_layers/crop.py:48-53 (slice the crop
 * rectangle, F.interpolate to H x W, align_corners=False) for UP-scaling geometries: per axis
 * 0.45 <= in/out <= 1 (the layer's crop rates are 0.5 .. 1), Wout % 4 == 0, frame width % 4 == 0.
 *   x: [N, Hsrc, Wsrc] planes (element strides x_sp, x_sh; 16-byte aligned base and strides), the
 *   rectangle [h0, h0+Hin) x [w0, w0+Win) is read in place; y: dense [N, Hout, Wout].
 *   bwd: gx dense [N, Hsrc, Wsrc], written completely (zeros outside the rectangle), deterministic.
 *   tables: wm_cropresize_table_words(...) 4-byte words of scratch, 16-byte aligned (band tables built
 *   by a small kernel in the same call; the rectangle changes every call).
 * Other geometries: wm_interp_fwd / wm_interp_bwd.  mode 0 = bilinear (the layer), 1 = bicubic. */
int wm_cropresize_ok(int Hin, int Win, int Hout, int Wout, int N, int mode);
int64_t wm_cropresize_table_words(int Hin, int Win, int Hout, int Wout, int mode);
int wm_cropresize_fwd(const float* x, int64_t x_sp, int64_t x_sh, int Hsrc, int Wsrc, int h0, int w0, int Hin, int Win,
                      float* y, int N, int Hout, int Wout, int mode, int32_t* tables, void* stream);
int wm_cropresize_bwd(const float* gy, float* gx, int Hsrc, int Wsrc, int h0, int w0, int Hin, int Win,
                      int N, int Hout, int Wout, int mode, int32_t* tables, void* stream);

/* ------------------------------------------------------------------------------------------
 * Real-codec round trip.  Replaces JpegTest.forward, noise_layers/jpeg.py:21-45
 * (PIL save(format="JPEG", quality, subsampling) to a temp file + Image.open, frame by frame).
 * Entropy coding is lossless, so the decoded pixels are an integer function of the input bytes:
 * libjpeg's fixed-point colour conversion, box downsampling, "islow" 8x8 DCT pair, quantisation
 * with the Annex-K tables scaled by `quality` (jpeg_set_quality, baseline), "fancy" chroma
 * upsampling.  Output is bit-identical to Pillow/libjpeg-turbo's.  Not differentiable (as upstream).
 *   x: [B,3,H,W], element strides x_sb/x_sc/x_sh (W stride 1); y: dense [B,3,H,W]; any H, W >= 1.
 *   mode 0: float32 in [-1,1] both sides, with JpegTest's own conversions
 *           (u8 = trunc((clamp(x,-1,1)+1)/2*255);  out = (u8/255 - 0.5)/0.5);
 *   mode 1: float32 in [0,1] (u8 = rint(clamp(x,0,1)*255); out = u8/255);
 *   mode 2: uint8 both sides.
 *   subsampling: Pillow's numbering, 0 = 4:4:4, 1 = 4:2:2, 2 = 4:2:0.
 *   scratch: wm_jpegcodec_scratch_bytes(B,H,W,subsampling) bytes, 16-byte aligned (decoded Y/Cb/Cr
 *   planes between the two kernels).  coef: optional int16 buffer of scratch_bytes ELEMENTS that
 *   receives the quantised coefficients in the same plane layout (what the entropy coder would
 *   see), or NULL.
 * ------------------------------------------------------------------------------------------ */
int64_t wm_jpegcodec_scratch_bytes(int B, int H, int W, int subsampling);
int wm_jpegcodec(const void* x, int64_t x_sb, int64_t x_sc, int64_t x_sh, void* y, int B, int H, int W,
                 int quality, int subsampling, int mode, uint8_t* scratch, int16_t* coef, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WM_ATTACK_H */
