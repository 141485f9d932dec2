/*
 * sc_b200.h -- C ABI of the B200-native spectral-cube hot path (libsc_b200.so).
 *
 * The reference (radio-astro-tools/spectral-cube @ cb6969e) is pure Python and has no
 * FFI of its own; the seams this library sits under are Python-level.
 * Each entry point below names the reference code it replaces (file:line under
 * /root/reference/spectral_cube/).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *   - Cubes are float32, numpy axis order (nchan, ny, nx), x contiguous (element
 *     stride 1); `stride_c` / `stride_y` are ELEMENT strides so views work
 *     (spectral_cube.py:1378, :1871 hand out views).
 *   - Bulk pointers (cubes, planes, masks, outputs) are DEVICE pointers owned by the
 *     caller.  Small parameter arrays (taps, per-channel coordinates, the mask
 *     descriptor) are HOST pointers; they are copied before the call returns.
 *     Entry points with the suffix `_host` take HOST bulk buffers and run the
 *     host<->device pipeline themselves.
 *   - Nothing is allocated or freed on the device by compute calls except by the
 *     `_host` pipelines (which own their staging buffers for the duration of the
 *     call).  Scratch comes from `workspace` (device, >= sc_workspace_bytes()).
 *   - Every call enqueues on `stream` (a cudaStream_t passed as void*) and returns
 *     without synchronising, except `_host` calls, which return when the result is
 *     in the caller's host buffers.
 *   - Return value: 0 = ok, negative = error; sc_last_error() gives the message
 *     (thread local).  Arguments are validated before anything is launched.
 *   - There is no CPU fallback anywhere in this library.
 */
#ifndef SC_B200_H
#define SC_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC_OK              0
#define SC_ERR_ARG        -1   /* null pointer, bad shape/stride, even kernel size ... */
#define SC_ERR_CUDA       -2   /* a CUDA runtime call or launch failed */
#define SC_ERR_UNSUPPORTED -3
#define SC_ERR_WORKSPACE  -4   /* workspace too small */

/* ---- mask descriptor ---------------------------------------------------------------
 * Device-side restatement of the reference mask tree (masks.py:101-803): the include
 * predicate of a voxel is evaluated in registers while the cube streams through the
 * kernel, so lazy masks cost no HBM traffic.  Nodes are stored children-first; node
 * `n_nodes-1` is the root.  n_nodes == 0 means "no mask" (mask is None,
 * base_class.py:408-409).
 *
 * Data-referencing nodes (FINITE, CMP_*) look at `data` -- the array the LazyMask was
 * BUILT on (masks.py:649-651), which need not be the cube being reduced
 * (spectral_cube.py:3043-3045 keeps the old mask after smoothing).  data == NULL means
 * "the cube passed to the call".  `s_c, s_y, s_x` are element strides of `data` /
 * `array` in the cube's index space; 0 on a broadcast axis (masks.py:520, :724-733).
 * Comparisons are done as (double)x OP value, as numpy does with the np.float64
 * threshold the reference produces (spectral_cube.py:2248-2252).
 */
enum sc_mask_kind {
    SC_MASK_FINITE     = 1,   /* LazyMask(np.isfinite)            io/fits.py:214       */
    SC_MASK_CMP_SCALAR = 2,   /* LazyComparisonMask, scalar       masks.py:731-733     */
    SC_MASK_CMP_ARRAY  = 3,   /* LazyComparisonMask, broadcast    masks.py:724-729     */
    SC_MASK_BOOL       = 4,   /* BooleanArrayMask (uint8 0/1)     masks.py:555-557     */
    SC_MASK_AND        = 5,   /* CompositeMask                    masks.py:425-435     */
    SC_MASK_OR         = 6,
    SC_MASK_XOR        = 7,
    SC_MASK_NOT        = 8    /* InvertedMask                     masks.py:346-347     */
};
enum sc_cmp_op { SC_GT = 0, SC_GE = 1, SC_LT = 2, SC_LE = 3, SC_EQ = 4, SC_NE = 5 };
enum sc_dtype  { SC_F32 = 0, SC_F64 = 1, SC_U8 = 2 };

#define SC_MASK_MAX_NODES 16

typedef struct sc_mask_node {
    int32_t     kind;          /* sc_mask_kind */
    int32_t     op;            /* sc_cmp_op for CMP_* */
    int32_t     a, b;          /* child node indices for AND/OR/XOR/NOT (b unused by NOT) */
    int32_t     array_dtype;   /* sc_dtype of `array` (CMP_ARRAY: F32/F64, BOOL: U8) */
    int32_t     reserved;
    double      value;         /* CMP_SCALAR threshold */
    const float *data;         /* FINITE / CMP_*: float32 data the mask was built on; NULL = the cube */
    int64_t     ds_c, ds_y;    /* element strides of `data` (x stride is 1); ignored if data == NULL */
    const void  *array;        /* CMP_ARRAY comparison array / BOOL mask array */
    int64_t     as_c, as_y, as_x;  /* element strides of `array`; 0 = broadcast axis */
} sc_mask_node;

typedef struct sc_mask_desc {
    int32_t      n_nodes;
    int32_t      reserved;
    sc_mask_node nodes[SC_MASK_MAX_NODES];
} sc_mask_desc;

/* ---- library ------------------------------------------------------------------------ */
const char *sc_last_error(void);
int         sc_version(void);                 /* ABI version, currently 1 */
/* Scratch a call may need, in bytes; op is one of the SC_OP_* below. */
enum sc_op { SC_OP_MOMENTS = 1, SC_OP_SPECTRAL_SMOOTH = 2, SC_OP_SPATIAL_SMOOTH = 3,
             SC_OP_SPECTRAL_INTERP = 4, SC_OP_REPROJECT = 5, SC_OP_SMOOTH_MOMENTS = 6, SC_OP_REDUCE = 7, SC_OP_INGEST = 8, SC_OP_POINTWISE = 9 };
size_t      sc_workspace_bytes(int op, int64_t nchan, int64_t ny, int64_t nx, int64_t aux);
/* Number of kernel launches this library has enqueued in this process (all streams). */
int64_t     sc_launch_count(void);

/* ---- mask evaluation / filled data ---------------------------------------------------
 * sc_mask_include: materialise `mask.include()` (masks.py:143-158) as uint8 (nchan,ny,nx).
 * sc_fill_masked:  `_get_filled_data(fill=...)` (base_class.py:389-417 -> masks.py:197-237):
 *                  out = include ? cube : fill, float32 contiguous (nchan,ny,nx). */
int sc_mask_include(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                    int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                    uint8_t *out, void *stream);
int sc_fill_masked(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                   int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask, double fill,
                   float *out, void *stream);

/* ---- moments along the spectral axis ------------------------------------------------
 * Replaces the strategy table `dispatch[how](cube, order, axis)` for axis 0
 * (spectral_cube.py:1682-1691 -> _moments.py:30-202; dask_spectral_cube.py:1083-1101):
 *   M0 = sum_c I dx ; M1 = sum_c I x dx / sum_c I dx ; M2 = sum_c I (x-M1)^2 dx / sum_c I dx
 * over voxels that the mask includes AND that are not NaN (nansum, _moments.py:176-193);
 * a ray with no such voxel gives NaN for every order (np_compat.py:20-24).
 * One pass over the cube computes any subset of {M0, M1, M2} (want_bits: 1|2|4) with
 * float64 accumulators.  `chan_offset[nchan]` = spectral offset of each channel from
 * channel 0 (`_pix_cen()[0]`, spectral_cube.py:1473-1475), `pix_size` =
 * `_pix_size_slice(0)` (:1526-1529), `m1_offset` = the world coordinate of channel 0
 * that `moment()` adds for order 1 (:1709-1710).  Outputs are (ny, nx) float64,
 * contiguous; unused outputs may be NULL.
 */
#define SC_WANT_M0 1
#define SC_WANT_M1 2
#define SC_WANT_M2 4
int sc_moments_axis0(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                     int64_t stride_c, int64_t stride_y,
                     const sc_mask_desc *mask,
                     const double *chan_offset, double pix_size, double m1_offset,
                     int want_bits, double *out_m0, double *out_m1, double *out_m2,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Higher orders (order >= 2 about a given centre plane): the second pass of
 * _moments.py:108-123 / :192-193.  centre is the (ny, nx) float64 offset-from-channel-0
 * first moment (i.e. M1 - m1_offset). */
int sc_moment_central_axis0(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                            int64_t stride_c, int64_t stride_y,
                            const sc_mask_desc *mask,
                            const double *chan_offset, const double *centre, int order,
                            double *out, void *workspace, size_t workspace_bytes, void *stream);

/* Moments along a spatial axis (axis = 1 or 2; _moments.py same functions with
 * pix_cen[axis], spectral_cube.py:1481-1492).  `offsets` is the (ny, nx) float64 plane of
 * cumulative angular offsets (DEVICE pointer); output is (nchan, nx) for axis 1 and
 * (nchan, ny) for axis 2, float64.  want_bits as above (M1 gets no world shift). */
int sc_moments_spatial(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                       int64_t stride_c, int64_t stride_y, int axis,
                       const sc_mask_desc *mask,
                       const double *offsets, double pix_size,
                       int want_bits, double *out_m0, double *out_m1, double *out_m2,
                       void *stream);

/* Cumulative great-circle pixel offsets from the cube face along numpy axis 1 (y) or 2 (x)
 * (`_pix_cen()[1]`, `[2]`, spectral_cube.py:1477-1492): float64 (ny, nx) DEVICE plane, in degrees.
 * wcs = the 12 doubles of sc_wcs_pixel_map; workspace >= ny*nx*16 + 256 bytes. */
int sc_pixel_offsets(const double *wcs, int64_t ny, int64_t nx, int axis, double *offsets,
                     void *workspace, size_t workspace_bytes, void *stream);

/* Host-buffer pipeline for the moment maps: `cube_host` is a HOST float32 array (pinned
 * or pageable) with the strides given; row blocks are streamed through two device
 * staging buffers (copy/compute overlap) and the three maps land in HOST float64
 * buffers.  Only self-referencing masks (data == NULL, no arrays) are accepted.
 * `staging_bytes` = device memory the call may allocate for staging (0 = default). */
int sc_moments_axis0_host(const float *cube_host, int64_t nchan, int64_t ny, int64_t nx,
                          int64_t stride_c, int64_t stride_y,
                          const sc_mask_desc *mask,
                          const double *chan_offset, double pix_size, double m1_offset,
                          int want_bits, double *out_m0_host, double *out_m1_host,
                          double *out_m2_host, size_t staging_bytes, int device);

/* ---- spectral smoothing ----------------------------------------------------------------
 * Replaces `convolve(spectrum, kernel, normalize_kernel=True)` applied to every spaxel
 * (spectral_cube.py:3186-3222 via :3103-3159 and :147-158; dask_spectral_cube.py:880-917
 * with an (n,1,1) kernel): true convolution, zero-filled boundary whose zeros count as
 * valid samples, NaN-interpolating (`top/bot`, bot == 0 keeps the input NaN), float64
 * arithmetic.  Masked voxels are replaced by `fill` first (`unitless_filled_data`,
 * base_class.py:439-450).  `taps` are the n (odd) float64 kernel values as the user's
 * Kernel1D holds them; normalisation by their sum is part of the call.
 * out_dtype SC_F64 mirrors the numpy class (:2953/:2963), SC_F32 the dask class (:829).
 * `spaxel_passthrough` != 0 reproduces `_apply_spectral_function` (:147-158): a spaxel
 * with no included voxel is copied through unchanged (only differs from the
 * convolution when `fill` is finite).  in == out (in place) is allowed for SC_F32.
 */
int sc_spectral_smooth(const float *in, void *out, int out_dtype,
                       int64_t nchan, int64_t ny, int64_t nx,
                       int64_t stride_c, int64_t stride_y,
                       int64_t out_stride_c, int64_t out_stride_y,
                       const sc_mask_desc *mask, double fill,
                       const double *taps, int ntaps, int spaxel_passthrough,
                       void *workspace, size_t workspace_bytes, void *stream);

/* Fused `spectral_smooth(kernel)` -> `moment(order)` (BASELINE config 3): the smoothed
 * cube is never written.  The smoothed value is rounded to `smooth_dtype` exactly where
 * the reference materialises it, and the moment uses the include mask of the ORIGINAL
 * data (the smoothed cube keeps the old mask, spectral_cube.py:3043-3045). */
int sc_smooth_moments_axis0(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                            int64_t stride_c, int64_t stride_y,
                            const sc_mask_desc *mask, double fill,
                            const double *taps, int ntaps, int smooth_dtype,
                            const double *chan_offset, double pix_size, double m1_offset,
                            int want_bits, double *out_m0, double *out_m1, double *out_m2,
                            void *workspace, size_t workspace_bytes, void *stream);

/* ---- spatial smoothing -----------------------------------------------------------------
 * Replaces `convolve(image, kernel2d, normalize_kernel=True)` applied to every channel
 * (spectral_cube.py:2808-2842 via :3049-3101 and :161-172; dask_spectral_cube.py:962-993).
 * Same convolution semantics as above, in 2-D.  The separable entry takes the two 1-D
 * factors of an outer-product kernel (Gaussian2DKernel); the 2-D entry takes any odd x odd
 * kernel (Tophat2DKernel, rotated elliptical beams).
 *
 * Row sharding: the caller owns rows [y0, y0+ny) of a taller image.
 * `halo_top` / `halo_bot` are float32 (nchan, halo_rows, nx) contiguous device buffers
 * holding the neighbouring ranks' FILLED edge rows (NULL = image boundary: zero fill).
 * `plane_passthrough` != 0 reproduces `_apply_spatial_function` (:161-172).
 */
int sc_spatial_smooth_sep(const float *in, void *out, int out_dtype,
                          int64_t nchan, int64_t ny, int64_t nx,
                          int64_t stride_c, int64_t stride_y,
                          int64_t out_stride_c, int64_t out_stride_y,
                          const sc_mask_desc *mask, double fill,
                          const double *taps_y, int ntaps_y, const double *taps_x, int ntaps_x,
                          const float *halo_top, const float *halo_bot, int halo_rows,
                          int plane_passthrough,
                          void *workspace, size_t workspace_bytes, void *stream);

/* Denominator strategy of the separable path.  `sc_spatial_smooth_sep` samples the cube itself (every
 * 8-row x 128-column block of up to 16 planes) and lets the device pick between the integer deficit (pipe
 * kernel) and the convolved float32 denominator (more than a third of the blocks "crowded": over 1/64 of a
 * block's samples missing, but not all).  A row-sharded job must take ONE decision for all shards to stay
 * bit-identical with the unsharded result: each rank calls `sc_spatial_missing_sample` (counts = device
 * uint32[2] {crowded blocks, blocks sampled}, accumulated, not zeroed),
 * sums the counts over the ranks and hands them to `sc_spatial_smooth_sep_ex`. */
int sc_spatial_missing_sample(const float *in, int64_t nchan, int64_t ny, int64_t nx,
                              int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                              unsigned int *counts, void *stream);
int sc_spatial_smooth_sep_ex(const float *in, void *out, int out_dtype,
                             int64_t nchan, int64_t ny, int64_t nx,
                             int64_t stride_c, int64_t stride_y,
                             int64_t out_stride_c, int64_t out_stride_y,
                             const sc_mask_desc *mask, double fill,
                             const double *taps_y, int ntaps_y, const double *taps_x, int ntaps_x,
                             const float *halo_top, const float *halo_bot, int halo_rows,
                             int plane_passthrough, const unsigned int *strategy_counts,
                             void *workspace, size_t workspace_bytes, void *stream);

int sc_spatial_smooth_2d(const float *in, void *out, int out_dtype,
                         int64_t nchan, int64_t ny, int64_t nx,
                         int64_t stride_c, int64_t stride_y,
                         int64_t out_stride_c, int64_t out_stride_y,
                         const sc_mask_desc *mask, double fill,
                         const double *taps, int ntaps_y, int ntaps_x,
                         const float *halo_top, const float *halo_bot, int halo_rows,
                         int plane_passthrough,
                         void *workspace, size_t workspace_bytes, void *stream);

/* In-place epilogue of `convolve_to` over a (nchan, ny, nx) float32/float64 block (x stride 1), one pass:
 * `data *= factor` -- the Jy/beam rescale by target.sr / beam.sr of each convolved channel image
 * (spectral_cube.py:3369-3378; VaryingResolutionSpectralCube :4207-4233); with `nan_to_zero`, NaN -> 0.0 --
 * what `astropy.convolution.convolve_fft` (the numpy class's default there) returns for output pixels whose
 * kernel window holds no valid input, where the direct `convolve` keeps the NaN; otherwise NaN stays NaN.
 * `skip_planes` (device uint8[nchan] or NULL): channels with a non-zero flag are left untouched (planes that
 * `_apply_spatial_function` copies through, spectral_cube.py:161-172). */
int sc_scale(void *data, int dtype, int64_t nchan, int64_t ny, int64_t nx,
             int64_t stride_c, int64_t stride_y, double factor, int nan_to_zero,
             const uint8_t *skip_planes, void *stream);

/* mosaic_cubes (cube_utils.py:791-856), the elementwise steps around the per-cube `reproject` calls, contiguous
 * (nchan, ny, nx) buffers.  sc_mosaic_accumulate: acc += nan_to_num(reprojected) (:838-841) and, when `weight` is given,
 * weight += footprint0 -- the reprojected cube's mask in channel 0, the 2-D coverage count of :834-836.
 * sc_mosaic_normalize: acc[c] /= weight for every channel (:847-849), IEEE division like numpy's. */
/* Rows -> channels re-shard over NVLink peer memory (no counterpart in the reference, whose parallelism is
 * one host): this rank's block `local` (nchan, rows, nx) -- rows [y0, y0 + rows) of the image -- is stored channel by
 * channel into the destination ranks' (chan_bounds[d+1] - chan_bounds[d], ny_total, nx) buffers, `peer_ptrs[d]` being rank d's
 * buffer as mapped into THIS process (device pointers in a host array of `world` entries, e.g. the `buffer_ptrs` of a
 * torch symmetric-memory handle).  One kernel, one read of the local block, peer stores of 16-byte vectors, destinations
 * interleaved line by line starting at `rank`'s neighbour so that no peer's ingress is a hot spot; the caller
 * orders it against the other ranks (barrier before: the buffers are free; barrier after: all stores have landed). */
int sc_reshard_scatter(const float *local, int64_t nchan, int64_t rows, int64_t nx,
                       int64_t stride_c, int64_t stride_y,
                       const uint64_t *peer_ptrs, int world, int rank, const int64_t *chan_bounds,
                       int64_t ny_total, int64_t y0, void *stream);

int sc_mosaic_accumulate(double *acc, double *weight, const void *reprojected, int dtype, const uint8_t *footprint0,
                         int64_t nchan, int64_t ny, int64_t nx, void *stream);
int sc_mosaic_normalize(double *acc, const double *weight, int64_t nchan, int64_t ny, int64_t nx, void *stream);

/* Write the FILLED edge rows a neighbour needs: rows [row0, row0+nrows) of every channel,
 * mask applied (excluded -> fill), into a contiguous (nchan, nrows, nx) float32 buffer. */
int sc_pack_filled_rows(const float *in, int64_t nchan, int64_t ny, int64_t nx,
                        int64_t stride_c, int64_t stride_y,
                        const sc_mask_desc *mask, double fill,
                        int64_t row0, int64_t nrows, float *out, void *stream);

/* ---- spectral interpolation ------------------------------------------------------------
 * Replaces the per-spaxel `np.interp(grid, inaxis, spectrum, left, right)` loop
 * (spectral_cube.py:3298-3315) and the per-block `interp1d` call
 * (dask_spectral_cube.py:1342-1349).  The host precomputes, per output channel, the
 * bracketing input channel `idx`, `xlo = inaxis[idx]`, `xhi = inaxis[idx+1]` and the
 * classification `kind`: 0 interior, 1 exact knot hit on idx (value = data[idx]),
 * 2 left of the axis, 3 right of the axis.  mode 0 = numpy class (clamp / fill_value,
 * new mask = interp(mask) > 0, rays with nothing included -> NaN/False, :3305-3313),
 * mode 1 = dask class (NaN outside unless fill_value, mask = ~isnan, :1364).
 * Inputs are addressed through `in_reversed` (data/mask flipped first, :3267-3268) and
 * outputs through `out_reversed` (:3259-3264).  out is (nchan_out, ny, nx) of out_dtype;
 * out_mask (uint8, may be NULL) the new boolean mask.
 */
int sc_spectral_interp(const float *in, void *out, int out_dtype, uint8_t *out_mask,
                       int64_t nchan, int64_t ny, int64_t nx,
                       int64_t stride_c, int64_t stride_y,
                       int64_t nchan_out,
                       const sc_mask_desc *mask, double fill,
                       const double *in_axis, const double *grid,
                       int has_fill_value, double fill_value,
                       int in_reversed, int out_reversed, int mode,
                       void *workspace, size_t workspace_bytes, void *stream);

/* The same interpolation with a SCATTERED output (the step between `spectral_interpolate` and `reproject`
 * of a row-sharded cube, which needs whole planes): output channel j of these rows is stored at
 * `(T *)out_chan_ptrs[j] + y * nx + x`.  `out_chan_ptrs` is a DEVICE array of nchan_out addresses, each 16-byte aligned
 * and pointing at row `y0` of channel j in the buffer of the rank that owns that channel -- this rank's own buffer or a
 * peer's mapped over NVLink (torch symmetric memory `buffer_ptrs`).  The interpolation kernel's stores ARE the exchange:
 * no row-sharded result is written and read again, no all-to-all follows.  The caller brackets the call with a barrier
 * among the ranks on both sides.  `out_mask` (may be NULL) stays local, (nchan_out, ny, nx).
 * `phase` / `nphases` (rank / world, nphases <= 16): CTA b starts its march over the spectrum at the fraction
 * ((phase + b) % nphases) / nphases of it and wraps around, so that at any moment the ranks of a job store to all channel
 * owners in equal shares instead of all to the same one; 0 / 1 = plain order. */
int sc_spectral_interp_scatter(const float *in, const uint64_t *out_chan_ptrs, int out_dtype, uint8_t *out_mask,
                               int64_t nchan, int64_t ny, int64_t nx,
                               int64_t stride_c, int64_t stride_y,
                               int64_t nchan_out,
                               const sc_mask_desc *mask, double fill,
                               const double *in_axis, const double *grid,
                               int has_fill_value, double fill_value,
                               int in_reversed, int out_reversed, int mode,
                               int phase, int nphases,
                               void *workspace, size_t workspace_bytes, void *stream);

/* ---- reprojection ----------------------------------------------------------------------
 * Replaces `reproject_interp((data, header), wcs_out, shape_out=..., order='bilinear')`
 * (spectral_cube.py:2726-2732): per output pixel the input pixel position (float64 planes
 * `yin`, `xin`, DEVICE, (ny_out, nx_out)) is sampled bilinearly in float64 on every
 * channel; NaN beyond half a pixel outside the input image; NaN neighbours poison the
 * sample.  footprint (uint8, may be NULL) = ~isnan(out).  order 0 = nearest-neighbour,
 * 1 = bilinear.
 */
int sc_reproject(const float *in, void *out, int out_dtype, uint8_t *footprint,
                 int64_t nchan, int64_t ny_in, int64_t nx_in,
                 int64_t stride_c, int64_t stride_y,
                 int64_t ny_out, int64_t nx_out,
                 const sc_mask_desc *mask, double fill,
                 const double *yin, const double *xin, int order, void *stream);

/* The same in one pass with the by-products the caller of `reproject_interp` derives afterwards
 * (spectral_cube.py:2733-2746): `out_f32` (may be NULL) receives the result rounded to float32
 * next to the `out` of dtype `out_dtype` (the new cube's float32 working copy), and
 * `*any_valid` (device int32, may be NULL; zeroed by the call) is set to 1 when at least one output
 * voxel is not NaN -- the "All values in reprojected cube are nan" test without another pass. */
int sc_reproject_ex(const float *in, void *out, int out_dtype, float *out_f32, uint8_t *footprint, int *any_valid,
                    int64_t nchan, int64_t ny_in, int64_t nx_in,
                    int64_t stride_c, int64_t stride_y,
                    int64_t ny_out, int64_t nx_out,
                    const sc_mask_desc *mask, double fill,
                    const double *yin, const double *xin, int order, void *stream);

/* ---- reductions along the spectral axis ------------------------------
 * Replaces `apply_numpy_function(np.nansum / nanmean / nanstd / nanmax / nanmin / nanargmax /
 * nanargmin, fill=..., axis=0)`: spectral_cube.py:361-470 (driver), `sum` :578-588, `mean` :592-652,
 * `std` :669-724, `max` :770-781, `min` :785-796, `argmax` :800-811, `argmin` :815-826;
 * dask_spectral_cube.py:641-767.  A voxel takes part iff the mask includes it and it is not NaN.
 * One read of the cube serves every requested output (each (ny, nx), any may be NULL):
 *   out_sum (f64; NaN when nothing takes part, np_compat.py:20-24), out_count (i32),
 *   out_m2 (f64: sum of squared deviations from the mean; std = sqrt(m2 / (count - ddof))),
 *   out_min / out_max (f32; NaN when nothing takes part), out_argmin / out_argmax (i32 channel of
 *   the FIRST extremum like numpy; 0 when nothing takes part -- "arbitrary" in the reference).
 */
int sc_reduce_axis0(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                    int64_t stride_c, int64_t stride_y, const sc_mask_desc *mask,
                    double *out_sum, int32_t *out_count, double *out_m2,
                    float *out_min, float *out_max, int32_t *out_argmin, int32_t *out_argmax,
                    void *stream);

/* The same seven statistics along a SPATIAL axis (numpy axis 1 = y -> outputs (nchan, nx); axis 2 = x -> outputs
 * (nchan, ny)); indices are positions along the reduced axis.  `apply_numpy_function(..., axis=1 or 2)`,
 * spectral_cube.py:361-470, 578-826. */
int sc_reduce_spatial(const float *cube, int64_t nchan, int64_t ny, int64_t nx,
                      int64_t stride_c, int64_t stride_y, int axis, const sc_mask_desc *mask,
                      double *out_sum, int32_t *out_count, double *out_m2,
                      float *out_min, float *out_max, int32_t *out_argmin, int32_t *out_argmax,
                      void *stream);

/* ---- FITS ingest -------------------------------------------------------
 * Decodes `n` big-endian FITS samples already on the device (the raw bytes of the data block as
 * astropy.io.fits maps them, io/fits.py:100-160 `read_data_fits`) to native float32:
 * out = BZERO + BSCALE * sample, integer samples equal to BLANK become NaN (FITS standard 4.4.2.5).
 * bitpix is -32, -64, 16, 32 or 8.  The host side streams the file through pinned buffers
 * (spectral_cube_b200/io_fits.py) and calls this per block. */
int sc_fits_decode(const void *raw_be, float *out, int64_t n, int bitpix,
                   double bscale, double bzero, int has_blank, long long blank, void *stream);

/* Celestial pixel->pixel map for two TAN/SIN WCSs (the part of `reproject_interp` that
 * runs through astropy.wcs; FITS WCS papers I/II).  wcs_* = 12 doubles:
 * crpix1, crpix2, crval1, crval2, cd11, cd12, cd21, cd22, lonpole, proj(0 TAN,1 SIN), 0, 0.
 * Writes float64 (ny_out, nx_out) planes yin, xin (0-based input pixel coordinates). */
int sc_wcs_pixel_map(const double *wcs_out, const double *wcs_in,
                     int64_t ny_out, int64_t nx_out, double *yin, double *xin, void *stream);

/* ---- synthetic cubes (bench / test inputs) -------------------------------
 * cube[c,y,x] = A(y,x) * profile[|16 (c - c0(y,x))|] + noise, A ~ U(0,10),
 * c0 ~ U(nchan/4, 3 nchan/4) on a 1/16-channel lattice, noise = Irwin-Hall(8) scaled to
 * unit variance, all from a counter-based integer hash of (seed, global voxel index), so
 * any sub-block can be regenerated bit-identically on the CPU (oracle/synth.py).
 * `profile` is a DEVICE float32 table of 16*nchan_total entries.  nan_permille random
 * voxels per thousand and a `border` pixel frame are NaN.  (y0, x0, ny_total, nx_total)
 * place this block inside the full plane so shards generate their own rows.
 */
int sc_synth_cube(float *cube, int64_t nchan, int64_t ny, int64_t nx,
                  int64_t y0, int64_t x0, int64_t ny_total, int64_t nx_total,
                  uint64_t seed, const float *profile, int nan_permille, int border,
                  void *stream);

/* Utility: per-kernel timing hooks used by bench.py -- elapsed ms between two events the
 * library records on `stream` around the LAST launch of each op (0 if none). */
float sc_last_kernel_ms(int op);
void  sc_enable_kernel_timing(int enable);

#ifdef __cplusplus
}
#endif
#endif /* SC_B200_H */
