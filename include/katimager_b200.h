/* katimager_b200.h -- C ABI of libkatimager_b200.so
 *
 * B200 (sm_100a) implementation of katsdpimager's imaging hot path:
 * W-projection gridding/degridding, grid<->image FFT stage with fused
 * taper/W-term/fftshift kernels, Hogbom CLEAN minor cycles, plus the small
 * weighting / prediction / image-arithmetic kernels that
 * katsdpimager.imaging.ImagingTemplate constructs unconditionally.
 *
 * The reference (ska-sa/katsdpimager) has no C ABI for this path: its kernels
 * are Mako templates JIT-compiled through katsdpsigproc/PyCUDA.  Each entry
 * point below therefore cites the reference *operation* (Python `_run` that
 * enqueues the kernel + the .mako kernel) that it replaces.  Paths are
 * relative to the reference root.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; the message for
 *    the last failure on the calling thread is returned by kib_last_error().
 *  - all kernels are enqueued asynchronously on the caller-supplied stream
 *    (an opaque handle from kib_stream_create; NULL = default stream).
 *  - the caller owns every buffer; device pointers are plain `void *`.
 *  - strides are in ELEMENTS of the array's dtype, not bytes.
 *  - `dtype` arguments select the precision of grid/image data:
 *    KIB_F32 (float / float2) or KIB_F64 (double / double2).  Visibilities,
 *    convolution kernel LUTs and weights are always single precision, as in
 *    the reference (grid.py:661-668, 426-447).
 */
#ifndef KATIMAGER_B200_H
#define KATIMAGER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KIB_VERSION 1

#define KIB_F32 0
#define KIB_F64 1

#define KIB_FFT_FORWARD 0
#define KIB_FFT_INVERSE 1

#define KIB_CLEAN_I 0      /* clean.py:29 */
#define KIB_CLEAN_SUMSQ 1  /* clean.py:31 */

typedef void *kib_stream_t;
typedef void *kib_event_t;
typedef void *kib_fft_plan_t;

/* ------------------------------------------------------------------ runtime
 * Replaces the katsdpsigproc.accel / PyCUDA layer used by every operation
 * (context, command queue, DeviceArray, HostArray: SURVEY.md section 8b.2). */
int kib_version(void);
const char *kib_last_error(void);
int kib_device_count(int *count);
int kib_set_device(int device);
int kib_get_device(int *device);
int kib_device_name(int device, char *buf, int buf_len);
/* attr: 0 = SM count, 1 = max shared memory per block (opt-in), 2 = L2 bytes,
 * 3 = SM clock kHz, 4 = compute capability major*10+minor, 5 = warp size */
int kib_device_attr(int device, int attr, int64_t *value);
int kib_mem_info(size_t *free_bytes, size_t *total_bytes);

int kib_stream_create(kib_stream_t *stream);
int kib_stream_destroy(kib_stream_t stream);
int kib_stream_sync(kib_stream_t stream);
int kib_stream_wait_event(kib_stream_t stream, kib_event_t event);

int kib_event_create(kib_event_t *event);
int kib_event_record(kib_event_t event, kib_stream_t stream);
int kib_event_sync(kib_event_t event);
int kib_event_query(kib_event_t event, int *done);
int kib_event_elapsed_ms(kib_event_t start, kib_event_t stop, float *ms);
int kib_event_destroy(kib_event_t event);

int kib_malloc(void **ptr, size_t bytes);
int kib_free(void *ptr);
int kib_host_alloc(void **ptr, size_t bytes);   /* pinned host memory */
int kib_host_free(void *ptr);
/* Page-lock an existing host range (e.g. the mapping of an output file) so that device copies
 * can target it directly; kib_host_unregister before unmapping it. */
int kib_host_register(void *ptr, size_t bytes);
int kib_host_unregister(void *ptr);
int kib_memset_async(void *ptr, int value, size_t bytes, kib_stream_t stream);
int kib_memcpy_h2d_async(void *dst, const void *src, size_t bytes, kib_stream_t stream);
int kib_memcpy_d2h_async(void *dst, const void *src, size_t bytes, kib_stream_t stream);
int kib_memcpy_d2d_async(void *dst, const void *src, size_t bytes, kib_stream_t stream);
/* Strided 3-D copy (DeviceArray.set_region/get_region/copy_region).
 * kind: 0 = host->device, 1 = device->host, 2 = device->device.
 * A "plane" is `height` rows of `width_bytes`; pitches are in bytes. */
int kib_memcpy3d_async(void *dst, size_t dst_row_pitch, size_t dst_plane_pitch,
                       const void *src, size_t src_row_pitch, size_t src_plane_pitch,
                       size_t width_bytes, size_t height, size_t depth,
                       int kind, kib_stream_t stream);

/* ---------------------------------------------------------------------- FFT
 * katsdpsigproc.fft.FftTemplate as used by image.py:585-600 (cuFFT C2C 2-D,
 * unnormalised, in place allowed). ny x nx complex elements, contiguous rows
 * of `row_stride` elements (row_stride == nx unless padded). */
int kib_fft_plan2d_create(kib_fft_plan_t *plan, int ny, int nx, int row_stride, int dtype);
/* Real <-> half-complex 2-D plans (katsdpsigproc.fft.FftTemplate with a real source or
 * destination, beam.py:327-330): ny x nx real rows of real_row_stride elements, ny x (nx/2+1)
 * complex rows of complex_row_stride elements; inverse = 0 real -> complex, 1 complex -> real
 * (unnormalised; the input of the inverse is overwritten).  Executed with
 * kib_fft_plan2d_exec (direction ignored). */
int kib_fft_plan2d_real_create(kib_fft_plan_t *plan, int ny, int nx, int real_row_stride,
                               int complex_row_stride, int inverse, int dtype);
int kib_fft_plan2d_exec(kib_fft_plan_t plan, void *src, void *dst, int direction,
                        kib_stream_t stream);
int kib_fft_plan2d_destroy(kib_fft_plan_t plan);
/* Batched 1-D C2C plan: `batch` transforms of length n; consecutive elements of one
 * transform are `stride` elements apart, consecutive transforms `dist` elements apart
 * (same layout for input and output).  Executed/destroyed with kib_fft_plan2d_exec /
 * kib_fft_plan2d_destroy.  Used to build the zero-row-skipping 2-D transform of the
 * grid -> image stage (the grid occupies only G of the N layer rows). */
int kib_fft_plan1d_create(kib_fft_plan_t *plan, int n, int64_t stride, int64_t dist,
                          int batch, int dtype);

/* ------------------------------------------------------------ grid / degrid
 * kib_grid replaces Gridder._run / static_run (grid.py:787-867) + grid.mako:63.
 * For each visibility i < num_vis:
 *   sample[p] = vis[i][p] * weights_grid[p][uv.y + G/2][uv.x + G/2]
 *   grid[p][v0 + j][u0 + k] += sample[p] * conj(lut[w][sub_v][j] * lut[w][sub_u][k])
 * with u0 = uv.x - ((K-1)/2 - G/2), v0 likewise, j,k < K  (oracle: grid.py:1033-1052).
 *   uv        int16[num_vis][4] = (u, v, sub_u, sub_v)       (grid.py:661-664)
 *   w_plane   int16[num_vis]
 *   vis       float2[num_vis][num_pols], pre-multiplied by statistical weights
 *   lut       float2[w_planes][oversample][lut_slice_stride]; taps start at
 *             lut_tap_offset within a slice (the reference pads the slice,
 *             grid.py:429-447; pass 0 for an unpadded LUT)
 *   grid      complex (dtype) [num_pols][grid_size][grid_size] (+strides)
 * Visibilities whose footprint would fall outside the grid are skipped and
 * counted in *num_rejected (device int32, may be NULL).
 */
int kib_grid(void *grid, int grid_row_stride, int64_t grid_pol_stride, int grid_size, int dtype,
             const float *weights_grid, int weights_row_stride, int64_t weights_pol_stride,
             const int16_t *uv, const int16_t *w_plane, const void *vis,
             const void *lut, int lut_slice_stride, int lut_tap_offset,
             int w_planes, int oversample, int kernel_width, int num_pols,
             int64_t num_vis, int32_t *num_rejected, kib_stream_t stream);

/* kib_degrid replaces Degridder._run (grid.py:986-1029) + degrid.mako:77:
 *   vis[i][p] -= weights[i][p] * sum_{j,k} lut[w][sub_v][j]*lut[w][sub_u][k]*grid[p][v0+j][u0+k]
 * (oracle: grid.py:1139-1154).  weights float[num_vis][num_pols]. */
int kib_degrid(const void *grid, int grid_row_stride, int64_t grid_pol_stride, int grid_size,
               int dtype,
               const int16_t *uv, const int16_t *w_plane, const float *weights, void *vis,
               const void *lut, int lut_slice_stride, int lut_tap_offset,
               int w_planes, int oversample, int kernel_width, int num_pols,
               int64_t num_vis, int32_t *num_rejected, kib_stream_t stream);

/* ------------------------------------------------------- grid <-> image stage
 * kib_grid_to_layer replaces the memset + 4 quadrant copy_region calls of
 * GridToImage._run (image.py:659-671): layer = zero-padded ifftshift of one
 * polarization plane of the centred grid, in one pass. */
int kib_grid_to_layer(void *layer, int layer_row_stride, int layer_size,
                      const void *grid_plane, int grid_row_stride, int grid_size,
                      int dtype, kib_stream_t stream);
/* inverse: centre crop + fftshift of the layer into one grid plane
 * (ImageToGrid._run, image.py:730-740). */
int kib_layer_to_grid(void *grid_plane, int grid_row_stride, int grid_size,
                      const void *layer, int layer_row_stride, int layer_size,
                      int dtype, kib_stream_t stream);
/* kib_layer_to_image replaces _LayerImage._run (image.py:153-180) +
 * layer_to_image.mako: for every pixel (y, x) of an size x size image plane
 *   n = sqrt(1 - l(x)^2 - m(y)^2), l(x) = x*lm_scale + lm_bias
 *   image[y][x] += Re(layer[ifftshift(y,x)] * exp(2 pi i w (n-1))) * n / (kernel1d[y]*kernel1d[x])
 * (oracle: image.py:781-799; note the host class multiplies np.fft.ifft2 by
 * size^2 to match the unnormalised cuFFT result). */
int kib_layer_to_image(void *image_plane, int image_row_stride,
                       const void *layer, int layer_row_stride, int size,
                       const void *kernel1d, double lm_scale, double lm_bias, double w,
                       int dtype, kib_stream_t stream);
/* kib_grid_to_image replaces the whole per-polarization body of GridToImage._run
 * (image.py:655-673: layer memset + 4 quadrant copies, the inverse cuFFT of
 * image.py:649-653, and LayerToImage) by a pruned, fused transform that never builds
 * the zero-padded layer:
 *   kib_grid_to_image_columns  inverse DFT along the rows of the grid_size^2 grid, only
 *                              for its non-zero columns -> scratch (size rows of
 *                              grid_size complex values, row stride scratch_row_stride).
 *                              A decimation-in-frequency fold of the grid followed by the
 *                              sub-transforms of its 64 KB tiles: one cluster kernel that
 *                              exchanges the fold through distributed shared memory for
 *                              sizes up to 8192, else two kernels with the tiles in
 *                              fold_scratch (kib_grid_to_image_fold_bytes bytes);
 *   kib_grid_to_image_rows     one size-point inverse FFT per image row in shared memory
 *                              with the layer_to_image arithmetic (see kib_layer_to_image)
 *                              applied from registers; accumulates into image_plane.
 *                              The per-pixel factor exp(2 pi i w (n-1)) n / (k1d[y] k1d[x])
 *                              does not depend on the polarization: factor_mode 1 also
 *                              stores it in `factors` (size x size complex, row stride
 *                              size), 2 loads it from there, 0 ignores `factors`.
 *                              factor_mode 3 / 4: as 1 / 2 when lm_bias = -size / 2 *
 *                              lm_scale and kernel1d[i] = kernel1d[size - i] (what the
 *                              reference's Imaging sets up, imaging.py:90-91, grid.py:404):
 *                              the factor is then symmetric about the image centre and
 *                              `factors` holds one quadrant, (size / 2 + 1)^2 complex
 *                              values at [|y - size/2|][|x - size/2|].
 * kib_grid_to_image runs both (factor_mode 0).  The reference's layer buffer is large
 * enough as scratch.  Single precision and size in {2048, 4096, 8192, 16384} only;
 * kib_grid_to_image_supported returns 1 for supported combinations and 0 otherwise
 * (callers then use kib_grid_to_layer + kib_fft_plan2d_exec + kib_layer_to_image). */
int kib_grid_to_image_supported(int size, int grid_size, int dtype);
int kib_grid_to_image_fold_bytes(int size, int grid_size, int64_t *bytes);
/* Kernels one kib_grid_to_image_columns call launches: 1 where the column pass runs as a single
 * thread-block-cluster kernel (fold butterflies exchanged through distributed shared memory,
 * no fold tiles in global memory: sizes up to 8192), 2 (fold + tile transforms) otherwise. */
int kib_grid_to_image_columns_kernels(int size);
int kib_grid_to_image_columns(void *scratch, int scratch_row_stride, int size,
                              const void *grid_plane, int grid_row_stride, int grid_size,
                              void *fold_scratch, int dtype, kib_stream_t stream);
int kib_grid_to_image_rows(void *image_plane, int image_row_stride,
                           const void *scratch, int scratch_row_stride, int grid_size, int size,
                           const void *kernel1d, double lm_scale, double lm_bias, double w,
                           void *factors, int factor_mode, int dtype, kib_stream_t stream);
int kib_grid_to_image(void *image_plane, int image_row_stride,
                      const void *grid_plane, int grid_row_stride, int grid_size,
                      void *scratch, int scratch_row_stride, void *fold_scratch, int size,
                      const void *kernel1d, double lm_scale, double lm_bias, double w,
                      int dtype, kib_stream_t stream);
/* kib_image_to_grid_rows + kib_image_to_grid_columns replace the per-polarization body of
 * ImageToGrid._run (image.py:716-740: image_to_layer.mako, the forward cuFFT and the four
 * fftshift copies of the centre of the layer into the grid) by the mirror image of the
 * transform above, with the same restrictions (kib_grid_to_image_supported):
 *   kib_image_to_grid_rows     image row -> image_to_layer arithmetic (see
 *                              kib_image_to_layer) -> forward size-point FFT in shared
 *                              memory -> the grid_size columns the grid keeps -> scratch
 *                              (size rows of grid_size complex values).  factor_mode as in
 *                              kib_grid_to_image_rows (the factor here is
 *                              exp(-2 pi i w (n-1)) / (k1d[y] k1d[x] n)); factor_mode 3 =
 *                              mode 0 plus: an image row that is entirely zero (a CLEAN
 *                              model is zero almost everywhere) is answered with zeros
 *                              without being transformed;
 *   kib_image_to_grid_columns  forward DFT along the rows of scratch, only for the
 *                              grid_size output rows the grid keeps -> grid_plane
 *                              (tile transforms into fold_scratch, then one butterfly
 *                              per output element). */
int kib_image_to_grid_rows(void *scratch, int scratch_row_stride, int grid_size, int size,
                           const void *image_plane, int image_row_stride,
                           const void *kernel1d, double lm_scale, double lm_bias, double w,
                           void *factors, int factor_mode, int dtype, kib_stream_t stream);
int kib_image_to_grid_columns(void *grid_plane, int grid_row_stride, int grid_size,
                              const void *scratch, int scratch_row_stride, int size,
                              void *fold_scratch, int dtype, kib_stream_t stream);
/* Image -> grid of a SPARSE image (a CLEAN model: zero except for a few hundred pixels; the
 * result is identical to the dense route for any image).  kib_image_to_grid_rows_sparse finds
 * the rows of the plane that hold a non-zero pixel (row_info: int32[2 * size + 1], owned by the
 * caller: [0] count, [1 .. size] layer rows, [size + 1 ..] one flag per layer row), transforms
 * only those (factor computed on the fly) and leaves the other rows of `scratch` untouched;
 * kib_image_to_grid_columns_sparse takes rows flagged empty as zero without reading them.
 * Available where the column pass is the single cluster kernel
 * (kib_image_to_grid_sparse_supported). */
int kib_image_to_grid_sparse_supported(int size, int grid_size, int dtype);
int kib_image_to_grid_rows_sparse(void *scratch, int scratch_row_stride, int grid_size, int size,
                                  const void *image_plane, int image_row_stride,
                                  const void *kernel1d, double lm_scale, double lm_bias, double w,
                                  int32_t *row_info, int dtype, kib_stream_t stream);
/* The same with a row_info that an earlier kib_image_to_grid_rows_sparse call filled for the
 * same, unchanged image plane (the model is transformed once per W slice between two batches of
 * CLEAN cycles: one classification pass over the plane instead of one per slice). */
int kib_image_to_grid_rows_classified(void *scratch, int scratch_row_stride, int grid_size,
                                      int size, const void *image_plane, int image_row_stride,
                                      const void *kernel1d, double lm_scale, double lm_bias,
                                      double w, int32_t *row_info, int dtype, kib_stream_t stream);
int kib_image_to_grid_columns_sparse(void *grid_plane, int grid_row_stride, int grid_size,
                                     const void *scratch, int scratch_row_stride, int size,
                                     const int32_t *row_info, int dtype, kib_stream_t stream);

/* Column occupancy of a W slice.  The reference transforms every column of the zero-padded layer
 * (image.py:649-673, :716-740) although a visibility only touches kernel_width columns of the
 * grid: the occupied groups of 8 columns are 3 - 87 % of a MeerKAT W slice.  `occupancy` is a
 * bit mask owned by the caller, (grid_size / 8 + 31) / 32 + 1 words, bit g = columns
 * [8 g, 8 g + 8) may hold data.  kib_column_occupancy ORs the footprints (origin as in kib_grid:
 * u - ((K - 1) / 2 - G / 2)) of num_vis coordinates into it; `uv` points at the first int16 u,
 * consecutive visibilities stride_bytes apart (8 for the uv slot, the record size for
 * preprocessed records).  The *_occ entry points are kib_grid_to_image_columns / _rows and
 * kib_image_to_grid_columns (row_info as in kib_image_to_grid_columns_sparse, or NULL) restricted
 * to the occupied groups: grid -> image takes every other column of the grid as zero without
 * reading it (the caller promises that it is), image -> grid computes the occupied groups and
 * leaves the other columns of the grid unspecified (untouched, or computed when they share a
 * tile with an occupied group; the caller promises not to read them).  Results on the occupied
 * columns are bit-identical to the dense entry points. */
int kib_column_occupancy(const void *uv, int64_t stride_bytes, int64_t num_vis, int kernel_width,
                         int grid_size, uint32_t *occupancy, kib_stream_t stream);
int kib_grid_to_image_columns_occ(void *scratch, int scratch_row_stride, int size,
                                  const void *grid_plane, int grid_row_stride, int grid_size,
                                  void *fold_scratch, const uint32_t *occupancy, int dtype,
                                  kib_stream_t stream);
/* kib_clear_columns zeroes the occupied column groups of a float32 grid (all polarizations) and
 * leaves the others alone: enough for a grid that is then filled by kib_grid with the same
 * visibilities and read through the *_occ transforms (replaces the whole-buffer memset of
 * Imaging.clear_grid, reference imaging.py:253-255). */
int kib_clear_columns(void *grid, int grid_row_stride, int64_t grid_pol_stride, int grid_size,
                      int num_pols, const uint32_t *occupancy, int dtype, kib_stream_t stream);
/* The row pass wants the mask per first-stage butterfly: kib_row_presence turns `occupancy` into
 * `presence` (size / 16 uint16, owned by the caller) for an image of size^2 pixels;
 * kib_grid_to_image_rows_occ takes that table. */
int kib_row_presence(const uint32_t *occupancy, int grid_size, int size, uint16_t *presence,
                     kib_stream_t stream);
int kib_grid_to_image_rows_occ(void *image_plane, int image_row_stride,
                               const void *scratch, int scratch_row_stride, int grid_size, int size,
                               const void *kernel1d, double lm_scale, double lm_bias, double w,
                               void *factors, int factor_mode, const uint16_t *presence,
                               int dtype, kib_stream_t stream);
int kib_image_to_grid_columns_occ(void *grid_plane, int grid_row_stride, int grid_size,
                                  const void *scratch, int scratch_row_stride, int size,
                                  void *fold_scratch, const int32_t *row_info,
                                  const uint32_t *occupancy, int dtype, kib_stream_t stream);

/* kib_image_to_layer replaces image_to_layer.mako (oracle image.py:836-843):
 *   layer[ifftshift(y,x)] = image[y][x] / (kernel1d[y]*kernel1d[x]*n) * exp(-2 pi i w (n-1)) */
int kib_image_to_layer(void *layer, int layer_row_stride,
                       const void *image_plane, int image_row_stride, int size,
                       const void *kernel1d, double lm_scale, double lm_bias, double w,
                       int dtype, kib_stream_t stream);

/* Scale._run (image.py:351-367, scale.mako): image[p] *= scale[p]; scale is a
 * HOST array of num_pols doubles (passed by value to the kernel). */
int kib_scale(void *image, int row_stride, int64_t pol_stride, int width, int height,
              int num_pols, const double *scale, int dtype, kib_stream_t stream);
/* AddImage._run (image.py:439-458, add_image.mako): dest += src */
int kib_add_image(void *dest, int dest_row_stride, int64_t dest_pol_stride,
                  const void *src, int src_row_stride, int64_t src_pol_stride,
                  int width, int height, int num_pols, int dtype, kib_stream_t stream);
/* ApplyPrimaryBeam._run (image.py:539-558, apply_primary_beam.mako):
 *   image[p] = beam < threshold ? replacement : image[p] / beam */
int kib_apply_primary_beam(void *image, int row_stride, int64_t pol_stride,
                           const void *beam_power, int width, int height, int num_pols,
                           double threshold, double replacement, int dtype,
                           kib_stream_t stream);

/* -------------------------------------------------------------------- CLEAN
 * Tiles are 32x32 pixels starting at `border` pixels from each image edge
 * (clean.py:431-433, 996-1001).  tile_max is real[tiles_y][tile_stride],
 * tile_pos is int32[tiles_y][tile_stride][2] holding (row, col).
 *
 * kib_update_tiles replaces _UpdateTiles.__call__ (clean.py:451-480) +
 * update_tiles.mako, but follows the HOST tie-break exactly so CLEAN component
 * indices are bit-exact (clean.py:947-968): strict `>` scan in row-major order
 * starting from best = 0; a tile with no positive metric stores value 0 and
 * position (x0, y0) [sic, clean.py:950].  SUMSQ metric is accumulated without
 * FMA contraction in polarization order.  Updates tiles [tx0,tx1) x [ty0,ty1). */
int kib_update_tiles(const void *dirty, int row_stride, int64_t pol_stride,
                     int width, int height, int num_pols, int border, int mode,
                     void *tile_max, int32_t *tile_pos, int tile_stride,
                     int tx0, int ty0, int tx1, int ty1, int dtype, kib_stream_t stream);
/* kib_find_peak replaces _FindPeak._run (clean.py:566-587) + find_peak.mako:
 * first maximum of tile_max in row-major tile order (np.argmax, clean.py:1062);
 * writes peak_value[1], peak_pos[2] = (row, col), peak_pixel[num_pols]. */
int kib_find_peak(const void *dirty, int row_stride, int64_t pol_stride, int num_pols,
                  const void *tile_max, const int32_t *tile_pos, int tile_stride,
                  int tiles_x, int tiles_y,
                  void *peak_value, int32_t *peak_pos, void *peak_pixel,
                  int dtype, kib_stream_t stream);
/* kib_subtract_psf replaces _SubtractPsf.__call__ (clean.py:683-726) +
 * subtract_psf.mako:  scale[p] = loop_gain * peak_pixel[p] (rounded to dtype),
 *   dirty[p][y][x] -= scale[p] * psf[p][...] over the patch centred on (pos_y, pos_x),
 *   clipped to the image; model[p][pos_y][pos_x] += scale[p].
 * Multiplication and subtraction are rounded separately (no FMA), as numpy
 * does in CleanHost._subtract_psf (clean.py:1044-1047). */
int kib_subtract_psf(void *dirty, void *model, int row_stride, int64_t pol_stride,
                     int width, int height, int num_pols,
                     const void *psf, int psf_row_stride, int64_t psf_pol_stride,
                     int psf_width, int psf_height,
                     int patch_width, int patch_height,
                     const void *peak_pixel, int pos_y, int pos_x, double loop_gain,
                     int dtype, kib_stream_t stream);

/* Device-resident minor-cycle loop (NEW; the reference round-trips to the host
 * every cycle, clean.py:870-891).  Runs up to max_cycles iterations of
 *   { test peak_value < threshold -> stop;  subtract;  update tiles under patch;
 *     find next peak }
 * entirely on the device.  On entry peak_value/peak_pos/peak_pixel must hold
 * the current peak (kib_find_peak).  Each executed cycle appends a record
 * to `components` (int32 y, int32 x, real value, real pixel[num_pols] laid out
 * as `component_stride` bytes per record) and increments state[0]; state[1] is
 * set to 1 when the threshold stopped the loop.  `state` is int32[4] device
 * memory zeroed by the caller before the first call of a batch.  `row_scratch` is
 * tiles_y * (sizeof(real) + 4) bytes of device scratch (per-row tile maxima, rebuilt by
 * every call, so the next peak is found without rescanning all tiles each cycle). */
int kib_clean_minor_cycles(void *dirty, void *model, int row_stride, int64_t pol_stride,
                           int width, int height, int num_pols, int border, int mode,
                           const void *psf, int psf_row_stride, int64_t psf_pol_stride,
                           int psf_width, int psf_height,
                           int patch_width, int patch_height,
                           void *tile_max, int32_t *tile_pos, int tile_stride,
                           int tiles_x, int tiles_y,
                           void *peak_value, int32_t *peak_pos, void *peak_pixel,
                           double loop_gain, double threshold, int max_cycles,
                           void *components, int component_stride, int32_t *state,
                           void *row_scratch, int dtype, kib_stream_t stream);
/* Kernels one kib_clean_minor_cycles call launches (launch accounting): the row-maxima
 * pass plus ONE cooperative persistent kernel that runs every cycle of the batch behind a grid
 * barrier (float32), or -- float64, KIB_CLEAN_ROUTE=pdl, or no cooperative launch -- one
 * kernel per cycle.  `state` must hold 16 int32 (zeroed by the caller before every call). */
int kib_clean_minor_cycles_launches(int max_cycles, int dtype);

/* kib_psf_patch replaces PsfPatch.__call__ (clean.py:123-163) + psf_patch.mako:
 * bound[0] = max |x - mid_x|, bound[1] = max |y - mid_y| over pixels in
 * [min_x,max_x] x [min_y,max_y] with |psf| >= threshold in any polarization.
 * `bound` is device int32[2], zeroed by this call. */
int kib_psf_patch(const void *psf, int row_stride, int64_t pol_stride, int num_pols,
                  int min_x, int min_y, int max_x, int max_y, int mid_x, int mid_y,
                  double threshold, int32_t *bound, int dtype, kib_stream_t stream);

/* Noise estimation (clean.py:295-353 + rank.mako use ~32 rank passes with a host
 * round trip each; oracle noise_est_host clean.py:938-943 takes an exact
 * median).  kib_abs_histogram is one radix pass of an exact selection over
 * |image| inside the border: counts, into hist (device uint32[1 << bits],
 * zeroed by the caller), bits [shift, shift+bits) of the IEEE bit pattern of
 * |pixel| for pixels whose higher bits equal `prefix` (ignored when
 * prefix_bits == 0).  f32 only. */
int kib_abs_histogram(const void *image, int row_stride, int64_t pol_stride,
                      int width, int height, int num_pols, int border,
                      uint32_t prefix, int prefix_bits, int shift, int bits,
                      uint32_t *hist, int dtype, kib_stream_t stream);
/* kib_abs_histogram_window: digit (shift, bits) of |pixel| inside the border for `window`
 * adjacent leading prefixes first_prefix ... (hist[window][1 << bits], at most 8192 bins in all)
 * and the count of values whose prefix is smaller (*below): the first two radix passes of the
 * exact median in one when the leading digit can be guessed (NoiseEst, clean.py:247-353). */
int kib_abs_histogram_window(const void *image, int row_stride, int64_t pol_stride,
                             int width, int height, int num_pols, int border,
                             uint32_t first_prefix, int window, int prefix_bits, int shift,
                             int bits, uint32_t *hist, unsigned long long *below, int dtype,
                             kib_stream_t stream);
/* The reference's own kernel, kept for API parity (rank.mako): number of
 * |pixels| strictly below `value` inside the border, accumulated into
 * rank (device uint64[1], zeroed by the caller). */
int kib_rank(const void *image, int row_stride, int64_t pol_stride,
             int width, int height, int num_pols, int border, double value,
             unsigned long long *rank, int dtype, kib_stream_t stream);

/* ------------------------------------------------------------------ weights
 * kib_grid_weights: GridWeights._run (weight.py:155-176) + grid_weights.mako:
 *   grid[p][uv.y + G_h/2][uv.x + G_w/2] += weights[i][p]   (uv = first two of 4 int16) */
int kib_grid_weights(float *grid, int row_stride, int64_t pol_stride, int width, int height,
                     const int16_t *uv, const float *weights, int num_pols, int64_t num_vis,
                     kib_stream_t stream);
/* kib_mean_weight: MeanWeight._run (weight.py:357-376) + mean_weight.mako:
 * sums[0] += sum W, sums[1] += sum W^2 over polarization 0 (device double[2]). */
int kib_mean_weight(const float *grid, int row_stride, int width, int height,
                    double *sums, kib_stream_t stream);
/* kib_density_weights: DensityWeights._run (weight.py:261-284) +
 * density_weights.mako: d = W != 0 ? 1/(a*W+b) : 0 written in place for every
 * polarization; sums[0..2] += sum W, sum d*W, sum d^2*W over polarization 0
 * (device double[3]). */
int kib_density_weights(float *grid, int row_stride, int64_t pol_stride, int width, int height,
                        int num_pols, float a, float b, double *sums, kib_stream_t stream);
/* kib_fourier_beam replaces FourierBeam._run (beam.py:271-301) + fourier_beam.mako: the
 * half-complex transform of an image times amplitude * exp(a v^2 + b u v + c u^2), u = column,
 * v = signed row frequency (restoring-beam convolution). */
int kib_fourier_beam(void *data, int row_stride, double amplitude, double a, double b, double c,
                     int width, int height, int dtype, kib_stream_t stream);
/* kib_fits_plane writes an image in the order the reference's FITS writer stores it
 * (io.py:186-200: l axis reversed, big-endian, rows packed): out[p][y][x] =
 * byteswap(image[p][y][width-1-x]).  float32 only. */
int kib_fits_plane(void *out, const void *image, int row_stride, int64_t pol_stride,
                   int width, int height, int num_pols, int dtype, kib_stream_t stream);
/* kib_fill: katsdpsigproc.fill.FillTemplate (weight.py:403,480-484). */
int kib_fill(void *data, int row_stride, int64_t pol_stride, int width, int height,
             int num_pols, double value, int dtype, kib_stream_t stream);

/* ------------------------------------------------------- visibility records
 * Splits preprocessed visibility records (the array-of-structures layout of
 * preprocess.cpp:39-52 minus w_slice: int16 uv[2], int16 sub_uv[2], int16 w_plane,
 * 2 bytes padding, float weights[P], complex64 vis[P]; record_bytes = 12 + 12 P) that
 * were uploaded in one block into the structure-of-arrays buffers the kernels use.
 * Replaces the host-side field extraction of Imaging._set_uv / _set_buffer
 * (imaging.py:269-314), which costs more than all device work of a step.
 * Any output pointer may be NULL.  If vis_from_weights is non-zero, `vis` receives
 * the weights as real numbers (the PSF pass grids the weights, frontend.py:511). */
int kib_unpack_records(const void *records, int record_bytes, int64_t num_vis, int num_pols,
                       int16_t *uv, int16_t *w_plane, float *weights, void *vis,
                       int vis_from_weights, kib_stream_t stream);

/* ------------------------------------------------------------------ predict
 * kib_predict replaces Predict._run (predict.py:386-416) + predict.mako:
 *   u = (uv.x*oversample + sub_u + 0.5)*uv_scale, v likewise, w = w_plane*w_scale + w_bias
 *   vis[i][p] -= weights[i][p] * sum_s flux[s][p] * exp(-2 pi i (l_s u + m_s v + n1_s w))
 * lmn float[num_sources][3] = (l, m, n-1); flux float[num_sources][num_pols]
 * (oracle: predict.py:420-438). */
int kib_predict(void *vis, const int16_t *uv, const int16_t *w_plane, const float *weights,
                const float *lmn, const float *flux, int64_t num_vis, int num_sources,
                int num_pols, int oversample, float uv_scale, float w_scale, float w_bias,
                kib_stream_t stream);

/* -------------------------------------------------------------- measurement
 * FFMA micro-benchmark used by bench.py for the FP32 roofline denominator
 * (MEASURED_PEAKS.json has no non-tensor FP32 figure).  Runs `iters` rounds
 * of 8 independent FFMA chains x 64 per thread on blocks x 256 threads and
 * returns the flop count in *flops; the caller times it with events. */
int kib_fp32_peak_kernel(float *sink, int blocks, int iters, double *flops, kib_stream_t stream);

/* ---- visibility preprocessing
 * kib_preprocess replaces visibility_collector<P>::add for ONE channel (reference
 * katsdpimager/preprocess.cpp:401-513 add_impl2 + :335-397 compress; pybind11 entry point
 * `_preprocess.VisibilityCollector.add`, preprocess.cpp:673): Mueller/Stokes transform with
 * MulZ arithmetic, weight transform, w < 0 flip + conjugate, quantisation (subpixel_coord
 * :313-323), flagged samples dropped, adjacent duplicates merged within each `capacity`-sample
 * buffer (0 = one buffer), stable bucket sort by W slice.
 *   uvw float[n][3] metres, weights float[n][Q], vis complex64[n][Q]           (device)
 *   feed_angle1/2 float[n] or NULL; mueller_stokes complex64 [P][Q] (no feed angles) or
 *   [P][4]; mueller_circular complex64 [4][Q] or NULL                            (device)
 *   records: n x (12 + 12 P) bytes {int16 uv[2], sub_uv[2], w_plane, w_slice; float
 *   weights[P]; complex64 vis[P]} in W-slice order (device); counts[w_slices]  (HOST, the
 *   call synchronises the stream).  scratch: kib_preprocess_scratch_bytes. */
int kib_preprocess_scratch_bytes(int64_t num_vis, int num_pols, int w_slices, int64_t *bytes);
int kib_preprocess(const float *uvw, const float *weights, const void *vis, int64_t num_vis,
                   int num_in_pols, const float *feed_angle1, const float *feed_angle2,
                   const void *mueller_stokes, const void *mueller_circular, int num_pols,
                   double cell_size, double max_w, int w_slices, int w_planes, int oversample,
                   int64_t capacity, void *records, int64_t *host_counts,
                   void *scratch, int64_t scratch_bytes, kib_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* KATIMAGER_B200_H */
