// Grid <-> image stage kernels for sm_100a (everything around the cuFFT call).
//
// Replaces GridToImage._run / ImageToGrid._run (reference katsdpimager/image.py:649-673,
// 716-740), layer_to_image.mako / image_to_layer.mako / layer_image.mako, scale.mako,
// add_image.mako and apply_primary_beam.mako.  Numerics follow GridToImageHost /
// ImageToGridHost (image.py:781-799, 836-848): direction cosines are evaluated in the
// image precision with separately rounded multiply and add (numpy semantics, no FMA),
// so n = sqrt(1 - (l^2 + m^2)) is bit-identical to the host; the W phase w*(n-1) is
// formed and range-reduced in double precision (the reference's frontend passes w as
// a float64 scalar, which promotes the host computation the same way).
//
// All kernels are HBM-bound streaming passes: one thread per (pair of) element(s),
// fully coalesced rows, no shared memory.
#include "kib_common.cuh"
#include "kib_imagemath.cuh"

namespace kib {

template <typename Real> struct Vec2;
template <> struct Vec2<float> { typedef float2 type; };
template <> struct Vec2<double> { typedef double2 type; };

// ------------------------------------------------------------ grid -> layer (pad + ifftshift)
// One thread writes two adjacent layer elements (one 16-byte store for complex64).  Layer
// index f in [0, half) holds non-negative frequencies (grid index f + half); [N - half, N)
// holds negative frequencies (grid index f - (N - half)); everything else is zero.
template <typename Complex> struct Pair;
template <> struct __align__(16) Pair<float2> { float2 a, b; };
template <> struct __align__(32) Pair<double2> { double2 a, b; };

template <typename Complex>
__global__ void __launch_bounds__(256)
grid_to_layer_kernel(Complex *__restrict__ layer, int layer_row_stride, int N,
                     const Complex *__restrict__ grid, int grid_row_stride, int G)
{
    const int x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y = blockIdx.y;
    if (x >= N) return;
    const int half = G / 2;
    int gy = -1;
    if (y < half) gy = y + half;
    else if (y >= N - half) gy = y - (N - half);
    Pair<Complex> out;
    out.a.x = out.a.y = out.b.x = out.b.y = 0;
    if (gy >= 0) {
        const Complex *row = grid + (long long) gy * grid_row_stride;
        int gx0 = -1, gx1 = -1;
        if (x < half) gx0 = x + half;
        else if (x >= N - half) gx0 = x - (N - half);
        if (x + 1 < half) gx1 = x + 1 + half;
        else if (x + 1 >= N - half) gx1 = x + 1 - (N - half);
        if (gx0 >= 0) out.a = __ldg(row + gx0);
        if (gx1 >= 0) out.b = __ldg(row + gx1);
    }
    Complex *dst = layer + (long long) y * layer_row_stride + x;
    if ((layer_row_stride & 1) == 0) {
        *reinterpret_cast<Pair<Complex> *>(dst) = out;
    } else {
        dst[0] = out.a;
        dst[1] = out.b;
    }
}

template <typename Complex>
__global__ void __launch_bounds__(256)
layer_to_grid_kernel(Complex *__restrict__ grid, int grid_row_stride, int G,
                     const Complex *__restrict__ layer, int layer_row_stride, int N)
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int gy = blockIdx.y;
    if (gx >= G) return;
    const int half = G / 2;
    const int y = gy >= half ? gy - half : gy + (N - half);
    const int x = gx >= half ? gx - half : gx + (N - half);
    grid[(long long) gy * grid_row_stride + gx] = layer[(long long) y * layer_row_stride + x];
}

// ------------------------------------------------------------ layer <-> image
// One thread handles the four pixels (x, y), (x + h, y), (x, y + h), (x + h, y + h),
// h = size / 2, which map to the four layer elements at the same offsets with the halves
// swapped (fftshift); they share kernel1d / l^2 / m^2 values.
template <typename Real, bool TO_IMAGE>
__global__ void __launch_bounds__(256)
layer_image_kernel(Real *__restrict__ image, int image_row_stride,
                   typename Vec2<Real>::type *__restrict__ layer, int layer_row_stride,
                   int half, const Real *__restrict__ kernel1d,
                   Real lm_scale, Real lm_bias, double w)
{
    typedef typename Vec2<Real>::type Complex;
    const int x0 = blockIdx.x * blockDim.x + threadIdx.x;
    const int y0 = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= half || y0 >= half) return;
    int xs[2] = {x0, x0 + half};
    int ys[2] = {y0, y0 + half};
    Real kx[2], ky[2], l2[2], m2[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        // reciprocal taper: the four pixels share two divisions per axis
        kx[i] = (Real) 1 / kernel1d[xs[i]];
        ky[i] = (Real) 1 / kernel1d[ys[i]];
        const Real l = add_rn(mul_rn((Real) xs[i], lm_scale), lm_bias);
        const Real m = add_rn(mul_rn((Real) ys[i], lm_scale), lm_bias);
        l2[i] = mul_rn(l, l);
        m2[i] = mul_rn(m, m);
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const long long iaddr = (long long) ys[i] * image_row_stride + xs[j];
            const long long laddr = (long long) ys[1 - i] * layer_row_stride + xs[1 - j];
            const Real n = sqrt(add_rn((Real) 1, -add_rn(m2[i], l2[j])));
            Real c, s;
            w_rotation<Real>(n, w, &c, &s);
            if (TO_IMAGE) {
                const Complex v = layer[laddr];
                // Re(v * exp(2 pi i w (n - 1))) * n / (ky * kx)
                const Real rotated = v.x * c - v.y * s;
                image[iaddr] += rotated * n * (ky[i] * kx[j]);
            } else {
                // image / (ky * kx * n) * exp(-2 pi i w (n - 1))
                const Real v = image[iaddr] * (ky[i] * kx[j]) / n;
                Complex out;
                out.x = v * c;
                out.y = -(v * s);
                layer[laddr] = out;
            }
        }
}

// Vectorised single-precision layer -> image: each thread handles two adjacent columns in
// each of the four mirrored quadrants (8 pixels): 16-byte layer loads, 8-byte image
// loads/stores.  Requires size / 2 even (all FFT-friendly sizes are multiples of 8).
__global__ void __launch_bounds__(256)
layer_to_image_x2_kernel(float *__restrict__ image, int image_row_stride,
                         const float2 *__restrict__ layer, int layer_row_stride,
                         int half, const float *__restrict__ kernel1d,
                         float lm_scale, float lm_bias, double w)
{
    const int x0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y0 = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= half || y0 >= half) return;
    const int xs[2] = {x0, x0 + half};
    const int ys[2] = {y0, y0 + half};
    // issue all loads first
    float4 lay[2][2];
    float2 img[2][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            lay[i][j] = *reinterpret_cast<const float4 *>(
                layer + (long long) ys[1 - i] * layer_row_stride + xs[1 - j]);
            img[i][j] = *reinterpret_cast<const float2 *>(
                image + (long long) ys[i] * image_row_stride + xs[j]);
        }
    float kx[2][2], ky[2], l2[2][2], m2[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float2 k2 = *reinterpret_cast<const float2 *>(kernel1d + xs[i]);
        kx[i][0] = 1.0f / k2.x;
        kx[i][1] = 1.0f / k2.y;
        ky[i] = 1.0f / kernel1d[ys[i]];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const float l = __fadd_rn(__fmul_rn((float) (xs[i] + q), lm_scale), lm_bias);
            l2[i][q] = __fmul_rn(l, l);
        }
        const float m = __fadd_rn(__fmul_rn((float) ys[i], lm_scale), lm_bias);
        m2[i] = __fmul_rn(m, m);
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            float out[2];
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float n = sqrtf(__fadd_rn(1.0f, -__fadd_rn(m2[i], l2[j][q])));
                float c, s;
                w_rotation<float>(n, w, &c, &s);
                const float vx = q == 0 ? lay[i][j].x : lay[i][j].z;
                const float vy = q == 0 ? lay[i][j].y : lay[i][j].w;
                const float rotated = vx * c - vy * s;
                out[q] = (q == 0 ? img[i][j].x : img[i][j].y) + rotated * n * (ky[i] * kx[j][q]);
            }
            *reinterpret_cast<float2 *>(image + (long long) ys[i] * image_row_stride + xs[j]) =
                make_float2(out[0], out[1]);
        }
}

// Vectorised single-precision image -> layer (mirror image of layer_to_image_x2_kernel).
__global__ void __launch_bounds__(256)
image_to_layer_x2_kernel(float2 *__restrict__ layer, int layer_row_stride,
                         const float *__restrict__ image, int image_row_stride,
                         int half, const float *__restrict__ kernel1d,
                         float lm_scale, float lm_bias, double w)
{
    const int x0 = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int y0 = blockIdx.y * blockDim.y + threadIdx.y;
    if (x0 >= half || y0 >= half) return;
    const int xs[2] = {x0, x0 + half};
    const int ys[2] = {y0, y0 + half};
    float2 img[2][2];
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++)
            img[i][j] = *reinterpret_cast<const float2 *>(
                image + (long long) ys[i] * image_row_stride + xs[j]);
    float kx[2][2], ky[2], l2[2][2], m2[2];
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const float2 k2 = *reinterpret_cast<const float2 *>(kernel1d + xs[i]);
        kx[i][0] = 1.0f / k2.x;
        kx[i][1] = 1.0f / k2.y;
        ky[i] = 1.0f / kernel1d[ys[i]];
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const float l = __fadd_rn(__fmul_rn((float) (xs[i] + q), lm_scale), lm_bias);
            l2[i][q] = __fmul_rn(l, l);
        }
        const float m = __fadd_rn(__fmul_rn((float) ys[i], lm_scale), lm_bias);
        m2[i] = __fmul_rn(m, m);
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) {
            float4 out;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const float n = sqrtf(__fadd_rn(1.0f, -__fadd_rn(m2[i], l2[j][q])));
                float c, s;
                w_rotation<float>(n, w, &c, &s);
                const float v = (q == 0 ? img[i][j].x : img[i][j].y) * (ky[i] * kx[j][q]) / n;
                if (q == 0) {
                    out.x = v * c;
                    out.y = -(v * s);
                } else {
                    out.z = v * c;
                    out.w = -(v * s);
                }
            }
            *reinterpret_cast<float4 *>(layer + (long long) ys[1 - i] * layer_row_stride + xs[1 - j]) = out;
        }
}

// ------------------------------------------------------------ elementwise image ops
struct ScaleFactors {
    double v[4];
};

template <typename Real>
__global__ void __launch_bounds__(256)
scale_kernel(Real *__restrict__ image, int row_stride, long long pol_stride, int width, int height,
             int num_pols, ScaleFactors scale)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= width) return;
    long long addr = (long long) y * row_stride + x;
    for (int p = 0; p < num_pols; p++, addr += pol_stride)
        image[addr] *= (Real) scale.v[p];
}

template <typename Real>
__global__ void __launch_bounds__(256)
add_image_kernel(Real *__restrict__ dest, int dest_row_stride, long long dest_pol_stride,
                 const Real *__restrict__ src, int src_row_stride, long long src_pol_stride,
                 int width, int height, int num_pols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= width) return;
    long long d = (long long) y * dest_row_stride + x;
    long long s = (long long) y * src_row_stride + x;
    for (int p = 0; p < num_pols; p++, d += dest_pol_stride, s += src_pol_stride)
        dest[d] += src[s];
}

// float4 variants of the two streaming kernels (16-byte aligned planes, strides and widths
// that are multiples of 4): one 16-byte read-modify-write per thread and polarization,
// 8 in flight per thread at 4 polarizations.
__global__ void __launch_bounds__(256)
scale_kernel_v4(float4 *__restrict__ image, int row_stride4, long long pol_stride4, int width4,
                int num_pols, ScaleFactors scale)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width4) return;
    float4 *ptr = image + (long long) blockIdx.y * row_stride4 + x;
    float4 v[4];
#pragma unroll
    for (int p = 0; p < 4; p++)
        if (p < num_pols) v[p] = ptr[p * pol_stride4];
#pragma unroll
    for (int p = 0; p < 4; p++)
        if (p < num_pols) {
            const float f = (float) scale.v[p];
            v[p].x *= f; v[p].y *= f; v[p].z *= f; v[p].w *= f;
            ptr[p * pol_stride4] = v[p];
        }
}

__global__ void __launch_bounds__(256)
add_image_kernel_v4(float4 *__restrict__ dest, int dest_row_stride4, long long dest_pol_stride4,
                    const float4 *__restrict__ src, int src_row_stride4, long long src_pol_stride4,
                    int width4, int num_pols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= width4) return;
    float4 *d = dest + (long long) blockIdx.y * dest_row_stride4 + x;
    const float4 *s = src + (long long) blockIdx.y * src_row_stride4 + x;
    float4 a[4], b[4];
#pragma unroll
    for (int p = 0; p < 4; p++)
        if (p < num_pols) {
            a[p] = d[p * dest_pol_stride4];
            b[p] = __ldg(s + p * src_pol_stride4);
        }
#pragma unroll
    for (int p = 0; p < 4; p++)
        if (p < num_pols) {
            a[p].x += b[p].x; a[p].y += b[p].y; a[p].z += b[p].z; a[p].w += b[p].w;
            d[p * dest_pol_stride4] = a[p];
        }
}

static bool vec4_ok(const void *ptr, int row_stride, long long pol_stride, int width, int num_pols)
{
    return (reinterpret_cast<size_t>(ptr) & 15) == 0 && (row_stride & 3) == 0
        && (pol_stride & 3) == 0 && (width & 3) == 0 && num_pols <= 4;
}

template <typename Real>
__global__ void __launch_bounds__(256)
apply_primary_beam_kernel(Real *__restrict__ image, int row_stride, long long pol_stride,
                          const Real *__restrict__ beam_power, int width, int height,
                          int num_pols, Real threshold, Real replacement)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= width) return;
    long long addr = (long long) y * row_stride + x;
    const Real beam = beam_power[addr];
    const bool low = beam < threshold;
    for (int p = 0; p < num_pols; p++, addr += pol_stride)
        image[addr] = low ? replacement : image[addr] / beam;
}

template <typename Real>
__global__ void __launch_bounds__(256)
fill_kernel(Real *__restrict__ data, int row_stride, long long pol_stride, int width, int height,
            int num_pols, Real value)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= width) return;
    long long addr = (long long) y * row_stride + x;
    for (int p = 0; p < num_pols; p++, addr += pol_stride)
        data[addr] = value;
}

// FITS plane order (reference io.py:186-200): the l axis reversed (RA increases to the left)
// and every value big-endian, packed rows.  Done on the device so that the D2H copy can land
// directly in the output file's mapping.
__global__ void __launch_bounds__(256)
fits_plane_kernel(unsigned *__restrict__ out, const unsigned *__restrict__ image, int row_stride,
                  long long pol_stride, int width, int height, int num_pols)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= width) return;
    const long long in = (long long) y * row_stride + (width - 1 - x);
    const long long o = (long long) y * width + x;
    for (int p = 0; p < num_pols; p++)
        out[p * (long long) width * height + o] = __byte_perm(image[p * pol_stride + in], 0, 0x0123);
}

// Fourier transform of the image (real-to-complex layout: height x (width / 2 + 1)) times the
// analytic transform of the Gaussian restoring beam, amplitude * exp(a v^2 + b u v + c u^2)
// with v the signed row frequency and u the column frequency (beam.py:271-301,
// fourier_beam.mako).
template <typename Real>
__global__ void __launch_bounds__(256)
fourier_beam_kernel(typename Complex2<Real>::type *__restrict__ data, int stride, Real amplitude,
                    Real a, Real b, Real c, int width, int height)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= width) return;
    const Real u = (Real) x;
    const Real v = (Real) (2 * y >= height ? y - height : y);
    const Real ft = amplitude * exp((a * v + b * u) * v + c * u * u);
    typename Complex2<Real>::type value = data[(long long) y * stride + x];
    value.x *= ft;
    value.y *= ft;
    data[(long long) y * stride + x] = value;
}

static dim3 row_grid(int width, int height) { return dim3(divup(width, 256), height, 1); }

}  // namespace kib

using namespace kib;

#define KIB_CHECK_DTYPE(name)                                                        \
    KIB_REQUIRE(dtype == KIB_F32 || dtype == KIB_F64, name ": bad dtype %d", dtype)

extern "C" {

int kib_grid_to_layer(void *layer, int layer_row_stride, int layer_size,
                      const void *grid_plane, int grid_row_stride, int grid_size,
                      int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_grid_to_layer");
    KIB_REQUIRE(grid_size % 2 == 0 && grid_size <= layer_size && grid_size > 0,
                "kib_grid_to_layer: grid size %d must be even and no larger than layer size %d",
                grid_size, layer_size);
    KIB_REQUIRE(layer_size % 2 == 0, "kib_grid_to_layer: odd layer size %d", layer_size);
    dim3 g = row_grid(layer_size / 2, layer_size);
    if (dtype == KIB_F32)
        grid_to_layer_kernel<float2><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<float2 *>(layer), layer_row_stride, layer_size,
            static_cast<const float2 *>(grid_plane), grid_row_stride, grid_size);
    else
        grid_to_layer_kernel<double2><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<double2 *>(layer), layer_row_stride, layer_size,
            static_cast<const double2 *>(grid_plane), grid_row_stride, grid_size);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_layer_to_grid(void *grid_plane, int grid_row_stride, int grid_size,
                      const void *layer, int layer_row_stride, int layer_size,
                      int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_layer_to_grid");
    KIB_REQUIRE(grid_size % 2 == 0 && grid_size <= layer_size && grid_size > 0,
                "kib_layer_to_grid: grid size %d must be even and no larger than layer size %d",
                grid_size, layer_size);
    dim3 g = row_grid(grid_size, grid_size);
    if (dtype == KIB_F32)
        layer_to_grid_kernel<float2><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<float2 *>(grid_plane), grid_row_stride, grid_size,
            static_cast<const float2 *>(layer), layer_row_stride, layer_size);
    else
        layer_to_grid_kernel<double2><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<double2 *>(grid_plane), grid_row_stride, grid_size,
            static_cast<const double2 *>(layer), layer_row_stride, layer_size);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_layer_to_image(void *image_plane, int image_row_stride,
                       const void *layer, int layer_row_stride, int size,
                       const void *kernel1d, double lm_scale, double lm_bias, double w,
                       int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_layer_to_image");
    KIB_REQUIRE(size > 0 && size % 2 == 0, "kib_layer_to_image: image size must be even, not %d", size);
    const int half = size / 2;
    dim3 block(32, 8, 1);
    dim3 g(divup(half, 32), divup(half, 8), 1);
    if (dtype == KIB_F32 && half % 2 == 0 && image_row_stride % 2 == 0 && layer_row_stride % 2 == 0
        && (reinterpret_cast<uintptr_t>(image_plane) & 7) == 0
        && (reinterpret_cast<uintptr_t>(layer) & 15) == 0
        && (reinterpret_cast<uintptr_t>(kernel1d) & 7) == 0) {
        dim3 block2(64, 4, 1);
        dim3 g2(divup(half / 2, 64), divup(half, 4), 1);
        layer_to_image_x2_kernel<<<g2, block2, 0, as_stream(stream)>>>(
            static_cast<float *>(image_plane), image_row_stride,
            static_cast<const float2 *>(layer), layer_row_stride, half,
            static_cast<const float *>(kernel1d), (float) lm_scale, (float) lm_bias, w);
    } else if (dtype == KIB_F32)
        layer_image_kernel<float, true><<<g, block, 0, as_stream(stream)>>>(
            static_cast<float *>(image_plane), image_row_stride,
            static_cast<float2 *>(const_cast<void *>(layer)), layer_row_stride, half,
            static_cast<const float *>(kernel1d), (float) lm_scale, (float) lm_bias, w);
    else
        layer_image_kernel<double, true><<<g, block, 0, as_stream(stream)>>>(
            static_cast<double *>(image_plane), image_row_stride,
            static_cast<double2 *>(const_cast<void *>(layer)), layer_row_stride, half,
            static_cast<const double *>(kernel1d), lm_scale, lm_bias, w);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_image_to_layer(void *layer, int layer_row_stride,
                       const void *image_plane, int image_row_stride, int size,
                       const void *kernel1d, double lm_scale, double lm_bias, double w,
                       int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_image_to_layer");
    KIB_REQUIRE(size > 0 && size % 2 == 0, "kib_image_to_layer: image size must be even, not %d", size);
    const int half = size / 2;
    dim3 block(32, 8, 1);
    dim3 g(divup(half, 32), divup(half, 8), 1);
    if (dtype == KIB_F32 && half % 2 == 0 && image_row_stride % 2 == 0 && layer_row_stride % 2 == 0
        && (reinterpret_cast<uintptr_t>(image_plane) & 7) == 0
        && (reinterpret_cast<uintptr_t>(layer) & 15) == 0
        && (reinterpret_cast<uintptr_t>(kernel1d) & 7) == 0) {
        dim3 block2(64, 4, 1);
        dim3 g2(divup(half / 2, 64), divup(half, 4), 1);
        image_to_layer_x2_kernel<<<g2, block2, 0, as_stream(stream)>>>(
            static_cast<float2 *>(layer), layer_row_stride,
            static_cast<const float *>(image_plane), image_row_stride, half,
            static_cast<const float *>(kernel1d), (float) lm_scale, (float) lm_bias, w);
    } else if (dtype == KIB_F32)
        layer_image_kernel<float, false><<<g, block, 0, as_stream(stream)>>>(
            static_cast<float *>(const_cast<void *>(image_plane)), image_row_stride,
            static_cast<float2 *>(layer), layer_row_stride, half,
            static_cast<const float *>(kernel1d), (float) lm_scale, (float) lm_bias, w);
    else
        layer_image_kernel<double, false><<<g, block, 0, as_stream(stream)>>>(
            static_cast<double *>(const_cast<void *>(image_plane)), image_row_stride,
            static_cast<double2 *>(layer), layer_row_stride, half,
            static_cast<const double *>(kernel1d), lm_scale, lm_bias, w);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_scale(void *image, int row_stride, int64_t pol_stride, int width, int height,
              int num_pols, const double *scale, int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_scale");
    KIB_REQUIRE(num_pols >= 1 && num_pols <= 4, "kib_scale: num_pols must be 1..4");
    KIB_REQUIRE(scale != nullptr, "kib_scale: null scale");
    if (width <= 0 || height <= 0) return 0;
    ScaleFactors f = {{0, 0, 0, 0}};
    for (int p = 0; p < num_pols; p++) f.v[p] = scale[p];
    dim3 g = row_grid(width, height);
    if (dtype == KIB_F32 && vec4_ok(image, row_stride, pol_stride, width, num_pols)) {
        dim3 g4 = row_grid(width / 4, height);
        scale_kernel_v4<<<g4, 256, 0, as_stream(stream)>>>(
            static_cast<float4 *>(image), row_stride / 4, pol_stride / 4, width / 4, num_pols, f);
    } else if (dtype == KIB_F32)
        scale_kernel<float><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<float *>(image), row_stride, pol_stride, width, height, num_pols, f);
    else
        scale_kernel<double><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<double *>(image), row_stride, pol_stride, width, height, num_pols, f);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_add_image(void *dest, int dest_row_stride, int64_t dest_pol_stride,
                  const void *src, int src_row_stride, int64_t src_pol_stride,
                  int width, int height, int num_pols, int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_add_image");
    if (width <= 0 || height <= 0) return 0;
    dim3 g = row_grid(width, height);
    if (dtype == KIB_F32 && vec4_ok(dest, dest_row_stride, dest_pol_stride, width, num_pols)
        && vec4_ok(src, src_row_stride, src_pol_stride, width, num_pols)) {
        dim3 g4 = row_grid(width / 4, height);
        add_image_kernel_v4<<<g4, 256, 0, as_stream(stream)>>>(
            static_cast<float4 *>(dest), dest_row_stride / 4, dest_pol_stride / 4,
            static_cast<const float4 *>(src), src_row_stride / 4, src_pol_stride / 4,
            width / 4, num_pols);
    } else if (dtype == KIB_F32)
        add_image_kernel<float><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<float *>(dest), dest_row_stride, dest_pol_stride,
            static_cast<const float *>(src), src_row_stride, src_pol_stride,
            width, height, num_pols);
    else
        add_image_kernel<double><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<double *>(dest), dest_row_stride, dest_pol_stride,
            static_cast<const double *>(src), src_row_stride, src_pol_stride,
            width, height, num_pols);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_apply_primary_beam(void *image, int row_stride, int64_t pol_stride,
                           const void *beam_power, int width, int height, int num_pols,
                           double threshold, double replacement, int dtype,
                           kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_apply_primary_beam");
    if (width <= 0 || height <= 0) return 0;
    dim3 g = row_grid(width, height);
    if (dtype == KIB_F32)
        apply_primary_beam_kernel<float><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<float *>(image), row_stride, pol_stride,
            static_cast<const float *>(beam_power), width, height, num_pols,
            (float) threshold, (float) replacement);
    else
        apply_primary_beam_kernel<double><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<double *>(image), row_stride, pol_stride,
            static_cast<const double *>(beam_power), width, height, num_pols,
            threshold, replacement);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_fourier_beam(void *data, int row_stride, double amplitude, double a, double b, double c,
                     int width, int height, int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_fourier_beam");
    if (width <= 0 || height <= 0) return 0;
    dim3 g = row_grid(width, height);
    if (dtype == KIB_F32)
        fourier_beam_kernel<float><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<float2 *>(data), row_stride, (float) amplitude, (float) a, (float) b,
            (float) c, width, height);
    else
        fourier_beam_kernel<double><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<double2 *>(data), row_stride, amplitude, a, b, c, width, height);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_fits_plane(void *out, const void *image, int row_stride, int64_t pol_stride,
                   int width, int height, int num_pols, int dtype, kib_stream_t stream)
{
    KIB_REQUIRE(dtype == KIB_F32, "kib_fits_plane: only float32 images are supported");
    if (width <= 0 || height <= 0) return 0;
    fits_plane_kernel<<<row_grid(width, height), 256, 0, as_stream(stream)>>>(
        static_cast<unsigned *>(out), static_cast<const unsigned *>(image), row_stride,
        pol_stride, width, height, num_pols);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_fill(void *data, int row_stride, int64_t pol_stride, int width, int height,
             int num_pols, double value, int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_fill");
    if (width <= 0 || height <= 0) return 0;
    dim3 g = row_grid(width, height);
    if (dtype == KIB_F32)
        fill_kernel<float><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<float *>(data), row_stride, pol_stride, width, height, num_pols,
            (float) value);
    else
        fill_kernel<double><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<double *>(data), row_stride, pol_stride, width, height, num_pols, value);
    KIB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
