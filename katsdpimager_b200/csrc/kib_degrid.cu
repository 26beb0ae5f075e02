// W-projection degridder (prediction from a model grid) for sm_100a.
//
// Replaces Degridder._run (reference katsdpimager/grid.py:986-1029) and
// imager_kernels/degrid.mako; numerics follow the host oracle `_degrid`
// (grid.py:1139-1154): taps are accumulated in j (v) major, k (u) minor order.
//
// Design: one thread per visibility.  Degridding is a gather, so no atomics or
// cross-thread reductions are needed at all when a thread owns a whole
// visibility; consecutive threads hold consecutive samples of a baseline track,
// whose footprints overlap almost entirely, so a warp's 32 gathers of tap (j,k)
// fall in one or two 128-byte lines and are served by L1.  The separable kernel
// row for v is hoisted out of the inner loop; all P polarizations share the
// weight product.
#include "kib_common.cuh"

namespace kib {

struct DegridParams {
    const void *grid;
    const short4 *uv;
    const short *w_plane;
    const float *weights;
    float2 *vis;
    const float2 *lut;
    int32_t *num_rejected;
    long long grid_pol_stride;
    long long num_vis;
    int grid_row_stride;
    int grid_size;
    int lut_slice_stride;
    int lut_w_stride;
    int lut_tap_offset;
    int w_planes;
    int oversample;
    int kernel_width;
    int uv_bias;
};

template <typename Real> struct Cplx;
template <> struct Cplx<float> { typedef float2 type; };
template <> struct Cplx<double> { typedef double2 type; };

template <typename Real, int P>
__global__ void __launch_bounds__(128)
degrid_kernel(const DegridParams prm)
{
    typedef typename Cplx<Real>::type Complex;
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.num_vis) return;
    const int K = prm.kernel_width;
    const short4 c = prm.uv[i];
    const int w = prm.w_plane[i];
    const int u0 = c.x - prm.uv_bias;
    const int v0 = c.y - prm.uv_bias;
    const bool ok = u0 >= 0 && v0 >= 0 && u0 + K <= prm.grid_size && v0 + K <= prm.grid_size
                    && w >= 0 && w < prm.w_planes
                    && c.z >= 0 && c.z < prm.oversample && c.w >= 0 && c.w < prm.oversample;
    if (!ok) {
        if (prm.num_rejected != nullptr) atomicAdd(prm.num_rejected, 1);
        return;
    }
    const float2 *ku = prm.lut + (long long) w * prm.lut_w_stride + c.z * prm.lut_slice_stride
                       + prm.lut_tap_offset;
    const float2 *kv = prm.lut + (long long) w * prm.lut_w_stride + c.w * prm.lut_slice_stride
                       + prm.lut_tap_offset;
    const Complex *base = static_cast<const Complex *>(prm.grid)
                          + (long long) v0 * prm.grid_row_stride + u0;
    Complex sum[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
        sum[p].x = 0;
        sum[p].y = 0;
    }
    for (int j = 0; j < K; j++) {
        const float2 wv = __ldg(kv + j);
        const Complex *row = base + (long long) j * prm.grid_row_stride;
#pragma unroll 4
        for (int k = 0; k < K; k++) {
            const float2 wu = __ldg(ku + k);
            // weight = lut_v[j] * lut_u[k], no conjugate (grid.py:1150)
            const Real wr = (Real) (wv.x * wu.x - wv.y * wu.y);
            const Real wi = (Real) (wv.x * wu.y + wv.y * wu.x);
#pragma unroll
            for (int p = 0; p < P; p++) {
                const Complex g = __ldg(row + p * prm.grid_pol_stride + k);
                sum[p].x = fma(wr, g.x, fma(-wi, g.y, sum[p].x));
                sum[p].y = fma(wr, g.y, fma(wi, g.x, sum[p].y));
            }
        }
    }
#pragma unroll
    for (int p = 0; p < P; p++) {
        const float wt = prm.weights[i * P + p];
        float2 v = prm.vis[i * P + p];
        v.x = (float) ((Real) v.x - (Real) wt * sum[p].x);
        v.y = (float) ((Real) v.y - (Real) wt * sum[p].y);
        prm.vis[i * P + p] = v;
    }
}

template <typename Real>
static int launch_degrid(const DegridParams &prm, int P, cudaStream_t stream)
{
    const int threads = 128;
    const unsigned blocks = (unsigned) ((prm.num_vis + threads - 1) / threads);
    switch (P) {
    case 1: degrid_kernel<Real, 1><<<blocks, threads, 0, stream>>>(prm); break;
    case 2: degrid_kernel<Real, 2><<<blocks, threads, 0, stream>>>(prm); break;
    case 3: degrid_kernel<Real, 3><<<blocks, threads, 0, stream>>>(prm); break;
    case 4: degrid_kernel<Real, 4><<<blocks, threads, 0, stream>>>(prm); break;
    default:
        set_error("kib_degrid: num_pols must be 1..4, not %d", P);
        return -1;
    }
    KIB_CHECK_LAUNCH();
    return 0;
}

}  // namespace kib

using namespace kib;

extern "C" int kib_degrid(const void *grid, int grid_row_stride, int64_t grid_pol_stride,
                          int grid_size, int dtype,
                          const int16_t *uv, const int16_t *w_plane, const float *weights,
                          void *vis,
                          const void *lut, int lut_slice_stride, int lut_tap_offset,
                          int w_planes, int oversample, int kernel_width, int num_pols,
                          int64_t num_vis, int32_t *num_rejected, kib_stream_t stream)
{
    KIB_REQUIRE(dtype == KIB_F32 || dtype == KIB_F64, "kib_degrid: bad dtype %d", dtype);
    KIB_REQUIRE(num_vis >= 0, "kib_degrid: negative num_vis");
    KIB_REQUIRE(kernel_width >= 1 && kernel_width <= grid_size,
                "kib_degrid: kernel width %d does not fit grid %d", kernel_width, grid_size);
    KIB_REQUIRE(w_planes >= 1 && oversample >= 1, "kib_degrid: bad LUT shape");
    KIB_REQUIRE(lut_slice_stride >= kernel_width + lut_tap_offset && lut_tap_offset >= 0,
                "kib_degrid: bad LUT strides");
    KIB_REQUIRE(grid_size % 2 == 0, "kib_degrid: odd grid size %d", grid_size);
    if (num_vis == 0) return 0;
    DegridParams prm;
    prm.grid = grid;
    prm.uv = reinterpret_cast<const short4 *>(uv);
    prm.w_plane = w_plane;
    prm.weights = weights;
    prm.vis = static_cast<float2 *>(vis);
    prm.lut = static_cast<const float2 *>(lut);
    prm.num_rejected = num_rejected;
    prm.grid_pol_stride = grid_pol_stride;
    prm.num_vis = num_vis;
    prm.grid_row_stride = grid_row_stride;
    prm.grid_size = grid_size;
    prm.lut_slice_stride = lut_slice_stride;
    prm.lut_w_stride = lut_slice_stride * oversample;
    prm.lut_tap_offset = lut_tap_offset;
    prm.w_planes = w_planes;
    prm.oversample = oversample;
    prm.kernel_width = kernel_width;
    prm.uv_bias = (kernel_width - 1) / 2 - grid_size / 2;
    if (dtype == KIB_F32) return launch_degrid<float>(prm, num_pols, as_stream(stream));
    return launch_degrid<double>(prm, num_pols, as_stream(stream));
}
