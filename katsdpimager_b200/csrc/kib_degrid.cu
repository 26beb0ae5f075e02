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
// fall in one or two 128-byte lines and are served by L1.  The L1 data pipe
// delivers 128 bytes per clock whether or not the lanes' addresses coincide, so
// every kernel here is bound by the number of bytes its lanes load per FMA: the
// routes below differ in how few loads they issue besides the K x K x P grid
// cells (taps in registers instead of one look-up per cell).  A register-cached
// mirror of the gridder (8-lane groups holding their footprint rows, partial sums
// reduced by shuffles) was built and measured at 0.61 ms against 0.31 ms for the
// kernel below on the dense W slice of config 2 (profiles/r02_degrid.md); it is
// in the history (commit 816e985), not in the library.
#include "kib_common.cuh"
#include <cstdlib>

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

// KT > 0: the support is KT (compile time): loops unrolled, the u taps held in registers.
template <typename Real, int P, int KT>
__global__ void __launch_bounds__(128)
degrid_kernel(const DegridParams prm)
{
    typedef typename Cplx<Real>::type Complex;
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.num_vis) return;
    const int K = KT > 0 ? KT : prm.kernel_width;
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
    if (KT > 0) {
        float2 wu[KT > 0 ? KT : 1];
#pragma unroll
        for (int k = 0; k < KT; k++) wu[k] = __ldg(ku + k);
#pragma unroll
        for (int j = 0; j < KT; j++) {
            const float2 wv = __ldg(kv + j);
            const Complex *row = base + (long long) j * prm.grid_row_stride;
#pragma unroll
            for (int k = 0; k < KT; k++) {
                // weight = lut_v[j] * lut_u[k], no conjugate (grid.py:1150)
                const Real wr = (Real) (wv.x * wu[k].x - wv.y * wu[k].y);
                const Real wi = (Real) (wv.x * wu[k].y + wv.y * wu[k].x);
#pragma unroll
                for (int p = 0; p < P; p++) {
                    const Complex g = __ldg(row + p * prm.grid_pol_stride + k);
                    sum[p].x = fma(wr, g.x, fma(-wi, g.y, sum[p].x));
                    sum[p].y = fma(wr, g.y, fma(wi, g.x, sum[p].y));
                }
            }
        }
    } else {
        for (int j = 0; j < K; j++) {
            const float2 wv = __ldg(kv + j);
            const Complex *row = base + (long long) j * prm.grid_row_stride;
#pragma unroll 4
            for (int k = 0; k < K; k++) {
                const float2 wu = __ldg(ku + k);
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

// =====================================================================================
// Vector-load route (float32): one thread per visibility like degrid_kernel, but
//  * the grid is read as aligned *pairs* of cells (LDG.128): the footprint row is widened to
//    the even column below u0, and the u taps are shifted by that parity, so the number of
//    gather instructions per tap halves (the scalar kernel keeps the L1 data pipe 85 % busy);
//  * the u taps live in registers in blocks of TB (all 8 for K <= 8, 32 at a time for wide
//    supports) instead of being fetched again for every footprint row;
//  * a row is summed with the u taps first and multiplied by its v tap once
//    (4 P FMAs per cell instead of 4 P + 4).
// Needs a 16-byte aligned grid with even row and plane strides.
template <int P, int TB, bool WIDE>
__global__ void __launch_bounds__(128)
degrid_vec_kernel(const DegridParams prm)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= prm.num_vis) return;
    const int K = prm.kernel_width, G = prm.grid_size;
    const short4 c = prm.uv[i];
    const int w = prm.w_plane[i];
    const int u0 = c.x - prm.uv_bias;
    const int v0 = c.y - prm.uv_bias;
    const bool ok = u0 >= 0 && v0 >= 0 && u0 + K <= G && v0 + K <= G
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
    const int par = u0 & 1;
    const int taps = K + par;               // taps in shifted coordinates k' = k + par
    // pair (v0, (u0 - par) / 2) of plane 0
    const float4 *base = static_cast<const float4 *>(prm.grid)
                         + (((long long) v0 * prm.grid_row_stride + (u0 - par)) >> 1);
    const long long row_pairs = prm.grid_row_stride >> 1;
    const long long pol_pairs = prm.grid_pol_stride >> 1;
    float2 sum[P];
#pragma unroll
    for (int p = 0; p < P; p++) sum[p] = make_float2(0.0f, 0.0f);
#pragma unroll 1
    for (int kb = 0; kb < taps; kb += TB) {
        float2 wu[TB];
#pragma unroll
        for (int q = 0; q < TB; q++) {
            const int k = kb + q - par;
            wu[q] = (k >= 0 && k < K) ? __ldg(ku + k) : make_float2(0.0f, 0.0f);
        }
        const float4 *blk = base + (kb >> 1);
#pragma unroll 1
        for (int j = 0; j < K; j++) {
            const float2 wv = __ldg(kv + j);
            const float4 *row = blk + j * row_pairs;
            float2 part[P];
#pragma unroll
            for (int p = 0; p < P; p++) part[p] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int q = 0; q < TB; q += 2) {
                if (kb + q < taps) {         // the pair holds a tap: its columns are inside the grid
                    float4 g[P];
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        if (WIDE) {
                            g[p] = __ldg(row + p * pol_pairs + (q >> 1));
                        } else {
                            const float2 *cells
                                = reinterpret_cast<const float2 *>(row + p * pol_pairs + (q >> 1));
                            const float2 a = __ldg(cells), b = __ldg(cells + 1);
                            g[p] = make_float4(a.x, a.y, b.x, b.y);
                        }
                    }
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        part[p].x = fmaf(wu[q].x, g[p].x, fmaf(-wu[q].y, g[p].y, part[p].x));
                        part[p].y = fmaf(wu[q].x, g[p].y, fmaf(wu[q].y, g[p].x, part[p].y));
                        part[p].x = fmaf(wu[q + 1].x, g[p].z, fmaf(-wu[q + 1].y, g[p].w, part[p].x));
                        part[p].y = fmaf(wu[q + 1].x, g[p].w, fmaf(wu[q + 1].y, g[p].z, part[p].y));
                    }
                }
            }
            // weight = lut_v[j] * lut_u[k], no conjugate (grid.py:1150)
#pragma unroll
            for (int p = 0; p < P; p++) {
                sum[p].x = fmaf(wv.x, part[p].x, fmaf(-wv.y, part[p].y, sum[p].x));
                sum[p].y = fmaf(wv.x, part[p].y, fmaf(wv.y, part[p].x, sum[p].y));
            }
        }
    }
#pragma unroll
    for (int p = 0; p < P; p++) {
        const float wt = prm.weights[i * P + p];
        float2 v = prm.vis[i * P + p];
        v.x = v.x - wt * sum[p].x;
        v.y = v.y - wt * sum[p].y;
        prm.vis[i * P + p] = v;
    }
}

template <int P>
static int launch_degrid_vec(const DegridParams &prm, bool wide, cudaStream_t stream)
{
    const int threads = 128;
    const unsigned blocks = (unsigned) ((prm.num_vis + threads - 1) / threads);
    if (prm.kernel_width <= 8) {
        if (wide) degrid_vec_kernel<P, 8, true><<<blocks, threads, 0, stream>>>(prm);
        else degrid_vec_kernel<P, 8, false><<<blocks, threads, 0, stream>>>(prm);
    } else {
        if (wide) degrid_vec_kernel<P, 32, true><<<blocks, threads, 0, stream>>>(prm);
        else degrid_vec_kernel<P, 32, false><<<blocks, threads, 0, stream>>>(prm);
    }
    KIB_CHECK_LAUNCH();
    return 0;
}

// Returns 1 when the launch was not taken (caller falls through to the scalar kernel).
static int try_degrid_vec(const DegridParams &prm, int P, bool wide, cudaStream_t stream)
{
    if ((reinterpret_cast<size_t>(prm.grid) & 15) != 0 || (prm.grid_row_stride & 1) != 0
        || (prm.grid_pol_stride & 1) != 0 || (prm.grid_size & 1) != 0)
        return 1;
    switch (P) {
    case 1: return launch_degrid_vec<1>(prm, wide, stream);
    case 2: return launch_degrid_vec<2>(prm, wide, stream);
    case 3: return launch_degrid_vec<3>(prm, wide, stream);
    case 4: return launch_degrid_vec<4>(prm, wide, stream);
    }
    return 1;
}

template <typename Real>
static int launch_degrid(const DegridParams &prm, int P, cudaStream_t stream)
{
    const int threads = 128;
    const unsigned blocks = (unsigned) ((prm.num_vis + threads - 1) / threads);
    const char *route = getenv("KIB_DEGRID_ROUTE");
    const bool k7 = prm.kernel_width == 7 && !(route && route[0] == 't' && route[1] == 'h'
                                               && route[2] == 'r' && route[3] == 'e'
                                               && route[4] == 'a' && route[5] == 'd'
                                               && route[6] == '0');
#define KIB_DEGRID(PP)                                                                     \
    do {                                                                                    \
        if (k7) degrid_kernel<Real, PP, 7><<<blocks, threads, 0, stream>>>(prm);            \
        else degrid_kernel<Real, PP, 0><<<blocks, threads, 0, stream>>>(prm);               \
    } while (0)
    switch (P) {
    case 1: KIB_DEGRID(1); break;
    case 2: KIB_DEGRID(2); break;
    case 3: KIB_DEGRID(3); break;
    case 4: KIB_DEGRID(4); break;
    default:
        set_error("kib_degrid: num_pols must be 1..4, not %d", P);
        return -1;
    }
#undef KIB_DEGRID
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
    if (dtype == KIB_F32) {
        // Measured on B200 (profiles/r02_degrid.md): supports up to 8 run the scalar kernel (K = 7
        // specialised: taps in registers); wider supports keep a block of u taps in registers and
        // read the grid as aligned 16-byte pairs when there is one polarization, as 8-byte cells
        // otherwise.  KIB_DEGRID_ROUTE = thread | vec | hoist overrides the choice (probes).
        const char *route = getenv("KIB_DEGRID_ROUTE");
        const bool pick_vec = route ? (route[0] == 'v' || route[0] == 'h') : kernel_width > 8;
        if (pick_vec) {
            const bool wide = route ? route[0] == 'v' : num_pols == 1;
            const int rc = try_degrid_vec(prm, num_pols, wide, as_stream(stream));
            if (rc != 1) return rc;
        }
        return launch_degrid<float>(prm, num_pols, as_stream(stream));
    }
    return launch_degrid<double>(prm, num_pols, as_stream(stream));
}
