// Imaging-weight kernels for sm_100a.
//
// Replaces GridWeights._run (reference katsdpimager/weight.py:155-176, grid_weights.mako),
// MeanWeight._run (:357-376, mean_weight.mako) and DensityWeights._run (:261-284,
// density_weights.mako).  Numerics follow WeightsHost (weight.py:541-605): the density
// weight is 1 / (a*W + b) with the multiply and add rounded separately in float32 (numpy),
// and 0 where W == 0.  Statistics are accumulated in double precision (the host uses
// numpy pairwise float32 sums; both are within 1e-6 of the exact value).
#include "kib_common.cuh"

namespace kib {

__global__ void __launch_bounds__(256)
grid_weights_kernel(float *__restrict__ grid, int row_stride, long long pol_stride,
                    int width, int height, const short4 *__restrict__ uv,
                    const float *__restrict__ weights, int P, long long num_vis)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_vis) return;
    const short4 c = uv[i];
    const int u = c.x + width / 2, v = c.y + height / 2;
    if (u < 0 || u >= width || v < 0 || v >= height) return;
    const long long addr = (long long) v * row_stride + u;
    for (int p = 0; p < P; p++)
        atomicAdd(grid + p * pol_stride + addr, weights[i * P + p]);
}

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) v += __shfl_xor_sync(0xffffffffu, v, offset);
    return v;
}

// Block-level sum of N running totals followed by one atomicAdd per total.
template <int N>
__device__ __forceinline__ void block_accumulate(double (&v)[N], double *sums)
{
    __shared__ double scratch[N][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; k++) {
        v[k] = warp_sum(v[k]);
        if (lane == 0) scratch[k][warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double total = 0;
        for (int w = 0; w < (int) (blockDim.x >> 5); w++) total += scratch[threadIdx.x][w];
        if (total != 0) atomicAdd(sums + threadIdx.x, total);
    }
}

__global__ void __launch_bounds__(256)
mean_weight_kernel(const float *__restrict__ grid, int row_stride, int width, int height,
                   double *__restrict__ sums)
{
    double v[2] = {0, 0};
    const long long total = (long long) width * height;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long) gridDim.x * blockDim.x) {
        const int x = (int) (i % width);
        const long long y = i / width;
        const double w = grid[y * row_stride + x];
        v[0] += w;
        v[1] += w * w;
    }
    block_accumulate<2>(v, sums);
}

__global__ void __launch_bounds__(256)
density_weights_kernel(float *__restrict__ grid, int row_stride, long long pol_stride,
                       int width, int height, int P, float a, float b, double *__restrict__ sums)
{
    double v[3] = {0, 0, 0};
    const long long total = (long long) width * height;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long) gridDim.x * blockDim.x) {
        const int x = (int) (i % width);
        const long long addr = (i / width) * row_stride + x;
        for (int p = 0; p < P; p++) {
            const float w = grid[p * pol_stride + addr];
            const float d = w != 0.0f ? 1.0f / __fadd_rn(__fmul_rn(a, w), b) : 0.0f;
            if (p == 0) {
                const double dw = (double) d * w;
                v[0] += w;
                v[1] += dw;
                v[2] += d * dw;
            }
            grid[p * pol_stride + addr] = d;
        }
    }
    block_accumulate<3>(v, sums);
}

// Array-of-structures visibility records -> structure-of-arrays device buffers
__global__ void __launch_bounds__(256)
unpack_records_kernel(const unsigned *__restrict__ records, int record_words, long long num_vis,
                      int P, short4 *__restrict__ uv, short *__restrict__ w_plane,
                      float *__restrict__ weights, float2 *__restrict__ vis, int vis_from_weights)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= num_vis) return;
    const unsigned *rec = records + i * record_words;
    if (uv != nullptr) {
        const unsigned a = rec[0], b = rec[1];
        uv[i] = make_short4((short) (a & 0xffff), (short) (a >> 16),
                            (short) (b & 0xffff), (short) (b >> 16));
    }
    if (w_plane != nullptr) w_plane[i] = (short) (rec[2] & 0xffff);
    for (int p = 0; p < P; p++) {
        const float wt = __uint_as_float(rec[3 + p]);
        if (weights != nullptr) weights[i * P + p] = wt;
        if (vis != nullptr) {
            vis[i * P + p] = vis_from_weights
                ? make_float2(wt, 0.0f)
                : make_float2(__uint_as_float(rec[3 + P + 2 * p]),
                              __uint_as_float(rec[4 + P + 2 * p]));
        }
    }
}

static int reduction_blocks(long long total)
{
    long long blocks = (total + 256 * 8 - 1) / (256 * 8);
    const long long max_blocks = (long long) sm_count() * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    return blocks < 1 ? 1 : (int) blocks;
}

}  // namespace kib

using namespace kib;

extern "C" {

int kib_grid_weights(float *grid, int row_stride, int64_t pol_stride, int width, int height,
                     const int16_t *uv, const float *weights, int num_pols, int64_t num_vis,
                     kib_stream_t stream)
{
    KIB_REQUIRE(num_pols >= 1 && num_pols <= 4, "kib_grid_weights: num_pols must be 1..4");
    KIB_REQUIRE(width % 2 == 0 && height % 2 == 0, "kib_grid_weights: odd-sized grid");
    if (num_vis <= 0) return 0;
    const unsigned blocks = (unsigned) ((num_vis + 255) / 256);
    grid_weights_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
        grid, row_stride, pol_stride, width, height, reinterpret_cast<const short4 *>(uv),
        weights, num_pols, num_vis);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_unpack_records(const void *records, int record_bytes, int64_t num_vis, int num_pols,
                       int16_t *uv, int16_t *w_plane, float *weights, void *vis,
                       int vis_from_weights, kib_stream_t stream)
{
    KIB_REQUIRE(num_pols >= 1 && num_pols <= 4, "kib_unpack_records: num_pols must be 1..4");
    KIB_REQUIRE(record_bytes == 12 + 12 * num_pols,
                "kib_unpack_records: record size %d does not match %d polarizations",
                record_bytes, num_pols);
    if (num_vis <= 0) return 0;
    const unsigned blocks = (unsigned) ((num_vis + 255) / 256);
    unpack_records_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
        static_cast<const unsigned *>(records), record_bytes / 4, num_vis, num_pols,
        reinterpret_cast<short4 *>(uv), w_plane, weights, static_cast<float2 *>(vis),
        vis_from_weights);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_mean_weight(const float *grid, int row_stride, int width, int height,
                    double *sums, kib_stream_t stream)
{
    KIB_REQUIRE(sums != nullptr, "kib_mean_weight: null sums");
    if (width <= 0 || height <= 0) return 0;
    mean_weight_kernel<<<reduction_blocks((long long) width * height), 256, 0, as_stream(stream)>>>(
        grid, row_stride, width, height, sums);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_density_weights(float *grid, int row_stride, int64_t pol_stride, int width, int height,
                        int num_pols, float a, float b, double *sums, kib_stream_t stream)
{
    KIB_REQUIRE(sums != nullptr, "kib_density_weights: null sums");
    KIB_REQUIRE(num_pols >= 1 && num_pols <= 4, "kib_density_weights: num_pols must be 1..4");
    if (width <= 0 || height <= 0) return 0;
    density_weights_kernel<<<reduction_blocks((long long) width * height), 256, 0,
                             as_stream(stream)>>>(
        grid, row_stride, pol_stride, width, height, num_pols, a, b, sums);
    KIB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
