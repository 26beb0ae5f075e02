// W-projection convolutional gridder for sm_100a.
//
// Replaces Gridder._run/static_run (reference katsdpimager/grid.py:787-867) and
// imager_kernels/grid.mako; numerics follow the host oracle `_grid`
// (grid.py:1033-1052).
//
// Design (see DESIGN.md "gridder"): the work is FP32-pipe bound, so the kernel is
// organised to maximise the share of issue slots that are FFMAs.
//  * The visibility list is cut into contiguous *runs*; a *group* of TX*TY threads
//    walks one run.  Consecutive visibilities of a run come from one baseline
//    track (the preprocessor keeps baseline-major, time-minor order), so their
//    K x K footprints overlap almost completely.
//  * Every thread owns MX column slots and MY row slots of the footprint with a
//    *cyclic* assignment: column slot s always holds the unique grid column
//    c >= u0 with c == s (mod BX), BX = MX*TX >= K.  A thread therefore keeps
//    MX*MY*P complex accumulators in registers, and an accumulator is only
//    flushed (vector red.global.add) when the footprint moves past its cell.
//    Unlike the reference's block-aligned bins no zero padding of the
//    convolution kernel is needed when MX*TX == K, and BX need not be a power
//    of two (the residues u0 mod BX are computed once per visibility by the
//    loading thread).
//  * Visibility records (footprint origin, LUT row offsets, weighted samples)
//    are staged in shared memory in batches by the group itself, so the inner
//    loop reads them with two LDS.128 + P/2 LDS.128 and fetches its MX+MY
//    convolution taps from the read-only LUT (L1 resident for small kernels).
#include "kib_common.cuh"

namespace kib {

constexpr int GRID_BATCH = 16;          // visibilities staged per group per batch
constexpr int GRID_MAX_GROUP = 512;      // largest thread group (one visibility stream)
constexpr int GRID_MAX_GROUPS_PER_BLOCK = 32;

struct GridParams {
    void *grid;
    const float *weights_grid;
    const short4 *uv;
    const short *w_plane;
    const float2 *vis;
    const float2 *lut;
    int32_t *num_rejected;
    long long grid_pol_stride;
    long long weights_pol_stride;
    long long num_vis;
    int grid_row_stride;
    int grid_size;
    int weights_row_stride;
    int lut_slice_stride;    // elements between sub-pixel rows
    int lut_w_stride;        // elements between w planes
    int lut_tap_offset;
    int w_planes;
    int oversample;
    int kernel_width;
    int uv_bias;             // (K-1)/2 - G/2
    int half_grid;           // G/2
    int tx, ty;              // threads per group in x / y
    int bx, by;              // MX*TX, MY*TY
    int group_size;          // TX*TY
    int groups_per_block;
    int run;                 // visibilities per group
};

template <typename Real> struct Acc;
template <> struct Acc<float> {
    typedef float2 type;
    static __device__ __forceinline__ void flush(float2 *addr, float2 v) { atomicAdd(addr, v); }
};
template <> struct Acc<double> {
    typedef double2 type;
    static __device__ __forceinline__ void flush(double2 *addr, double2 v)
    {
        atomicAdd(&addr->x, v.x);
        atomicAdd(&addr->y, v.y);
    }
};

template <typename Real, int P, int MX, int MY, int MAXT>
__global__ void __launch_bounds__(MAXT)
grid_kernel(const GridParams prm)
{
    typedef typename Acc<Real>::type Complex;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // Layout: int4 header[groups][BATCH]; int residues[groups][BATCH]; float2 sample[groups][BATCH][P]
    const int gpb = prm.groups_per_block;
    int4 *hdr_all = reinterpret_cast<int4 *>(smem_raw);
    float2 *sample_all = reinterpret_cast<float2 *>(hdr_all + gpb * GRID_BATCH);
    int *res_all = reinterpret_cast<int *>(sample_all + gpb * GRID_BATCH * P);

    const int tid = threadIdx.x;
    const int g = tid / prm.group_size;
    const int q = tid - g * prm.group_size;
    const int ty = q / prm.tx;
    const int tx = q - ty * prm.tx;
    int4 *hdr = hdr_all + g * GRID_BATCH;
    float2 *sample = sample_all + g * GRID_BATCH * P;
    int *res = res_all + g * GRID_BATCH;

    const int K = prm.kernel_width;
    const int BX = prm.bx, BY = prm.by;
    const int G = prm.grid_size;

    const long long group_id = (long long) blockIdx.x * gpb + g;
    const long long run_start = group_id * prm.run;
    long long run_end = run_start + prm.run;
    if (run_end > prm.num_vis) run_end = prm.num_vis;

    Complex acc[MY][MX][P];
    int cur_col[MX], cur_row[MY];
#pragma unroll
    for (int j = 0; j < MY; j++)
#pragma unroll
        for (int i = 0; i < MX; i++)
#pragma unroll
            for (int p = 0; p < P; p++) {
                acc[j][i][p].x = 0;
                acc[j][i][p].y = 0;
            }
#pragma unroll
    for (int i = 0; i < MX; i++) cur_col[i] = -1;
#pragma unroll
    for (int j = 0; j < MY; j++) cur_row[j] = -1;

    Complex *const grid = static_cast<Complex *>(prm.grid);
    int rejected = 0;

    for (int batch = 0; batch < prm.run; batch += GRID_BATCH) {
        const long long batch_start = run_start + batch;
        // ---- load phase: the group stages its next GRID_BATCH visibilities
        for (int e = q; e < GRID_BATCH; e += prm.group_size) {
            const long long idx = batch_start + e;
            if (idx < run_end) {
                const short4 c = prm.uv[idx];
                int w = prm.w_plane[idx];
                int u0 = c.x - prm.uv_bias;
                int v0 = c.y - prm.uv_bias;
                bool ok = u0 >= 0 && v0 >= 0 && u0 + K <= G && v0 + K <= G
                          && w >= 0 && w < prm.w_planes
                          && c.z >= 0 && c.z < prm.oversample && c.w >= 0 && c.w < prm.oversample;
                if (!ok) {
                    // Out-of-range visibility: contributes nothing (the reference rejects
                    // such data up front, grid.py:753-761).  Keep coordinates harmless.
                    u0 = 0; v0 = 0; w = 0;
                    rejected++;
                }
                const int lut_u = w * prm.lut_w_stride + (ok ? c.z : 0) * prm.lut_slice_stride
                                  + prm.lut_tap_offset;
                const int lut_v = w * prm.lut_w_stride + (ok ? c.w : 0) * prm.lut_slice_stride
                                  + prm.lut_tap_offset;
                hdr[e] = make_int4(u0, v0, lut_u, lut_v);
                res[e] = (u0 % BX) | ((v0 % BY) << 16);
                const long long waddr = (long long) (c.y + prm.half_grid) * prm.weights_row_stride
                                        + (c.x + prm.half_grid);
#pragma unroll
                for (int p = 0; p < P; p++) {
                    float2 v = make_float2(0.0f, 0.0f);
                    if (ok) {
                        const float wt = __ldg(prm.weights_grid + p * prm.weights_pol_stride + waddr);
                        v = __ldg(prm.vis + idx * P + p);
                        v.x *= wt;
                        v.y *= wt;
                    }
                    sample[e * P + p] = v;
                }
            }
        }
        __syncthreads();

        long long remaining = run_end - batch_start;
        const int count = remaining >= GRID_BATCH ? GRID_BATCH : (remaining > 0 ? (int) remaining : 0);
        // ---- process phase
        for (int e = 0; e < count; e++) {
            const int4 h = hdr[e];
            const int rr = res[e];
            const int ru = rr & 0xffff;
            const int rv = rr >> 16;
            float2 s[P];
#pragma unroll
            for (int p = 0; p < P; p++) s[p] = sample[e * P + p];

            float2 wu[MX], wv[MY];
            int col[MX], row[MY];
            bool changed = false;
#pragma unroll
            for (int i = 0; i < MX; i++) {
                int d = tx + i * prm.tx - ru;
                if (d < 0) d += BX;
                col[i] = h.x + d;
                wu[i] = d < K ? __ldg(prm.lut + h.z + d) : make_float2(0.0f, 0.0f);
                changed |= col[i] != cur_col[i];
            }
#pragma unroll
            for (int j = 0; j < MY; j++) {
                int d = ty + j * prm.ty - rv;
                if (d < 0) d += BY;
                row[j] = h.y + d;
                wv[j] = d < K ? __ldg(prm.lut + h.w + d) : make_float2(0.0f, 0.0f);
                changed |= row[j] != cur_row[j];
            }
            if (changed) {
#pragma unroll
                for (int j = 0; j < MY; j++)
#pragma unroll
                    for (int i = 0; i < MX; i++) {
                        if (col[i] != cur_col[i] || row[j] != cur_row[j]) {
                            if (cur_col[i] >= 0 && cur_row[j] >= 0 && cur_col[i] < G && cur_row[j] < G) {
                                Complex *ptr = grid + (long long) cur_row[j] * prm.grid_row_stride
                                               + cur_col[i];
#pragma unroll
                                for (int p = 0; p < P; p++)
                                    Acc<Real>::flush(ptr + p * prm.grid_pol_stride, acc[j][i][p]);
                            }
#pragma unroll
                            for (int p = 0; p < P; p++) {
                                acc[j][i][p].x = 0;
                                acc[j][i][p].y = 0;
                            }
                        }
                    }
#pragma unroll
                for (int i = 0; i < MX; i++) cur_col[i] = col[i];
#pragma unroll
                for (int j = 0; j < MY; j++) cur_row[j] = row[j];
            }
#pragma unroll
            for (int j = 0; j < MY; j++)
#pragma unroll
                for (int i = 0; i < MX; i++) {
                    // weight = lut_v[j] * lut_u[i]; the grid receives sample * conj(weight)
                    // (grid.py:1049-1052).
                    float2 wgt;
                    wgt.x = wv[j].x * wu[i].x - wv[j].y * wu[i].y;
                    wgt.y = wv[j].x * wu[i].y + wv[j].y * wu[i].x;
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        acc[j][i][p].x = fma((Real) s[p].x, (Real) wgt.x,
                                             fma((Real) s[p].y, (Real) wgt.y, acc[j][i][p].x));
                        acc[j][i][p].y = fma((Real) s[p].y, (Real) wgt.x,
                                             fma(-(Real) s[p].x, (Real) wgt.y, acc[j][i][p].y));
                    }
                }
        }
        __syncthreads();
    }

    // ---- final flush
#pragma unroll
    for (int j = 0; j < MY; j++)
#pragma unroll
        for (int i = 0; i < MX; i++) {
            if (cur_col[i] >= 0 && cur_row[j] >= 0 && cur_col[i] < G && cur_row[j] < G) {
                Complex *ptr = grid + (long long) cur_row[j] * prm.grid_row_stride + cur_col[i];
#pragma unroll
                for (int p = 0; p < P; p++)
                    Acc<Real>::flush(ptr + p * prm.grid_pol_stride, acc[j][i][p]);
            }
        }
    if (rejected != 0 && prm.num_rejected != nullptr) atomicAdd(prm.num_rejected, rejected);
}

struct GridConfig {
    int mx, my, tx, ty;
};

// Register budget: accumulators use MX*MY*P*2 (*2 for double) 32-bit registers.
constexpr int GRID_ACC_BUDGET = 64;
static const int kMxOptions[] = {4, 5, 6, 7, 8};
static const int kMyOptions[] = {1, 2, 4};

static bool config_allowed(int mx, int my, int P, int dtype)
{
    return mx * my * P * 2 * (dtype == KIB_F64 ? 2 : 1) <= GRID_ACC_BUDGET;
}

static bool choose_config(int K, int P, int dtype, GridConfig *out)
{
    double best_score = 1e30;
    bool found = false;
    for (int mx : kMxOptions)
        for (int my : kMyOptions) {
            if (!config_allowed(mx, my, P, dtype)) continue;
            int tx = (K + mx - 1) / mx;
            int ty = (K + my - 1) / my;
            if (tx * ty > GRID_MAX_GROUP) continue;
            double waste = (double) (mx * tx) * (my * ty) / ((double) K * K);
            // Loads per cell (mx + my LUT fetches amortised over mx*my cells), and a mild
            // preference for more cells per thread (amortises the record fetch).
            double overhead = (double) (mx + my + 6) / (mx * my * (4.0 + 4.0 * P));
            double score = waste * (1.0 + overhead);
            if (score < best_score - 1e-9) {
                best_score = score;
                out->mx = mx; out->my = my; out->tx = tx; out->ty = ty;
                found = true;
            }
        }
    return found;
}

template <typename Real, int P, int MX, int MY>
static int launch_grid(GridParams &prm, cudaStream_t stream)
{
    const int group = prm.tx * prm.ty;
    int gpb = 256 / group;
    if (gpb < 1) gpb = 1;
    if (gpb > GRID_MAX_GROUPS_PER_BLOCK) gpb = GRID_MAX_GROUPS_PER_BLOCK;
    prm.group_size = group;
    prm.groups_per_block = gpb;
    const int threads = gpb * group;
    // Enough groups to fill the machine a few times over, but runs long enough
    // that the final flush (MX*MY*P reds per thread) stays negligible.
    const long long blocks_per_sm = 2048 / (threads < 64 ? 64 : threads);
    const long long target_groups = (long long) sm_count() * 2 * blocks_per_sm * gpb;
    long long run = (prm.num_vis + target_groups - 1) / target_groups;
    if (run < 4 * GRID_BATCH) run = 4 * GRID_BATCH;
    if (run > 8192) run = 8192;
    run = (run + GRID_BATCH - 1) / GRID_BATCH * GRID_BATCH;
    prm.run = (int) run;
    const long long groups = (prm.num_vis + run - 1) / run;
    const long long blocks = (groups + gpb - 1) / gpb;
    const size_t smem = (size_t) gpb * GRID_BATCH * (sizeof(int4) + sizeof(int) + P * sizeof(float2));
    if (threads <= 256)
        grid_kernel<Real, P, MX, MY, 256><<<(unsigned) blocks, threads, smem, stream>>>(prm);
    else
        grid_kernel<Real, P, MX, MY, GRID_MAX_GROUP><<<(unsigned) blocks, threads, smem, stream>>>(prm);
    KIB_CHECK_LAUNCH();
    return 0;
}

template <typename Real, int P, int MX, int MY>
constexpr bool kernel_allowed()
{
    return MX * MY * P * 2 * (int) (sizeof(Real) / 4) <= GRID_ACC_BUDGET;
}

template <typename Real, int P, int MX, int MY>
static int launch_if_allowed(GridParams &prm, cudaStream_t stream)
{
    if constexpr (kernel_allowed<Real, P, MX, MY>()) {
        return launch_grid<Real, P, MX, MY>(prm, stream);
    } else {
        set_error("kib_grid: no kernel for MX=%d MY=%d P=%d", MX, MY, P);
        return -1;
    }
}

template <typename Real, int P, int MX>
static int dispatch_my(GridParams &prm, int my, int dtype, cudaStream_t stream)
{
    switch (my) {
    case 1: return launch_if_allowed<Real, P, MX, 1>(prm, stream);
    case 2: return launch_if_allowed<Real, P, MX, 2>(prm, stream);
    case 4: return launch_if_allowed<Real, P, MX, 4>(prm, stream);
    }
    set_error("kib_grid: no kernel for MX=%d MY=%d P=%d", MX, my, P);
    return -1;
}

template <typename Real, int P>
static int dispatch_mx(GridParams &prm, const GridConfig &cfg, int dtype, cudaStream_t stream)
{
    switch (cfg.mx) {
    case 4: return dispatch_my<Real, P, 4>(prm, cfg.my, dtype, stream);
    case 5: return dispatch_my<Real, P, 5>(prm, cfg.my, dtype, stream);
    case 6: return dispatch_my<Real, P, 6>(prm, cfg.my, dtype, stream);
    case 7: return dispatch_my<Real, P, 7>(prm, cfg.my, dtype, stream);
    case 8: return dispatch_my<Real, P, 8>(prm, cfg.my, dtype, stream);
    }
    set_error("kib_grid: unsupported MX=%d", cfg.mx);
    return -1;
}

template <typename Real>
static int dispatch_pols(GridParams &prm, const GridConfig &cfg, int P, int dtype, cudaStream_t stream)
{
    switch (P) {
    case 1: return dispatch_mx<Real, 1>(prm, cfg, dtype, stream);
    case 2: return dispatch_mx<Real, 2>(prm, cfg, dtype, stream);
    case 3: return dispatch_mx<Real, 3>(prm, cfg, dtype, stream);
    case 4: return dispatch_mx<Real, 4>(prm, cfg, dtype, stream);
    }
    set_error("kib_grid: num_pols must be 1..4, not %d", P);
    return -1;
}

}  // namespace kib

using namespace kib;

extern "C" int kib_grid(void *grid, int grid_row_stride, int64_t grid_pol_stride, int grid_size,
                        int dtype,
                        const float *weights_grid, int weights_row_stride,
                        int64_t weights_pol_stride,
                        const int16_t *uv, const int16_t *w_plane, const void *vis,
                        const void *lut, int lut_slice_stride, int lut_tap_offset,
                        int w_planes, int oversample, int kernel_width, int num_pols,
                        int64_t num_vis, int32_t *num_rejected, kib_stream_t stream)
{
    KIB_REQUIRE(dtype == KIB_F32 || dtype == KIB_F64, "kib_grid: bad dtype %d", dtype);
    KIB_REQUIRE(num_vis >= 0, "kib_grid: negative num_vis");
    KIB_REQUIRE(kernel_width >= 1 && kernel_width <= grid_size,
                "kib_grid: kernel width %d does not fit grid %d", kernel_width, grid_size);
    KIB_REQUIRE(w_planes >= 1 && oversample >= 1, "kib_grid: bad LUT shape");
    KIB_REQUIRE(lut_slice_stride >= kernel_width + lut_tap_offset && lut_tap_offset >= 0,
                "kib_grid: bad LUT strides");
    KIB_REQUIRE(grid_size % 2 == 0, "kib_grid: odd grid size %d", grid_size);
    if (num_vis == 0) return 0;
    GridConfig cfg;
    KIB_REQUIRE(choose_config(kernel_width, num_pols, dtype, &cfg),
                "kib_grid: no configuration for kernel width %d, %d pols", kernel_width, num_pols);
    GridParams prm;
    prm.grid = grid;
    prm.weights_grid = weights_grid;
    prm.uv = reinterpret_cast<const short4 *>(uv);
    prm.w_plane = w_plane;
    prm.vis = static_cast<const float2 *>(vis);
    prm.lut = static_cast<const float2 *>(lut);
    prm.num_rejected = num_rejected;
    prm.grid_pol_stride = grid_pol_stride;
    prm.weights_pol_stride = weights_pol_stride;
    prm.num_vis = num_vis;
    prm.grid_row_stride = grid_row_stride;
    prm.grid_size = grid_size;
    prm.weights_row_stride = weights_row_stride;
    prm.lut_slice_stride = lut_slice_stride;
    prm.lut_w_stride = lut_slice_stride * oversample;
    prm.lut_tap_offset = lut_tap_offset;
    prm.w_planes = w_planes;
    prm.oversample = oversample;
    prm.kernel_width = kernel_width;
    prm.uv_bias = (kernel_width - 1) / 2 - grid_size / 2;
    prm.half_grid = grid_size / 2;
    prm.tx = cfg.tx;
    prm.ty = cfg.ty;
    prm.bx = cfg.mx * cfg.tx;
    prm.by = cfg.my * cfg.ty;
    cudaStream_t s = as_stream(stream);
    if (dtype == KIB_F32) return dispatch_pols<float>(prm, cfg, num_pols, dtype, s);
    return dispatch_pols<double>(prm, cfg, num_pols, dtype, s);
}
