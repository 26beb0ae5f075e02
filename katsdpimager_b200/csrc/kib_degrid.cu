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
// Fast path (float32, K <= 8): the mirror image of the gridder's register scheme.
//
// A *group* of 8 lanes walks a contiguous run of visibilities (one baseline track, mostly).
// Lane t owns row slot t of the footprint and all MX >= K column slots of that row with the
// gridder's cyclic assignment (column slot s always holds the grid column c >= u0 with
// c == s mod MX, row slot t the row r >= v0 with r == t mod 8), and keeps those MX x P grid
// cells in registers: consecutive visibilities of a track share their footprint, so a cell is
// loaded from the grid once per *move* of the footprint instead of once per visibility.  Per
// visibility a lane forms  wv[t] * sum_s wu[s] * cell[s][p]  (the taps come from doubled
// tables in shared memory, as in grid_tma_kernel) and the 2 P partial sums are reduced over the
// 8 lanes with a transposing butterfly (4 + 2 + 1 shuffles), which leaves component c of the
// visibility in lane bitrev(c): the 8 lanes then update the 8 floats of vis[i] with one
// 32-byte access.  Headers (origin, table offsets, move masks) are computed by the group
// itself, 8 visibilities at a time, and passed through shared memory.
constexpr int DG_LANES = 8;                 // lanes per group = row slots = batch length
constexpr int DG_THREADS = 256;
constexpr int DG_GROUPS = DG_THREADS / DG_LANES;
constexpr int DG_LUT_SMEM_LIMIT = 96 * 1024;

__device__ __forceinline__ unsigned dg_change_mask(int old_pos, int new_pos, int B)
{
    const int d = new_pos - old_pos;
    if (d == 0) return 0u;
    if (d >= B || -d >= B) return 0xffffu;
    int start = (d > 0 ? old_pos : new_pos) % B;
    const int count = d > 0 ? d : -d;
    unsigned mask = 0;
    for (int i = 0; i < count; i++) {
        mask |= 1u << start;
        if (++start == B) start = 0;
    }
    return mask;
}

template <int P, int MX>
__global__ void __launch_bounds__(DG_THREADS, 2)
degrid_cached_kernel(const DegridParams prm, int run)
{
    constexpr int BX = MX, BY = DG_LANES;
    constexpr int NV = P == 1 ? 2 : (P == 2 ? 4 : 8);      // reduced values (2 P rounded up)
    extern __shared__ __align__(16) unsigned char dg_smem[];
    int4 *const hdr_all = reinterpret_cast<int4 *>(dg_smem);
    float2 *const tabu = reinterpret_cast<float2 *>(hdr_all + DG_GROUPS * DG_LANES);
    const int rows = prm.w_planes * prm.oversample;
    const int vbase = rows * 2 * BX;
    const int K = prm.kernel_width, G = prm.grid_size;
    const int tid = threadIdx.x;

    // doubled tap tables: entry e of a row holds tap e mod B (zero beyond K)
    for (int e = tid; e < rows * 2 * (BX + BY); e += DG_THREADS) {
        const bool is_v = e >= vbase;
        const int f = is_v ? e - vbase : e;
        const int period = is_v ? BY : BX;
        const int row = f / (2 * period);
        int d = f - row * 2 * period;
        if (d >= period) d -= period;
        tabu[e] = d < K ? __ldg(prm.lut + (long long) row * prm.lut_slice_stride
                                + prm.lut_tap_offset + d)
                        : make_float2(0.0f, 0.0f);
    }
    __syncthreads();

    const int g = tid / DG_LANES;
    const int t = tid % DG_LANES;
    int4 *const hdr = hdr_all + g * DG_LANES;
    const unsigned my_mask = ((1u << MX) - 1u) | (1u << (16 + t));
    // component of the reduced visibility this lane ends up with
    int comp = 0;
#pragma unroll
    for (int s = 0, half = NV / 2; half >= 1; s++, half /= 2)
        if ((t >> s) & 1) comp += half;
    const bool writer = t < NV && comp < 2 * P;

    const long long group_id = (long long) blockIdx.x * DG_GROUPS + g;
    const long long run_start = group_id * run;
    long long run_end = run_start + run;
    if (run_end > prm.num_vis) run_end = prm.num_vis;

    float2 cell[MX][P];
#pragma unroll
    for (int i = 0; i < MX; i++)
#pragma unroll
        for (int p = 0; p < P; p++) cell[i][p] = make_float2(0.0f, 0.0f);
    int carry_u = 0, carry_v = 0;
    int rejected = 0;
    const float2 *const grid = static_cast<const float2 *>(prm.grid);
    float *const vis_f = reinterpret_cast<float *>(prm.vis);

    for (int b = 0; b < run; b += DG_LANES) {
        const long long batch_start = run_start + b;
        {
            // ---- header of visibility batch_start + t
            const long long idx = batch_start + t;
            const bool live = idx < run_end;
            int u0 = 0, v0 = 0, w = 0, su = 0, sv = 0;
            bool ok = false;
            if (live) {
                const short4 c = prm.uv[idx];
                w = prm.w_plane[idx];
                u0 = c.x - prm.uv_bias;
                v0 = c.y - prm.uv_bias;
                su = c.z;
                sv = c.w;
                ok = u0 >= 0 && v0 >= 0 && u0 + K <= G && v0 + K <= G
                     && w >= 0 && w < prm.w_planes
                     && su >= 0 && su < prm.oversample && sv >= 0 && sv < prm.oversample;
                if (!ok) {
                    // leaves vis untouched; harmless stand-in coordinates
                    u0 = 0; v0 = 0; w = 0; su = 0; sv = 0;
                    rejected++;
                }
            }
            int pu = __shfl_up_sync(0xffffffffu, u0, 1, DG_LANES);
            int pv = __shfl_up_sync(0xffffffffu, v0, 1, DG_LANES);
            if (t == 0) { pu = carry_u; pv = carry_v; }
            unsigned xm = 0, ym = 0;
            if (live) {
                if (b == 0 && t == 0) {
                    xm = 0xffffu;           // first visibility of the run: every cell is new
                    ym = 0xffffu;
                } else {
                    xm = dg_change_mask(pu, u0, BX);
                    ym = dg_change_mask(pv, v0, BY);
                }
            }
            const int ru = u0 % BX, rv = v0 % BY;
            const int lutu = (w * prm.oversample + su) * 2 * BX + BX - ru;
            const int lutv = vbase + (w * prm.oversample + sv) * 2 * BY + BY - rv;
            hdr[t] = make_int4(u0 | (v0 << 16), lutu | (lutv << 16), (int) (xm | (ym << 16)),
                               ru | (rv << 8) | (ok ? 0 : 0x10000));
            // origin of the last live visibility of the batch (a dead tail ends the run)
            carry_u = __shfl_sync(0xffffffffu, u0, DG_LANES - 1, DG_LANES);
            carry_v = __shfl_sync(0xffffffffu, v0, DG_LANES - 1, DG_LANES);
        }
        __syncwarp();
#pragma unroll 1
        for (int e = 0; e < DG_LANES; e++) {
            const int4 h = hdr[e];
            const long long idx = batch_start + e;
            const bool store = writer && !(h.w & 0x10000);
            float vold = 0.0f, wt = 0.0f;
            if (store) {
                vold = vis_f[idx * (2 * P) + comp];
                wt = __ldg(prm.weights + idx * P + (comp >> 1));
            }
            if ((unsigned) h.z & my_mask) {
                // ---- the footprint moved: reload the cells that now hold another grid cell
                const int u0 = h.x & 0xffff, v0 = (int) ((unsigned) h.x >> 16);
                const int ru = h.w & 0xff, rv = (h.w >> 8) & 0xff;
                int dy = t - rv;
                if (dy < 0) dy += BY;
                const int row = v0 + dy;
                const bool row_changed = ((unsigned) h.z >> (16 + t)) & 1u;
                const float2 *rp = grid + (long long) row * prm.grid_row_stride;
#pragma unroll
                for (int i = 0; i < MX; i++) {
                    if (row_changed || (((unsigned) h.z >> i) & 1u)) {
                        int dx = i - ru;
                        if (dx < 0) dx += BX;
                        const int col = u0 + dx;
                        const bool inside = row < G && col < G;
#pragma unroll
                        for (int p = 0; p < P; p++)
                            cell[i][p] = inside ? __ldg(rp + p * prm.grid_pol_stride + col)
                                                : make_float2(0.0f, 0.0f);
                    }
                }
            }
            const float2 *urow = tabu + (h.y & 0xffff);
            const float2 wv = tabu[((unsigned) h.y >> 16) + t];
            float2 sum[P];
#pragma unroll
            for (int p = 0; p < P; p++) sum[p] = make_float2(0.0f, 0.0f);
#pragma unroll
            for (int i = 0; i < MX; i++) {
                const float2 wu = urow[i];
#pragma unroll
                for (int p = 0; p < P; p++) {
                    sum[p].x = fmaf(wu.x, cell[i][p].x, fmaf(-wu.y, cell[i][p].y, sum[p].x));
                    sum[p].y = fmaf(wu.x, cell[i][p].y, fmaf(wu.y, cell[i][p].x, sum[p].y));
                }
            }
            // weight = lut_v[j] * lut_u[k], no conjugate (grid.py:1150)
            float r[NV];
#pragma unroll
            for (int i = 0; i < NV; i++) r[i] = 0.0f;
#pragma unroll
            for (int p = 0; p < P; p++) {
                r[2 * p] = wv.x * sum[p].x - wv.y * sum[p].y;
                r[2 * p + 1] = wv.x * sum[p].y + wv.y * sum[p].x;
            }
            // transposing butterfly: after step s a lane keeps the half of its values selected
            // by bit s of its index, summed with its partner's
#pragma unroll
            for (int s = 0, half = NV / 2; half >= 1; s++, half /= 2) {
                const bool up = (t >> s) & 1;
#pragma unroll
                for (int i = 0; i < half; i++) {
                    const float send = up ? r[i] : r[i + half];
                    const float keep = up ? r[i + half] : r[i];
                    r[i] = keep + __shfl_xor_sync(0xffffffffu, send, 1 << s);
                }
            }
#pragma unroll
            for (int m = NV; m < DG_LANES; m *= 2)
                r[0] += __shfl_xor_sync(0xffffffffu, r[0], m);
            if (store) vis_f[idx * (2 * P) + comp] = vold - wt * r[0];
        }
        __syncwarp();
    }
    if (rejected != 0 && prm.num_rejected != nullptr) atomicAdd(prm.num_rejected, rejected);
}

template <int P, int MX>
static int launch_degrid_cached(const DegridParams &prm, size_t smem, cudaStream_t stream)
{
    // Runs long enough that the first load of a group's cells (MX x P per lane) is noise,
    // short enough for two waves of blocks; small launches get shorter runs down to one batch.
    const long long one_wave = (long long) sm_count() * 2 * DG_GROUPS;
    long long run = (prm.num_vis + 2 * one_wave - 1) / (2 * one_wave);
    if (run < 64) {
        run = (prm.num_vis + one_wave - 1) / one_wave;
        if (run > 64) run = 64;
    }
    if (run > 4096) run = 4096;
    run = (run + DG_LANES - 1) / DG_LANES * DG_LANES;
    const long long groups = (prm.num_vis + run - 1) / run;
    const unsigned blocks = (unsigned) ((groups + DG_GROUPS - 1) / DG_GROUPS);
    auto kernel = degrid_cached_kernel<P, MX>;
    if (smem > 48 * 1024)
        KIB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) smem));
    kernel<<<blocks, DG_THREADS, smem, stream>>>(prm, (int) run);
    KIB_CHECK_LAUNCH();
    return 0;
}

template <int P>
static int dispatch_degrid_cached(const DegridParams &prm, int mx, size_t smem, cudaStream_t stream)
{
    switch (mx) {
    case 4: return launch_degrid_cached<P, 4>(prm, smem, stream);
    case 5: return launch_degrid_cached<P, 5>(prm, smem, stream);
    case 6: return launch_degrid_cached<P, 6>(prm, smem, stream);
    case 7: return launch_degrid_cached<P, 7>(prm, smem, stream);
    case 8: return launch_degrid_cached<P, 8>(prm, smem, stream);
    }
    set_error("kib_degrid: no cached kernel for %d column slots", mx);
    return -1;
}

// Returns 1 when the launch was not taken (caller falls through to the generic kernel).
static int try_degrid_cached(const DegridParams &prm, int P, cudaStream_t stream)
{
    const int K = prm.kernel_width;
    if (K > 8 || prm.grid_size >= 32768 || prm.grid_size < 16) return 1;
    const int mx = K < 4 ? 4 : K;
    const size_t rows = (size_t) prm.w_planes * prm.oversample;
    const size_t entries = rows * 2 * (mx + DG_LANES);
    const size_t smem = (size_t) DG_GROUPS * DG_LANES * sizeof(int4) + entries * sizeof(float2);
    if (entries >= 65536 || entries * sizeof(float2) > (size_t) DG_LUT_SMEM_LIMIT) return 1;
    switch (P) {
    case 1: return dispatch_degrid_cached<1>(prm, mx, smem, stream);
    case 2: return dispatch_degrid_cached<2>(prm, mx, smem, stream);
    case 3: return dispatch_degrid_cached<3>(prm, mx, smem, stream);
    case 4: return dispatch_degrid_cached<4>(prm, mx, smem, stream);
    }
    return 1;
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
        // "thread": scalar kernel, "cached": register-cached groups (K <= 8), default "vec"
        const char *route = getenv("KIB_DEGRID_ROUTE");
        if (route && route[0] == 'c') {
            const int rc = try_degrid_cached(prm, num_pols, as_stream(stream));
            if (rc != 1) return rc;
        }
        if (!(route && route[0] == 't')) {
            const bool wide = !(route && route[0] == 'h');      // "hoist": taps in registers, 8-byte loads
            const int rc = try_degrid_vec(prm, num_pols, wide, as_stream(stream));
            if (rc != 1) return rc;
        }
        return launch_degrid<float>(prm, num_pols, as_stream(stream));
    }
    return launch_degrid<double>(prm, num_pols, as_stream(stream));
}
