// Hogbom CLEAN minor-cycle kernels for sm_100a.
//
// Replaces the operations of reference katsdpimager/clean.py (_UpdateTiles :398,
// _FindPeak :516, _SubtractPsf :625, Clean :756, PsfPatch :72, NoiseEst :247) and
// imager_kernels/clean/*.mako.  Results are bit-exact with the HOST classes
// (CleanHost clean.py:971-1075, _tile_peak :946-968):
//   * tile peak = first strict maximum in row-major order starting from best = 0;
//     a tile without a positive metric stores 0 and the position (x0, y0) [sic];
//   * global peak = first maximum over tiles in row-major tile order (np.argmax);
//   * the SUMSQ metric and dirty -= (gain*peak)*psf use separately rounded multiplies
//     and adds (numba / numpy do not contract to FMA).
//
// The reference syncs with the host on every minor cycle (clean.py:875-878).  Here a
// whole batch of cycles runs device-resident: one `clean_step_kernel` launch per cycle
// subtracts the PSF patch, recomputes the peaks of the 32x32 tiles it touched (from the
// values still in registers) and the last block to finish selects the next global
// peak, so no cycle ever waits for the host.
#include "kib_common.cuh"
#include <climits>
#include <cstdlib>
#include <cstring>

namespace kib {

constexpr int TILE = 32;
constexpr int CLEAN_THREADS = 256;      // 32 x 8 threads, 4 rows of a tile each

template <typename Real>
struct Best {
    Real value;
    int key;   // tie-break: smaller key wins
};

template <typename Real>
__device__ __forceinline__ Best<Real> better(Best<Real> a, Best<Real> b)
{
    return (b.value > a.value || (b.value == a.value && b.key < a.key)) ? b : a;
}

template <typename Real>
__device__ __forceinline__ Best<Real> warp_best(Best<Real> v)
{
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) {
        Best<Real> o;
        o.value = __shfl_xor_sync(0xffffffffu, v.value, offset);
        o.key = __shfl_xor_sync(0xffffffffu, v.key, offset);
        v = better(v, o);
    }
    return v;
}

// Reduction over a block of up to 1024 threads; result valid in every thread.
template <typename Real>
__device__ __forceinline__ Best<Real> block_best(Best<Real> v, Best<Real> *scratch /* [33] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps = (blockDim.x + 31) >> 5;
    v = warp_best(v);
    __syncthreads();        // protect scratch from a previous use
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        Best<Real> w;
        w.value = -1;
        w.key = INT_MAX;
        if (lane < warps) w = scratch[lane];
        w = warp_best(w);
        if (lane == 0) scratch[32] = w;
    }
    __syncthreads();
    return scratch[32];
}

// The same reduction with the result valid in thread 0 only and a single barrier.  Successive
// calls must alternate between two scratch buffers (8 entries each, blocks of 256 threads): a
// warp that writes buffer b for call n + 2 has passed the barrier of call n + 1, which warp 0
// only reaches after it has read buffer b for call n.
template <typename Real>
__device__ __forceinline__ Best<Real> block_best_thread0(Best<Real> v, Best<Real> *scratch /* [8] */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_best(v);
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        Best<Real> w;
        w.value = -1;
        w.key = INT_MAX;
        if (lane < CLEAN_THREADS / 32) w = scratch[lane];
        v = warp_best(w);
    }
    return v;
}

template <typename Real> __device__ __forceinline__ Real mul_rn_(Real a, Real b);
template <> __device__ __forceinline__ float mul_rn_(float a, float b) { return __fmul_rn(a, b); }
template <> __device__ __forceinline__ double mul_rn_(double a, double b) { return __dmul_rn(a, b); }
template <typename Real> __device__ __forceinline__ Real add_rn_(Real a, Real b);
template <> __device__ __forceinline__ float add_rn_(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double add_rn_(double a, double b) { return __dadd_rn(a, b); }

// Peak-finding metric of one pixel (clean.py:952-967; metric.mako)
template <typename Real, int MODE>
__device__ __forceinline__ Real clean_metric(const Real *pix, int P)
{
    if (MODE == KIB_CLEAN_I) return fabs(pix[0]);
    Real value = 0;
    for (int p = 0; p < P; p++) value = add_rn_(value, mul_rn_(pix[p], pix[p]));
    return value;
}

// --------------------------------------------------------------------- update_tiles
template <typename Real, int MODE>
__global__ void __launch_bounds__(CLEAN_THREADS)
update_tiles_kernel(const Real *__restrict__ dirty, int row_stride, long long pol_stride,
                    int width, int height, int P, int border,
                    Real *__restrict__ tile_max, int2 *__restrict__ tile_pos, int tile_stride,
                    int tx0, int ty0)
{
    __shared__ Best<Real> scratch[33];
    const int tx = tx0 + blockIdx.x, ty = ty0 + blockIdx.y;
    const int x0 = tx * TILE + border, y0 = ty * TILE + border;
    const int x1 = min(x0 + TILE, width - border), y1 = min(y0 + TILE, height - border);
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    Best<Real> best;
    best.value = 0;
    best.key = INT_MAX;
    const int x = x0 + lx;
#pragma unroll
    for (int k = 0; k < TILE / 8; k++) {
        const int y = y0 + ly + 8 * k;
        if (x < x1 && y < y1) {
            Real pix[4];
            const long long addr = (long long) y * row_stride + x;
            const int np = MODE == KIB_CLEAN_I ? 1 : P;
            for (int p = 0; p < np; p++) pix[p] = dirty[p * pol_stride + addr];
            const Real value = clean_metric<Real, MODE>(pix, P);
            if (value > best.value) {
                best.value = value;
                best.key = y * width + x;
            }
        }
    }
    best = block_best(best, scratch);
    if (threadIdx.x == 0) {
        const long long idx = (long long) ty * tile_stride + tx;
        tile_max[idx] = best.value;
        // clean.py:950: a tile with no positive value reports (x0, y0)
        tile_pos[idx] = best.key == INT_MAX ? make_int2(x0, y0)
                                            : make_int2(best.key / width, best.key % width);
    }
}

// ------------------------------------------------------------------------ find_peak
// Body shared by the stand-alone kernel and the tail of clean_step_kernel.
template <typename Real>
__device__ __forceinline__ void find_peak_body(
    const Real *dirty, int row_stride, long long pol_stride, int P,
    const Real *tile_max, const int2 *tile_pos, int tile_stride, int tiles_x, int tiles_y,
    Real *peak_value, int *peak_pos, Real *peak_pixel, Best<Real> *scratch)
{
    Best<Real> best;
    best.value = -1;
    best.key = INT_MAX;
    const int total = tiles_x * tiles_y;
#pragma unroll 4
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int ty = i / tiles_x, tx = i - ty * tiles_x;
        const Real value = __ldcg(tile_max + (long long) ty * tile_stride + tx);
        if (value > best.value) {      // i increases, so the first maximum is kept
            best.value = value;
            best.key = i;
        }
    }
    best = block_best(best, scratch);
    if (threadIdx.x == 0) {
        int2 pos = make_int2(0, 0);
        Real value = 0;
        if (best.key != INT_MAX) {
            const int ty = best.key / tiles_x, tx = best.key - ty * tiles_x;
            pos = __ldcg(tile_pos + (long long) ty * tile_stride + tx);
            value = best.value;
        }
        peak_value[0] = value;
        peak_pos[0] = pos.x;    // row
        peak_pos[1] = pos.y;    // column
        for (int p = 0; p < P; p++)
            peak_pixel[p] = __ldcg(dirty + p * pol_stride + (long long) pos.x * row_stride + pos.y);
    }
}

template <typename Real>
__global__ void __launch_bounds__(1024)
find_peak_kernel(const Real *dirty, int row_stride, long long pol_stride, int P,
                 const Real *tile_max, const int2 *tile_pos, int tile_stride,
                 int tiles_x, int tiles_y, Real *peak_value, int *peak_pos, Real *peak_pixel)
{
    __shared__ Best<Real> scratch[33];
    find_peak_body(dirty, row_stride, pol_stride, P, tile_max, tile_pos, tile_stride,
                   tiles_x, tiles_y, peak_value, peak_pos, peak_pixel, scratch);
}

// ------------------------------------------------------------------ per-row tile maxima
// row_max[ty] = max over tx of tile_max[ty][tx], row_arg[ty] = first tx attaining it.
// Lets the minor-cycle kernel find the global peak from (rows touched) x tiles_x + tiles_y
// loads instead of scanning all tiles_y x tiles_x tiles every cycle.
template <typename Real>
__device__ __forceinline__ void tile_row_max(const Real *tile_max, int tile_stride, int tiles_x,
                                             int ty, Real *row_max, int *row_arg)
{
    // one warp per row
    const int lane = threadIdx.x & 31;
    Best<Real> best;
    best.value = -1;
    best.key = INT_MAX;
    for (int tx = lane; tx < tiles_x; tx += 32) {
        const Real value = __ldcg(tile_max + (long long) ty * tile_stride + tx);
        if (value > best.value) {
            best.value = value;
            best.key = tx;
        }
    }
    best = warp_best(best);
    if (lane == 0) {
        row_max[ty] = best.value;
        row_arg[ty] = best.key;
    }
}

template <typename Real>
__global__ void __launch_bounds__(256)
tile_rows_kernel(const Real *tile_max, int tile_stride, int tiles_x, int tiles_y,
                 Real *row_max, int *row_arg)
{
    const int ty = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ty < tiles_y) tile_row_max(tile_max, tile_stride, tiles_x, ty, row_max, row_arg);
}

// --------------------------------------------------------------------- subtract_psf
template <typename Real>
__global__ void __launch_bounds__(256)
subtract_psf_kernel(Real *__restrict__ dirty, Real *__restrict__ model,
                    int row_stride, long long pol_stride, int width, int height, int P,
                    const Real *__restrict__ psf, int psf_row_stride, long long psf_pol_stride,
                    int psf_x0, int psf_y0, int patch_w, int patch_h,
                    const Real *__restrict__ peak_pixel, int pos_y, int pos_x,
                    int start_x, int start_y, Real loop_gain)
{
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int gy = blockIdx.y;
    if (gx == 0 && gy == 0) {
        for (int p = 0; p < P; p++) {
            const long long c = p * pol_stride + (long long) pos_y * row_stride + pos_x;
            model[c] = add_rn_(model[c], mul_rn_(loop_gain, peak_pixel[p]));
        }
    }
    if (gx >= patch_w) return;
    const int x = start_x + gx, y = start_y + gy;
    if (x < 0 || x >= width || y < 0 || y >= height) return;
    const long long addr = (long long) y * row_stride + x;
    const long long paddr = (long long) (psf_y0 + gy) * psf_row_stride + psf_x0 + gx;
    for (int p = 0; p < P; p++) {
        const Real scale = mul_rn_(loop_gain, peak_pixel[p]);
        const Real d = dirty[p * pol_stride + addr];
        dirty[p * pol_stride + addr] = add_rn_(d, -mul_rn_(scale, psf[p * psf_pol_stride + paddr]));
    }
}

// ------------------------------------------------------------- device-resident cycle
struct CleanStepParams {
    void *dirty;
    void *model;
    const void *psf;
    void *tile_max;
    int2 *tile_pos;
    void *peak_value;
    int *peak_pos;
    void *peak_pixel;
    void *components;
    void *row_max;        // per tile row: maximum of tile_max and the first column attaining it
    int *row_arg;
    int *state;           // [0] cycles done, [1] stopped by threshold, [2] block ticket
    long long pol_stride;
    long long psf_pol_stride;
    double loop_gain;
    double threshold;
    int row_stride, width, height, border;
    int psf_row_stride, psf_width, psf_height;
    int patch_w, patch_h;
    int tile_stride, tiles_x, tiles_y;
    int component_stride;   // bytes
    int max_components;
};

__device__ __forceinline__ int floordiv32(int a) { return a >> 5; }   // arithmetic shift = floor

// Geometry of one minor cycle: the patch rectangle (clean.py:1024-1043) clipped to the image
// and the lattice cells (32 x 32, aligned with the tiles, extended over the border) under it.
struct CleanCycle {
    int pos_x, pos_y;
    int cx0, cy0, cx1, cy1;       // clipped patch
    int psf_x0, psf_y0;           // psf x = image x + psf_x0
    int cell_x0, cell_y0;         // first lattice cell
    int cells_x, cells_y;
};

__device__ __forceinline__ CleanCycle clean_cycle_geometry(const CleanStepParams &prm,
                                                           int pos_y, int pos_x)
{
    CleanCycle g;
    g.pos_x = pos_x;
    g.pos_y = pos_y;
    const int px0 = pos_x - prm.patch_w / 2, py0 = pos_y - prm.patch_h / 2;
    g.cx0 = max(px0, 0);
    g.cy0 = max(py0, 0);
    g.cx1 = min(px0 + prm.patch_w, prm.width);
    g.cy1 = min(py0 + prm.patch_h, prm.height);
    g.psf_x0 = prm.psf_width / 2 - prm.patch_w / 2 - px0;
    g.psf_y0 = prm.psf_height / 2 - prm.patch_h / 2 - py0;
    g.cell_x0 = floordiv32(g.cx0 - prm.border);
    g.cell_y0 = floordiv32(g.cy0 - prm.border);
    g.cells_x = floordiv32(g.cx1 - 1 - prm.border) - g.cell_x0 + 1;
    g.cells_y = floordiv32(g.cy1 - 1 - prm.border) - g.cell_y0 + 1;
    return g;
}

// Subtract the patch from the pixels of lattice cell (cell_x, cell_y) and recompute the peak
// of its tile from the values still in registers.  One block of CLEAN_THREADS threads.
template <typename Real, int P, int MODE>
__device__ __forceinline__ void clean_cell(const CleanStepParams &prm, const CleanCycle &g,
                                           const Real (&scale)[P], int cell_x, int cell_y,
                                           Best<Real> (*scratch)[CLEAN_THREADS / 32], int &parity)
{
    Real *const dirty = static_cast<Real *>(prm.dirty);
    const Real *const psf = static_cast<const Real *>(prm.psf);
    const int W = prm.width, H = prm.height, border = prm.border;
    const int rx0 = border + cell_x * TILE, ry0 = border + cell_y * TILE;
    const bool is_tile = cell_x >= 0 && cell_x < prm.tiles_x && cell_y >= 0 && cell_y < prm.tiles_y;
    const bool touches = rx0 < g.cx1 && rx0 + TILE > g.cx0 && ry0 < g.cy1 && ry0 + TILE > g.cy0;
    if (!touches) return;
    const int lx = threadIdx.x & 31, ly = threadIdx.x >> 5;
    const int x = rx0 + lx;
    Best<Real> best;
    best.value = 0;
    best.key = INT_MAX;
    const bool x_in_patch = x >= g.cx0 && x < g.cx1;
    const bool x_in_tile = is_tile && x >= border && x < W - border;
    // All loads of the thread's four rows are issued before the first store: a load of `dirty`
    // may not pass an earlier store to `dirty`, so a loop of load / subtract / store per row
    // pays one L2 round trip per row (four per cell) instead of one.
    Real d[TILE / 8][P], psfv[TILE / 8][P];
    bool in_patch[TILE / 8], in_tile[TILE / 8];
#pragma unroll
    for (int k = 0; k < TILE / 8; k++) {
        const int y = ry0 + ly + 8 * k;
        in_patch[k] = x_in_patch && y >= g.cy0 && y < g.cy1;
        in_tile[k] = x_in_tile && y >= border && y < H - border;
        const long long addr = (long long) y * prm.row_stride + x;
        if (in_patch[k]) {
            const long long paddr = (long long) (y + g.psf_y0) * prm.psf_row_stride + x + g.psf_x0;
#pragma unroll
            for (int p = 0; p < P; p++) {
                d[k][p] = dirty[p * prm.pol_stride + addr];
                psfv[k][p] = __ldg(psf + p * prm.psf_pol_stride + paddr);
            }
        } else if (in_tile[k]) {
            const int np = MODE == KIB_CLEAN_I ? 1 : P;
#pragma unroll
            for (int p = 0; p < P; p++)
                if (p < np) d[k][p] = dirty[p * prm.pol_stride + addr];
        }
    }
#pragma unroll
    for (int k = 0; k < TILE / 8; k++) {
        const int y = ry0 + ly + 8 * k;
        if (in_patch[k] || in_tile[k]) {
            const long long addr = (long long) y * prm.row_stride + x;
            Real pix[P];
            if (in_patch[k]) {
#pragma unroll
                for (int p = 0; p < P; p++) {
                    pix[p] = add_rn_(d[k][p], -mul_rn_(scale[p], psfv[k][p]));
                    dirty[p * prm.pol_stride + addr] = pix[p];
                }
            } else {
                const int np = MODE == KIB_CLEAN_I ? 1 : P;
#pragma unroll
                for (int p = 0; p < P; p++)
                    if (p < np) pix[p] = d[k][p];
            }
            if (in_tile[k]) {
                const Real value = clean_metric<Real, MODE>(pix, P);
                if (value > best.value) {
                    best.value = value;
                    best.key = y * W + x;
                }
            }
        }
    }
    if (is_tile) {
        best = block_best_thread0(best, scratch[parity]);
        parity ^= 1;
        if (threadIdx.x == 0) {
            const long long idx = (long long) cell_y * prm.tile_stride + cell_x;
            static_cast<Real *>(prm.tile_max)[idx] = best.value;
            prm.tile_pos[idx] = best.key == INT_MAX ? make_int2(rx0, ry0)
                                                    : make_int2(best.key / W, best.key % W);
        }
    }
}

// Model image and component list (clean.py:1047, :882; imaging.py:389-396); one thread.
template <typename Real, int P>
__device__ __forceinline__ void clean_record(const CleanStepParams &prm, const CleanCycle &g,
                                             const Real (&scale)[P], Real peak_value, int done)
{
    Real *const model = static_cast<Real *>(prm.model);
    char *rec = static_cast<char *>(prm.components) + (long long) done * prm.component_stride;
    reinterpret_cast<int *>(rec)[0] = g.pos_y;
    reinterpret_cast<int *>(rec)[1] = g.pos_x;
    Real *vals = reinterpret_cast<Real *>(rec + 8);
    vals[0] = peak_value;
#pragma unroll
    for (int p = 0; p < P; p++) {
        const long long c = p * prm.pol_stride + (long long) g.pos_y * prm.row_stride + g.pos_x;
        model[c] = add_rn_(model[c], scale[p]);
        vals[1 + p] = scale[p];
    }
}

// Refresh the row maxima of the tile rows a cycle touched (one warp per row), then take the
// first maximum over rows -- the same answer as np.argmax over all tiles -- and publish the
// next peak.  Whole block; all tile updates of the cycle must be visible.
template <typename Real, int P>
__device__ __forceinline__ void clean_next_peak(const CleanStepParams &prm, const CleanCycle &g,
                                                Best<Real> *scratch)
{
    Real *const row_max = static_cast<Real *>(prm.row_max);
    const Real *const tile_max = static_cast<const Real *>(prm.tile_max);
    const Real *const dirty = static_cast<const Real *>(prm.dirty);
    int ty0 = g.cell_y0, ty1 = g.cell_y0 + g.cells_y;
    if (ty0 < 0) ty0 = 0;
    if (ty1 > prm.tiles_y) ty1 = prm.tiles_y;
    for (int ty = ty0 + (threadIdx.x >> 5); ty < ty1; ty += CLEAN_THREADS / 32)
        tile_row_max(tile_max, prm.tile_stride, prm.tiles_x, ty, row_max, prm.row_arg);
    __syncthreads();
    Best<Real> best;
    best.value = -1;
    best.key = INT_MAX;
    for (int ty = threadIdx.x; ty < prm.tiles_y; ty += CLEAN_THREADS) {
        const Real value = row_max[ty];
        if (value > best.value) {
            best.value = value;
            best.key = ty;
        }
    }
    best = block_best(best, scratch);
    if (threadIdx.x == 0) {
        int2 pos = make_int2(0, 0);
        Real value = 0;
        if (best.key != INT_MAX) {
            const int ty = best.key, tx = prm.row_arg[ty];
            pos = __ldcg(prm.tile_pos + (long long) ty * prm.tile_stride + tx);
            value = best.value;
        }
        static_cast<Real *>(prm.peak_value)[0] = value;
        prm.peak_pos[0] = pos.x;
        prm.peak_pos[1] = pos.y;
#pragma unroll
        for (int p = 0; p < P; p++)
            static_cast<Real *>(prm.peak_pixel)[p] =
                __ldcg(dirty + p * prm.pol_stride + (long long) pos.x * prm.row_stride + pos.y);
    }
}

// One launch per minor cycle, consecutive launches overlapped by programmatic dependent
// launch.  Used when a cooperative launch is not possible (see clean_persistent_kernel).
template <typename Real, int P, int MODE>
__global__ void __launch_bounds__(CLEAN_THREADS)
clean_step_kernel(const CleanStepParams prm)
{
    __shared__ Best<Real> scratch[33];
    __shared__ int is_last;
    int *state = prm.state;
    // Programmatic dependent launch: consecutive cycles are launched with stream serialization
    // relaxed, so this grid may already be resident while the previous cycle finishes.  Wait
    // for its completion (and the visibility of its writes) before touching any state, then
    // let the next cycle's blocks be scheduled behind this one.
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;");
    // Uniform across the grid: state[1] is only ever written by launches in which every
    // block takes the "below threshold" exit.
    if (__ldcg(state + 1) != 0) return;
    const Real pv = __ldcg(static_cast<Real *>(prm.peak_value));
    const int done = __ldcg(state);
    if ((double) pv < prm.threshold || done >= prm.max_components) {
        if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 && (double) pv < prm.threshold)
            state[1] = 1;
        return;
    }
    const CleanCycle g = clean_cycle_geometry(prm, __ldcg(prm.peak_pos), __ldcg(prm.peak_pos + 1));
    Real scale[P];
#pragma unroll
    for (int p = 0; p < P; p++)
        scale[p] = mul_rn_((Real) prm.loop_gain, __ldcg(static_cast<Real *>(prm.peak_pixel) + p));
    // the launch grid covers the largest number of cells a patch can touch; blocks beyond this
    // cycle's cells have nothing to subtract
    __shared__ Best<Real> cell_scratch[2][CLEAN_THREADS / 32];
    int parity = 0;
    if ((int) blockIdx.x < g.cells_x && (int) blockIdx.y < g.cells_y)
        clean_cell<Real, P, MODE>(prm, g, scale, g.cell_x0 + blockIdx.x, g.cell_y0 + blockIdx.y,
                                  cell_scratch, parity);
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
        clean_record<Real, P>(prm, g, scale, pv, done);
    // Last block to finish selects the next peak.
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const int ticket = atomicAdd(state + 2, 1);
        is_last = ticket == (int) (gridDim.x * gridDim.y) - 1;
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        clean_next_peak<Real, P>(prm, g, scratch);
        if (threadIdx.x == 0) {
            // The cycle count is published only here, after every block of this launch has
            // taken its ticket: a block that starts late must still see `done`, not done + 1
            // (it would skip its share of the subtraction on the last cycle of a batch).
            state[0] = done + 1;
            state[2] = 0;
        }
    }
}

// All cycles of a batch in ONE cooperative launch (single precision): the blocks stay
// resident and share out the lattice cells of each cycle.  A cycle is one chain of dependent
// round trips to L2, so the protocol is built to keep that chain short:
//   * arrival = one acq_rel atomic per block; the last arrival selects the next peak from one
//     parallel load of candidates (row maxima of untouched tile rows, tiles of touched rows);
//   * the next peak travels to the other blocks as two self-validating 16-byte packets
//     {cycle, y << 16 | x, value, pixel[0]}, {cycle, pixel[1..3]} that they poll with acquire
//     loads -- no separate flag, no fence on the critical path.
// state: [0] cycles done, [1] stopped by threshold, [2] arrivals, [4..7] / [8..11] packets.
__device__ __forceinline__ uint4 ld_acquire_v4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.acquire.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_v4(uint4 *p, uint4 v)
{
    asm volatile("st.release.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ int atom_add_acq_rel(int *p, int v)
{
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}

template <int P, int MODE>
__global__ void __launch_bounds__(CLEAN_THREADS, 4)
clean_persistent_kernel(const CleanStepParams prm)
{
    typedef float Real;
    __shared__ Best<Real> scratch[33];
    __shared__ int is_last;
    __shared__ unsigned s_peak[8];                       // y, x, value, pixel[0..3]
    __shared__ Best<Real> cell_scratch[2][CLEAN_THREADS / 32];
    int parity = 0;
    int *state = prm.state;
    uint4 *const packets = reinterpret_cast<uint4 *>(state + 4);
    const int nblocks = (int) gridDim.x;
    const Real *const dirty = static_cast<const Real *>(prm.dirty);
    for (int done = 0; done < prm.max_components; done++) {
        // ---- this cycle's peak
        if (threadIdx.x == 0) {
            if (done == 0) {
                s_peak[0] = (unsigned) __ldcg(prm.peak_pos);
                s_peak[1] = (unsigned) __ldcg(prm.peak_pos + 1);
                s_peak[2] = __float_as_uint(__ldcg(static_cast<const Real *>(prm.peak_value)));
                for (int p = 0; p < P; p++)
                    s_peak[3 + p] = __float_as_uint(__ldcg(static_cast<const Real *>(prm.peak_pixel) + p));
            } else {
                uint4 a, b;
                // (relaxed polling followed by one fence measured slower: 8.6 against 7.9 us
                // per cycle at a 255^2 patch)
                do {
                    a = ld_acquire_v4(packets);
                    b = P > 1 ? ld_acquire_v4(packets + 1) : a;
                } while (a.x != (unsigned) done || b.x != (unsigned) done);
                s_peak[0] = a.y >> 16;
                s_peak[1] = a.y & 0xffffu;
                s_peak[2] = a.z;
                s_peak[3] = a.w;
                s_peak[4] = b.y;
                s_peak[5] = b.z;
                s_peak[6] = b.w;
            }
        }
        __syncthreads();
        const Real pv = __uint_as_float(s_peak[2]);
        if ((double) pv < prm.threshold) {
            if (blockIdx.x == 0 && threadIdx.x == 0) state[1] = 1;
            break;                                       // uniform: every block saw the same peak
        }
        const CleanCycle g = clean_cycle_geometry(prm, (int) s_peak[0], (int) s_peak[1]);
        Real scale[P];
#pragma unroll
        for (int p = 0; p < P; p++)
            scale[p] = mul_rn_((Real) prm.loop_gain, __uint_as_float(s_peak[3 + p]));
        const int cells = g.cells_x * g.cells_y;
        for (int cell = blockIdx.x; cell < cells; cell += nblocks) {
            const int cy = cell / g.cells_x, cx = cell - cy * g.cells_x;
            clean_cell<Real, P, MODE>(prm, g, scale, g.cell_x0 + cx, g.cell_y0 + cy, cell_scratch,
                                      parity);
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) clean_record<Real, P>(prm, g, scale, pv, done);
        // ---- arrive; the last arrival finds the next peak and publishes it
        __syncthreads();
        if (threadIdx.x == 0) is_last = atom_add_acq_rel(state + 2, 1) == nblocks - 1;
        __syncthreads();
        if (!is_last) continue;

        Real *const row_max = static_cast<Real *>(prm.row_max);
        const Real *const tile_max = static_cast<const Real *>(prm.tile_max);
        int ty0 = g.cell_y0, ty1 = g.cell_y0 + g.cells_y;
        if (ty0 < 0) ty0 = 0;
        if (ty1 > prm.tiles_y) ty1 = prm.tiles_y;
        Best<Real> best;
        best.value = -1;
        best.key = INT_MAX;
        // tile rows this cycle did not touch: their stored maxima
        for (int ty = threadIdx.x; ty < prm.tiles_y; ty += CLEAN_THREADS)
            if (ty < ty0 || ty >= ty1) {
                Best<Real> c;
                c.value = __ldcg(row_max + ty);
                c.key = ty * prm.tiles_x + __ldcg(prm.row_arg + ty);
                best = better(best, c);
            }
        // touched rows: one warp per row rescans the row and refreshes its stored maximum
        const int lane = threadIdx.x & 31;
        for (int ty = ty0 + (threadIdx.x >> 5); ty < ty1; ty += CLEAN_THREADS / 32) {
            Best<Real> row;
            row.value = -1;
            row.key = INT_MAX;
            for (int tx = lane; tx < prm.tiles_x; tx += 32) {
                const Real value = __ldcg(tile_max + (long long) ty * prm.tile_stride + tx);
                if (value > row.value) {
                    row.value = value;
                    row.key = tx;
                }
            }
            row = warp_best(row);
            if (lane == 0) {
                row_max[ty] = row.value;
                prm.row_arg[ty] = row.key;
            }
            row.key += ty * prm.tiles_x;
            best = better(best, row);
        }
        best = block_best(best, scratch);
        if (threadIdx.x == 0) {
            int2 pos = make_int2(0, 0);
            Real value = 0;
            if (best.key != INT_MAX) {
                const int ty = best.key / prm.tiles_x, tx = best.key - ty * prm.tiles_x;
                pos = __ldcg(prm.tile_pos + (long long) ty * prm.tile_stride + tx);
                value = best.value;
            }
            Real pix[4] = {0, 0, 0, 0};
#pragma unroll
            for (int p = 0; p < P; p++)
                pix[p] = __ldcg(dirty + p * prm.pol_stride + (long long) pos.x * prm.row_stride + pos.y);
            state[0] = done + 1;
            state[2] = 0;
            const unsigned seq = (unsigned) (done + 1);
            if (P > 1)
                st_release_v4(packets + 1, make_uint4(seq, __float_as_uint(pix[1]),
                                                      __float_as_uint(pix[2]), __float_as_uint(pix[3])));
            st_release_v4(packets, make_uint4(seq, ((unsigned) pos.x << 16) | (unsigned) pos.y,
                                              __float_as_uint(value), __float_as_uint(pix[0])));
            // the reference's slots (off the critical path)
            static_cast<Real *>(prm.peak_value)[0] = value;
            prm.peak_pos[0] = pos.x;
            prm.peak_pos[1] = pos.y;
#pragma unroll
            for (int p = 0; p < P; p++) static_cast<Real *>(prm.peak_pixel)[p] = pix[p];
        }
    }
}

// ------------------------------------------------------------------------ psf_patch
template <typename Real>
__global__ void __launch_bounds__(256)
psf_patch_kernel(const Real *__restrict__ psf, int row_stride, long long pol_stride, int P,
                 int min_x, int min_y, int max_x, int max_y, int mid_x, int mid_y,
                 Real threshold, int *__restrict__ bound)
{
    const int x = min_x + blockIdx.x * blockDim.x + threadIdx.x;
    const int y = min_y + blockIdx.y;
    int dx = 0, dy = 0;
    if (x <= max_x && y <= max_y) {
        bool over = false;
        for (int p = 0; p < P; p++)
            over |= fabs(psf[p * pol_stride + (long long) y * row_stride + x]) >= threshold;
        if (over) {
            dx = abs(x - mid_x);
            dy = abs(y - mid_y);
        }
    }
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1) {
        dx = max(dx, __shfl_xor_sync(0xffffffffu, dx, offset));
        dy = max(dy, __shfl_xor_sync(0xffffffffu, dy, offset));
    }
    if ((threadIdx.x & 31) == 0) {
        if (dx > 0) atomicMax(bound, dx);
        if (dy > 0) atomicMax(bound + 1, dy);
    }
}

// ------------------------------------------------------------------- noise estimate
// One radix digit of |pixel| (float bits) histogrammed over the region inside the border,
// optionally restricted to values whose higher bits equal `prefix`.  Blocks walk whole image
// rows (no per-element division); equal digits within a warp -- the common case for the
// leading digit, where almost every value shares its exponent -- are combined with
// match.any so that one lane adds the group's count (the shared-memory atomics of 32 lanes
// on one bin used to serialise: 0.9 ms per pass at 8192^2 x 4, now memory-bound).
__global__ void __launch_bounds__(256)
abs_histogram_kernel(const float *__restrict__ image, int row_stride, long long pol_stride,
                     int inner_w, int inner_h, int P, int border,
                     unsigned prefix, int prefix_shift, int use_prefix, int shift, unsigned mask,
                     unsigned *__restrict__ hist)
{
    extern __shared__ unsigned local[];
    for (unsigned i = threadIdx.x; i <= mask; i += blockDim.x) local[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int rows = inner_h * P;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const int p = r / inner_h, y = r - p * inner_h;
        const float *row = image + p * pol_stride + (long long) (y + border) * row_stride + border;
#pragma unroll 4
        for (int x0 = 0; x0 < inner_w; x0 += 256) {
            const int x = x0 + threadIdx.x;
            bool ok = x < inner_w;
            const unsigned bits = ok ? __float_as_uint(fabsf(__ldg(row + x))) : 0u;
            ok = ok && (!use_prefix || (bits >> prefix_shift) == prefix);
            const unsigned bin = (bits >> shift) & mask;
            const unsigned active = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const unsigned peers = __match_any_sync(active, bin);
                if (lane == __ffs(peers) - 1) atomicAdd(&local[bin], (unsigned) __popc(peers));
            }
        }
    }
    __syncthreads();
    for (unsigned i = threadIdx.x; i <= mask; i += blockDim.x)
        if (local[i] != 0) atomicAdd(&hist[i], local[i]);
}

// Second radix digit (bits shift .. shift+bits) of |pixel| for `window` adjacent values of the
// leading digit (first_prefix ...) at once, plus the number of values whose leading digit is
// smaller.  With a guess of the leading digit of the median (the exponent of the previous
// estimate of the same image: the window spans a factor of 256) this replaces the expensive
// first pass: the second digit is spread evenly over its bins, so the shared-memory atomics do
// not collide.
__global__ void __launch_bounds__(256)
abs_histogram_window_kernel(const float *__restrict__ image, int row_stride, long long pol_stride,
                            int inner_w, int inner_h, int P, int border,
                            unsigned first_prefix, unsigned window, int prefix_shift, int shift,
                            unsigned mask, unsigned *__restrict__ hist /* [window][mask + 1] */,
                            unsigned long long *__restrict__ below)
{
    extern __shared__ unsigned local[];
    const unsigned bins = mask + 1;
    for (unsigned i = threadIdx.x; i < window * bins; i += blockDim.x) local[i] = 0;
    __syncthreads();
    const int rows = inner_h * P;
    unsigned count_below = 0;
    auto tally = [&](float value) {
        const unsigned bits = __float_as_uint(fabsf(value));
        const unsigned rel = (bits >> prefix_shift) - first_prefix;     // wraps when smaller
        if (rel < window)
            atomicAdd(&local[rel * bins + ((bits >> shift) & mask)], 1u);
        else
            count_below += (bits >> prefix_shift) < first_prefix;
    };
    // 16-byte loads, two per thread in flight, when every row segment is aligned (the pass is
    // latency-bound with 4-byte loads: 41 % of the DRAM bandwidth at full occupancy)
    const bool vec = ((reinterpret_cast<size_t>(image) | (size_t) border * 4 | (size_t) row_stride * 4
                       | (size_t) pol_stride * 4) & 15) == 0 && (inner_w & 3) == 0;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        const int p = r / inner_h, y = r - p * inner_h;
        const float *row = image + p * pol_stride + (long long) (y + border) * row_stride + border;
        if (vec) {
            const float4 *row4 = reinterpret_cast<const float4 *>(row);
            const int n4 = inner_w >> 2;
            int x = threadIdx.x;
            for (; x + 256 < n4; x += 512) {
                const float4 a = __ldg(row4 + x), b = __ldg(row4 + x + 256);
                tally(a.x); tally(a.y); tally(a.z); tally(a.w);
                tally(b.x); tally(b.y); tally(b.z); tally(b.w);
            }
            if (x < n4) {
                const float4 a = __ldg(row4 + x);
                tally(a.x); tally(a.y); tally(a.z); tally(a.w);
            }
        } else {
#pragma unroll 4
            for (int x = threadIdx.x; x < inner_w; x += 256) tally(__ldg(row + x));
        }
    }
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1)
        count_below += __shfl_xor_sync(0xffffffffu, count_below, offset);
    if ((threadIdx.x & 31) == 0 && count_below != 0)
        atomicAdd(below, (unsigned long long) count_below);
    __syncthreads();
    for (unsigned i = threadIdx.x; i < window * bins; i += blockDim.x)
        if (local[i] != 0) atomicAdd(&hist[i], local[i]);
}

template <typename Real>
__global__ void __launch_bounds__(256)
rank_kernel(const Real *__restrict__ image, int row_stride, long long pol_stride,
            int inner_w, int inner_h, int P, int border, Real value,
            unsigned long long *__restrict__ rank)
{
    const long long total = (long long) inner_w * inner_h * P;
    unsigned count = 0;
    for (long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long) gridDim.x * blockDim.x) {
        const int x = (int) (i % inner_w);
        const long long r = i / inner_w;
        const int y = (int) (r % inner_h);
        const int p = (int) (r / inner_h);
        const Real v = image[p * pol_stride + (long long) (y + border) * row_stride + x + border];
        count += fabs(v) < value;
    }
#pragma unroll
    for (int offset = 16; offset > 0; offset >>= 1)
        count += __shfl_xor_sync(0xffffffffu, count, offset);
    if ((threadIdx.x & 31) == 0 && count != 0) atomicAdd(rank, (unsigned long long) count);
}

template <typename Real, int P>
static void launch_step_mode(const CleanStepParams &prm, int mode, dim3 grid, cudaStream_t stream)
{
    // programmatic stream serialization: see the top of clean_step_kernel
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t config = {};
    config.gridDim = grid;
    config.blockDim = dim3(CLEAN_THREADS);
    config.dynamicSmemBytes = 0;
    config.stream = stream;
    config.attrs = &attr;
    config.numAttrs = 1;
    if (mode == KIB_CLEAN_I)
        cudaLaunchKernelEx(&config, clean_step_kernel<Real, P, KIB_CLEAN_I>, prm);
    else
        cudaLaunchKernelEx(&config, clean_step_kernel<Real, P, KIB_CLEAN_SUMSQ>, prm);
}

// KIB_CLEAN_ROUTE=pdl forces one launch per cycle
static bool clean_route_pdl()
{
    const char *route = getenv("KIB_CLEAN_ROUTE");
    return route && strcmp(route, "pdl") == 0;
}

// Cooperative launch of clean_persistent_kernel with at most `cells` blocks (one lattice cell
// per block and cycle when they all fit).  Returns 1 if the device cannot co-schedule it.
template <int P, int MODE>
static int launch_persistent_mode(const CleanStepParams &prm, int cells, cudaStream_t stream)
{
    auto kernel = clean_persistent_kernel<P, MODE>;
    static int max_blocks = -1;                          // per instantiation
    if (max_blocks < 0) {
        int device = 0, cooperative = 0, per_sm = 0;
        KIB_CUDA(cudaGetDevice(&device));
        KIB_CUDA(cudaDeviceGetAttribute(&cooperative, cudaDevAttrCooperativeLaunch, device));
        KIB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, CLEAN_THREADS, 0));
        max_blocks = cooperative ? per_sm * sm_count() : 0;
    }
    if (max_blocks < 1) return 1;
    int blocks = cells < max_blocks ? cells : max_blocks;
    void *args[] = {const_cast<CleanStepParams *>(&prm)};
    KIB_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<const void *>(kernel), dim3(blocks),
                                         dim3(CLEAN_THREADS), args, 0, stream));
    return 0;
}

static int launch_persistent(const CleanStepParams &prm, int P, int mode, int cells,
                             cudaStream_t stream)
{
#define KIB_PERSISTENT(PP)                                                                      \
    return mode == KIB_CLEAN_I ? launch_persistent_mode<PP, KIB_CLEAN_I>(prm, cells, stream)    \
                               : launch_persistent_mode<PP, KIB_CLEAN_SUMSQ>(prm, cells, stream)
    switch (P) {
    case 1: KIB_PERSISTENT(1);
    case 2: KIB_PERSISTENT(2);
    case 3: KIB_PERSISTENT(3);
    case 4: KIB_PERSISTENT(4);
    default:
        set_error("kib_clean_minor_cycles: num_pols must be 1..4, not %d", P);
        return -1;
    }
#undef KIB_PERSISTENT
}

template <typename Real>
static int launch_step(const CleanStepParams &prm, int P, int mode, dim3 grid, cudaStream_t stream)
{
    switch (P) {
    case 1: launch_step_mode<Real, 1>(prm, mode, grid, stream); break;
    case 2: launch_step_mode<Real, 2>(prm, mode, grid, stream); break;
    case 3: launch_step_mode<Real, 3>(prm, mode, grid, stream); break;
    case 4: launch_step_mode<Real, 4>(prm, mode, grid, stream); break;
    default:
        set_error("kib_clean_minor_cycles: num_pols must be 1..4, not %d", P);
        return -1;
    }
    return 0;
}

}  // namespace kib

using namespace kib;

#define KIB_CHECK_DTYPE(name)                                                        \
    KIB_REQUIRE(dtype == KIB_F32 || dtype == KIB_F64, name ": bad dtype %d", dtype)
#define KIB_CHECK_MODE(name)                                                         \
    KIB_REQUIRE(mode == KIB_CLEAN_I || mode == KIB_CLEAN_SUMSQ, name ": bad clean mode %d", mode)

extern "C" {

int kib_update_tiles(const void *dirty, int row_stride, int64_t pol_stride,
                     int width, int height, int num_pols, int border, int mode,
                     void *tile_max, int32_t *tile_pos, int tile_stride,
                     int tx0, int ty0, int tx1, int ty1, int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_update_tiles");
    KIB_CHECK_MODE("kib_update_tiles");
    KIB_REQUIRE(num_pols >= 1 && num_pols <= 4, "kib_update_tiles: num_pols must be 1..4");
    if (tx0 >= tx1 || ty0 >= ty1) return 0;
    dim3 g(tx1 - tx0, ty1 - ty0, 1);
    cudaStream_t s = as_stream(stream);
    int2 *pos = reinterpret_cast<int2 *>(tile_pos);
#define LAUNCH(REAL, MODE)                                                                   \
    update_tiles_kernel<REAL, MODE><<<g, CLEAN_THREADS, 0, s>>>(                             \
        static_cast<const REAL *>(dirty), row_stride, pol_stride, width, height, num_pols,   \
        border, static_cast<REAL *>(tile_max), pos, tile_stride, tx0, ty0)
    if (dtype == KIB_F32) {
        if (mode == KIB_CLEAN_I) LAUNCH(float, KIB_CLEAN_I); else LAUNCH(float, KIB_CLEAN_SUMSQ);
    } else {
        if (mode == KIB_CLEAN_I) LAUNCH(double, KIB_CLEAN_I); else LAUNCH(double, KIB_CLEAN_SUMSQ);
    }
#undef LAUNCH
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_find_peak(const void *dirty, int row_stride, int64_t pol_stride, int num_pols,
                  const void *tile_max, const int32_t *tile_pos, int tile_stride,
                  int tiles_x, int tiles_y,
                  void *peak_value, int32_t *peak_pos, void *peak_pixel,
                  int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_find_peak");
    KIB_REQUIRE(tiles_x > 0 && tiles_y > 0, "kib_find_peak: empty tile array");
    const int2 *pos = reinterpret_cast<const int2 *>(tile_pos);
    if (dtype == KIB_F32)
        find_peak_kernel<float><<<1, 1024, 0, as_stream(stream)>>>(
            static_cast<const float *>(dirty), row_stride, pol_stride, num_pols,
            static_cast<const float *>(tile_max), pos, tile_stride, tiles_x, tiles_y,
            static_cast<float *>(peak_value), peak_pos, static_cast<float *>(peak_pixel));
    else
        find_peak_kernel<double><<<1, 1024, 0, as_stream(stream)>>>(
            static_cast<const double *>(dirty), row_stride, pol_stride, num_pols,
            static_cast<const double *>(tile_max), pos, tile_stride, tiles_x, tiles_y,
            static_cast<double *>(peak_value), peak_pos, static_cast<double *>(peak_pixel));
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_subtract_psf(void *dirty, void *model, int row_stride, int64_t pol_stride,
                     int width, int height, int num_pols,
                     const void *psf, int psf_row_stride, int64_t psf_pol_stride,
                     int psf_width, int psf_height,
                     int patch_width, int patch_height,
                     const void *peak_pixel, int pos_y, int pos_x, double loop_gain,
                     int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_subtract_psf");
    KIB_REQUIRE(patch_width >= 1 && patch_height >= 1 && patch_width <= psf_width
                && patch_height <= psf_height, "kib_subtract_psf: bad patch %d x %d",
                patch_width, patch_height);
    const int psf_x0 = psf_width / 2 - patch_width / 2;
    const int psf_y0 = psf_height / 2 - patch_height / 2;
    const int start_x = pos_x - patch_width / 2, start_y = pos_y - patch_height / 2;
    dim3 g(divup(patch_width, 256), patch_height, 1);
    if (dtype == KIB_F32)
        subtract_psf_kernel<float><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<float *>(dirty), static_cast<float *>(model), row_stride, pol_stride,
            width, height, num_pols, static_cast<const float *>(psf), psf_row_stride,
            psf_pol_stride, psf_x0, psf_y0, patch_width, patch_height,
            static_cast<const float *>(peak_pixel), pos_y, pos_x, start_x, start_y,
            (float) loop_gain);
    else
        subtract_psf_kernel<double><<<g, 256, 0, as_stream(stream)>>>(
            static_cast<double *>(dirty), static_cast<double *>(model), row_stride, pol_stride,
            width, height, num_pols, static_cast<const double *>(psf), psf_row_stride,
            psf_pol_stride, psf_x0, psf_y0, patch_width, patch_height,
            static_cast<const double *>(peak_pixel), pos_y, pos_x, start_x, start_y, loop_gain);
    KIB_CHECK_LAUNCH();
    return 0;
}

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
                           void *row_scratch, int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_clean_minor_cycles");
    KIB_REQUIRE(row_scratch != nullptr, "kib_clean_minor_cycles: null row scratch");
    KIB_CHECK_MODE("kib_clean_minor_cycles");
    KIB_REQUIRE(patch_width >= 1 && patch_height >= 1 && patch_width <= psf_width
                && patch_height <= psf_height, "kib_clean_minor_cycles: bad patch %d x %d",
                patch_width, patch_height);
    const int real_size = dtype == KIB_F32 ? 4 : 8;
    KIB_REQUIRE(component_stride >= 8 + real_size * (1 + num_pols),
                "kib_clean_minor_cycles: component stride %d too small", component_stride);
    KIB_REQUIRE(state != nullptr && components != nullptr, "kib_clean_minor_cycles: null state");
    if (max_cycles <= 0) return 0;
    CleanStepParams prm;
    prm.dirty = dirty;
    prm.model = model;
    prm.psf = psf;
    prm.tile_max = tile_max;
    prm.tile_pos = reinterpret_cast<int2 *>(tile_pos);
    prm.peak_value = peak_value;
    prm.peak_pos = peak_pos;
    prm.peak_pixel = peak_pixel;
    prm.components = components;
    prm.row_max = row_scratch;
    prm.row_arg = reinterpret_cast<int *>(static_cast<char *>(row_scratch)
                                          + (size_t) tiles_y * (dtype == KIB_F32 ? 4 : 8));
    prm.state = state;
    prm.pol_stride = pol_stride;
    prm.psf_pol_stride = psf_pol_stride;
    prm.loop_gain = loop_gain;
    prm.threshold = threshold;
    prm.row_stride = row_stride;
    prm.width = width;
    prm.height = height;
    prm.border = border;
    prm.psf_row_stride = psf_row_stride;
    prm.psf_width = psf_width;
    prm.psf_height = psf_height;
    prm.patch_w = patch_width;
    prm.patch_h = patch_height;
    prm.tile_stride = tile_stride;
    prm.tiles_x = tiles_x;
    prm.tiles_y = tiles_y;
    prm.component_stride = component_stride;
    prm.max_components = max_cycles;
    // A patch of width w starting anywhere touches at most (w - 1) / 32 + 2 lattice cells.
    dim3 g((patch_width - 1) / TILE + 2, (patch_height - 1) / TILE + 2, 1);
    cudaStream_t s = as_stream(stream);
    {
        const int rows_per_block = 256 / 32;
        const int blocks = divup(tiles_y, rows_per_block);
        if (dtype == KIB_F32)
            tile_rows_kernel<float><<<blocks, 256, 0, s>>>(
                static_cast<const float *>(tile_max), tile_stride, tiles_x, tiles_y,
                static_cast<float *>(prm.row_max), prm.row_arg);
        else
            tile_rows_kernel<double><<<blocks, 256, 0, s>>>(
                static_cast<const double *>(tile_max), tile_stride, tiles_x, tiles_y,
                static_cast<double *>(prm.row_max), prm.row_arg);
    }
    if (dtype == KIB_F32 && width <= 65536 && height <= 65536 && !clean_route_pdl()) {
        const int rc = launch_persistent(prm, num_pols, mode, (int) (g.x * g.y), s);
        if (rc <= 0) return rc;
        // rc == 1: no cooperative launch on this device
    }
    for (int i = 0; i < max_cycles; i++) {
        int rc = dtype == KIB_F32 ? launch_step<float>(prm, num_pols, mode, g, s)
                                  : launch_step<double>(prm, num_pols, mode, g, s);
        if (rc != 0) return rc;
    }
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_clean_minor_cycles_launches(int max_cycles, int dtype)
{
    // row maxima + cycles
    return 1 + ((clean_route_pdl() || dtype != KIB_F32) ? max_cycles : 1);
}

int kib_psf_patch(const void *psf, int row_stride, int64_t pol_stride, int num_pols,
                  int min_x, int min_y, int max_x, int max_y, int mid_x, int mid_y,
                  double threshold, int32_t *bound, int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_psf_patch");
    KIB_REQUIRE(bound != nullptr, "kib_psf_patch: null bound");
    cudaStream_t s = as_stream(stream);
    KIB_CUDA(cudaMemsetAsync(bound, 0, 2 * sizeof(int32_t), s));
    if (max_x < min_x || max_y < min_y) return 0;
    dim3 g(divup(max_x - min_x + 1, 256), max_y - min_y + 1, 1);
    if (dtype == KIB_F32)
        psf_patch_kernel<float><<<g, 256, 0, s>>>(
            static_cast<const float *>(psf), row_stride, pol_stride, num_pols,
            min_x, min_y, max_x, max_y, mid_x, mid_y, (float) threshold, bound);
    else
        psf_patch_kernel<double><<<g, 256, 0, s>>>(
            static_cast<const double *>(psf), row_stride, pol_stride, num_pols,
            min_x, min_y, max_x, max_y, mid_x, mid_y, threshold, bound);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_abs_histogram(const void *image, int row_stride, int64_t pol_stride,
                      int width, int height, int num_pols, int border,
                      uint32_t prefix, int prefix_bits, int shift, int bits,
                      uint32_t *hist, int dtype, kib_stream_t stream)
{
    KIB_REQUIRE(dtype == KIB_F32, "kib_abs_histogram: only float32 images are supported");
    KIB_REQUIRE(bits >= 1 && bits <= 13 && shift >= 0 && shift + bits <= 32,
                "kib_abs_histogram: bad digit (shift %d, bits %d)", shift, bits);
    KIB_REQUIRE(prefix_bits >= 0 && prefix_bits + shift + bits <= 32,
                "kib_abs_histogram: bad prefix");
    const int inner_w = width - 2 * border, inner_h = height - 2 * border;
    if (inner_w <= 0 || inner_h <= 0) return 0;
    const unsigned mask = (1u << bits) - 1;
    const long long rows = (long long) inner_h * num_pols;
    int blocks = sm_count() * 8;
    if (blocks > rows) blocks = (int) rows;
    abs_histogram_kernel<<<blocks, 256, (mask + 1) * sizeof(unsigned), as_stream(stream)>>>(
        static_cast<const float *>(image), row_stride, pol_stride, inner_w, inner_h, num_pols,
        border, prefix, shift + bits, prefix_bits > 0, shift, mask, hist);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_abs_histogram_window(const void *image, int row_stride, int64_t pol_stride,
                             int width, int height, int num_pols, int border,
                             uint32_t first_prefix, int window, int prefix_bits, int shift,
                             int bits, uint32_t *hist, unsigned long long *below, int dtype,
                             kib_stream_t stream)
{
    KIB_REQUIRE(dtype == KIB_F32, "kib_abs_histogram_window: only float32 images are supported");
    KIB_REQUIRE(bits >= 1 && bits <= 13 && shift >= 0 && prefix_bits >= 1
                && prefix_bits + shift + bits == 32 && window >= 1 && (window << bits) <= 8192,
                "kib_abs_histogram_window: bad digits (prefix %d, shift %d, bits %d, window %d)",
                prefix_bits, shift, bits, window);
    KIB_REQUIRE(hist != nullptr && below != nullptr, "kib_abs_histogram_window: null output");
    const int inner_w = width - 2 * border, inner_h = height - 2 * border;
    if (inner_w <= 0 || inner_h <= 0) return 0;
    const unsigned mask = (1u << bits) - 1;
    const long long rows = (long long) inner_h * num_pols;
    int blocks = sm_count() * 8;
    if (blocks > rows) blocks = (int) rows;
    abs_histogram_window_kernel<<<blocks, 256, (size_t) window * (mask + 1) * sizeof(unsigned),
                                  as_stream(stream)>>>(
        static_cast<const float *>(image), row_stride, pol_stride, inner_w, inner_h, num_pols,
        border, first_prefix, (unsigned) window, shift + bits, shift, mask, hist, below);
    KIB_CHECK_LAUNCH();
    return 0;
}

int kib_rank(const void *image, int row_stride, int64_t pol_stride,
             int width, int height, int num_pols, int border, double value,
             unsigned long long *rank, int dtype, kib_stream_t stream)
{
    KIB_CHECK_DTYPE("kib_rank");
    const int inner_w = width - 2 * border, inner_h = height - 2 * border;
    if (inner_w <= 0 || inner_h <= 0) return 0;
    const long long total = (long long) inner_w * inner_h * num_pols;
    int blocks = (int) ((total + 256 * 16 - 1) / (256 * 16));
    const int max_blocks = sm_count() * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    if (blocks < 1) blocks = 1;
    if (dtype == KIB_F32)
        rank_kernel<float><<<blocks, 256, 0, as_stream(stream)>>>(
            static_cast<const float *>(image), row_stride, pol_stride, inner_w, inner_h,
            num_pols, border, (float) value, rank);
    else
        rank_kernel<double><<<blocks, 256, 0, as_stream(stream)>>>(
            static_cast<const double *>(image), row_stride, pol_stride, inner_w, inner_h,
            num_pols, border, value, rank);
    KIB_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
