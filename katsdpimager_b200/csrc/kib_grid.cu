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
#include <map>
#include <type_traits>
#include <mutex>
#include <utility>

namespace kib {

constexpr int GRID_BATCH = 16;          // visibilities staged per group per batch
constexpr int GRID_MAX_GROUP = 512;      // largest thread group (one visibility stream)
constexpr int GRID_MAX_GROUPS_PER_BLOCK = 64;

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
    int batch;               // fast path: records per group per ring stage (<= GRID_TMA_BATCH)
    unsigned magic_x, magic_y;   // floor(2^32 / BX) + 1: exact u % BX for u < 65536
    int num_units;               // work units (gpb groups x run visibilities each)
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
    extern __shared__ __align__(128) unsigned char smem_raw[];
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

// =====================================================================================
// Fast path: convolution-kernel table resident in shared memory, visibility records
// pre-staged by a streaming pass and pulled into shared memory by TMA bulk copies.
//
// Used whenever the (rotation-friendly, see below) copy of the table fits in shared
// memory, which covers the small-support / many-visibility regime (e.g. 7x7 support,
// 16 W planes: 14 KB).  Differences from the generic kernel above:
//  * `grid_stage_kernel` (one thread per visibility, fully parallel so the dependent
//    uv -> density-weight gather latency is hidden by occupancy) writes one compact
//    record per visibility: footprint origin, shared-memory offsets of its u and v tap
//    rows, bit masks of the column / row slots that moved to a new grid cell relative to
//    the previous visibility of the run, and the weighted samples.  Records are laid
//    out [block][batch][entry][group] so that a block's batch is one contiguous chunk
//    and the groups sharing a warp read adjacent records (no bank conflicts);
//  * the gridding kernel streams those chunks through a 3-stage shared-memory ring
//    with cp.async.bulk + mbarrier, so no thread ever waits on global memory;
//  * each table row is stored twice back to back with period BX (resp. BY) and zero
//    padding beyond K, so the tap of column slot s is at  row + (BX - u0 % BX) + s:
//    no modulo and no bounds test in the inner loop, and slot offsets are per-thread
//    constants;
//  * the inner loop runs in warp lock step over consecutive visibilities that keep all
//    of the warp's cells (loads + FMAs only); when any lane's cell moved, the whole
//    warp leaves it, the affected lanes flush, and the loop resumes.
constexpr int GRID_LUT_SMEM_LIMIT = 96 * 1024;    // bytes of shared memory for the table
constexpr int GRID_STAGES = 3;
constexpr int GRID_TMA_BATCH = 16;                // records per group per stage

// Record = header + weighted samples.  Narrow header (16 B): origin, two 16-bit table
// offsets, two 32-bit slot masks.  Wide header (32 B, tables too large for shared memory
// or more than 32 slots per axis): origin, two 32-bit table offsets, two 64-bit masks.
__host__ __device__ constexpr int grid_record_bytes(int P, bool wide = false)
{
    return (wide ? 32 : 16) + (8 * P + 15) / 16 * 16;
}

struct GridStageParams {
    unsigned char *records;
    const float *weights_grid;
    const short4 *uv;
    const short *w_plane;
    const float2 *vis;
    int32_t *num_rejected;
    long long weights_pol_stride;
    long long num_vis;
    long long total;             // records to write (num_vis rounded up to whole blocks)
    int weights_row_stride;
    int grid_size;
    int w_planes;
    int oversample;
    int kernel_width;
    int uv_bias;
    int half_grid;
    int bx, by;
    int groups_per_block;
    int run;
    int batch;
    int lutv_base;               // offset of the v table in 8-byte units
    // doubled kernel tables (see grid_tma_kernel), built by the first threads
    float2 *tables;
    const float2 *lut;
    int lut_slice_stride, lut_tap_offset;
    int lutx_count, luty_count;
};

__device__ __forceinline__ unsigned long long change_mask(int old_pos, int new_pos, int B)
{
    const int d = new_pos - old_pos;
    if (d == 0) return 0ull;
    if (d >= B || -d >= B) return ~0ull;
    // d > 0: columns old .. old + d - 1 leave; d < 0: columns new .. new - d - 1 enter
    // (each takes over the slot of the column B cells away).
    int start = (d > 0 ? old_pos : new_pos) % B;
    const int count = d > 0 ? d : -d;
    unsigned long long mask = 0;
    for (int i = 0; i < count; i++) {
        mask |= 1ull << start;
        if (++start == B) start = 0;
    }
    return mask;
}

template <int P, bool WIDE>
__global__ void __launch_bounds__(256)
grid_stage_kernel(const GridStageParams prm)
{
    constexpr int REC = grid_record_bytes(P, WIDE);
    constexpr int HDR = WIDE ? 32 : 16;
    const long long idx = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    const int K = prm.kernel_width, G = prm.grid_size, BX = prm.bx, BY = prm.by;
    if (idx < prm.lutx_count + prm.luty_count) {
        // table entry: row r of the u (period BX) or v (period BY) table, stored twice
        const bool is_v = idx >= prm.lutx_count;
        const int t = is_v ? (int) idx - prm.lutx_count : (int) idx;
        const int period = is_v ? BY : BX;
        const int row = t / (2 * period);
        int d = t - row * 2 * period;
        if (d >= period) d -= period;
        prm.tables[idx] = d < K ? __ldg(prm.lut + (long long) row * prm.lut_slice_stride
                                        + prm.lut_tap_offset + d)
                                : make_float2(0.0f, 0.0f);
    }
    if (idx >= prm.total) return;
    // position of this visibility in the [block][batch][entry][group] layout
    const long long group_id = idx / prm.run;
    const int r = (int) (idx - group_id * prm.run);
    const int batch = r / prm.batch, e = r - batch * prm.batch;
    const long long block = group_id / prm.groups_per_block;
    const int g = (int) (group_id - block * prm.groups_per_block);
    const int nbatches = prm.run / prm.batch;
    const long long slot = (((block * nbatches + batch) * prm.batch + e)
                            * prm.groups_per_block + g);
    unsigned char *rec = prm.records + slot * REC;

    // null record: origin 0, zero-weight taps (second half of a doubled row), no moves
    int pos = 0, lutu = BX, lutv = prm.lutv_base + BY;
    unsigned long long xm = 0, ym = 0;
    float2 v[P];
#pragma unroll
    for (int p = 0; p < P; p++) v[p] = make_float2(0.0f, 0.0f);
    if (idx < prm.num_vis) {
        const short4 c = prm.uv[idx];
        int w = prm.w_plane[idx];
        int u0 = c.x - prm.uv_bias, v0 = c.y - prm.uv_bias;
        int su = c.z, sv = c.w;
        const bool ok = u0 >= 0 && v0 >= 0 && u0 + K <= G && v0 + K <= G
                        && w >= 0 && w < prm.w_planes
                        && su >= 0 && su < prm.oversample && sv >= 0 && sv < prm.oversample;
        if (!ok) {
            // contributes nothing; harmless stand-in coordinates (stateless, so that the
            // next visibility's masks can be derived from them as well)
            u0 = 0; v0 = 0; w = 0; su = 0; sv = 0;
            if (prm.num_rejected != nullptr) atomicAdd(prm.num_rejected, 1);
        }
        pos = u0 | (v0 << 16);
        lutu = ((w * prm.oversample + su) * 2 * BX) + BX - u0 % BX;
        lutv = prm.lutv_base + ((w * prm.oversample + sv) * 2 * BY) + BY - v0 % BY;
        if (r == 0) {
            xm = ~0ull;     // first visibility of a run: every slot is new
            ym = ~0ull;
        } else {
            const short4 pc = prm.uv[idx - 1];
            const int pw = prm.w_plane[idx - 1];
            int pu = pc.x - prm.uv_bias, pv = pc.y - prm.uv_bias;
            const bool pok = pu >= 0 && pv >= 0 && pu + K <= G && pv + K <= G
                             && pw >= 0 && pw < prm.w_planes
                             && pc.z >= 0 && pc.z < prm.oversample
                             && pc.w >= 0 && pc.w < prm.oversample;
            if (!pok) { pu = 0; pv = 0; }
            xm = change_mask(pu, u0, BX);
            ym = change_mask(pv, v0, BY);
        }
        if (ok) {
            const long long waddr = (long long) (c.y + prm.half_grid) * prm.weights_row_stride
                                    + (c.x + prm.half_grid);
#pragma unroll
            for (int p = 0; p < P; p++) {
                const float wt = __ldg(prm.weights_grid + p * prm.weights_pol_stride + waddr);
                v[p] = __ldg(prm.vis + idx * P + p);
                v[p].x *= wt;
                v[p].y *= wt;
            }
        }
    }
    if (WIDE) {
        reinterpret_cast<int4 *>(rec)[0] = make_int4(pos, lutu, lutv, 0);
        reinterpret_cast<int4 *>(rec)[1] = make_int4((int) (unsigned) xm, (int) (unsigned) (xm >> 32),
                                                     (int) (unsigned) ym, (int) (unsigned) (ym >> 32));
    } else {
        *reinterpret_cast<int4 *>(rec) = make_int4(pos, lutu | (lutv << 16),
                                                   (int) (unsigned) xm, (int) (unsigned) ym);
    }
    float2 *out = reinterpret_cast<float2 *>(rec + HDR);
    if (P % 2 == 0) {
#pragma unroll
        for (int p = 0; p < P; p += 2)
            *reinterpret_cast<float4 *>(out + p) = make_float4(v[p].x, v[p].y, v[p + 1].x, v[p + 1].y);
    } else {
#pragma unroll
        for (int p = 0; p < P; p++) out[p] = v[p];
    }
}

// ---- mbarrier / bulk-copy helpers (sm_90+ PTX)
__device__ __forceinline__ unsigned smem_u32(const void *ptr)
{
    return (unsigned) __cvta_generic_to_shared(ptr);
}

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void *dst, const void *src, unsigned bytes,
                                              unsigned long long *bar)
{
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <typename Real, int P, int MX, int MY, int MAXT, int MINB, bool TX1, bool WIDE>
__global__ void __launch_bounds__(MAXT, MINB)
grid_tma_kernel(const GridParams prm, const unsigned char *__restrict__ records,
                const float2 *__restrict__ tables)
{
    typedef typename Acc<Real>::type Complex;
    typedef typename std::conditional<WIDE, unsigned long long, unsigned>::type Mask;
    constexpr int REC = grid_record_bytes(P, WIDE);
    constexpr int HDR = WIDE ? 32 : 16;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int gpb = prm.groups_per_block;
    const int BX = prm.bx, BY = prm.by;
    const int G = prm.grid_size;
    const int rows = prm.w_planes * prm.oversample;
    // Shared memory: barriers | record ring [STAGES][BATCH][gpb] | u table [rows][2 BX] |
    // v table [rows][2 BY] (aliases the u table when BX == BY)
    unsigned long long *const full = reinterpret_cast<unsigned long long *>(smem_raw);
    unsigned long long *const empty = full + GRID_STAGES;
    unsigned long long *const lut_bar = empty + GRID_STAGES;
    unsigned char *const ring = smem_raw + 128;
    const int batch_len = prm.batch;
    const int stage_bytes = batch_len * gpb * REC;
    float2 *const lutx = reinterpret_cast<float2 *>(ring + GRID_STAGES * stage_bytes);
    const int lutx_count = rows * 2 * BX;
    const int luty_count = BX == BY ? 0 : rows * 2 * BY;

    const int tid = threadIdx.x;
    const unsigned lanes = __activemask();      // a block need not be a whole number of warps
    const int nbatches = prm.run / batch_len;
    // Persistent blocks: block k walks work units k, k + gridDim.x, ...; a unit is what
    // `gpb` groups grid in one run (gpb * run visibilities, `nbatches` ring stages).
    // Virtual batch vb of this block is batch vb % nbatches of its (vb / nbatches)-th unit.
    const int my_units = prm.num_units > (int) blockIdx.x
        ? (prm.num_units - 1 - (int) blockIdx.x) / (int) gridDim.x + 1 : 0;
    const int my_batches = my_units * nbatches;
    auto chunk_of = [&](int vb) {
        const int k = vb / nbatches;
        const long long unit = (long long) blockIdx.x + (long long) k * gridDim.x;
        return records + (unit * nbatches + (vb - k * nbatches)) * stage_bytes;
    };

    if (tid == 0) {
        const unsigned warps = (blockDim.x + 31) / 32;
        for (int s = 0; s < GRID_STAGES; s++) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, warps);
        }
        mbar_init(lut_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if (!WIDE) {
            const unsigned table_bytes = (unsigned) ((lutx_count + luty_count) * sizeof(float2));
            mbar_expect_tx(lut_bar, table_bytes);
            bulk_copy_g2s(lutx, tables, table_bytes, lut_bar);
        }
        for (int s = 0; s < GRID_STAGES && s < my_batches; s++) {
            mbar_expect_tx(full + s, stage_bytes);
            bulk_copy_g2s(ring + s * stage_bytes, chunk_of(s), stage_bytes, full + s);
        }
    }
    __syncthreads();        // barriers initialised
    if (!WIDE) mbar_wait(lut_bar, 0);  // tables have landed

    const int g = tid / prm.group_size;
    const int q = tid - g * prm.group_size;
    const int ty = q / prm.tx;
    const int tx = q - ty * prm.tx;
    // Tap tables: shared memory (narrow) or the global copy, read through L1 (wide)
    const unsigned char *const lut_bytes = WIDE ? reinterpret_cast<const unsigned char *>(tables)
                                                : reinterpret_cast<const unsigned char *>(lutx);

    // Per-thread slot constants
    int xoff[MX], yoff[MY];
    Mask my_xmask = 0, my_ymask = 0;
#pragma unroll
    for (int i = 0; i < MX; i++) {
        const int s = TX1 ? i : tx + i * prm.tx;
        xoff[i] = s * (int) sizeof(float2);
        my_xmask |= (Mask) 1 << s;
    }
#pragma unroll
    for (int j = 0; j < MY; j++) {
        const int s = ty + j * prm.ty;
        yoff[j] = s * (int) sizeof(float2);
        my_ymask |= (Mask) 1 << s;
    }

    Complex acc[MY][MX][P];
#pragma unroll
    for (int j = 0; j < MY; j++)
#pragma unroll
        for (int i = 0; i < MX; i++)
#pragma unroll
            for (int p = 0; p < P; p++) {
                acc[j][i][p].x = 0;
                acc[j][i][p].y = 0;
            }
    int pu0 = 0, pv0 = 0;       // footprint origin at this thread's last flush check
    bool valid = false;
    Complex *const grid = static_cast<Complex *>(prm.grid);

    // Flush every accumulator selected by the masks to its cell under origin (pu0, pv0).
    // The reductions sit in branches but only read the accumulators; the zeroing is a
    // branch-free select so that the accumulators never change registers.
    auto flush_cells = [&](Mask xm, Mask ym) {
        const int pru = pu0 - BX * (int) __umulhi((unsigned) pu0, prm.magic_x);
        const int prv = pv0 - BY * (int) __umulhi((unsigned) pv0, prm.magic_y);
        bool hit[MY][MX];
#pragma unroll
        for (int i = 0; i < MX; i++) {
            const int sx = xoff[i] / (int) sizeof(float2);
            const bool cx = (xm >> sx) & 1;
            int dx = sx - pru;
            if (dx < 0) dx += BX;
            const int col = pu0 + dx;
#pragma unroll
            for (int j = 0; j < MY; j++) {
                const int sy = yoff[j] / (int) sizeof(float2);
                hit[j][i] = cx || ((ym >> sy) & 1);
                if (hit[j][i]) {
                    int dy = sy - prv;
                    if (dy < 0) dy += BY;
                    const int row = pv0 + dy;
                    if (col < G && row < G) {
                        Complex *ptr = grid + ((unsigned) row * (unsigned) prm.grid_row_stride
                                               + (unsigned) col);
#pragma unroll
                        for (int p = 0; p < P; p++) {
                            Acc<Real>::flush(ptr, acc[j][i][p]);
                            ptr += prm.grid_pol_stride;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < MY; j++)
#pragma unroll
            for (int i = 0; i < MX; i++)
                if (hit[j][i]) {
#pragma unroll
                    for (int p = 0; p < P; p++) {
                        acc[j][i][p].x = 0;
                        acc[j][i][p].y = 0;
                    }
                }
    };

    const int rec_stride = gpb * REC;       // bytes between consecutive entries of a group
    for (int b = 0; b < my_batches; b++) {
        const int stage = b % GRID_STAGES;
        if (b != 0 && b % nbatches == 0 && valid) {
            // next work unit: unrelated visibilities, start from clean accumulators
            flush_cells(~(Mask) 0, ~(Mask) 0);
            valid = false;
        }
        mbar_wait(full + stage, (b / GRID_STAGES) & 1);
        const unsigned char *rec = ring + stage * stage_bytes + g * REC;
        // ---- process the batch.  The flush lives in the outer loop so that the inner
        // loop contains nothing but loads and FMAs on the register accumulators.
        int e = 0;
        // header of record `rec`: origin, table offsets (bytes) and this thread's move test
        int pos, uoff, voff;
        Mask xm, ym;
        auto read_header = [&](const unsigned char *r) {
            const int4 h = *reinterpret_cast<const int4 *>(r);
            pos = h.x;
            if (WIDE) {
                const int4 m = *reinterpret_cast<const int4 *>(r + 16);
                uoff = h.y * (int) sizeof(float2);
                voff = h.z * (int) sizeof(float2);
                xm = (Mask) (((unsigned long long) (unsigned) m.y << 32) | (unsigned) m.x);
                ym = (Mask) (((unsigned long long) (unsigned) m.w << 32) | (unsigned) m.z);
            } else {
                uoff = (h.y & 0xffff) * (int) sizeof(float2);
                voff = (int) ((unsigned) h.y >> 16) * (int) sizeof(float2);
                xm = (Mask) (unsigned) h.z;
                ym = (Mask) (unsigned) h.w;
            }
            return ((xm & my_xmask) | (ym & my_ymask)) != 0;
        };
#pragma unroll 1
        while (e < batch_len) {
            if (read_header(rec)) {
                if (valid) flush_cells(xm, ym);
                pu0 = pos & 0xffff;
                pv0 = (unsigned) pos >> 16;
                valid = true;
            }
#pragma unroll 1
            for (;;) {
                float2 s[P];
                if (P % 2 == 0) {
#pragma unroll
                    for (int p = 0; p < P; p += 2) {
                        const float4 two = *reinterpret_cast<const float4 *>(rec + HDR + 8 * p);
                        s[p] = make_float2(two.x, two.y);
                        s[p + 1] = make_float2(two.z, two.w);
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < P; p++)
                        s[p] = *reinterpret_cast<const float2 *>(rec + HDR + 8 * p);
                }
                const unsigned char *const urow = lut_bytes + uoff;
                const unsigned char *const vrow = lut_bytes + voff;
                float2 wu[MX], wv[MY];
#pragma unroll
                for (int i = 0; i < MX; i++)
                    wu[i] = WIDE ? __ldg(reinterpret_cast<const float2 *>(urow + xoff[i]))
                                 : *reinterpret_cast<const float2 *>(urow + xoff[i]);
#pragma unroll
                for (int j = 0; j < MY; j++)
                    wv[j] = WIDE ? __ldg(reinterpret_cast<const float2 *>(vrow + yoff[j]))
                                 : *reinterpret_cast<const float2 *>(vrow + yoff[j]);
#pragma unroll
                for (int j = 0; j < MY; j++)
#pragma unroll
                    for (int i = 0; i < MX; i++) {
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
                rec += rec_stride;
                if (++e >= batch_len) break;
                // leave the inner loop as a whole warp, so the lanes stay in lock step
                if (__ballot_sync(lanes, read_header(rec)) != 0)
                    break;
            }
        }
        // This warp is done with the stage.  Warps are not synchronised with each other:
        // thread 0 refills the stage of the *previous* batch once every warp has left it,
        // which leaves the copy two batches to land.
        __syncwarp(lanes);
        if ((tid & 31) == 0) mbar_arrive(empty + stage);
        if (tid == 0 && b >= 1 && b - 1 + GRID_STAGES < my_batches) {
            const int ps = (b - 1) % GRID_STAGES;
            mbar_wait(empty + ps, ((b - 1) / GRID_STAGES) & 1);
            mbar_expect_tx(full + ps, stage_bytes);
            bulk_copy_g2s(ring + ps * stage_bytes, chunk_of(b - 1 + GRID_STAGES), stage_bytes,
                          full + ps);
        }
    }
    if (valid) flush_cells(~(Mask) 0, ~(Mask) 0);
}

// Library-managed scratch for the staged records: grows on demand, one per (device,
// stream) so that imagers running on different command queues do not share it.
struct GridScratch {
    unsigned char *data = nullptr;
    size_t bytes = 0;
};
static std::mutex g_scratch_mutex;
static std::map<std::pair<int, cudaStream_t>, GridScratch> g_scratch;

static int get_scratch(size_t bytes, cudaStream_t stream, unsigned char **out)
{
    int dev = 0;
    KIB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    GridScratch &scratch = g_scratch[std::make_pair(dev, stream)];
    if (scratch.bytes < bytes) {
        if (scratch.data != nullptr) {
            KIB_CUDA(cudaStreamSynchronize(stream));
            KIB_CUDA(cudaFree(scratch.data));
            scratch.data = nullptr;
            scratch.bytes = 0;
        }
        const size_t want = bytes + bytes / 4;
        KIB_CUDA(cudaMalloc(&scratch.data, want));
        scratch.bytes = want;
    }
    *out = scratch.data;
    return 0;
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

__attribute__((unused)) static bool choose_config(int K, int P, int dtype, GridConfig *out)
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

template <typename Kernel>
static int enable_large_smem(Kernel kernel, size_t bytes)
{
    if (bytes > 48 * 1024)
        KIB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int) bytes));
    return 0;
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
    // that the final flush (MX*MY*P reds per thread) stays negligible.  Small launches
    // (sparsely populated W slices) use short runs: their cost is latency, not work.
    const long long blocks_per_sm = 2048 / (threads < 64 ? 64 : threads);
    const long long target_groups = (long long) sm_count() * 2 * blocks_per_sm * gpb;
    long long run = (prm.num_vis + target_groups - 1) / target_groups;
    if (run < 4 * GRID_BATCH) {
        // at least one full wave of blocks before runs get longer than one batch
        const long long one_wave = (long long) sm_count() * gpb;
        run = (prm.num_vis + one_wave - 1) / one_wave;
        if (run > 4 * GRID_BATCH) run = 4 * GRID_BATCH;
        if (run < GRID_BATCH) run = GRID_BATCH;
    }
    if (run > 8192) run = 8192;
    run = (run + GRID_BATCH - 1) / GRID_BATCH * GRID_BATCH;
    if (threads <= 256) {
        // A few waves of blocks: pick the run length (within -25 % .. +50 %) that leaves the
        // last wave fullest, so no SM idles while a nearly empty wave drains.
        const long long slots = (long long) sm_count() * 2;
        double best_fill = -1.0;
        long long best_run = run;
        for (long long r = run - run / 4; r <= run + run / 2; r += GRID_BATCH) {
            const long long rr = (r + GRID_BATCH - 1) / GRID_BATCH * GRID_BATCH;
            if (rr < GRID_BATCH) continue;
            const long long b = ((prm.num_vis + rr - 1) / rr + gpb - 1) / gpb;
            const long long waves = (b + slots - 1) / slots;
            const double fill = (double) b / (double) (waves * slots);
            if (waves > 8) { best_run = run; break; }       // many waves: tail is negligible
            if (fill > best_fill + 1e-9) {
                best_fill = fill;
                best_run = rr;
            }
        }
        run = best_run;
    }
    // Small launches (less than one wave of blocks) are bound by the length of the
    // sequential chain of one group, not by throughput: give every group a shorter run,
    // down to 4 visibilities, as long as the blocks still fit in one wave.
    int batch = GRID_TMA_BATCH;
    if (threads <= 256 && run == GRID_BATCH) {
        const long long slots = (long long) sm_count() * 2;
        for (int b = 4; b < GRID_TMA_BATCH; b *= 2) {
            const long long blocks_b = ((prm.num_vis + b - 1) / b + gpb - 1) / gpb;
            if (blocks_b <= slots) {
                batch = b;
                run = b;
                break;
            }
        }
    }
    prm.batch = batch;
    prm.run = (int) run;
    const long long groups = (prm.num_vis + run - 1) / run;
    const long long blocks = (groups + gpb - 1) / gpb;

    // Fast path.  Narrow: the doubled kernel table fits in shared memory and a slot mask
    // in 32 bits.  Wide: the table stays in global memory (read through L1) and masks are
    // 64 bits, which covers every support up to 64 cells per axis.
    const size_t rows = (size_t) prm.w_planes * prm.oversample;
    const size_t lut_bytes = rows * 2 * (prm.bx + (prm.bx == prm.by ? 0 : prm.by)) * sizeof(float2);
    const bool addressable = prm.grid_size < 65536
        && (long long) prm.grid_size * prm.grid_row_stride < (1ll << 32);
    const bool narrow = addressable && lut_bytes <= (size_t) GRID_LUT_SMEM_LIMIT
        && lut_bytes / sizeof(float2) < 65536 && prm.bx <= 32 && prm.by <= 32;
    const bool wide = addressable && !narrow && prm.bx <= 64 && prm.by <= 64
        && lut_bytes / sizeof(float2) < (1u << 28);
    if (narrow || wide) {
        const int rec = grid_record_bytes(P, wide);
        const size_t stage_bytes = (size_t) prm.batch * gpb * rec;
        const size_t smem = 128 + GRID_STAGES * stage_bytes + (wide ? 0 : lut_bytes);
        const long long total = blocks * gpb * run;
        unsigned char *scratch = nullptr;
        const size_t table_bytes = (lut_bytes + 127) / 128 * 128;
        int rc = get_scratch(table_bytes + (size_t) total * rec, stream, &scratch);
        if (rc != 0) return rc;
        unsigned char *records = scratch + table_bytes;
        float2 *tables = reinterpret_cast<float2 *>(scratch);
        GridStageParams sp;
        sp.records = records;
        sp.weights_grid = prm.weights_grid;
        sp.uv = prm.uv;
        sp.w_plane = prm.w_plane;
        sp.vis = prm.vis;
        sp.num_rejected = prm.num_rejected;
        sp.weights_pol_stride = prm.weights_pol_stride;
        sp.num_vis = prm.num_vis;
        sp.total = total;
        sp.weights_row_stride = prm.weights_row_stride;
        sp.grid_size = prm.grid_size;
        sp.w_planes = prm.w_planes;
        sp.oversample = prm.oversample;
        sp.kernel_width = prm.kernel_width;
        sp.uv_bias = prm.uv_bias;
        sp.half_grid = prm.half_grid;
        sp.bx = prm.bx;
        sp.by = prm.by;
        sp.groups_per_block = gpb;
        sp.run = prm.run;
        sp.batch = prm.batch;
        sp.lutv_base = prm.bx == prm.by ? 0 : (int) (rows * 2 * prm.bx);
        sp.tables = tables;
        sp.lut = prm.lut;
        sp.lut_slice_stride = prm.lut_slice_stride;
        sp.lut_tap_offset = prm.lut_tap_offset;
        sp.lutx_count = (int) (rows * 2 * prm.bx);
        sp.luty_count = prm.bx == prm.by ? 0 : (int) (rows * 2 * prm.by);
        const long long stage_threads = total > sp.lutx_count + sp.luty_count
            ? total : sp.lutx_count + sp.luty_count;
        const unsigned stage_blocks = (unsigned) ((stage_threads + 255) / 256);
        if (wide)
            grid_stage_kernel<P, true><<<stage_blocks, 256, 0, stream>>>(sp);
        else
            grid_stage_kernel<P, false><<<stage_blocks, 256, 0, stream>>>(sp);
        prm.num_units = (int) blocks;
        // One unit per block: the hardware block scheduler balances units of unequal cost
        // (the kernel itself also supports fewer, persistent blocks).
        const unsigned launch_blocks = (unsigned) blocks;
#define KIB_LAUNCH_TMA(MAXT, MINB, TX1, WIDE)                                                \
        do {                                                                                 \
            auto kernel = grid_tma_kernel<Real, P, MX, MY, MAXT, MINB, TX1, WIDE>;            \
            rc = enable_large_smem(kernel, smem);                                            \
            if (rc != 0) return rc;                                                          \
            kernel<<<launch_blocks, threads, smem, stream>>>(prm, records, tables);          \
        } while (0)
        if (wide) {
            if (threads <= 256) KIB_LAUNCH_TMA(256, 2, false, true);
            else KIB_LAUNCH_TMA(GRID_MAX_GROUP, 1, false, true);
        } else if (threads <= 256) {
            if (prm.tx == 1) KIB_LAUNCH_TMA(256, 2, true, false);
            else KIB_LAUNCH_TMA(256, 2, false, false);
        } else {
            KIB_LAUNCH_TMA(GRID_MAX_GROUP, 1, false, false);
        }
#undef KIB_LAUNCH_TMA
        KIB_CHECK_LAUNCH();
        return 0;
    }

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

// The file is compiled twice (Makefile: -DKIB_GRID_PART=1 single precision + entry point,
// -DKIB_GRID_PART=2 double precision) so that the two sets of kernel instantiations build in
// parallel.
#ifndef KIB_GRID_PART
#define KIB_GRID_PART 1
#endif
int grid_dispatch_f32(GridParams &prm, const GridConfig &cfg, int P, cudaStream_t stream);
int grid_dispatch_f64(GridParams &prm, const GridConfig &cfg, int P, cudaStream_t stream);
#if KIB_GRID_PART == 1
int grid_dispatch_f32(GridParams &prm, const GridConfig &cfg, int P, cudaStream_t stream)
{
    return dispatch_pols<float>(prm, cfg, P, KIB_F32, stream);
}
#else
int grid_dispatch_f64(GridParams &prm, const GridConfig &cfg, int P, cudaStream_t stream)
{
    return dispatch_pols<double>(prm, cfg, P, KIB_F64, stream);
}
#endif

}  // namespace kib

#if KIB_GRID_PART == 1
namespace kib {
// Called by kib_stream_destroy: the staging scratch of a stream dies with it (a recycled
// stream handle must not inherit the buffer, and the memory must not leak).
void release_grid_scratch(cudaStream_t stream)
{
    std::lock_guard<std::mutex> lock(g_scratch_mutex);
    for (auto it = g_scratch.begin(); it != g_scratch.end();) {
        if (it->first.second == stream) {
            if (it->second.data != nullptr) cudaFree(it->second.data);
            it = g_scratch.erase(it);
        } else {
            ++it;
        }
    }
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
    prm.magic_x = (unsigned) ((1ull << 32) / prm.bx) + 1;
    prm.magic_y = (unsigned) ((1ull << 32) / prm.by) + 1;
    cudaStream_t s = as_stream(stream);
    if (dtype == KIB_F32) return grid_dispatch_f32(prm, cfg, num_pols, s);
    return grid_dispatch_f64(prm, cfg, num_pols, s);
}
#endif  // KIB_GRID_PART == 1
