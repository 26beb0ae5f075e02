// Visibility preprocessing on the device.
//
// Replaces visibility_collector<P>::add_impl2 and ::compress of the reference's only native
// component (katsdpimager/preprocess.cpp:401-513 and :335-397; subpixel_coord :313-323;
// mueller_generator_simple / _parallactic :136-182; MulZ arithmetic mulz.h:14-50) for one
// channel: Mueller / Stokes transform of the correlation products, weight transform, w < 0
// flip with conjugation, UVW quantisation to (cell, sub-cell, W plane, W slice), removal of
// flagged samples, merging of adjacent samples that fall on the same coordinates (within each
// buffer of `capacity` samples, as the host code does), and a stable bucket sort by W slice.
// The output is the record array the gridder's unpack kernel consumes, already in HBM: the raw
// visibilities cross PCIe once and nothing derived from them ever does.
//
//   quantise_kernel   one thread per sample -> staged records (flagged samples all-zero)
//   head_kernel       first sample of every run of equal coordinates (flagged ones skipped)
//   cub exclusive sum output position of every run
//   merge_kernel      one thread per run sums its samples in input order (the host's order:
//                     float sums are bit-identical)
//   cub radix sort    stable sort of run indices by W slice
//   gather_kernel     records in bucket order + per-slice counts
// All arithmetic that decides a coordinate or a value is single precision with separately
// rounded multiplies and adds, as the host's scalar C++ is.
#include "kib_common.cuh"
#include <cub/cub.cuh>

namespace kib {
namespace prep {

struct Config {
    const float *uvw;
    const float *weights;
    const float2 *vis;
    const float *feed1, *feed2;
    const float2 *stokes;
    const float2 *circular;
    long long n;
    long long capacity;
    float uv_scale, w_scale, w_bias;
    int Q, w_planes, oversample, max_slice_plane;
};

__device__ __forceinline__ float2 cmul_rn(float2 a, float2 b)
{
    return make_float2(__fsub_rn(__fmul_rn(a.x, b.x), __fmul_rn(a.y, b.y)),
                       __fadd_rn(__fmul_rn(a.x, b.y), __fmul_rn(a.y, b.x)));
}

// MulZ: a product with an exact zero is zero, whatever the other factor (mulz.h:38-41)
__device__ __forceinline__ float2 mulz(float2 a, float2 b)
{
    const bool nz = (a.x != 0.0f || a.y != 0.0f) && (b.x != 0.0f || b.y != 0.0f);
    return nz ? cmul_rn(a, b) : make_float2(0.0f, 0.0f);
}
__device__ __forceinline__ float mulz(float a, float b)
{
    return (a != 0.0f && b != 0.0f) ? __fmul_rn(a, b) : 0.0f;
}

__device__ __forceinline__ void subpixel_coord(float x, int oversample, short *pixel, short *sub)
{
    const int xs = (int) floorf(__fmul_rn(x, (float) oversample));
    int p = xs / oversample, s = xs % oversample;
    if (s < 0) {
        p--;
        s += oversample;
    }
    *pixel = (short) p;
    *sub = (short) s;
}

template <int P>
__global__ void __launch_bounds__(256)
quantise_kernel(const Config cfg, unsigned *__restrict__ staged)
{
    constexpr int WORDS = 3 + 3 * P;
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cfg.n) return;
    unsigned *out = staged + i * WORDS;
    const int Q = cfg.Q;
    float wt[4];
    float2 v[4];
    bool flagged = false;
    for (int q = 0; q < Q; q++) {
        wt[q] = cfg.weights[i * Q + q];
        v[q] = cfg.vis[i * Q + q];
        flagged |= wt[q] == 0.0f;
    }
    if (flagged) {
        for (int k = 0; k < WORDS; k++) out[k] = 0;
        return;
    }
    float2 conv[P][4];
    if (cfg.feed1 == nullptr) {
        for (int p = 0; p < P; p++)
            for (int q = 0; q < Q; q++) conv[p][q] = cfg.stokes[p * Q + q];
    } else {
        float s1, c1, s2, c2;
        sincosf(cfg.feed1[i], &s1, &c1);
        sincosf(cfg.feed2[i], &s2, &c2);
        const float2 rr = cmul_rn(make_float2(c1, s1), make_float2(c2, -s2));
        const float2 rl = cmul_rn(make_float2(c1, s1), make_float2(c2, s2));
        const float2 scale[4] = {rr, rl, make_float2(rl.x, -rl.y), make_float2(rr.x, -rr.y)};
        for (int p = 0; p < P; p++)
            for (int q = 0; q < Q; q++) {
                float2 acc = make_float2(0.0f, 0.0f);
                for (int r = 0; r < 4; r++) {
                    const float2 m = cmul_rn(cfg.circular[r * Q + q], scale[r]);
                    const float2 t = cmul_rn(cfg.stokes[p * 4 + r], m);
                    acc.x = __fadd_rn(acc.x, t.x);
                    acc.y = __fadd_rn(acc.y, t.y);
                }
                conv[p][q] = acc;
            }
    }
    float u = cfg.uvw[3 * i], vv = cfg.uvw[3 * i + 1], w = cfg.uvw[3 * i + 2];
    const bool flip = w < 0.0f;
    if (flip) {
        u = -u;
        vv = -vv;
        w = -w;
    }
    for (int p = 0; p < P; p++) {
        float2 acc = make_float2(0.0f, 0.0f);
        float var = 0.0f;
        for (int q = 0; q < Q; q++) {
            const float2 t = mulz(conv[p][q], v[q]);
            acc.x = __fadd_rn(acc.x, t.x);
            acc.y = __fadd_rn(acc.y, t.y);
            const float abs2 = __fadd_rn(__fmul_rn(conv[p][q].x, conv[p][q].x),
                                         __fmul_rn(conv[p][q].y, conv[p][q].y));
            var = __fadd_rn(var, mulz(abs2, __fdiv_rn(1.0f, fabsf(wt[q]))));
        }
        float weight = __fdiv_rn(1.0f, var);
        if (flip) acc.y = -acc.y;
        acc.x = __fmul_rn(acc.x, weight);
        acc.y = __fmul_rn(acc.y, weight);
        if (!isfinite(acc.x) || !isfinite(acc.y)) {
            // NaNs from calibration failures are squashed (preprocess.cpp:488-494)
            acc = make_float2(0.0f, 0.0f);
            weight = 0.0f;
        }
        out[3 + p] = __float_as_uint(weight);
        out[3 + P + 2 * p] = __float_as_uint(acc.x);
        out[4 + P + 2 * p] = __float_as_uint(acc.y);
    }
    u = __fmul_rn(u, cfg.uv_scale);
    vv = __fmul_rn(vv, cfg.uv_scale);
    // the plane number is biased by half a slice: the first slice is half-width, centred at w = 0
    w = truncf(__fadd_rn(__fmul_rn(w, cfg.w_scale), cfg.w_bias));
    int sp = (int) w;
    if (sp > cfg.max_slice_plane) sp = cfg.max_slice_plane;
    short pu, su, pv, sv;
    subpixel_coord(u, cfg.oversample, &pu, &su);
    subpixel_coord(vv, cfg.oversample, &pv, &sv);
    const unsigned w_plane = (unsigned) (sp % cfg.w_planes), w_slice = (unsigned) (sp / cfg.w_planes);
    out[0] = ((unsigned) (unsigned short) pu) | ((unsigned) (unsigned short) pv << 16);
    out[1] = ((unsigned) (unsigned short) su) | ((unsigned) (unsigned short) sv << 16);
    out[2] = w_plane | (w_slice << 16);
}

__device__ __forceinline__ bool is_valid(const unsigned *rec) { return __uint_as_float(rec[3]) != 0.0f; }
__device__ __forceinline__ bool same_key(const unsigned *a, const unsigned *b)
{
    return a[0] == b[0] && a[1] == b[1] && a[2] == b[2];
}

// head[i] = 1 if sample i starts a run: it is valid and the previous valid sample of its buffer
// (flagged ones in between are skipped, preprocess.cpp:349-352) has other coordinates.
__global__ void __launch_bounds__(256)
head_kernel(const unsigned *__restrict__ staged, int words, long long n, long long capacity,
            int *__restrict__ head)
{
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned *rec = staged + i * words;
    int h = 0;
    if (is_valid(rec)) {
        h = 1;
        const long long first = i - i % capacity;
        for (long long j = i - 1; j >= first; j--) {
            const unsigned *prev = staged + j * words;
            if (is_valid(prev)) {
                h = !same_key(prev, rec);
                break;
            }
        }
    }
    head[i] = h;
}

template <int P>
__global__ void __launch_bounds__(256)
merge_kernel(const unsigned *__restrict__ staged, const int *__restrict__ head,
             const int *__restrict__ pos, long long n, long long capacity,
             unsigned *__restrict__ merged, unsigned short *__restrict__ keys,
             unsigned *__restrict__ index)
{
    constexpr int WORDS = 3 + 3 * P;
    const long long i = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    const unsigned *rec = staged + i * WORDS;
    float acc[3 * P];
    for (int k = 0; k < 3 * P; k++) acc[k] = __uint_as_float(rec[3 + k]);
    const long long end = min(n, i - i % capacity + capacity);
    for (long long j = i + 1; j < end; j++) {
        if (head[j]) break;
        const unsigned *e = staged + j * WORDS;
        if (!is_valid(e)) continue;
        // the host adds the visibilities, then the weights, element by element in input order
        for (int k = 0; k < 3 * P; k++) acc[k] = __fadd_rn(acc[k], __uint_as_float(e[3 + k]));
    }
    const int o = pos[i];
    unsigned *out = merged + (long long) o * WORDS;
    out[0] = rec[0];
    out[1] = rec[1];
    out[2] = rec[2];
    for (int k = 0; k < 3 * P; k++) out[3 + k] = __float_as_uint(acc[k]);
    keys[o] = (unsigned short) (rec[2] >> 16);
    index[o] = (unsigned) o;
}

__global__ void __launch_bounds__(256)
gather_kernel(const unsigned *__restrict__ merged, const unsigned *__restrict__ order,
              const unsigned short *__restrict__ sorted_keys, int words, long long total,
              unsigned *__restrict__ out, long long *__restrict__ counts)
{
    const long long k = (long long) blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const unsigned *src = merged + (long long) order[k] * words;
    unsigned *dst = out + k * words;
    for (int w = 0; w < words; w++) dst[w] = src[w];
    // one atomic per boundary: the sorted keys change w_slices - 1 times at most
    const unsigned short key = sorted_keys[k];
    if (k + 1 == total || sorted_keys[k + 1] != key)
        counts[key] = k + 1;            // end of the slice's run; turned into counts on the host
}

struct Layout {
    size_t staged, merged, head, pos, keys, keys_out, index, index_out, counts, total, temp;
    size_t temp_bytes, bytes;
};

static int layout(long long n, int P, int w_slices, Layout *l)
{
    const size_t words = 3 + 3 * P;
    auto align = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t scan_bytes = 0, sort_bytes = 0;
    KIB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (const int *) nullptr,
                                           (int *) nullptr, (int) n));
    KIB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const unsigned short *) nullptr,
                                             (unsigned short *) nullptr, (const unsigned *) nullptr,
                                             (unsigned *) nullptr, (int) n, 0, 16));
    size_t off = 0;
    l->staged = off; off += align(n * words * 4);
    l->merged = off; off += align(n * words * 4);
    l->head = off; off += align(n * 4);
    l->pos = off; off += align(n * 4);
    l->keys = off; off += align(n * 2);
    l->keys_out = off; off += align(n * 2);
    l->index = off; off += align(n * 4);
    l->index_out = off; off += align(n * 4);
    l->counts = off; off += align((size_t) (w_slices + 1) * 8);
    l->total = off; off += 256;
    l->temp_bytes = scan_bytes > sort_bytes ? scan_bytes : sort_bytes;
    l->temp = off; off += align(l->temp_bytes);
    l->bytes = off;
    return 0;
}

}  // namespace prep
}  // namespace kib

using namespace kib;
using namespace kib::prep;

extern "C" {

int kib_preprocess_scratch_bytes(int64_t num_vis, int num_pols, int w_slices, int64_t *bytes)
{
    KIB_REQUIRE(bytes != nullptr && num_vis >= 0 && num_vis < (1ll << 31) && num_pols >= 1
                && num_pols <= 4 && w_slices >= 1, "kib_preprocess_scratch_bytes: bad arguments");
    Layout l;
    if (int rc = layout(num_vis > 0 ? num_vis : 1, num_pols, w_slices, &l)) return rc;
    *bytes = (int64_t) l.bytes;
    return 0;
}

int kib_preprocess(const float *uvw, const float *weights, const void *vis, int64_t num_vis,
                   int num_in_pols, const float *feed_angle1, const float *feed_angle2,
                   const void *mueller_stokes, const void *mueller_circular, int num_pols,
                   double cell_size, double max_w, int w_slices, int w_planes, int oversample,
                   int64_t capacity, void *records, int64_t *host_counts,
                   void *scratch, int64_t scratch_bytes, kib_stream_t stream)
{
    KIB_REQUIRE(num_pols >= 1 && num_pols <= 4 && num_in_pols >= 1 && num_in_pols <= 4,
                "kib_preprocess: polarization counts must be 1..4");
    KIB_REQUIRE((feed_angle1 == nullptr) == (feed_angle2 == nullptr)
                && (feed_angle1 == nullptr) == (mueller_circular == nullptr),
                "kib_preprocess: feed angles and mueller_circular go together");
    KIB_REQUIRE(w_slices >= 1 && w_slices < 65536 && w_planes >= 1 && oversample >= 1,
                "kib_preprocess: bad grid parameters");
    KIB_REQUIRE(num_vis >= 0 && num_vis < (1ll << 31), "kib_preprocess: too many samples");
    KIB_REQUIRE(host_counts != nullptr, "kib_preprocess: null counts");
    for (int s = 0; s < w_slices; s++) host_counts[s] = 0;
    if (num_vis == 0) return 0;
    Layout l;
    if (int rc = layout(num_vis, num_pols, w_slices, &l)) return rc;
    KIB_REQUIRE(scratch != nullptr && scratch_bytes >= (int64_t) l.bytes,
                "kib_preprocess: scratch of %lld bytes needed, %lld given",
                (long long) l.bytes, (long long) scratch_bytes);
    cudaStream_t s = as_stream(stream);
    char *base = static_cast<char *>(scratch);
    unsigned *staged = reinterpret_cast<unsigned *>(base + l.staged);
    unsigned *merged = reinterpret_cast<unsigned *>(base + l.merged);
    int *head = reinterpret_cast<int *>(base + l.head);
    int *pos = reinterpret_cast<int *>(base + l.pos);
    unsigned short *keys = reinterpret_cast<unsigned short *>(base + l.keys);
    unsigned short *keys_out = reinterpret_cast<unsigned short *>(base + l.keys_out);
    unsigned *index = reinterpret_cast<unsigned *>(base + l.index);
    unsigned *index_out = reinterpret_cast<unsigned *>(base + l.index_out);
    long long *counts = reinterpret_cast<long long *>(base + l.counts);
    void *temp = base + l.temp;

    Config cfg;
    cfg.uvw = uvw;
    cfg.weights = weights;
    cfg.vis = static_cast<const float2 *>(vis);
    cfg.feed1 = feed_angle1;
    cfg.feed2 = feed_angle2;
    cfg.stokes = static_cast<const float2 *>(mueller_stokes);
    cfg.circular = static_cast<const float2 *>(mueller_circular);
    cfg.n = num_vis;
    cfg.capacity = capacity > 0 ? capacity : num_vis;
    cfg.uv_scale = 1.0f / (float) cell_size;
    cfg.w_scale = ((float) w_slices - 0.5f) * (float) w_planes / (float) max_w;
    cfg.w_bias = (float) w_planes * 0.5f;
    cfg.Q = num_in_pols;
    cfg.w_planes = w_planes;
    cfg.oversample = oversample;
    cfg.max_slice_plane = w_slices * w_planes - 1;
    const int words = 3 + 3 * num_pols;
    const unsigned blocks = (unsigned) ((num_vis + 255) / 256);
    switch (num_pols) {
    case 1: quantise_kernel<1><<<blocks, 256, 0, s>>>(cfg, staged); break;
    case 2: quantise_kernel<2><<<blocks, 256, 0, s>>>(cfg, staged); break;
    case 3: quantise_kernel<3><<<blocks, 256, 0, s>>>(cfg, staged); break;
    default: quantise_kernel<4><<<blocks, 256, 0, s>>>(cfg, staged); break;
    }
    KIB_CHECK_LAUNCH();
    head_kernel<<<blocks, 256, 0, s>>>(staged, words, num_vis, cfg.capacity, head);
    KIB_CHECK_LAUNCH();
    size_t temp_bytes = l.temp_bytes;
    KIB_CUDA(cub::DeviceScan::ExclusiveSum(temp, temp_bytes, head, pos, (int) num_vis, s));
    switch (num_pols) {
    case 1: merge_kernel<1><<<blocks, 256, 0, s>>>(staged, head, pos, num_vis, cfg.capacity, merged, keys, index); break;
    case 2: merge_kernel<2><<<blocks, 256, 0, s>>>(staged, head, pos, num_vis, cfg.capacity, merged, keys, index); break;
    case 3: merge_kernel<3><<<blocks, 256, 0, s>>>(staged, head, pos, num_vis, cfg.capacity, merged, keys, index); break;
    default: merge_kernel<4><<<blocks, 256, 0, s>>>(staged, head, pos, num_vis, cfg.capacity, merged, keys, index); break;
    }
    KIB_CHECK_LAUNCH();
    // number of runs = pos[n - 1] + head[n - 1]
    int last_pos = 0, last_head = 0;
    KIB_CUDA(cudaMemcpyAsync(&last_pos, pos + num_vis - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    KIB_CUDA(cudaMemcpyAsync(&last_head, head + num_vis - 1, sizeof(int), cudaMemcpyDeviceToHost, s));
    KIB_CUDA(cudaStreamSynchronize(s));
    const long long total = (long long) last_pos + last_head;
    if (total == 0) return 0;
    int bits = 1;
    while ((1 << bits) < w_slices) bits++;
    temp_bytes = l.temp_bytes;
    KIB_CUDA(cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys, keys_out, index, index_out,
                                             (int) total, 0, bits, s));
    KIB_CUDA(cudaMemsetAsync(counts, 0, sizeof(long long) * (w_slices + 1), s));
    gather_kernel<<<(unsigned) ((total + 255) / 256), 256, 0, s>>>(
        merged, index_out, keys_out, words, total, static_cast<unsigned *>(records), counts);
    KIB_CHECK_LAUNCH();
    KIB_CUDA(cudaMemcpyAsync(host_counts, counts, sizeof(long long) * w_slices,
                             cudaMemcpyDeviceToHost, s));
    KIB_CUDA(cudaStreamSynchronize(s));
    // run ends -> counts (slices without records have end 0)
    long long previous = 0;
    for (int k = 0; k < w_slices; k++) {
        const long long end = host_counts[k];
        host_counts[k] = end > 0 ? end - previous : 0;
        if (end > 0) previous = end;
    }
    return 0;
}

}  // extern "C"
