// Runtime layer of libkatimager_b200.so: devices, streams, events, memory, cuFFT plans.
// Replaces the katsdpsigproc.accel/PyCUDA services the reference operations use
// (SURVEY.md section 8b.2); see include/katimager_b200.h.
#include "kib_common.cuh"
#include <cufft.h>
#include <cstring>
#include <string>

namespace kib {

static thread_local char g_error[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

struct FftPlan {
    cufftHandle handle;
    int dtype;
    int kind = 0;       // 0 complex <-> complex, 1 real -> complex, 2 complex -> real
};

static const char *cufft_error_string(cufftResult r)
{
    switch (r) {
    case CUFFT_SUCCESS: return "CUFFT_SUCCESS";
    case CUFFT_INVALID_PLAN: return "CUFFT_INVALID_PLAN";
    case CUFFT_ALLOC_FAILED: return "CUFFT_ALLOC_FAILED";
    case CUFFT_INVALID_TYPE: return "CUFFT_INVALID_TYPE";
    case CUFFT_INVALID_VALUE: return "CUFFT_INVALID_VALUE";
    case CUFFT_INTERNAL_ERROR: return "CUFFT_INTERNAL_ERROR";
    case CUFFT_EXEC_FAILED: return "CUFFT_EXEC_FAILED";
    case CUFFT_SETUP_FAILED: return "CUFFT_SETUP_FAILED";
    case CUFFT_INVALID_SIZE: return "CUFFT_INVALID_SIZE";
    default: return "CUFFT error";
    }
}

#define KIB_CUFFT(expr)                                                             \
    do {                                                                            \
        cufftResult kib_r__ = (expr);                                               \
        if (kib_r__ != CUFFT_SUCCESS) {                                             \
            ::kib::set_error("%s failed: %s (%d)", #expr,                           \
                             ::kib::cufft_error_string(kib_r__), (int) kib_r__);    \
            return 1000 + (int) kib_r__;                                            \
        }                                                                           \
    } while (0)

// Small device-to-host reads (scalars, peaks, histograms: what the host waits for between the
// stages of a channel) are written into the pinned destination by a kernel instead of going
// through the copy engine: the engine serves its queue in order, so a read issued while the
// previous channel's image (1 GB) is on its way out would wait for all of it -- 20 ms per
// channel in bench.py's end-to-end leg.
constexpr size_t SMALL_D2H_BYTES = 256 * 1024;

template <typename T>
__global__ void __launch_bounds__(256)
small_d2h_kernel(T *__restrict__ dst, size_t dst_row_pitch, size_t dst_plane_pitch,
                 const T *__restrict__ src, size_t src_row_pitch, size_t src_plane_pitch,
                 size_t width, size_t height, size_t depth)
{
    // pitches and width in units of T
    const size_t total = width * height * depth;
    for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t) gridDim.x * blockDim.x) {
        const size_t x = i % width, r = i / width;
        const size_t y = r % height, z = r / height;
        dst[z * dst_plane_pitch + y * dst_row_pitch + x] = src[z * src_plane_pitch + y * src_row_pitch + x];
    }
}

// Device-visible address of `host` if it is page-locked memory the device can write, else null.
static void *mapped_pointer(const void *host)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host) != cudaSuccess) {
        (void) cudaGetLastError();
        return nullptr;
    }
    return attr.type == cudaMemoryTypeHost ? attr.devicePointer : nullptr;
}

// Returns 1 if the copy was not taken (too large, destination not mapped).
static int small_d2h(void *dst, size_t dst_row_pitch, size_t dst_plane_pitch,
                     const void *src, size_t src_row_pitch, size_t src_plane_pitch,
                     size_t width_bytes, size_t height, size_t depth, cudaStream_t stream)
{
    const size_t total = width_bytes * height * depth;
    if (total > SMALL_D2H_BYTES) return 1;
    void *mapped = mapped_pointer(dst);
    if (mapped == nullptr) return 1;
    // whole rows of a few planes lie inside one pinned allocation; the last byte as a check
    if (mapped_pointer(static_cast<const char *>(dst) + (depth - 1) * dst_plane_pitch
                       + (height - 1) * dst_row_pitch + width_bytes - 1) == nullptr)
        return 1;
    const bool words = ((reinterpret_cast<size_t>(mapped) | reinterpret_cast<size_t>(src)
                         | dst_row_pitch | dst_plane_pitch | src_row_pitch | src_plane_pitch
                         | width_bytes) & 3) == 0;
    const size_t elems = words ? total / 4 : total;
    const unsigned blocks = (unsigned) ((elems + 255) / 256 > 64 ? 64 : (elems + 255) / 256);
    if (words)
        small_d2h_kernel<unsigned><<<blocks, 256, 0, stream>>>(
            static_cast<unsigned *>(mapped), dst_row_pitch / 4, dst_plane_pitch / 4,
            static_cast<const unsigned *>(src), src_row_pitch / 4, src_plane_pitch / 4,
            width_bytes / 4, height, depth);
    else
        small_d2h_kernel<unsigned char><<<blocks, 256, 0, stream>>>(
            static_cast<unsigned char *>(mapped), dst_row_pitch, dst_plane_pitch,
            static_cast<const unsigned char *>(src), src_row_pitch, src_plane_pitch,
            width_bytes, height, depth);
    KIB_CHECK_LAUNCH();
    return 0;
}

}  // namespace kib

using namespace kib;

extern "C" {

int kib_version(void) { return KIB_VERSION; }

const char *kib_last_error(void) { return g_error; }

int kib_device_count(int *count)
{
    KIB_REQUIRE(count != nullptr, "kib_device_count: null argument");
    KIB_CUDA(cudaGetDeviceCount(count));
    return 0;
}

int kib_set_device(int device)
{
    KIB_CUDA(cudaSetDevice(device));
    return 0;
}

int kib_get_device(int *device)
{
    KIB_REQUIRE(device != nullptr, "kib_get_device: null argument");
    KIB_CUDA(cudaGetDevice(device));
    return 0;
}

int kib_device_name(int device, char *buf, int buf_len)
{
    KIB_REQUIRE(buf != nullptr && buf_len > 0, "kib_device_name: bad buffer");
    cudaDeviceProp prop;
    KIB_CUDA(cudaGetDeviceProperties(&prop, device));
    strncpy(buf, prop.name, buf_len - 1);
    buf[buf_len - 1] = '\0';
    return 0;
}

int kib_device_attr(int device, int attr, int64_t *value)
{
    KIB_REQUIRE(value != nullptr, "kib_device_attr: null argument");
    int v = 0;
    switch (attr) {
    case 0: KIB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device)); break;
    case 1: KIB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)); break;
    case 2: KIB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrL2CacheSize, device)); break;
    case 3: KIB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device)); break;
    case 4: {
        int major = 0, minor = 0;
        KIB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
        KIB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
        v = major * 10 + minor;
        break;
    }
    case 5: KIB_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrWarpSize, device)); break;
    default:
        set_error("kib_device_attr: unknown attribute %d", attr);
        return -1;
    }
    *value = v;
    return 0;
}

int kib_mem_info(size_t *free_bytes, size_t *total_bytes)
{
    KIB_REQUIRE(free_bytes != nullptr && total_bytes != nullptr, "kib_mem_info: null argument");
    KIB_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
    return 0;
}

int kib_stream_create(kib_stream_t *stream)
{
    KIB_REQUIRE(stream != nullptr, "kib_stream_create: null argument");
    cudaStream_t s;
    KIB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = reinterpret_cast<kib_stream_t>(s);
    return 0;
}

int kib_stream_destroy(kib_stream_t stream)
{
    if (stream != nullptr) {
        KIB_CUDA(cudaStreamSynchronize(as_stream(stream)));
        release_grid_scratch(as_stream(stream));
        KIB_CUDA(cudaStreamDestroy(as_stream(stream)));
    }
    return 0;
}

int kib_stream_sync(kib_stream_t stream)
{
    KIB_CUDA(cudaStreamSynchronize(as_stream(stream)));
    return 0;
}

int kib_stream_wait_event(kib_stream_t stream, kib_event_t event)
{
    KIB_CUDA(cudaStreamWaitEvent(as_stream(stream), reinterpret_cast<cudaEvent_t>(event), 0));
    return 0;
}

int kib_event_create(kib_event_t *event)
{
    KIB_REQUIRE(event != nullptr, "kib_event_create: null argument");
    cudaEvent_t e;
    KIB_CUDA(cudaEventCreate(&e));
    *event = reinterpret_cast<kib_event_t>(e);
    return 0;
}

int kib_event_record(kib_event_t event, kib_stream_t stream)
{
    KIB_CUDA(cudaEventRecord(reinterpret_cast<cudaEvent_t>(event), as_stream(stream)));
    return 0;
}

int kib_event_sync(kib_event_t event)
{
    KIB_CUDA(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(event)));
    return 0;
}

int kib_event_query(kib_event_t event, int *done)
{
    KIB_REQUIRE(done != nullptr, "kib_event_query: null argument");
    cudaError_t err = cudaEventQuery(reinterpret_cast<cudaEvent_t>(event));
    if (err == cudaSuccess) { *done = 1; return 0; }
    if (err == cudaErrorNotReady) { *done = 0; return 0; }
    set_error("cudaEventQuery failed: %s", cudaGetErrorString(err));
    return (int) err;
}

int kib_event_elapsed_ms(kib_event_t start, kib_event_t stop, float *ms)
{
    KIB_REQUIRE(ms != nullptr, "kib_event_elapsed_ms: null argument");
    KIB_CUDA(cudaEventElapsedTime(ms, reinterpret_cast<cudaEvent_t>(start),
                                  reinterpret_cast<cudaEvent_t>(stop)));
    return 0;
}

int kib_event_destroy(kib_event_t event)
{
    if (event != nullptr) KIB_CUDA(cudaEventDestroy(reinterpret_cast<cudaEvent_t>(event)));
    return 0;
}

int kib_malloc(void **ptr, size_t bytes)
{
    KIB_REQUIRE(ptr != nullptr, "kib_malloc: null argument");
    if (bytes == 0) bytes = 1;
    KIB_CUDA(cudaMalloc(ptr, bytes));
    return 0;
}

int kib_free(void *ptr)
{
    if (ptr != nullptr) KIB_CUDA(cudaFree(ptr));
    return 0;
}

int kib_host_alloc(void **ptr, size_t bytes)
{
    KIB_REQUIRE(ptr != nullptr, "kib_host_alloc: null argument");
    if (bytes == 0) bytes = 1;
    KIB_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
    return 0;
}

int kib_host_free(void *ptr)
{
    if (ptr != nullptr) KIB_CUDA(cudaFreeHost(ptr));
    return 0;
}

int kib_host_register(void *ptr, size_t bytes)
{
    KIB_REQUIRE(ptr != nullptr && bytes > 0, "kib_host_register: empty range");
    KIB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return 0;
}

int kib_host_unregister(void *ptr)
{
    KIB_CUDA(cudaHostUnregister(ptr));
    return 0;
}

int kib_memset_async(void *ptr, int value, size_t bytes, kib_stream_t stream)
{
    if (bytes == 0) return 0;
    KIB_CUDA(cudaMemsetAsync(ptr, value, bytes, as_stream(stream)));
    return 0;
}

int kib_memcpy_h2d_async(void *dst, const void *src, size_t bytes, kib_stream_t stream)
{
    if (bytes == 0) return 0;
    KIB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, as_stream(stream)));
    return 0;
}

int kib_memcpy_d2h_async(void *dst, const void *src, size_t bytes, kib_stream_t stream)
{
    if (bytes == 0) return 0;
    const int rc = small_d2h(dst, bytes, bytes, src, bytes, bytes, bytes, 1, 1, as_stream(stream));
    if (rc != 1) return rc;
    KIB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, as_stream(stream)));
    return 0;
}

int kib_memcpy_d2d_async(void *dst, const void *src, size_t bytes, kib_stream_t stream)
{
    if (bytes == 0) return 0;
    KIB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
    return 0;
}

int kib_memcpy3d_async(void *dst, size_t dst_row_pitch, size_t dst_plane_pitch,
                       const void *src, size_t src_row_pitch, size_t src_plane_pitch,
                       size_t width_bytes, size_t height, size_t depth,
                       int kind, kib_stream_t stream)
{
    if (width_bytes == 0 || height == 0 || depth == 0) return 0;
    if (kind == 1) {
        const int rc = small_d2h(dst, dst_row_pitch, dst_plane_pitch, src, src_row_pitch,
                                 src_plane_pitch, width_bytes, height, depth, as_stream(stream));
        if (rc != 1) return rc;
    }
    cudaMemcpyKind k;
    switch (kind) {
    case 0: k = cudaMemcpyHostToDevice; break;
    case 1: k = cudaMemcpyDeviceToHost; break;
    case 2: k = cudaMemcpyDeviceToDevice; break;
    default:
        set_error("kib_memcpy3d_async: bad kind %d", kind);
        return -1;
    }
    // One 2-D copy per plane: planes of a padded array need not be a whole
    // number of rows apart, which a single pitched 3-D copy cannot express.
    for (size_t z = 0; z < depth; z++) {
        KIB_CUDA(cudaMemcpy2DAsync(
            static_cast<char *>(dst) + z * dst_plane_pitch, dst_row_pitch,
            static_cast<const char *>(src) + z * src_plane_pitch, src_row_pitch,
            width_bytes, height, k, as_stream(stream)));
    }
    return 0;
}

int kib_fft_plan2d_create(kib_fft_plan_t *plan, int ny, int nx, int row_stride, int dtype)
{
    KIB_REQUIRE(plan != nullptr, "kib_fft_plan2d_create: null argument");
    KIB_REQUIRE(ny > 0 && nx > 0 && row_stride >= nx, "kib_fft_plan2d_create: bad shape");
    KIB_REQUIRE(dtype == KIB_F32 || dtype == KIB_F64, "kib_fft_plan2d_create: bad dtype");
    FftPlan *p = new FftPlan;
    p->dtype = dtype;
    cufftResult r = cufftCreate(&p->handle);
    if (r != CUFFT_SUCCESS) {
        delete p;
        set_error("cufftCreate failed: %s", cufft_error_string(r));
        return 1000 + (int) r;
    }
    long long n[2] = {ny, nx};
    long long embed[2] = {ny, row_stride};
    size_t work_size = 0;
    r = cufftMakePlanMany64(p->handle, 2, n, embed, 1, (long long) ny * row_stride,
                            embed, 1, (long long) ny * row_stride,
                            dtype == KIB_F32 ? CUFFT_C2C : CUFFT_Z2Z, 1, &work_size);
    if (r != CUFFT_SUCCESS) {
        cufftDestroy(p->handle);
        delete p;
        set_error("cufftMakePlanMany64(%d x %d) failed: %s", ny, nx, cufft_error_string(r));
        return 1000 + (int) r;
    }
    *plan = reinterpret_cast<kib_fft_plan_t>(p);
    return 0;
}

int kib_fft_plan2d_real_create(kib_fft_plan_t *plan, int ny, int nx, int real_row_stride,
                               int complex_row_stride, int inverse, int dtype)
{
    KIB_REQUIRE(plan != nullptr, "kib_fft_plan2d_real_create: null argument");
    KIB_REQUIRE(ny > 0 && nx > 0 && real_row_stride >= nx && complex_row_stride >= nx / 2 + 1,
                "kib_fft_plan2d_real_create: bad shape");
    KIB_REQUIRE(dtype == KIB_F32 || dtype == KIB_F64, "kib_fft_plan2d_real_create: bad dtype");
    FftPlan *p = new FftPlan;
    p->dtype = dtype;
    p->kind = inverse ? 2 : 1;
    cufftResult r = cufftCreate(&p->handle);
    if (r != CUFFT_SUCCESS) {
        delete p;
        set_error("cufftCreate failed: %s", cufft_error_string(r));
        return 1000 + (int) r;
    }
    long long n[2] = {ny, nx};
    long long real_embed[2] = {ny, real_row_stride};
    long long complex_embed[2] = {ny, complex_row_stride};
    size_t work_size = 0;
    cufftType type = dtype == KIB_F32 ? (inverse ? CUFFT_C2R : CUFFT_R2C)
                                      : (inverse ? CUFFT_Z2D : CUFFT_D2Z);
    if (inverse)
        r = cufftMakePlanMany64(p->handle, 2, n, complex_embed, 1,
                                (long long) ny * complex_row_stride, real_embed, 1,
                                (long long) ny * real_row_stride, type, 1, &work_size);
    else
        r = cufftMakePlanMany64(p->handle, 2, n, real_embed, 1,
                                (long long) ny * real_row_stride, complex_embed, 1,
                                (long long) ny * complex_row_stride, type, 1, &work_size);
    if (r != CUFFT_SUCCESS) {
        cufftDestroy(p->handle);
        delete p;
        set_error("cufftMakePlanMany64(real %d x %d) failed: %s", ny, nx, cufft_error_string(r));
        return 1000 + (int) r;
    }
    *plan = reinterpret_cast<kib_fft_plan_t>(p);
    return 0;
}

int kib_fft_plan1d_create(kib_fft_plan_t *plan, int n, int64_t stride, int64_t dist,
                          int batch, int dtype)
{
    KIB_REQUIRE(plan != nullptr, "kib_fft_plan1d_create: null argument");
    KIB_REQUIRE(n > 0 && batch > 0 && stride > 0 && dist > 0, "kib_fft_plan1d_create: bad shape");
    KIB_REQUIRE(dtype == KIB_F32 || dtype == KIB_F64, "kib_fft_plan1d_create: bad dtype");
    FftPlan *p = new FftPlan;
    p->dtype = dtype;
    cufftResult r = cufftCreate(&p->handle);
    if (r != CUFFT_SUCCESS) {
        delete p;
        set_error("cufftCreate failed: %s", cufft_error_string(r));
        return 1000 + (int) r;
    }
    long long len[1] = {n};
    long long embed[1] = {n};
    size_t work_size = 0;
    r = cufftMakePlanMany64(p->handle, 1, len, embed, stride, dist, embed, stride, dist,
                            dtype == KIB_F32 ? CUFFT_C2C : CUFFT_Z2Z, batch, &work_size);
    if (r != CUFFT_SUCCESS) {
        cufftDestroy(p->handle);
        delete p;
        set_error("cufftMakePlanMany64(1-D %d x %d) failed: %s", n, batch, cufft_error_string(r));
        return 1000 + (int) r;
    }
    *plan = reinterpret_cast<kib_fft_plan_t>(p);
    return 0;
}

int kib_fft_plan2d_exec(kib_fft_plan_t plan, void *src, void *dst, int direction,
                        kib_stream_t stream)
{
    KIB_REQUIRE(plan != nullptr, "kib_fft_plan2d_exec: null plan");
    FftPlan *p = reinterpret_cast<FftPlan *>(plan);
    int dir = direction == KIB_FFT_FORWARD ? CUFFT_FORWARD : CUFFT_INVERSE;
    KIB_CUFFT(cufftSetStream(p->handle, as_stream(stream)));
    if (p->kind == 1) {
        if (p->dtype == KIB_F32)
            KIB_CUFFT(cufftExecR2C(p->handle, static_cast<cufftReal *>(src),
                                   static_cast<cufftComplex *>(dst)));
        else
            KIB_CUFFT(cufftExecD2Z(p->handle, static_cast<cufftDoubleReal *>(src),
                                   static_cast<cufftDoubleComplex *>(dst)));
        return 0;
    }
    if (p->kind == 2) {
        if (p->dtype == KIB_F32)
            KIB_CUFFT(cufftExecC2R(p->handle, static_cast<cufftComplex *>(src),
                                   static_cast<cufftReal *>(dst)));
        else
            KIB_CUFFT(cufftExecZ2D(p->handle, static_cast<cufftDoubleComplex *>(src),
                                   static_cast<cufftDoubleReal *>(dst)));
        return 0;
    }
    if (p->dtype == KIB_F32)
        KIB_CUFFT(cufftExecC2C(p->handle, static_cast<cufftComplex *>(src),
                               static_cast<cufftComplex *>(dst), dir));
    else
        KIB_CUFFT(cufftExecZ2Z(p->handle, static_cast<cufftDoubleComplex *>(src),
                               static_cast<cufftDoubleComplex *>(dst), dir));
    return 0;
}

int kib_fft_plan2d_destroy(kib_fft_plan_t plan)
{
    if (plan == nullptr) return 0;
    FftPlan *p = reinterpret_cast<FftPlan *>(plan);
    KIB_CUFFT(cufftDestroy(p->handle));
    delete p;
    return 0;
}

}  // extern "C"
