// Internal helpers shared by the translation units of libkatimager_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdarg>
#include "katimager_b200.h"

namespace kib {

void set_error(const char *fmt, ...);

inline cudaStream_t as_stream(kib_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define KIB_CUDA(expr)                                                              \
    do {                                                                            \
        cudaError_t kib_err__ = (expr);                                             \
        if (kib_err__ != cudaSuccess) {                                             \
            ::kib::set_error("%s failed: %s (%s:%d)", #expr,                        \
                             cudaGetErrorString(kib_err__), __FILE__, __LINE__);    \
            (void) cudaGetLastError();   /* reported here: do not leave it for a later launch check */ \
            return (int) kib_err__;                                                 \
        }                                                                           \
    } while (0)

#define KIB_REQUIRE(cond, ...)                                                      \
    do {                                                                            \
        if (!(cond)) {                                                              \
            ::kib::set_error(__VA_ARGS__);                                          \
            return -1;                                                              \
        }                                                                           \
    } while (0)

#define KIB_CHECK_LAUNCH() KIB_CUDA(cudaGetLastError())

inline int divup(int64_t a, int64_t b) { return (int) ((a + b - 1) / b); }

int sm_count();

// kib_grid.cu: frees the gridder's staging scratch of a stream (called when it is destroyed)
void release_grid_scratch(cudaStream_t stream);

template <typename Real> struct Complex2;
template <> struct Complex2<float> { typedef float2 type; };
template <> struct Complex2<double> { typedef double2 type; };

}  // namespace kib
