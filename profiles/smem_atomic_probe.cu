// Upper bound of a shared-memory-atomic sub-grid gridder (the design BASELINE.json's north_star
// sketches and round 1 turned down on a guide's number): how many visibilities per second can a
// B200 accumulate into a shared-memory tile with red.shared.add.f32 when everything else is
// free?  Best case on purpose: the visibilities of a block already belong to its tile (no
// binning pass), one thread per (footprint cell, polarization) so that the 2 atomics of a
// thread never collide inside a warp, taps from a shared-memory table, no global traffic in the
// loop, tile written out once at the end.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/_build/smem_atomic_probe \
//        profiles/smem_atomic_probe.cu && profiles/_build/smem_atomic_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

constexpr int K = 7, P = 4, OVERSAMPLE = 8;
constexpr int THREADS = 256;                 // 196 = 49 cells x 4 pols do the work

template <int TILE>
__global__ void __launch_bounds__(THREADS)
probe(float2 *out, int vis_per_block, int use_atomics)
{
    extern __shared__ float2 tile[];          // [P][TILE][TILE]
    __shared__ float2 taps[OVERSAMPLE][K];
    for (int i = threadIdx.x; i < P * TILE * TILE; i += THREADS) tile[i] = make_float2(0.f, 0.f);
    for (int i = threadIdx.x; i < OVERSAMPLE * K; i += THREADS)
        taps[i / K][i % K] = make_float2(1.0f / (1 + i), 0.5f / (2 + i));
    __syncthreads();
    const int t = threadIdx.x;
    const bool active = t < K * K * P;
    const int p = t / (K * K), cell = t % (K * K), j = cell / K, k = cell % K;
    unsigned state = 12345u + blockIdx.x * 977u;
    int u0 = 20, v0 = 20;
    float2 acc = make_float2(0.f, 0.f);
    for (int v = 0; v < vis_per_block; v++) {
        // the same pseudo-random track in every thread: footprint origin walks slowly, sub-pixel
        // position changes every visibility (as along a baseline track)
        state = state * 1664525u + 1013904223u;
        const int su = (state >> 8) & 7, sv = (state >> 12) & 7;
        if ((state >> 20 & 15) == 0) u0 += (state >> 24 & 1) ? 1 : -1;
        if ((state >> 16 & 15) == 0) v0 += (state >> 25 & 1) ? 1 : -1;
        u0 = min(max(u0, 0), TILE - K);
        v0 = min(max(v0, 0), TILE - K);
        if (active) {
            const float2 wu = taps[su][k], wv = taps[sv][j];
            const float wr = wv.x * wu.x - wv.y * wu.y, wi = wv.x * wu.y + wv.y * wu.x;
            const float sx = 1.0f + 0.001f * v, sy = 0.5f - 0.001f * p;      // the sample
            const float re = sx * wr + sy * wi, im = sy * wr - sx * wi;
            float2 *dst = tile + (p * TILE + v0 + j) * TILE + u0 + k;
            if (use_atomics) {
                atomicAdd(&dst->x, re);
                atomicAdd(&dst->y, im);
            } else {
                acc.x += re;                    // arithmetic only: the floor of the loop
                acc.y += im;
            }
        }
    }
    __syncthreads();
    if (!use_atomics && active) tile[t] = acc;
    __syncthreads();
    for (int i = threadIdx.x; i < P * TILE * TILE; i += THREADS)
        out[(size_t) blockIdx.x * P * TILE * TILE + i] = tile[i];
}

template <int TILE>
static int run(int sms, int blocks_per_sm)
{
    const int smem = P * TILE * TILE * (int) sizeof(float2);
    cudaFuncSetAttribute(probe<TILE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int blocks = sms * blocks_per_sm, vis_per_block = 200000 / blocks_per_sm;
    float2 *out;
    cudaMalloc(&out, (size_t) blocks * smem);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    int resident = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, probe<TILE>, THREADS, smem);
    printf("{\"tile\": %d, \"smem_kb\": %d, \"blocks\": %d, \"resident_blocks_per_sm\": %d, "
           "\"vis_per_block\": %d", TILE, smem / 1024, blocks, resident, vis_per_block);
    for (int use_atomics = 1; use_atomics >= 0; use_atomics--) {
        probe<TILE><<<blocks, THREADS, smem>>>(out, 1000, use_atomics);
        cudaDeviceSynchronize();
        cudaEventRecord(a);
        probe<TILE><<<blocks, THREADS, smem>>>(out, vis_per_block, use_atomics);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        const double vis = (double) blocks * vis_per_block;
        printf(", \"%s\": {\"ms\": %.3f, \"gvis_per_s\": %.3f, \"atomics_per_clk_per_sm\": %.2f}",
               use_atomics ? "smem_atomics" : "arithmetic_only", ms, vis / ms / 1e6,
               use_atomics ? vis * K * K * P * 2 / (ms * 1e-3) / 1.965e9 / sms : 0.0);
    }
    cudaError_t err = cudaGetLastError();
    printf(", \"cuda\": \"%s\"}\n", cudaGetErrorString(err));
    cudaFree(out);
    return err != cudaSuccess;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int rc = run<64>(sms, 1);         // 128 KB tile, one block per SM
    rc |= run<32>(sms, 6);            // 32 KB tiles, six blocks per SM
    rc |= run<16>(sms, 8);            // 8 KB tiles, eight blocks (2048 threads) per SM
    return rc;
}
