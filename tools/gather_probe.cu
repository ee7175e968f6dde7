// gather_probe.cu — microbenchmark (not product code): rate and DRAM traffic of independent random loads from a
// table far larger than L2, for several load flavours and access widths. Used to size the random-access roofline
// the BFT walk is bound by and to pick the cache operators of bft_ld_bucket().
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t i) {
    uint64_t x = (i + 0x9E3779B97F4A7C15ULL) * 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 31; x *= 0x94D049BB133111EBULL; x ^= x >> 29;
    return x;
}

template <int MODE>
__device__ __forceinline__ uint64_t load8(const uint64_t* p) {
    uint64_t v;
    if (MODE == 0) v = __ldg((const unsigned long long*)p);
    else if (MODE == 1) v = __ldcs((const unsigned long long*)p);
    else if (MODE == 2) asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else if (MODE == 4) asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}

// one 256-bit load (sm_100): MODE 0 plain, 1 L2::evict_first, 2 L2::evict_last
template <int MODE>
__device__ __forceinline__ uint64_t load32(const uint64_t* p) {
    uint64_t a, b, c, d;
    if (MODE == 0) asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    else if (MODE == 1) asm volatile("ld.global.nc.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    else asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    return a ^ b ^ c ^ d;
}

template <int MODE>
__global__ void __launch_bounds__(256) k_gather256(const uint64_t* __restrict__ table, size_t n_units, size_t n_loads, unsigned long long* sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_loads; i += stride) acc += load32<MODE>(table + (mix(i) % n_units) * 4);
    if (acc == 0x1234567ULL) *sink = acc;
}

template <int MODE>
static void run256(const char* name, const uint64_t* table, size_t bytes, size_t n_loads, unsigned long long* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k_gather256<MODE><<<148 * 8, 256>>>(table, bytes / 32, n_loads, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    printf("%-44s width  32 B: %7.2f G loads/s  (%6.1f GB/s useful)  %s\n", name, n_loads / best / 1e6, n_loads / best / 1e6 * 32,
           cudaGetErrorString(cudaGetLastError()));
}

// WIDTH = bytes read per random access (8, 32 = one aligned sector as 2x16B, 64 = aligned pair as 4x16B)
template <int MODE, int WIDTH>
__global__ void __launch_bounds__(256) k_gather(const uint64_t* __restrict__ table, size_t n_units, size_t n_loads, unsigned long long* sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_loads; i += stride) {
        const uint64_t* p = table + (mix(i) % n_units) * (WIDTH / 8);
        if (WIDTH == 8) acc += load8<MODE>(p);
        else {
#pragma unroll
            for (int j = 0; j < WIDTH / 16; j++) {
                ulonglong2 t;
                if (MODE == 1) t = __ldcs((const ulonglong2*)p + j);
                else t = __ldg((const ulonglong2*)p + j);
                acc += t.x ^ t.y;
            }
        }
    }
    if (acc == 0x1234567ULL) *sink = acc;
}

template <int MODE, int WIDTH>
static void run(const char* name, const uint64_t* table, size_t bytes, size_t n_loads, unsigned long long* sink) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k_gather<MODE, WIDTH><<<148 * 8, 256>>>(table, bytes / WIDTH, n_loads, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    printf("%-44s width %3d B: %7.2f G loads/s  (%6.1f GB/s useful)  %s\n", name, WIDTH, n_loads / best / 1e6, n_loads / best / 1e6 * WIDTH,
           cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
    size_t bytes = (size_t)((argc > 1 ? atof(argv[1]) : 4.0) * (double)(1ull << 30)) / 4096 * 4096;
    size_t n_loads = (size_t)1 << 28;
    uint64_t* table; unsigned long long* sink;
    cudaMalloc(&table, bytes); cudaMalloc(&sink, 8);
    cudaMemset(table, 1, bytes);
    int g = argc > 2 ? atoi(argv[2]) : 0;
    if (g) printf("cudaLimitMaxL2FetchGranularity=%d -> %s\n", g, cudaGetErrorString(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g)));
    size_t cur = 0; cudaDeviceGetLimit(&cur, cudaLimitMaxL2FetchGranularity);
    printf("table %.2f GiB, %zu loads, L2 fetch granularity limit %zu\n", bytes / 1073741824.0, n_loads, cur);
    run<0, 8>("ld.global.nc (__ldg)", table, bytes, n_loads, sink);
    run<1, 8>("ld.global.cs (__ldcs)", table, bytes, n_loads, sink);
    run<2, 8>("ld.global.nc.L2::64B", table, bytes, n_loads, sink);
    run<3, 8>("ld.global.nc.L1::no_allocate.L2::128B", table, bytes, n_loads, sink);
    run<4, 8>("ld.global.cv", table, bytes, n_loads, sink);
    run<5, 8>("ld.global.nc.L1::no_allocate", table, bytes, n_loads, sink);
    run<0, 32>("__ldg 2x16B (one sector)", table, bytes, n_loads, sink);
    run<1, 32>("__ldcs 2x16B (one sector)", table, bytes, n_loads, sink);
    run<0, 64>("__ldg 4x16B (sector pair)", table, bytes, n_loads, sink);
    run<1, 64>("__ldcs 4x16B (sector pair)", table, bytes, n_loads, sink);
    run256<0>("ld.global.nc.v4.u64 (256-bit)", table, bytes, n_loads, sink);
    run256<1>("ld.global.nc.L2::evict_first.v4.u64", table, bytes, n_loads, sink);
    run256<2>("ld...L1::no_allocate.L2::evict_first.v4.u64", table, bytes, n_loads, sink);
    return 0;
}
