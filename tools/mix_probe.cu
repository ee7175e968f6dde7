// mix_probe.cu — microbenchmark (not product code): the memory traffic of k_query_kmers_rows<1,4> on the C3 workload, replayed
// without the walk's arithmetic, to find what the memory system of a B200 sustains for that MIX of accesses — the joint ceiling
// the kernel's time is compared with in DESIGN.md §5 (the additive model "random accesses / probe rate + streamed bytes / copy
// bandwidth" is only an estimate: random row activations and streamed bursts share the HBM channels).
//
// Per item (one per thread, grid-stride, 148 x 8 CTAs of 256 like the product kernel):
//   stream   8 B in  (ld.global.cs)                                  | the packed k-mer
//   L2       8 B from a 2 MB table, 32 B sector from an F MB table   | root directory entry, stored-k-mer filter block
//   HBM      with probability P: one 256-bit load, L2::64B, from a T MB table   | the bucket
//   L2       with probability Q: 16 B from a 20 MB table             | the class row
//   stream   1 B + 16 B out (st.global.cs)                           | presence byte + colour row
// Table indices come from a hash of the streamed word, so every load depends on the input like in the walk; the bucket index
// additionally depends on the two L2 loads (the walk cannot issue the bucket load before the root entry arrived).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/mix_probe tools/mix_probe.cu
//   tools/mix_probe [n_items=125000000] [bucket_table_MB=771] [filter_MB=25] [P=0.54] [Q=0.51]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
    x = (x + 0x9E3779B97F4A7C15ULL) * 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 31; x *= 0x94D049BB133111EBULL; x ^= x >> 29;
    return x;
}

struct Tables {
    const uint64_t* in;      // n items
    const uint2* rootdir;    // 2 MB
    const uint64_t* filter;  // F MB, 32-byte blocks
    const uint64_t* buckets; // T MB, 32-byte buckets
    const uint4* rows;       // 20 MB
    uint8_t* present;
    uint4* out_rows;
    uint32_t n_filter, n_rows;
    uint64_t n_buckets;
    uint32_t p_thresh, q_thresh; // probabilities scaled to 2^32
};

// FLAGS: 1 = stream in/out, 2 = root directory entry (L2), 16 = filter block (L2), 4 = random HBM bucket, 8 = class row (L2),
//        32 = "filter first": only the items that go on to a bucket fetch their root directory entry
template <int FLAGS>
__global__ void __launch_bounds__(256) k_mix(const Tables t, size_t n, unsigned long long* sink) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        uint64_t x = (FLAGS & 1) ? __ldcs((const unsigned long long*)t.in + i) : (uint64_t)i;
        const uint64_t h = mix(x);
        uint64_t dep = 0;
        if ((FLAGS & 2) && !(FLAGS & 32)) {
            const uint2 e = __ldg(t.rootdir + (h & 0x3ffffu));
            dep ^= (e.x ^ e.y) & 1ULL; // tables are zero-filled: dep == 0, but the compiler cannot know
        }
        if (FLAGS & 16) {
            uint64_t a, b, c, d;
            const uint64_t* p = t.filter + (size_t)((uint32_t)(((h >> 32) * (uint64_t)t.n_filter) >> 32)) * 4;
            asm volatile("ld.global.nc.L2::evict_last.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
            dep ^= (a ^ b ^ c ^ d) & 1ULL;
        }
        uint64_t h2 = mix(h + dep);
        const bool to_bucket = (uint32_t)h2 < t.p_thresh;
        if ((FLAGS & 2) && (FLAGS & 32) && to_bucket) {
            const uint2 e = __ldg(t.rootdir + (h & 0x3ffffu));
            h2 += (e.x ^ e.y) & 1ULL;
        }
        uint64_t got = 0;
        if ((FLAGS & 4) && to_bucket) {
            const uint64_t* p = t.buckets + ((h2 >> 20) % t.n_buckets) * 4;
            uint64_t a, b, c, d;
            asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];"
                         : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
            got = a ^ b ^ c ^ d;
        }
        uint4 r = make_uint4(0, 0, 0, 0);
        const bool hit = (uint32_t)(h2 >> 32) < t.q_thresh;
        if ((FLAGS & 8) && hit) r = __ldg(t.rows + (uint32_t)((((h2 + got) & 0xffffffffu) * (uint64_t)t.n_rows) >> 32));
        r.y ^= (uint32_t)got; // zero tables: the stored value does not change, the loads stay
        if (FLAGS & 1) {
            t.present[i] = hit;
            __stcs(t.out_rows + i, r);
        } else {
            acc += r.x ^ r.y;
        }
    }
    if (acc == 0x1234567ULL) *sink = acc;
}

template <int FLAGS>
static float run(const char* name, const Tables& t, size_t n, unsigned long long* sink, double hbm_bytes_per_item) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        k_mix<FLAGS><<<148 * 8, 256>>>(t, n, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    printf("%-58s %7.3f ms  %7.2f G items/s  %7.1f GB/s HBM (algorithmic)  %s\n", name, best, n / best / 1e6, hbm_bytes_per_item * n / best / 1e6,
           cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
    return best;
}

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? (size_t)atof(argv[1]) : 125000000;
    const double t_mb = argc > 2 ? atof(argv[2]) : 771, f_mb = argc > 3 ? atof(argv[3]) : 25;
    const double P = argc > 4 ? atof(argv[4]) : 0.54, Q = argc > 5 ? atof(argv[5]) : 0.51;
    Tables t;
    uint64_t* in; uint8_t* present; uint4* out_rows; void *rootdir, *filter, *buckets, *rows; unsigned long long* sink;
    const size_t fb = (size_t)((f_mb > 0.001 ? f_mb : 0.001) * 1e6) / 32 * 32, tb = (size_t)(t_mb * 1e6) / 32 * 32, rb = 20000000 / 16 * 16;
    cudaMalloc(&in, n * 8); cudaMalloc(&present, n); cudaMalloc(&out_rows, n * 16);
    cudaMalloc(&rootdir, 2 << 20); cudaMalloc(&filter, fb); cudaMalloc(&buckets, tb); cudaMalloc(&rows, rb); cudaMalloc(&sink, 8);
    if (cudaGetLastError() != cudaSuccess) { printf("allocation failed\n"); return 1; }
    cudaMemset(rootdir, 0, 2 << 20); cudaMemset(filter, 0, fb); cudaMemset(buckets, 0, tb); cudaMemset(rows, 0, rb);
    // input: distinct words (their hash picks every table index)
    {
        uint64_t* h = (uint64_t*)malloc(n * 8);
        uint64_t s = 88172645463325252ULL;
        for (size_t i = 0; i < n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = s; }
        cudaMemcpy(in, h, n * 8, cudaMemcpyHostToDevice);
        free(h);
    }
    t.in = in; t.rootdir = (const uint2*)rootdir; t.filter = (const uint64_t*)filter; t.buckets = (const uint64_t*)buckets; t.rows = (const uint4*)rows;
    t.present = present; t.out_rows = out_rows;
    t.n_filter = (uint32_t)(fb / 32 ? fb / 32 : 1); t.n_rows = (uint32_t)(rb / 16); t.n_buckets = tb / 32;
    t.p_thresh = (uint32_t)(P * 4294967295.0); t.q_thresh = (uint32_t)(Q * 4294967295.0);
    printf("items %zu, bucket table %.0f MB, filter %.0f MB, P(bucket) %.3f, P(row) %.3f\n", n, t_mb, f_mb, P, Q);
    const float stream = run<1>("stream only (8 B in, 17 B out)", t, n, sink, 25);
    const float rnd = run<4>("random 64-byte HBM accesses only (P per item)", t, n, sink, 32 * P);
    run<1 | 4>("stream + random HBM access", t, n, sink, 25 + 32 * P);
    run<2>("root directory entry only (1 random L2 sector per item)", t, n, sink, 0);
    run<16>("filter block only (1 random L2 sector per item)", t, n, sink, 0);
    run<2 | 16>("root directory + filter (2 random L2 sectors per item)", t, n, sink, 0);
    run<1 | 2 | 16>("stream + root directory + filter", t, n, sink, 25);
    run<2 | 16 | 4>("root directory + filter + dependent random HBM access", t, n, sink, 32 * P);
    const float all = run<1 | 2 | 16 | 4 | 8>("ALL: the traffic of k_query_kmers_rows<1,4>", t, n, sink, 25 + 32 * P);
    run<1 | 2 | 16 | 4 | 8 | 32>("variant: filter first, root entry only for bucket-bound items", t, n, sink, 25 + 32 * P);
    run<1 | 16 | 4 | 8>("variant: no root directory (flat hash of the k-mer)", t, n, sink, 25 + 32 * P);
    run<1 | 16 | 4>("variant: no root directory, no class row", t, n, sink, 25 + 32 * P);
    run<1 | 4 | 8>("variant: no root directory, no filter (P applies)", t, n, sink, 25 + 32 * P);
    printf("additive (stream + random) %.3f ms, max %.3f ms, measured mix %.3f ms; random accesses in the mix: %.2f G/s\n", stream + rnd,
           stream > rnd ? stream : rnd, all, P * n / all / 1e6);
    return 0;
}
