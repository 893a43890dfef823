// Micro-benchmark behind DESIGN.md §5: how fast can a B200 do 4-byte random probes into a table, as a
// function of table size (L2-resident vs HBM), load flavour, loads in flight per thread, and the L2
// fetch granularity limit?  Also: fire-and-forget RED.ADD, load+CAS, and 32-stream scattered stores
// (the write pattern of a radix partition pass).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o probe_bench tools/probe_bench.cu && ./probe_bench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t ld_cg(const uint32_t* p) { uint32_t v; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ld_nc(const uint32_t* p) { uint32_t v; asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ld_ca(const uint32_t* p) { uint32_t v; asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ld_ef(const uint32_t* p) { uint32_t v; asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
// L2 prefetch-size hints and eviction policies: does anything make an HBM miss fetch less than 128 bytes?
__device__ __forceinline__ uint32_t ld_cg64(const uint32_t* p) { uint32_t v; asm volatile("ld.global.cg.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ld_64(const uint32_t* p) { uint32_t v; asm volatile("ld.global.L2::64B.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ld_cg128(const uint32_t* p) { uint32_t v; asm volatile("ld.global.cg.L2::128B.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
__device__ __forceinline__ uint32_t ld_efp(const uint32_t* p) {
    uint32_t v; uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_u8(const uint32_t* p) { uint32_t v; asm volatile("ld.global.cg.u8 %0, [%1];" : "=r"(v) : "l"(p)); return v; }

// mode 0 cg, 1 nc, 2 ca, 3 volatile, 4 red.add, 5 load+cas(2-bit sat), 6 scattered store to 32 streams,
// 7 cg.L2::64B, 8 L2::64B, 9 cg.L2::128B, 10 nc + evict_first policy, 11 cg.u8
template <int MODE, int U>
__global__ void __launch_bounds__(256) probe(uint32_t* table, uint64_t words_mask, uint64_t iters, uint32_t* sink, uint32_t* streams,
                                             uint64_t stream_cap) {
    uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    uint32_t cnt = 0;
    for (uint64_t it = 0; it < iters; ++it) {
        uint32_t h[U], v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) h[u] = mix((uint32_t)(tid * 2654435761u) + (uint32_t)(it * U + u) * 40503u + 0x9e3779b9u * (uint32_t)(tid >> 20));
        if (MODE <= 3 || MODE == 5 || MODE >= 7) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t* p = table + (((uint64_t)h[u]) & words_mask);
                v[u] = MODE == 0 || MODE == 5 ? ld_cg(p) : MODE == 1 ? ld_nc(p) : MODE == 2 ? ld_ca(p) : MODE == 3 ? ld_ef(p) :
                       MODE == 7 ? ld_cg64(p) : MODE == 8 ? ld_64(p) : MODE == 9 ? ld_cg128(p) : MODE == 10 ? ld_efp(p) : ld_u8(p);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += v[u];
            if (MODE == 5) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    uint32_t* p = table + (((uint64_t)h[u]) & words_mask);
                    int sh = (mix(h[u]) & 15u) * 2;
                    uint32_t seen = v[u];
                    while (((seen >> sh) & 3u) < 3u) {
                        uint32_t old = atomicCAS(p, seen, seen + (1u << sh));
                        if (old == seen) break;
                        seen = old;
                    }
                }
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int u = 0; u < U; ++u) atomicAdd(table + (((uint64_t)h[u]) & words_mask), 1u);   // result unused -> RED
        } else {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                uint32_t p = h[u] >> 27;
                // per-thread private slot sequence inside stream p (no counters: measures the store path only)
                uint64_t slot = ((uint64_t)tid * 64 + (cnt & 63)) % stream_cap;
                streams[(uint64_t)p * stream_cap + slot] = h[u];
                ++cnt;
            }
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int MODE, int U>
static void run(const char* name, uint32_t* table, uint64_t bytes, int blocks_per_sm, uint32_t* sink, uint32_t* streams, uint64_t stream_cap) {
    uint64_t words_mask = bytes / 4 - 1;
    int grid = 148 * blocks_per_sm;
    uint64_t threads = (uint64_t)grid * 256;
    uint64_t target = 1ull << 31;                                  // ~2 G probes
    uint64_t iters = target / (threads * U);
    if (iters < 1) iters = 1;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    probe<MODE, U><<<grid, 256>>>(table, words_mask, iters / 8 + 1, sink, streams, stream_cap);   // warm
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    probe<MODE, U><<<grid, 256>>>(table, words_mask, iters, sink, streams, stream_cap);
    CK(cudaEventRecord(b));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    double probes = (double)threads * U * iters;
    printf("%-10s table %7.0f MiB  U=%2d  blk/SM=%d  %8.2f ms  %7.1f Gprobe/s  (x32B = %6.0f GB/s)\n", name, bytes / 1048576.0, U, blocks_per_sm,
           ms, probes / ms / 1e6, probes * 32 / ms / 1e6);
    fflush(stdout);
}

int main(int argc, char** argv) {
    int gran = argc > 1 ? atoi(argv[1]) : 0;
    if (gran) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity <- %d: %s, now %zu\n", gran, cudaGetErrorString(e), got);
    } else {
        size_t got = 0;
        cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity default %zu\n", got);
    }
    uint32_t *table, *sink, *streams;
    uint64_t max_bytes = 4ull << 30;
    CK(cudaMalloc(&table, max_bytes));
    CK(cudaMemset(table, 0, max_bytes));
    CK(cudaMalloc(&sink, 4));
    uint64_t stream_cap = 1ull << 26;                              // 32 streams x 256 MiB
    CK(cudaMalloc(&streams, 32 * stream_cap * 4));
    if (argc > 2) {                                                 // ./probe_bench <gran> hints : the fill-size experiments only
        for (uint64_t s : {1ull << 30}) {
            run<0, 12>("ld.cg", table, s, 4, sink, streams, stream_cap);
            run<7, 12>("cg.L2::64B", table, s, 4, sink, streams, stream_cap);
            run<8, 12>("L2::64B", table, s, 4, sink, streams, stream_cap);
            run<9, 12>("cg.L2::128", table, s, 4, sink, streams, stream_cap);
            run<10, 12>("nc+evict1", table, s, 4, sink, streams, stream_cap);
            run<11, 12>("cg.u8", table, s, 4, sink, streams, stream_cap);
        }
        return 0;
    }
    uint64_t sizes[] = {16ull << 20, 32ull << 20, 48ull << 20, 64ull << 20, 96ull << 20, 128ull << 20, 1ull << 30, 4ull << 30};
    for (uint64_t s : sizes) run<0, 12>("ld.cg", table, s, 4, sink, streams, stream_cap);
    for (uint64_t s : {32ull << 20, 1ull << 30}) {
        run<0, 4>("ld.cg", table, s, 4, sink, streams, stream_cap);
        run<0, 24>("ld.cg", table, s, 4, sink, streams, stream_cap);
        run<0, 12>("ld.cg", table, s, 8, sink, streams, stream_cap);
        run<0, 12>("ld.cg", table, s, 2, sink, streams, stream_cap);
        run<1, 12>("ld.nc", table, s, 4, sink, streams, stream_cap);
        run<2, 12>("ld.ca", table, s, 4, sink, streams, stream_cap);
        run<3, 12>("ld.volat", table, s, 4, sink, streams, stream_cap);
        run<4, 12>("red.add", table, s, 4, sink, streams, stream_cap);
        CK(cudaMemset(table, 0, max_bytes));
        run<5, 12>("ld+cas", table, s, 4, sink, streams, stream_cap);
        CK(cudaMemset(table, 0, max_bytes));
    }
    run<4, 12>("red.add", table, 4ull << 30, 4, sink, streams, stream_cap);
    run<6, 12>("st.32strm", table, 1ull << 30, 4, sink, streams, stream_cap);
    run<6, 3>("st.32strm", table, 1ull << 30, 8, sink, streams, stream_cap);
    return 0;
}
