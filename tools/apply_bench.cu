// Micro-benchmark of S1 phase B (s1_apply_kernel) on a synthetic stream with the screen's own mix of probes:
// 70 % of the hashes come from a pool of recurring ("true") k-mers that saturate after three sightings, 30 % are
// one-off ("sequencing error") k-mers that need a 0 -> 1 compare-and-swap.  All hashes fall in one 64 MiB table
// slice.  Variants: probes per thread per tile, ring depth, resident CTAs, and the load-only bound.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/apply_bench tools/apply_bench.cu
#include "../localhgt_b200/csrc/lhgt_kernels.cu"
#include <cstdio>
#include <cstdlib>
using namespace lhgt;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

__global__ void make_stream(uint32_t* s, uint32_t n, uint32_t pool, uint32_t slice_lo, uint32_t slice_mask) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t r = mix32(i * 2654435761u + 12345u);
        uint32_t key = (r % 10u) < 7u ? mix32((mix32(i) % pool) * 40503u + 7u) : mix32(i ^ 0xabcdef01u);
        s[i] = slice_lo | (key & slice_mask);
    }
}

// ---- alternative stream readers (exploration) -------------------------------------------------------------
// V2: same TMA ring, but only lane 0 of each warp polls the mbarrier.
template <int PER, int STAGES, int MIN_CTAS, int MODE>
__global__ void __launch_bounds__(kApplyThreads, MIN_CTAS) apply_v2(const uint32_t* __restrict__ stream, const uint32_t* __restrict__ cursor,
                                                                    uint32_t cap, uint32_t* __restrict__ count) {
    constexpr int kTileN = kApplyThreads * PER;
    extern __shared__ __align__(128) uint32_t apply_smem[];
    uint32_t* tiles = apply_smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(apply_smem + STAGES * kTileN);
    uint64_t* empty = full + STAGES;
    const uint32_t n = min(*cursor, cap);
    const uint32_t ntiles = (n + kTileN - 1) / kTileN;
    if (blockIdx.x >= ntiles) return;
    const uint32_t mine = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int lane = threadIdx.x & 31;
    auto issue = [&](uint32_t j) {
        uint32_t stg = j % STAGES;
        if (j >= (uint32_t)STAGES) mbar_wait(&empty[stg], ((j / STAGES) - 1u) & 1u);
        uint32_t first = (blockIdx.x + j * gridDim.x) * kTileN;
        uint32_t bytes = (min((uint32_t)kTileN, n - first) * 4u + 15u) & ~15u;
        mbar_expect_tx(&full[stg], bytes);
        bulk_load_evict_first(tiles + stg * kTileN, stream + first, bytes, &full[stg]);
    };
    if (threadIdx.x == 0) {
        for (int s2 = 0; s2 < STAGES; ++s2) { mbar_init(&full[s2], 1); mbar_init(&empty[s2], kApplyThreads / 32); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (uint32_t j = 0; j < min(mine, (uint32_t)STAGES - 1); ++j) issue(j);
    }
    __syncthreads();
    uint32_t sink = 0;
    for (uint32_t i = 0; i < mine; ++i) {
        if (threadIdx.x == 0 && i + STAGES - 1 < mine) issue(i + STAGES - 1);
        uint32_t stg = i % STAGES;
        if (lane == 0) mbar_wait(&full[stg], (i / STAGES) & 1u);
        __syncwarp();
        uint32_t first = (blockIdx.x + i * gridDim.x) * kTileN;
        uint32_t valid = min((uint32_t)kTileN, n - first);
        const uint32_t* tile = tiles + stg * kTileN;
        uint32_t h[PER], seen[PER];
        bool ok[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) { uint32_t x = threadIdx.x + q * kApplyThreads; ok[q] = x < valid; h[q] = tile[x]; }
#pragma unroll
        for (int q = 0; q < PER; ++q) if (ok[q]) seen[q] = ld_table(count + (h[q] >> 4));
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[stg]);
        if (MODE == 0) bump_batch<PER>(count, h, seen, ok);
        else {
#pragma unroll
            for (int q = 0; q < PER; ++q) if (ok[q]) sink += seen[q];
        }
    }
    if (MODE != 0 && sink == 0x9e3779b9u) count[0] = sink;
}

// V3: no TMA, no shared memory.  Each thread reads its hashes with two 16-byte loads and a real instruction consumes
// every loaded register (xor with a run-time zero) before the first table probe is issued, so the stream loads and
// the probes never share a scoreboard.
template <int PER, int MIN_CTAS, int MODE>
__global__ void __launch_bounds__(kApplyThreads, MIN_CTAS) apply_v3(const uint32_t* __restrict__ stream, const uint32_t* __restrict__ cursor,
                                                                    uint32_t cap, uint32_t* __restrict__ count, uint32_t zero) {
    static_assert(PER % 4 == 0, "");
    const uint32_t n = min(*cursor, cap);
    const uint32_t n4 = n >> 2;                                   // tail ignored here (bench only)
    uint32_t tid = blockIdx.x * kApplyThreads + threadIdx.x, nthreads = gridDim.x * kApplyThreads;
    uint32_t sink = 0;
    const uint4* s4 = reinterpret_cast<const uint4*>(stream);
    for (uint32_t base = tid; base < n4; base += nthreads * (PER / 4)) {
        uint32_t h[PER], seen[PER];
        bool ok[PER];
#pragma unroll
        for (int v = 0; v < PER / 4; ++v) {
            uint32_t at = base + v * nthreads;
            bool in = at < n4;
            uint4 x = in ? __ldg(s4 + at) : make_uint4(0, 0, 0, 0);
            h[4 * v] = x.x; h[4 * v + 1] = x.y; h[4 * v + 2] = x.z; h[4 * v + 3] = x.w;
            ok[4 * v] = ok[4 * v + 1] = ok[4 * v + 2] = ok[4 * v + 3] = in;
        }
#pragma unroll
        for (int q = 0; q < PER; ++q) h[q] ^= zero;               // consume the stream registers here
#pragma unroll
        for (int q = 0; q < PER; ++q) if (ok[q]) seen[q] = ld_table(count + (h[q] >> 4));
        if (MODE == 0) bump_batch<PER>(count, h, seen, ok);
        else {
#pragma unroll
            for (int q = 0; q < PER; ++q) if (ok[q]) sink += seen[q];
        }
    }
    if (MODE != 0 && sink == 0x9e3779b9u) count[0] = sink;
}

template <class K, class... A>
static void time_kernel(const char* name, K kern, int grid, size_t smem, uint32_t* count, uint32_t n, A... args) {
    if (smem) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kApplyThreads, smem));
    float best = 1e9f;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(count, 0, (size_t)1 << 30));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(a));
        kern<<<grid, kApplyThreads, smem>>>(args...);
        CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaEventRecord(a));
    kern<<<grid, kApplyThreads, smem>>>(args...);
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms2; CK(cudaEventElapsedTime(&ms2, a, b));
    printf("%-44s occ=%d grid=%5d  zeroed table: %7.3f ms (%6.1f Gprobe/s)   populated: %7.3f ms (%6.1f Gprobe/s)\n", name, occ, grid, best,
           n / best / 1e6, ms2, n / ms2 / 1e6);
    fflush(stdout);
}

template <int PER, int STAGES, int MINB, int MODE>
static void run(const char* name, const uint32_t* stream, const uint32_t* cursor, uint32_t n, uint32_t* count, int grid_mult) {
    auto kern = s1_apply_kernel<PER, STAGES, MINB, MODE>;
    size_t smem = apply_smem_bytes<PER, STAGES>();
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kApplyThreads, smem));
    float best = 1e9f, first = 0;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(count, 0, (size_t)1 << 30));
        CK(cudaDeviceSynchronize());
        cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
        CK(cudaEventRecord(a));
        kern<<<148 * grid_mult, kApplyThreads, smem>>>(stream, cursor, n, count);
        CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (rep == 0) first = ms;
        if (ms < best) best = ms;
    }
    // second application on the now-populated table: almost every probe finds a saturated counter
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    CK(cudaEventRecord(a));
    kern<<<148 * grid_mult, kApplyThreads, smem>>>(stream, cursor, n, count);
    CK(cudaEventRecord(b)); CK(cudaDeviceSynchronize());
    float ms2; CK(cudaEventElapsedTime(&ms2, a, b));
    printf("%-26s per=%2d stages=%d minctas=%d occ=%d grid=148x%d  zeroed table: %7.3f ms (%6.1f Gprobe/s; first %7.3f)   populated: %7.3f ms (%6.1f Gprobe/s)\n",
           name, PER, STAGES, MINB, occ, grid_mult, best, n / best / 1e6, first, ms2, n / ms2 / 1e6);
    fflush(stdout);
}

int main() {
    uint32_t n = 200u << 20;                       // ~ one busy stream of a 5 M-read mate
    uint32_t *stream, *cursor, *count;
    CK(cudaMalloc(&stream, (size_t)n * 4)); CK(cudaMalloc(&cursor, 4)); CK(cudaMalloc(&count, (size_t)1 << 30));
    CK(cudaMemcpy(cursor, &n, 4, cudaMemcpyHostToDevice));
    make_stream<<<148 * 8, 256>>>(stream, n, 2500000u, 1u << 28, (1u << 28) - 1);
    CK(cudaDeviceSynchronize());
    run<8, 4, 4, 0>("product (CAS, checked)", stream, cursor, n, count, 4);
    run<8, 4, 4, 1>("loads only", stream, cursor, n, count, 4);
    run<4, 4, 8, 0>("product (CAS, checked)", stream, cursor, n, count, 8);
    run<4, 4, 8, 1>("loads only", stream, cursor, n, count, 8);
    time_kernel("v2 lane0 polls, per 8, 4 CTAs", apply_v2<8, 4, 4, 0>, 148 * 4, apply_smem_bytes<8, 4>(), count, n, stream, cursor, n, count);
    time_kernel("v2 lane0 polls, per 8, 4 CTAs, loads only", apply_v2<8, 4, 4, 1>, 148 * 4, apply_smem_bytes<8, 4>(), count, n, stream, cursor, n, count);
    time_kernel("v2 lane0 polls, per 4, 8 CTAs", apply_v2<4, 4, 8, 0>, 148 * 8, apply_smem_bytes<4, 4>(), count, n, stream, cursor, n, count);
    time_kernel("v2 lane0 polls, per 4, 8 CTAs, loads only", apply_v2<4, 4, 8, 1>, 148 * 8, apply_smem_bytes<4, 4>(), count, n, stream, cursor, n, count);
    time_kernel("v3 plain loads, per 8, 4 CTAs", apply_v3<8, 4, 0>, 148 * 4, 0, count, n, stream, cursor, n, count, 0u);
    time_kernel("v3 plain loads, per 8, 4 CTAs, loads only", apply_v3<8, 4, 1>, 148 * 4, 0, count, n, stream, cursor, n, count, 0u);
    time_kernel("v3 plain loads, per 8, 4 CTAs, grid x32", apply_v3<8, 4, 0>, 148 * 32, 0, count, n, stream, cursor, n, count, 0u);
    time_kernel("v3 plain loads, per 4, 8 CTAs", apply_v3<4, 8, 0>, 148 * 8, 0, count, n, stream, cursor, n, count, 0u);
    time_kernel("v3 plain loads, per 4, 8 CTAs, loads only", apply_v3<4, 8, 1>, 148 * 8, 0, count, n, stream, cursor, n, count, 0u);
    time_kernel("v3 plain loads, per 16, 2 CTAs", apply_v3<16, 2, 0>, 148 * 2, 0, count, n, stream, cursor, n, count, 0u);
    time_kernel("v3 plain loads, per 16, 4 CTAs", apply_v3<16, 4, 0>, 148 * 4, 0, count, n, stream, cursor, n, count, 0u);
    time_kernel("v3 plain loads, per 16, 4 CTAs, loads only", apply_v3<16, 4, 1>, 148 * 4, 0, count, n, stream, cursor, n, count, 0u);
    return 0;
}
