// sm_100a kernels of the LocalHGT k-mer screen.  Integer hashing + HBM gathers; no tensor cores.
// "E:" = reference src/extract_ref_normal_peak.cpp (cited for semantics only; nothing here is a
// translation of it — see DESIGN.md §4 for the data-parallel forms these kernels evaluate).
#include "lhgt_kernels.cuh"
#include <cstdlib>

namespace lhgt {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kSMs = 148;

// ------------------------------------------------------------------------------------------------
// bit-plane hashing (semantics E:786-813 == E:1052-1081 == E:430-452)
// ------------------------------------------------------------------------------------------------

// ASCII -> 4 bits: plane0 (A,T) | plane1 (A,C) << 1 | plane2 (A,G) << 2 | valid << 3   (E:1109-1154)
__device__ __forceinline__ uint32_t base_bits(uint32_t ch) {
    uint32_t up = ch & 0xDFu;
    uint32_t a = up == 65u, c = up == 67u, g = up == 71u, t = up == 84u;
    return (a | t) | ((a | c) << 1) | ((a | g) << 2) | ((a | c | g | t) << 3);
}

// Planes are stored big-endian in bit order: position p is bit (31 - p%32) of word p/32, so the
// 32 positions starting at j come out of one funnel shift with position j in bit 31.
__device__ __forceinline__ uint32_t window32(const uint32_t* plane, int j) {
    int q = j >> 5;
    return __funnelshift_l(plane[q + 1], plane[q], j & 31);
}

struct KmerWin {
    uint32_t w2, x0, x1;        // forward:  W2, W0^W2, W1^W2
    uint32_t nrw2, nrx0, rx1;   // mirrored: ~rev(W2), ~rev(W0^W2), rev(W1^W2)
    bool valid;
};

template <int STRIDE>
__device__ __forceinline__ KmerWin make_win(const uint32_t* planes, int j, const HashP& hp) {
    KmerWin kw;
    uint32_t a = window32(planes, j) >> hp.shr;
    uint32_t b = window32(planes + STRIDE, j) >> hp.shr;
    uint32_t c = window32(planes + 2 * STRIDE, j) >> hp.shr;
    uint32_t v = window32(planes + 3 * STRIDE, j) >> hp.shr;
    kw.valid = v == hp.kmask;
    kw.w2 = c;
    kw.x0 = a ^ c;
    kw.x1 = b ^ c;
    kw.nrw2 = ~(__brev(c) >> hp.shr);
    kw.nrx0 = ~(__brev(kw.x0) >> hp.shr);
    kw.rx1 = __brev(kw.x1) >> hp.shr;
    return kw;
}

__device__ __forceinline__ uint32_t hash_of(const KmerWin& kw, const HashP& hp, int i) {
    uint32_t f = kw.w2 ^ (kw.x0 & hp.m0[i]) ^ (kw.x1 & hp.m1[i]);
    uint32_t r = (kw.nrw2 ^ (kw.nrx0 & hp.m0[i]) ^ (kw.rx1 & hp.m1[i])) & hp.kmask;
    return min(f, r);
}

// ---- rolling pack + hash: no shared-memory planes ------------------------------------------------------------
// A warp walks a read 32 bases at a time.  Word w of the four planes (bit p%32 of word p/32 = position p: ballots
// deliver exactly that) lives in four warp-uniform registers; the 32-position chunk c needs words c and c+1, and
// lane l's window (positions 32c+l ...) is one funnel shift.  In this little-endian window position j+z sits at
// bit z, which is the form the reverse-complement hash wants; one BREV per plane gives the forward form
// (the two use the same masks).  Positions whose window runs past the read see validity bits 0.
struct Planes { uint32_t p0, p1, p2, pv; };

__device__ __forceinline__ void fill_base_lut(uint8_t* lut) {           // call with all threads, then __syncthreads
    for (int c = threadIdx.x; c < 256; c += blockDim.x) lut[c] = (uint8_t)base_bits((uint32_t)c);
}

__device__ __forceinline__ Planes pack_word(uint32_t ch, const uint8_t* lut) {
    uint32_t bits = lut[ch];
    Planes w;
    w.p0 = __ballot_sync(kFull, bits & 1u);
    w.p1 = __ballot_sync(kFull, bits & 2u);
    w.p2 = __ballot_sync(kFull, bits & 4u);
    w.pv = __ballot_sync(kFull, bits & 8u);
    return w;
}

struct LeWin {                       // per-lane k-mer window, both orientations
    uint32_t w2, wx0, wx1;           // forward:   W2, W0^W2, W1^W2   (base z at bit k-1-z)
    uint32_t n2, nx0, x1;            // mirrored: ~L2, ~(L0^L2), L1^L2 (base z at bit z; masked by kmask at the end)
    bool valid;
};

__device__ __forceinline__ LeWin le_window(const Planes& lo, const Planes& hi, int lane, const HashP& hp) {
    uint32_t l0 = __funnelshift_r(lo.p0, hi.p0, lane), l1 = __funnelshift_r(lo.p1, hi.p1, lane);
    uint32_t l2 = __funnelshift_r(lo.p2, hi.p2, lane), lv = __funnelshift_r(lo.pv, hi.pv, lane);
    LeWin k;
    k.valid = (lv & hp.kmask) == hp.kmask;
    uint32_t x0 = l0 ^ l2, x1 = l1 ^ l2;
    k.n2 = ~l2; k.nx0 = ~x0; k.x1 = x1;
    k.w2 = __brev(l2) >> hp.shr; k.wx0 = __brev(x0) >> hp.shr; k.wx1 = __brev(x1) >> hp.shr;
    return k;
}

__device__ __forceinline__ uint32_t le_hash(const LeWin& k, const HashP& hp, int i) {
    uint32_t f = k.w2 ^ (k.wx0 & hp.m0[i]) ^ (k.wx1 & hp.m1[i]);
    uint32_t r = (k.n2 ^ (k.nx0 & hp.m0[i]) ^ (k.x1 & hp.m1[i])) & hp.kmask;
    return min(f, r);
}

// ------------------------------------------------------------------------------------------------
// 2-bit saturating count table: entry g = tbl_index(h) lives in bits [2*(g&15), +2) of word g>>4
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_table(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- TMA (cp.async.bulk) + mbarrier plumbing ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LHGT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LHGT_DONE;\n"
        "bra LHGT_WAIT;\n"
        "LHGT_DONE:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// 1-D bulk copy shared -> global (bulk async-group completion)
__device__ __forceinline__ void bulk_store(void* dst, const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_store_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- read staging -------------------------------------------------------------------------------------------
// A warp has the bytes of the read it will hash NEXT copied into its own shared-memory slot by the TMA unit (one
// cp.async.bulk issued by lane 0, completion counted on the slot's mbarrier), so no lane ever waits on a FASTQ byte
// in DRAM: by the time a read is hashed its bytes have been on chip for a whole read's worth of work, and no
// register or load scoreboard is held meanwhile.  A slot receives the 16-byte granules that cover [start, start+len).
constexpr int kStageBytes = 528;                                       // 15 bytes of misalignment + kMaxReadLen, in granules
static_assert(kStageBytes % 16 == 0 && kStageBytes >= 15 + kMaxReadLen, "stage slot too small");

// fq must be 16-byte aligned (device allocations are; lhgt_reads_attach_device checks).  Returns the offset of the
// read's first byte inside the slot.  The last granule may run past start + len: it stays inside the 16-byte block
// that holds the read's last byte.  Call with len > 0; `bytes_out` accumulates what the barrier has to expect.
__device__ __forceinline__ uint32_t stage_granules(uint64_t start, uint32_t len) { return ((uint32_t)start & 15u) + len + 15u >> 4; }
__device__ __forceinline__ uint32_t stage_read(uint8_t* slot, uint64_t* bar, const uint8_t* __restrict__ fq, uint64_t start, uint32_t len, int lane) {
    uint32_t mis = (uint32_t)start & 15u;
    if (lane == 0) {
        uint32_t bytes = stage_granules(start, len) << 4;
        mbar_expect_tx(bar, bytes);
        bulk_load(slot, fq + (start - mis), bytes, bar);
    }
    return mis;
}
// S3 stages the two mates of a pair behind ONE barrier
__device__ __forceinline__ void stage_pair_reads(uint8_t* slot1, uint8_t* slot2, uint64_t* bar, const uint8_t* __restrict__ fq1, uint64_t s1,
                                                 uint32_t l1, const uint8_t* __restrict__ fq2, uint64_t s2, uint32_t l2, int lane) {
    if (lane == 0) {
        uint32_t b1 = l1 ? stage_granules(s1, l1) << 4 : 0u, b2 = l2 ? stage_granules(s2, l2) << 4 : 0u;
        mbar_expect_tx(bar, b1 + b2);
        if (b1) bulk_load(slot1, fq1 + (s1 & ~15ull), b1, bar);
        if (b2) bulk_load(slot2, fq2 + (s2 & ~15ull), b2, bar);
    }
}

// count[h] = min(3, count[h] + 1), exact under any interleaving (E:1082-1084 made race-free).  Here and in
// bump_batch `h` is the TABLE INDEX of the hash (tbl_index), not the hash itself.
__device__ __forceinline__ void bump(uint32_t* count, uint32_t h, uint32_t seen) {
    uint32_t* addr = count + (h >> 4);
    int sh = (h & 15u) * 2;
    while (((seen >> sh) & 3u) < 3u) {
        uint32_t old = atomicCAS(addr, seen, seen + (1u << sh));
        if (old == seen) break;
        seen = old;
    }
}

// Saturating increment of N counters whose words were loaded into seen[].  Every round issues the compare-and-swaps
// of all still-pending probes back to back and only then looks at the results, so a round costs one L2 round trip
// however many probes it carries; a probe that lost its word to a neighbour (16 counters share a word) re-decides on
// the value the failed CAS returned and goes again in the next round.
template <int N>
__device__ __forceinline__ void bump_batch(uint32_t* count, const uint32_t (&h)[N], uint32_t (&seen)[N], const bool (&ok)[N]) {
    uint32_t pend = 0;
#pragma unroll
    for (int q = 0; q < N; ++q)
        if (ok[q] && ((seen[q] >> ((h[q] & 15u) * 2)) & 3u) < 3u) pend |= 1u << q;
    while (pend) {
        uint32_t got[N];
#pragma unroll
        for (int q = 0; q < N; ++q)
            if (pend & (1u << q)) got[q] = atomicCAS(count + (h[q] >> 4), seen[q], seen[q] + (1u << ((h[q] & 15u) * 2)));
#pragma unroll
        for (int q = 0; q < N; ++q)
            if (pend & (1u << q)) {
                if (got[q] == seen[q] || ((got[q] >> ((h[q] & 15u) * 2)) & 3u) == 3u) pend &= ~(1u << q);
                seen[q] = got[q];
            }
    }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of u32 (tile counts): 1024 elements per block, recursive over block sums
// ------------------------------------------------------------------------------------------------
constexpr int kScanBlock = 256, kScanItems = 4, kScanSpan = kScanBlock * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total, uint32_t* smem /*>=9*/) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < (int)(blockDim.x >> 5) ? smem[lane] : 0;
        uint32_t winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(kFull, winc, d);
            if (lane >= d) winc += t;
        }
        if (lane < (int)(blockDim.x >> 5)) smem[lane] = winc - w;
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    uint32_t res = smem[warp] + inc - v;
    *total = smem[32];
    __syncthreads();
    return res;
}

template <class T>
__global__ void __launch_bounds__(kScanBlock) scan_block_kernel(const T* __restrict__ in, T* __restrict__ out,
                                                                uint64_t n, T* __restrict__ block_sums) {
    __shared__ uint32_t sm[33];
    __shared__ T wide[9];
    uint64_t base = (uint64_t)blockIdx.x * kScanSpan + (uint64_t)threadIdx.x * kScanItems;
    T v[kScanItems], sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { v[i] = base + i < n ? in[base + i] : (T)0; sum += v[i]; }
    T ex;
    if constexpr (sizeof(T) == 4) {
        uint32_t total, e32 = block_exclusive_scan((uint32_t)sum, &total, sm);
        ex = e32;
        if (threadIdx.x == 0 && block_sums) block_sums[blockIdx.x] = total;
    } else {                                                     // 64-bit sums: warp scan in registers, warp totals through shared memory
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        T inc = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T t = __shfl_up_sync(kFull, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) wide[warp] = inc;
        __syncthreads();
        T before = 0, total = 0;
#pragma unroll
        for (int q = 0; q < kScanBlock / 32; ++q) { T w = wide[q]; before += q < warp ? w : (T)0; total += w; }
        ex = before + inc - sum;
        if (threadIdx.x == 0 && block_sums) block_sums[blockIdx.x] = total;
    }
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
}

template <class T>
__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(T* __restrict__ out, uint64_t n, const T* __restrict__ block_prefix) {
    uint64_t base = (uint64_t)blockIdx.x * kScanSpan + (uint64_t)threadIdx.x * kScanItems;
    T add = block_prefix[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) if (base + i < n) out[base + i] += add;
}

size_t scan_tmp_words(uint64_t n) {
    size_t words = 0;
    while (n > 1) { n = (n + kScanSpan - 1) / kScanSpan; words += 2 * n; }
    return words + 2;
}

template <class T>
static int scan_exclusive_t(const T* in, T* out, uint64_t n, T* tmp, cudaStream_t st) {
    if (n == 0) return 0;
    uint64_t blocks = (n + kScanSpan - 1) / kScanSpan;
    int launches = 0;
    if (blocks == 1) {
        scan_block_kernel<T><<<1, kScanBlock, 0, st>>>(in, out, n, (T*)nullptr);
        return 1;
    }
    T* sums = tmp;
    T* sums_scanned = tmp + blocks;
    scan_block_kernel<T><<<(unsigned)blocks, kScanBlock, 0, st>>>(in, out, n, sums);
    launches += 1;
    launches += scan_exclusive_t<T>(sums, sums_scanned, blocks, tmp + 2 * blocks, st);
    scan_add_kernel<T><<<(unsigned)blocks, kScanBlock, 0, st>>>(out, n, sums_scanned);
    return launches + 1;
}

int launch_scan_exclusive(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* tmp, cudaStream_t st) {
    return scan_exclusive_t<uint32_t>(in, out, n, tmp, st);
}
// 64-bit sums (FASTA files beyond 4 GB of sequence); tmp holds scan_tmp_words(n) elements of 8 bytes
int launch_scan_exclusive64(const uint64_t* in, uint64_t* out, uint64_t n, uint64_t* tmp, cudaStream_t st) {
    return scan_exclusive_t<unsigned long long>((const unsigned long long*)in, (unsigned long long*)out, n, (unsigned long long*)tmp, st);
}

// ------------------------------------------------------------------------------------------------
// FASTQ record location (the getline loops of E:1020-1034 / E:356-409 as a newline scan)
// ------------------------------------------------------------------------------------------------
constexpr int kFqThreads = 256, kFqIter = 4, kFqChunk = 16, kFqSub = kFqThreads * kFqChunk, kFqTile = kFqSub * kFqIter;

uint64_t fastq_index_tiles(uint64_t n) { return (n + kFqTile - 1) / kFqTile; }

__device__ __forceinline__ void load16(const uint8_t* __restrict__ fq, uint64_t off, uint64_t n, uint32_t w[4]) {
    if (off + 16 <= n) {
        uint4 v = *reinterpret_cast<const uint4*>(fq + off);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        w[0] = w[1] = w[2] = w[3] = 0;
        for (int b = 0; b < 16; ++b)
            if (off + b < n) w[b >> 2] |= (uint32_t)fq[off + b] << ((b & 3) * 8);
    }
}

__global__ void __launch_bounds__(kFqThreads) fq_count_kernel(const uint8_t* __restrict__ fq, uint64_t n,
                                                              uint32_t* __restrict__ tile_cnt) {
    __shared__ uint32_t sm[8];
    uint64_t tile0 = (uint64_t)blockIdx.x * kFqTile;
    uint32_t c = 0;
#pragma unroll
    for (int it = 0; it < kFqIter; ++it) {
        uint64_t off = tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk;
        if (off < n) {
            uint32_t w[4];
            load16(fq, off, n, w);
#pragma unroll
            for (int q = 0; q < 4; ++q) c += __popc(__vcmpeq4(w[q], 0x0a0a0a0au)) >> 3;
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(kFull, c, d);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kFqThreads / 32; ++w) t += sm[w];
        tile_cnt[blockIdx.x] = t;
    }
}

// newline with global rank g ends line g: line 4r+1 is the sequence of record r.
// A thread owns kFqIter 16-byte chunks, iteration-major (coalesced); the ranks of its newlines need the exclusive scan
// of the per-chunk counts in file order = (iteration, thread) order.  All kFqIter scans run as ONE block scan over
// kFqIter-vectors (same barriers as a scalar scan), then the iteration totals are chained.
__global__ void __launch_bounds__(kFqThreads) fq_assign_kernel(const uint8_t* __restrict__ fq, uint64_t n,
                                                               const uint32_t* __restrict__ tile_base,
                                                               uint64_t* __restrict__ rec_start,
                                                               uint64_t* __restrict__ rec_end, uint64_t rec_cap) {
    __shared__ uint32_t wsum[kFqIter][kFqThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kFqTile;
    uint32_t m[kFqIter][4], c[kFqIter], inc[kFqIter];
#pragma unroll
    for (int it = 0; it < kFqIter; ++it) {
        uint64_t off = tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk;
        uint32_t w[4] = {0, 0, 0, 0};
        if (off < n) load16(fq, off, n, w);
        c[it] = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { m[it][q] = off < n ? __vcmpeq4(w[q], 0x0a0a0a0au) : 0u; c[it] += __popc(m[it][q]) >> 3; }
        inc[it] = c[it];
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
#pragma unroll
        for (int it = 0; it < kFqIter; ++it) {
            uint32_t t = __shfl_up_sync(kFull, inc[it], d);
            if (lane >= d) inc[it] += t;
        }
    if (lane == 31)
#pragma unroll
        for (int it = 0; it < kFqIter; ++it) wsum[it][warp] = inc[it];
    __syncthreads();
    uint64_t running = tile_base[blockIdx.x];
#pragma unroll
    for (int it = 0; it < kFqIter; ++it) {
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int q = 0; q < kFqThreads / 32; ++q) { uint32_t v = wsum[it][q]; before += q < warp ? v : 0u; total += v; }
        uint64_t g = running + before + inc[it] - c[it];
        uint64_t off = tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t mm = m[it][q];
            while (mm) {
                int bit = __ffs(mm) - 1;
                mm &= ~(0xffu << (bit & ~7));
                uint64_t pos = off + q * 4 + (bit >> 3);
                uint64_t r = g >> 2;
                if (r < rec_cap) {
                    if ((g & 3) == 0) rec_start[r] = pos + 1;
                    else if ((g & 3) == 1) rec_end[r] = pos;
                }
                ++g;
            }
        }
        running += total;
    }
}

int launch_fastq_index(const uint8_t* fq, uint64_t n, uint32_t* tile_cnt, uint32_t* tile_base, uint32_t* scan_tmp,
                       uint64_t* rec_start, uint64_t* rec_end, uint64_t rec_cap, int phase, cudaStream_t st) {
    uint64_t tiles = fastq_index_tiles(n);
    if (tiles == 0) return 0;
    if (phase == 0) {
        fq_count_kernel<<<(unsigned)tiles, kFqThreads, 0, st>>>(fq, n, tile_cnt);
        return 1 + launch_scan_exclusive(tile_cnt, tile_base, tiles, scan_tmp, st);
    }
    fq_assign_kernel<<<(unsigned)tiles, kFqThreads, 0, st>>>(fq, n, tile_base, rec_start, rec_end, rec_cap);
    return 1;
}

// out[0] += sum of sequence-line lengths, out[1] = max(out[1], longest)
__global__ void sum_lengths_kernel(const uint64_t* __restrict__ s, const uint64_t* __restrict__ e, uint64_t n,
                                   unsigned long long* out) {
    unsigned long long acc = 0, longest = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        unsigned long long l = e[i] - s[i];
        acc += l;
        longest = l > longest ? l : longest;
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        acc += __shfl_xor_sync(kFull, acc, d);
        unsigned long long o = __shfl_xor_sync(kFull, longest, d);
        longest = o > longest ? o : longest;
    }
    if ((threadIdx.x & 31) == 0 && acc) { atomicAdd(out, acc); atomicMax(out + 1, longest); }
}

// ------------------------------------------------------------------------------------------------
// FASTA ingest on the device (the getline loop of read_ref, E:761-831, as a stream compaction): the raw file bytes
// are uploaded once; a byte is a sequence byte iff it is not '\n' and not inside a header line.  Header lines are
// few: the host finds them (memchr) and passes their byte spans.  Same tile geometry as the FASTQ scan.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t fa_keep_mask(const uint8_t* __restrict__ fa, uint64_t off, uint64_t n,
                                                 const ByteSpan* __restrict__ spans, uint32_t nspans) {   // bit b: byte off+b is kept
    if (off >= n) return 0u;
    uint32_t w[4];
    load16(fa, off, n, w);
    uint32_t keep = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        uint32_t nl = __vcmpeq4(w[q], 0x0a0a0a0au);             // 0xff per newline byte
        uint32_t bits = ((~nl) & 0x01010101u) * 0x01020408u >> 24;   // 4 bits: byte j kept -> bit j
        keep |= (bits & 0xfu) << (q * 4);
    }
    if (off + 16 > n) keep &= (1u << (uint32_t)(n - off)) - 1u;
    // header spans that overlap [off, off + 16): first span with hi >= off
    uint32_t lo = 0, hi = nspans;
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (spans[mid].hi < off) lo = mid + 1; else hi = mid;
    }
    for (; lo < nspans && spans[lo].lo < off + 16; ++lo) {
        uint64_t a = spans[lo].lo > off ? spans[lo].lo - off : 0, b = spans[lo].hi - off;   // local [a, b], b may exceed 15
        uint32_t upto = b >= 15 ? 0xffffu : (1u << (uint32_t)(b + 1)) - 1u;
        keep &= ~(upto & ~((1u << (uint32_t)a) - 1u));
    }
    return keep;
}

__global__ void __launch_bounds__(kFqThreads) fa_count_kernel(const uint8_t* __restrict__ fa, uint64_t n, const ByteSpan* __restrict__ spans,
                                                              uint32_t nspans, uint64_t* __restrict__ tile_cnt) {
    __shared__ uint32_t sm[8];
    uint64_t tile0 = (uint64_t)blockIdx.x * kFqTile;
    uint32_t c = 0;
#pragma unroll
    for (int it = 0; it < kFqIter; ++it)
        c += __popc(fa_keep_mask(fa, tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk, n, spans, nspans));
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(kFull, c, d);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kFqThreads / 32; ++w) t += sm[w];
        tile_cnt[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(kFqThreads) fa_compact_kernel(const uint8_t* __restrict__ fa, uint64_t n, const ByteSpan* __restrict__ spans,
                                                                uint32_t nspans, const uint64_t* __restrict__ tile_base,
                                                                uint8_t* __restrict__ out) {
    __shared__ uint32_t wsum[kFqIter][kFqThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t tile0 = (uint64_t)blockIdx.x * kFqTile;
    uint32_t keep[kFqIter], c[kFqIter], inc[kFqIter];
#pragma unroll
    for (int it = 0; it < kFqIter; ++it) {
        keep[it] = fa_keep_mask(fa, tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk, n, spans, nspans);
        c[it] = __popc(keep[it]);
        inc[it] = c[it];
    }
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
#pragma unroll
        for (int it = 0; it < kFqIter; ++it) {
            uint32_t t = __shfl_up_sync(kFull, inc[it], d);
            if (lane >= d) inc[it] += t;
        }
    if (lane == 31)
#pragma unroll
        for (int it = 0; it < kFqIter; ++it) wsum[it][warp] = inc[it];
    __syncthreads();
    uint64_t running = tile_base[blockIdx.x];
#pragma unroll
    for (int it = 0; it < kFqIter; ++it) {
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int q = 0; q < kFqThreads / 32; ++q) { uint32_t v = wsum[it][q]; before += q < warp ? v : 0u; total += v; }
        uint64_t g = running + before + inc[it] - c[it];
        uint64_t off = tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk;
        uint32_t mm = keep[it];
        if (mm == 0xffffu && (g & 15u) == 0) {                  // a whole kept chunk landing on a 16-byte boundary: one vector store
            *reinterpret_cast<uint4*>(out + g) = *reinterpret_cast<const uint4*>(fa + off);
        } else {
            while (mm) {
                int b = __ffs(mm) - 1;
                mm &= mm - 1;
                out[g++] = fa[off + b];
            }
        }
        running += total;
    }
}

// out[i] = number of kept bytes before byte position pos[i] (pos[i] <= n); one CTA per query
__global__ void __launch_bounds__(kFqThreads) fa_offsets_kernel(const uint8_t* __restrict__ fa, uint64_t n, const ByteSpan* __restrict__ spans,
                                                                uint32_t nspans, const uint64_t* __restrict__ tile_base, uint64_t ntiles,
                                                                uint64_t* __restrict__ out) {
    __shared__ uint32_t sm[8];
    uint64_t q = blockIdx.x < nspans ? spans[blockIdx.x].lo : n;   // query i < nspans: header i's first byte; query nspans: end of file
    uint64_t tile = q / kFqTile;
    uint32_t c = 0;
    uint64_t base = 0;
    if (tile < ntiles) {
        base = tile_base[tile];
        uint64_t tile0 = tile * kFqTile;
        for (int it = 0; it < kFqIter; ++it) {
            uint64_t off = tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk;
            if (off >= q) continue;
            uint32_t m = fa_keep_mask(fa, off, n, spans, nspans);
            if (q - off < 16) m &= (1u << (uint32_t)(q - off)) - 1u;
            c += __popc(m);
        }
    } else if (ntiles) {                                       // q == n on a tile boundary: everything
        base = tile_base[ntiles - 1];
        uint64_t tile0 = (ntiles - 1) * kFqTile;
        for (int it = 0; it < kFqIter; ++it)
            c += __popc(fa_keep_mask(fa, tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk, n, spans, nspans));
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) c += __shfl_xor_sync(kFull, c, d);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t t = base;
        for (int w = 0; w < kFqThreads / 32; ++w) t += sm[w];
        out[blockIdx.x] = t;
    }
}

// Header lines (E:762: a line whose first byte is '>'): every '>' that follows a newline or opens the file.  They are
// few, so the thread that meets one walks to the end of its line itself and appends the span; the host sorts the list.
__global__ void __launch_bounds__(kFqThreads) fa_headers_kernel(const uint8_t* __restrict__ fa, uint64_t n, ByteSpan* __restrict__ spans,
                                                                uint32_t cap, unsigned long long* __restrict__ count) {
    uint64_t tile0 = (uint64_t)blockIdx.x * kFqTile;
#pragma unroll
    for (int it = 0; it < kFqIter; ++it) {
        uint64_t off = tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk;
        if (off >= n) continue;
        uint32_t w[4];
        load16(fa, off, n, w);
        uint32_t gt = 0, nl = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            gt |= ((__vcmpeq4(w[q], 0x3e3e3e3eu) & 0x01010101u) * 0x01020408u >> 24 & 0xfu) << (q * 4);
            nl |= ((__vcmpeq4(w[q], 0x0a0a0a0au) & 0x01010101u) * 0x01020408u >> 24 & 0xfu) << (q * 4);
        }
        if (!gt) continue;
        uint32_t prev_nl = off == 0 ? 1u : (uint32_t)(fa[off - 1] == '\n');
        uint32_t starts = gt & ((nl << 1) | prev_nl) & 0xffffu;
        if (off + 16 > n) starts &= (1u << (uint32_t)(n - off)) - 1u;
        while (starts) {
            int b = __ffs(starts) - 1;
            starts &= starts - 1;
            uint64_t lo = off + b, hi = lo;
            while (hi < n && fa[hi] != '\n') ++hi;
            if (hi == n) hi = n - 1;                             // last line without a newline
            unsigned long long at = atomicAdd(count, 1ull);
            if (at < cap) spans[at] = ByteSpan{lo, hi};
        }
    }
}

// packed[off[i] .. ) = the bytes of header i (its newline excluded), at most `limit` of them; one CTA per header
__global__ void __launch_bounds__(128) fa_header_text_kernel(const uint8_t* __restrict__ fa, const ByteSpan* __restrict__ spans,
                                                             const uint64_t* __restrict__ off, uint8_t* __restrict__ packed) {
    ByteSpan sp = spans[blockIdx.x];
    uint64_t len = off[blockIdx.x + 1] - off[blockIdx.x];
    for (uint64_t i = threadIdx.x; i < len; i += blockDim.x) packed[off[blockIdx.x] + i] = fa[sp.lo + i];
}

int launch_fasta_headers(const uint8_t* fa, uint64_t n, ByteSpan* spans, uint32_t cap, unsigned long long* count, cudaStream_t st) {
    uint64_t tiles = fastq_index_tiles(n);
    if (tiles == 0) return 0;
    fa_headers_kernel<<<(unsigned)tiles, kFqThreads, 0, st>>>(fa, n, spans, cap, count);
    return 1;
}

int launch_fasta_header_text(const uint8_t* fa, const ByteSpan* spans, uint32_t nspans, const uint64_t* off, uint8_t* packed, cudaStream_t st) {
    if (!nspans) return 0;
    fa_header_text_kernel<<<nspans, 128, 0, st>>>(fa, spans, off, packed);
    return 1;
}

int launch_fasta_compact(const uint8_t* fa, uint64_t n, const ByteSpan* spans, uint32_t nspans, uint64_t* tile_cnt, uint64_t* tile_base,
                         uint64_t* scan_tmp, uint8_t* out, uint64_t* pos_out, int phase, cudaStream_t st) {
    uint64_t tiles = fastq_index_tiles(n);
    if (tiles == 0) return 0;
    if (phase == 0) {
        fa_count_kernel<<<(unsigned)tiles, kFqThreads, 0, st>>>(fa, n, spans, nspans, tile_cnt);
        int l = 1 + launch_scan_exclusive64(tile_cnt, tile_base, tiles, scan_tmp, st);
        fa_offsets_kernel<<<nspans + 1, kFqThreads, 0, st>>>(fa, n, spans, nspans, tile_base, tiles, pos_out);
        return l + 1;
    }
    fa_compact_kernel<<<(unsigned)tiles, kFqThreads, 0, st>>>(fa, n, spans, nspans, tile_base, out);
    return 1;
}

int launch_sum_lengths(const uint64_t* rec_start, const uint64_t* rec_end, uint64_t nrec, unsigned long long* out,
                       cudaStream_t st) {
    if (nrec == 0) return 0;
    sum_lengths_kernel<<<kSMs * 4, 256, 0, st>>>(rec_start, rec_end, nrec, out);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// IB: index build.  One block per 1024-position tile: ASCII -> planes in shared memory -> e hashes
// per position, stored position-major/hash-minor exactly as the file holds them (E:785-813).
// ------------------------------------------------------------------------------------------------
constexpr int kIbPlane = kTileWords + 4;   // 33 words used (+1 zero word read by window32)

template <int E>
__global__ void __launch_bounds__(256) index_build_kernel(const uint8_t* __restrict__ seq, const Contig* __restrict__ contigs,
                                                          const Tile* __restrict__ tiles, HashP hp,
                                                          uint32_t* __restrict__ image, uint8_t* __restrict__ valid_out) {
    __shared__ uint32_t planes[4 * kIbPlane];
    const int e = E ? E : hp.e;
    Tile t = tiles[blockIdx.x];
    Contig c = contigs[t.contig];
    long np = (long)c.len - hp.k + 1;
    long j0 = t.j0;
    if (j0 >= np) return;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint8_t* src = seq + c.seq_off + j0;
    long avail = (long)c.len - j0;
    for (int w = warp; w < kTileWords + 2; w += 8) {
        int p = w * 32 + lane;
        uint32_t bits = p < avail ? base_bits(src[p]) : 0u;
        uint32_t b0 = __brev(__ballot_sync(kFull, bits & 1u));
        uint32_t b1 = __brev(__ballot_sync(kFull, bits & 2u));
        uint32_t b2 = __brev(__ballot_sync(kFull, bits & 4u));
        uint32_t b3 = __brev(__ballot_sync(kFull, bits & 8u));
        if (lane == 0) {
            planes[w] = b0; planes[kIbPlane + w] = b1; planes[2 * kIbPlane + w] = b2; planes[3 * kIbPlane + w] = b3;
        }
    }
    __syncthreads();
    if (j0 == 0 && threadIdx.x == 0) image[c.hash_word - 1] = c.len;
    // (staging the tile's hashes in shared memory and storing 16-byte vectors was measured: 142 against 164 Gbp/s)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int jl = r * 256 + threadIdx.x;
        long j = j0 + jl;
        if (j < np) {
            KmerWin kw = make_win<kIbPlane>(planes, jl, hp);
            uint32_t* dst = image + c.hash_word + (size_t)j * e;
#pragma unroll
            for (int i = 0; i < (E ? E : kMaxE); ++i)
                if (i < e) dst[i] = kw.valid ? hash_of(kw, hp, i) : 0u;
            if (valid_out) valid_out[j] = kw.valid;
        }
    }
}

template <int E>
static void ib_launch(const uint8_t* seq, const Contig* contigs, const Tile* tiles, uint64_t ntiles, const HashP& hp,
                      uint32_t* image, uint8_t* valid_out, cudaStream_t st) {
    index_build_kernel<E><<<(unsigned)ntiles, 256, 0, st>>>(seq, contigs, tiles, hp, image, valid_out);
}

int launch_index_build(const uint8_t* seq, const Contig* contigs, const Tile* tiles, uint64_t ntiles, const HashP& hp,
                       uint32_t* image, uint8_t* valid_out, cudaStream_t st) {
    if (ntiles == 0) return 0;
    switch (hp.e) {
        case 1: ib_launch<1>(seq, contigs, tiles, ntiles, hp, image, valid_out, st); break;
        case 2: ib_launch<2>(seq, contigs, tiles, ntiles, hp, image, valid_out, st); break;
        case 3: ib_launch<3>(seq, contigs, tiles, ntiles, hp, image, valid_out, st); break;
        case 4: ib_launch<4>(seq, contigs, tiles, ntiles, hp, image, valid_out, st); break;
        default: ib_launch<0>(seq, contigs, tiles, ntiles, hp, image, valid_out, st); break;
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------
// S1: sampled-read k-mer counting (E:1037-1087).  One warp per read; the read is 2-bit(+valid)
// packed into shared memory by ballots; each lane hashes positions lane, lane+32, ... and issues all
// its table loads before the first compare-and-swap so ~12 independent sectors per lane are in flight.
// ------------------------------------------------------------------------------------------------
constexpr int kS1Warps = 8;

__device__ __forceinline__ bool is_sampled(const uint32_t* __restrict__ sample_bits, uint64_t ordinal) {
    if (!sample_bits) return true;
    uint32_t o = (uint32_t)(ordinal % (uint64_t)kRandomArray);
    return (sample_bits[o >> 5] >> (o & 31)) & 1u;
}

template <int E>
__global__ void __launch_bounds__(kS1Warps * 32, 4) s1_count_kernel(
    const uint8_t* __restrict__ fq, const uint64_t* __restrict__ rec_start, const uint64_t* __restrict__ rec_end,
    uint64_t nrec, uint64_t budget, const uint32_t* __restrict__ sample_bits, uint64_t ordinal_base, HashP hp,
    uint32_t* __restrict__ count, unsigned long long* __restrict__ n_sampled, int* __restrict__ err) {
    __shared__ uint8_t lut[256];
    const int e = E ? E : hp.e;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fill_base_lut(lut);
    __syncthreads();
    unsigned long long mine = 0;
    uint64_t stride = (uint64_t)gridDim.x * kS1Warps;
    for (uint64_t r = (uint64_t)blockIdx.x * kS1Warps + warp; r < nrec; r += stride) {
        uint64_t start = rec_start[r];
        if (start > budget) continue;                       // Q15 (E:1022-1025 with end = size(fq1))
        if (!is_sampled(sample_bits, r + ordinal_base)) continue;
        uint64_t len64 = rec_end[r] - start;
        if (len64 > (uint64_t)kMaxReadLen) { if (lane == 0) atomicExch(err, 1); continue; }
        int len = (int)len64;
        ++mine;
        int np = len - hp.k + 1;
        if (np <= 0) continue;
        const uint8_t* src = fq + start;
        int nch = (np + 31) >> 5;
        Planes prev = pack_word(lane < len ? src[lane] : 0u, lut);
        uint32_t chn = 32 + lane < len ? src[32 + lane] : 0u;
        for (int w = 1; w <= nch; ++w) {
            Planes cur = pack_word(chn, lut);
            int pn = (w + 1) * 32 + lane;
            chn = pn < len ? src[pn] : 0u;
            LeWin kw = le_window(prev, cur, lane, hp);
            prev = cur;
            uint32_t h[E ? E : kMaxE], seen[E ? E : kMaxE];
            bool ok[E ? E : kMaxE];
#pragma unroll
            for (int i = 0; i < (E ? E : kMaxE); ++i) {
                ok[i] = i < e && kw.valid;
                h[i] = i < e ? tbl_index(le_hash(kw, hp, i), hp) : 0u;   // table index from here on
                seen[i] = 0u;
                if (ok[i]) seen[i] = ld_table(count + (h[i] >> 4));
            }
            bump_batch<(E ? E : kMaxE)>(count, h, seen, ok);
        }
    }
    if (lane == 0 && mine) atomicAdd(n_sampled, mine);
}

template <int E>
static void s1_launch(const uint8_t* fq, const uint64_t* rs, const uint64_t* re, uint64_t nrec, uint64_t budget,
                      const uint32_t* sb, uint64_t ob, const HashP& hp, uint32_t* count, unsigned long long* ns, int* err,
                      cudaStream_t st) {
    uint64_t want = (nrec + kS1Warps - 1) / kS1Warps;
    unsigned grid = (unsigned)(want < (uint64_t)kSMs * 4 ? want : (uint64_t)kSMs * 4);
    s1_count_kernel<E><<<grid, kS1Warps * 32, 0, st>>>(fq, rs, re, nrec, budget, sb, ob, hp, count, ns, err);
}

int launch_s1(const uint8_t* fq, const uint64_t* rec_start, const uint64_t* rec_end, uint64_t nrec, uint64_t budget,
              const uint32_t* sample_bits, uint64_t ordinal_base, const HashP& hp, uint32_t* count,
              unsigned long long* n_sampled, int* err, cudaStream_t st) {
    if (nrec == 0) return 0;
    switch (hp.e) {
        case 1: s1_launch<1>(fq, rec_start, rec_end, nrec, budget, sample_bits, ordinal_base, hp, count, n_sampled, err, st); break;
        case 2: s1_launch<2>(fq, rec_start, rec_end, nrec, budget, sample_bits, ordinal_base, hp, count, n_sampled, err, st); break;
        case 3: s1_launch<3>(fq, rec_start, rec_end, nrec, budget, sample_bits, ordinal_base, hp, count, n_sampled, err, st); break;
        case 4: s1_launch<4>(fq, rec_start, rec_end, nrec, budget, sample_bits, ordinal_base, hp, count, n_sampled, err, st); break;
        default: s1_launch<0>(fq, rec_start, rec_end, nrec, budget, sample_bits, ordinal_base, hp, count, n_sampled, err, st); break;
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------
// S1, streamed form.  A 2^k-entry table far larger than L2 makes every direct probe a
// 128-byte DRAM fill (profiles/r01_probe_bench_*: 43 G probes/s), and even an L2-resident slice is capped by
// the L2's atomic rate (profiles/r01_apply_bench_*: ~125 G updates/s).  So counting moves the hashes, not the
// table: the table is stored leaf-major (HashP: a leaf = the hashes sharing a field of middle bits) and
//   P1  s1_bin_kernel   hashes the sampled reads and appends every hash to one of 2^b1 streams chosen by its
//                       first b1 leaf bits (per-CTA shared-memory buckets, flushed as coalesced runs);
//   P2  s1_split_kernel splits every stream by the other b2 bits of the leaf field into leaf streams (tile histogram -> one
//                       reservation per leaf -> ordered scatter through shared memory -> coalesced runs);
//   P3  s1_leaf_kernel  one CTA per leaf: the leaf's table slice (<= 64 KiB) is bulk-copied into shared memory,
//                       the leaf stream is applied to it with shared-memory compare-and-swap, and the slice
//                       is copied back.  No global atomics, no random DRAM access.
// Saturating increments commute, so the table is bit-identical to the direct form's.  Middle bits of a canonical
// hash are uniform, so streams and leaves fill evenly; whatever does not fit a bucket or a stream region
// (low-complexity input) is applied directly with the global compare-and-swap -- exact either way.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void bump_direct(uint32_t* count, uint32_t h, const HashP& hp) {
    uint32_t g = tbl_index(h, hp);
    bump(count, g, ld_table(count + (g >> 4)));
}

template <int E>
__global__ void __launch_bounds__(kBinWarps * 32, kBinCtas) s1_bin_kernel(
    const uint8_t* __restrict__ fq, const uint64_t* __restrict__ rec_start, const uint64_t* __restrict__ rec_end,
    uint64_t rec_lo, uint64_t rec_hi, uint64_t budget, const uint32_t* __restrict__ sample_bits, uint64_t ordinal_base,
    HashP hp, BinP bp, uint32_t* __restrict__ count, unsigned long long* __restrict__ n_sampled, int* __restrict__ err) {
    extern __shared__ __align__(16) uint32_t dyn[];           // two bucket sets: one fills while the other drains
    __shared__ uint32_t cnt[2][1 << kMaxB1];
    __shared__ uint8_t lut[256];
    __shared__ __align__(16) uint8_t stage[kBinWarps][2][kStageBytes];
    __shared__ __align__(8) uint64_t sbar[kBinWarps][2];      // "slot filled", one per stage slot
    const int e = E ? E : hp.e;
    const int nbins = 1 << bp.b1;
    const uint32_t bin_mask = (uint32_t)nbins - 1u;
    const int bin_lo = hp.leaf_lo;
    const uint32_t bcap = bp.bcap;
    const uint32_t set_entries = bcap << bp.b1;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < (1 << kMaxB1)) { cnt[0][threadIdx.x] = 0; cnt[1][threadIdx.x] = 0; }
    if (lane == 0) { mbar_init(&sbar[warp][0], 1); mbar_init(&sbar[warp][1], 1); }
    fill_base_lut(lut);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // Draining: this warp owns streams warp, warp + 8, ...; they are drained side by side, a group of `lpb` lanes each
    // (64 streams: 8 per warp, 4 lanes each), so no lane idles on a short run and there is no per-stream loop.
    const int per_warp = (nbins + kBinWarps - 1) / kBinWarps;             // 1, 2, 4 or 8
    const int lpb = 32 / per_warp;
    const int grp = lane / lpb, sub = lane - grp * lpb;
    const int my_bin = warp + kBinWarps * grp;
    const bool owner = my_bin < nbins;
    unsigned long long mine = 0;
    const uint64_t stride = (uint64_t)gridDim.x * kBinWarps;
    auto next_sampled = [&](uint64_t q) {                     // warp-uniform
        while (q < rec_hi && !is_sampled(sample_bits, q + ordinal_base)) q += stride;
        return q;
    };
    // Two records ahead: the offsets of the record after next are in flight while the bytes of the next record are
    // being staged (both issued when the current read was accepted), so a warp never sits through the
    // start -> end -> bytes load chain between reads.
    uint64_t rn = next_sampled(rec_lo + (uint64_t)blockIdx.x * kBinWarps + warp);   // next record: offsets known (ns, ne)
    uint64_t ns = 0, ne = 0;
    if (rn < rec_hi) { ns = rec_start[rn]; ne = rec_end[rn]; }
    uint64_t rnn = rn < rec_hi ? next_sampled(rn + stride) : rn;   // the one after: offsets loading (nns, nne)
    uint64_t nns = 0, nne = 0;
    if (rnn < rec_hi) { nns = rec_start[rnn]; nne = rec_end[rnn]; }
    bool nstaged = false;                                     // next record's bytes are on their way into stage[warp][slot ^ 1]
    uint32_t nmis = 0, phase = 0;                             // phase bit s: parity the next fill of slot s completes
    int slot = 0;
    auto stageable = [&](uint64_t s0, uint64_t e0) { return s0 <= budget && e0 - s0 <= (uint64_t)kMaxReadLen && e0 - s0 >= (uint64_t)hp.k; };
    if (rn < rec_hi && stageable(ns, ne)) { nmis = stage_read(stage[warp][1], &sbar[warp][1], fq, ns, (uint32_t)(ne - ns), lane); nstaged = true; }
    const uint8_t* src = stage[warp][0];
    int len = 0, w = 0, nch = 0;                              // current read: length, next word to pack, hash chunks
    uint32_t chn = 0;                                         // this lane's byte of word w, read one chunk ahead
    Planes prev{0, 0, 0, 0};
    bool have = false;
    int set = 0;                                              // the set this round fills; set^1 drains
    for (bool first_round = true;; first_round = false) {
        while (!have && rn < rec_hi) {                        // turn the next record into the current read
            uint64_t start = ns, len64 = ne - ns;
            bool staged = nstaged;
            uint32_t mis = nmis;
            nstaged = false;
            bool too_long = len64 > (uint64_t)kMaxReadLen;
            bool counted = start <= budget && !too_long;      // Q15
            bool accept = counted && len64 >= (uint64_t)hp.k;
            if (accept) {
                slot ^= 1;                                    // the slot the bytes were (or now are) staged into
                if (!staged) mis = stage_read(stage[warp][slot], &sbar[warp][slot], fq, start, (uint32_t)len64, lane);   // only after a skipped record
                mbar_wait(&sbar[warp][slot], (phase >> slot) & 1u);
                phase ^= 1u << slot;
            }
            rn = rnn; ns = nns; ne = nne;
            if (rnn < rec_hi) {
                rnn = next_sampled(rnn + stride);
                if (rnn < rec_hi) { nns = rec_start[rnn]; nne = rec_end[rnn]; }
            }
            if (start <= budget && too_long && lane == 0) atomicExch(err, 1);
            mine += counted;
            if (!accept) continue;
            len = (int)len64;
            int np = len - hp.k + 1;
            src = stage[warp][slot] + mis;
            __syncwarp();                                     // every lane is done with the other slot
            if (rn < rec_hi && stageable(ns, ne)) { nmis = stage_read(stage[warp][slot ^ 1], &sbar[warp][slot ^ 1], fq, ns, (uint32_t)(ne - ns), lane); nstaged = true; }
            nch = (np + 31) >> 5;
            prev = pack_word(lane < len ? src[lane] : 0u, lut);
            chn = 32 + lane < len ? src[32 + lane] : 0u;
            w = 1;
            have = true;
        }
        // One barrier per round: behind it the set filled last round is complete and the set drained last round is free.
        bool any = __syncthreads_or(have);
        // reserve room in the global streams for what last round produced; the answers are picked up after this
        // round's hashing, so the atomics' round trip costs nothing
        // Streams only ever grow by whole 32-byte sectors (8 hashes): a bucket hands over the multiple of 8 it holds and
        // carries the rest into its next fill, so every store below is a full, aligned sector.
        uint32_t dn = 0, n8 = 0, dg = 0;
        if (!first_round && owner) {
            dn = min(cnt[set ^ 1][my_bin], bcap);
            n8 = dn & ~7u;
            if (n8 && sub == 0) dg = atomicAdd(bp.cursor_a + my_bin * kCursorStride, n8);
        }
        if (any) {
            uint32_t* buckets = dyn + set * set_entries;
            uint32_t* fill = cnt[set];
#pragma unroll 1
            for (int it = 0; it < bp.round_chunks && have; ++it) {  // chunk w-1 = words w-1 (prev) and w (cur)
                Planes cur = pack_word(chn, lut);
                int pn = (w + 1) * 32 + lane;
                chn = pn < len ? src[pn] : 0u;
                LeWin kw = le_window(prev, cur, lane, hp);
                if (kw.valid) {
                    uint32_t h[E ? E : kMaxE], at[E ? E : kMaxE];
#pragma unroll
                    for (int i = 0; i < (E ? E : kMaxE); ++i)
                        if (i < e) { h[i] = le_hash(kw, hp, i); at[i] = atomicAdd(&fill[(h[i] >> bin_lo) & bin_mask], 1u); }
#pragma unroll
                    for (int i = 0; i < (E ? E : kMaxE); ++i)
                        if (i < e) {
                            if (at[i] < bcap) buckets[((h[i] >> bin_lo) & bin_mask) * bcap + at[i]] = h[i];
                            else bump_direct(count, h[i], hp);   // bucket full: rare, exact either way
                        }
                }
                prev = cur;
                if (++w > nch) { have = false; }
            }
        }
        if (!first_round) {                                   // drain last round's set: all streams of this warp side by side
            dg = __shfl_sync(kFull, dg, grp * lpb);
            uint32_t* from = dyn + (set ^ 1) * set_entries + (uint32_t)(owner ? my_bin : 0) * bcap;
            if (owner && n8) {
                uint32_t fit = dg < bp.cap_a ? min(n8, bp.cap_a - dg) : 0u;            // a multiple of 8, like dg and cap_a
                const uint2* from2 = reinterpret_cast<const uint2*>(from);
                uint2* dst2 = reinterpret_cast<uint2*>(bp.pool_a + (size_t)my_bin * bp.cap_a + dg);
                for (uint32_t x = sub; x < fit / 2; x += lpb) dst2[x] = from2[x];
                for (uint32_t x = fit + sub; x < n8; x += lpb) bump_direct(count, from[x], hp);   // stream region full
            }
            __syncwarp();
            if (owner && n8)
                for (uint32_t j = sub; j < dn - n8; j += lpb) from[j] = from[n8 + j];  // carry (disjoint: n8 >= 8 > dn - n8)
            __syncwarp();
            if (owner && sub == 0) cnt[set ^ 1][my_bin] = dn - n8;
        }
        if (!any) break;
        set ^= 1;
    }
    // what the buckets still carry (< 8 hashes each) goes to the table directly
    __syncthreads();
    if (owner)
        for (int s2 = 0; s2 < 2; ++s2) {
            const uint32_t* from = dyn + s2 * set_entries + (uint32_t)my_bin * bcap;
            uint32_t left = min(cnt[s2][my_bin], bcap);
            for (uint32_t j = sub; j < left; j += lpb) bump_direct(count, from[j], hp);
        }
    if (lane == 0 && mine) atomicAdd(n_sampled, mine);
}

// P2.  Grid (tiles, 2^b1): CTA (x, y) splits tile x of stream y by bits [b1, b1 + b2) of the hash.
// (the LHGT_* macros exist for tools/sweep.sh: variants built with -D and timed side by side)
#ifndef LHGT_SPLIT_THREADS
#define LHGT_SPLIT_THREADS 512
#endif
#ifndef LHGT_SPLIT_PER
#define LHGT_SPLIT_PER 16
#endif
#ifndef LHGT_SPLIT_CTAS
#define LHGT_SPLIT_CTAS 3
#endif
constexpr int kSplitThreads = LHGT_SPLIT_THREADS, kSplitPer = LHGT_SPLIT_PER, kSplitTile = kSplitThreads * kSplitPer;   // 8192 hashes = 32 KiB

__global__ void __launch_bounds__(kSplitThreads, LHGT_SPLIT_CTAS) s1_split_kernel(BinP bp, HashP hp, uint32_t* __restrict__ count) {
    __shared__ uint32_t tile[kSplitTile];                      // the tile again, grouped by sub-stream
    __shared__ uint32_t hist[1 << kMaxB2], off[1 << kMaxB2];
    __shared__ uint2 route[1 << kMaxB2];                       // per sub-stream: {pool_b index of grouped position 0, first position that does not fit}
    __shared__ uint32_t wsum[8];
    const int nsub = 1 << bp.b2;
    const uint32_t sub_mask = (uint32_t)nsub - 1u;
    const int sub_lo = hp.leaf_lo + bp.b1;
    const uint32_t y = blockIdx.y;
    const uint32_t n = min(bp.cursor_a[y * kCursorStride], bp.cap_a);
    const uint32_t first = blockIdx.x * kSplitTile;
    if (first >= n) return;
    const uint32_t valid = min((uint32_t)kSplitTile, n - first);
    const uint32_t* __restrict__ in = bp.pool_a + (size_t)y * bp.cap_a + first;
    if (threadIdx.x < nsub) hist[threadIdx.x] = 0;
    uint32_t h[kSplitPer], rank[kSplitPer / 2];                // two 16-bit ranks per register
#pragma unroll
    for (int q = 0; q < kSplitPer; ++q) {
        uint32_t x = threadIdx.x + q * kSplitThreads;
        h[q] = x < valid ? ld_stream(in + x) : 0u;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kSplitPer; ++q) {
        uint32_t x = threadIdx.x + q * kSplitThreads;
        uint32_t r = x < valid ? atomicAdd(&hist[(h[q] >> sub_lo) & sub_mask], 1u) : 0u;
        if (q & 1) rank[q >> 1] |= r << 16; else rank[q >> 1] = r;
    }
    __syncthreads();
    // exclusive scan of the histogram (<= 256 entries: threads 0..255, one each) + one reservation per leaf
    {
        uint32_t v = threadIdx.x < nsub ? hist[threadIdx.x] : 0u;
        if (threadIdx.x < (1 << kMaxB2)) {
            int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
            uint32_t inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(kFull, inc, d);
                if (lane >= d) inc += t;
            }
            if (lane == 31) wsum[wp] = inc;
            asm volatile("bar.sync 1, 256;");                  // the first eight warps only
            uint32_t before = 0;
            for (int q = 0; q < wp; ++q) before += wsum[q];
            if (threadIdx.x < nsub) {
                uint32_t o = before + inc - v;
                off[threadIdx.x] = o;
                uint32_t leaf = y | ((uint32_t)threadIdx.x << bp.b1);
                uint32_t g = v ? atomicAdd(bp.cursor_b + leaf, v) : 0u;
                // grouped position p of this sub-stream lands at pool_b[leaf * cap_b + g + (p - o)] while g + (p - o) < cap_b
                route[threadIdx.x] = make_uint2(leaf * bp.cap_b + g - o, g < bp.cap_b ? o + (bp.cap_b - g) : o);
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kSplitPer; ++q) {
        uint32_t x = threadIdx.x + q * kSplitThreads;
        uint32_t r = (q & 1) ? rank[q >> 1] >> 16 : rank[q >> 1] & 0xffffu;
        if (x < valid) tile[off[(h[q] >> sub_lo) & sub_mask] + r] = h[q];
    }
    __syncthreads();
    // neighbours in the grouped tile are neighbours in their leaf stream: coalesced runs
    for (uint32_t p = threadIdx.x; p < valid; p += kSplitThreads) {
        uint32_t hh = tile[p];
        uint2 rt = route[(hh >> sub_lo) & sub_mask];
        if (p < rt.y) bp.pool_b[rt.x + p] = tbl_idx(hh, hp);      // the leaf is implied by the stream: keep the counter index only
        else bump_direct(count, hh, hp);                       // leaf region full
    }
}

// P3.  One CTA per leaf.
#ifndef LHGT_LEAF_THREADS
#define LHGT_LEAF_THREADS 640          // profiles/r01w_sweep.txt: 256x2 6.7 ms, 384 5.4, 512 5.0, 640 4.8, 1024 (2 CTAs) 5.3 per step
#endif
#ifndef LHGT_LEAF_PER
#define LHGT_LEAF_PER 1
#endif
#ifndef LHGT_LEAF_CTAS
#define LHGT_LEAF_CTAS 3
#endif
constexpr int kLeafThreads = LHGT_LEAF_THREADS, kLeafPer = LHGT_LEAF_PER;   // 16-byte loads per thread per batch
constexpr int kLeafMaxLog2 = 18;                               // 2^18 counters = 64 KiB of shared memory

__global__ void __launch_bounds__(kLeafThreads, LHGT_LEAF_CTAS) s1_leaf_kernel(BinP bp, HashP hp, uint32_t* __restrict__ count) {
    extern __shared__ __align__(128) uint32_t slice[];         // 2^(idx_bits - 4) words
    __shared__ __align__(8) uint64_t bar;
    const uint32_t leaf = blockIdx.x;
    const uint32_t n = min(bp.cursor_b[leaf], bp.cap_b);
    if (n == 0) return;
    const uint32_t words = 1u << (hp.idx_bits - 4);
    uint32_t* __restrict__ home = count + (size_t)leaf * words;
    const uint32_t* __restrict__ in = bp.pool_b + (size_t)leaf * bp.cap_b;
    const bool bulk = words >= 4;                              // bulk copies move multiples of 16 bytes
    if (bulk) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_expect_tx(&bar, words * 4u);
            for (uint32_t o = 0; o < words; o += 4096u)        // 16 KiB pieces
                bulk_load(slice + o, home + o, min(4096u, words - o) * 4u, &bar);
        }
    } else {
        for (uint32_t x = threadIdx.x; x < words; x += kLeafThreads) slice[x] = home[x];
    }
    // The stream is read four entries per load (leaf regions start on 16-byte boundaries), the next batch in flight
    // while this one is applied; the first loads travel while the slice does.
    const uint4* __restrict__ in4 = reinterpret_cast<const uint4*>(in);
    const uint32_t nvec = n >> 2;
    auto fetch = [&](uint32_t v) {
        uint4 r = make_uint4(0u, 0u, 0u, 0u);
        if (v < nvec) asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(in4 + v));
        return r;
    };
    auto apply = [&](uint32_t idx) {                           // idx = tbl_idx(hash), stored by s1_split_kernel
        uint32_t* addr = slice + (idx >> 4);
        int sh = (idx & 15u) * 2;
        uint32_t seen = *addr;
        while (((seen >> sh) & 3u) < 3u) {
            uint32_t old = atomicCAS(addr, seen, seen + (1u << sh));
            if (old == seen) break;
            seen = old;
        }
    };
    uint4 cur[kLeafPer];
#pragma unroll
    for (int q = 0; q < kLeafPer; ++q) cur[q] = fetch(threadIdx.x + q * kLeafThreads);
    __syncthreads();
    if (bulk) mbar_wait(&bar, 0);
    for (uint32_t base = 0; base < nvec; base += kLeafThreads * kLeafPer) {
        uint4 nx[kLeafPer];
#pragma unroll
        for (int q = 0; q < kLeafPer; ++q) nx[q] = fetch(base + kLeafThreads * kLeafPer + threadIdx.x + q * kLeafThreads);
#pragma unroll
        for (int q = 0; q < kLeafPer; ++q)
            if (base + threadIdx.x + q * kLeafThreads < nvec) { apply(cur[q].x); apply(cur[q].y); apply(cur[q].z); apply(cur[q].w); }
#pragma unroll
        for (int q = 0; q < kLeafPer; ++q) cur[q] = nx[q];
    }
    for (uint32_t x = (nvec << 2) + threadIdx.x; x < n; x += kLeafThreads) apply(in[x]);   // n % 4 entries
    __syncthreads();
    if (bulk) {
        if (threadIdx.x == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (uint32_t o = 0; o < words; o += 4096u) bulk_store(home + o, slice + o, min(4096u, words - o) * 4u);
            bulk_store_commit_wait();
        }
    } else {
        for (uint32_t x = threadIdx.x; x < words; x += kLeafThreads) home[x] = slice[x];
    }
}

size_t s1_bin_smem_bytes(const BinP& bp) { return (size_t)2 * ((size_t)bp.bcap << bp.b1) * sizeof(uint32_t); }   // two bucket sets
int s1_leaf_max_log2() { return kLeafMaxLog2; }

template <int E>
static cudaError_t s1_bin_launch(const uint8_t* fq, const uint64_t* rs, const uint64_t* re, uint64_t lo, uint64_t hi, uint64_t budget,
                                 const uint32_t* sb, uint64_t ob, const HashP& hp, const BinP& bp, uint32_t* count,
                                 unsigned long long* ns, int* err, cudaStream_t st) {
    size_t smem = s1_bin_smem_bytes(bp);
    cudaError_t rc = cudaFuncSetAttribute(s1_bin_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    uint64_t want = (hi - lo + kBinWarps - 1) / kBinWarps;
    unsigned grid = (unsigned)(want < (uint64_t)kSMs * kBinCtas ? want : (uint64_t)kSMs * kBinCtas);
    s1_bin_kernel<E><<<grid, kBinWarps * 32, smem, st>>>(fq, rs, re, lo, hi, budget, sb, ob, hp, bp, count, ns, err);
    return cudaGetLastError();
}

int launch_s1_binned(const uint8_t* fq, const uint64_t* rec_start, const uint64_t* rec_end, uint64_t rec_lo, uint64_t rec_hi,
                     uint64_t budget, const uint32_t* sample_bits, uint64_t ordinal_base, const HashP& hp, const BinP& bp,
                     uint32_t* count, unsigned long long* n_sampled, int* err, int phase, cudaStream_t st) {
    if (rec_hi <= rec_lo) return 0;
    if (phase == 1) {
        unsigned tiles = (bp.cap_a + kSplitTile - 1) / kSplitTile;
        s1_split_kernel<<<dim3(tiles, 1u << bp.b1), kSplitThreads, 0, st>>>(bp, hp, count);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    if (phase == 2) {
        size_t smem = (size_t)4 << (hp.idx_bits - 4);
        if (cudaFuncSetAttribute(s1_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        s1_leaf_kernel<<<1u << hp.leaf_bits, kLeafThreads, smem, st>>>(bp, hp, count);
        return cudaGetLastError() == cudaSuccess ? 1 : -1;
    }
    cudaError_t rc;
    switch (hp.e) {
        case 1: rc = s1_bin_launch<1>(fq, rec_start, rec_end, rec_lo, rec_hi, budget, sample_bits, ordinal_base, hp, bp, count, n_sampled, err, st); break;
        case 2: rc = s1_bin_launch<2>(fq, rec_start, rec_end, rec_lo, rec_hi, budget, sample_bits, ordinal_base, hp, bp, count, n_sampled, err, st); break;
        case 3: rc = s1_bin_launch<3>(fq, rec_start, rec_end, rec_lo, rec_hi, budget, sample_bits, ordinal_base, hp, bp, count, n_sampled, err, st); break;
        case 4: rc = s1_bin_launch<4>(fq, rec_start, rec_end, rec_lo, rec_hi, budget, sample_bits, ordinal_base, hp, bp, count, n_sampled, err, st); break;
        default: rc = s1_bin_launch<0>(fq, rec_start, rec_end, rec_lo, rec_hi, budget, sample_bits, ordinal_base, hp, bp, count, n_sampled, err, st); break;
    }
    return rc != cudaSuccess ? -1 : 1;
}

// ------------------------------------------------------------------------------------------------
// S2 (E:888-979 + E:550-725 + E:239-301) as data-parallel passes over 1024-position tiles: gather (direct form below, sliced
// form further down), mark, complete, windows (good / flag / count_new), ids (scan), register (direct or through buckets).
// Bit arrays are little-endian in bit order: tile t, local position x -> word t*32 + x/32, bit x%32.
// ------------------------------------------------------------------------------------------------

// pass a: trio = ALL e stored hashes of the position saturated (E:573-595; stored 0 = no hit, Q4, E:936-941), evaluated
// as a short-circuit AND: hash i+1 is only looked up where hashes 0..i were saturated.  A table probe is a 128-byte DRAM
// fill on this part (profiles/r01p), so the pass costs ~1.1 probes per position instead of e.  `sat0` (hash 0 saturated)
// is a lower bound of `single` (SOME hash saturated) and becomes exact in pass d, only where it can matter.
template <int E>
__global__ void __launch_bounds__(256) s2_trio_kernel(const uint32_t* __restrict__ image, const Contig* __restrict__ contigs,
                                                      const Tile* __restrict__ tiles, uint64_t tile_begin, HashP hp,
                                                      const uint32_t* __restrict__ count, uint32_t* __restrict__ single,
                                                      uint32_t* __restrict__ trio) {
    const int e = E ? E : hp.e;
    uint64_t tix = tile_begin + blockIdx.x;
    Tile t = tiles[tix];
    Contig c = contigs[t.contig];
    long np = (long)c.len - hp.k + 1;
    const uint32_t* hashes = image + c.hash_word;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t h[4];
    bool sat[4], first[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        long j = (long)t.j0 + r * 256 + threadIdx.x;
        h[r] = j < np ? ld_stream(hashes + (size_t)j * e) : 0u;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t g = tbl_index(h[r], hp);
        sat[r] = h[r] && ((ld_stream(count + (g >> 4)) >> ((g & 15u) * 2)) & 3u) == 3u;
        first[r] = sat[r];
    }
    for (int i = 1; i < e; ++i) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            long j = (long)t.j0 + r * 256 + threadIdx.x;
            h[r] = sat[r] ? ld_stream(hashes + (size_t)j * e + i) : 0u;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (sat[r]) {
                uint32_t g = tbl_index(h[r], hp);
                sat[r] = h[r] && ((ld_stream(count + (g >> 4)) >> ((g & 15u) * 2)) & 3u) == 3u;
            }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t ws = __ballot_sync(kFull, first[r]);
        uint32_t wt = __ballot_sync(kFull, sat[r]);
        if (lane == 0) {
            size_t word = (size_t)tix * kTileWords + r * 8 + warp;
            single[word] = ws;
            trio[word] = wt;
        }
    }
}

// pass b: a tile is HOT when some position j of it has three[j] = sum trio[j-499..j] >= three_min -- the only positions
// that can be good windows (E:610).  One warp per tile: lane l holds word l of the tile and of its predecessor.
__global__ void __launch_bounds__(256) s2_hot_kernel(const Contig* __restrict__ contigs, const Tile* __restrict__ tiles, uint64_t ntiles,
                                                     const uint32_t* __restrict__ trio, int three_min, uint8_t* __restrict__ hot) {
    int lane = threadIdx.x & 31;
    uint64_t tix = (uint64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (tix >= ntiles) return;
    Tile t = tiles[tix];
    uint32_t len = contigs[t.contig].len;
    uint32_t cur = trio[tix * kTileWords + lane];
    uint32_t prev = t.j0 > 0 ? trio[(tix - 1) * kTileWords + lane] : 0u;
    int pc = __popc(cur), pp = __popc(prev);
    int ic = pc, ip = pp;                                      // inclusive scans
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int a = __shfl_up_sync(kFull, ic, d), b = __shfl_up_sync(kFull, ip, d);
        if (lane >= d) { ic += a; ip += b; }
    }
    int tot_p = __shfl_sync(kFull, ip, 31), tot_c = __shfl_sync(kFull, ic, 31);
    bool is_hot = false;
    if (three_min <= 0) is_hot = true;
    else if (tot_p + tot_c >= three_min) {
        int cum_c = tot_p + ic - pc, cum_p = ip - pp;          // set bits before this lane's word, counted from the start of prev
        // position x = 32 * lane + bit of the tile sits at b = 1024 + x of the two-tile window; b - 500 = 32 * (16 + lane) + 12 + bit
        int sa = (lane + 16) & 31, sb = (lane + 17) & 31;
        uint32_t pa = __shfl_sync(kFull, prev, sa), ca = __shfl_sync(kFull, cur, sa);
        uint32_t pb = __shfl_sync(kFull, prev, sb), cb = __shfl_sync(kFull, cur, sb);
        int cpa = __shfl_sync(kFull, cum_p, sa), cca = __shfl_sync(kFull, cum_c, sa);
        int cpb = __shfl_sync(kFull, cum_p, sb), ccb = __shfl_sync(kFull, cum_c, sb);
        uint32_t wa = lane < 16 ? pa : ca, wb = lane < 15 ? pb : cb;
        int ca_ = lane < 16 ? cpa : cca, cb_ = lane < 15 ? cpb : ccb;
        long j = (long)t.j0 + 32 * lane;
#pragma unroll 4
        for (int bit = 0; bit < 32; ++bit) {
            int hi = cum_c + __popc(cur & (0xffffffffu >> (31 - bit)));
            int o = 12 + bit;
            int lo = o < 32 ? ca_ + __popc(wa & (0xffffffffu >> (31 - o))) : cb_ + __popc(wb & (0xffffffffu >> (63 - o)));
            is_hot |= j + bit < (long)len && hi - lo >= three_min;
        }
    }
    uint32_t any = __ballot_sync(kFull, is_hot);
    if (lane == 0) hot[tix] = any != 0u;
}

// pass c: the tiles the remaining passes have to look at = those within two tiles of a hot one on the same contig: a good
// window reaches 499 positions back for `one` (E:597-608), an interval 1000 positions either side of a good window
// (E:617-638) and the coverage-edge test 2k+9 further (E:640-671).  Everywhere else good = flagged = 0 whatever `single`
// holds.  The list is compacted IN TILE ORDER (count per 1024 tiles, scan, write): every rank of a multi-GPU run holds the
// same list and takes an equal share of it, which is a contiguous range of tiles.
__device__ __forceinline__ bool tile_needed(const Tile* __restrict__ tiles, uint64_t ntiles, const uint8_t* __restrict__ hot, uint64_t tix) {
    if (tix >= ntiles) return false;
    uint32_t contig = tiles[tix].contig;
    for (long u = (long)tix - 2; u <= (long)tix + 2; ++u)
        if (u >= 0 && u < (long)ntiles && hot[u] && tiles[u].contig == contig) return true;
    return false;
}

__global__ void __launch_bounds__(256) s2_need_count_kernel(const Tile* __restrict__ tiles, uint64_t ntiles, const uint8_t* __restrict__ hot,
                                                            uint32_t* __restrict__ block_cnt) {
    __shared__ uint32_t sm[33];
    uint64_t base = (uint64_t)blockIdx.x * 1024 + threadIdx.x * 4;
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) c += tile_needed(tiles, ntiles, hot, base + q);
    uint32_t total;
    block_exclusive_scan(c, &total, sm);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) s2_need_write_kernel(const Tile* __restrict__ tiles, uint64_t ntiles, const uint8_t* __restrict__ hot,
                                                            const uint32_t* __restrict__ block_cnt, const uint32_t* __restrict__ block_base,
                                                            uint32_t* __restrict__ need_list, uint32_t* __restrict__ n_need) {
    __shared__ uint32_t sm[33];
    uint64_t base = (uint64_t)blockIdx.x * 1024 + threadIdx.x * 4;
    bool need[4];
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { need[q] = tile_needed(tiles, ntiles, hot, base + q); c += need[q]; }
    uint32_t total, at = block_exclusive_scan(c, &total, sm) + block_base[blockIdx.x];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (need[q]) need_list[at++] = (uint32_t)(base + q);
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) *n_need = block_base[blockIdx.x] + block_cnt[blockIdx.x];
}

// pass d: `single` made exact on the needed tiles of [tile_begin, tile_end): a short-circuit OR over hashes 1..e-1 for the
// positions whose hash 0 was not saturated.
template <int E>
__global__ void __launch_bounds__(256) s2_single_kernel(const uint32_t* __restrict__ image, const Contig* __restrict__ contigs,
                                                        const Tile* __restrict__ tiles, const uint32_t* __restrict__ need_list,
                                                        const uint32_t* __restrict__ n_need, uint64_t tile_begin, uint64_t tile_end,
                                                        HashP hp, const uint32_t* __restrict__ count, uint32_t* __restrict__ single) {
    const int e = E ? E : hp.e;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = *n_need;
    for (uint32_t it = blockIdx.x; it < n; it += gridDim.x) {
        uint64_t tix = need_list[it];
        if (tix < tile_begin || tix >= tile_end) continue;
        Tile t = tiles[tix];
        Contig c = contigs[t.contig];
        long np = (long)c.len - hp.k + 1;
        const uint32_t* hashes = image + c.hash_word;
        bool open[4], sat[4];
        uint32_t h[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int xl = r * 256 + threadIdx.x;
            long j = (long)t.j0 + xl;
            uint32_t w = single[(size_t)tix * kTileWords + (xl >> 5)];
            sat[r] = (w >> (xl & 31)) & 1u;
            open[r] = !sat[r] && j < np;
        }
        for (int i = 1; i < e; ++i) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                long j = (long)t.j0 + r * 256 + threadIdx.x;
                h[r] = open[r] ? ld_stream(hashes + (size_t)j * e + i) : 0u;
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (open[r] && h[r]) {
                    uint32_t g = tbl_index(h[r], hp);
                    if (((ld_stream(count + (g >> 4)) >> ((g & 15u) * 2)) & 3u) == 3u) { sat[r] = true; open[r] = false; }
                }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            uint32_t ws = __ballot_sync(kFull, sat[r]);
            if (lane == 0) single[(size_t)tix * kTileWords + r * 8 + warp] = ws;
        }
    }
}

// inclusive count of set bits in local bit positions [0, x] of `words` given per-word exclusive prefix `cum`
__device__ __forceinline__ int bits_upto(const uint32_t* words, const int* cum, int x) {
    if (x < 0) return 0;
    int q = x >> 5;
    return cum[q] + __popc(words[q] & (0xffffffffu >> (31 - (x & 31))));
}

// cum[i] = set bits in words[0..i), cum[n] = total; n <= 96.  Call with all threads: warp `w` does the work (a shuffle scan).
__device__ __forceinline__ void prefix_words(const uint32_t* words, int* cum, int n, int w = 0) {
    if ((int)(threadIdx.x >> 5) != w) return;
    int lane = threadIdx.x & 31;
    int c[3], sum = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int idx = lane * 3 + q;
        c[q] = idx < n ? __popc(words[idx]) : 0;
        sum += c[q];
    }
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(kFull, inc, d);
        if (lane >= d) inc += t;
    }
    int ex = inc - sum;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        int idx = lane * 3 + q;
        if (idx < n) cum[idx] = ex;
        ex += c[q];
    }
    if (lane == 31) cum[n] = inc;
}

// pass e: 500-wide window sums and the good-window flag (E:597-615), on the needed tiles
__global__ void __launch_bounds__(256) s2_good_kernel(const Contig* __restrict__ contigs, const Tile* __restrict__ tiles,
                                                      const uint32_t* __restrict__ need_list, const uint32_t* __restrict__ n_need,
                                                      uint64_t t_lo, uint64_t t_hi,
                                                      const uint32_t* __restrict__ single, const uint32_t* __restrict__ trio,
                                                      int one_min, int three_min, uint32_t* __restrict__ good) {
    __shared__ uint32_t ws[2 * kTileWords], wt[2 * kTileWords];
    __shared__ int cs[2 * kTileWords + 1], ct[2 * kTileWords + 1];
    const uint32_t n = *n_need;
    for (uint32_t it = blockIdx.x; it < n; it += gridDim.x) {
        const uint64_t tix = need_list[it];
        if (tix < t_lo || tix >= t_hi) continue;
        Tile t = tiles[tix];
        Contig c = contigs[t.contig];
        bool has_prev = t.j0 > 0;
        if (threadIdx.x < 2 * kTileWords) {
            int q = threadIdx.x;
            bool take = q >= kTileWords || has_prev;
            size_t word = (size_t)tix * kTileWords + q - kTileWords;
            ws[q] = take ? single[word] : 0u;
            wt[q] = take ? trio[word] : 0u;
        }
        __syncthreads();
        prefix_words(ws, cs, 2 * kTileWords, 0);
        prefix_words(wt, ct, 2 * kTileWords, 1);
        __syncthreads();
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int xl = r * 256 + threadIdx.x;
            long j = (long)t.j0 + xl;
            int b = kTile + xl;
            int one = bits_upto(ws, cs, b) - bits_upto(ws, cs, b - 500);
            int three = bits_upto(wt, ct, b) - bits_upto(wt, ct, b - 500);
            bool g = j < (long)c.len && one >= one_min && three >= three_min;
            uint32_t wg = __ballot_sync(kFull, g);
            if (lane == 0) good[(size_t)tix * kTileWords + r * 8 + warp] = wg;
        }
        __syncthreads();
    }
}

// min of the unsigned bytes / max of the signed bytes a[lo .. lo + n), 1 <= n <= 32, `a` 4-byte aligned in shared memory and
// readable up to lo + 35: nine word loads re-aligned with funnel shifts, then byte-SIMD min/max -- instead of n byte loads.
__device__ __forceinline__ uint32_t window_min_u8(const signed char* a, int lo, int n) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a) + (lo >> 2);
    const int sh = (lo & 3) * 8;
    uint32_t acc = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t v = __funnelshift_r(w[j], w[j + 1], sh);             // bytes lo + 4j .. lo + 4j + 3
        int left = n - 4 * j;                                        // how many of them belong to the window
        if (left <= 0) break;
        if (left < 4) v |= 0xffffffffu << (8 * left);
        acc = __vminu4(acc, v);
    }
    acc = __vminu4(acc, acc >> 16);
    acc = __vminu4(acc, acc >> 8);
    return acc & 0xffu;
}
__device__ __forceinline__ int window_max_s8(const signed char* a, int lo, int n) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(a) + (lo >> 2);
    const int sh = (lo & 3) * 8;
    uint32_t acc = 0x80808080u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        uint32_t v = __funnelshift_r(w[j], w[j + 1], sh);
        int left = n - 4 * j;
        if (left <= 0) break;
        if (left < 4) v = (v & ~(0xffffffffu << (8 * left))) | (0x80808080u << (8 * left));
        acc = __vmaxs4(acc, v);
    }
    acc = __vmaxs4(acc, acc >> 16);
    acc = __vmaxs4(acc, acc >> 8);
    return (int)(signed char)(acc & 0xffu);
}

// pass c: interval membership (E:617-638, 675-686 collapse to "a good window within +-1000") and the
// coverage-edge peak test (E:640-671) in its closed form:
//   D(x) = sum single[x-4..x],  C(j) = D(j-5) - D(j-k-5) - D(j),  diff_t(j) = C(j) + D(j-k-5-t), t in [0,k)
//   peak[j] if some diff_t(j) <= -2;  peak[j-k-5-t] if diff_t(j) >= 2;  only for 2k+10 < j < len.
__global__ void __launch_bounds__(256) s2_flag_kernel(const Contig* __restrict__ contigs, const Tile* __restrict__ tiles,
                                                      uint64_t ntiles, const uint32_t* __restrict__ need_list,
                                                      const uint32_t* __restrict__ n_need, uint64_t t_lo, uint64_t t_hi, int k,
                                                      const uint32_t* __restrict__ single,
                                                      const uint32_t* __restrict__ good, uint32_t* __restrict__ flagged) {
    __shared__ uint32_t ws[3 * kTileWords + 1], wg[3 * kTileWords];
    __shared__ int cg[3 * kTileWords + 1];
    __shared__ __align__(4) signed char D[3 * kTile + 8], C[3 * kTile + 8];
    const uint32_t n = *n_need;
    for (uint32_t it = blockIdx.x; it < n; it += gridDim.x) {
        const uint64_t tix = need_list[it];
        if (tix < t_lo || tix >= t_hi) continue;
        Tile t = tiles[tix];
        Contig c = contigs[t.contig];
        bool has_prev = t.j0 > 0;
        bool has_next = tix + 1 < ntiles && tiles[tix + 1].contig == t.contig;
        if (threadIdx.x < 3 * kTileWords) {
            int q = threadIdx.x;
            bool take = (q >= kTileWords || has_prev) && (q < 2 * kTileWords || has_next);
            size_t word = (size_t)tix * kTileWords + q - kTileWords;
            ws[q] = take ? single[word] : 0u;
            wg[q] = take ? good[word] : 0u;
        }
        if (threadIdx.x == 0) ws[3 * kTileWords] = 0;
        __syncthreads();
        prefix_words(wg, cg, 3 * kTileWords);
        __syncthreads();
        int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (cg[3 * kTileWords] == 0) {                          // no good window within reach: nothing is flagged here
            if (threadIdx.x < kTileWords) flagged[(size_t)tix * kTileWords + threadIdx.x] = 0u;
            __syncthreads();
            continue;
        }
        for (int x = threadIdx.x; x < 3 * kTile; x += 256) {
            int d = 0;
            if (x >= 4) {
                int lo = x - 4, q = lo >> 5, s = lo & 31;
                uint64_t two = ((uint64_t)ws[q + 1] << 32) | ws[q];
                d = __popc((uint32_t)(two >> s) & 31u);
            }
            D[x] = (signed char)d;
        }
        __syncthreads();
        for (int x = kTile + threadIdx.x; x < 3 * kTile; x += 256) {
            long j = (long)t.j0 - kTile + x;
            bool okj = j > 2 * k + 10 && j < (long)c.len;
            C[x] = okj ? (signed char)(D[x - 5] - D[x - k - 5] - D[x]) : (signed char)-100;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int xl = r * 256 + threadIdx.x;
            int x = kTile + xl;
            long j = (long)t.j0 + xl;
            bool f = false;
            if (j < (long)c.len && j >= 1) {
                int lo = x - 1000, hi = x + 1000;
                bool in_iv = bits_upto(wg, cg, hi) - bits_upto(wg, cg, lo - 1) > 0;
                if (in_iv) {
                    int cj = C[x], dq = D[x];
                    bool pk = false;
                    if (cj != -100) pk = cj + (int)window_min_u8(D, x - 2 * k - 4, k) <= -2;        // min over D[x-k-5-t], t in [0, k)
                    if (!pk) pk = window_max_s8(C, x + k + 5, k) + dq >= 2;                          // some C[x+k+5+t] + D[x] >= 2
                    f = pk;
                }
            }
            uint32_t wf = __ballot_sync(kFull, f);
            if (lane == 0) flagged[(size_t)tix * kTileWords + r * 8 + warp] = wf;
        }
        __syncthreads();
    }
}

// A flagged position opens a new peak iff no flagged position precedes it in its 50-bp bucket (E:288-301).
__device__ __forceinline__ bool flagged_at(const uint32_t* __restrict__ flagged, uint64_t tix, long j0, long j) {
    long rel = j - j0;                 // may be negative (previous tile of the same contig)
    long bit = (long)tix * kTile + rel;
    return (flagged[bit >> 5] >> (bit & 31)) & 1u;
}

__device__ __forceinline__ bool opens_peak(const uint32_t* __restrict__ flagged, uint64_t tix, long j0, long j) {
    long bs = (j / 50) * 50;
    for (long q = bs; q < j; ++q)
        if (flagged_at(flagged, tix, j0, q)) return false;
    return true;
}

__global__ void __launch_bounds__(256) s2_count_new_kernel(const Tile* __restrict__ tiles, const uint32_t* __restrict__ need_list,
                                                           const uint32_t* __restrict__ n_need, uint64_t t_lo, uint64_t t_hi,
                                                           const uint32_t* __restrict__ flagged,
                                                           uint32_t* __restrict__ tile_new, unsigned long long* __restrict__ flagged_total) {
    __shared__ uint32_t n_new, n_flag;
    const uint32_t n = *n_need;
    for (uint32_t it = blockIdx.x; it < n; it += gridDim.x) {
        const uint64_t tix = need_list[it];
        if (tix < t_lo || tix >= t_hi) continue;
        if (threadIdx.x == 0) { n_new = 0; n_flag = 0; }
        __syncthreads();
        Tile t = tiles[tix];
        uint32_t mine = 0, mine_f = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int xl = r * 256 + threadIdx.x;
            uint32_t w = flagged[(size_t)tix * kTileWords + (xl >> 5)];
            if ((w >> (xl & 31)) & 1u) {
                ++mine_f;
                mine += opens_peak(flagged, tix, t.j0, (long)t.j0 + xl);
            }
        }
        if (mine) atomicAdd(&n_new, mine);
        if (mine_f) atomicAdd(&n_flag, mine_f);
        __syncthreads();
        if (threadIdx.x == 0) {
            tile_new[tix] = n_new;
            if (n_flag) atomicAdd(flagged_total, (unsigned long long)n_flag);
        }
    }
}

// Peak ids follow (contig, position) order = tile order: id = tile_base + (#openers at or before me) - 1.
// Every flagged position stamps its id on the k-mers it holds with count > 0; later ids win, i.e.
// peak_kmer[h] = max id (E:246-270 executed in order).  Id 0 is the reference's "none" (Q10).
template <int E>
__global__ void __launch_bounds__(256) s2_register_kernel(const uint32_t* __restrict__ image, const Contig* __restrict__ contigs,
                                                          const Tile* __restrict__ tiles, const uint32_t* __restrict__ need_list,
                                                          const uint32_t* __restrict__ n_need, uint64_t t_lo, uint64_t t_hi, HashP hp,
                                                          const uint32_t* __restrict__ count, const uint32_t* __restrict__ flagged,
                                                          const uint32_t* __restrict__ tile_base, int32_t* __restrict__ loci, uint32_t loci_cap,
                                                          uint32_t* __restrict__ peak_kmer, uint32_t* __restrict__ prefilter,
                                                          int mode) {
    __shared__ uint32_t opener[kTileWords];
    __shared__ int cum[kTileWords + 1];
    const int e = E ? E : hp.e;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t n = *n_need;
    for (uint32_t it = blockIdx.x; it < n; it += gridDim.x) {
        const uint64_t tix = need_list[it];
        if (tix < t_lo || tix >= t_hi) continue;
        Tile t = tiles[tix];
        Contig c = contigs[t.contig];
        bool fl[4], op[4];
        bool any = false;
        __syncthreads();                                        // the previous tile's readers of opener/cum are done
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int xl = r * 256 + threadIdx.x;
            uint32_t w = flagged[(size_t)tix * kTileWords + (xl >> 5)];
            fl[r] = (w >> (xl & 31)) & 1u;
            op[r] = fl[r] && opens_peak(flagged, tix, t.j0, (long)t.j0 + xl);
            uint32_t wo = __ballot_sync(kFull, op[r]);
            if (lane == 0) opener[r * 8 + warp] = wo;
            any |= fl[r];
        }
        if (!__syncthreads_or(any)) continue;
        prefix_words(opener, cum, kTileWords);
        __syncthreads();
        long np = (long)c.len - hp.k + 1;
        uint32_t base = tile_base[tix];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (!fl[r]) continue;
            int xl = r * 256 + threadIdx.x;
            long j = (long)t.j0 + xl;
            uint32_t id = base + (uint32_t)bits_upto(opener, cum, xl) - 1u;   // wraps to 0xffffffff only if no opener yet: impossible for a flagged bit
            if (op[r] && mode == 0 && id < loci_cap) { loci[2 * (size_t)id] = (int32_t)t.contig + 1; loci[2 * (size_t)id + 1] = (int32_t)j; }
            if (j < np && id != 0u) {                                         // j = len-k+1 reads the zero tail (Q6)
                const uint32_t* hashes = image + c.hash_word + (size_t)j * e;
                for (int i = 0; i < e; ++i) {
                    uint32_t h = hashes[i];
                    if (!h) continue;
                    uint32_t g = tbl_index(h, hp);
                    uint32_t cnt = (count[g >> 4] >> ((g & 15u) * 2)) & 3u;
                    if (!cnt) continue;                                        // E:250,265: hit > 0
                    uint32_t slot = prefilter_slot(g);
                    if (mode == 0) {
                        atomicMax(peak_kmer + g, id);                           // the peak table is laid out like the count table (tbl_index)
                        if (prefilter) atomicOr(prefilter + (slot >> 5), 1u << (slot & 31));
                    } else {
                        peak_kmer[g] = 0u;
                        prefilter[slot >> 5] = 0u;
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// S2 gather through table slices (DESIGN.md 4.5a').  A direct probe of a count table far larger than L2 is a 128-byte DRAM
// fill for 2 useful bits.  Here the stored hashes of a chunk of the reference are turned into (position, table index)
// records, appended to 128 buckets = 128 contiguous slices of the table (8 MiB each at k = 32: the two or three slices in
// flight, the record stream and the planes share L2 comfortably), and answered bucket by bucket while that slice sits in L2; a saturated counter sets the position's bit in the hash's own plane (records keep their
// position order inside a bucket, so those atomics stay in L2 too).  Everything that touches DRAM is then a stream: the
// image once, the records once out and once in, the table once per chunk.
// ------------------------------------------------------------------------------------------------
constexpr int kGsBuckets = 128, kGsLog2 = 7, kGsStage = 40, kGsCursorStride = 32;

struct GsSink { uint2* pool; uint32_t* cursor; uint32_t cap; };   // bucket b: pool[b * cap .. +cap), cursor[b * kGsCursorStride]

__device__ __forceinline__ void gs_answer(uint2 r, uint32_t bucket, int slice_shift, uint64_t chunk_bit0, size_t plane_words,
                                          const uint32_t* __restrict__ count, uint32_t* __restrict__ sat) {
    uint32_t g = (bucket << slice_shift) | (r.x & 0x0fffffffu), i = r.x >> 28;
    if (((ld_table(count + (g >> 4)) >> ((g & 15u) * 2)) & 3u) != 3u) return;
    uint64_t bit = chunk_bit0 + r.y;
    atomicOr(sat + (size_t)i * plane_words + (bit >> 5), 1u << (bit & 31));
}

template <int E>
__global__ void __launch_bounds__(256) s2_gsemit_kernel(const uint32_t* __restrict__ image, const Contig* __restrict__ contigs,
                                                        const Tile* __restrict__ tiles, uint64_t tile_begin, uint64_t chunk_tile0, HashP hp,
                                                        int slice_shift, size_t plane_words, const uint32_t* __restrict__ count,
                                                        uint32_t* __restrict__ sat, GsSink sink) {
    __shared__ uint2 stage[kGsBuckets][kGsStage];
    __shared__ uint32_t cnt[kGsBuckets];
    const int e = E ? E : hp.e;
    uint64_t tix = tile_begin + blockIdx.x;
    Tile t = tiles[tix];
    Contig c = contigs[t.contig];
    long np = (long)c.len - hp.k + 1;
    const uint32_t* hashes = image + c.hash_word;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kGsBuckets) cnt[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t chunk_bit0 = chunk_tile0 * kTile;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int xl = r * 256 + threadIdx.x;
        long j = (long)t.j0 + xl;
        if (j >= np) continue;
        uint32_t pos = (uint32_t)((tix - chunk_tile0) * kTile + xl);        // < 2^28: the caller bounds the chunk
#pragma unroll
        for (int i = 0; i < (E ? E : 4); ++i) {
            if (i >= e) break;
            uint32_t h = ld_stream(hashes + (size_t)j * e + i);
            if (!h) continue;                                                 // stored 0 = no hit (Q4, E:936-941)
            uint32_t g = tbl_index(h, hp), b = g >> slice_shift;
            uint2 rec = make_uint2((g & ((1u << slice_shift) - 1u)) | ((uint32_t)i << 28), pos);
            uint32_t slot = atomicAdd(&cnt[b], 1u);
            if (slot < (uint32_t)kGsStage) stage[b][slot] = rec;
            else {                                                            // stage full (a skewed tile): straight to the region
                uint32_t gq = atomicAdd(sink.cursor + b * kGsCursorStride, 1u);
                if (gq < sink.cap) sink.pool[(size_t)b * sink.cap + gq] = rec;
                else gs_answer(rec, b, slice_shift, chunk_bit0, plane_words, count, sat);
            }
        }
    }
    __syncthreads();
    // a warp hands its 16 buckets over: all reservations first (one atomic per lane, in flight together), then one coalesced
    // run per bucket
    static_assert(kGsBuckets == 128 && kGsStage <= 64, "flush layout");
    {
        const int b_mine = warp * 16 + (lane & 15);
        uint32_t n_mine = min(cnt[b_mine], (uint32_t)kGsStage), g_mine = 0;
        if (lane < 16 && n_mine) g_mine = atomicAdd(sink.cursor + b_mine * kGsCursorStride, n_mine);
#pragma unroll 1
        for (int u = 0; u < 16; ++u) {
            const int b = warp * 16 + u;
            uint32_t n = __shfl_sync(kFull, n_mine, u), g0 = __shfl_sync(kFull, g_mine, u);
            for (uint32_t q = lane; q < n; q += 32) {
                if (g0 + q < sink.cap) sink.pool[(size_t)b * sink.cap + g0 + q] = stage[b][q];
                else gs_answer(stage[b][q], b, slice_shift, chunk_bit0, plane_words, count, sat);
            }
        }
    }
}

// grid (parts, kGsBuckets), dispatched in index order: the resident CTAs share one or two slices of the table
__global__ void __launch_bounds__(256, 4) s2_gsapply_kernel(GsSink sink, int slice_shift, uint64_t chunk_bit0, size_t plane_words,
                                                            const uint32_t* __restrict__ count, uint32_t* __restrict__ sat) {
    const uint32_t b = blockIdx.y;
    const uint32_t n = min(sink.cursor[b * kGsCursorStride], sink.cap);
    const uint2* __restrict__ in = sink.pool + (size_t)b * sink.cap;
    for (uint32_t x = blockIdx.x * 256 + threadIdx.x; x < n; x += gridDim.x * 256) {
        uint2 r;
        asm volatile("ld.global.cs.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(in + x));   // evict-first: the table slice is what L2 is for
        gs_answer(r, b, slice_shift, chunk_bit0, plane_words, count, sat);
    }
}

// single = some plane set, trio = all e planes set, word by word over tiles [tile_begin, tile_end)
__global__ void s2_gscombine_kernel(const uint32_t* __restrict__ sat, size_t plane_words, int e, size_t w_lo, size_t w_hi,
                                    uint32_t* __restrict__ single, uint32_t* __restrict__ trio) {
    for (size_t w = w_lo + (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < w_hi; w += (size_t)gridDim.x * blockDim.x) {
        uint32_t any = 0u, all = 0xffffffffu;
        for (int i = 0; i < e; ++i) { uint32_t v = sat[(size_t)i * plane_words + w]; any |= v; all &= v; }
        single[w] = any; trio[w] = all;
    }
}

int s2_gs_buckets() { return kGsBuckets; }
int s2_gs_cursor_words() { return kGsBuckets * kGsCursorStride; }

// one chunk of tiles [tile_begin, tile_end) (at most 2^18 tiles: positions inside a chunk fit 28 bits); cursor zeroed by the caller
int launch_s2_gather_sliced(const uint32_t* image, const Contig* contigs, const Tile* tiles, uint64_t tile_begin, uint64_t tile_end,
                            const HashP& hp, const uint32_t* count, uint32_t* sat, size_t plane_words, uint2* pool, uint32_t* cursor,
                            uint32_t cap, cudaStream_t st) {
    if (tile_end <= tile_begin) return 0;
    GsSink sink{pool, cursor, cap};
    int slice_shift = hp.k - kGsLog2;
    unsigned grid = (unsigned)(tile_end - tile_begin);
    uint64_t bit0 = tile_begin * kTile;
    switch (hp.e) {
        case 1: s2_gsemit_kernel<1><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, tile_begin, hp, slice_shift, plane_words, count, sat, sink); break;
        case 2: s2_gsemit_kernel<2><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, tile_begin, hp, slice_shift, plane_words, count, sat, sink); break;
        case 3: s2_gsemit_kernel<3><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, tile_begin, hp, slice_shift, plane_words, count, sat, sink); break;
        default: s2_gsemit_kernel<4><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, tile_begin, hp, slice_shift, plane_words, count, sat, sink); break;
    }
    s2_gsapply_kernel<<<dim3(kSMs * 4, kGsBuckets), 256, 0, st>>>(sink, slice_shift, bit0, plane_words, count, sat);
    return 2;
}

int launch_s2_gather_combine(const uint32_t* sat, size_t plane_words, int e, uint64_t tile_begin, uint64_t tile_end, uint32_t* single,
                             uint32_t* trio, cudaStream_t st) {
    if (tile_end <= tile_begin) return 0;
    s2_gscombine_kernel<<<kSMs * 8, 256, 0, st>>>(sat, plane_words, e, (size_t)tile_begin * kTileWords, (size_t)tile_end * kTileWords, single, trio);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// Peak registration through buckets.  Every flagged position stamps its peak id on up to e entries of the
// 2^k-entry peak table: as direct atomics that is one random 128-byte DRAM read-modify-write per k-mer (cfg4: ~6 G of
// them, 466 ms).  Instead (table index, id) records are appended to 512 buckets = the top 9 bits of the table index -- the
// peak table shares the count table's leaf-major layout, whose top bits are the uniform middle bits of the
// hash -- through per-CTA shared-memory stages flushed in runs, and applied bucket by bucket: a bucket is one contiguous
// 1/512 of both tables (32 MiB and 2 MiB at k = 32), so while it is being applied the scatter-max and the count > 0 test
// (E:250,265) run against L2, not DRAM.
// ------------------------------------------------------------------------------------------------
constexpr int kRegBuckets = 512, kRegStage = 16, kRegFlushMin = 4;    // 66 KiB of stage per CTA: three CTAs per SM (24 / 8 gave two: 25 % warps active, profiles/r02g)
constexpr int kRegCursorStride = 32;                           // words between bucket cursors (own 128-byte line each)

struct RegSink {
    uint2* pool[2]; uint32_t* cursor; uint32_t cap;            // bucket b occupies pool[b >> 8][(b & 255) * cap .. +cap) (two pools: the two S1
    int shift;                                                 // stream pools, both idle now); cursor[b * kRegCursorStride]; bucket = g >> shift
    __device__ __forceinline__ uint2* region(uint32_t b) const { return pool[b >> 8] + (size_t)(b & 255u) * cap; }
};

// record = (table index g, peak id)
__device__ __forceinline__ void reg_apply_one(uint32_t g, uint32_t id, const uint32_t* __restrict__ count,
                                              uint32_t* __restrict__ peak_kmer, uint32_t* __restrict__ prefilter) {
    if (((ld_table(count + (g >> 4)) >> ((g & 15u) * 2)) & 3u) == 0u) return;   // E:250,265: hit > 0
    atomicMax(peak_kmer + g, id);
    if (prefilter) { uint32_t slot = prefilter_slot(g); atomicOr(prefilter + (slot >> 5), 1u << (slot & 31)); }
}

template <int E>
__global__ void __launch_bounds__(256, 3) s2_regemit_kernel(const uint32_t* __restrict__ image, const Contig* __restrict__ contigs,
                                                            const Tile* __restrict__ tiles, const uint32_t* __restrict__ need_list,
                                                            const uint32_t* __restrict__ n_need, uint32_t it_lo, uint32_t it_hi,
                                                            uint64_t t_lo, uint64_t t_hi, HashP hp,
                                                            const uint32_t* __restrict__ count, const uint32_t* __restrict__ flagged,
                                                            const uint32_t* __restrict__ tile_base, int32_t* __restrict__ loci, uint32_t loci_cap,
                                                            uint32_t* __restrict__ peak_kmer, uint32_t* __restrict__ prefilter, RegSink sink) {
    extern __shared__ __align__(16) uint32_t dyn[];
    uint2* stage = reinterpret_cast<uint2*>(dyn);               // [kRegBuckets][kRegStage]
    uint32_t* cnt = dyn + 2 * kRegBuckets * kRegStage;          // [kRegBuckets]
    __shared__ uint32_t opener[kTileWords];
    __shared__ int cum[kTileWords + 1];
    const int e = E ? E : hp.e;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < kRegBuckets; b += 256) cnt[b] = 0;
    const uint32_t hi = min(it_hi, *n_need);
    auto direct = [&](uint32_t b, uint2 r) {                     // past a stage or a bucket region: rare, exact either way
        uint32_t g = atomicAdd(sink.cursor + b * kRegCursorStride, 1u);
        if (g < sink.cap) sink.region(b)[g] = r;
        else reg_apply_one(r.x, r.y, count, peak_kmer, prefilter);
    };
    // A warp hands over its share of the buckets, 32 at a time: lane l looks at bucket base + l, the reservations are one
    // atomic per flushing bucket, and every run leaves as one coalesced store per 32 records (a lane writing its own
    // bucket's run would make each store instruction touch 32 different sectors).
    auto flush = [&](bool final) {
        for (int base = warp * 32; base < kRegBuckets; base += 256) {
            int b = base + lane;
            uint32_t n = min(cnt[b], (uint32_t)kRegStage);
            bool go = n >= (uint32_t)kRegFlushMin || (final && n);
            uint32_t g = go ? atomicAdd(sink.cursor + b * kRegCursorStride, n) : 0u;
            if (go) cnt[b] = 0; else cnt[b] = n;
            uint32_t m = __ballot_sync(kFull, go);
            // four flushing buckets per step, eight lanes each (a run is at most 16 records: two stores per lane): with ~6
            // records arriving per bucket per tile nearly every bucket flushes every round, and one bucket per step made
            // this loop most of the kernel's instructions (profiles/r02g_s2_regemit_kernel.md)
            const int grp = lane >> 3, sub = lane & 7;
            while (m) {
                uint32_t pos = __fns(m, 0, grp + 1);              // the (grp+1)-th flushing bucket of the group, 0xffffffff if none
                int src = pos < 32u ? (int)pos : 0;
                uint32_t nb = __shfl_sync(kFull, n, src), gb = __shfl_sync(kFull, g, src);
                if (pos < 32u) {
                    int bb = base + src;
#pragma unroll
                    for (int q = sub; q < kRegStage; q += 8)
                        if ((uint32_t)q < nb) {
                            uint2 r = stage[bb * kRegStage + q];
                            if (gb + q < sink.cap) sink.region(bb)[gb + q] = r;
                            else reg_apply_one(r.x, r.y, count, peak_kmer, prefilter);
                        }
                }
                // drop the (up to) four lowest set bits
                m &= m - 1; m &= m - 1; m &= m - 1; m &= m - 1;
            }
        }
    };
    static_assert(kRegStage <= 16, "a run is written by eight lanes, two records each");
    for (uint32_t it = it_lo + blockIdx.x; it < hi; it += gridDim.x) {
        const uint64_t tix = need_list[it];
        if (tix < t_lo || tix >= t_hi) continue;
        Tile t = tiles[tix];
        Contig c = contigs[t.contig];
        bool fl[4], op[4];
        bool any = false;
        __syncthreads();                                        // previous round: stage writers and opener/cum readers are done
        flush(false);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            int xl = r * 256 + threadIdx.x;
            uint32_t w = flagged[(size_t)tix * kTileWords + (xl >> 5)];
            fl[r] = (w >> (xl & 31)) & 1u;
            op[r] = fl[r] && opens_peak(flagged, tix, t.j0, (long)t.j0 + xl);
            uint32_t wo = __ballot_sync(kFull, op[r]);
            if (lane == 0) opener[r * 8 + warp] = wo;
            any |= fl[r];
        }
        if (!__syncthreads_or(any)) continue;                   // (also orders the flush before this round's stage writes)
        prefix_words(opener, cum, kTileWords);
        __syncthreads();
        long np = (long)c.len - hp.k + 1;
        uint32_t base = tile_base[tix];
        uint32_t ids[4], hv[4][E ? E : kMaxE];
#pragma unroll
        for (int r = 0; r < 4; ++r) {                                          // every hash load of this thread goes out before any is used
            int xl = r * 256 + threadIdx.x;
            long j = (long)t.j0 + xl;
            ids[r] = fl[r] ? base + (uint32_t)bits_upto(opener, cum, xl) - 1u : 0u;
            if (fl[r] && op[r] && ids[r] < loci_cap) { loci[2 * (size_t)ids[r]] = (int32_t)t.contig + 1; loci[2 * (size_t)ids[r] + 1] = (int32_t)j; }
            bool reg = fl[r] && j < np && ids[r] != 0u;                        // j = len-k+1 reads the zero tail (Q6); id 0 never registers (Q10)
            const uint32_t* hashes = image + c.hash_word + (size_t)j * e;
#pragma unroll
            for (int i = 0; i < (E ? E : kMaxE); ++i) hv[r][i] = (reg && i < e) ? ld_stream(hashes + i) : 0u;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < (E ? E : kMaxE); ++i) {
                uint32_t h = hv[r][i];
                if (!h) continue;
                uint32_t g = tbl_index(h, hp), b = g >> sink.shift;
                uint32_t slot = atomicAdd(&cnt[b], 1u);
                if (slot < (uint32_t)kRegStage) stage[b * kRegStage + slot] = make_uint2(g, ids[r]);
                else direct(b, make_uint2(g, ids[r]));
            }
        }
    }
    __syncthreads();
    flush(true);
}

// grid (parts, kRegBuckets): blocks are dispatched in index order, so at any time the resident CTAs work on one or two
// adjacent buckets; a bucket is a contiguous 1/512 of both tables (peak table: 32 MiB, count table: 2 MiB at k = 32)
__global__ void __launch_bounds__(256, 4) s2_regapply_kernel(RegSink sink, const uint32_t* __restrict__ count,
                                                             uint32_t* __restrict__ peak_kmer, uint32_t* __restrict__ prefilter) {
    const uint32_t b = blockIdx.y;
    const uint32_t n = min(sink.cursor[b * kRegCursorStride], sink.cap);
    const uint2* __restrict__ in = sink.region(b);
    for (uint32_t x = blockIdx.x * 256 + threadIdx.x; x < n; x += gridDim.x * 256) {
        uint2 r;
        asm volatile("ld.global.cs.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(in + x));   // evict-first: the table slices are what L2 is for
        reg_apply_one(r.x, r.y, count, peak_kmer, prefilter);
    }
}

size_t s2_regemit_smem() { return (size_t)kRegBuckets * kRegStage * 8 + (size_t)kRegBuckets * 4; }
int s2_reg_buckets() { return kRegBuckets; }
int s2_reg_cursor_words() { return kRegBuckets * kRegCursorStride; }

// one chunk of the needed-tile list [it_lo, it_hi): emit, then apply (cursor zeroed by the caller)
int launch_s2_register_bucketed(const uint32_t* image, const Contig* contigs, const Tile* tiles, const uint32_t* need_list,
                                const uint32_t* n_need, uint32_t it_lo, uint32_t it_hi, uint64_t t_lo, uint64_t t_hi, const HashP& hp, const uint32_t* count,
                                const uint32_t* flagged, const uint32_t* tile_base, int32_t* loci, uint32_t loci_cap, uint32_t* peak_kmer,
                                uint32_t* prefilter, uint2* pool_lo, uint2* pool_hi, uint32_t* cursor, uint32_t cap, cudaStream_t st) {
    if (it_hi <= it_lo) return 0;
    RegSink sink{{pool_lo, pool_hi}, cursor, cap, hp.k > 9 ? hp.k - 9 : 0};
    size_t smem = s2_regemit_smem();
    unsigned grid = it_hi - it_lo < (uint32_t)kSMs * 3 ? it_hi - it_lo : (uint32_t)kSMs * 3;
#define LHGT_REGEMIT(EE)                                                                                                        \
    do {                                                                                                                        \
        if (cudaFuncSetAttribute(s2_regemit_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) \
            return -1;                                                                                                          \
        s2_regemit_kernel<EE><<<grid, 256, smem, st>>>(image, contigs, tiles, need_list, n_need, it_lo, it_hi, t_lo, t_hi, hp, count, flagged, \
                                                      tile_base, loci, loci_cap, peak_kmer, prefilter, sink);                  \
    } while (0)
    if (hp.e == 3) LHGT_REGEMIT(3); else LHGT_REGEMIT(0);
#undef LHGT_REGEMIT
    s2_regapply_kernel<<<dim3(kSMs * 4, kRegBuckets), 256, 0, st>>>(sink, count, peak_kmer, prefilter);
    return 2;
}

constexpr int kS2Grid = kSMs * 8;                              // persistent grids over the needed-tile list

int launch_s2_gather(const uint32_t* image, const Contig* contigs, const Tile* tiles, uint64_t tile_begin, uint64_t tile_end,
                     const HashP& hp, const uint32_t* count, uint32_t* single, uint32_t* trio, cudaStream_t st) {
    if (tile_end <= tile_begin) return 0;
    unsigned grid = (unsigned)(tile_end - tile_begin);
    switch (hp.e) {
        case 1: s2_trio_kernel<1><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
        case 2: s2_trio_kernel<2><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
        case 3: s2_trio_kernel<3><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
        case 4: s2_trio_kernel<4><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
        default: s2_trio_kernel<0><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
    }
    return 1;
}

// scratch: 2 * ceil(ntiles / 1024) + scan_tmp_words(ceil(ntiles / 1024)) words
size_t s2_mark_scratch_words(uint64_t ntiles) { uint64_t b = (ntiles + 1023) / 1024; return 2 * b + scan_tmp_words(b) + 2; }

int launch_s2_mark(const Contig* contigs, const Tile* tiles, uint64_t ntiles, const uint32_t* trio, int three_min, uint8_t* hot,
                   uint32_t* need_list, uint32_t* n_need, uint32_t* scratch, cudaStream_t st) {
    if (!ntiles) return 0;
    uint64_t blocks = (ntiles + 1023) / 1024;
    uint32_t *cnt = scratch, *base = scratch + blocks, *tmp = scratch + 2 * blocks;
    s2_hot_kernel<<<(unsigned)((ntiles + 7) / 8), 256, 0, st>>>(contigs, tiles, ntiles, trio, three_min, hot);
    s2_need_count_kernel<<<(unsigned)blocks, 256, 0, st>>>(tiles, ntiles, hot, cnt);
    int l = 2 + launch_scan_exclusive(cnt, base, blocks, tmp, st);
    s2_need_write_kernel<<<(unsigned)blocks, 256, 0, st>>>(tiles, ntiles, hot, cnt, base, need_list, n_need);
    return l + 1;
}

int launch_s2_single(const uint32_t* image, const Contig* contigs, const Tile* tiles, const uint32_t* need_list, const uint32_t* n_need,
                     uint64_t tile_begin, uint64_t tile_end, const HashP& hp, const uint32_t* count, uint32_t* single, cudaStream_t st) {
    if (tile_end <= tile_begin || hp.e < 2) return 0;
    switch (hp.e) {
        case 2: s2_single_kernel<2><<<kS2Grid, 256, 0, st>>>(image, contigs, tiles, need_list, n_need, tile_begin, tile_end, hp, count, single); break;
        case 3: s2_single_kernel<3><<<kS2Grid, 256, 0, st>>>(image, contigs, tiles, need_list, n_need, tile_begin, tile_end, hp, count, single); break;
        case 4: s2_single_kernel<4><<<kS2Grid, 256, 0, st>>>(image, contigs, tiles, need_list, n_need, tile_begin, tile_end, hp, count, single); break;
        default: s2_single_kernel<0><<<kS2Grid, 256, 0, st>>>(image, contigs, tiles, need_list, n_need, tile_begin, tile_end, hp, count, single); break;
    }
    return 1;
}

int launch_s2_good(const Contig* contigs, const Tile* tiles, const uint32_t* need_list, const uint32_t* n_need, uint64_t t_lo, uint64_t t_hi,
                   const uint32_t* single, const uint32_t* trio, int one_min, int three_min, uint32_t* good, cudaStream_t st) {
    s2_good_kernel<<<kS2Grid, 256, 0, st>>>(contigs, tiles, need_list, n_need, t_lo, t_hi, single, trio, one_min, three_min, good);
    return 1;
}

int launch_s2_flag(const Contig* contigs, const Tile* tiles, uint64_t ntiles, const uint32_t* need_list, const uint32_t* n_need,
                   uint64_t t_lo, uint64_t t_hi, int k, const uint32_t* single, const uint32_t* good, uint32_t* flagged, cudaStream_t st) {
    s2_flag_kernel<<<kS2Grid / 2, 256, 0, st>>>(contigs, tiles, ntiles, need_list, n_need, t_lo, t_hi, k, single, good, flagged);
    return 1;
}

int launch_s2_count_new(const Tile* tiles, const uint32_t* need_list, const uint32_t* n_need, uint64_t t_lo, uint64_t t_hi,
                        const uint32_t* flagged, uint32_t* tile_new, unsigned long long* flagged_total, cudaStream_t st) {
    s2_count_new_kernel<<<kS2Grid, 256, 0, st>>>(tiles, need_list, n_need, t_lo, t_hi, flagged, tile_new, flagged_total);
    return 1;
}

int launch_s2_register(const uint32_t* image, const Contig* contigs, const Tile* tiles, const uint32_t* need_list, const uint32_t* n_need,
                       uint64_t t_lo, uint64_t t_hi, const HashP& hp, const uint32_t* count, const uint32_t* flagged, const uint32_t* tile_base,
                       int32_t* loci, uint32_t loci_cap, uint32_t* peak_kmer, uint32_t* prefilter, int mode, cudaStream_t st) {
    s2_register_kernel<0><<<kS2Grid, 256, 0, st>>>(image, contigs, tiles, need_list, n_need, t_lo, t_hi, hp, count, flagged, tile_base, loci, loci_cap,
                                                  peak_kmer, prefilter, mode);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// S3: read-pair confirmation (E:419-499 + Split_reads E:109-202).  Scan: one warp per pair; lanes hash
// positions in parallel, test the L2-resident pre-filter (sparse results only) and probe the 2^k-entry peak
// table; positions that hold a peak k-mer are appended, in read order, to a per-warp list.  Vote: pairs that
// pass the exact pre-check (s3_may_split) hand their list to s3_vote_kernel (one thread per pair); when the
// hand-over is not possible the warp runs the order-dependent vote (judge_base) itself.
// ------------------------------------------------------------------------------------------------
#ifndef LHGT_S3_WARPS
#define LHGT_S3_WARPS 16
#endif
#ifndef LHGT_S3_CTAS
#define LHGT_S3_CTAS 2
#endif
constexpr int kS3Warps = LHGT_S3_WARPS;
int s3_warps_per_block() { return kS3Warps; }
int s3_grid_blocks(int) { return kSMs * LHGT_S3_CTAS; }

// Peak ids are handed out in (contig, position) order, so the contig of a peak is a search in the first-id-per-contig
// array (n_contigs + 1 entries, a few KB: on chip) instead of a random gather from the per-peak loci.  Returns the 1-based
// record ordinal = the largest c with contig_first[c - 1] <= id (contigs without peaks share their successor's first id).
__device__ __forceinline__ int contig_of_peak(const uint32_t* contig_first, uint32_t n_contigs, uint32_t id) {
    uint32_t lo = 0, hi = n_contigs;                             // invariant: contig_first[lo] <= id (peak 0 opens contig_first[.] = 0)
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (contig_first[mid] <= id) lo = mid; else hi = mid;
    }
    return (int)lo + 1;
}

// src points into the warp's shared-memory stage (s3_pairs_kernel).
template <int E>
__device__ __forceinline__ int s3_scan_mate(const uint8_t* src, int len, const uint8_t* lut, const HashP& hp,
                                            const uint32_t* __restrict__ prefilter,
                                            const uint32_t* __restrict__ peak_kmer, const uint32_t* contig_first, uint32_t n_contigs,
                                            uint32_t* __restrict__ cands, int32_t* __restrict__ cont, int n_listed, int lane) {
    const int e = E ? E : hp.e;
    int np = len - hp.k + 1;
    if (np <= 0) return n_listed;
    int nch = (np + 31) >> 5;
    Planes prev = pack_word(lane < len ? src[lane] : 0u, lut);
    uint32_t c1 = 32 + lane < len ? src[32 + lane] : 0u;      // bytes run one word ahead of the hashing
    for (int w = 1; w <= nch; ++w) {                          // chunk w-1: positions 32(w-1) + lane, ascending
        Planes cur = pack_word(c1, lut);
        int pn = (w + 1) * 32 + lane;
        c1 = pn < len ? src[pn] : 0u;
        LeWin kw = le_window(prev, cur, lane, hp);
        prev = cur;
        uint32_t h[E ? E : kMaxE], pk[E ? E : kMaxE], fw[E ? E : kMaxE];
        // all e filter words are requested back to back; .cg: a probe must not claim an L1 line (the lines in
        // flight, not the warps, would bound the probes in flight)
#pragma unroll
        for (int i = 0; i < (E ? E : kMaxE); ++i)
            if (i < e) {
                h[i] = tbl_index(le_hash(kw, hp, i), hp);                 // table index from here on (count and peak tables share the layout)
                uint32_t slot = prefilter_slot(h[i]);
                fw[i] = !kw.valid ? 0u : prefilter ? ld_table(prefilter + (slot >> 5)) >> (slot & 31) : 1u;   // no filter: it is saturated
            }
        bool any = false;
#pragma unroll
        for (int i = 0; i < (E ? E : kMaxE); ++i)
            if (i < e) {
                pk[i] = (fw[i] & 1u) ? ld_table(peak_kmer + h[i]) : 0u;
                any |= pk[i] != 0u;
            }
        uint32_t mask = __ballot_sync(kFull, any);
        if (mask) {
            if (any) {
                int slot = n_listed + __popc(mask & ((1u << lane) - 1u));
#pragma unroll
                for (int i = 0; i < (E ? E : kMaxE); ++i)
                    if (i < e) {                                  // the lane that found the peak also names its contig
                        cands[(size_t)slot * e + i] = pk[i];
                        cont[(size_t)slot * e + i] = pk[i] ? contig_of_peak(contig_first, n_contigs, pk[i]) : 0;
                    }
            }
            n_listed += __popc(mask);
        }
    }
    __syncwarp();
    return n_listed;
}

// judge_base (E:118-159) + check_split (E:161-202) for one pair, run by a single lane.
__device__ void s3_vote(const uint32_t* cands, int n_listed, int e, const int32_t* __restrict__ loci, int32_t* tally,
                        uint8_t* __restrict__ peak_filter) {
    int n_t = 0;    // tally[3*t] = contig, [3*t+1] = votes, [3*t+2] = first peak id
    for (int f = 0; f < n_listed; ++f) {
        uint32_t sel_peak = 0; int sel_contig = 0, sel_votes = 0, sel_t = -1;
        for (int i = 0; i < e; ++i) {
            uint32_t pk = cands[(size_t)f * e + i];
            if (!pk) continue;
            int contig = loci[2 * (size_t)pk];
            int at = -1;
            for (int t = 0; t < n_t; ++t) if (tally[3 * t] == contig) { at = t; break; }
            if (at >= 0) {
                if (tally[3 * at + 1] >= sel_votes) { sel_peak = pk; sel_contig = contig; sel_votes = tally[3 * at + 1]; sel_t = at; }
            } else if (sel_peak == 0) { sel_peak = pk; sel_contig = contig; sel_votes = 0; sel_t = -1; }
        }
        if (sel_t >= 0) tally[3 * sel_t + 1] += 1;
        else { tally[3 * n_t] = sel_contig; tally[3 * n_t + 1] = 1; tally[3 * n_t + 2] = (int32_t)sel_peak; ++n_t; }
    }
    int largest = 0, second = 0, strong = 0;
    for (int t = 0; t < n_t; ++t) {
        int n = tally[3 * t + 1];
        if (n < 6) continue;
        ++strong;
        if (n >= largest) { second = largest; largest = n; }
        else if (n >= second) second = n;
    }
    if (strong < 2) return;
    for (int t = 0; t < n_t; ++t) {
        int n = tally[3 * t + 1];
        if (n >= 6 && (n == largest || n == second)) peak_filter[(uint32_t)tally[3 * t + 2]] = 1;   // only >= 1 is consumed (E:526)
    }
}

// The single-lane vote with the tally direct-addressed by contig (one L1-resident load per candidate instead of a search
// through the contigs seen so far): for pairs that touch many contigs -- a dense peak table makes nearly every position
// vote.  Entries carry the epoch of the pair that wrote them, so nothing is cleared between pairs.
__device__ void s3_vote_table(const uint32_t* cands, const int32_t* cont, int n_listed, int e, uint32_t* vt, uint32_t n_contigs,
                              int32_t* touched, uint8_t* __restrict__ peak_filter) {
    uint32_t* votes = vt + 1;
    uint32_t* first = vt + 1 + (n_contigs + 1);
    uint32_t epoch = vt[0] + 1;
    if (epoch >= (1u << 22)) {                                  // tags would wrap: start over
        for (uint32_t c = 0; c <= n_contigs; ++c) votes[c] = 0;
        epoch = 1;
    }
    vt[0] = epoch;
    int n_t = 0;
    for (int f = 0; f < n_listed; ++f) {
        uint32_t sel_peak = 0, sel_tv = 0;
        int sel_contig = 0, sel_votes = 0;
        bool sel_seen = false;
        for (int i = 0; i < e; ++i) {
            uint32_t pk = __ldcg(cands + (size_t)f * e + i);
            if (!pk) continue;
            int contig = __ldcg(cont + (size_t)f * e + i);
            uint32_t tv = votes[contig];
            if ((tv >> 10) == epoch) {
                int v = (int)(tv & 1023u);
                if (v >= sel_votes) { sel_peak = pk; sel_contig = contig; sel_votes = v; sel_seen = true; sel_tv = tv; }
            } else if (sel_peak == 0) { sel_peak = pk; sel_contig = contig; sel_votes = 0; sel_seen = false; }
        }
        if (sel_seen) votes[sel_contig] = sel_tv + 1;           // <= 2 * kMaxReadLen votes: fits the 10 bits
        else { votes[sel_contig] = (epoch << 10) | 1u; first[sel_contig] = sel_peak; touched[n_t++] = sel_contig; }
    }
    int largest = 0, second = 0, strong = 0;
    for (int t = 0; t < n_t; ++t) {
        int n = (int)(votes[touched[t]] & 1023u);
        if (n < 6) continue;
        ++strong;
        if (n >= largest) { second = largest; largest = n; }
        else if (n >= second) second = n;
    }
    if (strong < 2) return;
    for (int t = 0; t < n_t; ++t) {
        int n = (int)(votes[touched[t]] & 1023u);
        if (n >= 6 && (n == largest || n == second)) peak_filter[first[touched[t]]] = 1;
    }
}

// The same vote run by the whole warp: lane t keeps tally entry t (contig, votes, first peak) in registers, so
// "has this contig been seen, and with how many votes" is one ballot + one shuffle instead of a search, and the
// candidates of 32 listed positions are fetched with one coalesced load each.  The walk over positions stays sequential
// (every vote depends on the tally so far).  Returns false when a pair touches more than 32 contigs: the caller then
// runs the single-lane form.
template <int E>
__device__ __forceinline__ bool s3_vote_warp(const uint32_t* cands, const int32_t* cont, int n_listed, int e,
                                             uint8_t* __restrict__ peak_filter, int lane) {
    int t_contig = 0, t_votes = 0, n_t = 0;
    uint32_t t_first = 0;
    for (int base = 0; base < n_listed; base += 32) {
        const int f = base + lane;
        uint32_t pk[E ? E : kMaxE];
        int ct[E ? E : kMaxE];
#pragma unroll
        for (int i = 0; i < (E ? E : kMaxE); ++i) {
            bool on = i < e && f < n_listed;
            pk[i] = on ? __ldcg(cands + (size_t)f * e + i) : 0u;         // written by other lanes of this warp: read at L2
            ct[i] = on ? __ldcg(cont + (size_t)f * e + i) : 0;
        }
        const int lim = min(32, n_listed - base);
        for (int j = 0; j < lim; ++j) {
            uint32_t sel_peak = 0;
            int sel_contig = 0, sel_votes = 0, sel_t = -1;
#pragma unroll
            for (int i = 0; i < (E ? E : kMaxE); ++i) {
                uint32_t p = __shfl_sync(kFull, pk[i], j);
                int c = __shfl_sync(kFull, ct[i], j);
                if (i >= e || !p) continue;                               // warp-uniform
                uint32_t m = __ballot_sync(kFull, lane < n_t && t_contig == c);
                if (m) {
                    int at = __ffs(m) - 1;
                    int v = __shfl_sync(kFull, t_votes, at);
                    if (v >= sel_votes) { sel_peak = p; sel_contig = c; sel_votes = v; sel_t = at; }
                } else if (sel_peak == 0) { sel_peak = p; sel_contig = c; sel_votes = 0; sel_t = -1; }
            }
            if (sel_t >= 0) { if (lane == sel_t) ++t_votes; }
            else {
                if (n_t == 32) return false;
                if (lane == n_t) { t_contig = sel_contig; t_votes = 1; t_first = sel_peak; }
                ++n_t;
            }
        }
    }
    int v = lane < n_t && t_votes >= 6 ? t_votes : 0;                     // check_split (E:161-202)
    if (__popc(__ballot_sync(kFull, v > 0)) < 2) return true;
    int largest = v;
#pragma unroll
    for (int d = 16; d; d >>= 1) largest = max(largest, __shfl_xor_sync(kFull, largest, d));
    int second = largest;
    if (__popc(__ballot_sync(kFull, v == largest)) < 2) {
        second = v == largest ? 0 : v;
#pragma unroll
        for (int d = 16; d; d >>= 1) second = max(second, __shfl_xor_sync(kFull, second, d));
    }
    if (v > 0 && (v == largest || v == second)) peak_filter[t_first] = 1;     // only >= 1 is consumed (E:526)
    return true;
}

// Can this pair's vote mark anything?  check_split (E:161-202) needs TWO contigs with >= 6 votes, and a contig cannot get
// more votes than the positions at which it is a candidate.  So: take the contig c* most common among 32 sampled
// candidates, count its candidates exactly, and count the others in 1024 hashed 16-bit slots (a slot may merge contigs:
// an upper bound); unless [c* has >= 6] + sum over slots of floor(count / 6) reaches 2 the vote cannot have two such
// contigs and is skipped.  Exact (a necessary condition), and in a dense result -- where every position of every pair
// holds candidates, nearly all of them of the read's own genome plus scattered collisions -- it spares almost every
// pair the serial vote and the trip through the arena.
constexpr int kMaySlots = 1024;                                // (256 slots let cfg4's ~500 collision candidates per pair fake a second contig in every pair)
constexpr int kCfShared = 4096;                                // contig_first entries kept in shared memory (s3_pairs_kernel)
__device__ __forceinline__ bool s3_may_split(const int32_t* __restrict__ cont, int n_entries, uint32_t* __restrict__ tw /* kMaySlots / 2 words */, int lane) {
    for (int x = lane; x < kMaySlots / 2; x += 32) tw[x] = 0u;
    int probe = __ldcg(cont + (int)(((long)lane * n_entries) >> 5));               // 32 evenly spaced entries (some are 0 = no candidate)
    uint32_t same = __match_any_sync(kFull, probe);
    int votes = probe ? __popc(same) : 0, best = votes;
#pragma unroll
    for (int d = 16; d; d >>= 1) best = max(best, __shfl_xor_sync(kFull, best, d));
    uint32_t who = __ballot_sync(kFull, votes == best && probe != 0);
    if (!who) return false;                                                          // (no candidate among the samples: cannot happen with n_listed >= 6 ... but then nothing to vote on either way)
    const int cstar = __shfl_sync(kFull, probe, __ffs(who) - 1);
    __syncwarp();
    int mine = 0;
    for (int x = lane; x < n_entries; x += 32) {
        int c = __ldcg(cont + x);
        if (c == cstar) ++mine;
        else if (c) {
            uint32_t slot = ((uint32_t)c * 2654435761u) >> 22;                       // 10 bits
            atomicAdd(tw + (slot >> 1), 1u << ((slot & 1u) * 16));                   // < 2^16 candidates per pair: no carry into the neighbour
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) mine += __shfl_xor_sync(kFull, mine, d);
    __syncwarp();
    int potential = 0;
    for (int x = lane; x < kMaySlots / 2; x += 32) { uint32_t w = tw[x]; potential += (int)((w & 0xffffu) / 6u) + (int)((w >> 16) / 6u); }
#pragma unroll
    for (int d = 16; d; d >>= 1) potential += __shfl_xor_sync(kFull, potential, d);
    return (mine >= 6 ? 1 : 0) + potential >= 2;
}

struct PairOff { uint64_t a0, b0; uint32_t l1, l2; bool ok; };

// One warp per pair, two pairs ahead: while a pair is voted on, the TMA unit is staging the bytes of the next sampled
// pair into the warp's other shared-memory slots and the offsets of the one after are in flight (see stage_read).
template <int E>
__global__ void __launch_bounds__(kS3Warps * 32, LHGT_S3_CTAS) s3_pairs_kernel(
    const uint8_t* __restrict__ fq1, const uint64_t* __restrict__ s1, const uint64_t* __restrict__ e1, uint64_t nrec1,
    const uint8_t* __restrict__ fq2, const uint64_t* __restrict__ s2, const uint64_t* __restrict__ e2, uint64_t nrec2,
    uint64_t tail_start, uint64_t tail_len, uint64_t first, uint64_t count, const uint32_t* __restrict__ sample_bits,
    uint64_t ordinal_base, HashP hp, const uint32_t* __restrict__ prefilter, const uint32_t* __restrict__ peak_kmer,
    const int32_t* __restrict__ loci, uint8_t* __restrict__ peak_filter, S3Scratch scratch,
    unsigned long long* __restrict__ n_sampled, int* __restrict__ err) {
    const uint32_t n_contigs = scratch.n_contigs;
    __shared__ uint8_t lut[256];
    __shared__ __align__(16) uint8_t stage[kS3Warps][4][kStageBytes];
    __shared__ __align__(8) uint64_t sbar[kS3Warps][2];        // "pair staged", one per slot pair
    extern __shared__ uint32_t may_dyn[];                      // s3_may_split's slots: kS3Warps x kMaySlots / 2 words, then contig_first
    // The peak -> contig search runs up to e times per position (11 dependent loads each at 2 000 contigs): from shared
    // memory when the array fits, because what is left of L1 beside this kernel's shared memory does not keep it (measured:
    // the scan went from 148 to 209 ms when the slots above grew and squeezed L1).
    uint32_t* cf_s = may_dyn + kS3Warps * (kMaySlots / 2);
    const bool cf_fits = n_contigs + 1 <= (uint32_t)kCfShared;
    if (cf_fits) for (uint32_t x = threadIdx.x; x <= n_contigs; x += blockDim.x) cf_s[x] = scratch.contig_first[x];
    const uint32_t* contig_first = cf_fits ? cf_s : scratch.contig_first;
    const int e = E ? E : hp.e;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { mbar_init(&sbar[warp][0], 1); mbar_init(&sbar[warp][1], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fill_base_lut(lut);
    __syncthreads();
    uint64_t gwarp = (uint64_t)blockIdx.x * kS3Warps + warp;
    uint32_t* cands = scratch.cands + gwarp * scratch.cands_stride;            // first half: peak ids, second half: their contigs
    int32_t* cont = reinterpret_cast<int32_t*>(cands + scratch.cands_stride / 2);
    int32_t* tally = scratch.tally + gwarp * scratch.tally_stride;
    unsigned long long mine = 0;
    const uint64_t stride = (uint64_t)gridDim.x * kS3Warps;
    const uint64_t last = first + count < nrec1 ? first + count : nrec1;
    auto next_sampled = [&](uint64_t q) {
        while (q < last && !is_sampled(sample_bits, q + ordinal_base)) q += stride;
        return q;
    };
    auto load_off = [&](uint64_t r) {
        PairOff o;
        uint64_t l1, l2;
        o.a0 = s1[r]; l1 = e1[r] - o.a0;
        if (r < nrec2) { o.b0 = s2[r]; l2 = e2[r] - o.b0; }
        else { o.b0 = tail_start; l2 = tail_len; }            // fq2 exhausted: std::getline leaves its last string
        o.ok = l1 <= (uint64_t)kMaxReadLen && l2 <= (uint64_t)kMaxReadLen;
        o.l1 = (uint32_t)l1; o.l2 = (uint32_t)l2;
        return o;
    };
    auto stage_pair = [&](const PairOff& o, int sl) {
        stage_pair_reads(stage[warp][2 * sl], stage[warp][2 * sl + 1], &sbar[warp][sl], fq1, o.a0, o.l1, fq2, o.b0, o.l2, lane);
    };
    uint64_t rn = next_sampled(first + gwarp);
    PairOff on{0, 0, 0, 0, false}, onn{0, 0, 0, 0, false};
    if (rn < last) on = load_off(rn);
    uint64_t rnn = rn < last ? next_sampled(rn + stride) : rn;
    if (rnn < last) onn = load_off(rnn);
    int slot = 0;
    bool nstaged = false;
    uint32_t phase = 0;                                         // bit s: parity the next fill of slot pair s completes
    if (rn < last && on.ok) { stage_pair(on, 1); nstaged = true; }
    while (rn < last) {
        PairOff cur = on;
        bool staged = nstaged;
        nstaged = false;
        if (cur.ok) {
            slot ^= 1;
            if (!staged) stage_pair(cur, slot);                 // only after a refused pair
            mbar_wait(&sbar[warp][slot], (phase >> slot) & 1u);
            phase ^= 1u << slot;
        }
        rn = rnn; on = onn;
        if (rnn < last) {
            rnn = next_sampled(rnn + stride);
            if (rnn < last) onn = load_off(rnn);
        }
        if (!cur.ok) { if (lane == 0) atomicExch(err, 1); continue; }
        ++mine;
        __syncwarp();                                           // every lane is done with the other slot pair
        if (rn < last && on.ok) { stage_pair(on, slot ^ 1); nstaged = true; }
        uint32_t m1 = (uint32_t)cur.a0 & 15u, m2 = (uint32_t)cur.b0 & 15u;
        int n_listed = s3_scan_mate<E>(stage[warp][2 * slot] + m1, (int)cur.l1, lut, hp, prefilter, peak_kmer, contig_first, n_contigs, cands, cont, 0, lane);
        n_listed = s3_scan_mate<E>(stage[warp][2 * slot + 1] + m2, (int)cur.l2, lut, hp, prefilter, peak_kmer, contig_first, n_contigs, cands, cont, n_listed, lane);
        if (n_listed >= 6 && (__syncwarp(), s3_may_split(cont, n_listed * e, may_dyn + warp * (kMaySlots / 2), lane))) {   // base_hits >= MIN_BASE_NUM (E:496)
            __syncwarp();
            // The order-dependent vote is handed to s3_vote_kernel (one THREAD per pair, 32 pairs per warp in flight): the
            // pair's candidates move to the arena and the pair joins the queue.  When either is full the warp votes itself.
            bool queued = false;
            if (scratch.arena) {
                uint32_t words = (uint32_t)n_listed * (uint32_t)e, off = 0, q = 0;
                if (lane == 0) {
                    off = atomicAdd(scratch.arena_cursor, words);
                    q = off + words <= scratch.arena_cap ? atomicAdd(scratch.queue_count, 1u) : 0xffffffffu;
                }
                off = __shfl_sync(kFull, off, 0); q = __shfl_sync(kFull, q, 0);
                if (q < scratch.queue_cap) {
                    uint2* dst = scratch.arena + off;
                    for (uint32_t x = lane; x < words; x += 32) dst[x] = make_uint2(__ldcg(cands + x), (uint32_t)__ldcg(cont + x));
                    if (lane == 0) scratch.queue[q] = make_uint2(off, (uint32_t)n_listed);
                    queued = true;
                }
            }
            if (!queued && !s3_vote_warp<E>(cands, cont, n_listed, e, peak_filter, lane)) {        // more than 32 contigs in one pair
                if (lane == 0) {
                    if (scratch.vote_table) s3_vote_table(cands, cont, n_listed, e, scratch.vote_table + gwarp * scratch.vote_stride,
                                                          scratch.vote_contigs, tally, peak_filter);
                    else s3_vote(cands, n_listed, e, loci, tally, peak_filter);
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0 && mine) atomicAdd(n_sampled, mine);
}

// The vote of the queued pairs (judge_base + check_split, E:118-202), one thread per pair.  The tally is an open-addressed
// table of `tsize` 32-bit slots per thread (contig << 10 | votes; 0 = empty; tsize a power of two >= twice the longest
// list).  SMEM = true keeps the tables in shared memory, slot s of thread t at tab[s * blockDim.x + t] (conflict-free
// whatever the slots); otherwise they live in global memory, sized to stay L2-resident.  A table is left clean: the slots a
// pair fills are remembered, together with the first peak voted for that contig, in the arena space of candidates
// already consumed, and zeroed at the end.  The e candidates of a position are looked up side by side (their first
// probes are independent loads); only the update is serial.
template <bool SMEM, int E>
__global__ void __launch_bounds__(128) s3_vote_kernel(uint2* __restrict__ arena, const uint2* __restrict__ queue,
                                                      const uint32_t* __restrict__ queue_count, uint32_t queue_cap, int e_rt,
                                                      uint32_t* __restrict__ tables, uint32_t tsize, uint8_t* __restrict__ peak_filter) {
    extern __shared__ uint32_t tab_s[];
    const int e = E ? E : e_rt;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
    uint32_t* table = SMEM ? tab_s + threadIdx.x : tables + (size_t)tid * tsize;
    const uint32_t stride = SMEM ? blockDim.x : 1u;
    const uint32_t mask = tsize - 1u;
    const int shift = 32 - (31 - __clz(tsize));
    if (SMEM) for (uint32_t x = 0; x < tsize; ++x) table[x * stride] = 0u;
    const uint32_t n = min(*queue_count, queue_cap);
    for (uint32_t q = tid; q < n; q += nthreads) {
        uint2 job = queue[q];
        uint2* list = arena + job.x;
        const int n_listed = (int)job.y;
        int n_t = 0;
        // candidates are fetched a block of kAhead positions ahead of the vote: a thread walks its own list, so every load is
        // its own sector and would otherwise put an L2 (or DRAM) round trip into each step of the serial chain
        constexpr int kAhead = 4, EE = E ? E : kMaxE;
        uint2 nxt[kAhead][EE];
        auto fetch = [&](int f0) {
#pragma unroll
            for (int a = 0; a < kAhead; ++a)
#pragma unroll
                for (int i = 0; i < EE; ++i)
                    nxt[a][i] = (i < e && f0 + a < n_listed) ? list[(size_t)(f0 + a) * e + i] : make_uint2(0u, 0u);
        };
        fetch(0);
        for (int f0 = 0; f0 < n_listed; f0 += kAhead) {
            uint2 cur[kAhead][EE];
#pragma unroll
            for (int a = 0; a < kAhead; ++a)
#pragma unroll
                for (int i = 0; i < EE; ++i) cur[a][i] = nxt[a][i];
            fetch(f0 + kAhead);
#pragma unroll
            for (int a = 0; a < kAhead; ++a) {
                if (f0 + a >= n_listed) break;
                uint32_t idx[EE], ent[EE];
#pragma unroll
                for (int i = 0; i < EE; ++i)
                    if (i < e) idx[i] = (cur[a][i].y * 2654435761u) >> shift;
#pragma unroll
                for (int i = 0; i < EE; ++i)
                    if (i < e) ent[i] = cur[a][i].x ? table[idx[i] * stride] : 0u;
                uint32_t sel_peak = 0, sel_slot = 0, sel_entry = 0;
                int sel_votes = 0;
                bool sel_seen = false;
#pragma unroll
                for (int i = 0; i < EE; ++i) {
                    if (i >= e || !cur[a][i].x) continue;
                    uint32_t en = ent[i], ix = idx[i];
                    while (en != 0u && (en >> 10) != cur[a][i].y) { ix = (ix + 1u) & mask; en = table[ix * stride]; }
                    if (en) {
                        int v = (int)(en & 1023u);
                        if (v >= sel_votes) { sel_peak = cur[a][i].x; sel_votes = v; sel_seen = true; sel_slot = ix; sel_entry = en; }
                    } else if (sel_peak == 0) { sel_peak = cur[a][i].x; sel_votes = 0; sel_seen = false; sel_slot = ix; sel_entry = cur[a][i].y << 10; }
                }
                table[sel_slot * stride] = sel_entry + 1u;
                if (!sel_seen) list[n_t++] = make_uint2(sel_slot, sel_peak);  // n_t <= f + 1: that entry has been consumed (fetched) already
            }
        }
        int largest = 0, second = 0, strong = 0;
        for (int t = 0; t < n_t; ++t) {
            int v = (int)(table[list[t].x * stride] & 1023u);
            if (v < 6) continue;
            ++strong;
            if (v >= largest) { second = largest; largest = v; }
            else if (v >= second) second = v;
        }
        for (int t = 0; t < n_t; ++t) {
            uint2 rec = list[t];
            int v = (int)(table[rec.x * stride] & 1023u);
            if (strong >= 2 && v >= 6 && (v == largest || v == second)) peak_filter[rec.y] = 1;   // only >= 1 is consumed (E:526)
            table[rec.x * stride] = 0u;
        }
    }
}

__global__ void contig_first_kernel(const Contig* __restrict__ contigs, uint32_t n_contigs, const uint32_t* __restrict__ tile_base,
                                    uint32_t total, uint32_t* __restrict__ contig_first) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_contigs) contig_first[c] = tile_base[contigs[c].tile0];
    else if (c == n_contigs) contig_first[c] = total;
}

int launch_contig_first(const Contig* contigs, uint32_t n_contigs, const uint32_t* tile_base, uint32_t total, uint32_t* contig_first,
                        cudaStream_t st) {
    contig_first_kernel<<<(n_contigs + 1 + 255) / 256, 256, 0, st>>>(contigs, n_contigs, tile_base, total, contig_first);
    return 1;
}

int s3_vote_threads() { return kSMs * 128; }

// tables in shared memory when `tsize` slots per thread leave room for at least one warp per CTA
int launch_s3_vote(const S3Scratch& sc, int e, uint32_t* tables, uint32_t tsize, uint8_t* peak_filter, cudaStream_t st) {
    const size_t budget = 192 << 10;
    int warps = (int)(budget / ((size_t)tsize * 4 * 32));
    if (warps >= 1) {
        warps = warps > 4 ? 4 : warps;
        size_t smem = (size_t)warps * 32 * tsize * 4;
        auto kern = e == 3 ? s3_vote_kernel<true, 3> : s3_vote_kernel<true, 0>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return -1;
        int ctas = (int)((220u << 10) / smem);
        kern<<<kSMs * (ctas < 1 ? 1 : ctas), warps * 32, smem, st>>>(sc.arena, sc.queue, sc.queue_count, sc.queue_cap, e, tables, tsize, peak_filter);
        return 1;
    }
    if (e == 3) s3_vote_kernel<false, 3><<<kSMs, 128, 0, st>>>(sc.arena, sc.queue, sc.queue_count, sc.queue_cap, e, tables, tsize, peak_filter);
    else s3_vote_kernel<false, 0><<<kSMs, 128, 0, st>>>(sc.arena, sc.queue, sc.queue_count, sc.queue_cap, e, tables, tsize, peak_filter);
    return 1;
}

int launch_s3(const uint8_t* fq1, const uint64_t* s1, const uint64_t* e1, uint64_t nrec1, const uint8_t* fq2,
              const uint64_t* s2, const uint64_t* e2, uint64_t nrec2, uint64_t tail_start, uint64_t tail_len,
              uint64_t first, uint64_t count, const uint32_t* sample_bits, uint64_t ordinal_base, const HashP& hp,
              const uint32_t* prefilter,
              const uint32_t* peak_kmer, const int32_t* loci, uint8_t* peak_filter, S3Scratch scratch, int grid_blocks,
              unsigned long long* n_sampled, int* err, cudaStream_t st) {
    if (count == 0 || nrec1 == 0) return 0;
    const size_t may_smem = ((size_t)kS3Warps * (kMaySlots / 2) + (scratch.n_contigs + 1 <= (uint32_t)kCfShared ? scratch.n_contigs + 1 : 0)) * sizeof(uint32_t);
#define LHGT_S3(EE)                                                                                                   \
    if (cudaFuncSetAttribute(s3_pairs_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)may_smem) != cudaSuccess) return -1; \
    s3_pairs_kernel<EE><<<grid_blocks, kS3Warps * 32, may_smem, st>>>(fq1, s1, e1, nrec1, fq2, s2, e2, nrec2, tail_start,    \
                                                               tail_len, first, count, sample_bits, ordinal_base, hp, prefilter,   \
                                                               peak_kmer, loci, peak_filter, scratch, n_sampled, err)
    switch (hp.e) {
        case 1: LHGT_S3(1); break;
        case 2: LHGT_S3(2); break;
        case 3: LHGT_S3(3); break;
        case 4: LHGT_S3(4); break;
        default: LHGT_S3(0); break;
    }
#undef LHGT_S3
    return 1;
}

// ------------------------------------------------------------------------------------------------
// sampling decisions (E:1037-1044 / E:413-419 with random_array reduced to its integer part, see lhgt_set_sampling)
// ------------------------------------------------------------------------------------------------
__global__ void sample_bits_kernel(const uint32_t* __restrict__ m, uint64_t n, uint32_t m_star, uint32_t* __restrict__ bits,
                                   uint64_t words) {
    int lane = threadIdx.x & 31;
    uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t w = warp; w < words; w += nwarps) {
        uint64_t o = w * 32 + lane;
        uint32_t word = __ballot_sync(kFull, o < n && m[o] < m_star);
        if (lane == 0) bits[w] = word;
    }
}

int launch_sample_bits(const uint32_t* m, uint64_t n, uint32_t m_star, uint32_t* bits, uint64_t words, cudaStream_t st) {
    sample_bits_kernel<<<kSMs * 8, 256, 0, st>>>(m, n, m_star, bits, words);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// OUT: the kept peaks (peak_filter >= 1, E:526) in id order, compacted on the device so that only they cross PCIe
// ------------------------------------------------------------------------------------------------
constexpr int kKeepBlock = 1024;
__global__ void __launch_bounds__(256) peaks_count_kernel(const uint8_t* __restrict__ filter, uint64_t n, uint32_t* __restrict__ block_cnt) {
    __shared__ uint32_t sm[33];
    uint64_t base = (uint64_t)blockIdx.x * kKeepBlock + threadIdx.x * 4;
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) c += base + q < n && filter[base + q] != 0;
    uint32_t total;
    block_exclusive_scan(c, &total, sm);
    if (threadIdx.x == 0) block_cnt[blockIdx.x] = total;
}

__global__ void __launch_bounds__(256) peaks_write_kernel(const uint8_t* __restrict__ filter, const int32_t* __restrict__ loci, uint64_t n,
                                                          const uint32_t* __restrict__ block_base, int32_t* __restrict__ out) {
    __shared__ uint32_t sm[33];
    uint64_t base = (uint64_t)blockIdx.x * kKeepBlock + threadIdx.x * 4;
    bool keep[4];
    uint32_t c = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) { keep[q] = base + q < n && filter[base + q] != 0; c += keep[q]; }
    uint32_t total, at = block_exclusive_scan(c, &total, sm) + block_base[blockIdx.x];
#pragma unroll
    for (int q = 0; q < 4; ++q)
        if (keep[q]) { out[2 * (size_t)at] = loci[2 * (base + q)]; out[2 * (size_t)at + 1] = loci[2 * (base + q) + 1]; ++at; }
}

uint64_t peaks_keep_blocks(uint64_t n) { return (n + kKeepBlock - 1) / kKeepBlock; }

int launch_peaks_compact(const uint8_t* filter, const int32_t* loci, uint64_t n, uint32_t* block_cnt, uint32_t* block_base,
                         uint32_t* scan_tmp, int32_t* out, int phase, cudaStream_t st) {
    uint64_t blocks = peaks_keep_blocks(n);
    if (!blocks) return 0;
    if (phase == 0) {
        peaks_count_kernel<<<(unsigned)blocks, 256, 0, st>>>(filter, n, block_cnt);
        return 1 + launch_scan_exclusive(block_cnt, block_base, blocks, scan_tmp, st);
    }
    peaks_write_kernel<<<(unsigned)blocks, 256, 0, st>>>(filter, loci, n, block_base, out);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// table utilities
// ------------------------------------------------------------------------------------------------
__global__ void count_unpack_kernel(const uint32_t* __restrict__ count, uint64_t entries, HashP hp, uint8_t* __restrict__ out) {
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < entries; h += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t g = tbl_index((uint32_t)h, hp);
        out[h] = (count[g >> 4] >> ((g & 15u) * 2)) & 3u;
    }
}

// out[h - h0] = peak_kmer[tbl_index(h)] for h in [h0, h0 + n): the peak table in hash order (test hook)
__global__ void peak_unpack_kernel(const uint32_t* __restrict__ peak_kmer, uint64_t h0, uint64_t n, HashP hp, uint32_t* __restrict__ out) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = peak_kmer[tbl_index((uint32_t)(h0 + i), hp)];
}

int launch_peak_unpack(const uint32_t* peak_kmer, uint64_t h0, uint64_t n, const HashP& hp, uint32_t* out, cudaStream_t st) {
    peak_unpack_kernel<<<kSMs * 8, 256, 0, st>>>(peak_kmer, h0, n, hp, out);
    return 1;
}

int launch_count_unpack(const uint32_t* count, uint64_t entries, const HashP& hp, uint8_t* out, cudaStream_t st) {
    count_unpack_kernel<<<kSMs * 8, 256, 0, st>>>(count, entries, hp, out);
    return 1;
}

// field-wise min(3, a + b) on sixteen 2-bit counters per word
__device__ __forceinline__ uint32_t sat_add2(uint32_t a, uint32_t b) {
    const uint32_t lo = 0x55555555u;
    uint32_t a0 = a & lo, a1 = (a >> 1) & lo, b0 = b & lo, b1 = (b >> 1) & lo;
    uint32_t s0 = a0 ^ b0, c0 = a0 & b0;           // bit 0 of the sum, carry into bit 1
    uint32_t s1 = a1 ^ b1 ^ c0;
    uint32_t over = (a1 & b1) | (c0 & (a1 ^ b1));  // sum >= 4
    return ((s0 | over) & lo) | (((s1 | over) & lo) << 1);
}

__global__ void count_merge_kernel(uint32_t* __restrict__ count, const uint32_t* __restrict__ other, uint64_t words) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x)
        count[i] = sat_add2(count[i], other[i]);
}

// how many counters hold 0, 1, 2, 3 (the occupancy figures of src/count_diff_kmer.cpp:26-50: empty = [0], weak = [0]+[1]+[2])
__global__ void count_histogram_kernel(const uint32_t* __restrict__ count, uint64_t words, unsigned long long* __restrict__ out4) {
    unsigned long long n1 = 0, n2 = 0, n3 = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t w = count[i], lo = w & 0x55555555u, hi = (w >> 1) & 0x55555555u;
        n1 += __popc(lo & ~hi); n2 += __popc(~lo & hi); n3 += __popc(lo & hi);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) {
        n1 += __shfl_xor_sync(kFull, n1, d); n2 += __shfl_xor_sync(kFull, n2, d); n3 += __shfl_xor_sync(kFull, n3, d);
    }
    if ((threadIdx.x & 31) == 0) {
        if (n1) atomicAdd(out4 + 1, n1);
        if (n2) atomicAdd(out4 + 2, n2);
        if (n3) atomicAdd(out4 + 3, n3);
    }
}

int launch_count_histogram(const uint32_t* count, uint64_t words, unsigned long long* out4, cudaStream_t st) {
    count_histogram_kernel<<<kSMs * 8, 256, 0, st>>>(count, words, out4);
    return 1;
}

int launch_count_merge(uint32_t* count, const uint32_t* other, uint64_t words, cudaStream_t st) {
    count_merge_kernel<<<kSMs * 8, 256, 0, st>>>(count, other, words);
    return 1;
}

// The multi-GPU count exchange as ONE kernel over peer memory (NVLink loads and stores, no staging buffer): this rank
// owns slice `rank` of every rank's table; it reads that slice from all `world` tables, combines the 2-bit fields with
// min(3, sum) and stores the result back into all of them.  Ranks touch disjoint slices, so the only synchronisation
// is a barrier before (every table final) and after (every slice written everywhere).
__global__ void __launch_bounds__(256) count_exchange_kernel(PeerTables pt, int world, uint64_t lo /* first uint4 of this rank's slice */,
                                                             uint64_t slice_vec /* uint4s in it */) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < slice_vec; i += (uint64_t)gridDim.x * blockDim.x) {
        uint4 v[kMaxPeers];
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < world) v[r] = reinterpret_cast<const uint4*>(pt.table[r])[lo + i];
        uint4 acc = v[0];
#pragma unroll
        for (int r = 1; r < kMaxPeers; ++r)
            if (r < world) {
                acc.x = sat_add2(acc.x, v[r].x); acc.y = sat_add2(acc.y, v[r].y);
                acc.z = sat_add2(acc.z, v[r].z); acc.w = sat_add2(acc.w, v[r].w);
            }
#pragma unroll
        for (int r = 0; r < kMaxPeers; ++r)
            if (r < world) reinterpret_cast<uint4*>(pt.table[r])[lo + i] = acc;
    }
}

// slices are whole 16-byte vectors and differ by at most one vector when `world` does not divide the table
int launch_count_exchange(const PeerTables& pt, int world, int rank, uint64_t table_words, cudaStream_t st) {
    uint64_t vecs = table_words / 4, base = vecs / (uint64_t)world, extra = vecs % (uint64_t)world;
    uint64_t lo = (uint64_t)rank * base + ((uint64_t)rank < extra ? (uint64_t)rank : extra);
    uint64_t n = base + ((uint64_t)rank < extra ? 1 : 0);
    if (n == 0) return 0;
    count_exchange_kernel<<<kSMs * 8, 256, 0, st>>>(pt, world, lo, n);
    return 1;
}

}  // namespace lhgt
