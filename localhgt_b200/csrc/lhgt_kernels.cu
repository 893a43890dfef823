// sm_100a kernels of the LocalHGT k-mer screen.  Integer hashing + HBM gathers; no tensor cores.
// "E:" = reference src/extract_ref_normal_peak.cpp (cited for semantics only; nothing here is a
// translation of it — see DESIGN.md §4 for the data-parallel forms these kernels evaluate).
#include "lhgt_kernels.cuh"

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
// 2-bit saturating count table: entry h lives in bits [2*(h&15), +2) of word h>>4
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

// count[h] = min(3, count[h] + 1), exact under any interleaving (E:1082-1084 made race-free).
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

__global__ void __launch_bounds__(kScanBlock) scan_block_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                                uint64_t n, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t sm[33];
    uint64_t base = (uint64_t)blockIdx.x * kScanSpan + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems], sum = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { v[i] = base + i < n ? in[base + i] : 0u; sum += v[i]; }
    uint32_t total, ex = block_exclusive_scan(sum, &total, sm);
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) { if (base + i < n) out[base + i] = ex; ex += v[i]; }
    if (threadIdx.x == 0 && block_sums) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanBlock) scan_add_kernel(uint32_t* __restrict__ out, uint64_t n,
                                                              const uint32_t* __restrict__ block_prefix) {
    uint64_t base = (uint64_t)blockIdx.x * kScanSpan + (uint64_t)threadIdx.x * kScanItems;
    uint32_t add = block_prefix[blockIdx.x];
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) if (base + i < n) out[base + i] += add;
}

size_t scan_tmp_words(uint64_t n) {
    size_t words = 0;
    while (n > 1) { n = (n + kScanSpan - 1) / kScanSpan; words += 2 * n; }
    return words + 2;
}

int launch_scan_exclusive(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* tmp, cudaStream_t st) {
    if (n == 0) return 0;
    uint64_t blocks = (n + kScanSpan - 1) / kScanSpan;
    int launches = 0;
    if (blocks == 1) {
        scan_block_kernel<<<1, kScanBlock, 0, st>>>(in, out, n, nullptr);
        return 1;
    }
    uint32_t* sums = tmp;
    uint32_t* sums_scanned = tmp + blocks;
    scan_block_kernel<<<(unsigned)blocks, kScanBlock, 0, st>>>(in, out, n, sums);
    launches += 1;
    launches += launch_scan_exclusive(sums, sums_scanned, blocks, tmp + 2 * blocks, st);
    scan_add_kernel<<<(unsigned)blocks, kScanBlock, 0, st>>>(out, n, sums_scanned);
    return launches + 1;
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

// newline with global rank g ends line g: line 4r+1 is the sequence of record r
__global__ void __launch_bounds__(kFqThreads) fq_assign_kernel(const uint8_t* __restrict__ fq, uint64_t n,
                                                               const uint32_t* __restrict__ tile_base,
                                                               uint64_t* __restrict__ rec_start,
                                                               uint64_t* __restrict__ rec_end, uint64_t rec_cap) {
    __shared__ uint32_t sm[33];
    uint64_t tile0 = (uint64_t)blockIdx.x * kFqTile;
    uint64_t running = tile_base[blockIdx.x];
    for (int it = 0; it < kFqIter; ++it) {
        uint64_t off = tile0 + (uint64_t)it * kFqSub + (uint64_t)threadIdx.x * kFqChunk;
        uint32_t w[4] = {0, 0, 0, 0}, m[4], c = 0;
        if (off < n) load16(fq, off, n, w);
#pragma unroll
        for (int q = 0; q < 4; ++q) { m[q] = off < n ? __vcmpeq4(w[q], 0x0a0a0a0au) : 0u; c += __popc(m[q]) >> 3; }
        uint32_t total, ex = block_exclusive_scan(c, &total, sm);
        uint64_t g = running + ex;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t mm = m[q];
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

__global__ void sum_lengths_kernel(const uint64_t* __restrict__ s, const uint64_t* __restrict__ e, uint64_t n,
                                   unsigned long long* out) {
    unsigned long long acc = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        acc += e[i] - s[i];
#pragma unroll
    for (int d = 16; d; d >>= 1) acc += __shfl_xor_sync(kFull, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
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
constexpr int kS1Warps = 8, kS1Unroll = 4;

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
                h[i] = i < e ? le_hash(kw, hp, i) : 0u;
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
// S1, binned form.  A 2^k-entry table far larger than L2 makes every direct probe a
// 128-byte DRAM fill (profiles/r01_probe_bench_*: 43 G probes/s), while a table slice that fits L2
// sustains 290 G probes/s.  So counting runs in two phases:
//   A  s1_bin_kernel: hash the sampled reads and append every hash to one of 2^bin_log2 streams
//      chosen by its top bits (per-CTA shared-memory buckets, flushed as coalesced runs);
//   B  s1_apply_kernel, once per stream: the 64 MiB table slice the stream addresses stays L2-resident
//      while the stream is read back sequentially and applied with the same load + CAS update.
// Saturating increments commute, so the table is bit-identical to the direct form's.  A stream that
// would overflow its region (pathological, low-complexity input) applies the surplus directly.
// ------------------------------------------------------------------------------------------------
constexpr int kBinWarps = 8;

__device__ __forceinline__ void bump_direct(uint32_t* count, uint32_t h) {
    bump(count, h, ld_table(count + (h >> 4)));
}

template <int E>
__global__ void __launch_bounds__(kBinWarps * 32, 4) s1_bin_kernel(
    const uint8_t* __restrict__ fq, const uint64_t* __restrict__ rec_start, const uint64_t* __restrict__ rec_end,
    uint64_t rec_lo, uint64_t rec_hi, uint64_t budget, const uint32_t* __restrict__ sample_bits, uint64_t ordinal_base,
    HashP hp, BinP bp, uint32_t* __restrict__ count, unsigned long long* __restrict__ n_sampled, int* __restrict__ err) {
    extern __shared__ uint32_t dyn[];                         // two bucket sets: one fills while the other drains
    __shared__ uint32_t cnt[2][kMaxBins];
    __shared__ uint2 bnd[kMaxBins];                           // bucket b of a set = set base + [bnd[b].x, bnd[b].y)
    __shared__ uint8_t lut[256];
    const int e = E ? E : hp.e;
    const int nbins = 1 << bp.log2;
    const uint32_t set_entries = bp.boff[kMaxBins];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < kMaxBins) {
        bnd[threadIdx.x] = make_uint2(bp.boff[threadIdx.x], bp.boff[threadIdx.x + 1]);
        cnt[0][threadIdx.x] = 0; cnt[1][threadIdx.x] = 0;
    }
    fill_base_lut(lut);
    __syncthreads();
    // the two streams this warp drains: b and nbins-1-b together carry an equal share of the hashes (stream_share)
    int mine_b[2] = {warp < nbins ? warp : -1, nbins - 1 - warp >= kBinWarps ? nbins - 1 - warp : -1};
    unsigned long long mine = 0;
    const uint64_t stride = (uint64_t)gridDim.x * kBinWarps;
    auto next_sampled = [&](uint64_t q) {                     // warp-uniform
        while (q < rec_hi && !is_sampled(sample_bits, q + ordinal_base)) q += stride;
        return q;
    };
    // The candidate record (next sampled one of this warp) is always one step ahead: its offsets are loaded when the
    // current read is accepted, its first 64 bases after the current read's first chunk, so a warp never sits
    // through the start -> end -> bytes load chain between reads.
    uint64_t r = next_sampled(rec_lo + (uint64_t)blockIdx.x * kBinWarps + warp);
    uint64_t cs = 0, ce = 0;
    if (r < rec_hi) { cs = rec_start[r]; ce = rec_end[r]; }
    uint32_t pb0 = 0, pb1 = 0;
    bool pb_ready = false;
    const uint8_t* src = fq;
    int len = 0, w = 0, nch = 0;                              // current read: length, next word to pack, hash chunks
    uint32_t chn = 0;                                         // this lane's byte of word w, loaded one chunk ahead
    Planes prev{0, 0, 0, 0};
    bool have = false;
    int set = 0;                                              // the set this round fills; set^1 drains
    for (bool first_round = true;; first_round = false) {
        while (!have && r < rec_hi) {                         // turn the candidate into the current read
            uint64_t start = cs, len64 = ce - cs;
            uint32_t b0 = pb0, b1 = pb1;
            bool fetched = pb_ready;
            r = next_sampled(r + stride);
            if (r < rec_hi) { cs = rec_start[r]; ce = rec_end[r]; }
            pb_ready = false;
            if (start > budget) continue;                     // Q15
            if (len64 > (uint64_t)kMaxReadLen) { if (lane == 0) atomicExch(err, 1); continue; }
            ++mine;
            len = (int)len64;
            int np = len - hp.k + 1;
            if (np <= 0) continue;
            src = fq + start;
            if (!fetched) { b0 = lane < len ? src[lane] : 0u; b1 = 32 + lane < len ? src[32 + lane] : 0u; }
            nch = (np + 31) >> 5;
            prev = pack_word(b0, lut);
            chn = b1;
            w = 1;
            have = true;
        }
        // One barrier per round: behind it the set filled last round is complete and the set drained last round is free.
        bool any = __syncthreads_or(have);
        // reserve room in the global streams for what last round produced; the answers are picked up after this
        // round's hashing, so the atomics' round trip costs nothing
        uint32_t dn[2] = {0, 0}, dg[2] = {0, 0};
        if (!first_round) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                int b = mine_b[q];
                if (b < 0) continue;
                dn[q] = min(cnt[set ^ 1][b], bnd[b].y - bnd[b].x);
                if (dn[q] && lane == 0) dg[q] = atomicAdd(bp.cursor + b, dn[q]);
            }
        }
        if (any) {
            uint32_t* buckets = dyn + set * set_entries;
            uint32_t* fill = cnt[set];
#pragma unroll 1
            for (int it = 0; it < kS1Unroll && have; ++it) {  // chunk w-1 = words w-1 (prev) and w (cur)
                Planes cur = pack_word(chn, lut);
                int pn = (w + 1) * 32 + lane;
                chn = pn < len ? src[pn] : 0u;
                LeWin kw = le_window(prev, cur, lane, hp);
                if (kw.valid) {
#pragma unroll
                    for (int i = 0; i < (E ? E : kMaxE); ++i)
                        if (i < e) {
                            uint32_t h = le_hash(kw, hp, i);
                            uint32_t b = h >> bp.shift;
                            uint2 lim = bnd[b];
                            uint32_t slot = lim.x + atomicAdd(&fill[b], 1u);
                            if (slot < lim.y) buckets[slot] = h;
                            else bump_direct(count, h);       // bucket full: rare, exact either way
                        }
                }
                prev = cur;
                if (++w > nch) { have = false; }
                if (!pb_ready && r < rec_hi) {                // candidate's first two words, consumed a read later
                    uint64_t nlen = ce - cs;
                    const uint8_t* nsrc = fq + cs;
                    pb0 = (uint64_t)lane < nlen ? nsrc[lane] : 0u;
                    pb1 = (uint64_t)(32 + lane) < nlen ? nsrc[32 + lane] : 0u;
                    pb_ready = true;
                }
            }
        }
        if (!first_round) {                                   // drain last round's set: one coalesced run per stream
            const uint32_t* drain = dyn + (set ^ 1) * set_entries;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                int b = mine_b[q];
                if (b < 0) continue;
                uint32_t n = dn[q];
                if (n) {
                    uint32_t g = __shfl_sync(kFull, dg[q], 0);
                    uint32_t b0 = bnd[b].x;
                    uint32_t* dst = bp.pool + bp.off[b];
                    uint32_t cap = bp.off[b + 1] - bp.off[b];
                    if (g + n <= cap) {
                        dst += g;
                        for (uint32_t x = lane; x < n; x += 32) dst[x] = drain[b0 + x];
                    } else {
                        for (uint32_t x = lane; x < n; x += 32) {
                            uint32_t h = drain[b0 + x];
                            if (g + x < cap) dst[g + x] = h;
                            else bump_direct(count, h);       // stream region full
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) cnt[set ^ 1][b] = 0;
            }
        }
        if (!any) break;
        set ^= 1;
    }
    if (lane == 0 && mine) atomicAdd(n_sampled, mine);
}

// ---- TMA (cp.async.bulk) + mbarrier plumbing for the stream reader ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`; the stream is read once: evict-first in L2
__device__ __forceinline__ void bulk_load_evict_first(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Phase B.  A CTA walks tiles of 2048 hashes of one stream.  An elected thread keeps a ring of TMA bulk copies in
// flight (UBLKCP; "full" mbarriers count the bytes landed, "empty" mbarriers count the warps done with a slot), so
// the stream never occupies load scoreboards and every thread spends its own on the 8 table probes it issues per
// tile.  No CTA-wide barrier inside the loop.
constexpr int kApplyThreads = 256;
template <int PER, int STAGES> constexpr size_t apply_smem_bytes() {
    return (size_t)STAGES * kApplyThreads * PER * sizeof(uint32_t) + 2 * STAGES * sizeof(uint64_t);
}

// MODE 0: the product.  Other modes exist for tools/apply_bench.cu only (cost probes): 1 = probes without updates.
template <int PER, int STAGES, int MIN_CTAS, int MODE>
__global__ void __launch_bounds__(kApplyThreads, MIN_CTAS) s1_apply_kernel(const uint32_t* __restrict__ stream,
                                                                           const uint32_t* __restrict__ cursor, uint32_t cap,
                                                                           uint32_t* __restrict__ count) {
    constexpr int kTileN = kApplyThreads * PER;
    extern __shared__ __align__(128) uint32_t apply_smem[];
    uint32_t* tiles = apply_smem;                                              // [STAGES][kTileN]
    uint64_t* full = reinterpret_cast<uint64_t*>(apply_smem + STAGES * kTileN);
    uint64_t* empty = full + STAGES;
    const uint32_t n = min(*cursor, cap);
    const uint32_t ntiles = (n + kTileN - 1) / kTileN;
    if (blockIdx.x >= ntiles) return;
    const uint32_t mine = (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;    // tiles blockIdx.x + i * gridDim.x
    const int lane = threadIdx.x & 31;
    auto issue = [&](uint32_t j) {                                             // thread 0 only
        uint32_t stg = j % STAGES;
        if (j >= (uint32_t)STAGES) mbar_wait(&empty[stg], ((j / STAGES) - 1u) & 1u);   // previous tenant drained
        uint32_t first = (blockIdx.x + j * gridDim.x) * kTileN;
        uint32_t bytes = (min((uint32_t)kTileN, n - first) * 4u + 15u) & ~15u;  // regions are multiples of 8 entries
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
        mbar_wait(&full[stg], (i / STAGES) & 1u);
        uint32_t first = (blockIdx.x + i * gridDim.x) * kTileN;
        uint32_t valid = min((uint32_t)kTileN, n - first);
        const uint32_t* tile = tiles + stg * kTileN;
        uint32_t h[PER], seen[PER];
        bool ok[PER];
#pragma unroll
        for (int q = 0; q < PER; ++q) {
            uint32_t x = threadIdx.x + q * kApplyThreads;
            ok[q] = x < valid;
            h[q] = tile[x];
        }
#pragma unroll
        for (int q = 0; q < PER; ++q)
            if (ok[q]) seen[q] = ld_table(count + (h[q] >> 4));
        __syncwarp();                                                          // every lane's tile words are in registers
        if (lane == 0) mbar_arrive(&empty[stg]);
        if (MODE == 0) bump_batch<PER>(count, h, seen, ok);
        else if (MODE == 1) {
#pragma unroll
            for (int q = 0; q < PER; ++q) if (ok[q]) sink += seen[q];
        } else {                                                               // cost probes, NOT exact: tools/apply_bench.cu
#pragma unroll
            for (int q = 0; q < PER; ++q) {
                int sh = (h[q] & 15u) * 2;
                if (ok[q] && ((seen[q] >> sh) & 3u) < 3u) {
                    uint32_t* addr = count + (h[q] >> 4);
                    if (MODE == 2) sink += atomicAdd(addr, 1u << sh);          // returning add
                    else if (MODE == 3) atomicAdd(addr, 1u << sh);             // fire-and-forget (RED)
                    else if (MODE == 4) atomicOr(addr, 1u << sh);              // RED.OR
                    else if (MODE == 5) atomicCAS(addr, seen[q], seen[q] + (1u << sh));   // CAS, result dropped
                }
            }
        }
    }
    if (MODE != 0 && sink == 0x9e3779b9u) count[0] = sink;
}

constexpr int kApplyPer = 8, kApplyStages = 4, kApplyMinCtas = 4;
constexpr size_t kApplySmem = apply_smem_bytes<kApplyPer, kApplyStages>();

size_t s1_bin_smem_bytes(const BinP& bp) { return (size_t)2 * bp.boff[kMaxBins] * sizeof(uint32_t); }   // two bucket sets

template <int E>
static cudaError_t s1_bin_launch(const uint8_t* fq, const uint64_t* rs, const uint64_t* re, uint64_t lo, uint64_t hi, uint64_t budget,
                                 const uint32_t* sb, uint64_t ob, const HashP& hp, const BinP& bp, uint32_t* count,
                                 unsigned long long* ns, int* err, cudaStream_t st) {
    size_t smem = s1_bin_smem_bytes(bp);
    cudaError_t rc = cudaFuncSetAttribute(s1_bin_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (rc != cudaSuccess) return rc;
    uint64_t want = (hi - lo + kBinWarps - 1) / kBinWarps;
    unsigned grid = (unsigned)(want < (uint64_t)kSMs * 4 ? want : (uint64_t)kSMs * 4);
    s1_bin_kernel<E><<<grid, kBinWarps * 32, smem, st>>>(fq, rs, re, lo, hi, budget, sb, ob, hp, bp, count, ns, err);
    return cudaGetLastError();
}

int launch_s1_binned(const uint8_t* fq, const uint64_t* rec_start, const uint64_t* rec_end, uint64_t rec_lo, uint64_t rec_hi,
                     uint64_t budget, const uint32_t* sample_bits, uint64_t ordinal_base, const HashP& hp, const BinP& bp,
                     uint32_t* count, unsigned long long* n_sampled, int* err, int phase, cudaStream_t st) {
    if (rec_hi <= rec_lo) return 0;
    int nbins = 1 << bp.log2;
    if (phase == 1) {
        auto kern = s1_apply_kernel<kApplyPer, kApplyStages, kApplyMinCtas, 0>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kApplySmem) != cudaSuccess) return -1;
        for (int b = 0; b < nbins; ++b)
            kern<<<kSMs * kApplyMinCtas, kApplyThreads, kApplySmem, st>>>(bp.pool + bp.off[b], bp.cursor + b, bp.off[b + 1] - bp.off[b], count);
        return nbins;
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
// S2 (E:888-979 + E:550-725 + E:239-301) as five data-parallel passes over 1024-position tiles.
// Bit arrays are little-endian in bit order: tile t, local position x -> word t*32 + x/32, bit x%32.
// ------------------------------------------------------------------------------------------------

// pass a: gather the count table at the stored hashes; single = some hash saturated, trio = all (E:573-595)
template <int E>
__global__ void __launch_bounds__(256) s2_gather_kernel(const uint32_t* __restrict__ image, const Contig* __restrict__ contigs,
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
    uint32_t h[4][E ? E : kMaxE], w[4][E ? E : kMaxE];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        long j = (long)t.j0 + r * 256 + threadIdx.x;
#pragma unroll
        for (int i = 0; i < (E ? E : kMaxE); ++i)
            if (i < e) h[r][i] = j < np ? ld_stream(hashes + (size_t)j * e + i) : 0u;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < (E ? E : kMaxE); ++i)
            if (i < e) w[r][i] = h[r][i] ? ld_stream(count + (h[r][i] >> 4)) : 0u;   // stored 0 = no hit (Q4, E:936-941)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int full = 0;
#pragma unroll
        for (int i = 0; i < (E ? E : kMaxE); ++i)
            if (i < e) full += h[r][i] && ((w[r][i] >> ((h[r][i] & 15u) * 2)) & 3u) == 3u;
        uint32_t ws = __ballot_sync(kFull, full > 0);
        uint32_t wt = __ballot_sync(kFull, full == e);
        if (lane == 0) {
            size_t word = (size_t)tix * kTileWords + r * 8 + warp;
            single[word] = ws;
            trio[word] = wt;
        }
    }
}

// inclusive count of set bits in local bit positions [0, x] of `words` given per-word exclusive prefix `cum`
__device__ __forceinline__ int bits_upto(const uint32_t* words, const int* cum, int x) {
    if (x < 0) return 0;
    int q = x >> 5;
    return cum[q] + __popc(words[q] & (0xffffffffu >> (31 - (x & 31))));
}

__device__ __forceinline__ void prefix_words(const uint32_t* words, int* cum, int n) {
    // n <= 96: one warp's worth of serial work is cheaper than a scan here
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < n; ++i) { cum[i] = acc; acc += __popc(words[i]); }
        cum[n] = acc;
    }
}

// pass b: 500-wide window sums and the good-window flag (E:597-615)
__global__ void __launch_bounds__(256) s2_good_kernel(const Contig* __restrict__ contigs, const Tile* __restrict__ tiles,
                                                      const uint32_t* __restrict__ single, const uint32_t* __restrict__ trio,
                                                      int one_min, int three_min, uint32_t* __restrict__ good) {
    __shared__ uint32_t ws[2 * kTileWords], wt[2 * kTileWords];
    __shared__ int cs[2 * kTileWords + 1], ct[2 * kTileWords + 1];
    Tile t = tiles[blockIdx.x];
    Contig c = contigs[t.contig];
    bool has_prev = t.j0 > 0;
    if (threadIdx.x < 2 * kTileWords) {
        int q = threadIdx.x;
        bool take = q >= kTileWords || has_prev;
        size_t word = (size_t)blockIdx.x * kTileWords + q - kTileWords;
        ws[q] = take ? single[word] : 0u;
        wt[q] = take ? trio[word] : 0u;
    }
    __syncthreads();
    prefix_words(ws, cs, 2 * kTileWords);
    if (threadIdx.x == 32) {
        int acc = 0;
        for (int i = 0; i < 2 * kTileWords; ++i) { ct[i] = acc; acc += __popc(wt[i]); }
    }
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
        if (lane == 0) good[(size_t)blockIdx.x * kTileWords + r * 8 + warp] = wg;
    }
}

// pass c: interval membership (E:617-638, 675-686 collapse to "a good window within +-1000") and the
// coverage-edge peak test (E:640-671) in its closed form:
//   D(x) = sum single[x-4..x],  C(j) = D(j-5) - D(j-k-5) - D(j),  diff_t(j) = C(j) + D(j-k-5-t), t in [0,k)
//   peak[j] if some diff_t(j) <= -2;  peak[j-k-5-t] if diff_t(j) >= 2;  only for 2k+10 < j < len.
__global__ void __launch_bounds__(256) s2_flag_kernel(const Contig* __restrict__ contigs, const Tile* __restrict__ tiles,
                                                      uint64_t ntiles, int k, const uint32_t* __restrict__ single,
                                                      const uint32_t* __restrict__ good, uint32_t* __restrict__ flagged) {
    __shared__ uint32_t ws[3 * kTileWords + 1], wg[3 * kTileWords];
    __shared__ int cg[3 * kTileWords + 1];
    __shared__ signed char D[3 * kTile], C[3 * kTile];
    Tile t = tiles[blockIdx.x];
    Contig c = contigs[t.contig];
    bool has_prev = t.j0 > 0;
    bool has_next = blockIdx.x + 1 < ntiles && tiles[blockIdx.x + 1].contig == t.contig;
    if (threadIdx.x < 3 * kTileWords) {
        int q = threadIdx.x;
        bool take = (q >= kTileWords || has_prev) && (q < 2 * kTileWords || has_next);
        size_t word = (size_t)blockIdx.x * kTileWords + q - kTileWords;
        ws[q] = take ? single[word] : 0u;
        wg[q] = take ? good[word] : 0u;
    }
    if (threadIdx.x == 0) ws[3 * kTileWords] = 0;
    __syncthreads();
    prefix_words(wg, cg, 3 * kTileWords);
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
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
                if (cj != -100) {
                    int m = 100;
                    for (int tt = 0; tt < k; ++tt) m = min(m, (int)D[x - k - 5 - tt]);
                    pk = cj + m <= -2;
                }
                for (int tt = 0; tt < k && !pk; ++tt) pk = (int)C[x + k + 5 + tt] + dq >= 2;
                f = pk;
            }
        }
        uint32_t wf = __ballot_sync(kFull, f);
        if (lane == 0) flagged[(size_t)blockIdx.x * kTileWords + r * 8 + warp] = wf;
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

__global__ void __launch_bounds__(256) s2_count_new_kernel(const Contig* __restrict__ contigs, const Tile* __restrict__ tiles,
                                                           const uint32_t* __restrict__ flagged, uint32_t* __restrict__ tile_new,
                                                           unsigned long long* __restrict__ flagged_total) {
    __shared__ uint32_t n_new, n_flag;
    if (threadIdx.x == 0) { n_new = 0; n_flag = 0; }
    __syncthreads();
    Tile t = tiles[blockIdx.x];
    uint32_t mine = 0, mine_f = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int xl = r * 256 + threadIdx.x;
        uint32_t w = flagged[(size_t)blockIdx.x * kTileWords + (xl >> 5)];
        if ((w >> (xl & 31)) & 1u) {
            ++mine_f;
            mine += opens_peak(flagged, blockIdx.x, t.j0, (long)t.j0 + xl);
        }
    }
    if (mine) atomicAdd(&n_new, mine);
    if (mine_f) atomicAdd(&n_flag, mine_f);
    __syncthreads();
    if (threadIdx.x == 0) {
        tile_new[blockIdx.x] = n_new;
        if (n_flag) atomicAdd(flagged_total, (unsigned long long)n_flag);
    }
}

// Peak ids follow (contig, position) order = tile order: id = tile_base + (#openers at or before me) - 1.
// Every flagged position stamps its id on the k-mers it holds with count > 0; later ids win, i.e.
// peak_kmer[h] = max id (E:246-270 executed in order).  Id 0 is the reference's "none" (Q10).
template <int E>
__global__ void __launch_bounds__(256) s2_register_kernel(const uint32_t* __restrict__ image, const Contig* __restrict__ contigs,
                                                          const Tile* __restrict__ tiles, HashP hp,
                                                          const uint32_t* __restrict__ count, const uint32_t* __restrict__ flagged,
                                                          const uint32_t* __restrict__ tile_base, int32_t* __restrict__ loci,
                                                          uint32_t* __restrict__ peak_kmer, uint32_t* __restrict__ prefilter,
                                                          int mode) {
    __shared__ uint32_t opener[kTileWords];
    __shared__ int cum[kTileWords + 1];
    const int e = E ? E : hp.e;
    Tile t = tiles[blockIdx.x];
    Contig c = contigs[t.contig];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    bool fl[4], op[4];
    bool any = false;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int xl = r * 256 + threadIdx.x;
        uint32_t w = flagged[(size_t)blockIdx.x * kTileWords + (xl >> 5)];
        fl[r] = (w >> (xl & 31)) & 1u;
        op[r] = fl[r] && opens_peak(flagged, blockIdx.x, t.j0, (long)t.j0 + xl);
        uint32_t wo = __ballot_sync(kFull, op[r]);
        if (lane == 0) opener[r * 8 + warp] = wo;
        any |= fl[r];
    }
    if (!__syncthreads_or(any)) return;
    prefix_words(opener, cum, kTileWords);
    __syncthreads();
    long np = (long)c.len - hp.k + 1;
    uint32_t base = tile_base[blockIdx.x];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        if (!fl[r]) continue;
        int xl = r * 256 + threadIdx.x;
        long j = (long)t.j0 + xl;
        uint32_t id = base + (uint32_t)bits_upto(opener, cum, xl) - 1u;   // wraps to 0xffffffff only if no opener yet: impossible for a flagged bit
        if (op[r] && mode == 0) { loci[2 * (size_t)id] = (int32_t)t.contig + 1; loci[2 * (size_t)id + 1] = (int32_t)j; }
        if (j < np && id != 0u) {                                         // j = len-k+1 reads the zero tail (Q6)
            const uint32_t* hashes = image + c.hash_word + (size_t)j * e;
            for (int i = 0; i < e; ++i) {
                uint32_t h = hashes[i];
                if (!h) continue;
                uint32_t cnt = (count[h >> 4] >> ((h & 15u) * 2)) & 3u;
                if (!cnt) continue;                                        // E:250,265: hit > 0
                uint32_t slot = prefilter_slot(h);
                if (mode == 0) {
                    atomicMax(peak_kmer + h, id);
                    atomicOr(prefilter + (slot >> 5), 1u << (slot & 31));
                } else {
                    peak_kmer[h] = 0u;
                    prefilter[slot >> 5] = 0u;
                }
            }
        }
    }
}

int launch_s2_gather(const uint32_t* image, const Contig* contigs, const Tile* tiles, uint64_t tile_begin, uint64_t tile_end,
                     const HashP& hp, const uint32_t* count, uint32_t* single, uint32_t* trio, cudaStream_t st) {
    if (tile_end <= tile_begin) return 0;
    unsigned grid = (unsigned)(tile_end - tile_begin);
    switch (hp.e) {
        case 1: s2_gather_kernel<1><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
        case 2: s2_gather_kernel<2><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
        case 3: s2_gather_kernel<3><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
        case 4: s2_gather_kernel<4><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
        default: s2_gather_kernel<0><<<grid, 256, 0, st>>>(image, contigs, tiles, tile_begin, hp, count, single, trio); break;
    }
    return 1;
}

int launch_s2_good(const Contig* contigs, const Tile* tiles, uint64_t ntiles, const uint32_t* single, const uint32_t* trio,
                   int one_min, int three_min, uint32_t* good, cudaStream_t st) {
    if (!ntiles) return 0;
    s2_good_kernel<<<(unsigned)ntiles, 256, 0, st>>>(contigs, tiles, single, trio, one_min, three_min, good);
    return 1;
}

int launch_s2_flag(const Contig* contigs, const Tile* tiles, uint64_t ntiles, int k, const uint32_t* single,
                   const uint32_t* good, uint32_t* flagged, cudaStream_t st) {
    if (!ntiles) return 0;
    s2_flag_kernel<<<(unsigned)ntiles, 256, 0, st>>>(contigs, tiles, ntiles, k, single, good, flagged);
    return 1;
}

int launch_s2_count_new(const Contig* contigs, const Tile* tiles, uint64_t ntiles, const uint32_t* flagged,
                        uint32_t* tile_new, unsigned long long* flagged_total, cudaStream_t st) {
    if (!ntiles) return 0;
    s2_count_new_kernel<<<(unsigned)ntiles, 256, 0, st>>>(contigs, tiles, flagged, tile_new, flagged_total);
    return 1;
}

int launch_s2_register(const uint32_t* image, const Contig* contigs, const Tile* tiles, uint64_t ntiles, const HashP& hp,
                       const uint32_t* count, const uint32_t* flagged, const uint32_t* tile_base, int32_t* loci,
                       uint32_t* peak_kmer, uint32_t* prefilter, int mode, cudaStream_t st) {
    if (!ntiles) return 0;
    s2_register_kernel<0><<<(unsigned)ntiles, 256, 0, st>>>(image, contigs, tiles, hp, count, flagged, tile_base, loci,
                                                           peak_kmer, prefilter, mode);
    return 1;
}

// ------------------------------------------------------------------------------------------------
// S3: read-pair confirmation (E:419-499 + Split_reads E:109-202).  One warp per pair.  Lanes hash
// positions in parallel and test an L2-resident pre-filter; only k-mers that pass it touch the
// 2^k-entry peak table.  Positions that hold a peak k-mer are appended, in read order, to a per-warp
// list; the order-dependent vote (judge_base) then runs on lane 0 over that (usually empty) list.
// ------------------------------------------------------------------------------------------------
constexpr int kS3Warps = 8;
int s3_warps_per_block() { return kS3Warps; }
int s3_grid_blocks(int) { return kSMs * 4; }

template <int E>
__device__ __forceinline__ int s3_scan_mate(const uint8_t* __restrict__ src, int len, const uint8_t* lut, const HashP& hp,
                                            const uint32_t* __restrict__ prefilter, const uint32_t* __restrict__ peak_kmer,
                                            uint32_t* __restrict__ cands, int n_listed, int lane) {
    const int e = E ? E : hp.e;
    int np = len - hp.k + 1;
    if (np <= 0) return n_listed;
    int nch = (np + 31) >> 5;
    Planes prev = pack_word(lane < len ? src[lane] : 0u, lut);
    uint32_t c1 = 32 + lane < len ? src[32 + lane] : 0u;      // bytes run two words ahead of the hashing
    uint32_t c2 = 64 + lane < len ? src[64 + lane] : 0u;
    for (int w = 1; w <= nch; ++w) {                          // chunk w-1: positions 32(w-1) + lane, ascending
        Planes cur = pack_word(c1, lut);
        c1 = c2;
        int pn = (w + 2) * 32 + lane;
        c2 = pn < len ? src[pn] : 0u;
        LeWin kw = le_window(prev, cur, lane, hp);
        prev = cur;
        uint32_t h[E ? E : kMaxE], pk[E ? E : kMaxE], fw[E ? E : kMaxE];
#pragma unroll
        for (int i = 0; i < (E ? E : kMaxE); ++i)
            if (i < e) {
                h[i] = le_hash(kw, hp, i);
                uint32_t slot = prefilter_slot(h[i]);
                fw[i] = kw.valid ? __ldg(prefilter + (slot >> 5)) >> (slot & 31) : 0u;
            }
        bool any = false;
#pragma unroll
        for (int i = 0; i < (E ? E : kMaxE); ++i)
            if (i < e) {
                pk[i] = (fw[i] & 1u) ? __ldg(peak_kmer + h[i]) : 0u;
                any |= pk[i] != 0u;
            }
        uint32_t mask = __ballot_sync(kFull, any);
        if (mask) {
            if (any) {
                int slot = n_listed + __popc(mask & ((1u << lane) - 1u));
#pragma unroll
                for (int i = 0; i < (E ? E : kMaxE); ++i)
                    if (i < e) cands[(size_t)slot * e + i] = pk[i];
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

template <int E>
__global__ void __launch_bounds__(kS3Warps * 32, 4) s3_pairs_kernel(
    const uint8_t* __restrict__ fq1, const uint64_t* __restrict__ s1, const uint64_t* __restrict__ e1, uint64_t nrec1,
    const uint8_t* __restrict__ fq2, const uint64_t* __restrict__ s2, const uint64_t* __restrict__ e2, uint64_t nrec2,
    uint64_t tail_start, uint64_t tail_len, uint64_t first, uint64_t count, const uint32_t* __restrict__ sample_bits,
    uint64_t ordinal_base, HashP hp, const uint32_t* __restrict__ prefilter, const uint32_t* __restrict__ peak_kmer,
    const int32_t* __restrict__ loci, uint8_t* __restrict__ peak_filter, S3Scratch scratch,
    unsigned long long* __restrict__ n_sampled, int* __restrict__ err) {
    __shared__ uint8_t lut[256];
    const int e = E ? E : hp.e;
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    fill_base_lut(lut);
    __syncthreads();
    uint64_t gwarp = (uint64_t)blockIdx.x * kS3Warps + warp;
    uint32_t* cands = scratch.cands + gwarp * scratch.cands_stride;
    int32_t* tally = scratch.tally + gwarp * scratch.tally_stride;
    unsigned long long mine = 0;
    uint64_t stride = (uint64_t)gridDim.x * kS3Warps;
    uint64_t last = first + count < nrec1 ? first + count : nrec1;
    for (uint64_t r = first + gwarp; r < last; r += stride) {
        if (!is_sampled(sample_bits, r + ordinal_base)) continue;
        uint64_t a0 = s1[r], l1 = e1[r] - a0, b0, l2;
        if (r < nrec2) { b0 = s2[r]; l2 = e2[r] - b0; }
        else { b0 = tail_start; l2 = tail_len; }            // fq2 exhausted: std::getline leaves its last string
        if (l1 > (uint64_t)kMaxReadLen || l2 > (uint64_t)kMaxReadLen) { if (lane == 0) atomicExch(err, 1); continue; }
        ++mine;
        int n_listed = s3_scan_mate<E>(fq1 + a0, (int)l1, lut, hp, prefilter, peak_kmer, cands, 0, lane);
        n_listed = s3_scan_mate<E>(fq2 + b0, (int)l2, lut, hp, prefilter, peak_kmer, cands, n_listed, lane);
        if (n_listed >= 6) {                                 // base_hits >= MIN_BASE_NUM (E:496)
            __syncwarp();
            if (lane == 0) s3_vote(cands, n_listed, e, loci, tally, peak_filter);
            __syncwarp();
        }
    }
    if (lane == 0 && mine) atomicAdd(n_sampled, mine);
}

int launch_s3(const uint8_t* fq1, const uint64_t* s1, const uint64_t* e1, uint64_t nrec1, const uint8_t* fq2,
              const uint64_t* s2, const uint64_t* e2, uint64_t nrec2, uint64_t tail_start, uint64_t tail_len,
              uint64_t first, uint64_t count, const uint32_t* sample_bits, uint64_t ordinal_base, const HashP& hp,
              const uint32_t* prefilter,
              const uint32_t* peak_kmer, const int32_t* loci, uint8_t* peak_filter, S3Scratch scratch, int grid_blocks,
              unsigned long long* n_sampled, int* err, cudaStream_t st) {
    if (count == 0 || nrec1 == 0) return 0;
#define LHGT_S3(EE)                                                                                                   \
    s3_pairs_kernel<EE><<<grid_blocks, kS3Warps * 32, 0, st>>>(fq1, s1, e1, nrec1, fq2, s2, e2, nrec2, tail_start,    \
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
// table utilities
// ------------------------------------------------------------------------------------------------
__global__ void count_unpack_kernel(const uint32_t* __restrict__ count, uint64_t entries, uint8_t* __restrict__ out) {
    for (uint64_t h = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; h < entries; h += (uint64_t)gridDim.x * blockDim.x)
        out[h] = (count[h >> 4] >> ((h & 15u) * 2)) & 3u;
}

int launch_count_unpack(const uint32_t* count, uint64_t entries, uint8_t* out, cudaStream_t st) {
    count_unpack_kernel<<<kSMs * 8, 256, 0, st>>>(count, entries, out);
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

int launch_count_merge(uint32_t* count, const uint32_t* other, uint64_t words, cudaStream_t st) {
    count_merge_kernel<<<kSMs * 8, 256, 0, st>>>(count, other, words);
    return 1;
}

}  // namespace lhgt
