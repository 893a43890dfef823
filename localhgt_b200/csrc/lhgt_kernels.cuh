// Internal interface between the host side (lhgt_api.cu) and the sm_100a kernels (lhgt_kernels.cu).
// Nothing here is part of the C ABI (include/lhgt.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lhgt {

constexpr int kMaxE = 10;
constexpr int kMaxReadLen = 500;
constexpr int kTile = 1024;            // reference positions per S2 / IB tile
constexpr int kTileWords = kTile / 32;
constexpr int kRandomArray = 50000000; // E:40
constexpr int kFilterLog2 = 28;        // S3 pre-filter: 2^28 bits = 32 MiB, L2-resident on B200
constexpr int kFilterWords = 1 << (kFilterLog2 - 5);

// Everything a kernel needs to hash: F_i = W2 ^ ((W0^W2)&m0[i]) ^ ((W1^W2)&m1[i]) and the mirrored
// form for the reverse complement.  m2 is implied (the three masks partition kmask).
struct HashP {
    int k, e;
    int shr;                           // 32 - k
    uint32_t kmask;
    uint32_t m0[kMaxE], m1[kMaxE];
    // Count-table layout: the table is stored leaf-major.  The leaf of hash h is its bit field
    // [leaf_lo, leaf_lo + leaf_bits); the remaining idx_bits = k - leaf_bits bits, closed up, index the counter inside
    // the leaf's contiguous slice:  g(h) = leaf << idx_bits | idx.  The field sits in the MIDDLE of the hash: a canonical
    // hash is min(forward, reverse-complement) and that choice is decided by the leading bases (top bits of the forward
    // code) and the trailing bases (top bits of the other, which are also the LOW bits of the forward code), so only
    // middle bits are uniform.  leaf_bits = 0 (all fields 0) is the identity.
    int leaf_bits, idx_bits, leaf_lo;
    uint32_t leaf_mask, lo_mask;
};

__host__ __device__ inline uint32_t tbl_leaf(uint32_t h, const HashP& hp) { return (h >> hp.leaf_lo) & hp.leaf_mask; }
__host__ __device__ inline uint32_t tbl_idx(uint32_t h, const HashP& hp) {
    return ((h >> (hp.leaf_lo + hp.leaf_bits)) << hp.leaf_lo) | (h & hp.lo_mask);
}
__host__ __device__ inline uint32_t tbl_index(uint32_t h, const HashP& hp) {
    return (tbl_leaf(h, hp) << hp.idx_bits) | tbl_idx(h, hp);
}

struct Tile { uint32_t contig; uint32_t j0; };

// Per indexed contig.
struct Contig {
    uint64_t hash_word;                // word offset of its first hash in the index image
    uint64_t seq_off;                  // byte offset in the compacted sequence buffer (IB only)
    uint32_t len;
    uint32_t tile0;                    // first tile
};

// S1, streamed form: hashes are partitioned by their leaf field in two steps -- 2^b1 streams written
// by the hashing kernel (low b1 bits of the leaf), each split into 2^b2 leaf streams (the other b2 bits) -- and every
// leaf stream is applied to its table slice in shared memory.  leaf_bits = b1 + b2.
constexpr int kMaxB1 = 6, kMaxB2 = 8;
#ifndef LHGT_BIN_WARPS                 // the LHGT_* macros exist for tools/sweep.sh (variants built with -D, timed side by side)
#define LHGT_BIN_WARPS 32              // profiles/r01w_sweep.txt: 8 warps x 4 CTAs/SM 13.1 ms per step, 16 x 2 11.7, 32 x 1 11.0 (longer
#endif                                 // runs per stream, fewer reservations and barriers per hash)
#ifndef LHGT_BIN_ROUND_CHUNKS
#define LHGT_BIN_ROUND_CHUNKS 4
#endif
#ifndef LHGT_BIN_CTAS
#define LHGT_BIN_CTAS 1
#endif
constexpr int kBinWarps = LHGT_BIN_WARPS;   // warps per CTA of s1_bin_kernel
constexpr int kCursorStride = 64;      // words between stream cursors: every CTA bumps every cursor every round, and atomics on one
                                       // cache line are served one per clock by one L2 slice (profiles/r01g) -- give each its own lines
constexpr int kBinRoundChunks = LHGT_BIN_ROUND_CHUNKS;   // 32-position chunks a warp hashes per round
constexpr int kBinCtas = LHGT_BIN_CTAS;     // resident CTAs per SM the kernel is built for

struct BinP {
    uint32_t* pool_a;                  // stream b (b < 2^b1) occupies pool_a[b * cap_a .. +cap_a)
    uint32_t* cursor_a;                // cursor_a[b * kCursorStride]: entries appended to stream b so far (may run past cap_a: the
                                       // surplus was applied directly)
    uint32_t* pool_b;                  // leaf stream l (l < 2^(b1+b2)) occupies pool_b[l * cap_b .. +cap_b)
    uint32_t* cursor_b;                // [2^(b1+b2)]
    uint32_t cap_a, cap_b;             // entries; cap_a is a multiple of 8
    uint32_t bcap;                     // shared-memory bucket entries per stream in s1_bin_kernel
    int round_chunks;                  // 32-position chunks a warp hashes per round (<= kBinRoundChunks)
    int b1, b2;
};

struct S3Scratch {                     // per resident warp
    uint32_t* cands;                   // [2*(kMaxReadLen)] * e peak ids, then as many contigs
    int32_t* tally;                    // [3 * 2*kMaxReadLen]
    size_t cands_stride, tally_stride; // elements per warp
    // direct-addressed vote table for pairs that touch many contigs (s3_vote_table): per warp [0] = epoch counter,
    // [1 + c] = epoch << 10 | votes of contig c, [1 + n_contigs + 1 + c] = first peak voted for contig c
    uint32_t* vote_table;              // nullptr: not available (too many contigs) -> linear search
    size_t vote_stride;                // 2 * (vote_contigs + 1) + 1
    uint32_t vote_contigs;
    // peak id -> contig: first peak id of every indexed contig, n_contigs + 1 entries (contig_of_peak)
    const uint32_t* contig_first;
    uint32_t n_contigs;
    // hand-over to s3_vote_kernel: candidate lists ({peak id, contig} per (listed position, hash)) bump-allocated in `arena`,
    // one {offset, listed positions} job per pair in `queue`; arena = nullptr: every warp votes itself
    uint2* arena; uint32_t arena_cap; uint32_t* arena_cursor;
    uint2* queue; uint32_t queue_cap; uint32_t* queue_count;
};

// ---- launchers (all asynchronous on `st`; return the number of kernels launched) ----
int launch_fastq_index(const uint8_t* fq, uint64_t n, uint32_t* tile_cnt, uint32_t* tile_base,
                       uint32_t* scan_tmp, uint64_t* rec_start, uint64_t* rec_end, uint64_t rec_cap,
                       int phase, cudaStream_t st);
uint64_t fastq_index_tiles(uint64_t n);
size_t scan_tmp_words(uint64_t n);
int launch_scan_exclusive(const uint32_t* in, uint32_t* out, uint64_t n, uint32_t* tmp, cudaStream_t st);
int launch_scan_exclusive64(const uint64_t* in, uint64_t* out, uint64_t n, uint64_t* tmp, cudaStream_t st);
// FASTA ingest, all on the device.  launch_fasta_headers appends the byte span of every header line (unordered; *count may
// exceed cap: enlarge and repeat).  With the spans sorted by position: phase 0 counts the sequence bytes per 16 KiB tile,
// scans (64-bit), and answers nspans + 1 "sequence bytes before header i / before the end of the file" queries; phase 1
// writes the sequence bytes back to back into `out`.  scan_tmp holds scan_tmp_words(tiles) 8-byte elements.
struct ByteSpan { uint64_t lo, hi; };  // a header line: bytes [lo, hi] (hi = its newline, or the last byte of the file)
int launch_fasta_headers(const uint8_t* fa, uint64_t n, ByteSpan* spans, uint32_t cap, unsigned long long* count, cudaStream_t st);
int launch_fasta_header_text(const uint8_t* fa, const ByteSpan* spans, uint32_t nspans, const uint64_t* off, uint8_t* packed,
                             cudaStream_t st);
int launch_fasta_compact(const uint8_t* fa, uint64_t n, const ByteSpan* spans, uint32_t nspans, uint64_t* tile_cnt,
                         uint64_t* tile_base, uint64_t* scan_tmp, uint8_t* out, uint64_t* pos_out, int phase, cudaStream_t st);
int launch_sum_lengths(const uint64_t* rec_start, const uint64_t* rec_end, uint64_t nrec,
                       unsigned long long* out, cudaStream_t st);

int launch_index_build(const uint8_t* seq, const Contig* contigs, const Tile* tiles, uint64_t ntiles,
                       const HashP& hp, uint32_t* image, uint8_t* valid_out, cudaStream_t st);

int launch_s1(const uint8_t* fq, const uint64_t* rec_start, const uint64_t* rec_end, uint64_t nrec,
              uint64_t budget, const uint32_t* sample_bits, uint64_t ordinal_base, const HashP& hp, uint32_t* count,
              unsigned long long* n_sampled, int* err, cudaStream_t st);

// phase 0: stream the hashes of records [rec_lo, rec_hi) (cursors must be zero on entry);
// phase 1: split the streams into leaf streams; phase 2: apply the leaf streams
int launch_s1_binned(const uint8_t* fq, const uint64_t* rec_start, const uint64_t* rec_end, uint64_t rec_lo,
                     uint64_t rec_hi, uint64_t budget, const uint32_t* sample_bits, uint64_t ordinal_base,
                     const HashP& hp, const BinP& bp, uint32_t* count, unsigned long long* n_sampled, int* err,
                     int phase, cudaStream_t st);
size_t s1_bin_smem_bytes(const BinP& bp);
int s1_leaf_max_log2();              // largest table slice (log2 counters) s1_leaf_kernel holds in shared memory

// S2.  gather: trio (exact) + hash-0 hits as the lower bound of single, for tiles [tile_begin, tile_end);
// mark: hot tiles and the list, in tile order, of the tiles the remaining passes must visit;
// single: `single` made exact on the needed tiles of [tile_begin, tile_end); the rest run over the needed tiles only
// (good / flagged / tile_new must be zero elsewhere: the caller clears them).
int launch_s2_gather(const uint32_t* image, const Contig* contigs, const Tile* tiles, uint64_t tile_begin,
                     uint64_t tile_end, const HashP& hp, const uint32_t* count, uint32_t* single,
                     uint32_t* trio, cudaStream_t st);
size_t s2_mark_scratch_words(uint64_t ntiles);
int launch_s2_mark(const Contig* contigs, const Tile* tiles, uint64_t ntiles, const uint32_t* trio, int three_min, uint8_t* hot,
                   uint32_t* need_list, uint32_t* n_need, uint32_t* scratch, cudaStream_t st);
int launch_s2_single(const uint32_t* image, const Contig* contigs, const Tile* tiles, const uint32_t* need_list, const uint32_t* n_need,
                     uint64_t tile_begin, uint64_t tile_end, const HashP& hp, const uint32_t* count, uint32_t* single, cudaStream_t st);
// [t_lo, t_hi): only the needed tiles inside that range are visited (multi-GPU: each rank its own block of tiles)
int launch_s2_good(const Contig* contigs, const Tile* tiles, const uint32_t* need_list, const uint32_t* n_need, uint64_t t_lo, uint64_t t_hi,
                   const uint32_t* single, const uint32_t* trio, int one_min, int three_min, uint32_t* good, cudaStream_t st);
int launch_s2_flag(const Contig* contigs, const Tile* tiles, uint64_t ntiles, const uint32_t* need_list, const uint32_t* n_need,
                   uint64_t t_lo, uint64_t t_hi, int k, const uint32_t* single, const uint32_t* good, uint32_t* flagged, cudaStream_t st);
int launch_s2_count_new(const Tile* tiles, const uint32_t* need_list, const uint32_t* n_need, uint64_t t_lo, uint64_t t_hi,
                        const uint32_t* flagged, uint32_t* tile_new, unsigned long long* flagged_total, cudaStream_t st);
// mode 0: write loci + scatter-max peak ids + pre-filter bits; mode 1: clear what mode 0 wrote
int launch_s2_register(const uint32_t* image, const Contig* contigs, const Tile* tiles, const uint32_t* need_list, const uint32_t* n_need,
                       uint64_t t_lo, uint64_t t_hi, const HashP& hp, const uint32_t* count, const uint32_t* flagged, const uint32_t* tile_base,
                       int32_t* loci, uint32_t loci_cap, uint32_t* peak_kmer, uint32_t* prefilter, int mode, cudaStream_t st);

// S2 gather through table slices (tables > 128 MiB, e <= 4): one chunk of at most 2^18 tiles per call -- records into
// s2_gs_buckets() regions of `cap` in `pool` (cursor: s2_gs_cursor_words() zeroed words), answered slice by slice into the e
// bit planes `sat` (plane_words apart, zeroed by the caller); then launch_s2_gather_combine derives single and trio (both exact).
int s2_gs_buckets();
int s2_gs_cursor_words();
int launch_s2_gather_sliced(const uint32_t* image, const Contig* contigs, const Tile* tiles, uint64_t tile_begin, uint64_t tile_end,
                            const HashP& hp, const uint32_t* count, uint32_t* sat, size_t plane_words, uint2* pool, uint32_t* cursor,
                            uint32_t cap, cudaStream_t st);
int launch_s2_gather_combine(const uint32_t* sat, size_t plane_words, int e, uint64_t tile_begin, uint64_t tile_end, uint32_t* single,
                             uint32_t* trio, cudaStream_t st);

// Registration through buckets, one chunk [it_lo, it_hi) of the needed-tile list: (hash, id) records are appended to
// s2_reg_buckets() regions of `cap` records each, the first half of them in pool_lo, the second in pool_hi (cursor: s2_reg_cursor_words() zeroed words) and applied bucket by
// bucket against L2-resident slices of the tables.  Whatever does not fit is applied directly: exact either way.
size_t s2_regemit_smem();
int s2_reg_buckets();
int s2_reg_cursor_words();
int launch_s2_register_bucketed(const uint32_t* image, const Contig* contigs, const Tile* tiles, const uint32_t* need_list,
                                const uint32_t* n_need, uint32_t it_lo, uint32_t it_hi, uint64_t t_lo, uint64_t t_hi, const HashP& hp, const uint32_t* count,
                                const uint32_t* flagged, const uint32_t* tile_base, int32_t* loci, uint32_t loci_cap, uint32_t* peak_kmer,
                                uint32_t* prefilter, uint2* pool_lo, uint2* pool_hi, uint32_t* cursor, uint32_t cap, cudaStream_t st);

// kept peaks (filter != 0) in id order: phase 0 counts per 1024-peak block and scans (block_cnt / block_base: peaks_keep_blocks(n)
// words, scan_tmp: scan_tmp_words of that), phase 1 writes (contig, position) pairs into `out`
uint64_t peaks_keep_blocks(uint64_t n);
int launch_peaks_compact(const uint8_t* filter, const int32_t* loci, uint64_t n, uint32_t* block_cnt, uint32_t* block_base,
                         uint32_t* scan_tmp, int32_t* out, int phase, cudaStream_t st);

int launch_s3(const uint8_t* fq1, const uint64_t* s1, const uint64_t* e1, uint64_t nrec1,
              const uint8_t* fq2, const uint64_t* s2, const uint64_t* e2, uint64_t nrec2,
              uint64_t mate2_tail_start, uint64_t mate2_tail_len,
              uint64_t first, uint64_t count, const uint32_t* sample_bits, uint64_t ordinal_base, const HashP& hp,
              const uint32_t* prefilter, const uint32_t* peak_kmer, const int32_t* loci,
              uint8_t* peak_filter, S3Scratch scratch, int grid_blocks, unsigned long long* n_sampled,
              int* err, cudaStream_t st);
int launch_contig_first(const Contig* contigs, uint32_t n_contigs, const uint32_t* tile_base, uint32_t total, uint32_t* contig_first,
                        cudaStream_t st);
// votes of the pairs queued by launch_s3: `tables` holds s3_vote_threads() * tsize zeroed words (left zeroed)
int s3_vote_threads();
int launch_s3_vote(const S3Scratch& sc, int e, uint32_t* tables, uint32_t tsize, uint8_t* peak_filter, cudaStream_t st);
int s3_grid_blocks(int device);
int s3_warps_per_block();

// bit o of `bits` := o < n && m[o] < m_star (the sampling decision of ordinal o); the other bits of the 50 M-bit array := 0
int launch_sample_bits(const uint32_t* m, uint64_t n, uint32_t m_star, uint32_t* bits, uint64_t words, cudaStream_t st);
int launch_peak_unpack(const uint32_t* peak_kmer, uint64_t h0, uint64_t n, const HashP& hp, uint32_t* out, cudaStream_t st);
int launch_count_unpack(const uint32_t* count, uint64_t entries, const HashP& hp, uint8_t* out, cudaStream_t st);
int launch_count_merge(uint32_t* count, const uint32_t* other, uint64_t words, cudaStream_t st);
// out4[1..3] += number of counters equal to 1, 2, 3 (out4 zeroed by the caller; [0] follows from the table size)
int launch_count_histogram(const uint32_t* count, uint64_t words, unsigned long long* out4, cudaStream_t st);

constexpr int kMaxPeers = 8;           // GPUs of one box
struct PeerTables { uint32_t* table[kMaxPeers]; };   // every rank's count table, own included, as mapped in this process
// table_words must divide by 4 * world; slice `rank` of every table := min(3, sum over ranks)
int launch_count_exchange(const PeerTables& pt, int world, int rank, uint64_t table_words, cudaStream_t st);

__host__ __device__ inline uint32_t prefilter_slot(uint32_t h) {   // h: the TABLE INDEX of the hash (tbl_index), like the peak table's
    // any function of it is exact here (the filter only gates the exact lookup); fold the high bits in
    return (h ^ (h >> kFilterLog2)) & ((1u << kFilterLog2) - 1u);
}

}  // namespace lhgt
