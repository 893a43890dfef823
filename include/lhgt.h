/* lhgt.h — C ABI of liblhgt.so: the B200-native k-mer screening stage of LocalHGT (`extract_ref`).
 *
 * The reference exposes this path only as a PROCESS boundary (scripts/pipeline.sh:35 execs
 * `extract_ref` with 12 positional arguments, src/extract_ref_normal_peak.cpp:1352-1364).  The
 * drop-in for that boundary is the `extract_ref` executable built from csrc/extract_ref_main.cpp,
 * whose whole body is lhgt_main().  Everything else below is the thin layer underneath it that a
 * ctypes / cgo / JNI binding can call stage by stage; each entry cites the reference
 * code it replaces ("E:" = src/extract_ref_normal_peak.cpp).
 *
 * Conventions: plain pointers and sizes; caller owns every buffer it passes; no exceptions cross
 * the boundary; return 0 (or a non-negative count) on success and a negative LHGT_E_* code on
 * failure, with a message in lhgt_last_error() (thread-local).  A context is bound to one CUDA
 * device and is not re-entrant.  Results follow the reference's `-t 1` semantics bit for bit
 * (SURVEY.md Appendix A); there is no CPU fallback: without a CUDA device every GPU entry fails
 * with LHGT_E_CUDA.
 */
#ifndef LHGT_H
#define LHGT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LHGT_ABI_VERSION 1
#define LHGT_CODER_SLOTS 300          /* short choose_coder[300], E:1186 */
#define LHGT_MAX_E 10                 /* int base_kmer[10], E:99-100 */
#define LHGT_MAX_READ_LEN 500         /* int reads_int[500], E:322-323, 1004-1005 */
#define LHGT_RANDOM_ARRAY 50000000    /* MAX_RANDOM_NUM, E:40 */

enum {
    LHGT_OK = 0,
    LHGT_E_ARG = -1,             /* bad argument (k, e, null pointer, ...) */
    LHGT_E_IO = -2,              /* file cannot be opened / read / written */
    LHGT_E_CUDA = -3,            /* CUDA runtime error or no device */
    LHGT_E_FORMAT = -4,          /* malformed index image / FASTQ */
    LHGT_E_TOO_MANY_PEAKS = -5,  /* E:272-274 prints and overruns; we stop (Q13) */
    LHGT_E_READ_TOO_LONG = -6,   /* E:322-323 would overflow its stack arrays */
    LHGT_E_UNPAIRED = -7,        /* first records of fq1/fq2 carry different read ids (E:368-399) */
    LHGT_E_STATE = -8,           /* stage called before its inputs exist */
    LHGT_E_NOMEM = -9
};

typedef struct lhgt_ctx lhgt_ctx;

int         lhgt_abi_version(void);
const char* lhgt_last_error(void);

/* ------------------------------------------------------------------ host-only helpers (no GPU) */

/* glibc srand(seed)/rand() stream (E:1386; consumed at E:1199 and E:1336): writes draws
 * skip .. skip+n-1 to out. */
int lhgt_rand_stream(unsigned seed, long skip, long n, int32_t* out);

/* random_coder (E:1182-1222) after srand(seed): fills cc[300]; returns the number of rand() draws
 * it consumed (k * (e/3 + 1)) so the caller can position the sampling stream (Q3). */
int lhgt_random_coder(unsigned seed, int k, int e, int16_t* cc);

/* The 300-word index header as the reference writes it (E:755-757, quirk Q1) and its inverse
 * (saved_random_coder, E:1224-1242). */
int lhgt_coder_to_header(const int16_t* cc, uint32_t* words300);
int lhgt_header_to_coder(const uint32_t* words300, int16_t* cc);

/* ------------------------------------------------------------------ context */

/* Allocates the 2^k saturating count table (2 bits per k-mer) on `device`.  E:1375-1378. */
int  lhgt_create(lhgt_ctx** out, int device, int k, int e);
void lhgt_destroy(lhgt_ctx* c);
int  lhgt_set_coder(lhgt_ctx* c, const int16_t* cc);
int  lhgt_get_coder(const lhgt_ctx* c, int16_t* cc);
/* Run every launch of this context on an existing CUDA stream (cudaStream_t as an integer);
 * 0 restores the context's own stream. */
int  lhgt_set_stream(lhgt_ctx* c, uintptr_t cuda_stream);
int  lhgt_sync(lhgt_ctx* c);

/* Test hook: canonical hashes of every k-mer of ascii[0..n), out[j*e+i], 0 for k-mers holding a
 * non-ACGTacgt byte; valid[j] (nullable) tells those zeros from a true zero hash.  Runs the same
 * device code as the index build (E:786-813). */
int lhgt_hash_seq(lhgt_ctx* c, const uint8_t* ascii, size_t n, uint32_t* out, uint8_t* valid);

/* ------------------------------------------------------------------ IB: index build (read_ref, E:727-886) */

/* Builds the index image in HBM from FASTA text held in host memory.  The image is byte-identical
 * to <ref>.k<k>.h<e>.index.dat.  genome.len.txt text is returned through lhgt_index_len_text(). */
int      lhgt_index_build(lhgt_ctx* c, const uint8_t* fasta, size_t n);
/* The same from FASTA text that already lives on this device (16-byte aligned; the caller keeps it alive during the
 * call).  The whole of read_ref's line loop (E:761-831) runs on the device either way: header lines are found,
 * sequence bytes compacted and contig lengths derived there, with 64-bit offsets (a 5 Gbp reference is a 5.06 GB file). */
int      lhgt_index_build_device(lhgt_ctx* c, const void* dev_fasta, size_t n);
/* Starts the host->device copy of FASTA text on the copy stream; a later lhgt_index_build with the SAME pointer and
 * size adopts it (see lhgt_reads_prefetch).  A context that was given a prefetch keeps its FASTA buffer between builds:
 * re-building a 60 GB image from 5 GB of FASTA per sample is cheaper than shipping the image over PCIe. */
int      lhgt_fasta_prefetch(lhgt_ctx* c, const uint8_t* fasta, size_t n);
uint64_t lhgt_index_bytes(const lhgt_ctx* c);
uint64_t lhgt_index_bases(const lhgt_ctx* c);            /* sum of indexed contig lengths */
long     lhgt_index_contigs(const lhgt_ctx* c);
int      lhgt_index_download(lhgt_ctx* c, uint8_t* dst, uint64_t cap);
int      lhgt_index_len_text(const lhgt_ctx* c, char* dst, size_t cap, size_t* n);
/* One record of the resident image (E:921-931: u32 len, then (len-k+1)*e hashes) by its ordinal among the indexed
 * contigs; dst = NULL returns the number of 32-bit words. */
long     lhgt_index_record(lhgt_ctx* c, long record, uint32_t* dst, uint64_t cap_words);
/* Images larger than one GPU (row e-S): the ranks of a box each keep ONE block of the image -- equal runs of 1024-position
 * tiles, block 0 with the header -- and build only that (lhgt_set_image_block before lhgt_index_build*; the FASTA is still
 * parsed whole, it is 1/12 of the image at e = 3).  Every stage that reads stored hashes (S2 gather / complete / register)
 * then accepts tile ranges inside the block only; the multi-GPU plan (localhgt_b200/multi.py) exchanges hit bits, flagged bits
 * and peak tables so that no rank ever needs another rank's hashes.  lhgt_index_download / lhgt_index_write_block move the
 * resident block; lhgt_index_block says where it sits in the file. */
int      lhgt_set_image_block(lhgt_ctx* c, int part, int parts);
int      lhgt_index_block(const lhgt_ctx* c, uint64_t* byte_offset, uint64_t* bytes, long* tile_begin, long* tile_end);
int      lhgt_index_write_block(lhgt_ctx* c, const char* index_path);
/* Makes an existing image HBM-resident (read_index's input, E:888-979) and adopts its coder
 * (E:1417).  lhgt_index_attach_device uses an image that already lives on this device. */
int      lhgt_index_upload(lhgt_ctx* c, const uint8_t* image, uint64_t n);
/* Overlap of the host->device copies with the stages (the reference reads its inputs from disk stage by stage,
 * E:1426-1507; here the bytes cross PCIe once, and may do so while an earlier stage computes).  A prefetch starts
 * the copy of a host buffer (pinned memory for a truly asynchronous copy) on the context's copy stream and
 * returns; the later lhgt_reads_upload / lhgt_index_upload call with the SAME pointer and size adopts the copy
 * instead of repeating it.  The host buffer must stay valid and unchanged until that call returns. */
int      lhgt_reads_prefetch(lhgt_ctx* c, int mate, const uint8_t* fq, uint64_t n);
/* Across samples: while sample i is screened (index resident), sample i+1's FASTQ images cross PCIe into the mates'
 * alternate buffers; lhgt_reads_upload of sample i+1 (same pointer and size) swaps buffers instead of copying.  Call after
 * both mates of sample i were uploaded.  Costs a second pair of image buffers in HBM. */
int      lhgt_reads_prefetch_next(lhgt_ctx* c, int mate, const uint8_t* fq, uint64_t n);
int      lhgt_index_prefetch(lhgt_ctx* c, const uint8_t* image, uint64_t n);
/* File-level forms (what main() does at E:1401-1417). */
int      lhgt_index_build_file(lhgt_ctx* c, const char* fasta_path, const char* index_path,
                               const char* len_path);
int      lhgt_index_load_file(lhgt_ctx* c, const char* index_path);

/* ------------------------------------------------------------------ reads */

/* Copies one FASTQ file image (mate 0 = fq1, 1 = fq2) from host memory to HBM and locates its
 * records there (newline scan).  lhgt_reads_attach_device does the same for bytes that are already
 * resident on this device (no copy; the caller keeps them alive).  The device buffer must be 16-byte aligned and extend
 * to the next 16-byte boundary at or beyond n: reads are staged in whole 16-byte granules. */
int      lhgt_reads_upload(lhgt_ctx* c, int mate, const uint8_t* fq, uint64_t n);
int      lhgt_reads_attach_device(lhgt_ctx* c, int mate, const void* dev_fq, uint64_t n);
/* The same from a file, streamed through a ring of three 64 MiB pinned staging buffers (disk read, host->device copy and
 * whatever the GPU is computing overlap; pinned host memory stays bounded by the ring whatever the file size).
 * lhgt_index_build_file and lhgt_index_load_file read their files the same way. */
int      lhgt_reads_upload_file(lhgt_ctx* c, int mate, const char* path);
long     lhgt_reads_records(const lhgt_ctx* c, int mate);
uint64_t lhgt_reads_seq_bases(const lhgt_ctx* c, int mate);   /* sum of sequence-line lengths */
uint64_t lhgt_reads_bytes(const lhgt_ctx* c, int mate);       /* size of the resident FASTQ image */

/* down_sam_ratio in percent (E:1392-1398, cal_sam_ratio E:1244-1270); needs fq1 uploaded when
 * sample_arg > 1. */
double lhgt_sample_ratio(lhgt_ctx* c, double sample_arg);
/* Fixes the sampling rule `random_array[ordinal % 50M] < ratio` (E:1037-1044, 413-419) where
 * random_array is drawn from srand(seed) after rand_skip earlier draws (get_random, E:1332-1340). */
int    lhgt_set_sampling(lhgt_ctx* c, double ratio_percent, unsigned seed, long rand_skip);
/* Multi-GPU: this context holds records [base, base+n) of the whole sample, so the sampling ordinal
 * of its record r is base + r (the reference's `lines/4`, E:1037-1044 / E:413-419, counted over the
 * whole file).  Call before lhgt_set_sampling.  Default 0. */
int    lhgt_set_ordinal_base(lhgt_ctx* c, uint64_t base);

/* ------------------------------------------------------------------ stages */

/* S1 (read_fastq, E:981-1107): saturating k-mer counts of the sampled reads of one file.  Records
 * whose sequence line starts beyond byte_budget are ignored (quirk Q15: main() passes size(fq1)
 * for both files, E:1419-1444).  Returns the number of sampled reads. */
long lhgt_s1_count(lhgt_ctx* c, int mate, uint64_t byte_budget);
/* How S1 addresses the table: 0 = automatic (tables larger than 64 MiB are counted through hash
 * streams: the hashes are partitioned by their low bits until each partition's table slice fits in
 * shared memory, where it is updated; smaller tables are probed directly), 1 = direct probes,
 * 2 = streams regardless of table size (needs k > 18, or LHGT_LEAF_LOG2 in the environment at
 * lhgt_create time).  The counts are identical either way. */
int  lhgt_set_s1_mode(lhgt_ctx* c, int mode);
/* S2 (read_index + slide_window + Peaks::add_peak, E:888-979, 550-725, 239-301).  Returns the
 * number of peaks.  [tile_begin, tile_end) restricts the table gather to a slice of the reference
 * (multi-GPU); pass 0, -1 for everything. */
long lhgt_s2_peaks(lhgt_ctx* c, float hit_ratio, float match_ratio, long max_peak);
/* S3 (Peaks::slide_reads + Split_reads, E:313-506, 91-202) over record pairs [first, first+count)
 * (count < 0: all).  Returns the number of sampled pairs. */
long lhgt_s3_pairs(lhgt_ctx* c, long first, long count);
/* OUT (count_filtered_peak, E:515-548): the text of <interval_file>. */
int  lhgt_intervals(lhgt_ctx* c, char* dst, size_t cap, size_t* n);
/* Clears the count table and the peak tables so the context can screen another sample against the
 * same index. */
int  lhgt_reset(lhgt_ctx* c);

/* ------------------------------------------------------------------ state (parity checks, multi-GPU) */

int  lhgt_count_table_copy(lhgt_ctx* c, uint8_t* dst /* 2^k bytes, one counter per byte */);
long lhgt_peaks_copy(lhgt_ctx* c, int32_t* loci /* 2 per peak */, uint8_t* filter /* 0/1 */, long cap);
long lhgt_flagged_positions(const lhgt_ctx* c);          /* positions fed to add_peak */
int  lhgt_peak_kmer_copy(lhgt_ctx* c, uint32_t* dst /* 2^k */);

/* S2 in its steps (lhgt_s2_peaks runs them all over the whole reference); tiles are runs of 1024 reference positions.
 *   gather   [tile_begin, tile_end): per position, trio = all e stored hashes saturated (E:573-595), found with a
 *            short-circuit AND (~1.1 table probes per position instead of e), and hash 0's hit as the lower bound of single
 *   mark     hot tiles (some window sum of trio reaches floor(500 * match_ratio), E:560,610) and the tiles within reach of
 *            one (two either side): the only ones where good windows, intervals or peaks can exist
 *   complete [tile_begin, tile_end): single = some hash saturated, made exact on the marked tiles
 *   finish   window sums, intervals, coverage-edge peaks, peak ids, registration in peak_kmer -- on the marked tiles
 * Multi-GPU: gather and complete run on a tile range per rank, with the trio/single bit arrays exchanged after each
 * (lhgt_dev_hit_bits); mark and finish run replicated. */
long     lhgt_s2_tiles(const lhgt_ctx* c);
int      lhgt_s2_gather(lhgt_ctx* c, long tile_begin, long tile_end);
int      lhgt_s2_mark(lhgt_ctx* c, float match_ratio);
int      lhgt_s2_complete(lhgt_ctx* c, long tile_begin, long tile_end);
int      lhgt_s2_finish(lhgt_ctx* c, float hit_ratio, float match_ratio, long max_peak, long* n_peaks);
long     lhgt_s2_needed_tiles(const lhgt_ctx* c);        /* marked tiles of the last finish */
/* finish = windows + ids + register, each of which the multi-GPU plan runs by tile block:
 *   windows  [tile_begin, tile_end): good windows, flagged positions, new peaks per tile (slide_window E:597-712, the opening
 *            rule of add_peak E:288-301); halos are re-evaluated locally, nothing is needed from the neighbours
 *   ids      exclusive scan of the per-tile new-peak counts over ALL tiles (all-gather lhgt_dev_tile_new first), buffers;
 *            flagged_total = flagged positions over all ranks (lhgt_s2_flagged_in_range summed), or < 0 for this context's
 *   register [tile_begin, tile_end): peak_loci + peak_kmer[hash] = max id (E:246-270) for the flagged positions of the range.
 *            With lhgt_s2_dense() the ranks then combine their peak tables and loci with an element-wise MAX (ids grow
 *            with position, so the maximum is the last writer of the sequential loop); otherwise they all-gather the
 *            flagged bits and every rank registers everything. */
/* The marked tiles are listed in tile order; lhgt_s2_need_range gives the tile range holding share `part` of `parts`
 * equal shares of them (the shares tile the reference): the block a rank takes for windows and register -- the marked tiles
 * cluster where the sample's genomes are, so equal blocks of ALL tiles would leave most ranks idle there. */
int      lhgt_s2_need_range(lhgt_ctx* c, int part, int parts, long* tile_begin, long* tile_end);
int      lhgt_s2_windows(lhgt_ctx* c, float hit_ratio, float match_ratio, long tile_begin, long tile_end);
long     lhgt_s2_flagged_in_range(lhgt_ctx* c);
int      lhgt_s2_ids(lhgt_ctx* c, long max_peak, long flagged_total, long* n_peaks);
int      lhgt_s2_register(lhgt_ctx* c, long tile_begin, long tile_end);
int      lhgt_s2_dense(const lhgt_ctx* c);
/* Occupancy of the count table (the diagnostic of src/count_diff_kmer.cpp:26-50): out4[v] = number of the
 * 2^k counters holding v.  Its "empty" figure is out4[0] / 2^k, its "weak" figure (out4[0] + out4[1] + out4[2]) / 2^k. */
int      lhgt_count_table_histogram(lhgt_ctx* c, uint64_t* out4);
/* Raw device pointers for the exchange steps (NCCL / peer loads run by the caller). */
void*    lhgt_dev_count_table(lhgt_ctx* c, uint64_t* bytes);      /* packed 2-bit counters */
/* 128 bytes per tile (1024 positions, one bit each), tile-major; *bytes = the allocation, which covers lhgt_s2_tiles() + 8
 * tiles so that equal per-rank tile blocks can be all-gathered in place. */
void*    lhgt_dev_hit_bits(lhgt_ctx* c, int which /*0 single, 1 trio*/, uint64_t* bytes);
void*    lhgt_dev_peak_filter(lhgt_ctx* c, uint64_t* bytes);
void*    lhgt_dev_tile_new(lhgt_ctx* c, uint64_t* bytes);         /* u32 per tile (+ 8 of padding), new peaks opened in it */
void*    lhgt_dev_flagged(lhgt_ctx* c, uint64_t* bytes);          /* flagged-position bits, laid out like the hit bits */
void*    lhgt_dev_peak_table(lhgt_ctx* c, uint64_t* bytes);       /* 2^k u32 peak ids, in the count table's entry order */
void*    lhgt_dev_loci(lhgt_ctx* c, uint64_t* bytes);             /* 2 x i32 per peak */
/* count := min(3, count + other) field-wise on packed tables (other: device pointer, same size). */
int      lhgt_count_merge(lhgt_ctx* c, const void* dev_other, uint64_t bytes, uint64_t word_offset);

/* The same exchange as ONE kernel over peer memory, for the ranks of one box (one process per GPU): every rank passes
 * the CUDA IPC handle of its table (64 bytes, lhgt_count_table_ipc) to the others (any transport), maps theirs once
 * with lhgt_peers_open (handles = world x 64 bytes in rank order; the own slot is ignored), and after S1 -- between two
 * barriers the caller provides -- lhgt_count_exchange_p2p makes slice `rank` of EVERY rank's table min(3, sum over
 * ranks) with NVLink loads and stores, no staging buffer.  Once all ranks have run it every table holds the total. */
int      lhgt_count_table_ipc(lhgt_ctx* c, void* handle64);
int      lhgt_peers_open(lhgt_ctx* c, int rank, int world, const void* handles);
int      lhgt_count_exchange_p2p(lhgt_ctx* c);

/* Deferred counters: with on != 0, lhgt_s1_count and lhgt_s3_pairs only enqueue their kernels and return 0; the sampled
 * read / pair counts and the read-too-long flags stay on the device until lhgt_deferred_counts copies them back
 * (out3 = sampled reads of mate 0, of mate 1, sampled pairs of S3) with the step's one synchronisation. */
int      lhgt_set_deferred(lhgt_ctx* c, int on);
int      lhgt_deferred_counts(lhgt_ctx* c, long* out3);

/* Device time spent in each stage since the previous lhgt_stage_ms call (read-and-clear), measured
 * with CUDA events on the context's stream:
 * [0] FASTQ record scan  [1] S1  [2] S2 gather  [3] S2 finish  [4] S3  [5] IB kernel. */
int  lhgt_stage_ms(const lhgt_ctx* c, float* ms6);
/* Same with n_stages slots: [6] S1 hash-stream kernel  [7] S1 stream-split kernel  [8] S1 leaf-apply kernel
 * ([1] holds their sum)  [9] peer-memory count exchange  [10] S2 peak registration (not in [3])  [11] S3 vote (in [4]). */
#define LHGT_STAGES 12
int  lhgt_stage_ms_ex(const lhgt_ctx* c, float* ms, int n_stages);
/* Kernels launched by this context since creation. */
long lhgt_launch_count(const lhgt_ctx* c);

/* ------------------------------------------------------------------ the whole program */

typedef struct lhgt_args {
    const char* fq1;            /* argv[1]  */
    const char* fq2;            /* argv[2]  */
    const char* fasta;          /* argv[3]  */
    const char* interval;       /* argv[4]  */
    double hit_ratio;           /* argv[5]  */
    double match_ratio;         /* argv[6]  */
    int    threads;             /* argv[7]: accepted; results are those of -t 1 */
    int    k;                   /* argv[8]  */
    long   max_peak;            /* argv[9]  */
    int    e;                   /* argv[10] */
    unsigned seed;              /* argv[11] */
    double sample;              /* argv[12] */
    int    device;              /* CUDA device ordinal */
    int    quiet;               /* suppress the stdout log */
} lhgt_args;

typedef struct lhgt_stats {
    long   reads_s1[2];         /* sampled reads per file in S1 */
    long   flagged_positions;   /* positions fed to add_peak */
    long   peaks;
    long   pairs_s3;
    long   kept_peaks;
    int    index_built;         /* 1 when the index was created in this run */
    double ratio_percent;
    double seconds[8];          /* wall: total, io_in, index, s1, s2, s3, out, (spare) */
} lhgt_stats;

/* main() of the reference (E:1342-1519) at -t 1. */
int lhgt_extract_ref(const lhgt_args* a, lhgt_stats* stats /* nullable */);
/* The CLI body: same 12 positional arguments, every numeric parsed like stod (E:1359-1371).
 * Returns the process exit status. */
int lhgt_main(int argc, char** argv);

/* ------------------------------------------------------------------ post-screen glue (host-side text; no GPU work)
 *
 * What scripts/pipeline.sh:36-37 runs right after extract_ref.  Buffers are caller-owned; pass dst = NULL to size. */

/* scripts/get_bed_file.py:8-23,46-53: every interval line `ref_index start end` becomes `name:start-end`, the name looked
 * up through <ref>.genome.len.txt (column 2 -> column 1; later lines win); start < 1 becomes 1; |end - start| < 50 is
 * dropped.  *extract_len (nullable) = the script's "extracted ref length".  A ref_index the table does not list is
 * LHGT_E_FORMAT (the script dies with KeyError there). */
int lhgt_bed_text(const char* interval_text, size_t n_interval, const char* len_text, size_t n_len,
                  char* dst, size_t cap, size_t* n, long* extract_len);
/* `samtools faidx -r <bed> <ref.fa>` (pipeline.sh:37): per region `>name:start-end` and the bases start..end (1-based,
 * inclusive, clipped at the end of the sequence), 60 per line.  Sequence names are the header text up to the first
 * white space.  PARITY UNPINNED: samtools is not available where this was written; the format follows its
 * documentation, not a run. */
int lhgt_regions_fasta(const uint8_t* fasta, size_t n_fasta, const char* bed_text, size_t n_bed,
                       char* dst, size_t cap, size_t* n);
/* Both on files: reads <fasta>.genome.len.txt and <interval>, writes <interval>.bed and, if out_fasta is not NULL, the
 * extracted reference. */
int lhgt_extract_regions_files(const char* fasta_path, const char* interval_path, const char* out_fasta,
                               long* extract_len);

#ifdef __cplusplus
}
#endif
#endif /* LHGT_H */
