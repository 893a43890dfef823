/* TEST INFRASTRUCTURE ONLY — scalar C restatement of LocalHGT's extract_ref hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library, and only
 * as the checker.  The product (localhgt_b200/csrc, liblhgt.so, extract_ref) never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this restatement against outputs of the
 * unmodified reference binary (oracle/_ref/extract_ref_z, built by oracle/Makefile from
 * /root/reference/src/extract_ref_normal_peak.cpp) that are committed under tests/golden/.
 *
 * "E:" below = /root/reference/src/extract_ref_normal_peak.cpp.  Semantics are those of `-t 1`.
 */
#ifndef LHGT_ORACLE_H
#define LHGT_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_ctx orc_ctx;

orc_ctx* orc_create(int k, int e);                       /* E:1359-1378 */
void     orc_destroy(orc_ctx* c);

/* glibc srandom/random TYPE_3 (what E:1386 srand / E:1199,1336 rand resolve to on Linux) */
void orc_srand(orc_ctx* c, unsigned seed);
int  orc_rand(orc_ctx* c);

void orc_random_coder(orc_ctx* c);                       /* E:1182-1222, draws from the ctx stream */
void orc_set_coder(orc_ctx* c, const int16_t* cc300);
const int16_t* orc_coder(const orc_ctx* c);
int  orc_load_coder_from_index(orc_ctx* c, const char* index_path);   /* E:1224-1242 */

/* hashes of every k-mer of s[0..n): out[(j*e)+i]; 0 when the k-mer holds an invalid byte
 * (E:786-813); valid[j] tells the two zeros apart.  Returns n-k+1 (or 0). */
long orc_hash_seq(const orc_ctx* c, const uint8_t* s, long n, uint32_t* out, uint8_t* valid);

int    orc_index_build(orc_ctx* c, const char* fasta, const char* index_path, const char* len_path); /* E:727-886 */
double orc_sample_ratio(const char* fq1, double sample_arg);          /* E:1392-1398, 1244-1270 */
void   orc_fill_random_array(orc_ctx* c, long n);                     /* E:1332-1340 (first n of 50M) */
long   orc_s1_count(orc_ctx* c, const char* fq, long byte_budget, double ratio);   /* E:981-1107 */
long   orc_s2_peaks(orc_ctx* c, const char* index_path, float hit_ratio, float match_ratio,
                    long max_peak);                                   /* E:888-979, 550-725, 239-301 */
long   orc_s3_pairs(orc_ctx* c, const char* fq1, const char* fq2, double ratio);   /* E:313-506, 91-202 */
int    orc_write_intervals(orc_ctx* c, const char* path);             /* E:515-548 */

/* main() at -t 1 (E:1342-1519).  Returns 0, or negative on I/O error. */
int orc_extract_ref(const char* fq1, const char* fq2, const char* fasta, const char* interval_path,
                    double hit_ratio, double match_ratio, int k, long max_peak, int e, unsigned seed,
                    double sample_arg, long* stats6);

/* state accessors for stage-level parity checks */
const uint8_t*  orc_count_table(const orc_ctx* c);    /* 2^k saturating counters 0..3 */
long            orc_n_peaks(const orc_ctx* c);
const int32_t*  orc_peak_loci(const orc_ctx* c);      /* 2 ints per peak: ref_index, pos */
const uint8_t*  orc_peak_filter(const orc_ctx* c);
const uint32_t* orc_peak_kmer(const orc_ctx* c);      /* 2^k entries */
long            orc_raw_peak_positions(const orc_ctx* c); /* flagged positions fed to add_peak */

#ifdef __cplusplus
}
#endif
#endif
