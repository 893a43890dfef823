/* TEST INFRASTRUCTURE ONLY — see lhgt_oracle.h.  Scalar, single-threaded, written for obviousness.
 * "E:" = /root/reference/src/extract_ref_normal_peak.cpp.  Quirk numbers Q1..Q15 = SURVEY.md Appendix A.
 */
#define _FILE_OFFSET_BITS 64
#define _POSIX_C_SOURCE 200809L
#include "lhgt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

enum { CC_SLOTS = 300, RAND_CAP = 50000000, WINDOW = 500, PEAK_W = 5, DIFF_MIN = 2,
       MIN_VOTES = 6, BUCKET = 50, REF_NEAR = 500, REF_GAP = 500, LEAST_DEPTH = 3 };

struct orc_ctx {
    int k, e;
    int16_t cc[CC_SLOTS];
    /* glibc TYPE_3 state: 31-word ring, front/rear cursors 3 apart */
    int rf, rr;
    int32_t ring[31];
    /* tables */
    uint64_t table_len;
    uint8_t* count;             /* 2^k */
    uint32_t* peak_kmer;        /* 2^k, lazily allocated */
    float* rnd; long n_rnd;
    int32_t* loci; long n_peaks, cap_peaks;
    uint8_t* filter;
    long raw_positions;
};

/* ---------- small file helpers ---------- */
typedef struct { uint8_t* p; long n; } blob;

static blob slurp(const char* path) {
    blob b = {NULL, -1};
    FILE* f = fopen(path, "rb");
    if (!f) return b;
    struct stat st;
    if (fstat(fileno(f), &st) != 0) { fclose(f); return b; }
    b.n = (long)st.st_size;
    b.p = (uint8_t*)malloc((size_t)b.n + 1);
    if (b.n && fread(b.p, 1, (size_t)b.n, f) != (size_t)b.n) { free(b.p); b.p = NULL; b.n = -1; }
    fclose(f);
    return b;
}

/* std::getline over a buffer: returns 0 at end; line = [*s, *s + *len) ; advances *pos */
static int next_line(const blob* b, long* pos, long* s, long* len) {
    if (*pos >= b->n) return 0;
    const uint8_t* nl = (const uint8_t*)memchr(b->p + *pos, '\n', (size_t)(b->n - *pos));
    *s = *pos;
    if (nl) { *len = (long)(nl - b->p) - *pos; *pos = (long)(nl - b->p) + 1; }
    else    { *len = b->n - *pos; *pos = b->n; }
    return 1;
}

/* ---------- glibc random_r TYPE_3 (x^31 + x^3 + 1 additive feedback) ---------- */
void orc_srand(orc_ctx* c, unsigned seed) {
    int32_t word = (int32_t)(seed ? seed : 1u);
    c->ring[0] = word;
    for (int i = 1; i < 31; i++) {
        long hi = word / 127773, lo = word % 127773;
        long w = 16807 * lo - 2836 * hi;
        if (w < 0) w += 2147483647;
        word = (int32_t)w;
        c->ring[i] = word;
    }
    c->rf = 3; c->rr = 0;
    for (int i = 0; i < 310; i++) (void)orc_rand(c);
}

int orc_rand(orc_ctx* c) {
    uint32_t v = (uint32_t)c->ring[c->rf] + (uint32_t)c->ring[c->rr];
    c->ring[c->rf] = (int32_t)v;
    if (++c->rf == 31) c->rf = 0;
    if (++c->rr == 31) c->rr = 0;
    return (int)(v >> 1);
}

/* ---------- coders ---------- */
/* E:1109-1154: three base->bit maps; anything that is not ACGTacgt is invalid (5). */
static int coder_bit(int which, uint8_t ch) {
    int b;
    switch (ch) {
        case 'A': case 'a': b = 0; break;
        case 'C': case 'c': b = 1; break;
        case 'G': case 'g': b = 2; break;
        case 'T': case 't': b = 3; break;
        default: return 5;
    }
    /*            A  C  G  T */
    static const int m[3][4] = {{1, 0, 0, 1},    /* coder0: A,T -> 1 */
                                {1, 1, 0, 0},    /* coder1: A,C -> 1 */
                                {1, 0, 1, 0}};   /* coder2: A,G -> 1 */
    return m[which][b];
}

/* E:1165-1180 */
static uint8_t complement(uint8_t ch) {
    switch (ch) {
        case 'A': case 'a': return 'T';
        case 'T': case 't': return 'A';
        case 'C': case 'c': return 'G';
        case 'G': case 'g': return 'C';
        default: return 0;
    }
}

/* E:1182-1222 */
void orc_random_coder(orc_ctx* c) {
    static const int16_t perms[6][3] = {{0,1,2},{0,2,1},{1,2,0},{1,0,2},{2,0,1},{2,1,0}};
    for (int i = 0; i < CC_SLOTS; i++) c->cc[i] = 100;
    int groups = c->e / 3 + 1;
    for (int j = 0; j < c->k; j++) {
        int16_t row[3 * 4 + 3];
        for (int g = 0; g < groups; g++) {
            int r = orc_rand(c) % 6;
            for (int w = 0; w < 3; w++) row[3 * g + w] = perms[r][w];
        }
        for (int i = 0; i < c->e; i++) c->cc[j * c->e + i] = row[i];
    }
}

void orc_set_coder(orc_ctx* c, const int16_t* cc300) { memcpy(c->cc, cc300, sizeof c->cc); }
const int16_t* orc_coder(const orc_ctx* c) { return c->cc; }

int orc_load_coder_from_index(orc_ctx* c, const char* index_path) {
    FILE* f = fopen(index_path, "rb");
    if (!f) return -1;
    uint32_t w[CC_SLOTS];
    size_t got = fread(w, 4, CC_SLOTS, f);
    fclose(f);
    if (got != CC_SLOTS) return -2;
    for (int i = 0; i < CC_SLOTS; i++) c->cc[i] = (int16_t)w[i];     /* E:1233 truncation */
    return 0;
}

/* ---------- the hash (E:786-813 == E:1052-1081 == E:430-452) ---------- */
static int hash_one(const orc_ctx* c, const uint8_t* s, int i, uint32_t* out) {
    const int k = c->k, e = c->e;
    uint32_t fwd = 0, rev = 0;
    for (int z = 0; z < k; z++) {
        int m = coder_bit(c->cc[z * e + i], s[z]);
        if (m == 5) return 0;
        int n = coder_bit(c->cc[(k - 1 - z) * e + i], complement(s[z]));
        fwd += (uint32_t)m << (k - 1 - z);
        rev += (uint32_t)n << z;
    }
    *out = fwd < rev ? fwd : rev;
    return 1;
}

long orc_hash_seq(const orc_ctx* c, const uint8_t* s, long n, uint32_t* out, uint8_t* valid) {
    long np = n - c->k + 1;
    if (np <= 0) return 0;
    for (long j = 0; j < np; j++) {
        int ok = 1;
        for (int i = 0; i < c->e; i++) {
            uint32_t h = 0;
            ok = hash_one(c, s + j, i, &h);
            out[j * c->e + i] = ok ? h : 0;
        }
        if (valid) valid[j] = (uint8_t)ok;
    }
    return np;
}

/* ---------- ctx ---------- */
orc_ctx* orc_create(int k, int e) {
    if (k < 2 || k > 32 || e < 1 || e > 10 || k * e > CC_SLOTS) return NULL;
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof *c);
    if (!c) return NULL;
    c->k = k; c->e = e;
    c->table_len = 1ull << k;
    c->count = (uint8_t*)calloc(c->table_len, 1);
    for (int i = 0; i < CC_SLOTS; i++) c->cc[i] = 100;
    orc_srand(c, 1);
    if (!c->count) { free(c); return NULL; }
    return c;
}

void orc_destroy(orc_ctx* c) {
    if (!c) return;
    free(c->count); free(c->peak_kmer); free(c->rnd); free(c->loci); free(c->filter); free(c);
}

/* ---------- names (E:303-311) ---------- */
static long read_id_len(const uint8_t* s, long n) {
    long m = n;
    for (long i = 0; i < m; i++) if (s[i] == '/') { m = i; break; }
    for (long i = 0; i < m; i++) if (s[i] == ' ') { m = i; break; }
    for (long i = 0; i < m; i++) if (s[i] == '\t') { m = i; break; }
    return m;
}

/* ---------- index build (E:727-886) ---------- */
static void emit_contig(orc_ctx* c, FILE* idx, FILE* lenf, const uint8_t* name, long name_len,
                        long ordinal, const uint8_t* seq, long len, long cumulative) {
    if (len <= c->k) return;                                   /* E:772, 836 */
    fwrite(name, 1, (size_t)name_len, lenf);
    fprintf(lenf, "\t%ld\t%ld\t%ld\n", ordinal, len, cumulative);
    uint32_t l32 = (uint32_t)len;
    fwrite(&l32, 4, 1, idx);
    long np = len - c->k + 1;
    uint32_t* h = (uint32_t*)malloc((size_t)np * c->e * 4);
    orc_hash_seq(c, seq, len, h, NULL);
    fwrite(h, 4, (size_t)np * c->e, idx);
    free(h);
}

int orc_index_build(orc_ctx* c, const char* fasta, const char* index_path, const char* len_path) {
    blob fa = slurp(fasta);
    if (fa.n < 0) return -1;
    FILE* idx = fopen(index_path, "wb");
    FILE* lenf = fopen(len_path, "wb");
    if (!idx || !lenf) { free(fa.p); return -2; }
    /* Q1: 300 x 4-byte writes starting at &cc[j] -> word j = cc[j] | cc[j+1] << 16 (E:755-757) */
    for (int j = 0; j < CC_SLOTS; j++) {
        uint32_t lo = (uint16_t)c->cc[j], hi = j + 1 < CC_SLOTS ? (uint16_t)c->cc[j + 1] : 0;
        uint32_t w = lo | hi << 16;
        fwrite(&w, 4, 1, idx);
    }
    uint8_t* seq = (uint8_t*)malloc((size_t)fa.n + 1);
    long seq_len = 0, cumulative = 0, ordinal = 0, pos = 0, s, n;
    const uint8_t* name = (const uint8_t*)"start"; long name_len = 5;   /* E:747 */
    while (next_line(&fa, &pos, &s, &n)) {
        if (n > 0 && fa.p[s] == '>') {
            cumulative += seq_len;
            emit_contig(c, idx, lenf, name, name_len, ordinal, seq, seq_len, cumulative);
            ordinal++;                                          /* E:825: counts every header */
            seq_len = 0;
            long idl = read_id_len(fa.p + s, n);                /* E:764: get_read_ID(line).substr(1) */
            name = fa.p + s + 1; name_len = idl > 0 ? idl - 1 : 0;
        } else {
            memcpy(seq + seq_len, fa.p + s, (size_t)n);
            seq_len += n;
        }
    }
    cumulative += seq_len;
    emit_contig(c, idx, lenf, name, name_len, ordinal, seq, seq_len, cumulative);   /* E:833-880 */
    free(seq); free(fa.p);
    fclose(idx); fclose(lenf);
    return 0;
}

/* ---------- sampling (E:1392-1398, 1244-1270, 1332-1340) ---------- */
double orc_sample_ratio(const char* fq1, double sample_arg) {
    if (sample_arg <= 1) return 100 * sample_arg;
    blob fq = slurp(fq1);
    if (fq.n < 0) return -1;
    long pos = 0, s, n, i = 0, total = 0;
    while (next_line(&fq, &pos, &s, &n)) { if (i % 4 == 1) total += n; i++; }
    free(fq.p);
    total *= 2;
    return 100 * sample_arg / (double)total;
}

void orc_fill_random_array(orc_ctx* c, long n) {
    if (n > RAND_CAP) n = RAND_CAP;
    free(c->rnd);
    c->rnd = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    c->n_rnd = n;
    for (long i = 0; i < n; i++) c->rnd[i] = (float)((orc_rand(c) % 100000) / 1000.0);
}

static int sampled(const orc_ctx* c, unsigned long ordinal, double ratio) {
    if (ratio >= 100) return 1;               /* every drawn value is <= 99.999 (E:1336) */
    long idx = (long)(ordinal % RAND_CAP);
    if (idx >= c->n_rnd) return -1;           /* caller did not draw enough values */
    return (double)c->rnd[idx] < ratio;
}

/* ---------- S1 (E:981-1107) ---------- */
long orc_s1_count(orc_ctx* c, const char* path, long byte_budget, double ratio) {
    blob fq = slurp(path);
    if (fq.n < 0) return -1;
    long pos = 0, s, n, add = 0, used = 0;
    unsigned long lines = 0;
    uint32_t* h = (uint32_t*)malloc(4u * 512 * (size_t)c->e);
    uint8_t* ok = (uint8_t*)malloc(512);
    while (next_line(&fq, &pos, &s, &n)) {
        if (add > byte_budget) break;                          /* Q15: budget is size(fq1) for both files */
        add += n + 1;
        if (lines % 4 == 1) {
            if (n > 500) { used = -5; break; }                 /* E:1004 stack arrays */
            int take = sampled(c, lines / 4, ratio);
            if (take < 0) { used = -6; break; }
            if (take) {
                used++;
                long np = orc_hash_seq(c, fq.p + s, n, h, ok);
                for (long j = 0; j < np; j++) {
                    if (!ok[j]) continue;
                    for (int i = 0; i < c->e; i++) {
                        uint8_t* slot = &c->count[h[j * c->e + i]];
                        if (*slot < LEAST_DEPTH) (*slot)++;    /* E:1082-1084 */
                    }
                }
            }
        }
        lines++;
    }
    free(h); free(ok); free(fq.p);
    return used;
}

/* ---------- S2 (E:888-979 driver, E:550-725 window scan, E:239-301 peak registry) ---------- */
static int register_peak(orc_ctx* c, int contig, int pos, const uint32_t* hashes, const uint8_t* hit,
                         int len, long max_peak) {
    long id = c->n_peaks;
    int merged = id > 0 && c->loci[2 * (id - 1)] == contig &&
                 pos / BUCKET == c->loci[2 * (id - 1) + 1] / BUCKET;          /* E:288-301 */
    long write_id = merged ? id - 1 : id;
    if (!merged) {
        if (id >= max_peak) return -4;                                        /* Q13 */
        if (id == c->cap_peaks) {
            c->cap_peaks = c->cap_peaks ? c->cap_peaks * 2 : 1024;
            c->loci = (int32_t*)realloc(c->loci, sizeof(int32_t) * 2 * (size_t)c->cap_peaks);
        }
        c->loci[2 * id] = contig; c->loci[2 * id + 1] = pos;
        c->n_peaks++;
    }
    if (pos <= len - c->k + 1)                                                /* E:247,262 (Q6) */
        for (int p = 0; p < c->e; p++) {
            long at = (long)c->e * pos + p;
            if (hit[at] > 0) c->peak_kmer[hashes[at]] = (uint32_t)write_id;   /* E:250-251,265-266 */
        }
    return 0;
}

static int scan_contig(orc_ctx* c, int contig, int len, const uint32_t* hashes, const uint8_t* hit,
                       int one_min, int three_min, long max_peak) {
    const int e = c->e, k = c->k;
    uint8_t* single = (uint8_t*)calloc((size_t)len, 1);
    uint8_t* trio   = (uint8_t*)calloc((size_t)len, 1);
    uint8_t* peak   = (uint8_t*)calloc((size_t)len, 1);
    int* iv = (int*)malloc(sizeof(int) * 2 * ((size_t)len / WINDOW + 2));    /* Q11: "large enough" */
    int n_iv = 0, one = 0, three = 0, in_run = 0, good = 0, start = 0, rc = 0;
    for (int j = 0; j < len; j++) {
        int full = 0;
        for (int p = 0; p < e; p++) full += hit[(long)e * j + p] == LEAST_DEPTH;   /* E:580 */
        single[j] = full > 0; trio[j] = full == e;
        one += single[j]; three += trio[j];
        if (j >= WINDOW) { one -= single[j - WINDOW]; three -= trio[j - WINDOW]; }  /* E:597-608 */
        good = one >= one_min && three >= three_min;
        if (!in_run && good) { start = j - 2 * WINDOW; if (start < 1) start = 1; in_run = 1; }
        if (in_run && !good) {
            int end = j + 2 * WINDOW; if (end > len) end = len;
            if (n_iv > 0 && start - iv[2 * n_iv - 1] < WINDOW) iv[2 * n_iv - 1] = end;
            else { iv[2 * n_iv] = start; iv[2 * n_iv + 1] = end; n_iv++; }
            in_run = 0;
        }
        if (j > 2 * k + 2 * PEAK_W) {                                              /* E:644-671, Q12 */
            int right = 0, left = 0;
            for (int n = 0; n < PEAK_W; n++) right += single[j - n];
            for (int m = k; m < 2 * k; m++) {
                if (m == k) for (int n = 0; n < PEAK_W; n++) left += single[j - PEAK_W - n];
                else left += single[j - 2 * PEAK_W + 1 - m] - single[j - m - PEAK_W + 1];
                int diff = left - right;
                if (diff >= DIFF_MIN) peak[j - m - PEAK_W] = 1;
                if (diff <= -DIFF_MIN) peak[j] = 1;
            }
        }
    }
    if (in_run && good) {                                                          /* E:675-686 */
        if (n_iv > 0 && start - iv[2 * n_iv - 1] < WINDOW) iv[2 * n_iv - 1] = len;
        else { iv[2 * n_iv] = start; iv[2 * n_iv + 1] = len; n_iv++; }
    }
    for (int i = 0; i < n_iv && rc == 0; i++)
        for (int j = iv[2 * i]; j < iv[2 * i + 1] && rc == 0; j++)
            if (peak[j]) { c->raw_positions++; rc = register_peak(c, contig, j, hashes, hit, len, max_peak); }
    free(single); free(trio); free(peak); free(iv);
    return rc;
}

long orc_s2_peaks(orc_ctx* c, const char* index_path, float hit_ratio, float match_ratio, long max_peak) {
    blob ix = slurp(index_path);
    if (ix.n < 0) return -1;
    if (!c->peak_kmer) c->peak_kmer = (uint32_t*)calloc(c->table_len, 4);
    int one_min = (int)(WINDOW * hit_ratio), three_min = (int)(WINDOW * match_ratio);   /* E:559-560, fp32 */
    long at = 4 * CC_SLOTS; int contig = 1, rc = 0;                                 /* E:1294,1297 */
    c->n_peaks = 0; c->raw_positions = 0;
    while (at + 4 <= ix.n && rc == 0) {
        uint32_t len; memcpy(&len, ix.p + at, 4); at += 4;
        long np = (long)len - c->k + 1, words = np * c->e;
        if (at + 4 * words > ix.n) { rc = -2; break; }
        uint32_t* hashes = (uint32_t*)calloc((size_t)len * c->e, 4);              /* Q5: zero tail */
        uint8_t* hit = (uint8_t*)calloc((size_t)len * c->e, 1);
        memcpy(hashes, ix.p + at, (size_t)words * 4); at += 4 * words;
        for (long j = 0; j < words; j++) hit[j] = hashes[j] ? c->count[hashes[j]] : 0;   /* E:936-941, Q4 */
        rc = scan_contig(c, contig, (int)len, hashes, hit, one_min, three_min, max_peak);
        free(hashes); free(hit);
        contig++;                                                                  /* Q2: record ordinal */
    }
    free(ix.p);
    free(c->filter);
    c->filter = (uint8_t*)calloc((size_t)(c->n_peaks > 0 ? c->n_peaks : 1), 1);
    return rc < 0 ? rc : c->n_peaks;
}

/* ---------- S3 (E:313-506 + Split_reads E:91-202) ---------- */
typedef struct { int contig, votes; uint32_t first_peak; } tally;
typedef struct { tally t[1024]; int n, base_hits; } pair_votes;

static tally* find_tally(pair_votes* v, int contig) {
    for (int i = 0; i < v->n; i++) if (v->t[i].contig == contig) return &v->t[i];
    return NULL;
}

static void vote_mate(const orc_ctx* c, pair_votes* v, const uint8_t* s, long n, uint32_t* h, uint8_t* ok) {
    long np = orc_hash_seq(c, s, n, h, ok);
    for (long j = 0; j < np; j++) {
        uint32_t sel_peak = 0; int sel_contig = 0, sel_votes = 0, any = 0;
        for (int i = 0; i < c->e && ok[j]; i++) {                                  /* E:118-147 */
            uint32_t pk = c->peak_kmer[h[j * c->e + i]];
            if (pk == 0) continue;                                                 /* Q10 */
            any = 1;
            int contig = c->loci[2 * pk];
            tally* t = find_tally(v, contig);
            if (t) { if (t->votes >= sel_votes) { sel_peak = pk; sel_contig = contig; sel_votes = t->votes; } }
            else if (sel_peak == 0) { sel_peak = pk; sel_contig = contig; sel_votes = 0; }
        }
        if (!any) continue;
        tally* t = find_tally(v, sel_contig);                                       /* E:149-158 */
        if (t) t->votes++;
        else { v->t[v->n].contig = sel_contig; v->t[v->n].votes = 1; v->t[v->n].first_peak = sel_peak; v->n++; }
        v->base_hits++;
    }
}

static void settle_pair(orc_ctx* c, const pair_votes* v) {                          /* E:161-202 */
    if (v->base_hits < MIN_VOTES) return;
    int largest = 0, second = 0, strong = 0;
    for (int i = 0; i < v->n; i++) {
        int n = v->t[i].votes;
        if (n < MIN_VOTES) continue;
        strong++;
        if (n >= largest) { second = largest; largest = n; }
        else if (n >= second) second = n;
    }
    if (strong < 2) return;
    for (int i = 0; i < v->n; i++) {
        int n = v->t[i].votes;
        if (n >= MIN_VOTES && (n == largest || n == second) && c->filter[v->t[i].first_peak] < 254)
            c->filter[v->t[i].first_peak]++;
    }
}

long orc_s3_pairs(orc_ctx* c, const char* fq1, const char* fq2, double ratio) {
    blob a = slurp(fq1), b = slurp(fq2);
    if (a.n < 0 || b.n < 0) { free(a.p); free(b.p); return -1; }
    long pa = 0, pb = 0, sa, na, sb = 0, nb = 0, add = 0, used = 0;
    int b_eof = 0, b_fail = 0;
    unsigned long lines = 0;
    uint32_t* h = (uint32_t*)malloc(4u * 512 * (size_t)c->e);
    uint8_t* ok = (uint8_t*)malloc(512);
    pair_votes* v = (pair_votes*)malloc(sizeof *v);
    while (next_line(&a, &pa, &sa, &na)) {
        /* std::getline(fq2): a failed read leaves the previous string in place */
        if (!b_fail && !b_eof) {
            if (pb >= b.n) { nb = 0; b_fail = b_eof = 1; }
            else { next_line(&b, &pb, &sb, &nb); if (pb >= b.n && b.p[b.n - 1] != '\n') b_eof = 1; }
        } else b_fail = 1;
        if (add > a.n) break;
        add += na + 1;
        if (lines == 0) {                                                           /* E:368-399 */
            long ia = read_id_len(a.p + sa, na), ib = read_id_len(b.p + sb, nb);
            if (ia != ib || memcmp(a.p + sa, b.p + sb, (size_t)ia) != 0) { used = -3; break; }
        }
        if (lines % 4 == 1) {
            if (na > 500 || nb > 500) { used = -5; break; }
            int take = sampled(c, lines / 4, ratio);
            if (take < 0) { used = -6; break; }
            if (take) {
                used++;
                v->n = 0; v->base_hits = 0;
                vote_mate(c, v, a.p + sa, na, h, ok);
                vote_mate(c, v, b.p + sb, nb, h, ok);
                settle_pair(c, v);
            }
        }
        lines++;
    }
    free(h); free(ok); free(v); free(a.p); free(b.p);
    return used;
}

/* ---------- OUT (E:515-548) ---------- */
int orc_write_intervals(orc_ctx* c, const char* path) {
    FILE* f = fopen(path, "wb");
    if (!f) return -1;
    int chr = 1, start = 1, end = 1;                                                /* Q9 */
    for (long i = 0; i < c->n_peaks; i++) {
        if (c->filter[i] < 1) continue;
        int contig = c->loci[2 * i], pos = c->loci[2 * i + 1];
        if (chr == contig && pos - REF_NEAR - end < REF_GAP) end = pos + REF_NEAR;
        else {
            fprintf(f, "%d\t%d\t%d\n", chr, start, end);
            chr = contig; start = pos - REF_NEAR; end = pos + REF_NEAR;
        }
    }
    fprintf(f, "%d\t%d\t%d\n", chr, start, end);
    fclose(f);
    return 0;
}

/* ---------- main() at -t 1 (E:1342-1519) ---------- */
static long file_bytes(const char* p) { struct stat st; return stat(p, &st) == 0 ? (long)st.st_size : -1; }

int orc_extract_ref(const char* fq1, const char* fq2, const char* fasta, const char* interval_path,
                    double hit_ratio, double match_ratio, int k, long max_peak, int e, unsigned seed,
                    double sample_arg, long* stats6) {
    orc_ctx* c = orc_create(k, e);
    if (!c) return -10;
    int rc = 0; long r;
    long st[6] = {0, 0, 0, 0, 0, 0};
    orc_srand(c, seed);                                                             /* E:1386 */
    double ratio = orc_sample_ratio(fq1, sample_arg);
    char index_path[4096], len_path[4096];
    snprintf(index_path, sizeof index_path, "%s.k%d.h%d.index.dat", fasta, k, e);   /* E:1401 */
    snprintf(len_path, sizeof len_path, "%s.genome.len.txt", fasta);
    if (file_bytes(index_path) < 0) {                                               /* E:1403-1410, Q3 */
        orc_random_coder(c);
        rc = orc_index_build(c, fasta, index_path, len_path);
    }
    if (rc == 0) rc = orc_load_coder_from_index(c, index_path);                     /* E:1417 */
    long budget = file_bytes(fq1);                                                  /* E:1419 */
    if (rc == 0 && ratio < 100) orc_fill_random_array(c, RAND_CAP);                 /* E:1422 */
    if (rc == 0) { r = orc_s1_count(c, fq1, budget, ratio); if (r < 0) rc = (int)r; else st[0] = r; }
    if (rc == 0) { r = orc_s1_count(c, fq2, budget, ratio); if (r < 0) rc = (int)r; else st[1] = r; }
    if (rc == 0) { r = orc_s2_peaks(c, index_path, (float)hit_ratio, (float)match_ratio, max_peak);
                   if (r < 0) rc = (int)r; else { st[2] = c->raw_positions; st[3] = r; } }
    if (rc == 0) { r = orc_s3_pairs(c, fq1, fq2, ratio); if (r < 0) rc = (int)r; else st[4] = r; }
    if (rc == 0) {
        for (long i = 0; i < c->n_peaks; i++) st[5] += c->filter[i] >= 1;
        rc = orc_write_intervals(c, interval_path);
    }
    if (stats6) memcpy(stats6, st, sizeof st);
    orc_destroy(c);
    return rc;
}

/* ---------- accessors ---------- */
const uint8_t*  orc_count_table(const orc_ctx* c) { return c->count; }
long            orc_n_peaks(const orc_ctx* c) { return c->n_peaks; }
const int32_t*  orc_peak_loci(const orc_ctx* c) { return c->loci; }
const uint8_t*  orc_peak_filter(const orc_ctx* c) { return c->filter; }
const uint32_t* orc_peak_kmer(const orc_ctx* c) { return c->peak_kmer; }
long            orc_raw_peak_positions(const orc_ctx* c) { return c->raw_positions; }
