/*
 * TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it.
 *
 * Plain-C restatement of PEPPER's two pileup-summary encoders, written from the
 * behaviour of the reference (file:line cited per function; paths relative to
 * the reference repo).  It is pinned against the reference's own C++ (compiled
 * unmodified into oracle/_ref by oracle/Makefile) by tests/test_oracle_*.py and
 * against the golden fixtures under tests/golden/ generated from that build.
 *
 *   port_variant_*  <->  pepper_variant/modules/cpp/region_summary.cpp
 *   port_polish_*   <->  pepper/modules/src/pileup_summary/summary_generator.cpp
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/pepper_b200.h"

static const char NT16[] = "=ACMGRSVTWYHKDBN";

static inline int seq_code(const pb_reads_t *R, int64_t nib) {
    uint8_t b = R->seq[nib >> 1];
    return (nib & 1) ? (b & 15) : (b >> 4);
}

/* region_summary.cpp:193-199 check_ref_base (after toupper, :203) */
static inline int valid_ref(char c) {
    return c == 'A' || c == 'C' || c == 'G' || c == 'T' || c == 'a' || c == 'c' || c == 'g' || c == 't';
}
static inline char up(char c) { return (c >= 'a' && c <= 'z') ? (char) (c - 32) : c; }

/* region_summary.cpp:201-230 get_feature_index */
static int feature_index(char ref_base, char base, int is_reverse) {
    if (!valid_ref(ref_base)) return -1;
    int start = is_reverse ? 18 : 7;
    switch (up(base)) {
        case 'A': return start + 1;
        case 'C': return start + 2;
        case 'G': return start + 3;
        case 'T': return start + 4;
        case 'I': return start + 5;
        case 'D': return start + 6;
        default:  return start + 7;
    }
}
/* region_summary.cpp:165-172 get_reference_feature_value */
static int ref_value(char b) {
    switch (up(b)) { case 'A': return 1; case 'C': return 2; case 'G': return 3; case 'T': return 4; default: return 5; }
}

/* ---- per-position allele tallies (the four std::map / std::set of
 *      region_summary.cpp:590-593) as small dynamic arrays ---- */
typedef struct { char key[PB_ALLELE_STRIDE]; int total, fwd, rev; } allele_t;
typedef struct { allele_t *a; int n, cap; } allele_list_t;

static void tally(allele_list_t *L, const char *key, int is_reverse) {
    for (int i = 0; i < L->n; i++)
        if (strcmp(L->a[i].key, key) == 0) {
            L->a[i].total++;
            if (is_reverse) L->a[i].rev++; else L->a[i].fwd++;
            return;
        }
    if (L->n == L->cap) { L->cap = L->cap ? 2 * L->cap : 4; L->a = (allele_t *) realloc(L->a, sizeof(allele_t) * L->cap); }
    allele_t *e = &L->a[L->n++];
    memset(e, 0, sizeof(*e));
    strncpy(e->key, key, PB_ALLELE_STRIDE - 1);
    e->total = 1; e->fwd = !is_reverse; e->rev = !!is_reverse;
}
static int allele_cmp(const void *x, const void *y) {   /* std::set<string> order */
    return strcmp(((const allele_t *) x)->key, ((const allele_t *) y)->key);
}

typedef struct {
    int64_t L1;                 /* ref_end - ref_start + 1 */
    int32_t *matrix;            /* [(L1+1)][26] */
    int32_t *cov, *snp, *ins, *del;
    allele_list_t *alleles;     /* [L1] */
    /* candidates */
    int64_t n_cand, cap_cand;
    int32_t *images; int64_t *positions; int32_t *depths, *freqs; char *keys;
} vstate_t;

static vstate_t g_v;

static void vstate_free(void) {
    if (g_v.alleles) for (int64_t i = 0; i < g_v.L1; i++) free(g_v.alleles[i].a);
    free(g_v.alleles); free(g_v.matrix); free(g_v.cov); free(g_v.snp); free(g_v.ins); free(g_v.del);
    free(g_v.images); free(g_v.positions); free(g_v.depths); free(g_v.freqs); free(g_v.keys);
    memset(&g_v, 0, sizeof(g_v));
}

/* region_summary.cpp:337-566 populate_summary_matrix for one read */
static void variant_walk_read(const pb_reads_t *R, int64_t r, const pb_region_t *reg, const char *ref,
                              const pb_variant_params_t *p) {
    const int64_t ref_start = reg->ref_start, ref_end = reg->ref_end;
    const int64_t so = R->seq_off[r];
    const int64_t lseq = R->seq_off[r + 1] - so;
    const int rev = R->flags[r] & 1;
    const int64_t c0 = R->cigar_off[r], c1 = R->cigar_off[r + 1];
    int64_t read_index = 0, ref_position = R->pos[r];
    int32_t *M = g_v.matrix;
    char key[PB_ALLELE_STRIDE + 8];

    for (int64_t ci = c0; ci < c1; ci++) {
        const int op = R->cigar[ci] & 15;
        const int64_t len = R->cigar[ci] >> 4;
        if (ref_position > ref_end) break;                                         /* :355 */
        switch (op) {
        case 7: case 8: case 0: {                                                  /* :357-430 */
            int64_t i0 = 0;
            if (ref_position < ref_start) {
                i0 = ref_start - ref_position; if (i0 > len) i0 = len;
                read_index += i0; ref_position += i0;
            }
            for (int64_t i = i0; i < len; i++) {
                if (ref_position >= ref_start && ref_position <= ref_end) {
                    const double q = R->qual[so + read_index];
                    const char base = NT16[seq_code(R, so + read_index)];
                    const int64_t x = ref_position - ref_start;
                    const char ref_base = (x < reg->ref_len) ? ref[reg->ref_off + x] : '\0';
                    const int f = feature_index(ref_base, base, rev);
                    if (q >= p->min_snp_baseq) {
                        g_v.cov[x] += 1;                                            /* :379 */
                        int anchor = 0;                                             /* :381-391 */
                        if (i == len - 1 && ci != c1 - 1) {
                            int nop = R->cigar[ci + 1] & 15;
                            if (nop == 1 || nop == 2) anchor = 1;
                        }
                        if (!anchor) M[x * 26 + (rev ? 15 : 4)] -= 1;
                        if (ref_base != base) {                                     /* :394-421 */
                            g_v.snp[x] += 1;
                            if (f >= 0) M[x * 26 + f] -= 1;
                            key[0] = '1'; key[1] = base; key[2] = 0;
                            tally(&g_v.alleles[x], key, rev);
                        } else if (f >= 0) {
                            M[x * 26 + f] -= 1;                                     /* :423 */
                        }
                    }
                }
                read_index += 1; ref_position += 1;
            }
            break; }
        case 1: {                                                                  /* :431-490 */
            if (ref_position - 1 >= ref_start && ref_position - 1 <= ref_end && read_index - 1 >= 0) {
                const int64_t x = ref_position - 1 - ref_start;
                const char ref_base = (x < reg->ref_len) ? ref[reg->ref_off + x] : '\0';
                const int f = feature_index(ref_base, 'I', rev);
                const int64_t n = len + 1;
                const int64_t s = read_index - 1;
                double qsum = 0;
                for (int64_t i = s; i < s + n; i++) qsum += R->qual[so + i];          /* :448-450 */
                if (qsum >= p->min_indel_baseq * n && R->qual[so + s] < p->min_snp_baseq)
                    g_v.cov[x] += 1;                                                /* :453 */
                /* key = "2" + sequence.substr(read_index-1, len+1); substr clamps at the end */
                int64_t klen = n; if (s + klen > lseq) klen = lseq - s;
                if (1 + klen <= 61 && qsum >= p->min_indel_baseq * n) {              /* :461 */
                    if (f >= 0) M[x * 26 + f] -= 1;
                    g_v.ins[x] += 1;
                    key[0] = '2';
                    for (int64_t i = 0; i < klen; i++) key[1 + i] = NT16[seq_code(R, so + s + i)];
                    key[1 + klen] = 0;
                    tally(&g_v.alleles[x], key, rev);
                }
            }
            read_index += len;
            break; }
        case 2: {                                                                  /* :491-555 */
            if (ref_position - 1 >= ref_start && ref_position - 1 <= ref_end) {
                const int64_t x = ref_position - 1 - ref_start;
                const char ref_base = (x < reg->ref_len) ? ref[reg->ref_off + x] : '\0';
                const int f = feature_index(ref_base, 'D', rev);
                if (f >= 0) M[x * 26 + f] -= 1;                                     /* :497 */
                /* key = "3" + reference.substr(x, len+1) (clamped at the string end) */
                int64_t klen = len + 1; if (x + klen > reg->ref_len) klen = reg->ref_len - x;
                if (klen < 0) klen = 0;
                if (1 + klen <= 61) {                                               /* :511 */
                    g_v.del[x] += 1;
                    key[0] = '3';
                    memcpy(key + 1, ref + reg->ref_off + x, (size_t) klen);
                    key[1 + klen] = 0;
                    tally(&g_v.alleles[x], key, rev);
                }
            }
            for (int64_t i = 0; i < len; i++) {                                     /* :542-552 */
                const int64_t q = ref_position + i;
                if (q >= ref_start && q <= ref_end) {
                    const int64_t x = q - ref_start;
                    const char ref_base = (x < reg->ref_len) ? ref[reg->ref_off + x] : '\0';
                    const int f = feature_index(ref_base, '*', rev);
                    if (f >= 0) M[x * 26 + f] -= 1;
                }
            }
            ref_position += len;
            break; }
        case 3: case 6:                                                             /* :556-558, falls through */
            ref_position += len;
            /* fallthrough */
        case 4:
            read_index += len;                                                      /* :560 */
            break;
        default:                                                                    /* H and unknown ops */
            break;
        }
    }
}

static void cand_reserve(int64_t need) {
    if (need <= g_v.cap_cand) return;
    int64_t cap = g_v.cap_cand ? g_v.cap_cand : 256;
    while (cap < need) cap *= 2;
    g_v.images = (int32_t *) realloc(g_v.images, sizeof(int32_t) * cap * 33 * 26);
    g_v.positions = (int64_t *) realloc(g_v.positions, sizeof(int64_t) * cap);
    g_v.depths = (int32_t *) realloc(g_v.depths, sizeof(int32_t) * cap);
    g_v.freqs = (int32_t *) realloc(g_v.freqs, sizeof(int32_t) * cap);
    g_v.keys = (char *) realloc(g_v.keys, (size_t) cap * PB_ALLELE_STRIDE);
    g_v.cap_cand = cap;
}
static inline int imin(int a, int b) { return a < b ? a : b; }

/* region_summary.cpp:568-916 generate_summary (inference mode) for ONE region.
 * Returns the number of candidates. */
int64_t port_variant_run(const pb_reads_t *R, const pb_region_t *reg, const char *ref,
                         const pb_variant_params_t *p) {
    vstate_free();
    const int64_t L1 = reg->ref_end - reg->ref_start + 1;
    g_v.L1 = L1;
    g_v.matrix = (int32_t *) calloc((size_t) (L1 + 1) * 26, sizeof(int32_t));
    g_v.cov = (int32_t *) calloc((size_t) L1, sizeof(int32_t));
    g_v.snp = (int32_t *) calloc((size_t) L1, sizeof(int32_t));
    g_v.ins = (int32_t *) calloc((size_t) L1, sizeof(int32_t));
    g_v.del = (int32_t *) calloc((size_t) L1, sizeof(int32_t));
    g_v.alleles = (allele_list_t *) calloc((size_t) L1, sizeof(allele_list_t));
    int32_t *M = g_v.matrix;

    for (int64_t x = 0; x < L1; x++)                                                /* :174-191 */
        M[x * 26] = ref_value(x < reg->ref_len ? ref[reg->ref_off + x] : '\0');

    for (int64_t r = reg->read_begin; r < reg->read_end; r++)                        /* :617-623 */
        if (R->mapq[r] > 0 && R->seq_off[r + 1] > R->seq_off[r]) variant_walk_read(R, r, reg, ref, p);

    uint8_t *pass = (uint8_t *) calloc((size_t) L1, 1);                              /* bit0 snp, 1 ins, 2 del, 3 site */
    for (int64_t x = 0; x < L1; x++) {                                              /* :634-654 */
        const double c = g_v.cov[x] > 1 ? (double) g_v.cov[x] : 1.0;
        const double fs = g_v.snp[x] / c, fi = g_v.ins[x] / c, fd = g_v.del[x] / c;
        const int64_t pos = reg->ref_start + x;
        if (fs >= p->snp_freq_threshold || fi >= p->insert_freq_threshold || fd >= p->delete_freq_threshold)
            if (pos >= reg->cand_start && pos <= reg->cand_end && g_v.cov[x] >= p->min_coverage_threshold) {
                pass[x] = 8;
                if (fs >= p->snp_freq_threshold) pass[x] |= 1;
                if (fi >= p->insert_freq_threshold) pass[x] |= 2;
                if (fd >= p->delete_freq_threshold) pass[x] |= 4;
            }
        for (int j = 11; j < 25; j++) {                                             /* :648-653 */
            int32_t v = M[x * 26 + j];
            M[x * 26 + j] = v >= 0 ? (v > 125 ? 125 : v) : (v < -125 ? -125 : v);
        }
    }

    const int64_t region_size = L1;   /* :584 with GENERATE_INDELS=false */
    for (int64_t x = 0; x < L1; x++) {                                              /* :669-912 */
        if (!(pass[x] & 8)) continue;
        allele_list_t *A = &g_v.alleles[x];
        qsort(A->a, (size_t) A->n, sizeof(allele_t), allele_cmp);
        const int depth = imin(g_v.cov[x], 125);                                    /* :682 */
        const char ref_base = x < reg->ref_len ? ref[reg->ref_off + x] : '\0';
        for (int k = 0; k < A->n; k++) {
            const allele_t *al = &A->a[k];
            const char t = al->key[0];
            const double freq = (double) al->total / (depth > 1 ? (double) depth : 1.0);
            if (al->total < p->candidate_support_threshold) continue;               /* :693 */
            if (t != '1' && freq < p->indel_candidate_freq_threshold) continue;     /* :697 */
            if (t == '1' && freq < p->snp_candidate_freq_threshold) continue;       /* :700 */
            if (t != '1' && p->skip_indels) continue;                               /* :704 */
            if ((t == '1' && !(pass[x] & 1)) || (t == '2' && !(pass[x] & 2)) || (t == '3' && !(pass[x] & 4)))
                continue;                                                           /* :708-712 */
            cand_reserve(g_v.n_cand + 1);
            int32_t *img = g_v.images + g_v.n_cand * 33 * 26;
            for (int i = 0; i < 33; i++) {                                           /* :828-841 */
                const int64_t row = x - 16 + i;
                for (int j = 0; j < 26; j++)
                    img[i * 26 + j] = (row < 0 || row > region_size) ? 0 : M[row * 26 + j];
            }
            const int mid = 16;
            const int klen = (int) strlen(al->key);
            if (t == '1') {                                                         /* :848-862 */
                const int ff = feature_index(ref_base, al->key[1], 0), fr = feature_index(ref_base, al->key[1], 1);
                img[mid * 26 + 1] = ref_value(al->key[1]);
                img[mid * 26 + 5] = imin(al->fwd, 125);
                img[mid * 26 + 16] = imin(al->rev, 125);
                /* the reference indexes image_matrix[mid][-1] when the ref base is not ACGT
                   (undefined behaviour); the restatement leaves the row untouched then */
                if (ff >= 0) img[mid * 26 + ff] = -img[mid * 26 + ff];
                if (fr >= 0) img[mid * 26 + fr] = -img[mid * 26 + fr];
            } else if (t == '2') {                                                  /* :863-877 */
                const int ff = feature_index(ref_base, 'I', 0), fr = feature_index(ref_base, 'I', 1);
                img[mid * 26 + 2] = imin(klen - 1, 125);
                img[mid * 26 + 6] = imin(al->fwd, 125);
                img[mid * 26 + 17] = imin(al->rev, 125);
                if (ff >= 0) img[mid * 26 + ff] = -img[mid * 26 + ff];
                if (fr >= 0) img[mid * 26 + fr] = -img[mid * 26 + fr];
            } else {                                                                /* :878-905 */
                const int del_len = klen - 1;
                const int end_index = imin(mid + del_len - 1, 31);
                int ff = feature_index(ref_base, 'D', 0), fr = feature_index(ref_base, 'D', 1);
                img[mid * 26 + 3] = imin(del_len, 125);
                img[mid * 26 + 7] = imin(al->fwd, 125);
                img[mid * 26 + 18] = imin(al->rev, 125);
                if (ff >= 0) img[mid * 26 + ff] = -img[mid * 26 + ff];
                if (fr >= 0) img[mid * 26 + fr] = -img[mid * 26 + fr];
                ff = feature_index(ref_base, '*', 0); fr = feature_index(ref_base, '*', 1);
                for (int idx = mid + 1; idx <= end_index; idx++) {
                    img[idx * 26 + 3] = imin(klen - 1, 125);
                    img[idx * 26 + 7] = imin(al->fwd, 125);
                    img[idx * 26 + 18] = imin(al->rev, 125);
                    if (ff >= 0) img[idx * 26 + ff] = -img[idx * 26 + ff];
                    if (fr >= 0) img[idx * 26 + fr] = -img[idx * 26 + fr];
                }
            }
            g_v.positions[g_v.n_cand] = reg->ref_start + x;
            g_v.depths[g_v.n_cand] = depth;
            g_v.freqs[g_v.n_cand] = imin(al->total, 125);
            memset(g_v.keys + g_v.n_cand * PB_ALLELE_STRIDE, 0, PB_ALLELE_STRIDE);
            strncpy(g_v.keys + g_v.n_cand * PB_ALLELE_STRIDE, al->key, PB_ALLELE_STRIDE - 1);
            g_v.n_cand++;
        }
    }
    free(pass);
    return g_v.n_cand;
}

void port_variant_fetch(int32_t *images, int64_t *positions, int32_t *depths, int32_t *freqs, char *keys) {
    memcpy(images, g_v.images, sizeof(int32_t) * g_v.n_cand * 33 * 26);
    memcpy(positions, g_v.positions, sizeof(int64_t) * g_v.n_cand);
    memcpy(depths, g_v.depths, sizeof(int32_t) * g_v.n_cand);
    memcpy(freqs, g_v.freqs, sizeof(int32_t) * g_v.n_cand);
    memcpy(keys, g_v.keys, (size_t) g_v.n_cand * PB_ALLELE_STRIDE);
}

/* intermediates of the last port_variant_run: matrix int32 [L1][26] (clamped), and the four count vectors */
void port_variant_debug(int32_t *matrix, int32_t *cov, int32_t *snp, int32_t *ins, int32_t *del) {
    if (matrix) memcpy(matrix, g_v.matrix, sizeof(int32_t) * g_v.L1 * 26);
    if (cov) memcpy(cov, g_v.cov, sizeof(int32_t) * g_v.L1);
    if (snp) memcpy(snp, g_v.snp, sizeof(int32_t) * g_v.L1);
    if (ins) memcpy(ins, g_v.ins, sizeof(int32_t) * g_v.L1);
    if (del) memcpy(del, g_v.del, sizeof(int32_t) * g_v.L1);
}

/* ======================================================================
 * Polish encoder: summary_generator.cpp
 * ==================================================================== */
typedef struct {
    int64_t L1;
    double *base;        /* [L1][10]   base_summaries  */
    double *cov;         /* [L1]       coverage        */
    int64_t *longest;    /* [L1]       longest_insert_count */
    double **insr;       /* [L1] -> [longest_cap][10] insert_summaries */
    int64_t *ins_cap;
    int64_t n_cols;
} pstate_t;
static pstate_t g_p;

static void pstate_free(void) {
    if (g_p.insr) for (int64_t i = 0; i < g_p.L1; i++) free(g_p.insr[i]);
    free(g_p.insr); free(g_p.ins_cap); free(g_p.base); free(g_p.cov); free(g_p.longest);
    memset(&g_p, 0, sizeof(g_p));
}

/* summary_generator.cpp:16-32 */
static int polish_feature(char base, int rev) {
    switch (up(base)) {
        case 'A': return rev ? 0 : 4;
        case 'C': return rev ? 1 : 5;
        case 'G': return rev ? 2 : 6;
        case 'T': return rev ? 3 : 7;
        default:  return rev ? 8 : 9;
    }
}

/* summary_generator.cpp:47-121 iterate_over_read */
static void polish_walk_read(const pb_reads_t *R, int64_t r, const pb_region_t *reg) {
    const int64_t ref_start = reg->ref_start, ref_end = reg->ref_end;
    const int64_t so = R->seq_off[r];
    const int rev = R->flags[r] & 1;
    int64_t read_index = 0, ref_position = R->pos[r];
    for (int64_t ci = R->cigar_off[r]; ci < R->cigar_off[r + 1]; ci++) {
        const int op = R->cigar[ci] & 15;
        const int64_t len = R->cigar[ci] >> 4;
        if (ref_position > ref_end) break;                                          /* :54 */
        switch (op) {
        case 7: case 8: case 0: {                                                   /* :56-78 */
            int64_t i0 = 0;
            if (ref_position < ref_start) {
                i0 = ref_start - ref_position; if (i0 > len) i0 = len;
                read_index += i0; ref_position += i0;
            }
            for (int64_t i = i0; i < len; i++) {
                if (ref_position >= ref_start && ref_position <= ref_end) {
                    const int64_t x = ref_position - ref_start;
                    g_p.base[x * 10 + polish_feature(NT16[seq_code(R, so + read_index)], rev)] += 1.0;
                    g_p.cov[x] += 1.0;
                }
                read_index += 1; ref_position += 1;
            }
            break; }
        case 1: {                                                                   /* :80-98 */
            if (ref_position - 1 >= ref_start && ref_position - 1 <= ref_end) {
                const int64_t x = ref_position - 1 - ref_start;
                const int64_t lseq = R->seq_off[r + 1] - so;
                int64_t n = len; if (read_index + n > lseq) n = lseq - read_index;  /* substr clamp */
                if (n < 0) n = 0;
                if (n > g_p.ins_cap[x]) {
                    int64_t cap = g_p.ins_cap[x] ? g_p.ins_cap[x] : 4; while (cap < n) cap *= 2;
                    g_p.insr[x] = (double *) realloc(g_p.insr[x], sizeof(double) * cap * 10);
                    memset(g_p.insr[x] + g_p.ins_cap[x] * 10, 0, sizeof(double) * (cap - g_p.ins_cap[x]) * 10);
                    g_p.ins_cap[x] = cap;
                }
                for (int64_t i = 0; i < n; i++)
                    g_p.insr[x][i * 10 + polish_feature(NT16[seq_code(R, so + read_index + i)], rev)] += 1.0;
                if (n > g_p.longest[x]) g_p.longest[x] = n;
            }
            read_index += len;
            break; }
        case 3: case 6: case 2: {                                                   /* :99-114 */
            for (int64_t i = 0; i < len; i++) {
                const int64_t q = ref_position + i;
                if (q >= ref_start && q <= ref_end) {
                    g_p.base[(q - ref_start) * 10 + polish_feature('*', rev)] += 1.0;
                    /* coverage[ref_position] (the FIRST deleted position), :109 */
                    if (ref_position >= ref_start && ref_position <= ref_end) g_p.cov[ref_position - ref_start] += 1.0;
                }
            }
            ref_position += len;
            break; }
        case 4:
            read_index += len;                                                      /* :115-117 */
            break;
        default: break;
        }
    }
}

/* summary_generator.cpp:370-393 generate_summary; returns number of columns */
int64_t port_polish_run(const pb_reads_t *R, const pb_region_t *reg) {
    pstate_free();
    const int64_t L1 = reg->ref_end - reg->ref_start + 1;
    g_p.L1 = L1;
    g_p.base = (double *) calloc((size_t) L1 * 10, sizeof(double));
    g_p.cov = (double *) calloc((size_t) L1, sizeof(double));
    g_p.longest = (int64_t *) calloc((size_t) L1, sizeof(int64_t));
    g_p.insr = (double **) calloc((size_t) L1, sizeof(double *));
    g_p.ins_cap = (int64_t *) calloc((size_t) L1, sizeof(int64_t));
    for (int64_t r = reg->read_begin; r < reg->read_end; r++)
        if (R->mapq[r] > 0) polish_walk_read(R, r, reg);                             /* :375 */
    int64_t n = 0;
    for (int64_t x = 0; x < L1; x++) n += 1 + g_p.longest[x];                        /* :381-388 */
    g_p.n_cols = n;
    return n;
}

/* summary_generator.cpp:274-306 generate_image: (uint8_t)((count / max(1.0, cov)) * 254).
 * The double -> uint8_t conversion of an out-of-range value is what gcc/x86-64 emits
 * (cvttsd2si to int32, low byte kept); restated explicitly. */
static inline uint8_t pixel(double count, double cov) {
    double v = (count / (cov > 1.0 ? cov : 1.0)) * 254.0;
    return (uint8_t) (int32_t) v;
}
void port_polish_fetch(uint8_t *image, int64_t *pos, int32_t *idx) {
    int64_t c = 0;
    for (int64_t x = 0; x < g_p.L1; x++) {
        for (int j = 0; j < 10; j++) image[c * 10 + j] = pixel(g_p.base[x * 10 + j], g_p.cov[x]);
        pos[c] = x; idx[c] = 0; c++;
        for (int64_t k = 0; k < g_p.longest[x]; k++) {
            for (int j = 0; j < 10; j++) image[c * 10 + j] = pixel(g_p.insr[x][k * 10 + j], g_p.cov[x]);
            pos[c] = x; idx[c] = (int32_t) (k + 1); c++;
        }
    }
}
/* positions are returned relative to ref_start by port_polish_fetch; helper adds it */
void port_polish_fix_positions(int64_t *pos, int64_t n, int64_t ref_start) {
    for (int64_t i = 0; i < n; i++) pos[i] += ref_start;
}
