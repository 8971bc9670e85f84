/* TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
 *
 * Plain-C restatement of the reference's read -> reference realignment:
 *   ReadAligner::align_reads_to_reference   pepper/modules/src/local_reassembly/simple_aligner.cpp:66-106
 *   Aligner::Align_cpp / ConvertAlignment / CalculateNumberMismatch   ssw_cpp.cpp:52-215, 330-362
 *   ssw_align / sw_sse2_byte / sw_sse2_word / banded_sw               ssw.c:161-367, 393-569, 571-757, 801-891
 * The SSE2 striped kernels are restated as scalar recurrences that keep their observable quirks:
 *   (1) E(i+1,p) opens from the H computed in the striped main loop, which only sees the F chain of its OWN segment
 *       (segLen = ceil(readLen / lanes), lanes = 16 in byte mode, 8 in word mode); the Lazy-F pass fixes H but not E
 *       ("disallow adjacent insertion and then deletion", ssw.c:282, 487);
 *   (2) byte mode first, word mode when the byte score saturates (max + bias >= 255, ssw.c:330, 826-830);
 *   (3) best cell = highest score, then first column in iteration order, then smallest read index (ssw.c:519-531);
 *   (4) banded_sw's band arithmetic including the zeroing of index min(end+1, width-1) (ssw.c:624) and the
 *       trace-back that starts at the bottom-right corner and stops at read row 0 (ssw.c:665-733).
 * Pinned against the unmodified reference sources compiled into oracle/_ref/libref_realign.so
 * (tests/test_oracle_realign.py). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../include/pepper_b200.h"

#define MATCH 4
#define MISMATCH 6
#define GAP_O 8
#define GAP_E 2
#define BIAS 6

static int base_code(int ch) {            /* kBaseTranslation, ssw_cpp.cpp:12-30 */
    switch (ch) {
        case 'A': case 'a': case 'U': case 'u': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}
static int score_of(int a, int b) { return (a < 4 && b < 4 && a == b) ? MATCH : -MISMATCH; }   /* BuildSwScoreMatrix */

typedef struct { int score, ref, read; } sw_best;

/* one striped pass (forward: ref_dir 0, or over the reversed reference: ref_dir 1) */
static sw_best sw_striped(const int8_t *ref, int ref_dir, int refLen, const int8_t *read, int readLen, int lanes, int byte_mode,
                          int terminate) {
    const int S = (readLen + lanes - 1) / lanes, P = S * lanes;
    int *H = (int *) calloc(P + 1, sizeof(int)), *Hn = (int *) calloc(P + 1, sizeof(int)), *E = (int *) calloc(P + 1, sizeof(int));
    int *Hmax = (int *) calloc(P + 1, sizeof(int));
    int max = 0, end_ref = byte_mode ? -1 : 0, end_read = readLen - 1, overflow = 0;
    int i = ref_dir ? refLen - 1 : 0;
    const int stop = ref_dir ? -1 : refLen, step = ref_dir ? -1 : 1;
    for (; i != stop; i += step) {
        int floc = 0, ffull = 0, colmax = 0;
        for (int p = 0; p < P; p++) {
            if (p % S == 0) floc = 0;                                     /* vF = 0 at the head of every lane */
            const int s = p < readLen ? score_of(ref[i], read[p]) : 0;    /* padded profile entries score 0 */
            int h = (p ? H[p - 1] : 0) + s;
            if (h < 0) h = 0;
            if (E[p] > h) h = E[p];
            int hm = floc > h ? floc : h;                                 /* what the main loop stores */
            int open = hm - GAP_O; if (open < 0) open = 0;
            int e = E[p] - GAP_E; if (e < 0) e = 0;
            E[p] = e > open ? e : open;                                   /* E never sees the Lazy-F correction */
            floc -= GAP_E; if (floc < 0) floc = 0;
            if (open > floc) floc = open;
            const int hf = ffull > hm ? ffull : hm;                       /* after the Lazy-F pass */
            ffull -= GAP_E; if (ffull < 0) ffull = 0;
            if (open > ffull) ffull = open;
            Hn[p] = hf;
            if (hf > colmax) colmax = hf;
        }
        int *t = H; H = Hn; Hn = t;
        if (colmax > max) {
            max = colmax;
            if (byte_mode && max + BIAS >= 255) { overflow = 1; break; }
            end_ref = i;
            memcpy(Hmax, H, sizeof(int) * P);
        }
        if (colmax == terminate) break;
    }
    for (int p = 0; p < P; p++) if (Hmax[p] == max && p < end_read) end_read = p;
    free(H); free(Hn); free(E); free(Hmax);
    sw_best b;
    b.score = overflow ? 255 : max; b.ref = end_ref; b.read = end_read;
    return b;
}

/* banded_sw (ssw.c:571-757) with its arrays in band coordinates; returns the number of cigar words (len<<4|op with
 * op 0=M 1=I 2=D) written to `cig`, or -1 on a trace-back error */
static int banded_trace(const int8_t *ref, const int8_t *read, int refLen, int readLen, int score, int band_width, uint32_t *cig,
                        int cap) {
    int max = 0, width = 0, width_d = 0;                                  /* max is kept across band retries, as in the reference */
    int8_t *dir = NULL;
    int *h_b = NULL, *e_b = NULL, *h_c = NULL;
    do {
        width = band_width * 2 + 3; width_d = band_width * 2 + 1;
        free(h_b); free(e_b); free(h_c); free(dir);
        h_b = (int *) calloc(width + 2, sizeof(int)); e_b = (int *) calloc(width + 2, sizeof(int)); h_c = (int *) calloc(width + 2, sizeof(int));
        dir = (int8_t *) calloc((size_t) width_d * readLen * 3 + 8, 1);
        for (int i = 0; i < readLen; i++) {
            int beg = i - band_width > 0 ? i - band_width : 0;
            int end = i + band_width < refLen - 1 ? i + band_width : refLen - 1;
            const int edge = end + 1 < width - 1 ? end + 1 : width - 1;
            const int x = beg, xp = (i - 1 - band_width) > 0 ? i - 1 - band_width : 0;
            int f = 0, u = 0;
            h_b[0] = e_b[0] = h_b[edge] = e_b[edge] = h_c[0] = 0;
            int8_t *line = dir + (size_t) width_d * i * 3;
            for (int j = beg; j <= end; j++) {
                u = j - x + 1;
                const int e_i = j - xp + 1, d_i = j - 1 - xp + 1, b_i = u - 1, dd = (j - x) * 3;
                int t1 = i == 0 ? -GAP_O : h_b[e_i] - GAP_O;
                int t2 = i == 0 ? -GAP_E : e_b[e_i] - GAP_E;
                e_b[u] = t1 > t2 ? t1 : t2;
                line[dd + 0] = t1 > t2 ? 3 : 2;
                t1 = h_c[b_i] - GAP_O; t2 = f - GAP_E;
                f = t1 > t2 ? t1 : t2;
                line[dd + 1] = t1 > t2 ? 5 : 4;
                const int e1 = e_b[u] > 0 ? e_b[u] : 0, f1 = f > 0 ? f : 0;
                t1 = e1 > f1 ? e1 : f1;
                t2 = h_b[d_i] + score_of(ref[j], read[i]);
                h_c[u] = t1 > t2 ? t1 : t2;
                if (h_c[u] > max) max = h_c[u];
                if (t1 <= t2) line[dd + 2] = 1;
                else line[dd + 2] = e1 > f1 ? line[dd + 0] : line[dd + 1];
            }
            for (int j = 1; j <= u; j++) h_b[j] = h_c[j];
        }
        band_width *= 2;
    } while (max < score && band_width < (1 << 24));
    band_width /= 2;
    /* trace back from the bottom-right corner */
    int i = readLen - 1, j = refLen - 1, e = 0, l = 0, state = 2, rc = 0;
    char op = 'M', prev = 'M';
    uint32_t *tmp = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (readLen + refLen + 4));
    while (i > 0) {
        const int x = (i - band_width) > 0 ? i - band_width : 0;
        const int col = j - x;
        if (col < 0 || col >= width_d) { rc = -1; break; }
        const int code = dir[(size_t) width_d * i * 3 + col * 3 + state];
        if (code == 1) { i--; j--; state = 2; op = 'M'; }
        else if (code == 2) { i--; state = 0; op = 'I'; }
        else if (code == 3) { i--; state = 2; op = 'I'; }
        else if (code == 4) { j--; state = 1; op = 'D'; }
        else if (code == 5) { j--; state = 2; op = 'D'; }
        else { rc = -1; break; }
        if (op == prev) e++;
        else { tmp[l++] = (uint32_t) e << 4 | (prev == 'M' ? 0 : prev == 'I' ? 1 : 2); prev = op; e = 1; }
    }
    if (rc == 0) {
        if (op == 'M') tmp[l++] = (uint32_t) (e + 1) << 4 | 0;
        else { tmp[l++] = (uint32_t) e << 4 | (op == 'I' ? 1 : 2); tmp[l++] = 1u << 4 | 0; }
        if (l > cap) rc = -1;
        else { for (int k = 0; k < l; k++) cig[k] = tmp[l - 1 - k]; rc = l; }
    }
    free(tmp); free(h_b); free(e_b); free(h_c); free(dir);
    return rc;
}

/* Aligner::Align_cpp on translated sequences.  out: score, ref_begin, ref_end, query_begin, query_end, mismatches.
 * cig: final cigar (S, =, X, I, D as BAM ops 4, 7, 8, 1, 2); returns its length (0 when no cigar was produced). */
int port_ssw_align(const char *query, int query_len, const char *refseq, int ref_len, int32_t *out, uint32_t *cig, int cap) {
    memset(out, 0, sizeof(int32_t) * 6);
    if (query_len == 0 || ref_len == 0) return 0;
    int8_t *q = (int8_t *) malloc(query_len), *r = (int8_t *) malloc(ref_len);
    for (int i = 0; i < query_len; i++) q[i] = (int8_t) base_code((unsigned char) query[i]);
    for (int i = 0; i < ref_len; i++) r[i] = (int8_t) base_code((unsigned char) refseq[i]);
    int word = 0;
    sw_best b = sw_striped(r, 0, ref_len, q, query_len, 16, 1, 255);
    if (b.score == 255) { b = sw_striped(r, 0, ref_len, q, query_len, 8, 0, 65535); word = 1; }
    out[0] = b.score; out[2] = b.ref; out[4] = b.read;
    out[1] = -1; out[3] = -1;
    int n_cig = 0;
    if (b.score > 0 && b.ref >= 0) {
        const int rl = b.ref + 1, ql = b.read + 1;
        int8_t *qr = (int8_t *) malloc(ql);
        for (int i = 0; i < ql; i++) qr[i] = q[ql - 1 - i];
        const sw_best rb = word ? sw_striped(r, 1, rl, qr, ql, 8, 0, b.score) : sw_striped(r, 1, rl, qr, ql, 16, 1, b.score);
        free(qr);
        out[1] = rb.ref; out[3] = b.read - rb.read;
        const int refLen = b.ref - out[1] + 1, readLen = b.read - out[3] + 1;
        if (refLen > 0 && readLen > 0) {
            uint32_t *path = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (refLen + readLen + 8));
            int bw = refLen - readLen; if (bw < 0) bw = -bw; bw += 1;
            const int pl = banded_trace(r + out[1], q + out[3], refLen, readLen, b.score, bw, path, refLen + readLen + 8);
            if (pl > 0) {
                /* ConvertAlignment + CalculateNumberMismatch: soft clips, M -> runs of = / X */
                const int8_t *rp = r + out[1], *qp = q + out[3];
                int mism = 0, in_m = 0, in_x = 0, len_m = 0, len_x = 0;
                if (out[3] > 0 && n_cig < cap) cig[n_cig++] = (uint32_t) out[3] << 4 | 4;
                for (int k = 0; k < pl; k++) {
                    const int op = (int) (path[k] & 15), len = (int) (path[k] >> 4);
                    if (op == 0) {
                        for (int t = 0; t < len; t++, rp++, qp++) {
                            if (*rp != *qp) {
                                mism++;
                                if (in_m && n_cig < cap) cig[n_cig++] = (uint32_t) len_m << 4 | 7;
                                len_m = 0; len_x++; in_m = 0; in_x = 1;
                            } else {
                                if (in_x && n_cig < cap) cig[n_cig++] = (uint32_t) len_x << 4 | 8;
                                len_m++; len_x = 0; in_m = 1; in_x = 0;
                            }
                        }
                    } else {
                        if (op == 1) qp += len; else rp += len;
                        mism += len;
                        if (in_m && n_cig < cap) cig[n_cig++] = (uint32_t) len_m << 4 | 7;
                        else if (in_x && n_cig < cap) cig[n_cig++] = (uint32_t) len_x << 4 | 8;
                        in_m = in_x = 0; len_m = len_x = 0;
                        if (n_cig < cap) cig[n_cig++] = path[k];
                    }
                }
                if (in_m && n_cig < cap) cig[n_cig++] = (uint32_t) len_m << 4 | 7;
                else if (in_x && n_cig < cap) cig[n_cig++] = (uint32_t) len_x << 4 | 8;
                const int tail = query_len - b.read - 1;
                if (tail > 0 && n_cig < cap) cig[n_cig++] = (uint32_t) tail << 4 | 4;
                out[5] = mism;
            }
            free(path);
        }
    }
    free(q); free(r);
    return n_cig;
}

/* align_reads_to_reference over reads [rb, re) of a pb_reads_t: new pos / pos_end / cigar (ops: M for = and X, S, I, D);
 * reads that start before region_start are dropped (simple_aligner.cpp:73-77).  kept[i] = index of output read i in
 * the input.  Returns the number of output reads; o_cigar_off has n+1 entries. */
static const char NT16P[] = "=ACMGRSVTWYHKDBN";
int64_t port_realign(const pb_reads_t *R, int64_t rb, int64_t re, int64_t region_start, const char *ref_seq, int64_t ref_len,
                     int64_t *o_pos, int64_t *o_pos_end, int64_t *o_cigar_off, uint32_t *o_cigar, int64_t *kept, int32_t *o_score) {
    int64_t n = 0, nc = 0;
    for (int64_t r = rb; r < re; r++) {
        if (R->pos[r] < region_start) continue;
        const int64_t so = R->seq_off[r], l = R->seq_off[r + 1] - so;
        const int64_t start_index = R->pos[r] - region_start;
        char *q = (char *) malloc((size_t) l + 1);
        for (int64_t i = 0; i < l; i++) {
            const int64_t k = so + i;
            q[i] = NT16P[(k & 1) ? (R->seq[k >> 1] & 15) : (R->seq[k >> 1] >> 4)];
        }
        uint32_t *cig = (uint32_t *) malloc(sizeof(uint32_t) * (size_t) (2 * l + ref_len + 16));
        int32_t out[6] = {0, 0, 0, 0, 0, 0};
        int nk = 0;
        if (start_index <= ref_len) nk = port_ssw_align(q, (int) l, ref_seq + start_index, (int) (ref_len - start_index), out, cig, (int) (2 * l + ref_len + 16));
        long long rl = 0;
        for (int64_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; c++) {
            const int op = R->cigar[c] & 15;
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += R->cigar[c] >> 4;
        }
        o_cigar_off[n] = nc;
        kept[n] = r;
        o_score[n] = out[0];
        if (out[0] > 1 && nk > 0) {
            for (int k = 0; k < nk; k++) {
                int op = (int) (cig[k] & 15);
                if (op == 7 || op == 8) op = 0;                                   /* CigarOperationFromChar: '=' and 'X' -> MATCH */
                o_cigar[nc++] = (cig[k] & ~15u) | (uint32_t) op;
            }
            o_pos[n] = R->pos[r] + out[1];
            o_pos_end[n] = R->pos[r] + out[2];
        } else {
            for (int64_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; c++) o_cigar[nc++] = R->cigar[c];
            o_pos[n] = R->pos[r];
            o_pos_end[n] = R->pos[r] + rl;
        }
        n++;
        free(q); free(cig);
    }
    o_cigar_off[n] = nc;
    return n;
}
