/* TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
 *
 * Plain-C restatement of BAM_handler::get_reads (pepper/modules/src/dataio/bam_handler.cpp:115-451; the
 * pepper_variant copy is identical) over the SoA record batch of include/pepper_b200.h.  It follows the reference loop
 * statement by statement (per-base inner loops included); the htslib 1.9 region iterator it sits on is restated as the
 * overlap rule documented in oracle/stub/sam.h.  Pinned against the unmodified reference function compiled into
 * oracle/_ref/libref_getreads.so (tests/test_oracle_getreads.py).
 * Output is a pb_reads_t-compatible batch (BAM-native packing) plus pos_end; query_name / bad_indicies / hp_tag are
 * not materialised (no encoder reads them), n_bad is returned so the bad-base rule (:216-222, :307) is still pinned. */
#include <stdint.h>
#include <string.h>
#include "../include/pepper_b200.h"

static int code_at(const uint8_t *seq, int64_t n) { return (n & 1) ? (seq[n >> 1] & 15) : (seq[n >> 1] >> 4); }
static void put_code(uint8_t *seq, int64_t n, int code) {
    if (n & 1) seq[n >> 1] = (uint8_t) ((seq[n >> 1] & 0xf0) | code);
    else seq[n >> 1] = (uint8_t) ((seq[n >> 1] & 0x0f) | (code << 4));
}

/* returns the number of reads; sizes[0..2] = reads, bases, cigar ops.  Output arrays must hold the whole input. */
int64_t port_get_reads(const pb_records_t *R, int64_t start, int64_t stop, int include_supplementary, int min_mapq, int min_baseq,
                       int64_t *o_pos, int64_t *o_pos_end, int64_t *o_seq_off, int64_t *o_cigar_off, uint8_t *o_flags,
                       uint8_t *o_mapq, uint8_t *o_seq, uint8_t *o_qual, uint32_t *o_cigar, int64_t *o_n_bad,
                       int64_t *sizes) {
    int64_t n_out = 0, nb = 0, nc = 0;
    for (int64_t r = 0; r < R->n_records; r++) {
        const int64_t pos = R->pos[r];
        const int64_t so = R->seq_off[r];
        const int64_t co = R->cigar_off[r], n_cigar = R->cigar_off[r + 1] - co;
        /* htslib iterator: overlap of [pos, pos + rlen) with [start, stop) */
        int64_t rlen = 0;
        for (int64_t k = 0; k < n_cigar; k++) {
            const int op = (int) (R->cigar[co + k] & 15);
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += R->cigar[co + k] >> 4;
        }
        if (n_cigar == 0) rlen = 1;
        if (pos >= stop) break;                       /* coordinate sorted */
        if (!(pos + rlen > start)) continue;
        /* flags (:139-146) and MAPQ (:148) */
        const int flag = R->flag[r];
        if ((flag & 512) || (flag & 1024) || (flag & 256) || (flag & 4)) continue;
        if (!include_supplementary && (flag & 2048)) continue;
        if ((int) R->mapq[r] < min_mapq) continue;

        long long pos_start = -1, pos_end = -1;
        long long current_read_pos = pos;
        int64_t current_read_index = 0;
        int64_t n_bad = 0;
        const int64_t nb0 = nb, nc0 = nc;
        for (int64_t k = 0; k < n_cigar; k++) {
            const int cigar_op = (int) (R->cigar[co + k] & 15);
            const int64_t cigar_len = R->cigar[co + k] >> 4;
            int64_t modified = 0, cigar_index;
            if (current_read_pos > stop) break;                                               /* :186 */
            switch (cigar_op) {
                case 0: case 8: case 7:                                                       /* M X = (:191-238) */
                    cigar_index = 0;
                    if (current_read_pos < start) {
                        cigar_index = (start - current_read_pos < cigar_len) ? start - current_read_pos : cigar_len;
                        current_read_index += cigar_index;
                        current_read_pos += cigar_index;
                    }
                    for (int64_t i = cigar_index; i < cigar_len; i++) {
                        if (current_read_pos <= stop) {
                            if (pos_start == -1) { pos_start = current_read_pos; pos_end = pos_start; }
                            const int q = R->qual[so + current_read_index];
                            const int code = code_at(R->seq, so + current_read_index);
                            o_qual[nb] = (uint8_t) q;
                            put_code(o_seq, nb, code);
                            if (q < min_baseq || !(code == 1 || code == 2 || code == 4 || code == 8)) n_bad++;
                            nb++; modified++; pos_end++;
                        } else break;
                        current_read_index++; current_read_pos++;
                    }
                    if (modified > 0) o_cigar[nc++] = (uint32_t) (modified << 4 | cigar_op);
                    break;
                case 4: case 1:                                                               /* S I (:239-271) */
                    if (current_read_pos >= start && current_read_pos <= stop && pos_start != -1) {
                        for (int64_t i = 0; i < cigar_len; i++) {
                            const int q = R->qual[so + current_read_index];
                            const int code = code_at(R->seq, so + current_read_index);
                            o_qual[nb] = (uint8_t) q;
                            put_code(o_seq, nb, code);
                            if (q < min_baseq || !(code == 1 || code == 2 || code == 4 || code == 8)) n_bad++;
                            nb++; modified++; current_read_index++;
                        }
                    } else current_read_index += cigar_len;
                    if (modified > 0) o_cigar[nc++] = (uint32_t) (modified << 4 | cigar_op);
                    break;
                case 3: case 2:                                                               /* N D (:272-299) */
                    if (current_read_pos >= start && current_read_pos <= stop && pos_start != -1) {
                        for (int64_t i = 0; i < cigar_len; i++) {
                            if (current_read_pos <= stop) { modified++; pos_end++; } else break;
                            current_read_pos++;
                        }
                    } else current_read_pos += cigar_len;
                    if (modified > 0) o_cigar[nc++] = (uint32_t) (modified << 4 | cigar_op);
                    break;
                default:                                                                      /* H (:300), P/B: no case */
                    break;
            }
        }
        n_bad++;                                                                              /* sentinel len+1 (:307) */
        if (nb > nb0) {                                                                       /* :432 */
            o_pos[n_out] = pos_start; o_pos_end[n_out] = pos_end;
            o_seq_off[n_out] = nb0; o_cigar_off[n_out] = nc0;
            o_flags[n_out] = (flag & 16) ? 1 : 0; o_mapq[n_out] = R->mapq[r]; o_n_bad[n_out] = n_bad;
            n_out++;
        } else { nb = nb0; nc = nc0; }
    }
    o_seq_off[n_out] = nb; o_cigar_off[n_out] = nc;
    sizes[0] = n_out; sizes[1] = nb; sizes[2] = nc;
    return n_out;
}
