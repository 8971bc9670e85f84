// TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
//
// Compiles the UNMODIFIED reference BAM_handler
//   /root/reference/pepper/modules/src/dataio/bam_handler.cpp
// (included from where it lies; the pepper_variant copy is identical) against the in-memory htslib stand-in of
// oracle/stub/sam.h and exposes get_reads() over the SoA record batch of include/pepper_b200.h (pb_records_t).
// Built by oracle/Makefile into oracle/_ref/libref_getreads.so.
#include <vector>
#include <map>
#include <set>
#include <string>
#include <iostream>
#include <cstring>
#include <cstdint>
using namespace std;
#include "dataio/bam_handler.cpp"

#include "../include/pepper_b200.h"

static MemBam g_bam;
static vector<type_read> g_reads;

extern "C" {

// install the record batch as the "BAM file" (one contig, tid 0)
void ref_getreads_load(const pb_records_t *R) {
    g_bam.targets.assign(1, "contig");
    g_bam.target_len.assign(1, 0x7fffffff);
    g_bam.records.clear();
    g_bam.records.resize(R->n_records);
    for (int64_t r = 0; r < R->n_records; r++) {
        MemBamRecord &m = g_bam.records[r];
        memset(&m.core, 0, sizeof(m.core));
        const int64_t so = R->seq_off[r], l = R->seq_off[r + 1] - so;
        const int64_t co = R->cigar_off[r], nc = R->cigar_off[r + 1] - co;
        m.core.tid = 0; m.core.pos = (int32_t) R->pos[r]; m.core.qual = R->mapq[r]; m.core.flag = R->flag[r];
        m.core.l_qname = 4; m.core.n_cigar = (uint32_t) nc; m.core.l_qseq = (int32_t) l;
        m.data.assign(4 + 4 * nc + (l + 1) / 2 + l, 0);
        memcpy(m.data.data(), "r\0\0\0", 4);
        memcpy(m.data.data() + 4, R->cigar + co, 4 * nc);
        uint8_t *s = m.data.data() + 4 + 4 * nc;
        for (int64_t i = 0; i < l; i++) {
            const int64_t n = so + i;
            const int code = (n & 1) ? (R->seq[n >> 1] & 15) : (R->seq[n >> 1] >> 4);
            s[i >> 1] |= (uint8_t) (code << ((~i & 1) << 2));
        }
        memcpy(s + (l + 1) / 2, R->qual + so, l);
    }
    membam_current() = &g_bam;
}

// BAM_handler("mem").get_reads("contig", start, stop, include_supplementary, min_mapq, min_baseq);
// sizes[0..2] = reads, bases, cigar ops of the result (kept until the next call).
void ref_getreads_query(int64_t start, int64_t stop, int include_supplementary, int min_mapq, int min_baseq, int64_t *sizes) {
    BAM_handler h("mem");
    g_reads = h.get_reads("contig", start, stop, include_supplementary != 0, min_mapq, min_baseq);
    int64_t nb = 0, nc = 0;
    for (auto &r : g_reads) { nb += (int64_t) r.sequence.size(); nc += (int64_t) r.cigar_tuples.size(); }
    sizes[0] = (int64_t) g_reads.size(); sizes[1] = nb; sizes[2] = nc;
}

// flatten the last result: ASCII sequence, qualities, (op,len) pairs, pos/pos_end, flags (bit0 reverse), mapq
void ref_getreads_fetch(int64_t *pos, int64_t *pos_end, int64_t *seq_off, int64_t *cigar_off, uint8_t *flags, uint8_t *mapq,
                        char *seq_ascii, uint8_t *qual, int32_t *cigar_op, int32_t *cigar_len, int64_t *n_bad) {
    int64_t so = 0, co = 0;
    for (size_t i = 0; i < g_reads.size(); i++) {
        const type_read &r = g_reads[i];
        pos[i] = r.pos; pos_end[i] = r.pos_end; seq_off[i] = so; cigar_off[i] = co;
        flags[i] = r.flags.is_reverse ? 1 : 0; mapq[i] = (uint8_t) r.mapping_quality;
        n_bad[i] = (int64_t) r.bad_indicies.size();
        for (size_t k = 0; k < r.sequence.size(); k++) { seq_ascii[so + k] = r.sequence[k]; qual[so + k] = (uint8_t) r.base_qualities[k]; }
        so += (int64_t) r.sequence.size();
        for (auto &c : r.cigar_tuples) { cigar_op[co] = c.operation; cigar_len[co] = c.length; co++; }
    }
    seq_off[g_reads.size()] = so; cigar_off[g_reads.size()] = co;
}
}
