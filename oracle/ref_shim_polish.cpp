// TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
//
// Compiles the UNMODIFIED reference polish encoder
//   /root/reference/pepper/modules/src/pileup_summary/summary_generator.cpp
// (included from where it lies) behind a flat C interface taking the SoA read
// batch of include/pepper_b200.h.  Built by oracle/Makefile into oracle/_ref/.
#include <vector>
#include <map>
#include <set>
#include <string>
#include <iostream>
#include <cstring>
#include <cstdint>
using namespace std;
#include "pileup_summary/summary_generator.cpp"

#include "../include/pepper_b200.h"

static const char NT16[] = "=ACMGRSVTWYHKDBN";

static vector<type_read> build_reads(const pb_reads_t *R, int64_t rb, int64_t re) {
    vector<type_read> reads;
    reads.reserve(re - rb);
    for (int64_t r = rb; r < re; r++) {
        type_read rd;
        rd.pos = R->pos[r];
        int64_t so = R->seq_off[r], l = R->seq_off[r + 1] - so;
        rd.sequence.resize(l);
        rd.base_qualities.resize(l);
        for (int64_t i = 0; i < l; i++) {
            int64_t n = so + i;
            uint8_t b = R->seq[n >> 1];
            int code = (n & 1) ? (b & 15) : (b >> 4);
            rd.sequence[i] = NT16[code];
            rd.base_qualities[i] = R->qual[n];
        }
        long long ref_len = 0;
        for (int64_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; c++) {
            int op = R->cigar[c] & 15, len = R->cigar[c] >> 4;
            rd.cigar_tuples.push_back(CigarOp(op, len));
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ref_len += len;
        }
        rd.pos_end = rd.pos + ref_len;
        rd.flags.is_reverse = (R->flags[r] & 1) != 0;
        rd.mapping_quality = R->mapq[r];
        rd.hp_tag = 0;
        rd.read_id = (int) r;
        reads.push_back(rd);
    }
    return reads;
}

static SummaryGenerator *g_gen = nullptr;

extern "C" {

// SummaryGenerator(ref, chrom, start, end).generate_summary(reads, start, end)
// as called by pepper/modules/python/AlignmentSummarizer.py:341-348.
// Returns the number of image columns.
int64_t ref_polish_run(const pb_reads_t *reads, const pb_region_t *region) {
    vector<type_read> rd = build_reads(reads, region->read_begin, region->read_end);
    delete g_gen;
    g_gen = new SummaryGenerator(string(), "contig", region->ref_start, region->ref_end);
    g_gen->generate_summary(rd, region->ref_start, region->ref_end);
    return (int64_t) g_gen->genomic_pos.size();
}

void ref_polish_fetch(uint8_t *image, int64_t *pos, int32_t *idx) {
    for (size_t c = 0; c < g_gen->genomic_pos.size(); c++) {
        for (int j = 0; j < PB_POLISH_FEATURES; j++) image[c * PB_POLISH_FEATURES + j] = g_gen->image[c][j];
        pos[c] = g_gen->genomic_pos[c].first;
        idx[c] = g_gen->genomic_pos[c].second;
    }
}

}  // extern "C"
