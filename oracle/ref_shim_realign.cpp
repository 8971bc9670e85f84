// TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
//
// Compiles the UNMODIFIED reference read->reference realigner
//   /root/reference/pepper/modules/src/local_reassembly/simple_aligner.cpp   (ReadAligner::align_reads_to_reference)
// which itself #includes ssw_cpp.cpp and ssw.c (the SSW library vendored in the reference tree), from where they lie,
// behind a flat C interface over the SoA read batch of include/pepper_b200.h.
// Built by oracle/Makefile into oracle/_ref/libref_realign.so.
#include <vector>
#include <map>
#include <set>
#include <string>
#include <iostream>
#include <cstring>
#include <cstdint>
using namespace std;
#include "local_reassembly/simple_aligner.cpp"

#include "../include/pepper_b200.h"

static const char NT16R[] = "=ACMGRSVTWYHKDBN";
static vector<type_read> g_out;

extern "C" {

// ReadAligner(ref_start, ref_end, ref_seq).align_reads_to_reference(reads[rb:re]) as called by
// pepper/modules/python/AlignmentSummarizer.py:159-177.  sizes[0..1] = reads, cigar ops of the result.
void ref_realign_run(const pb_reads_t *R, int64_t rb, int64_t re, int64_t ref_start, int64_t ref_end, const char *ref_seq,
                     int64_t ref_len, int64_t *sizes) {
    vector<type_read> reads;
    for (int64_t r = rb; r < re; r++) {
        type_read rd;
        rd.pos = R->pos[r];
        const int64_t so = R->seq_off[r], l = R->seq_off[r + 1] - so;
        rd.sequence.resize(l);
        rd.base_qualities.resize(l);
        for (int64_t i = 0; i < l; i++) {
            const int64_t n = so + i;
            const int code = (n & 1) ? (R->seq[n >> 1] & 15) : (R->seq[n >> 1] >> 4);
            rd.sequence[i] = NT16R[code];
            rd.base_qualities[i] = R->qual[n];
        }
        long long rl = 0;
        for (int64_t c = R->cigar_off[r]; c < R->cigar_off[r + 1]; c++) {
            const int op = R->cigar[c] & 15, len = R->cigar[c] >> 4;
            rd.cigar_tuples.push_back(CigarOp(op, len));
            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rl += len;
        }
        rd.pos_end = rd.pos + rl;
        rd.flags.is_reverse = (R->flags[r] & 1) != 0;
        rd.mapping_quality = R->mapq[r];
        reads.push_back(rd);
    }
    ReadAligner aligner((int) ref_start, (int) ref_end, string(ref_seq, ref_seq + ref_len));
    g_out = aligner.align_reads_to_reference(reads);
    int64_t nc = 0;
    for (auto &r : g_out) nc += (int64_t) r.cigar_tuples.size();
    sizes[0] = (int64_t) g_out.size(); sizes[1] = nc;
}

void ref_realign_fetch(int64_t *pos, int64_t *pos_end, int64_t *cigar_off, int32_t *op, int32_t *len) {
    int64_t co = 0;
    for (size_t i = 0; i < g_out.size(); i++) {
        pos[i] = g_out[i].pos; pos_end[i] = g_out[i].pos_end; cigar_off[i] = co;
        for (auto &c : g_out[i].cigar_tuples) { op[co] = c.operation; len[co] = c.length; co++; }
    }
    cigar_off[g_out.size()] = co;
}

// raw SSW result of one query against one reference string (Aligner::SetReferenceSequence + Align_cpp), for unit pinning
void ref_ssw_align(const char *query, const char *ref, int32_t ref_len, int32_t *out /* score, ref_begin, ref_end, query_begin, query_end, mismatches */,
                   char *cigar_string, int32_t cap) {
    LibSSWPairwiseAligner a;
    a.set_reference(string(ref, ref + ref_len));
    Alignment al = a.align(string(query));
    out[0] = al.sw_score; out[1] = al.ref_begin; out[2] = al.ref_end; out[3] = al.query_begin; out[4] = al.query_end; out[5] = al.mismatches;
    strncpy(cigar_string, al.cigar_string.c_str(), cap - 1);
    cigar_string[cap - 1] = 0;
}
}
