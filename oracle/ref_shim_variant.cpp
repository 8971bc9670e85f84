// TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path.
//
// Compiles the UNMODIFIED reference variant encoder
//   /root/reference/pepper_variant/modules/cpp/region_summary.cpp
// (included from where it lies; nothing is copied into this repo) behind a
// flat C interface so tests / bench can drive it with the same SoA read batch
// the CUDA library takes (include/pepper_b200.h).  htslib headers are stubbed
// (oracle/stub): BAM_handler / FASTA_handler are declared but never defined or
// called in this TU.  Built by oracle/Makefile into oracle/_ref/.
#include <vector>
#include <map>
#include <set>
#include <string>
#include <iomanip>
#include <iostream>
#include <cstring>
#include <cstdint>
using namespace std;
#include "bam_handler.h"
#include "candidate_finder.h"
#include "region_summary.cpp"

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

static vector<CandidateImageSummary> g_last;

extern "C" {

// Runs RegionalSummaryGenerator exactly as AlignmentSummarizer.create_summary
// does (pepper_variant/modules/python/AlignmentSummarizer.py:220-238).
// Returns the number of candidates; fetch them with ref_variant_fetch.
int64_t ref_variant_run(const pb_reads_t *reads, const pb_region_t *region,
                        const char *ref, const pb_variant_params_t *p) {
    vector<type_read> rd = build_reads(reads, region->read_begin, region->read_end);
    string ref_seq(ref + region->ref_off, (size_t) region->ref_len);
    RegionalSummaryGenerator gen("contig", region->ref_start, region->ref_end, ref_seq);
    gen.generate_max_insert_summary(rd);
    g_last = gen.generate_summary(rd, p->min_snp_baseq, p->min_indel_baseq,
                                  p->snp_freq_threshold, p->insert_freq_threshold,
                                  p->delete_freq_threshold, p->min_coverage_threshold,
                                  p->snp_candidate_freq_threshold,
                                  p->indel_candidate_freq_threshold,
                                  p->candidate_support_threshold, p->skip_indels != 0,
                                  region->cand_start, region->cand_end,
                                  PB_VARIANT_WINDOW - 1, PB_VARIANT_FEATURES, false);
    return (int64_t) g_last.size();
}

// images int32 [n][33][26] (unwrapped ints, as the pybind object holds them)
void ref_variant_fetch(int32_t *images, int64_t *positions, int32_t *depths,
                       int32_t *freqs, char *keys) {
    for (size_t k = 0; k < g_last.size(); k++) {
        const CandidateImageSummary &c = g_last[k];
        for (int i = 0; i < PB_VARIANT_WINDOW; i++)
            for (int j = 0; j < PB_VARIANT_FEATURES; j++)
                images[(k * PB_VARIANT_WINDOW + i) * PB_VARIANT_FEATURES + j] = c.image_matrix[i][j];
        positions[k] = c.position;
        depths[k] = c.depth;
        freqs[k] = c.candidate_frequency.empty() ? -1 : c.candidate_frequency[0];
        memset(keys + k * PB_ALLELE_STRIDE, 0, PB_ALLELE_STRIDE);
        if (!c.candidates.empty())
            strncpy(keys + k * PB_ALLELE_STRIDE, c.candidates[0].c_str(), PB_ALLELE_STRIDE - 1);
    }
}

}  // extern "C"
