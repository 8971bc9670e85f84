// TEST INFRASTRUCTURE ONLY (oracle).  Not part of the product path, not htslib code.
//
// A minimal IN-MEMORY stand-in for the handful of htslib 1.9 (pepper/modules/htslib.cmake:9) entry points the
// reference's BAM_handler (pepper/modules/src/dataio/bam_handler.cpp) calls, so that the UNMODIFIED get_reads() can be
// compiled and run here without htslib.  Record layout / flag bits / CIGAR encoding are the SAM/BAM specification's
// (SAMv1 section 4.2); the "file" is a vector of records the test shim installs before constructing BAM_handler, and
// the region iterator restates htslib's overlap rule: a record is returned iff
//        tid matches  &&  pos < end  &&  pos + (n_cigar ? reference length of the CIGAR : 1) > beg
// in file order (hts.c hts_itr_next / sam.c bam_readrec of htslib 1.9).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cerrno>
#include <string>
#include <vector>

#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048

#define BAM_CMATCH 0
#define BAM_CINS 1
#define BAM_CDEL 2
#define BAM_CREF_SKIP 3
#define BAM_CSOFT_CLIP 4
#define BAM_CHARD_CLIP 5
#define BAM_CPAD 6
#define BAM_CEQUAL 7
#define BAM_CDIFF 8
#define BAM_CBACK 9
#define bam_cigar_op(c) ((c) & 0xf)
#define bam_cigar_oplen(c) ((c) >> 4)

struct bam1_core_t {
    int32_t tid, pos;
    uint16_t bin;
    uint8_t qual, l_qname;
    uint16_t flag, unused1;
    uint32_t n_cigar;
    int32_t l_qseq, mtid, mpos, isize;
};
struct bam1_t {
    bam1_core_t core;
    int l_data;
    uint32_t m_data;
    uint8_t *data;      // qname | cigar | seq (4-bit) | qual | aux   (BAM record body, SAMv1 4.2)
};
#define bam_get_qname(b) ((char *) (b)->data)
#define bam_get_cigar(b) ((uint32_t *) ((b)->data + (b)->core.l_qname))
#define bam_get_seq(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname)
#define bam_get_qual(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1))
#define bam_get_aux(b) ((b)->data + ((b)->core.n_cigar << 2) + (b)->core.l_qname + (((b)->core.l_qseq + 1) >> 1) + (b)->core.l_qseq)
#define bam_seqi(s, i) ((s)[(i) >> 1] >> ((~(i) & 1) << 2) & 0xf)
static const char seq_nt16_str[] = "=ACMGRSVTWYHKDBN";

struct bam_hdr_t {
    int32_t n_targets;
    uint32_t l_text;
    uint32_t *target_len;
    char **target_name;
    char *text;
};

struct MemBamRecord { bam1_core_t core; std::vector<uint8_t> data; };
struct MemBam { std::vector<std::string> targets; std::vector<uint32_t> target_len; std::vector<MemBamRecord> records; };
inline MemBam *&membam_current() { static MemBam *p = nullptr; return p; }

struct htsFile { MemBam *mem; };
struct hts_idx_t { MemBam *mem; };
struct hts_itr_t { MemBam *mem; int tid; int64_t beg, end; size_t next; };
typedef htsFile samFile;

inline htsFile *sam_open(const char *, const char *) { return membam_current() ? new htsFile{membam_current()} : nullptr; }
inline int sam_close(htsFile *f) { delete f; return 0; }
inline hts_idx_t *sam_index_load(htsFile *f, const char *) { return new hts_idx_t{f->mem}; }
inline void hts_idx_destroy(hts_idx_t *i) { delete i; }
inline bam_hdr_t *sam_hdr_read(htsFile *f) {
    bam_hdr_t *h = new bam_hdr_t();
    h->n_targets = (int32_t) f->mem->targets.size();
    h->target_len = new uint32_t[h->n_targets + 1];
    h->target_name = new char *[h->n_targets + 1];
    for (int i = 0; i < h->n_targets; i++) {
        h->target_len[i] = f->mem->target_len[i];
        h->target_name[i] = strdup(f->mem->targets[i].c_str());
    }
    h->text = strdup("");
    h->l_text = 0;
    return h;
}
inline void bam_hdr_destroy(bam_hdr_t *h) {
    if (!h) return;
    for (int i = 0; i < h->n_targets; i++) free(h->target_name[i]);
    delete[] h->target_name; delete[] h->target_len; free(h->text); delete h;
}
inline int bam_name2id(bam_hdr_t *h, const char *name) {
    for (int i = 0; i < h->n_targets; i++) if (!strcmp(h->target_name[i], name)) return i;
    return -1;
}
inline bam1_t *bam_init1() { return new bam1_t(); }
inline void bam_destroy1(bam1_t *b) { delete b; }     // data is borrowed from the MemBam
inline hts_itr_t *sam_itr_queryi(hts_idx_t *idx, int tid, int64_t beg, int64_t end) { return new hts_itr_t{idx->mem, tid, beg, end, 0}; }
inline void hts_itr_destroy(hts_itr_t *it) { delete it; }
inline int64_t membam_rlen(const MemBamRecord &r) {
    if (r.core.n_cigar == 0) return 1;
    const uint32_t *c = (const uint32_t *) (r.data.data() + r.core.l_qname);
    int64_t l = 0;
    for (uint32_t k = 0; k < r.core.n_cigar; k++) {
        int op = bam_cigar_op(c[k]);
        if (op == BAM_CMATCH || op == BAM_CDEL || op == BAM_CREF_SKIP || op == BAM_CEQUAL || op == BAM_CDIFF) l += bam_cigar_oplen(c[k]);
    }
    return l;
}
inline int sam_itr_next(htsFile *, hts_itr_t *it, bam1_t *b) {
    while (it->next < it->mem->records.size()) {
        MemBamRecord &r = it->mem->records[it->next++];
        if (r.core.tid != it->tid) continue;
        if (r.core.pos >= it->end) return -1;                         // coordinate sorted: nothing further can overlap
        if ((int64_t) r.core.pos + membam_rlen(r) > it->beg) {
            b->core = r.core; b->data = r.data.data(); b->l_data = (int) r.data.size(); b->m_data = (uint32_t) r.data.size();
            return 1;
        }
    }
    return -1;
}
inline int64_t bam_aux2i(const uint8_t *s) {
    int type = *s++;
    switch (type) {
        case 'c': return (int8_t) s[0];
        case 'C': return s[0];
        case 's': { int16_t v; memcpy(&v, s, 2); return v; }
        case 'S': { uint16_t v; memcpy(&v, s, 2); return v; }
        case 'i': { int32_t v; memcpy(&v, s, 4); return v; }
        case 'I': { uint32_t v; memcpy(&v, s, 4); return v; }
    }
    errno = EINVAL;
    return 0;
}
