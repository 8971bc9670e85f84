#pragma once
#include <cstdint>
struct htsFile; struct hts_idx_t; struct bam_hdr_t;
