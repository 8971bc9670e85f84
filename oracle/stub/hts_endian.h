// TEST INFRASTRUCTURE ONLY (oracle): little-endian readers BAM_handler::get_reads uses on aux fields.
#pragma once
#include <cstdint>
#include <cstring>
static inline uint32_t le_to_u32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline float le_to_float(const uint8_t *p) { float v; memcpy(&v, p, 4); return v; }
