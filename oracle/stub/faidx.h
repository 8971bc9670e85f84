#pragma once
struct faidx_t;
