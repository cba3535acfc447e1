#pragma once
#include "common.cuh"
int cts_make_tmap_2d(cts_ctx* ctx, CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld_elems,
                     int box_rows, int is_bf16);
