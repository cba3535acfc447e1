#pragma once
#include "common.cuh"
// 2-D row-major [rows, cols] tensor of 2-byte elements; box = [box_rows, 64 cols] (128 bytes), 128-byte swizzle
int cts_make_tmap_2d(cts_ctx* ctx, CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld_elems,
                     int box_rows, int is_bf16);
// same tensor, un-swizzled box = [box_rows, box_cols] (dense rows of box_cols elements in shared memory)
int cts_make_tmap_2d_dense(cts_ctx* ctx, CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld_elems,
                           int box_rows, int box_cols, int is_bf16);
