// TEST INFRASTRUCTURE ONLY -- tensor maps for the "CUDA on CPU" shim: just the description of the tensor (the shimmed mainloop reads it directly)
#pragma once
#include "common.cuh"
static inline int cts_make_tmap_2d(cts_ctx*, CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld_elems, int box_rows,
                                   int is_bf16) {
  *tm = CUtensorMap{base, rows, cols, ld_elems, box_rows, is_bf16};
  return CTS_OK;
}
