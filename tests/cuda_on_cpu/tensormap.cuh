// TEST INFRASTRUCTURE ONLY -- tensor maps for the "CUDA on CPU" shim: the description of the tensor and of the box (see common.cuh)
#pragma once
#include "common.cuh"
// the argument checks of the real encoder (ctx.cu): 16-byte aligned base and row pitch
static inline int shim_tmap_check(cts_ctx* ctx, const void* base, long long ld_elems, int box_rows) {
  if (((uintptr_t)base & 15) != 0) return cts_set_error(ctx, CTS_ERR_BAD_ARG, "tensor map: base not 16-byte aligned");
  if ((ld_elems * 2) % 16 != 0) return cts_set_error(ctx, CTS_ERR_BAD_ARG, "tensor map: row pitch %lld B not a multiple of 16", ld_elems * 2);
  if (box_rows < 1 || box_rows > 256) return cts_set_error(ctx, CTS_ERR_BAD_ARG, "tensor map: box rows %d", box_rows);
  return CTS_OK;
}
static inline int cts_make_tmap_2d(cts_ctx* ctx, CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld_elems, int box_rows,
                                   int is_bf16) {
  if (int rc = shim_tmap_check(ctx, base, ld_elems, box_rows)) return rc;
  *tm = CUtensorMap{base, rows, cols, ld_elems, box_rows, 64, 1, is_bf16};
  return CTS_OK;
}
static inline int cts_make_tmap_2d_dense(cts_ctx* ctx, CUtensorMap* tm, const void* base, long long rows, long long cols, long long ld_elems,
                                         int box_rows, int box_cols, int is_bf16) {
  if (int rc = shim_tmap_check(ctx, base, ld_elems, box_rows)) return rc;
  *tm = CUtensorMap{base, rows, cols, ld_elems, box_rows, box_cols, 0, is_bf16};
  return CTS_OK;
}
