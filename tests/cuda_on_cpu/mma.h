// TEST INFRASTRUCTURE ONLY -- nvcuda::wmma for the "CUDA on CPU" shim.  The fragment layout is opaque in CUDA; here every lane simply
// holds the WHOLE 16x16 tile and performs the whole operation (32-fold redundant, and therefore independent of the other lanes):
// load / mma / store have exactly the API semantics, fp32 accumulation in k order.
#pragma once
#include "common.cuh"
namespace nvcuda { namespace wmma {
struct matrix_a {}; struct matrix_b {}; struct accumulator {};
struct row_major {}; struct col_major {};
enum layout_t { mem_row_major, mem_col_major };
template <typename Use, int M, int N, int K, typename T, typename Layout = void> struct fragment {
  static_assert(M == 16 && N == 16 && K == 16, "m16n16k16 only");
  float m[256];
};
template <typename F> static inline void fill_fragment(F& f, float v) { for (int i = 0; i < 256; ++i) f.m[i] = v; }
template <typename T, typename L> static inline void load_matrix_sync(fragment<matrix_a, 16, 16, 16, T, L>& f, const T* p, unsigned ldm) {
  for (int i = 0; i < 16; ++i)
    for (int k = 0; k < 16; ++k) f.m[i * 16 + k] = DT<T>::to_f(std::is_same<L, row_major>::value ? p[i * ldm + k] : p[k * ldm + i]);
}
template <typename T, typename L> static inline void load_matrix_sync(fragment<matrix_b, 16, 16, 16, T, L>& f, const T* p, unsigned ldm) {
  for (int k = 0; k < 16; ++k)
    for (int n = 0; n < 16; ++n) f.m[k * 16 + n] = DT<T>::to_f(std::is_same<L, row_major>::value ? p[k * ldm + n] : p[n * ldm + k]);
}
template <typename T, typename LA, typename LB>
static inline void mma_sync(fragment<accumulator, 16, 16, 16, float>& d, const fragment<matrix_a, 16, 16, 16, T, LA>& a,
                            const fragment<matrix_b, 16, 16, 16, T, LB>& b, const fragment<accumulator, 16, 16, 16, float>& c) {
  float out[256];
  for (int i = 0; i < 16; ++i)
    for (int n = 0; n < 16; ++n) {
      float s = c.m[i * 16 + n];
      for (int k = 0; k < 16; ++k) s += a.m[i * 16 + k] * b.m[k * 16 + n];
      out[i * 16 + n] = s;
    }
  for (int i = 0; i < 256; ++i) d.m[i] = out[i];
}
static inline void store_matrix_sync(float* p, const fragment<accumulator, 16, 16, 16, float>& f, unsigned ldm, layout_t l) {
  for (int i = 0; i < 16; ++i)
    for (int n = 0; n < 16; ++n) (l == mem_row_major ? p[i * ldm + n] : p[n * ldm + i]) = f.m[i * 16 + n];
}
}}  // namespace nvcuda::wmma
