// chatts_b200 -- assembly of one TS patch row (chatts/vllm/chatts_vllm.py:107-183), shared by the stand-alone patchify kernel
// (csrc/ts_frontend.cu) and the fused encoder (csrc/ts_encoder_fused.cu).  A row is produced in 16-byte chunks so that every global
// store -- and, for the position-embedding part, every gather from the table -- is one 128-bit access.
#pragma once
#include "common.cuh"

// One 16-byte chunk (8 elements from column j8) of the patch row (series row `xrow`, first point p0):
//   mode 0: the patch values; mode 1: [patch values | patch x emb_dim position embeddings (point-major)], padding id = max_seq_len
//   (:76,:128,:163-182); mode 2: (value, position / max(1, max_valid - 1)) pairs, padding position -1 (:145-154).
// Values past the valid length repeat the LAST valid value (:121-125).
template <typename T>
__device__ __forceinline__ uint4 ts_row_chunk(const T* __restrict__ xrow, int nf, int patch, int mode, const T* __restrict__ pos_table,
                                              int emb_dim, int max_seq_len, int vl, int p0, int j8, float denom) {
  if (mode == 1 && j8 >= patch && (emb_dim & 7) == 0) {          // 8 consecutive embedding dims of ONE point: a 16-byte gather
    const int e = j8 - patch;
    const int pt = p0 + e / emb_dim;
    const int id = pt < vl ? (pt < max_seq_len ? pt : max_seq_len) : max_seq_len;
    return *reinterpret_cast<const uint4*>(pos_table + (size_t)id * emb_dim + (e % emb_dim));
  }
  T v[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int col = j8 + u;
    if (mode == 2) {
      const int pt = p0 + (col >> 1);
      if ((col & 1) == 0) v[u] = xrow[(size_t)(pt < vl ? pt : vl - 1) * nf];
      else v[u] = DT<T>::from_f(pt < vl ? (float)pt / denom : -1.0f);
    } else if (col < patch) {
      const int pt = p0 + col;
      v[u] = xrow[(size_t)(pt < vl ? pt : vl - 1) * nf];
    } else {                                                     // mode 1, embedding dims not a multiple of 8: element by element
      const int e = col - patch;
      const int pt = p0 + e / emb_dim;
      const int id = pt < vl ? (pt < max_seq_len ? pt : max_seq_len) : max_seq_len;
      v[u] = pos_table[(size_t)id * emb_dim + (e % emb_dim)];
    }
  }
  return *reinterpret_cast<const uint4*>(v);
}

