// chatts_b200 -- TS-encoder front end: mask -> valid length -> patch count -> patch rows.
//
// Replaces TimeSeriesEmbedding.forward up to the MLP call (chatts/vllm/chatts_vllm.py:94-183) and
// get_patch_cnt (:198-207).  The reference walks the series in a Python loop with 2-3 device->host syncs
// and ~10 tiny launches PER SERIES (:107-158); here it is two launches for the whole batch and no sync:
//   1. ts_count_kernel : one CTA per series, vectorised read of the interleaved (value, mask) row, warp-shuffle
//                        reduction of long(mask) -> valid_len, patch_cnt                       (:98-100)
//      ts_scan_kernel  : exclusive scan of patch_cnt -> first output row of each series (row order :187), max(valid)
//   2. ts_patchify_kernel : one CTA per (series, patch): stages the patch's 16 values in shared memory,
//                        pads with the LAST VALID value (:121-125), gathers the position embedding of every
//                        point (padding id = max_sequence_length, :76,:128) and writes the [P, in0] row coalesced.
// HBM-bound: bytes = N*2L*elt (read once) + sum(P)*in0*elt (written once) + the touched pos-table rows.
#include "common.cuh"
#include "ts_rows.cuh"

namespace {

template <typename T>
__global__ void ts_count_kernel(const T* __restrict__ x, int row_len, int nf, int patch, int* __restrict__ valid_len,
                                int* __restrict__ patch_cnt) {
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.x;
  const T* row = x + (size_t)s * row_len;
  const int npts = row_len / nf;
  long long acc = 0;
  if (nf == 2 && (row_len % 8) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0)) {
    // 8 elements = 4 (value, mask) pairs per 16-byte load
    const uint4* v = reinterpret_cast<const uint4*>(row);
    for (int i = threadIdx.x; i < row_len / 8; i += blockDim.x) {
      uint4 u = __ldg(v + i);
      const T* e = reinterpret_cast<const T*>(&u);
#pragma unroll
      for (int j = 1; j < 8; j += 2) acc += (long long)DT<T>::to_f(e[j]);   // .long() truncates toward zero (:98)
    }
  } else {
    for (int i = threadIdx.x; i < npts; i += blockDim.x) acc += (long long)DT<T>::to_f(row[(size_t)i * nf + (nf - 1)]);
  }
  // block reduction (values are small non-negative integers: use int shuffles)
  int a = (int)acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  __shared__ int wsum[32];
  if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x < 32) {
    int v = threadIdx.x < (blockDim.x >> 5) ? wsum[threadIdx.x] : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (threadIdx.x == 0) {
      valid_len[s] = v;
      patch_cnt[s] = (v + patch - 1) / patch;
    }
  }
}

__global__ void ts_scan_kernel(const int* __restrict__ valid_len, const int* __restrict__ patch_cnt, int n,
                               int* __restrict__ row_offset, int* __restrict__ max_valid) {
  pdl_trigger();
  pdl_wait();
  // single CTA, chunked inclusive scan via warp shuffles
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  __shared__ int maxv_s;
  if (threadIdx.x == 0) { carry_s = 0; maxv_s = 0; }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  int local_max = 0;
  for (int base = 0; base < n; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const int v = i < n ? patch_cnt[i] : 0;
    if (i < n) local_max = max(local_max, valid_len[i]);
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    int woff = 0;
    for (int w = 0; w < warp; ++w) woff += warp_tot[w];
    const int carry = carry_s;
    if (i < n) row_offset[i] = carry + woff + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = carry + woff + incl;
    __syncthreads();
  }
  (void)nw;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local_max = max(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
  if (lane == 0) atomicMax(&maxv_s, local_max);
  __syncthreads();
  if (threadIdx.x == 0) {
    row_offset[n] = carry_s;
    max_valid[0] = maxv_s;
  }
}

// one CTA per (patch, series); in0 % 8 == 0: the row goes out in 16-byte chunks (ts_rows.cuh), else element by element
template <typename T>
__global__ void ts_patchify_kernel(const T* __restrict__ x, int row_len, int nf, int patch, int mode,
                                   const T* __restrict__ pos_table, int emb_dim, int max_seq_len,
                                   const int* __restrict__ valid_len, const int* __restrict__ row_offset,
                                   const int* __restrict__ max_valid, T* __restrict__ rows_out, int in0) {
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.y;
  const int pidx = blockIdx.x;
  const int vl = valid_len[s];
  const int cnt = (vl + patch - 1) / patch;
  if (pidx >= cnt) return;
  const T* row = x + (size_t)s * row_len;
  const int p0 = pidx * patch;
  T* out = rows_out + (size_t)(row_offset[s] + pidx) * in0;
  const float denom = (float)max(1, max_valid[0] - 1);
  if ((in0 & 7) == 0 && (((uintptr_t)rows_out) & 15) == 0 && (mode != 1 || (((uintptr_t)pos_table) & 15) == 0)) {
    for (int j = threadIdx.x; j < (in0 >> 3); j += blockDim.x)
      *reinterpret_cast<uint4*>(out + j * 8) = ts_row_chunk<T>(row, nf, patch, mode, pos_table, emb_dim, max_seq_len, vl, p0, j * 8, denom);
    return;
  }
  for (int col = threadIdx.x; col < in0; col += blockDim.x) {
    T v;
    if (mode == 2) {
      const int pt = p0 + (col >> 1);
      v = (col & 1) == 0 ? row[(size_t)(pt < vl ? pt : vl - 1) * nf] : DT<T>::from_f(pt < vl ? (float)pt / denom : -1.0f);
    } else if (col < patch) {
      const int pt = p0 + col;
      v = row[(size_t)(pt < vl ? pt : vl - 1) * nf];           // pad with the last valid value (:121-125)
    } else {
      const int e = col - patch;
      const int pt = p0 + e / emb_dim;
      // padding id (:76,:128); a point index beyond the table is rejected on the host (IndexError, as nn.Embedding raises) and clamped here
      const int id = pt < vl ? min(pt, max_seq_len) : max_seq_len;
      v = pos_table[(size_t)id * emb_dim + (e % emb_dim)];
    }
    out[col] = v;
  }
}

}  // namespace

extern "C" int cts_ts_patch_count(cts_ctx* ctx, const void* x, int dtype, int n_series, int row_len, int num_features,
                                  int patch_size, int* valid_len, int* patch_cnt, int* row_offset, int* max_valid,
                                  void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, n_series >= 0 && row_len >= 0, "negative sizes");
  CTS_CHECK_ARG(ctx, num_features >= 1 && patch_size >= 1, "num_features / patch_size must be >= 1");
  CTS_CHECK_ARG(ctx, row_len % num_features == 0, "row_len not a multiple of num_features");
  CTS_CHECK_ARG(ctx, valid_len && patch_cnt && row_offset && max_valid, "null output");
  CTS_CHECK_ARG(ctx, dtype == CTS_BF16 || dtype == CTS_F16, "dtype");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_series > 0) {
    CTS_CHECK_ARG(ctx, x != nullptr, "null x");
    if (dtype == CTS_BF16)
      CTS_CUDA(ctx, launch_pdl(ts_count_kernel<__nv_bfloat16>, dim3(n_series), dim3(128), 0, st, 1, (const __nv_bfloat16*)x, row_len,
                               num_features, patch_size, valid_len, patch_cnt));
    else
      CTS_CUDA(ctx, launch_pdl(ts_count_kernel<__half>, dim3(n_series), dim3(128), 0, st, 1, (const __half*)x, row_len, num_features,
                               patch_size, valid_len, patch_cnt));
  }
  CTS_CUDA(ctx, launch_pdl(ts_scan_kernel, dim3(1), dim3(256), 0, st, 1, (const int*)valid_len, (const int*)patch_cnt, n_series,
                           row_offset, max_valid));
  return CTS_OK;
}

extern "C" int cts_ts_patchify(cts_ctx* ctx, const void* x, int dtype, int n_series, int row_len, int num_features,
                               int patch_size, int mode, const void* pos_table, int emb_dim, int max_seq_len,
                               const int* valid_len, const int* row_offset, const int* max_valid, int max_patches,
                               void* rows_out, int in0, void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, mode >= 0 && mode <= 2, "mode must be 0, 1 or 2");
  CTS_CHECK_ARG(ctx, dtype == CTS_BF16 || dtype == CTS_F16, "dtype");
  CTS_CHECK_ARG(ctx, num_features >= 1 && patch_size >= 1 && row_len % num_features == 0, "shape");
  const int expect = mode == 0 ? patch_size : mode == 1 ? patch_size * (1 + emb_dim) : 2 * patch_size;
  CTS_CHECK_ARG(ctx, in0 == expect, "in0 does not match mode / patch_size / emb_dim");
  CTS_CHECK_ARG(ctx, mode != 1 || (pos_table != nullptr && emb_dim >= 1), "mode 1 needs pos_table");
  CTS_CHECK_ARG(ctx, valid_len && row_offset && max_valid, "null metadata");
  if (n_series == 0 || max_patches == 0) return CTS_OK;
  CTS_CHECK_ARG(ctx, x && rows_out, "null x / rows_out");
  CTS_CHECK_ARG(ctx, n_series <= 65535, "more than 65535 series in one call");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)max_patches, (unsigned)n_series);
  const int threads = in0 >= 256 ? 128 : 64;
  const size_t smem = 0;
  if (dtype == CTS_BF16)
    CTS_CUDA(ctx, launch_pdl(ts_patchify_kernel<__nv_bfloat16>, grid, dim3(threads), smem, st, 1, (const __nv_bfloat16*)x, row_len,
                             num_features, patch_size, mode, (const __nv_bfloat16*)pos_table, emb_dim, max_seq_len, valid_len,
                             row_offset, max_valid, (__nv_bfloat16*)rows_out, in0));
  else
    CTS_CUDA(ctx, launch_pdl(ts_patchify_kernel<__half>, grid, dim3(threads), smem, st, 1, (const __half*)x, row_len, num_features,
                             patch_size, mode, (const __half*)pos_table, emb_dim, max_seq_len, valid_len, row_offset, max_valid,
                             (__half*)rows_out, in0));
  return CTS_OK;
}
