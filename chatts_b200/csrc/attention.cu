// chatts_b200 -- causal GQA attention (the arithmetic of transformers qwen2/modeling_qwen2.py:161-184, which the
// reference reaches through vLLM's Attention layer, chatts/vllm/chatts_vllm.py:595-598).
//
//  cts_attn_prefill : variable-length causal attention over the tokens of one call (no prefix).  Tensor-core
//                     (HMMA via nvcuda::wmma) flash attention in two passes over the KV tiles: pass 1 finds the
//                     row max / row sum, pass 2 accumulates exp(s - m) V in fragments, so no accumulator rescale
//                     is needed.  Round-1 kernel: prefill attention is ~1% of the prefill FLOPs at the benchmark
//                     prompt (576 tokens); the tcgen05/TMEM version is the round-2 item in DESIGN.md.
//  cts_attn_decode  : one query token per sequence against the PAGED KV cache, flash-decoding split over pages.
//                     HBM-bound (GQA intensity = nh/nkv flop/B): K/V pages are pulled with cp.async.bulk into a
//                     double-buffered shared-memory ring (mbarrier complete_tx), all q heads of a kv head share
//                     each page; partial (m, l, o) per split are merged by a second tiny kernel.
#include <mma.h>

#include "common.cuh"

namespace {

using namespace nvcuda;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem)), "l"(gmem), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T> struct WT;
template <> struct WT<__nv_bfloat16> { using type = __nv_bfloat16; };
template <> struct WT<__half> { using type = __half; };

// ================================================================================================
// prefill
// ================================================================================================
constexpr int kPfQ = 64;        // query rows per CTA (16 per warp)
constexpr int kPfKV = 64;       // kv tokens per tile
constexpr int kPfThreads = 128;

template <int HD> struct PfSmem {
  static constexpr int LD = HD + 8;                       // padded row, elements
  static constexpr int SLD = kPfKV + 8;                   // padded score row, floats
  static constexpr size_t q_bytes = (size_t)kPfQ * LD * 2;
  static constexpr size_t kv_bytes = (size_t)kPfKV * LD * 2;
  static constexpr size_t s_bytes = (size_t)4 * 16 * SLD * 4;   // per-warp 16 x 64 fp32
  static constexpr size_t p_bytes = (size_t)4 * 16 * SLD * 2;   // per-warp 16 x 64 model dtype
  static constexpr size_t o_bytes = (size_t)4 * 16 * (HD + 8) * 4;
  // layout: Q | K0 | V0 | K1 | V1 | S | P ;  the O staging reuses the K0.. region at the end
  static constexpr size_t total = q_bytes + 4 * kv_bytes + s_bytes + p_bytes;
};

template <typename T, int HD>
__device__ __forceinline__ void pf_load_tile(T* dst, const T* __restrict__ src, long long row_stride, int rows_valid) {
  // 64 rows x HD elements, 16 B per cp.async
  constexpr int LD = PfSmem<HD>::LD;
  constexpr int CH = HD / 8;
  for (int i = threadIdx.x; i < kPfKV * CH; i += kPfThreads) {
    const int r = i / CH, c = i % CH;
    const bool ok = r < rows_valid;
    cp_async16(dst + r * LD + c * 8, src + (ok ? (long long)r * row_stride + c * 8 : 0), ok);
  }
}

template <typename T, int HD>
__global__ void __launch_bounds__(kPfThreads)
attn_prefill_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                    const int* __restrict__ cu_seqlens, int nh, int nkv, float scale, T* __restrict__ out) {
  using SM = PfSmem<HD>;
  constexpr int LD = SM::LD, SLD = SM::SLD;
  extern __shared__ __align__(128) uint8_t pf_smem[];
  T* q_s = reinterpret_cast<T*>(pf_smem);
  T* kv_s = reinterpret_cast<T*>(pf_smem + SM::q_bytes);                       // [2][K|V][64][LD]
  float* s_all = reinterpret_cast<float*>(pf_smem + SM::q_bytes + 4 * SM::kv_bytes);
  T* p_all = reinterpret_cast<T*>(pf_smem + SM::q_bytes + 4 * SM::kv_bytes + SM::s_bytes);

  const int b = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int seq0 = cu_seqlens[b], len = cu_seqlens[b + 1] - seq0;
  const int q0 = qt * kPfQ;
  if (q0 >= len) return;
  const int kvh = head / (nh / nkv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long q_stride = (long long)nh * HD, kv_stride = (long long)nkv * HD;
  const T* q_g = q + ((long long)seq0 + q0) * q_stride + (long long)head * HD;
  const T* k_g = k + (long long)seq0 * kv_stride + (long long)kvh * HD;
  const T* v_g = v + (long long)seq0 * kv_stride + (long long)kvh * HD;

  const int n_tiles = (min(q0 + kPfQ, len) + kPfKV - 1) / kPfKV;   // causal: kv <= last q row of this CTA

  auto kbuf = [&](int st) { return kv_s + (size_t)st * 2 * kPfKV * LD; };
  auto vbuf = [&](int st) { return kv_s + (size_t)st * 2 * kPfKV * LD + (size_t)kPfKV * LD; };
  auto issue_tile = [&](int j, int st, bool with_v) {
    const int kv0 = j * kPfKV;
    const int valid = min(kPfKV, len - kv0);
    pf_load_tile<T, HD>(kbuf(st), k_g + (long long)kv0 * kv_stride, kv_stride, valid);
    if (with_v) pf_load_tile<T, HD>(vbuf(st), v_g + (long long)kv0 * kv_stride, kv_stride, valid);
    cp_async_commit();
  };

  // Q tile (rows beyond the sequence are zero-filled)
  pf_load_tile<T, HD>(q_s, q_g, q_stride, min(kPfQ, len - q0));
  cp_async_commit();
  issue_tile(0, 0, false);

  float* s_w = s_all + (size_t)warp * 16 * SLD;
  T* p_w = p_all + (size_t)warp * 16 * SLD;
  // row statistics: lanes (2r, 2r+1) own row r of this warp's 16 rows; each covers 32 of the 64 columns
  const int r_loc = lane >> 1, c_half = lane & 1;
  const int row_g = q0 + warp * 16 + r_loc;             // query index inside the sequence
  float m_run = -INFINITY, l_run = 0.f;
  const float sl2 = scale * 1.4426950408889634f;        // scores are used as s * scale * log2(e) inside exp2f

  wmma::fragment<wmma::matrix_a, 16, 16, 16, T, wmma::row_major> qf[HD / 16];

  // ------------------------------- pass 1: row max and row sum -------------------------------
  for (int j = 0; j < n_tiles; ++j) {
    const int st = j & 1;
    if (j + 1 < n_tiles) issue_tile(j + 1, st ^ 1, false);
    if (j + 1 < n_tiles) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) wmma::load_matrix_sync(qf[kk], q_s + (size_t)warp * 16 * LD + kk * 16, LD);
    }
    const int kv0 = j * kPfKV;
    if (kv0 <= q0 + warp * 16 + 15) {                   // warp-uniform: tile intersects this warp's causal range
#pragma unroll
      for (int nt = 0; nt < kPfKV / 16; ++nt) {
        wmma::fragment<wmma::accumulator, 16, 16, 16, float> sf;
        wmma::fill_fragment(sf, 0.f);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          wmma::fragment<wmma::matrix_b, 16, 16, 16, T, wmma::col_major> kf;
          wmma::load_matrix_sync(kf, kbuf(st) + (size_t)nt * 16 * LD + kk * 16, LD);
          wmma::mma_sync(sf, qf[kk], kf, sf);
        }
        wmma::store_matrix_sync(s_w + nt * 16, sf, SLD, wmma::mem_row_major);
      }
      __syncwarp();
      float mx = -INFINITY;
      for (int c = 0; c < 32; ++c) {
        const int col = c_half * 32 + c;
        const float s = (kv0 + col <= row_g) ? s_w[r_loc * SLD + col] : -INFINITY;
        mx = fmaxf(mx, s);
      }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      const float m_new = fmaxf(m_run, mx);
      float sum = 0.f;
      if (m_new > -INFINITY) {
        for (int c = 0; c < 32; ++c) {
          const int col = c_half * 32 + c;
          if (kv0 + col <= row_g) sum += exp2f((s_w[r_loc * SLD + col] - m_new) * sl2);
        }
      }
      sum += __shfl_xor_sync(0xffffffffu, sum, 1);
      l_run = (m_run > -INFINITY ? l_run * exp2f((m_run - m_new) * sl2) : 0.f) + sum;
      m_run = m_new;
      __syncwarp();
    }
    __syncthreads();
  }

  // ------------------------------- pass 2: O = sum exp(s - m) V -------------------------------
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> of[HD / 16];
#pragma unroll
  for (int i = 0; i < HD / 16; ++i) wmma::fill_fragment(of[i], 0.f);
  issue_tile(0, 0, true);
  for (int j = 0; j < n_tiles; ++j) {
    const int st = j & 1;
    if (j + 1 < n_tiles) issue_tile(j + 1, st ^ 1, true);
    if (j + 1 < n_tiles) cp_async_wait<1>(); else cp_async_wait<0>();
    __syncthreads();
    const int kv0 = j * kPfKV;
    if (kv0 <= q0 + warp * 16 + 15) {
#pragma unroll
      for (int nt = 0; nt < kPfKV / 16; ++nt) {
        wmma::fragment<wmma::accumulator, 16, 16, 16, float> sf;
        wmma::fill_fragment(sf, 0.f);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          wmma::fragment<wmma::matrix_b, 16, 16, 16, T, wmma::col_major> kf;
          wmma::load_matrix_sync(kf, kbuf(st) + (size_t)nt * 16 * LD + kk * 16, LD);
          wmma::mma_sync(sf, qf[kk], kf, sf);
        }
        wmma::store_matrix_sync(s_w + nt * 16, sf, SLD, wmma::mem_row_major);
      }
      __syncwarp();
      for (int c = 0; c < 32; ++c) {
        const int col = c_half * 32 + c;
        float p = 0.f;
        if (kv0 + col <= row_g && m_run > -INFINITY) p = exp2f((s_w[r_loc * SLD + col] - m_run) * sl2);
        p_w[r_loc * SLD + col] = DT<T>::from_f(p);
      }
      __syncwarp();
#pragma unroll
      for (int kt = 0; kt < kPfKV / 16; ++kt) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, T, wmma::row_major> pf;
        wmma::load_matrix_sync(pf, p_w + kt * 16, SLD);
#pragma unroll
        for (int dn = 0; dn < HD / 16; ++dn) {
          wmma::fragment<wmma::matrix_b, 16, 16, 16, T, wmma::row_major> vf;
          wmma::load_matrix_sync(vf, vbuf(st) + (size_t)kt * 16 * LD + dn * 16, LD);
          wmma::mma_sync(of[dn], pf, vf, of[dn]);
        }
      }
      __syncwarp();
    }
    __syncthreads();
  }

  // ------------------------------- epilogue: O / l -> out -------------------------------
  // stage this warp's 16 x HD fp32 block through shared memory (the KV ring is free now)
  float* o_w = reinterpret_cast<float*>(pf_smem + SM::q_bytes) + (size_t)warp * 16 * (HD + 8);
#pragma unroll
  for (int dn = 0; dn < HD / 16; ++dn) wmma::store_matrix_sync(o_w + dn * 16, of[dn], HD + 8, wmma::mem_row_major);
  __syncwarp();
  if (row_g < len) {
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
    T* o_g = out + ((long long)seq0 + row_g) * q_stride + (long long)head * HD;
    for (int c = 0; c < HD / 2; c += 8) {
      const int col = c_half * (HD / 2) + c;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = o_w[r_loc * (HD + 8) + col + e] * inv;
      *reinterpret_cast<uint4*>(o_g + col) = pack8<T>(f);
    }
  }
}

// ================================================================================================
// decode
// ================================================================================================
constexpr int kDecThreads = 128;
constexpr int kMaxGroup = 8;    // q heads per kv head

// workspace per (batch, head, split): [m, l, o[HD]]
__host__ __device__ inline long long dec_ws_stride(int hd) { return hd + 2; }

template <typename T, int HD>
__global__ void __launch_bounds__(kDecThreads)
attn_decode_kernel(const T* __restrict__ q, const T* __restrict__ k_cache, const T* __restrict__ v_cache,
                   const int* __restrict__ page_table, int max_pages, const int* __restrict__ seq_lens, int nh, int nkv,
                   int page_size, float scale, int num_splits, float* __restrict__ ws) {
  extern __shared__ __align__(128) uint8_t dec_smem[];
  __shared__ uint64_t bar[2];
  __shared__ float q_s[kMaxGroup][HD];
  __shared__ float m_s[kMaxGroup], l_s[kMaxGroup], alpha_s[kMaxGroup];

  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int G = nh / nkv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int seq_len = seq_lens[b];
  const int n_pages = (seq_len + page_size - 1) / page_size;
  const int pps = (n_pages + num_splits - 1) / num_splits;
  const int pg0 = split * pps, pg1 = min(n_pages, pg0 + pps);

  const size_t tile_elems = (size_t)page_size * HD;
  T* k_s = reinterpret_cast<T*>(dec_smem);                       // [2][page][HD]
  T* v_s = k_s + 2 * tile_elems;                                  // [2][page][HD]
  float* sc = reinterpret_cast<float*>(v_s + 2 * tile_elems);     // [page][kMaxGroup]

  // out accumulators: thread owns dim pair dp and a token stripe ts of every page
  constexpr int NP = HD / 2;                 // dim pairs
  constexpr int NSTRIPE = kDecThreads / NP;  // 2 for HD=128, 4 for HD=64
  const int dp = tid % NP, stripe = tid / NP;
  float acc[kMaxGroup][2];
#pragma unroll
  for (int g = 0; g < kMaxGroup; ++g) acc[g][0] = acc[g][1] = 0.f;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    fence_mbar_init();
  }
  for (int i = tid; i < G * HD; i += kDecThreads) {
    const int g = i / HD, d = i % HD;
    q_s[g][d] = DT<T>::to_f(q[((long long)b * nh + (long long)kvh * G + g) * HD + d]);
  }
  if (tid < kMaxGroup) { m_s[tid] = -INFINITY; l_s[tid] = 0.f; }
  __syncthreads();

  const uint32_t tile_bytes = (uint32_t)(tile_elems * sizeof(T));
  auto issue = [&](int pg, int st) {
    const long long page = page_table[(long long)b * max_pages + pg];
    const T* kg = k_cache + ((page * nkv + kvh) * (long long)page_size) * HD;
    const T* vg = v_cache + ((page * nkv + kvh) * (long long)page_size) * HD;
    mbar_expect_tx(&bar[st], 2 * tile_bytes);
    bulk_load_1d(k_s + (size_t)st * tile_elems, kg, tile_bytes, &bar[st]);
    bulk_load_1d(v_s + (size_t)st * tile_elems, vg, tile_bytes, &bar[st]);
  };
  if (tid == 0 && pg0 < pg1) issue(pg0, 0);

  const int n_groups = kDecThreads / page_size;    // thread groups over tokens (page_size in {16,32,64,128})
  const int tk = tid % page_size, hh = tid / page_size;
  constexpr int NCH = HD / 8;

  for (int pg = pg0; pg < pg1; ++pg) {
    const int it = pg - pg0, st = it & 1;
    if (tid == 0 && pg + 1 < pg1) {                           // the other buffer was released by the barrier below
      fence_proxy_async_smem();
      issue(pg + 1, st ^ 1);
    }
    mbar_wait(&bar[st], (uint32_t)(it >> 1) & 1u);
    const T* kt = k_s + (size_t)st * tile_elems;
    const T* vt = v_s + (size_t)st * tile_elems;
    // ---- scores: thread (tk, hh) -> heads hh, hh + n_groups, ...
    {
      float dot[4] = {0.f, 0.f, 0.f, 0.f};
      const uint4* krow = reinterpret_cast<const uint4*>(kt + (size_t)tk * HD);
#pragma unroll 4
      for (int c = 0; c < NCH; ++c) {
        const int ch = (c + tk) % NCH;                // rotate the chunk order: conflict-free 16 B reads
        float kf[8];
        unpack8<T>(krow[ch], kf);
#pragma unroll
        for (int slot = 0; slot < 4; ++slot) {
          const int g = hh + slot * n_groups;
          if (g < G) {
            const float4 qa = *reinterpret_cast<const float4*>(&q_s[g][ch * 8]);
            const float4 qb = *reinterpret_cast<const float4*>(&q_s[g][ch * 8 + 4]);
            dot[slot] += kf[0] * qa.x + kf[1] * qa.y + kf[2] * qa.z + kf[3] * qa.w + kf[4] * qb.x + kf[5] * qb.y +
                         kf[6] * qb.z + kf[7] * qb.w;
          }
        }
      }
      const bool valid = pg * page_size + tk < seq_len;
#pragma unroll
      for (int slot = 0; slot < 4; ++slot) {
        const int g = hh + slot * n_groups;
        if (g < G) sc[tk * kMaxGroup + g] = valid ? dot[slot] * scale : -INFINITY;
      }
    }
    __syncthreads();
    // ---- online softmax: warp w handles heads w, w+4
    for (int g = warp; g < G; g += kDecThreads / 32) {
      float mx = -INFINITY;
      for (int t = lane; t < page_size; t += 32) mx = fmaxf(mx, sc[t * kMaxGroup + g]);
      mx = warp_max(mx);
      const float m_old = m_s[g];
      const float m_new = fmaxf(m_old, mx);
      float sum = 0.f;
      for (int t = lane; t < page_size; t += 32) {
        const float s = sc[t * kMaxGroup + g];
        const float p = (m_new > -INFINITY) ? __expf(s - m_new) : 0.f;
        sc[t * kMaxGroup + g] = p;
        sum += p;
      }
      sum = warp_sum(sum);
      if (lane == 0) {
        const float a = (m_old > -INFINITY) ? __expf(m_old - m_new) : 0.f;
        alpha_s[g] = a;
        l_s[g] = l_s[g] * a + sum;
        m_s[g] = m_new;
      }
    }
    __syncthreads();
    // ---- O update: thread (dp, stripe)
    {
#pragma unroll
      for (int g = 0; g < kMaxGroup; ++g) {
        if (g < G) { const float a = alpha_s[g]; acc[g][0] *= a; acc[g][1] *= a; }
      }
      for (int t = stripe; t < page_size; t += NSTRIPE) {
        const uint32_t vraw = *reinterpret_cast<const uint32_t*>(vt + (size_t)t * HD + 2 * dp);
        const T* vp = reinterpret_cast<const T*>(&vraw);
        const float v0 = DT<T>::to_f(vp[0]), v1 = DT<T>::to_f(vp[1]);
        const float4 pa = *reinterpret_cast<const float4*>(&sc[t * kMaxGroup]);
        const float4 pb = *reinterpret_cast<const float4*>(&sc[t * kMaxGroup + 4]);
        const float pr[8] = {pa.x, pa.y, pa.z, pa.w, pb.x, pb.y, pb.z, pb.w};
#pragma unroll
        for (int g = 0; g < kMaxGroup; ++g) {
          if (g < G) { acc[g][0] += pr[g] * v0; acc[g][1] += pr[g] * v1; }
        }
      }
    }
    __syncthreads();    // everyone is done with buffer st and with sc before they are overwritten
  }

  // ---- merge the token stripes and write this split's partial (m, l, o)
  float* red = reinterpret_cast<float*>(dec_smem);      // reuse: [NSTRIPE][G][HD]
  for (int g = 0; g < G; ++g) {
    red[((size_t)stripe * kMaxGroup + g) * HD + 2 * dp] = acc[g][0];
    red[((size_t)stripe * kMaxGroup + g) * HD + 2 * dp + 1] = acc[g][1];
  }
  __syncthreads();
  const long long wstride = dec_ws_stride(HD);
  for (int i = tid; i < G * HD; i += kDecThreads) {
    const int g = i / HD, d = i % HD;
    float o = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < NSTRIPE; ++s2) o += red[((size_t)s2 * kMaxGroup + g) * HD + d];
    float* w = ws + (((long long)b * nh + (long long)kvh * G + g) * num_splits + split) * wstride;
    w[2 + d] = o;
    if (d == 0) { w[0] = m_s[g]; w[1] = l_s[g]; }
  }
}

template <typename T, int HD>
__global__ void attn_decode_combine_kernel(const float* __restrict__ ws, int num_splits, T* __restrict__ out) {
  // grid (nh, batch); block HD threads
  const long long bh = (long long)blockIdx.y * gridDim.x + blockIdx.x;
  const int d = threadIdx.x;
  const long long wstride = dec_ws_stride(HD);
  const float* w = ws + bh * num_splits * wstride;
  float m = -INFINITY;
  for (int s = 0; s < num_splits; ++s) m = fmaxf(m, w[s * wstride]);
  float l = 0.f, o = 0.f;
  for (int s = 0; s < num_splits; ++s) {
    const float ms = w[s * wstride];
    const float sc = (ms > -INFINITY) ? __expf(ms - m) : 0.f;
    l += w[s * wstride + 1] * sc;
    o += w[s * wstride + 2 + d] * sc;
  }
  out[bh * HD + d] = DT<T>::from_f(l > 0.f ? o / l : 0.f);
}

}  // namespace

extern "C" int cts_attn_prefill(cts_ctx* ctx, const void* q, const void* k, const void* v, const int* cu_seqlens, int batch,
                                int max_seqlen, int nh, int nkv, int head_dim, float scale, void* out, int dtype,
                                void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, q && k && v && cu_seqlens && out, "null pointer");
  CTS_CHECK_ARG(ctx, nh > 0 && nkv > 0 && nh % nkv == 0, "nh must be a positive multiple of nkv");
  CTS_CHECK_ARG(ctx, dtype == CTS_BF16 || dtype == CTS_F16, "dtype");
  if (head_dim != 64 && head_dim != 128) return cts_set_error(ctx, CTS_ERR_UNSUPPORTED, "cts_attn_prefill: head_dim %d (64 or 128)", head_dim);
  if (batch == 0 || max_seqlen == 0) return CTS_OK;
  CTS_CHECK_ARG(ctx, batch <= 65535 && nh <= 65535, "grid");
  dim3 grid((unsigned)((max_seqlen + kPfQ - 1) / kPfQ), (unsigned)nh, (unsigned)batch);
  cudaStream_t st = (cudaStream_t)stream;
#define PF_LAUNCH(TT, HDV)                                                                                   \
  {                                                                                                          \
    auto kern = attn_prefill_kernel<TT, HDV>;                                                                \
    const size_t smem = PfSmem<HDV>::total;                                                                  \
    CTS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    kern<<<grid, kPfThreads, smem, st>>>((const TT*)q, (const TT*)k, (const TT*)v, cu_seqlens, nh, nkv, scale, (TT*)out); \
  }
  if (dtype == CTS_BF16) {
    if (head_dim == 128) PF_LAUNCH(__nv_bfloat16, 128) else PF_LAUNCH(__nv_bfloat16, 64)
  } else {
    if (head_dim == 128) PF_LAUNCH(__half, 128) else PF_LAUNCH(__half, 64)
  }
#undef PF_LAUNCH
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

extern "C" long long cts_attn_decode_workspace_floats(int batch, int nh, int head_dim, int num_splits) {
  return (long long)batch * nh * num_splits * dec_ws_stride(head_dim);
}

extern "C" int cts_attn_decode(cts_ctx* ctx, const void* q, const void* k_cache, const void* v_cache, const int* page_table,
                               int max_pages, const int* seq_lens, int batch, int nh, int nkv, int head_dim, int page_size,
                               float scale, int num_splits, float* workspace, void* out, int dtype, void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, q && k_cache && v_cache && page_table && seq_lens && workspace && out, "null pointer");
  CTS_CHECK_ARG(ctx, nh > 0 && nkv > 0 && nh % nkv == 0 && nh / nkv <= kMaxGroup, "nh/nkv must be an integer <= 8");
  CTS_CHECK_ARG(ctx, page_size == 16 || page_size == 32 || page_size == 64 || page_size == 128, "page_size must be 16, 32, 64 or 128");
  CTS_CHECK_ARG(ctx, num_splits >= 1 && num_splits <= 65535, "num_splits");
  CTS_CHECK_ARG(ctx, dtype == CTS_BF16 || dtype == CTS_F16, "dtype");
  if (head_dim != 64 && head_dim != 128) return cts_set_error(ctx, CTS_ERR_UNSUPPORTED, "cts_attn_decode: head_dim %d (64 or 128)", head_dim);
  // 4 score slots per thread must cover all q heads of a kv head
  CTS_CHECK_ARG(ctx, (nh / nkv) <= 4 * (kDecThreads / page_size), "page_size too large for this GQA group");
  if (batch == 0) return CTS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)num_splits, (unsigned)nkv, (unsigned)batch);
  dim3 cgrid((unsigned)nh, (unsigned)batch);
  size_t smem = (size_t)4 * page_size * head_dim * 2 + (size_t)page_size * kMaxGroup * 4;
  const size_t red_bytes = (size_t)(kDecThreads / (head_dim / 2)) * kMaxGroup * head_dim * 4;
  if (smem < red_bytes) smem = red_bytes;
#define DEC_LAUNCH(TT, HDV)                                                                                    \
  {                                                                                                            \
    auto kern = attn_decode_kernel<TT, HDV>;                                                                   \
    CTS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    kern<<<grid, kDecThreads, smem, st>>>((const TT*)q, (const TT*)k_cache, (const TT*)v_cache, page_table, max_pages, \
                                          seq_lens, nh, nkv, page_size, scale, num_splits, workspace);        \
    CTS_LAUNCH_CHECK(ctx);                                                                                     \
    attn_decode_combine_kernel<TT, HDV><<<cgrid, HDV, 0, st>>>(workspace, num_splits, (TT*)out);               \
  }
  if (dtype == CTS_BF16) {
    if (head_dim == 128) DEC_LAUNCH(__nv_bfloat16, 128) else DEC_LAUNCH(__nv_bfloat16, 64)
  } else {
    if (head_dim == 128) DEC_LAUNCH(__half, 128) else DEC_LAUNCH(__half, 64)
  }
#undef DEC_LAUNCH
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}
