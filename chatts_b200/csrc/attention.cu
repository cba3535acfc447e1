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

#include <type_traits>

#include "common.cuh"
#include "trace.cuh"
#include "tensormap.cuh"

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

// LSE = true (the training forward, cts_attn_prefill_lse) adds one store per query row: lse[token][head] = log sum_j exp(scale s_j).
template <typename T, int HD, bool LSE>
__global__ void __launch_bounds__(kPfThreads)
attn_prefill_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                    const int* __restrict__ cu_seqlens, int nh, int nkv, float scale, T* __restrict__ out, float* __restrict__ lse) {
  using SM = PfSmem<HD>;
  constexpr int LD = SM::LD, SLD = SM::SLD;
  extern __shared__ __align__(128) uint8_t pf_smem[];
  T* q_s = reinterpret_cast<T*>(pf_smem);
  T* kv_s = reinterpret_cast<T*>(pf_smem + SM::q_bytes);                       // [2][K|V][64][LD]
  float* s_all = reinterpret_cast<float*>(pf_smem + SM::q_bytes + 4 * SM::kv_bytes);
  T* p_all = reinterpret_cast<T*>(pf_smem + SM::q_bytes + 4 * SM::kv_bytes + SM::s_bytes);

  pdl_trigger();
  pdl_wait();
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
    if constexpr (LSE) {
      if (c_half == 0) lse[((long long)seq0 + row_g) * nh + head] = l_run > 0.f ? m_run * scale + logf(l_run) : -INFINITY;
    }
  }
}

// ================================================================================================
// prefill on tcgen05 (head_dim 128): S = Q K^T and O += P V on the 5th-gen tensor cores, accumulators in TMEM
// ================================================================================================
// One CTA = 128 query rows of one (sequence, head).  Q, K, V tiles arrive by TMA (128-byte swizzle); the MMA warp
// issues tcgen05.mma (M=128, N=64, K=16) for S[128 x 64] = Q K^T into one of two TMEM buffers; the four softmax warps
// (thread = query row, lane-aligned with TMEM) read S with tcgen05.ld and run a SINGLE-PASS online softmax:
//   * the row maximum used for the exponent is refreshed lazily -- only when the tile's maximum exceeds it by more than 8
//     (in log2 units; P then stays below 2^8, far inside bf16 / fp16 / fp32 range) -- so the O accumulator in TMEM needs a
//     rescale only in the first tiles of a row block; when any row of a warp needs one, the warp multiplies its 32 TMEM
//     lanes of O in place (tcgen05.ld -> scale -> tcgen05.st) between P.V of the previous tile and P.V of this one;
//   * P = exp2(s * scale*log2e - m) as bf16/fp16 into a swizzled K-major shared tile, the row sum kept in fp32, and the
//     MMA warp issues O[128 x 128] += P V into TMEM columns 128..255 (V is the MN-major B operand).
// Round 1-2 ran two passes (row maxima first, so that O never needed a rescale): 1.5 x the tensor work and twice the
// TMEM reads of S per tile; 254 TFLOP/s at 32 x 576 (profiles/r2_prefill_attention_vs_installed.json).  Causal masking
// only touches the tiles on the diagonal (query tiles are 128, key tiles 64, both aligned).  Epilogue: O / l from TMEM to global.
constexpr int kTcQ = 128, kTcKV = 64, kTcThreads = 192;
constexpr int kTcQHalf = kTcQ * 128;            // one [128 q rows x 128 B] swizzled half of the Q tile (16 KiB)
constexpr int kTcKHalf = kTcKV * 128;           // one [64 kv rows x 128 B] half of a K or V tile (8 KiB)
constexpr int kTcStages = 3;                    // K/V ring: tiles of 2 x 8 KiB, consumed in the order they are loaded
constexpr int kTcSmem = 2 * kTcQHalf + kTcQ * 128 + kTcStages * 2 * kTcKHalf;   // Q 32 + P 16 + ring 48 = 96 KiB -> 2 CTAs / SM

// MN-major (N contiguous) B operand, 128B swizzle: atoms of [8 K-rows x 64 N-elements]; LBO = stride between the
// 64-element N atoms (the two d-halves of the V tile), SBO = stride between consecutive 8-row K groups.
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Pipeline (per CTA; two CTAs share an SM): K and V tiles travel through ONE ring of kTcStages slots in exactly the order the MMA
// warp consumes them (K_0, K_1, V_0, K_2, V_1, ..., V_{n-1}), so the TMA producer runs up to three tiles ahead; S is double-buffered
// in TMEM (2 x 64 columns) and the MMA warp issues QK^T of tile i BEFORE P.V of tile i-1, so the softmax warps work on tile i while
// the tensor core runs P.V(i-1); P is single-buffered behind a p_free barrier, which is also what orders an O rescale (done by the
// softmax warps after p_free = "P.V(i-1) has completed" and before their p_full arrival = "P.V(i) may be issued").
// 2^x on the MUFU (ex2.approx.ftz: relative error 2^-22, flushes denormal results -- P values that small do not matter); exp2f()
// without fast-math wraps the same instruction in a denormal range fix-up of three more
__device__ __forceinline__ float tc_ex2(float x) {
#ifndef CTS_HOST_SHIM
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
#else
  return exp2f(x);
#endif
}

// Persistent CTAs (two per SM) over the (query tile, head, sequence) work items, heaviest query tiles first: at the benchmark shape an
// item is 2..9 key tiles, and with one CTA per item (round 2's first single-pass version) barrier set-up, the 256-column TMEM
// allocation, the first Q / K round trips and the O epilogue were ~45 % of a CTA's life.  Here the pipeline state simply carries on
// from item to item: the producer requests the next item's Q as soon as the last QK^T of the current one has completed (q_free) and its
// K tiles as ring slots free up; the MMA warp starts the next item's QK^T while the softmax warps still store the current O, and only
// the first P.V of an item waits for those warps to have read O (o_free).
struct TcItem {
  int seq0, len, q0, head, kvh, nt;
};
__device__ __forceinline__ bool tc_item(int idx, int nqt, int nh, int nkv, int batch, const int* __restrict__ cu_seqlens, TcItem& it) {
  const int per_qt = nh * batch;
  const int qt = nqt - 1 - idx / per_qt;              // long rows of work first
  const int rem = idx - (idx / per_qt) * per_qt;
  const int b = rem / nh;
  it.head = rem - b * nh;
  it.seq0 = cu_seqlens[b];
  it.len = cu_seqlens[b + 1] - it.seq0;
  it.q0 = qt * kTcQ;
  it.kvh = it.head / (nh / nkv);
  const int kv_end = min(it.len, it.q0 + kTcQ);       // causal: keys 0 .. kv_end-1
  it.nt = it.q0 < it.len ? (kv_end + kTcKV - 1) / kTcKV : 0;
  return it.nt > 0;
}

// does any lane of the warp want it?  One vote instruction (the shuffle tree the first version used was five dependent shuffles on the
// softmax warps' critical path: 8 % of the kernel's stall samples); tests/cuda_on_cpu has shuffles only
__device__ __forceinline__ bool tc_warp_any(bool p) {
#ifndef CTS_HOST_SHIM
  return __any_sync(0xffffffffu, p) != 0;
#else
  int v = p ? 1 : 0;
  for (int o = 16; o > 0; o >>= 1) v |= __shfl_xor_sync(0xffffffffu, v, o);
  return v != 0;
#endif
}

// Measured and rejected (profiles/r2_attention_two_warps_per_row_block_rejected.json, git history): two softmax warps per 32-row block,
// each taking half of a tile's key columns and exchanging the half-row maxima over a 64-thread named barrier -- half the instructions per
// warp and tile, and 5 % SLOWER (301 vs 286 us at 32 x 576, 891 vs 853 us at 4 x 4096): the softmax instruction stream is not the
// critical path any more; what remains is the chain of barrier round trips (S ready -> P ready -> P.V done) per 64-key tile.
template <typename T, bool LSE>
__global__ void __launch_bounds__(kTcThreads, 2)
attn_prefill_tc5_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_k,
                        const __grid_constant__ CUtensorMap tm_v, const int* __restrict__ cu_seqlens, int nh, int nkv, int batch, int nqt,
                        float scale, T* __restrict__ out, float* __restrict__ lse) {
  constexpr int HD = 128;
  extern __shared__ uint8_t tc_raw[];
  __shared__ uint64_t q_bar, q_free, kv_full[kTcStages], kv_empty[kTcStages], s_full[2], s_free[2], p_full, p_free, o_done, o_free;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(tc_raw);
  uint8_t* smem = tc_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* q_s = smem;                               // 32 KiB: 2 d-halves x [128 x 128 B]
  uint8_t* p_s = q_s + 2 * kTcQHalf;                 // 16 KiB: [128 x 128 B] (64 kv columns)
  uint8_t* ring = p_s + kTcQ * 128;                  // kTcStages x 16 KiB: a K or V tile = 2 d-halves x [64 x 128 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_items = nqt * nh * batch;
  pdl_trigger();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(&q_bar, 1);
    mbar_init(&q_free, 1);
    for (int s = 0; s < kTcStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 4); }
    mbar_init(&p_full, 4);
    mbar_init(&p_free, 1);
    mbar_init(&o_done, 1);
    mbar_init(&o_free, 4);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;               // S0: cols 0..63, S1: 64..127, O: 128..255
  pdl_wait();

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      uint32_t n = 0;                                   // K / V tiles requested so far (ring slot n % kTcStages), over all items
      uint32_t k = 0;                                   // live items started by this CTA
      for (int idx = blockIdx.x; idx < total_items; idx += gridDim.x) {
        TcItem it;
        if (!tc_item(idx, nqt, nh, nkv, batch, cu_seqlens, it)) continue;
        if (k > 0) mbar_wait(&q_free, (k - 1) & 1u);     // every QK^T of the previous item has read the Q tile
        mbar_expect_tx(&q_bar, (uint32_t)(2 * kTcQHalf));
        tma_load_2d(q_s, &tm_q, &q_bar, it.head * HD, it.seq0 + it.q0, CTS_L2_EVICT_FIRST);
        tma_load_2d(q_s + kTcQHalf, &tm_q, &q_bar, it.head * HD + 64, it.seq0 + it.q0, CTS_L2_EVICT_FIRST);
        auto load_tile = [&](const CUtensorMap* tm, int j) {
          const uint32_t sl = n % kTcStages;
          mbar_wait(&kv_empty[sl], ((n / kTcStages) & 1u) ^ 1u);
          mbar_expect_tx(&kv_full[sl], (uint32_t)(2 * kTcKHalf));
          uint8_t* dst = ring + (size_t)sl * 2 * kTcKHalf;
          const int row = it.seq0 + j * kTcKV;
          tma_load_2d(dst, tm, &kv_full[sl], it.kvh * HD, row, CTS_L2_EVICT_LAST);
          tma_load_2d(dst + kTcKHalf, tm, &kv_full[sl], it.kvh * HD + 64, row, CTS_L2_EVICT_LAST);
          ++n;
        };
        for (int j = 0; j < it.nt; ++j) {                // K_0, K_1, V_0, K_2, V_1, ..., V_{nt-1}
          load_tile(&tm_k, j);
          if (j > 0) load_tile(&tm_v, j - 1);
        }
        load_tile(&tm_v, it.nt - 1);
        ++k;
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0) {
      constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
      const uint32_t idesc_qk = umma_idesc_f16(kBf16 ? 1 : 0, kTcKV, 128);
      const uint32_t idesc_pv = umma_idesc_f16(kBf16 ? 1 : 0, 128, 128) | (1u << 16);   // B (= V) is MN-major
      const uint32_t q_addr = smem_u32(q_s), p_addr = smem_u32(p_s), ring_addr = smem_u32(ring);
      uint32_t c = 0;                                     // ring tiles consumed so far (same sequence as the producer's)
      uint32_t t = 0;                                     // S tiles issued so far, over all items (S buffer t & 1)
      uint32_t pv = 0;                                    // P.V products issued so far, over all items
      uint32_t k = 0;                                     // live items started
      auto take = [&]() {
        const uint32_t sl = c % kTcStages;
        mbar_wait(&kv_full[sl], (c / kTcStages) & 1u);
        ++c;
        return sl;
      };
      auto issue_pv = [&](bool first_of_item) {           // O (+)= P V for the oldest P not yet multiplied
        mbar_wait(&p_full, pv & 1u);
        const uint32_t sl = take();
        const uint32_t v_addr = ring_addr + sl * 2 * kTcKHalf;
        if (first_of_item && k > 0) mbar_wait(&o_free, (k - 1) & 1u);      // the epilogue warps have read the previous item's O
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < kTcKV / 16; ++kk) {
          const uint64_t bdesc = umma_desc_mn_sw128(v_addr + (uint32_t)kk * 16 * 128, 2 * kTcKHalf / 2, 1024);
          umma_f16(tmem_base + 128, umma_desc_k_sw128(p_addr + (uint32_t)kk * 32), bdesc, idesc_pv, (!first_of_item || kk > 0) ? 1u : 0u);
        }
        umma_commit(&kv_empty[sl]);
        umma_commit(&p_free);
        ++pv;
      };
      for (int idx = blockIdx.x; idx < total_items; idx += gridDim.x) {
        TcItem it;
        if (!tc_item(idx, nqt, nh, nkv, batch, cu_seqlens, it)) continue;
        mbar_wait(&q_bar, k & 1u);
        for (int i = 0; i < it.nt; ++i, ++t) {
          const uint32_t sb = t & 1u;
          const uint32_t sl = take();
          const uint32_t k_addr = ring_addr + sl * 2 * kTcKHalf;
          if (t >= 2) mbar_wait(&s_free[sb], ((t - 2) >> 1) & 1u);      // softmax warps finished reading S[sb]
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint32_t qoff = (uint32_t)(kk >> 2) * kTcQHalf + (uint32_t)(kk & 3) * 32;
            const uint32_t koff = (uint32_t)(kk >> 2) * kTcKHalf + (uint32_t)(kk & 3) * 32;
            umma_f16(tmem_base + sb * kTcKV, umma_desc_k_sw128(q_addr + qoff), umma_desc_k_sw128(k_addr + koff), idesc_qk,
                     kk > 0 ? 1u : 0u);
          }
          umma_commit(&s_full[sb]);
          umma_commit(&kv_empty[sl]);
          if (i + 1 == it.nt) umma_commit(&q_free);        // the Q tile may be replaced once these MMAs have completed
          if (i > 0) issue_pv(i == 1);                     // P.V of the previous tile, behind this tile's QK^T
        }
        issue_pv(it.nt == 1);
        umma_commit(&o_done);
        ++k;
      }
    }
  } else {
    // ------------------------------ softmax / epilogue warps (thread = query row) ------------------------------
    const int q = warp & 3;
    const int r = q * 32 + lane;                          // row inside the tile == TMEM lane
    const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
    const float sl2 = scale * 1.4426950408889634f;
    uint32_t t = 0;                                       // S tiles consumed so far, over all items
    uint32_t k = 0;                                       // live items started
    for (int idx = blockIdx.x; idx < total_items; idx += gridDim.x) {
      TcItem it;
      if (!tc_item(idx, nqt, nh, nkv, batch, cu_seqlens, it)) continue;
      const int q0 = it.q0, qi = q0 + r;                  // query index inside the sequence
      float m_run = -INFINITY, l_run = 0.f;               // m_run: the (lazily refreshed) maximum the exponents are taken against, in log2 units
      for (int j = 0; j < it.nt; ++j, ++t) {
        const int kv0 = j * kTcKV;
        const bool need_mask = kv0 + kTcKV - 1 > q0;      // CTA-uniform: the tile reaches past the first query row
        const uint32_t sb = t & 1u;
        const uint32_t s_addr = lane_base + sb * kTcKV;
        mbar_wait(&s_full[sb], (t >> 1) & 1u);
        tc_fence_after();
        uint32_t sv[kTcKV];
#pragma unroll
        for (int c = 0; c < kTcKV; c += 16) tmem_ld_32x32b_x16(s_addr + (uint32_t)c, sv + c);     // all four loads in flight, ONE wait
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[sb]);          // S[sb] is in registers: QK^T of tile t+2 may overwrite it
        // The softmax warps are instruction-issue bound (ncu of the first single-pass version: ~1 500 instructions per warp and tile,
        // 45 % issue-active, tensor pipe 15 %), so the per-element work is what the arithmetic needs and no more: FMNMX on the raw score,
        // one FFMA + one MUFU.EX2 + one FADD, half a conversion; the causal select only in the (CTA-uniform) diagonal tiles.
        float mx = -INFINITY;
        if (need_mask) {
#pragma unroll
          for (int e = 0; e < kTcKV; ++e) mx = fmaxf(mx, kv0 + e <= qi ? __uint_as_float(sv[e]) : -INFINITY);
        } else {
#pragma unroll
          for (int e = 0; e < kTcKV; ++e) mx = fmaxf(mx, __uint_as_float(sv[e]));
        }
        mx *= sl2;                                        // sl2 > 0: scaling commutes with the maximum
        // lazy maximum: keep the old one unless the tile exceeds it by more than 8 (then P <= 2^8 everywhere)
        float alpha = 1.f;
        bool rescale = false;
        if (mx > m_run + 8.f) {
          rescale = m_run > -INFINITY;                    // the first finite maximum needs no rescale: l and O are still zero
          alpha = rescale ? tc_ex2(m_run - mx) : 1.f;
          m_run = mx;
        }
        const float mneg = m_run > -INFINITY ? -m_run : 0.f;
        l_run *= alpha;
        uint4 pk[kTcKV / 8];
        float ls0 = 0.f, ls1 = 0.f;                       // two partial row sums: half the length of the dependent FADD chain
        if (need_mask) {
#pragma unroll
          for (int c = 0; c < kTcKV; c += 8) {
            float pv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) pv[e] = kv0 + c + e <= qi ? tc_ex2(fmaf(__uint_as_float(sv[c + e]), sl2, mneg)) : 0.f;
            ls0 += (pv[0] + pv[1]) + (pv[2] + pv[3]);
            ls1 += (pv[4] + pv[5]) + (pv[6] + pv[7]);
            pk[c / 8] = pack8<T>(pv);
          }
        } else {
#pragma unroll
          for (int c = 0; c < kTcKV; c += 8) {
            float pv[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) pv[e] = tc_ex2(fmaf(__uint_as_float(sv[c + e]), sl2, mneg));
            ls0 += (pv[0] + pv[1]) + (pv[2] + pv[3]);
            ls1 += (pv[4] + pv[5]) + (pv[6] + pv[7]);
            pk[c / 8] = pack8<T>(pv);
          }
        }
        l_run += ls0 + ls1;
        if (t > 0) mbar_wait(&p_free, (t - 1) & 1u);      // the previous P.V (of this item or the one before) has completed: P may be overwritten, O rescaled
        if (tc_warp_any(rescale)) {                       // warp-collective: rows that keep their maximum multiply by 1
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < HD; c += 16) {
            uint32_t v[16];
            tmem_ld_32x32b_x16(lane_base + 128u + (uint32_t)c, v);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
            tmem_st_32x32b_x16(lane_base + 128u + (uint32_t)c, v);
          }
          tmem_st_wait();
          tc_fence_before();
        }
        // P[r][0..63] -> K-major swizzled tile (row r, 16-byte chunk ch at ch ^ (r & 7))
#pragma unroll
        for (int ch = 0; ch < kTcKV / 8; ++ch)
          *reinterpret_cast<uint4*>(p_s + (uint32_t)r * 128 + ((ch ^ (r & 7)) << 4)) = pk[ch];
        fence_proxy_async_smem();                         // P must be visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full);
      }
      // ---- epilogue: O / l   (tcgen05.ld is warp-collective: every lane loads, rows inside the sequence are stored)
      mbar_wait(&o_done, k & 1u);
      tc_fence_after();
      {
        const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
        const bool row_ok = qi < it.len;
        T* o_g = out + ((long long)it.seq0 + (row_ok ? qi : 0)) * nh * HD + (long long)it.head * HD;
#pragma unroll 1
        for (int c = 0; c < HD; c += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x16(lane_base + 128u + (uint32_t)c, v);
          tmem_ld_32x32b_x16(lane_base + 128u + (uint32_t)c + 16u, v + 16);
          tmem_ld_wait();
          if (c + 32 == HD) {                             // O is in registers: the next item's first P.V may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free);
          }
          if (row_ok) {
            float f[32];
#pragma unroll
            for (int e = 0; e < 32; ++e) f[e] = __uint_as_float(v[e]) * inv;
#pragma unroll
            for (int e = 0; e < 32; e += 8) *reinterpret_cast<uint4*>(o_g + c + e) = pack8<T>(f + e);
          }
        }
        if constexpr (LSE) {
          if (row_ok) lse[((long long)it.seq0 + qi) * nh + it.head] = l_run > 0.f ? m_run * 0.6931471805599453f + logf(l_run) : -INFINITY;   // m_run is in log2 units
        }
      }
      ++k;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ================================================================================================
// decode: tensor-core flash-decoding over TMA-staged, 128B-swizzled KV pages
// ================================================================================================
// grid (num_splits, nkv, batch); 5 warps: warps 0..3 consume (each owns 16 of the 64 tokens of a tile and runs its
// own online softmax, merged once at the end), warp 4 is the TMA producer.  A tile = 64 tokens = 64/page_size pages;
// every page of the cache is [page_size rows x head_dim] contiguous, fetched as head_dim/64 boxes of
// [page_size x 128 B] with the 128-byte swizzle so that ldmatrix reads are bank-conflict free.
//   S[16 x 16] = Q[16(heads, G valid) x d] K^T   mma.sync m16n8k16, K via ldmatrix
//   O[16 x d] += P[16 x 16] V                    V via ldmatrix.trans
// The tensor cores are used to cut instruction count, not for FLOPs: the kernel is bound by the KV stream
// (GQA intensity nh/nkv flop per byte).  The last CTA of a (sequence, kv head) to finish merges the split partials
// (self-resetting counter), so there is no separate combine launch.
constexpr int kDecConsumers = 128;
constexpr int kDecThreads = 160;
constexpr int kDecStages = 3;
constexpr int kDecTile = 64;
constexpr int kDecMaxPagesSm = 256;   // page ids a CTA keeps in shared memory (tiles x pages per tile; beyond that: global reads)
constexpr int kMaxGroup = 8;    // q heads per kv head (rows 8..15 of the MMA tile are padding)

__host__ __device__ inline long long dec_ws_stride(int hd) { return hd + 2; }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t* r) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T> __device__ __forceinline__ void mma16816(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                               uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<__nv_bfloat16>(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                                     uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<__half>(float* c, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                              uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  T v[2] = {DT<T>::from_f(lo), DT<T>::from_f(hi)};
  return *reinterpret_cast<uint32_t*>(v);
}

template <typename T, int HD>
__global__ void __launch_bounds__(kDecThreads)
attn_decode_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v, const T* __restrict__ q,
                   const int* __restrict__ page_table, int max_pages, const int* __restrict__ seq_lens, int nh, int nkv,
                   int page_size, float scale, int num_splits, float* __restrict__ ws, int* __restrict__ counters,
                   T* __restrict__ out) {
  constexpr int NHALF = HD / 64;                 // 128-byte column blocks per row
  constexpr int KSTEPS = HD / 16;
  constexpr int HALF_BYTES = kDecTile * 128;     // one [64 rows x 128 B] swizzled block
  constexpr int TILE_BYTES = NHALF * HALF_BYTES; // K (or V) tile
  extern __shared__ uint8_t dec_raw[];
  __shared__ uint64_t full_bar[kDecStages];
  __shared__ uint64_t empty_bar[kDecStages];
  __shared__ int is_last_s;
  const uint32_t raw = smem_u32(dec_raw);
  uint8_t* smem = dec_raw + (((raw + 1023u) & ~1023u) - raw);      // [stage][K tile | V tile]

  const int split = blockIdx.x, kvh = blockIdx.y, b = blockIdx.z;
  const int G = nh / nkv;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  pdl_trigger();
  if (tid == 0) {
    tma_prefetch_desc(&tm_k);
    tma_prefetch_desc(&tm_v);
    for (int s = 0; s < kDecStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kDecConsumers / 32);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) CTS_TRACE(CTS_TK_ATTN_DECODE, 0);

  // seq_lens / page_table are STEP STATE: written by the advance kernel at the end of the previous step (or by the host before
  // the step) and by nobody inside a step, whose first kernel (embed_gather) waits for everything before it BEFORE it lets its
  // successors start -- so they may be read ahead of the dependency wait.  So may every KV row of an earlier position: the
  // predecessor (RoPE / KV write of THIS layer) writes only row seq_len - 1.
  const int seq_len = seq_lens[b];
  const int n_tiles = (seq_len + kDecTile - 1) / kDecTile;
  const int tps = (n_tiles + num_splits - 1) / num_splits;
  const int tl0 = split * tps, tl1 = min(n_tiles, tl0 + tps);
  const int ntl = max(tl1 - tl0, 0);
  const int npp = kDecTile / page_size;          // pages per tile
  const int t_new = (seq_len - 1) / kDecTile;    // the tile that holds the row written by the predecessor
  // page ids of this CTA's tiles, fetched once by the producer warp (one L2 round trip instead of one per tile inside the ring loop)
  __shared__ int pg_s[kDecMaxPagesSm];
  int n_pre = 0;                                 // tiles whose loads were issued before the dependency wait
  if (warp == kDecConsumers / 32) {
    const int n_pg = min(ntl * npp, kDecMaxPagesSm);
    for (int i = lane; i < n_pg; i += 32) {
      int pg = tl0 * npp + i;
      pg = pg < max_pages ? pg : max_pages - 1;
      pg_s[i] = page_table[(long long)b * max_pages + pg];
    }
    __syncwarp();
    if (lane == 0) {
      for (int it = 0; it < ntl && it < kDecStages && tl0 + it < t_new; ++it) {       // stage it is free: first use of the ring
        mbar_expect_tx(&full_bar[it], (uint32_t)(2 * TILE_BYTES));
        uint8_t* ks = smem + (size_t)it * 2 * TILE_BYTES;
        uint8_t* vs = ks + TILE_BYTES;
        for (int j = 0; j < npp; ++j) {
          const int row0 = (pg_s[it * npp + j] * nkv + kvh) * page_size;
#pragma unroll
          for (int h = 0; h < NHALF; ++h) {
            tma_load_2d(ks + h * HALF_BYTES + j * page_size * 128, &tm_k, &full_bar[it], h * 64, row0, CTS_L2_EVICT_FIRST);
            tma_load_2d(vs + h * HALF_BYTES + j * page_size * 128, &tm_v, &full_bar[it], h * 64, row0, CTS_L2_EVICT_FIRST);
          }
        }
        n_pre = it + 1;
      }
    }
  }
  pdl_wait();          // q and the freshly written KV row come from the predecessor; the split workspace from the previous launch
  if (threadIdx.x == 0) CTS_TRACE(CTS_TK_ATTN_DECODE, 1);

  // accumulators of this warp (rows g = lane/4 are real heads when g < G)
  const int g = lane >> 2, tq = lane & 3;
  float oacc[HD / 8][4];
#pragma unroll
  for (int i = 0; i < HD / 8; ++i) oacc[i][0] = oacc[i][1] = oacc[i][2] = oacc[i][3] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  if (warp == kDecConsumers / 32) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      for (int it = n_pre; it < ntl; ++it) {
        const int s = it % kDecStages;
        const uint32_t ph = (uint32_t)(it / kDecStages) & 1u;
        mbar_wait(&empty_bar[s], ph ^ 1u);
        mbar_expect_tx(&full_bar[s], (uint32_t)(2 * TILE_BYTES));
        uint8_t* ks = smem + (size_t)s * 2 * TILE_BYTES;
        uint8_t* vs = ks + TILE_BYTES;
        for (int j = 0; j < npp; ++j) {
          int page;
          if (it * npp + j < kDecMaxPagesSm) {
            page = pg_s[it * npp + j];
          } else {
            int pg = (tl0 + it) * npp + j;
            pg = pg < max_pages ? pg : max_pages - 1;        // pages past the sequence are masked; keep the read in bounds
            page = page_table[(long long)b * max_pages + pg];
          }
          const int row0 = (page * nkv + kvh) * page_size;
#pragma unroll
          for (int h = 0; h < NHALF; ++h) {
            tma_load_2d(ks + h * HALF_BYTES + j * page_size * 128, &tm_k, &full_bar[s], h * 64, row0, CTS_L2_EVICT_FIRST);
            tma_load_2d(vs + h * HALF_BYTES + j * page_size * 128, &tm_v, &full_bar[s], h * 64, row0, CTS_L2_EVICT_FIRST);
          }
        }
      }
    }
  } else {
    // ------------------------------ consumers ------------------------------
    // Q as the A operand: row g = head, the padding rows (g+8) are zero
    uint32_t qa[KSTEPS][2];
    {
      const T* qrow = q + ((long long)b * nh + (long long)kvh * G + (g < G ? g : 0)) * HD;
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        qa[kk][0] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + kk * 16 + tq * 2) : 0u;
        qa[kk][1] = g < G ? *reinterpret_cast<const uint32_t*>(qrow + kk * 16 + 8 + tq * 2) : 0u;
      }
    }
    const float sl2 = scale * 1.4426950408889634f;
    const int lm = lane >> 3, lr = lane & 7;       // ldmatrix: this lane supplies row lr of matrix lm
    for (int it = 0; it < ntl; ++it) {
      const int s = it % kDecStages;
      mbar_wait(&full_bar[s], (uint32_t)(it / kDecStages) & 1u);
      const uint32_t kbase = smem_u32(smem + (size_t)s * 2 * TILE_BYTES);
      const uint32_t vbase = kbase + TILE_BYTES;
      // ---- S = Q K^T for this warp's 16 tokens (2 n-tiles of 8)
      float sacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        const int row = warp * 16 + (lm >> 1) * 8 + lr;          // matrices: (nt0,k0-7) (nt0,k8-15) (nt1,k0-7) (nt1,k8-15)
        const int chunk = (kk & 3) * 2 + (lm & 1);
        uint32_t bf[4];
        ldsm_x4(kbase + (kk >> 2) * HALF_BYTES + row * 128 + ((chunk ^ (row & 7)) << 4), bf);
        mma16816<T>(sacc[0], qa[kk][0], 0u, qa[kk][1], 0u, bf[0], bf[1]);
        mma16816<T>(sacc[1], qa[kk][0], 0u, qa[kk][1], 0u, bf[2], bf[3]);
      }
      // ---- online softmax on row g (values of a row live in the 4 lanes of a quad)
      const int tok0 = (tl0 + it) * kDecTile + warp * 16 + tq * 2;
      float sv[2][2];
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const bool valid = tok0 + nt * 8 + j < seq_len;
          sv[nt][j] = valid ? sacc[nt][j] * sl2 : -INFINITY;
          mx = fmaxf(mx, sv[nt][j]);
        }
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      const float m_new = fmaxf(m_run, mx);
      float alpha = 1.f, p[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
      if (m_new > -INFINITY) {
        alpha = m_run > -INFINITY ? exp2f(m_run - m_new) : 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
          for (int j = 0; j < 2; ++j) p[nt][j] = exp2f(sv[nt][j] - m_new);      // exp2(-inf) = 0 for masked tokens
      }
      float ps = p[0][0] + p[0][1] + p[1][0] + p[1][1];
      ps += __shfl_xor_sync(0xffffffffu, ps, 1);
      ps += __shfl_xor_sync(0xffffffffu, ps, 2);
      l_run = l_run * alpha + ps;
      m_run = m_new;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) { oacc[i][0] *= alpha; oacc[i][1] *= alpha; }
      // ---- O += P V ; P (row g x 16 tokens) is the A operand straight from the S accumulator layout
      const uint32_t pa0 = pack2<T>(p[0][0], p[0][1]), pa2 = pack2<T>(p[1][0], p[1][1]);
#pragma unroll
      for (int dp = 0; dp < HD / 16; ++dp) {
        const int row = warp * 16 + (lm & 1) * 8 + lr;           // matrices: (tok0-7,dn) (tok8-15,dn) (tok0-7,dn+1) (tok8-15,dn+1)
        const int chunk = 2 * dp + (lm >> 1);
        uint32_t bf[4];
        ldsm_x4_trans(vbase + (chunk >> 3) * HALF_BYTES + row * 128 + (((chunk & 7) ^ (row & 7)) << 4), bf);
        mma16816<T>(oacc[2 * dp], pa0, 0u, pa2, 0u, bf[0], bf[1]);
        mma16816<T>(oacc[2 * dp + 1], pa0, 0u, pa2, 0u, bf[2], bf[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
  }
  __syncthreads();     // every tile consumed: the ring can be reused as merge scratch

  // ---- merge the 4 consumer warps (disjoint token subsets) -> this split's (m, l, o)
  float* m_sm = reinterpret_cast<float*>(smem);                 // [4][8]
  float* l_sm = m_sm + 32;                                      // [4][8]
  float* o_sm = l_sm + 32;                                      // [4][8][HD]
  if (warp < kDecConsumers / 32 && g < G) {
    if (tq == 0) { m_sm[warp * 8 + g] = m_run; l_sm[warp * 8 + g] = l_run; }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      o_sm[(warp * 8 + g) * HD + i * 8 + tq * 2] = oacc[i][0];
      o_sm[(warp * 8 + g) * HD + i * 8 + tq * 2 + 1] = oacc[i][1];
    }
  }
  __syncthreads();
  const long long wstride = dec_ws_stride(HD);
  const long long head0 = (long long)b * nh + (long long)kvh * G;
  for (int i = tid; i < G * HD; i += kDecThreads) {
    const int gg = i / HD, d = i % HD;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < 4; ++w) M = fmaxf(M, m_sm[w * 8 + gg]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
      const float mw = m_sm[w * 8 + gg];
      const float f = mw > -INFINITY ? exp2f(mw - M) : 0.f;
      L += l_sm[w * 8 + gg] * f;
      O += o_sm[(w * 8 + gg) * HD + d] * f;
    }
    if (num_splits == 1) {
      out[(head0 + gg) * HD + d] = DT<T>::from_f(L > 0.f ? O / L : 0.f);
    } else {
      float* w = ws + ((head0 + gg) * num_splits + split) * wstride;
      w[2 + d] = O;
      if (d == 0) { w[0] = M; w[1] = L; }
    }
  }
  if (num_splits == 1) return;
  // ---- the last split of this (sequence, kv head) to arrive merges all partials
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const int prev = atomicAdd(&counters[b * nkv + kvh], 1);
    is_last_s = prev == num_splits - 1;
  }
  __syncthreads();
  if (!is_last_s) return;
  __threadfence();
  for (int i = tid; i < G * HD; i += kDecThreads) {
    const int gg = i / HD, d = i % HD;
    const float* w = ws + (head0 + gg) * num_splits * wstride;
    // all loads of a batch of splits are issued before the first use: a run-time-bounded loop of dependent uses costs one L2 round
    // trip per split (16 splits at b = 1: ~20 us of pure latency, measured with csrc/trace.cuh).  Same arithmetic, same order.
    constexpr int kMB = 8;
    float M = -INFINITY;
    for (int s0 = 0; s0 < num_splits; s0 += kMB) {
      float mv[kMB];
#pragma unroll
      for (int u = 0; u < kMB; ++u) mv[u] = s0 + u < num_splits ? __ldcg(w + (s0 + u) * wstride) : -INFINITY;
#pragma unroll
      for (int u = 0; u < kMB; ++u) M = fmaxf(M, mv[u]);
    }
    float L = 0.f, O = 0.f;
    for (int s0 = 0; s0 < num_splits; s0 += kMB) {
      float mv[kMB], lv[kMB], ov[kMB];
#pragma unroll
      for (int u = 0; u < kMB; ++u) {
        const bool ok = s0 + u < num_splits;
        mv[u] = ok ? __ldcg(w + (s0 + u) * wstride) : -INFINITY;
        lv[u] = ok ? __ldcg(w + (s0 + u) * wstride + 1) : 0.f;
        ov[u] = ok ? __ldcg(w + (s0 + u) * wstride + 2 + d) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < kMB; ++u)
        if (s0 + u < num_splits) {
          const float f = mv[u] > -INFINITY ? exp2f(mv[u] - M) : 0.f;
          L += lv[u] * f;
          O += ov[u] * f;
        }
    }
    out[(head0 + gg) * HD + d] = DT<T>::from_f(L > 0.f ? O / L : 0.f);
  }
  if (tid == 0) counters[b * nkv + kvh] = 0;       // self-resetting for the next launch
}

}  // namespace

static int attn_prefill_impl(cts_ctx* ctx, const void* q, const void* k, const void* v, const int* cu_seqlens, int batch,
                             int max_seqlen, long long total_tokens_hint, int nh, int nkv, int head_dim, float scale, void* out,
                             float* lse, int dtype, void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, q && k && v && cu_seqlens && out, "null pointer");
  CTS_CHECK_ARG(ctx, nh > 0 && nkv > 0 && nh % nkv == 0, "nh must be a positive multiple of nkv");
  CTS_CHECK_ARG(ctx, dtype == CTS_BF16 || dtype == CTS_F16, "dtype");
  if (head_dim != 64 && head_dim != 128) return cts_set_error(ctx, CTS_ERR_UNSUPPORTED, "cts_attn_prefill: head_dim %d (64 or 128)", head_dim);
  if (batch == 0 || max_seqlen == 0) return CTS_OK;
  CTS_CHECK_ARG(ctx, batch <= 65535 && nh <= 65535, "grid");
  cudaStream_t st = (cudaStream_t)stream;
  if (head_dim == 128 && !ctx->force_wmma_attention) {
    // tcgen05 path: Q/K/V through TMA tensor maps over the packed [tokens, heads*128] activations
    const long long total_tokens = total_tokens_hint > 0 ? total_tokens_hint : (long long)batch * max_seqlen;
    CUtensorMap tm_q, tm_k, tm_v;
    int rc = cts_make_tmap_2d(ctx, &tm_q, q, total_tokens, (long long)nh * 128, (long long)nh * 128, kTcQ, dtype == CTS_BF16);
    if (rc) return rc;
    rc = cts_make_tmap_2d(ctx, &tm_k, k, total_tokens, (long long)nkv * 128, (long long)nkv * 128, kTcKV, dtype == CTS_BF16);
    if (rc) return rc;
    rc = cts_make_tmap_2d(ctx, &tm_v, v, total_tokens, (long long)nkv * 128, (long long)nkv * 128, kTcKV, dtype == CTS_BF16);
    if (rc) return rc;
    const int nqt = (max_seqlen + kTcQ - 1) / kTcQ;
    long long items = (long long)nqt * nh * batch;
    CTS_CHECK_ARG(ctx, items < 2147483647LL, "too many (query tile, head, sequence) items");
    long long ctas = 2LL * ctx->sm_count;                 // persistent: two CTAs per SM (shared memory and TMEM columns allow exactly two)
    if (ctas > items) ctas = items;
    dim3 g5((unsigned)ctas, 1, 1);
    const size_t smem5 = (size_t)kTcSmem + 1024;
#define TC5_LAUNCH(TT, LSEV)                                                                                 \
  {                                                                                                          \
    auto kern = attn_prefill_tc5_kernel<TT, LSEV>;                                                           \
    CTS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem5));      \
    CTS_CUDA(ctx, launch_pdl(kern, g5, dim3(kTcThreads), smem5, st, 1, tm_q, tm_k, tm_v, cu_seqlens, nh, nkv, batch, nqt, scale, (TT*)out, lse)); \
  }
    if (dtype == CTS_BF16) {
      if (lse) TC5_LAUNCH(__nv_bfloat16, true) else TC5_LAUNCH(__nv_bfloat16, false)
    } else {
      if (lse) TC5_LAUNCH(__half, true) else TC5_LAUNCH(__half, false)
    }
#undef TC5_LAUNCH
    return CTS_OK;
  }
  dim3 grid((unsigned)((max_seqlen + kPfQ - 1) / kPfQ), (unsigned)nh, (unsigned)batch);
#define PF_LAUNCH2(TT, HDV, LSEV)                                                                            \
  {                                                                                                          \
    auto kern = attn_prefill_kernel<TT, HDV, LSEV>;                                                          \
    const size_t smem = PfSmem<HDV>::total;                                                                  \
    CTS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    CTS_CUDA(ctx, launch_pdl(kern, grid, dim3(kPfThreads), smem, st, 1, (const TT*)q, (const TT*)k, (const TT*)v, cu_seqlens, nh, nkv, \
                             scale, (TT*)out, lse));                                                         \
  }
#define PF_LAUNCH(TT, HDV) { if (lse) PF_LAUNCH2(TT, HDV, true) else PF_LAUNCH2(TT, HDV, false) }
  if (dtype == CTS_BF16) {
    if (head_dim == 128) PF_LAUNCH(__nv_bfloat16, 128) else PF_LAUNCH(__nv_bfloat16, 64)
  } else {
    if (head_dim == 128) PF_LAUNCH(__half, 128) else PF_LAUNCH(__half, 64)
  }
#undef PF_LAUNCH
#undef PF_LAUNCH2
  CTS_LAUNCH_CHECK(ctx);
  return CTS_OK;
}

extern "C" int cts_attn_prefill(cts_ctx* ctx, const void* q, const void* k, const void* v, const int* cu_seqlens, int batch,
                                int max_seqlen, long long total_tokens_hint, int nh, int nkv, int head_dim, float scale, void* out,
                                int dtype, void* stream) {
  return attn_prefill_impl(ctx, q, k, v, cu_seqlens, batch, max_seqlen, total_tokens_hint, nh, nkv, head_dim, scale, out, nullptr,
                           dtype, stream);
}

// training forward: same kernels + the softmax statistics the backward needs (csrc/attention_bwd.cu)
extern "C" int cts_attn_prefill_lse(cts_ctx* ctx, const void* q, const void* k, const void* v, const int* cu_seqlens, int batch,
                                    int max_seqlen, long long total_tokens_hint, int nh, int nkv, int head_dim, float scale,
                                    void* out, float* lse, int dtype, void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, lse != nullptr, "null lse");
  return attn_prefill_impl(ctx, q, k, v, cu_seqlens, batch, max_seqlen, total_tokens_hint, nh, nkv, head_dim, scale, out, lse, dtype,
                           stream);
}

extern "C" long long cts_attn_decode_workspace_floats(int batch, int nh, int head_dim, int num_splits) {
  // fp32 partials [batch, nh, splits, head_dim + 2] followed by int32 arrival counters [batch * nh] (upper bound on
  // batch * nkv).  The caller zero-fills the buffer ONCE; the counters reset themselves after every launch.
  return (long long)batch * nh * num_splits * dec_ws_stride(head_dim) + (long long)batch * nh;
}

extern "C" int cts_attn_decode(cts_ctx* ctx, const void* q, const void* k_cache, const void* v_cache, int num_pages,
                               const int* page_table, int max_pages, const int* seq_lens, int batch, int nh, int nkv,
                               int head_dim, int page_size, float scale, int num_splits, float* workspace, void* out, int dtype,
                               void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, q && k_cache && v_cache && page_table && seq_lens && workspace && out, "null pointer");
  CTS_CHECK_ARG(ctx, nh > 0 && nkv > 0 && nh % nkv == 0 && nh / nkv <= kMaxGroup, "nh/nkv must be an integer <= 8");
  CTS_CHECK_ARG(ctx, page_size == 16 || page_size == 32 || page_size == 64, "page_size must be 16, 32 or 64");
  CTS_CHECK_ARG(ctx, num_splits >= 1 && num_splits <= 65535, "num_splits");
  CTS_CHECK_ARG(ctx, num_pages > 0 && max_pages > 0, "num_pages / max_pages");
  CTS_CHECK_ARG(ctx, (long long)num_pages * nkv * page_size < 2147483647LL, "KV cache too large for 32-bit TMA row coordinates");
  CTS_CHECK_ARG(ctx, dtype == CTS_BF16 || dtype == CTS_F16, "dtype");
  if (head_dim != 64 && head_dim != 128) return cts_set_error(ctx, CTS_ERR_UNSUPPORTED, "cts_attn_decode: head_dim %d (64 or 128)", head_dim);
  if (batch == 0) return CTS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const long long rows = (long long)num_pages * nkv * page_size;
  CUtensorMap tm_k, tm_v;
  int rc = cts_make_tmap_2d(ctx, &tm_k, k_cache, rows, head_dim, head_dim, page_size, dtype == CTS_BF16);
  if (rc) return rc;
  rc = cts_make_tmap_2d(ctx, &tm_v, v_cache, rows, head_dim, head_dim, page_size, dtype == CTS_BF16);
  if (rc) return rc;
  int* counters = reinterpret_cast<int*>(workspace + (long long)batch * nh * num_splits * dec_ws_stride(head_dim));
  dim3 grid((unsigned)num_splits, (unsigned)nkv, (unsigned)batch);
  const size_t smem = (size_t)kDecStages * 2 * (head_dim / 64) * kDecTile * 128 + 1024;
#define DEC_LAUNCH(TT, HDV)                                                                                    \
  {                                                                                                            \
    auto kern = attn_decode_kernel<TT, HDV>;                                                                   \
    CTS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    CTS_CUDA(ctx, launch_pdl(kern, grid, dim3(kDecThreads), smem, st, 1, tm_k, tm_v, (const TT*)q, page_table, max_pages, seq_lens, \
                             nh, nkv, page_size, scale, num_splits, workspace, counters, (TT*)out));           \
  }
  if (dtype == CTS_BF16) {
    if (head_dim == 128) DEC_LAUNCH(__nv_bfloat16, 128) else DEC_LAUNCH(__nv_bfloat16, 64)
  } else {
    if (head_dim == 128) DEC_LAUNCH(__half, 128) else DEC_LAUNCH(__half, 64)
  }
#undef DEC_LAUNCH
  return CTS_OK;
}

CTS_TRACE_SETTER(cts_trace_set_attention)
