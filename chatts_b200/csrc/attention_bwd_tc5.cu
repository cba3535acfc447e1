// chatts_b200 -- attention backward on tcgen05 (head_dim 128): the same FlashAttention-2 recomputation as attention_bwd.cu
// (dV = P^T dO, dP = dO V^T, dS = P o (dP - delta) * scale, dQ = dS K, dK = dS^T Q) with every product on the 5th-gen tensor
// cores and every accumulator in TMEM, built from the operand layouts of the GPU-validated forward kernel
// (attention.cu: attn_prefill_tc5_kernel) -- K-major 128B-swizzled Q/K/V/dO tiles by TMA, P-like tiles written by the
// softmax warps as swizzled K-major A operands, V-like tiles consumed as MN-major B operands.
//
//   attn_bwd_dq_tc5_kernel   CTA = 128 query rows x one head; per 64-key tile:  S = Q K^T and dP = dO V^T (TMEM cols 0..63 /
//                            64..127) -> softmax warps (thread = query row = TMEM lane) form dS -> dQ += dS K (TMEM cols 128..255)
//   attn_bwd_dkv_tc5_kernel  CTA = 128 key rows x one kv head; per (q head of the group, 64-query tile at or below the
//                            diagonal):  S^T = K Q^T, dP^T = V dO^T -> softmax warps (thread = key row) form P^T and dS^T ->
//                            dV += P^T dO (cols 128..255), dK += dS^T Q (cols 256..383): summed over the GQA group in TMEM,
//                            no atomics, fixed order.
// Round-1 pipeline: the streamed operand tiles (K/V in the dQ kernel, Q/dO in the dK/dV kernel) are double-buffered, so the TMA
// of tile j+1 is in flight while tile j is consumed; S / dP and the P-like tiles are single-buffered (the tensor core idles
// while the softmax warps work -- the next tuning step).  Every mbarrier wait is bounded (common.cuh: a protocol bug traps, it
// cannot hang the box).  OPT-IN (CTS_ATTN_BWD_TC5=1): written after the round-1 GPU budget was spent, not yet executed on a
// B200; the HMMA kernels of attention_bwd.cu stay the default until this one has passed tests/test_gpu_zz_d_attn_bwd_tc5.py.
#include <type_traits>

#include "common.cuh"
#include "tensormap.cuh"

namespace {

constexpr int kHD = 128;
constexpr int kBigRows = 128, kSmallRows = 64, kThreads = 192;
constexpr int kBigHalf = kBigRows * 128;        // one [128 rows x 128 B] swizzled d-half (16 KiB)
constexpr int kSmallHalf = kSmallRows * 128;    // one [64 rows x 128 B] d-half (8 KiB)
constexpr int kPTile = kBigRows * 128;          // [128 rows x 64 bf16] P-like operand tile (16 KiB)
constexpr float kLog2e = 1.4426950408889634f;

// MN-major (N contiguous) B operand, 128B swizzle (see attention.cu): LBO = stride between the two 64-element N atoms (the
// d-halves of a [rows x 128] tile), SBO = stride between consecutive 8-row K groups.
__device__ __forceinline__ uint64_t desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// D[128 x 64] (=|+=) A[128 x 128] B[64 x 128]^T : both K-major tiles of two d-halves (A rows 128, B rows 64)
__device__ __forceinline__ void mma_rows_x_rowsT(uint32_t d_tmem, uint32_t a_addr, uint32_t b_addr, uint32_t idesc) {
#pragma unroll
  for (int kk = 0; kk < kHD / 16; ++kk) {
    const uint32_t aoff = (uint32_t)(kk >> 2) * kBigHalf + (uint32_t)(kk & 3) * 32;
    const uint32_t boff = (uint32_t)(kk >> 2) * kSmallHalf + (uint32_t)(kk & 3) * 32;
    umma_f16(d_tmem, umma_desc_k_sw128(a_addr + aoff), umma_desc_k_sw128(b_addr + boff), idesc, kk > 0 ? 1u : 0u);
  }
}
// D[128 x 128] (+)= A[128 x 64] B[64 x 128] : A = P-like K-major tile, B = a [64 rows x 128] tile read MN-major
__device__ __forceinline__ void mma_p_x_tile(uint32_t d_tmem, uint32_t p_addr, uint32_t b_addr, uint32_t idesc, bool accumulate) {
#pragma unroll
  for (int kk = 0; kk < kSmallRows / 16; ++kk) {
    const uint64_t bdesc = desc_mn_sw128(b_addr + (uint32_t)kk * 16 * 128, kSmallHalf, 1024);
    umma_f16(d_tmem, umma_desc_k_sw128(p_addr + (uint32_t)kk * 32), bdesc, idesc, (accumulate || kk > 0) ? 1u : 0u);
  }
}
// this thread's 64 values -> row r of a swizzled K-major [128 x 64] tile (16-byte chunk ch lands at ch ^ (r & 7))
__device__ __forceinline__ void store_p_row(uint8_t* tile, int r, const uint4* pk) {
#pragma unroll
  for (int ch = 0; ch < kSmallRows / 8; ++ch) *reinterpret_cast<uint4*>(tile + (uint32_t)r * 128 + ((ch ^ (r & 7)) << 4)) = pk[ch];
}

// ================================================================================================ dQ
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dq_tc5_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                       const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                       const float* __restrict__ lse, const float* __restrict__ delta, const int* __restrict__ cu_seqlens, int nh,
                       int nkv, float scale, T* __restrict__ dq) {
  extern __shared__ uint8_t dq_raw[];
  __shared__ uint64_t qdo_bar, kv_full[2], kv_empty[2], s_full, s_free, ds_full, ds_free, o_done;
  __shared__ uint32_t tmem_slot;
  const uint32_t raw = smem_u32(dq_raw);
  uint8_t* smem = dq_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* q_s = smem;                       // 32 KiB
  uint8_t* do_s = q_s + 2 * kBigHalf;        // 32 KiB
  uint8_t* ds_s = do_s + 2 * kBigHalf;       // 16 KiB
  uint8_t* kv_s = ds_s + kPTile;             // 2 stages x (K 16 KiB | V 16 KiB): tile j+1 lands while tile j is being consumed
  constexpr int kKvStage = 4 * kSmallHalf;

  const int b = blockIdx.z, head = blockIdx.y, qt = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(&qdo_bar, 1);
    for (int st = 0; st < 2; ++st) { mbar_init(&kv_full[st], 1); mbar_init(&kv_empty[st], 1); }
    mbar_init(&s_full, 1); mbar_init(&s_free, 4);
    mbar_init(&ds_full, 4); mbar_init(&ds_free, 1);
    mbar_init(&o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<256>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;      // S: cols 0..63, dP: 64..127, dQ: 128..255
  pdl_wait();

  const int seq0 = cu_seqlens[b], len = cu_seqlens[b + 1] - seq0;
  const int q0 = qt * kBigRows;
  const bool live = q0 < len;                // CTA-uniform
  const int kvh = head / (nh / nkv);
  const int kv_end = min(len, q0 + kBigRows);
  const int nt = live ? (kv_end + kSmallRows - 1) / kSmallRows : 0;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0 && live) {
      mbar_expect_tx(&qdo_bar, (uint32_t)(4 * kBigHalf));
      tma_load_2d(q_s, &tm_q, &qdo_bar, head * kHD, seq0 + q0, CTS_L2_EVICT_FIRST);
      tma_load_2d(q_s + kBigHalf, &tm_q, &qdo_bar, head * kHD + 64, seq0 + q0, CTS_L2_EVICT_FIRST);
      tma_load_2d(do_s, &tm_do, &qdo_bar, head * kHD, seq0 + q0, CTS_L2_EVICT_FIRST);
      tma_load_2d(do_s + kBigHalf, &tm_do, &qdo_bar, head * kHD + 64, seq0 + q0, CTS_L2_EVICT_FIRST);
      for (int j = 0; j < nt; ++j) {
        const int row = seq0 + j * kSmallRows;
        const int st = j & 1;
        uint8_t* k_s = kv_s + st * kKvStage;
        uint8_t* v_s = k_s + 2 * kSmallHalf;
        mbar_wait(&kv_empty[st], (((uint32_t)j >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&kv_full[st], (uint32_t)kKvStage);
        tma_load_2d(k_s, &tm_k, &kv_full[st], kvh * kHD, row, CTS_L2_EVICT_LAST);
        tma_load_2d(k_s + kSmallHalf, &tm_k, &kv_full[st], kvh * kHD + 64, row, CTS_L2_EVICT_LAST);
        tma_load_2d(v_s, &tm_v, &kv_full[st], kvh * kHD, row, CTS_L2_EVICT_LAST);
        tma_load_2d(v_s + kSmallHalf, &tm_v, &kv_full[st], kvh * kHD + 64, row, CTS_L2_EVICT_LAST);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0 && live) {
      constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
      const uint32_t idesc_s = umma_idesc_f16(kBf16 ? 1 : 0, kSmallRows, 128);
      const uint32_t idesc_o = umma_idesc_f16(kBf16 ? 1 : 0, 128, 128) | (1u << 16);      // B is MN-major
      const uint32_t q_addr = smem_u32(q_s), do_addr = smem_u32(do_s), ds_addr = smem_u32(ds_s);
      mbar_wait(&qdo_bar, 0);
      for (int j = 0; j < nt; ++j) {
        const int st = j & 1;
        const uint32_t k_addr = smem_u32(kv_s + st * kKvStage), v_addr = k_addr + 2 * kSmallHalf;
        mbar_wait(&kv_full[st], ((uint32_t)j >> 1) & 1u);
        if (j > 0) mbar_wait(&s_free, (uint32_t)(j - 1) & 1u);      // softmax warps finished reading S / dP of tile j-1
        tc_fence_after();
        mma_rows_x_rowsT(tmem_base, q_addr, k_addr, idesc_s);        // S  = Q  K^T
        mma_rows_x_rowsT(tmem_base + 64, do_addr, v_addr, idesc_s);  // dP = dO V^T
        umma_commit(&s_full);
        mbar_wait(&ds_full, (uint32_t)j & 1u);                       // dS(j) is in shared memory
        tc_fence_after();
        mma_p_x_tile(tmem_base + 128, ds_addr, k_addr, idesc_o, j > 0);   // dQ += dS K
        umma_commit(&kv_empty[st]);                                  // this K/V stage may be reloaded
        umma_commit(&ds_free);                                       // dS tile may be rewritten
      }
      umma_commit(&o_done);
    }
  } else if (live) {
    // ------------------------------ softmax / epilogue warps (thread = query row = TMEM lane) ------------------------------
    const int qw = warp & 3;
    const int r = qw * 32 + lane;
    const int qi = q0 + r;
    const bool row_ok = qi < len;
    const uint32_t lane_base = tmem_base + ((uint32_t)(qw * 32) << 16);
    const float sl2 = scale * kLog2e;
    const long long sidx = ((long long)seq0 + (row_ok ? qi : 0)) * nh + head;
    const float lse2 = row_ok ? lse[sidx] * kLog2e : 0.f;
    const float dlt = row_ok ? delta[sidx] : 0.f;
    for (int j = 0; j < nt; ++j) {
      const int kv0 = j * kSmallRows;
      mbar_wait(&s_full, (uint32_t)j & 1u);
      tc_fence_after();
      uint4 pk[kSmallRows / 8];
#pragma unroll
      for (int c = 0; c < kSmallRows; c += 16) {
        uint32_t sv[16], dv[16];
        tmem_ld_32x32b_x16(lane_base + (uint32_t)c, sv);
        tmem_ld_32x32b_x16(lane_base + 64u + (uint32_t)c, dv);
        tmem_ld_wait();
        float ds[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const bool ok = row_ok && kv0 + c + e <= qi;
          const float p = ok ? exp2f(__uint_as_float(sv[e]) * sl2 - lse2) : 0.f;
          ds[e] = ok ? p * (__uint_as_float(dv[e]) - dlt) * scale : 0.f;
        }
        pk[c / 8] = pack8<T>(ds);
        pk[c / 8 + 1] = pack8<T>(ds + 8);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free);
      if (j > 0) mbar_wait(&ds_free, (uint32_t)(j - 1) & 1u);        // dQ MMA of tile j-1 has consumed the dS tile
      store_p_row(ds_s, r, pk);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full);
    }
    mbar_wait(&o_done, 0);
    tc_fence_after();
    T* o_g = dq + ((long long)seq0 + (row_ok ? qi : 0)) * nh * kHD + (long long)head * kHD;
#pragma unroll 1
    for (int c = 0; c < kHD; c += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(lane_base + 128u + (uint32_t)c, v);         // warp-collective: every lane loads
      tmem_ld_wait();
      if (row_ok) {
        float f[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]);
        *reinterpret_cast<uint4*>(o_g + c) = pack8<T>(f);
        *reinterpret_cast<uint4*>(o_g + c + 8) = pack8<T>(f + 8);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ================================================================================================ dK, dV
template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dkv_tc5_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_do,
                        const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_v,
                        const float* __restrict__ lse, const float* __restrict__ delta, const int* __restrict__ cu_seqlens, int nh,
                        int nkv, float scale, T* __restrict__ dk, T* __restrict__ dv) {
  extern __shared__ uint8_t dkv_raw[];
  __shared__ uint64_t kv_bar, qd_full[2], qd_empty[2], s_full, s_free, pd_full, pd_free, done_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ float stat_s[2][2][kSmallRows];          // [iteration parity][lse * log2e | delta][query of the tile]
  const uint32_t raw = smem_u32(dkv_raw);
  uint8_t* smem = dkv_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* k_s = smem;                        // 32 KiB
  uint8_t* v_s = k_s + 2 * kBigHalf;          // 32 KiB
  uint8_t* qd_s = v_s + 2 * kBigHalf;         // 2 stages x (Q 16 KiB | dO 16 KiB)
  constexpr int kQdStage = 4 * kSmallHalf;
  uint8_t* pt_s = qd_s + 2 * kQdStage;        // 16 KiB  P^T
  uint8_t* dst_s = pt_s + kPTile;             // 16 KiB  dS^T

  const int b = blockIdx.z, kvh = blockIdx.y, kt = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tm_q); tma_prefetch_desc(&tm_do); tma_prefetch_desc(&tm_k); tma_prefetch_desc(&tm_v);
    mbar_init(&kv_bar, 1);
    for (int st = 0; st < 2; ++st) { mbar_init(&qd_full[st], 1); mbar_init(&qd_empty[st], 1); }
    mbar_init(&s_full, 1); mbar_init(&s_free, 4);
    mbar_init(&pd_full, 4); mbar_init(&pd_free, 1);
    mbar_init(&done_bar, 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(&tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;       // S^T: cols 0..63, dP^T: 64..127, dV: 128..255, dK: 256..383
  pdl_wait();

  const int seq0 = cu_seqlens[b], len = cu_seqlens[b + 1] - seq0;
  const int kv0 = kt * kBigRows;
  const bool live = kv0 < len;                // CTA-uniform
  const int G = nh / nkv;
  const int it0 = kv0 / kSmallRows;           // first 64-query tile at or below the diagonal
  const int nq = (len + kSmallRows - 1) / kSmallRows;
  const int nI = live ? nq - it0 : 0;
  const int n_iter = G * nI;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0 && live) {
      mbar_expect_tx(&kv_bar, (uint32_t)(4 * kBigHalf));
      tma_load_2d(k_s, &tm_k, &kv_bar, kvh * kHD, seq0 + kv0, CTS_L2_EVICT_FIRST);
      tma_load_2d(k_s + kBigHalf, &tm_k, &kv_bar, kvh * kHD + 64, seq0 + kv0, CTS_L2_EVICT_FIRST);
      tma_load_2d(v_s, &tm_v, &kv_bar, kvh * kHD, seq0 + kv0, CTS_L2_EVICT_FIRST);
      tma_load_2d(v_s + kBigHalf, &tm_v, &kv_bar, kvh * kHD + 64, seq0 + kv0, CTS_L2_EVICT_FIRST);
      for (int n = 0; n < n_iter; ++n) {
        const int head = kvh * G + n / nI;
        const int row = seq0 + (it0 + n % nI) * kSmallRows;
        const int st = n & 1;
        uint8_t* q_s = qd_s + st * kQdStage;
        uint8_t* do_s = q_s + 2 * kSmallHalf;
        mbar_wait(&qd_empty[st], (((uint32_t)n >> 1) & 1u) ^ 1u);
        mbar_expect_tx(&qd_full[st], (uint32_t)kQdStage);
        tma_load_2d(q_s, &tm_q, &qd_full[st], head * kHD, row, CTS_L2_EVICT_LAST);
        tma_load_2d(q_s + kSmallHalf, &tm_q, &qd_full[st], head * kHD + 64, row, CTS_L2_EVICT_LAST);
        tma_load_2d(do_s, &tm_do, &qd_full[st], head * kHD, row, CTS_L2_EVICT_LAST);
        tma_load_2d(do_s + kSmallHalf, &tm_do, &qd_full[st], head * kHD + 64, row, CTS_L2_EVICT_LAST);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ------------------------------
    if (lane == 0 && live) {
      constexpr bool kBf16 = std::is_same<T, __nv_bfloat16>::value;
      const uint32_t idesc_s = umma_idesc_f16(kBf16 ? 1 : 0, kSmallRows, 128);
      const uint32_t idesc_o = umma_idesc_f16(kBf16 ? 1 : 0, 128, 128) | (1u << 16);      // B is MN-major
      const uint32_t k_addr = smem_u32(k_s), v_addr = smem_u32(v_s), pt_addr = smem_u32(pt_s), dst_addr = smem_u32(dst_s);
      mbar_wait(&kv_bar, 0);
      for (int n = 0; n < n_iter; ++n) {
        const int st = n & 1;
        const uint32_t q_addr = smem_u32(qd_s + st * kQdStage), do_addr = q_addr + 2 * kSmallHalf;
        mbar_wait(&qd_full[st], ((uint32_t)n >> 1) & 1u);
        if (n > 0) mbar_wait(&s_free, (uint32_t)(n - 1) & 1u);
        tc_fence_after();
        mma_rows_x_rowsT(tmem_base, k_addr, q_addr, idesc_s);         // S^T  = K Q^T
        mma_rows_x_rowsT(tmem_base + 64, v_addr, do_addr, idesc_s);   // dP^T = V dO^T
        umma_commit(&s_full);
        mbar_wait(&pd_full, (uint32_t)n & 1u);                        // P^T and dS^T are in shared memory
        tc_fence_after();
        mma_p_x_tile(tmem_base + 128, pt_addr, do_addr, idesc_o, n > 0);    // dV += P^T  dO
        mma_p_x_tile(tmem_base + 256, dst_addr, q_addr, idesc_o, n > 0);    // dK += dS^T Q
        umma_commit(&qd_empty[st]);                                   // this Q/dO stage may be reloaded
        umma_commit(&pd_free);                                        // P^T, dS^T may be rewritten
      }
      umma_commit(&done_bar);
    }
  } else if (live) {
    // ------------------------------ softmax / epilogue warps (thread = key row = TMEM lane) ------------------------------
    const int kw = warp & 3;
    const int r = kw * 32 + lane;
    const int st = (int)threadIdx.x - 64;       // 0..127 inside the softmax group
    const int kv_g = kv0 + r;
    const bool kv_ok = kv_g < len;
    const uint32_t lane_base = tmem_base + ((uint32_t)(kw * 32) << 16);
    const float sl2 = scale * kLog2e;
    for (int n = 0; n < n_iter; ++n) {
      const int head = kvh * G + n / nI;
      const int q0 = (it0 + n % nI) * kSmallRows;
      float* lse_s = stat_s[n & 1][0];
      float* dlt_s = stat_s[n & 1][1];
      if (st < kSmallRows) {                    // stage the 64 queries' statistics (double-buffered by iteration parity)
        const int qi = q0 + st;
        const bool ok = qi < len;
        const long long sidx = ((long long)seq0 + (ok ? qi : 0)) * nh + head;
        lse_s[st] = ok ? lse[sidx] * kLog2e : 0.f;
        dlt_s[st] = ok ? delta[sidx] : 0.f;
      }
      named_bar_sync(1, 128);
      mbar_wait(&s_full, (uint32_t)n & 1u);
      tc_fence_after();
      uint4 ppk[kSmallRows / 8], dpk[kSmallRows / 8];
#pragma unroll
      for (int c = 0; c < kSmallRows; c += 16) {
        uint32_t sv[16], dvv[16];
        tmem_ld_32x32b_x16(lane_base + (uint32_t)c, sv);
        tmem_ld_32x32b_x16(lane_base + 64u + (uint32_t)c, dvv);
        tmem_ld_wait();
        float pp[16], ds[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const int qi = q0 + c + e;
          const bool ok = kv_ok && qi < len && kv_g <= qi;
          const float p = ok ? exp2f(__uint_as_float(sv[e]) * sl2 - lse_s[c + e]) : 0.f;
          pp[e] = p;
          ds[e] = ok ? p * (__uint_as_float(dvv[e]) - dlt_s[c + e]) * scale : 0.f;
        }
        ppk[c / 8] = pack8<T>(pp);
        ppk[c / 8 + 1] = pack8<T>(pp + 8);
        dpk[c / 8] = pack8<T>(ds);
        dpk[c / 8 + 1] = pack8<T>(ds + 8);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free);
      if (n > 0) mbar_wait(&pd_free, (uint32_t)(n - 1) & 1u);        // the MMAs of iteration n-1 have consumed both tiles
      store_p_row(pt_s, r, ppk);
      store_p_row(dst_s, r, dpk);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pd_full);
    }
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    const long long orow = ((long long)seq0 + (kv_ok ? kv_g : 0)) * nkv * kHD + (long long)kvh * kHD;
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
      T* o_g = (which == 0 ? dv : dk) + orow;
      const uint32_t col0 = which == 0 ? 128u : 256u;
#pragma unroll 1
      for (int c = 0; c < kHD; c += 16) {
        uint32_t v[16];
        tmem_ld_32x32b_x16(lane_base + col0 + (uint32_t)c, v);
        tmem_ld_wait();
        if (kv_ok) {
          float f[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(v[e]);
          *reinterpret_cast<uint4*>(o_g + c) = pack8<T>(f);
          *reinterpret_cast<uint4*>(o_g + c + 8) = pack8<T>(f + 8);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

constexpr int kDqSmem = 4 * kBigHalf + kPTile + 8 * kSmallHalf;          // Q 32 + dO 32 + dS 16 + 2 x (K 16 + V 16) = 144 KiB
constexpr int kDkvSmem = 4 * kBigHalf + 8 * kSmallHalf + 2 * kPTile;     // K 32 + V 32 + 2 x (Q 16 + dO 16) + P^T 16 + dS^T 16 = 160 KiB

}  // namespace

// called by cts_attn_bwd (attention_bwd.cu) after the delta kernel, head_dim 128 only
int cts_attn_bwd_tc5_launch(cts_ctx* ctx, const void* q, const void* k, const void* v, const void* dout, const float* lse,
                            const float* delta, const int* cu_seqlens, int batch, int max_seqlen, long long total_tokens, int nh,
                            int nkv, float scale, void* dq, void* dk, void* dv, int dtype, cudaStream_t st) {
  const bool bf = dtype == CTS_BF16;
  const long long qc = (long long)nh * kHD, kc = (long long)nkv * kHD;
  CUtensorMap q128, do128, k64, v64, q64, do64, k128, v128;
  int rc;
  if ((rc = cts_make_tmap_2d(ctx, &q128, q, total_tokens, qc, qc, kBigRows, bf))) return rc;
  if ((rc = cts_make_tmap_2d(ctx, &do128, dout, total_tokens, qc, qc, kBigRows, bf))) return rc;
  if ((rc = cts_make_tmap_2d(ctx, &k64, k, total_tokens, kc, kc, kSmallRows, bf))) return rc;
  if ((rc = cts_make_tmap_2d(ctx, &v64, v, total_tokens, kc, kc, kSmallRows, bf))) return rc;
  if ((rc = cts_make_tmap_2d(ctx, &q64, q, total_tokens, qc, qc, kSmallRows, bf))) return rc;
  if ((rc = cts_make_tmap_2d(ctx, &do64, dout, total_tokens, qc, qc, kSmallRows, bf))) return rc;
  if ((rc = cts_make_tmap_2d(ctx, &k128, k, total_tokens, kc, kc, kBigRows, bf))) return rc;
  if ((rc = cts_make_tmap_2d(ctx, &v128, v, total_tokens, kc, kc, kBigRows, bf))) return rc;
  const unsigned tiles = (unsigned)((max_seqlen + kBigRows - 1) / kBigRows);
  const size_t smem_q = (size_t)kDqSmem + 1024, smem_kv = (size_t)kDkvSmem + 1024;
#define TC5_BW(TT)                                                                                                          \
  {                                                                                                                         \
    auto kq = attn_bwd_dq_tc5_kernel<TT>;                                                                                   \
    CTS_CUDA(ctx, cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));                      \
    CTS_CUDA(ctx, launch_pdl(kq, dim3(tiles, (unsigned)nh, (unsigned)batch), dim3(kThreads), smem_q, st, 1, q128, do128, k64, v64, \
                             lse, delta, cu_seqlens, nh, nkv, scale, (TT*)dq));                                              \
    auto kkv = attn_bwd_dkv_tc5_kernel<TT>;                                                                                 \
    CTS_CUDA(ctx, cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));                    \
    CTS_CUDA(ctx, launch_pdl(kkv, dim3(tiles, (unsigned)nkv, (unsigned)batch), dim3(kThreads), smem_kv, st, 1, q64, do64, k128,   \
                             v128, lse, delta, cu_seqlens, nh, nkv, scale, (TT*)dk, (TT*)dv));                               \
  }
  if (bf) TC5_BW(__nv_bfloat16) else TC5_BW(__half)
#undef TC5_BW
  return CTS_OK;
}
