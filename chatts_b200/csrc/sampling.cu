// chatts_b200 -- stochastic sampling + device-side advance of the decode loop (K13, sampled variant).
// Replaces, for temperature / top-k / top-p decoding (chatts/utils/inference_tsmllm_deepspeed.py:95-100 calls
// generate(temperature=0.2); chatts/utils/llm_utils.py:166-170 builds vLLM SamplingParams(temperature, top_p)):
// transformers' TemperatureLogitsWarper -> TopKLogitsWarper -> TopPLogitsWarper -> softmax -> multinomial, and the same
// host bookkeeping as cts_greedy_advance, so that a SAMPLED decode step is also one graph-replayable launch sequence
// with no host round trip.
//
// One CTA per sequence, no sort: the kept set {i : z_i >= tau} is found by bisection over the 16-bit ordered key of the
// (bf16 / fp16) logit -- at most 16 counting passes per filter over a row that stays in L2 (304 KB at V = 152k):
//   top-k : the largest tau with  count(z_i >= tau) >= k
//   top-p : the largest tau (>= tau_k) with  sum_{z_i >= tau} p_i >= top_p * mass(top-k set)     (ties at tau are all kept)
// then one inverse-CDF pass in INDEX order over the kept set with a counter-based uniform u(seed, step, sequence)
// (splitmix64), so a (seed, step) pair always gives the same token -- restated on the CPU in tests/cabi_double.py.
#include "common.cuh"

namespace {

constexpr int kSaThreads = 1024;

__host__ __device__ inline unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// monotone map of a 16-bit float pattern to an unsigned key (larger value <=> larger key); NaN patterns sort above +inf
__device__ __forceinline__ unsigned key16(unsigned short b) { return (b & 0x8000u) ? (unsigned)(unsigned short)~b : (unsigned)(b | 0x8000u); }

template <typename T> __device__ __forceinline__ unsigned short bits16(T v) { return *reinterpret_cast<unsigned short*>(&v); }

__device__ __forceinline__ float block_sum_f(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 32 ? red[threadIdx.x] : 0.f;      // 32 warps
  if (threadIdx.x < 32) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}
__device__ __forceinline__ float block_max_f(float v, float* red) {
  v = warp_max(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = threadIdx.x < 32 ? red[threadIdx.x] : -INFINITY;
  if (threadIdx.x < 32) t = warp_max(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

template <typename T>
__global__ void __launch_bounds__(kSaThreads)
sample_advance_kernel(const T* __restrict__ logits, long long vocab, float inv_temp, int top_k, float top_p, unsigned long long seed,
                      int* __restrict__ out_tokens, int out_ld, int* __restrict__ step_ptr, int* __restrict__ cur_ids,
                      int* __restrict__ positions, int* __restrict__ seq_lens, int* __restrict__ slot_map,
                      const int* __restrict__ page_table, int max_pages, int page_size) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  __shared__ float scan_s[32];
  __shared__ int winner_s;
  const int b = blockIdx.x;
  const T* row = logits + (long long)b * vocab;
  // contiguous chunk of indices per thread (multiple of 8 so the vector path stays aligned)
  const long long per = ((vocab + kSaThreads - 1) / kSaThreads + 7) / 8 * 8;
  const long long i0 = (long long)threadIdx.x * per;
  const long long i1 = i0 + per < vocab ? i0 + per : vocab;
  const bool vec = (vocab & 7) == 0 && (reinterpret_cast<uintptr_t>(row) & 15) == 0;

  // visit(fn): fn(index, value as float, 16-bit pattern) over this thread's chunk
  auto visit = [&](auto&& fn) {
    if (vec) {
      for (long long i = i0; i < i1; i += 8) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + i);
        const T* p = reinterpret_cast<const T*>(&u);
#pragma unroll
        for (int j = 0; j < 8; ++j) fn(i + j, DT<T>::to_f(p[j]), bits16<T>(p[j]));
      }
    } else {
      for (long long i = i0; i < i1; ++i) fn(i, DT<T>::to_f(row[i]), bits16<T>(row[i]));
    }
  };

  // ---- max and partition function of z = logit / temperature
  float m = -INFINITY;
  visit([&](long long, float x, unsigned short) { if (x > m) m = x; });      // NaN never wins
  m = block_max_f(m, red);
  auto prob = [&](float x) { return x == x ? __expf((x - m) * inv_temp) : 0.f; };   // unnormalised; NaN logits get mass 0
  unsigned kmax = 0;
  visit([&](long long, float x, unsigned short bb) { if (x == x) { const unsigned k = key16(bb); if (k > kmax) kmax = k; } });
  kmax = (unsigned)block_max_f((float)kmax, red);                            // keys < 2^16 are exact in fp32
  auto mass_ge = [&](unsigned tau) {
    float s = 0.f;
    visit([&](long long, float x, unsigned short bb) { if (x == x && key16(bb) >= tau) s += prob(x); });
    return block_sum_f(s, red);
  };
  auto count_ge = [&](unsigned tau) {
    float c = 0.f;
    visit([&](long long, float x, unsigned short bb) { if (x == x && key16(bb) >= tau) c += 1.f; });
    return block_sum_f(c, red);                                              // counts <= 2^24 are exact in fp32
  };
  unsigned tau = 0;
  if (top_k > 0 && (long long)top_k < vocab) {
    unsigned lo = 0, hi = kmax + 1;                                          // count(>= lo) >= k holds, count(>= hi) = 0 does not
    while (hi - lo > 1) {
      const unsigned mid = lo + ((hi - lo) >> 1);
      if (count_ge(mid) >= (float)top_k) lo = mid; else hi = mid;
    }
    tau = lo;
  }
  float mass = mass_ge(tau);
  if (top_p > 0.f && top_p < 1.f) {
    const float target = top_p * mass;
    unsigned lo = tau, hi = kmax + 1;
    while (hi - lo > 1) {
      const unsigned mid = lo + ((hi - lo) >> 1);
      if (mass_ge(mid) >= target) lo = mid; else hi = mid;
    }
    if (lo != tau) { tau = lo; mass = mass_ge(tau); }
  }
  // ---- inverse CDF in index order over the kept set
  const int step = step_ptr ? step_ptr[0] : 0;
  const unsigned long long h = splitmix64(seed ^ splitmix64(((unsigned long long)(unsigned)step << 32) | (unsigned)b));
  const float u = (float)(h >> 40) * (1.0f / 16777216.0f);                   // 24 random bits -> [0, 1)
  float loc = 0.f;
  visit([&](long long, float x, unsigned short bb) { if (x == x && key16(bb) >= tau) loc += prob(x); });
  // block exclusive scan of the per-thread sums (threads own increasing index ranges)
  float incl = loc;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += n;
  }
  if (threadIdx.x == 0) winner_s = -1;
  __syncthreads();
  if ((threadIdx.x & 31) == 31) scan_s[threadIdx.x >> 5] = incl;
  __syncthreads();
  if (threadIdx.x < 32) {
    float w = scan_s[threadIdx.x], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, wi, o);
      if (threadIdx.x >= o) wi += n;
    }
    scan_s[threadIdx.x] = wi - w;                                            // exclusive prefix of the warp totals
    if (threadIdx.x == 31) red[0] = wi;                                      // grand total (same summation tree every launch)
  }
  __syncthreads();
  const float total = red[0];
  const float excl = scan_s[threadIdx.x >> 5] + incl - loc;
  float target = u * total;
  if (target >= total) target = total * 0.99999994f;
  if (loc > 0.f && target >= excl && target < excl + loc) {
    float acc = excl;
    int pick = -1;
    long long last = -1;
    visit([&](long long i, float x, unsigned short bb) {
      if (pick < 0 && x == x && key16(bb) >= tau) {
        acc += prob(x);
        last = i;
        if (target < acc) pick = (int)i;
      }
    });
    if (pick < 0) pick = (int)last;                                          // rounding at the chunk's upper edge
    winner_s = pick;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int tok = winner_s;
    if (tok < 0) {                                                           // target fell into a rounding gap between chunks
      float best = -INFINITY;                                                // -> most probable token (thread 0 scans the row)
      for (long long i = 0; i < vocab; ++i) {
        const float x = DT<T>::to_f(row[i]);
        if (x > best) { best = x; tok = (int)i; }
      }
      if (tok < 0) tok = 0;
    }
    if (out_tokens) out_tokens[(long long)b * out_ld + step] = tok;
    if (cur_ids) cur_ids[b] = tok;
    if (positions) {
      const int np = positions[b] + 1;
      positions[b] = np;
      if (seq_lens) seq_lens[b] = seq_lens[b] + 1;
      if (slot_map && page_table) {
        const int pg = np / page_size;
        slot_map[b] = pg < max_pages ? page_table[(long long)b * max_pages + pg] * page_size + np % page_size : -1;
      }
    }
    if (step_ptr) {
      __threadfence();
      const int done = atomicAdd(&step_ptr[1], 1) + 1;                       // step_ptr = {step, done-counter}
      if (done == (int)gridDim.x) {
        step_ptr[1] = 0;
        step_ptr[0] = step + 1;
      }
    }
  }
}

}  // namespace

extern "C" int cts_sample_advance(cts_ctx* ctx, const void* logits, long long vocab, int batch, float temperature, int top_k,
                                  float top_p, unsigned long long seed, int* out_tokens, int out_ld, int* step_ptr, int* cur_ids,
                                  int* positions, int* seq_lens, int* slot_map, const int* page_table, int max_pages, int page_size,
                                  int dtype, void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, logits != nullptr && vocab > 0 && vocab <= (1LL << 24) && batch >= 0, "logits / vocab (<= 2^24) / batch");
  CTS_CHECK_ARG(ctx, temperature > 0.f, "temperature must be > 0 (greedy decoding is cts_greedy_advance)");
  CTS_CHECK_ARG(ctx, dtype == CTS_BF16 || dtype == CTS_F16, "dtype");
  CTS_CHECK_ARG(ctx, page_size > 0 || slot_map == nullptr, "page_size");
  if (batch == 0) return CTS_OK;
  const float inv_t = 1.0f / temperature;
  if (dtype == CTS_BF16) {
    CTS_CUDA(ctx, launch_pdl(sample_advance_kernel<__nv_bfloat16>, dim3((unsigned)batch), dim3(kSaThreads), 0, (cudaStream_t)stream, 1,
                             (const __nv_bfloat16*)logits, vocab, inv_t, top_k, top_p, seed, out_tokens, out_ld, step_ptr, cur_ids,
                             positions, seq_lens, slot_map, page_table, max_pages, page_size));
  } else {
    CTS_CUDA(ctx, launch_pdl(sample_advance_kernel<__half>, dim3((unsigned)batch), dim3(kSaThreads), 0, (cudaStream_t)stream, 1,
                             (const __half*)logits, vocab, inv_t, top_k, top_p, seed, out_tokens, out_ld, step_ptr, cur_ids, positions,
                             seq_lens, slot_map, page_table, max_pages, page_size));
  }
  return CTS_OK;
}

// ------------------------------------------------------------------------------------------------
// Repetition penalty (transformers RepetitionPenaltyLogitsProcessor; the checkpoint's generation_config.json may set it): for every
// token id that already occurs in the row's sequence -- prompt and generated part -- logit = logit / penalty if logit > 0 else
// logit * penalty, applied ONCE however often the token occurred.  The occurrence set is a bit mask [batch, ceil(vocab / 32)]:
//   cts_rep_penalty_mark  sets the bits of (row, token) pairs (the prompt once, then the new token of every step)
//   cts_rep_penalty_apply rewrites the logits of the marked tokens in place, before the argmax / sampling kernel reads them.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void rep_mark_kernel(const int* __restrict__ tokens, const int* __restrict__ rows, int n, unsigned* __restrict__ seen,
                                int words_per_row, long long vocab) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tokens[i];
  const int r = rows != nullptr ? rows[i] : i;
  if (t < 0 || t >= vocab || r < 0) return;
  atomicOr(&seen[(long long)r * words_per_row + (t >> 5)], 1u << (t & 31));
}

template <typename T>
__global__ void rep_apply_kernel(T* __restrict__ logits, long long vocab, long long ld, const unsigned* __restrict__ seen,
                                 int words_per_row, float penalty) {
  pdl_trigger();
  pdl_wait();
  const long long b = blockIdx.y;
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= words_per_row) return;
  unsigned bits = seen[b * words_per_row + w];
  while (bits) {
    const int j = __ffs(bits) - 1;
    bits &= bits - 1;
    const long long t = (long long)w * 32 + j;
    if (t < vocab) {
      const float v = DT<T>::to_f(logits[b * ld + t]);
      logits[b * ld + t] = DT<T>::from_f(v > 0.f ? v / penalty : v * penalty);      // the processor works in the logits' dtype
    }
  }
}
}  // namespace

extern "C" int cts_rep_penalty_mark(cts_ctx* ctx, const int* tokens, const int* rows, int n, unsigned* seen, int words_per_row,
                                    long long vocab, void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, n >= 0 && words_per_row > 0 && vocab > 0 && vocab <= 32LL * words_per_row, "n / words_per_row / vocab");
  if (n == 0) return CTS_OK;
  CTS_CHECK_ARG(ctx, tokens != nullptr && seen != nullptr, "null tokens / seen");
  CTS_CUDA(ctx, launch_pdl(rep_mark_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 1, tokens, rows, n, seen,
                           words_per_row, vocab));
  return CTS_OK;
}

extern "C" int cts_rep_penalty_apply(cts_ctx* ctx, void* logits, long long vocab, long long ld, int batch, const unsigned* seen,
                                     int words_per_row, float penalty, int dtype, void* stream) {
  if (!ctx) return CTS_ERR_BAD_ARG;
  CTS_CHECK_ARG(ctx, batch >= 0 && batch <= 65535 && vocab > 0 && ld >= vocab && words_per_row > 0 && vocab <= 32LL * words_per_row, "shape");
  CTS_CHECK_ARG(ctx, penalty > 0.f, "penalty must be > 0");
  CTS_CHECK_ARG(ctx, dtype == CTS_BF16 || dtype == CTS_F16, "dtype");
  if (batch == 0) return CTS_OK;
  CTS_CHECK_ARG(ctx, logits != nullptr && seen != nullptr, "null logits / seen");
  dim3 grid((unsigned)((words_per_row + 255) / 256), (unsigned)batch);
  if (dtype == CTS_BF16)
    CTS_CUDA(ctx, launch_pdl(rep_apply_kernel<__nv_bfloat16>, grid, dim3(256), 0, (cudaStream_t)stream, 1, (__nv_bfloat16*)logits, vocab, ld, seen,
                             words_per_row, penalty));
  else
    CTS_CUDA(ctx, launch_pdl(rep_apply_kernel<__half>, grid, dim3(256), 0, (cudaStream_t)stream, 1, (__half*)logits, vocab, ld, seen, words_per_row,
                             penalty));
  return CTS_OK;
}
