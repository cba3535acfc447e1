// chatts_b200 -- optional device-side timeline (debug / profiling aid, off unless cts_trace_enable() installed a buffer).
//
// ncu serialises kernels and runs them cold, so it cannot show how the ~440 PDL-chained kernels of a decode step OVERLAP inside a
// CUDA-graph replay.  With a buffer installed, instrumented kernels append {tag, %globaltimer} records at a few points (entry,
// dependency wait released, last load issued, exit); tools/trace_decode_step.py turns one replay into a per-kernel timeline (when
// each grid's first CTA started, when its wait was released, when its last CTA left) and lists the HBM-idle gaps between the
// weight streams.  Without a buffer a mark is one constant-bank load and a predicated branch.
#pragma once
#ifdef CTS_HOST_SHIM
#define CTS_TRACE(kind, phase) ((void)0)
#define CTS_TRACE_SETTER(name) extern "C" int name(unsigned long long* buf) { (void)buf; return 0; }
#else
static __constant__ unsigned long long* cts_trace_c = nullptr;      // [0] cursor, [1] capacity (records), records {tag, time} from [2]
__device__ __forceinline__ void cts_trace_mark(unsigned kind, unsigned phase) {
  unsigned long long* b = cts_trace_c;
  if (b == nullptr) return;
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const unsigned long long i = atomicAdd(b, 1ull);
  if (i >= b[1]) return;
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const unsigned long long cta = blockIdx.x + gridDim.x * (blockIdx.y + (unsigned long long)gridDim.y * blockIdx.z);
  b[2 + 2 * i] = ((unsigned long long)(kind & 0xff) << 56) | ((unsigned long long)(phase & 0xf) << 52) | ((unsigned long long)(gridDim.x & 0xfff) << 40) |
                 ((unsigned long long)(gridDim.z & 0xff) << 32) | ((unsigned long long)(smid & 0xff) << 24) | (cta & 0xffffffull);
  b[3 + 2 * i] = t;
}
#define CTS_TRACE(kind, phase) cts_trace_mark((kind), (phase))
#define CTS_TRACE_SETTER(name)                                                                             \
  extern "C" int name(unsigned long long* buf) {                                                           \
    return cudaMemcpyToSymbol(cts_trace_c, &buf, sizeof(buf)) == cudaSuccess ? 0 : -1;                     \
  }
#endif
// kinds
#define CTS_TK_GEMM 1
#define CTS_TK_ATTN_DECODE 2
#define CTS_TK_NORM 3
#define CTS_TK_SWIGLU 4
#define CTS_TK_ROPE 5
#define CTS_TK_FUSED 6
#define CTS_TK_OTHER 7
