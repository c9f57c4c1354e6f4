// Shared device/host helpers for libfq_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "fq.h"

namespace fq {

// ---------------------------------------------------------------------------
// host side: errors, tensor views, device properties
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
int sm_count();                      // SMs of the current device (148 on B200)

struct View {
  void* data = nullptr;
  int64_t numel = 0;
  int code = 0, bits = 0;
  bool null = true;
  template <class T> T* as() const { return reinterpret_cast<T*>(data); }
  bool is_f32() const { return code == kDLFloat && bits == 32; }
};
// Validates a borrowed DLTensor (CUDA device, lanes 1, compact).  Returns false + sets the error.
bool view_of(const DLTensor* t, const char* name, bool allow_null, View* out);

#define FQ_TRY(expr)                 \
  do {                               \
    if (!(expr)) return -1;          \
  } while (0)
#define FQ_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      ::fq::set_error(__VA_ARGS__);  \
      return -1;                     \
    }                                \
  } while (0)
#define FQ_CUDA(call)                                                            \
  do {                                                                           \
    cudaError_t e__ = (call);                                                    \
    if (e__ != cudaSuccess) {                                                    \
      ::fq::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
      return -1;                                                                 \
    }                                                                            \
  } while (0)
#define FQ_LAUNCH_CHECK(name)                                                    \
  do {                                                                           \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      ::fq::set_error("launch of %s failed: %s", name, cudaGetErrorString(e__)); \
      return -1;                                                                 \
    }                                                                            \
  } while (0)

// ---------------------------------------------------------------------------
// workspace layout (device memory, zero-filled once; kernels restore the zeros)
// ---------------------------------------------------------------------------
struct Workspace {
  unsigned int ticket;          // "last block done" counter
  unsigned int ticket2;
  unsigned int pad[30];
  unsigned int rowmax[FQ_MAX_ROWS];   // |x| bit patterns, atomicMax target
  float minmax_part[2 * 4096];        // per-block partials of fq_minmax
  double stats_part[2 * FQ_MAX_STAT_BLOCKS];   // per-block {S1, S2} of fq_channel_stats
};

constexpr int kThreads = 256;
constexpr int kUnroll = 4;            // independent 16 B loads in flight per thread

#ifdef __CUDACC__
// Launch `kernel` as a programmatic dependent of the previous kernel in `st` (see pdl_wait below).
template <class... KArgs, class... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
// keeps the line in L2 with normal priority (first pass of a two-pass kernel)
__device__ __forceinline__ float4 ld_keep(const float4* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// Programmatic dependent launch (sm_90+): a kernel launched with launch_dependent() below may start while the
// kernel before it in the stream is still running.  It must call pdl_wait() before touching anything that
// kernel writes (and before writing anything that kernel reads); everything above the wait -- index math, loads
// of data the earlier kernel only reads -- overlaps with the earlier kernel's tail.  The earlier kernel calls
// pdl_launch_dependents() as its first instruction.  Both are no-ops under a plain launch.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// mshadow_op::round == roundf (half away from zero); rintf + tie fix-up, branch free.
__device__ __forceinline__ float round_half_away(float q) {
  float r = rintf(q);                              // half to even
  float diff = __fsub_rn(q, r);                    // exact
  // a tie that went towards zero has diff == +-0.5 with the sign of q
  if (fabsf(diff) == 0.5f && (diff > 0.f) == (q > 0.f)) r = __fadd_rn(r, copysignf(1.0f, q));
  return r;
}

// mshadow_op::clip: x > hi -> hi ; x < lo -> lo ; else x  (NaN passes through)
__device__ __forceinline__ float clipf(float x, float lo, float hi) {
  return x > hi ? hi : (x < lo ? lo : x);
}

// IEEE divide -> roundf: the integer-valued "code", by the book.
__device__ __forceinline__ float quant_code(float x, float d) {
  return round_half_away(__fdiv_rn(x, d));
}

// The same code in ~6 instructions instead of ~20.  With r = RN(1/d) and q0 = RN(v * r),
//   |q0 - RN(v / d)| <= (3u + u^2) |v / d|  <  0.38 * 2^-21 * |q0|        (u = 2^-24, no underflow in r),
// so whenever q0 is farther than 2^-21 |q0| from every rounding boundary k + 0.5, q0 and the IEEE
// quotient lie between the same two boundaries and roundf(v / d) == rintf(q0) (neither is a tie).
// Anything closer (a few elements per million), NaN/Inf, and divisors outside [1e-30, 1e30] take
// the IEEE divide.  Bit-exactness therefore does not rest on the fast path being correctly rounded.
struct QDiv {
  float d, r, guard;       // guard = 2^-21 on the fast path, +inf to force the IEEE divide
  __device__ __forceinline__ static QDiv make(float d) {
    QDiv q;
    q.d = d;
    q.r = __frcp_rn(d);
    q.guard = (d >= 1e-30f && d <= 1e30f) ? 4.76837158203125e-7f : INFINITY;
    return q;
  }
  // near(q0, t): true when q0 is within the guard band of a rounding boundary (or NaN)
  __device__ __forceinline__ bool near_tie(float q0, float t) const {
    const float e = __fsub_rn(q0, t);                                  // |e| = distance to the integer, <= 0.5
    return !(__fmaf_rn(fabsf(q0), guard, fabsf(e)) < 0.5f);            // true for NaN as well
  }
  __device__ __forceinline__ float code(float v) const {
    const float q0 = __fmul_rn(v, r);
    const float t = rintf(q0);
    if (!near_tie(q0, t)) return t;
    return quant_code(v, d);
  }
  // Four at once: straight-line fast path for the whole vector, one branch for the rare fix-up.
  __device__ __forceinline__ float4 code4(float4 v) const {
    const float q0 = __fmul_rn(v.x, r), q1 = __fmul_rn(v.y, r), q2 = __fmul_rn(v.z, r), q3 = __fmul_rn(v.w, r);
    float4 t = make_float4(rintf(q0), rintf(q1), rintf(q2), rintf(q3));
    const bool n0 = near_tie(q0, t.x), n1 = near_tie(q1, t.y), n2 = near_tie(q2, t.z), n3 = near_tie(q3, t.w);
    if (n0 | n1 | n2 | n3) {
      if (n0) t.x = quant_code(v.x, d);
      if (n1) t.y = quant_code(v.y, d);
      if (n2) t.z = quant_code(v.z, d);
      if (n3) t.w = quant_code(v.w, d);
    }
    return t;
  }
};

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// max over the block of a non-negative value; valid in thread 0.  `red` = 32 floats of smem.
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                       // protect `red` from the previous use
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    v = lane < (int)(blockDim.x >> 5) ? red[lane] : 0.f;
    v = warp_max(v);
  }
  return v;
}

// Block-cooperative walk over x[begin, end): scalar head up to 16 B alignment, float4 body with
// kUnroll independent loads in flight, scalar tail.  vf(elem_index, float4), sf(elem_index, float).
// REVERSE walks the body from the top (second pass of a two-pass kernel: most recently read first).
template <bool REVERSE, bool KEEP, class VF, class SF>
__device__ __forceinline__ void for_range(const float* __restrict__ x, int64_t begin, int64_t end, VF vf, SF sf) {
  const int64_t len = end - begin;
  if (len <= 0) return;
  const int tid = threadIdx.x, nt = blockDim.x;
  const float* p = x + begin;
  int64_t head = (int64_t)(((16u - (unsigned)((uintptr_t)p & 15u)) & 15u) >> 2);
  if (head > len) head = len;
  const int64_t nvec = (len - head) >> 2;
  const int64_t tail0 = head + 4 * nvec;
  if (tid < head) sf(begin + tid, p[tid]);
  if (tid < len - tail0) sf(begin + tail0 + tid, p[tail0 + tid]);
  const float4* p4 = reinterpret_cast<const float4*>(p + head);
  const int64_t base = begin + head;
  int64_t i = tid;
  for (; i + (int64_t)(kUnroll - 1) * nt < nvec; i += (int64_t)kUnroll * nt) {
    float4 v[kUnroll];
    int64_t j[kUnroll];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      j[k] = REVERSE ? (nvec - 1 - (i + (int64_t)k * nt)) : (i + (int64_t)k * nt);
      v[k] = KEEP ? ld_keep(p4 + j[k]) : ld_stream(p4 + j[k]);
    }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) vf(base + 4 * j[k], v[k]);
  }
  for (; i < nvec; i += nt) {
    const int64_t j = REVERSE ? (nvec - 1 - i) : i;
    vf(base + 4 * j, KEEP ? ld_keep(p4 + j) : ld_stream(p4 + j));
  }
}

// Tile-interleaved walk over a 16 B aligned x[0, n): tile k of 4 * kThreads * kUnroll elements goes to
// block k % gridDim.x, so at any moment the resident blocks stream one contiguous window of the
// tensor (DRAM row locality for the mixed read/write streams of the elementwise kernels) instead of
// gridDim.x far-apart slices.  REVERSE starts from the last tile (second pass: newest lines first).
constexpr int64_t kTileElems = 4LL * kThreads * kUnroll;

template <bool REVERSE, bool KEEP, class VF, class SF>
__device__ __forceinline__ void for_tiles(const float* __restrict__ x, int64_t n, VF vf, SF sf) {
  const int64_t nvec = n >> 2;
  const int64_t ntiles = (n + kTileElems - 1) / kTileElems;
  const float4* p4 = reinterpret_cast<const float4*>(x);
  for (int64_t k = blockIdx.x; k < ntiles; k += gridDim.x) {
    const int64_t tile = REVERSE ? (ntiles - 1 - k) : k;
    const int64_t v0 = tile * (kTileElems / 4) + threadIdx.x;
    float4 v[kUnroll];
    if (v0 + (int64_t)(kUnroll - 1) * kThreads < nvec) {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) v[u] = KEEP ? ld_keep(p4 + v0 + u * kThreads) : ld_stream(p4 + v0 + u * kThreads);
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) vf(4 * (v0 + u * kThreads), v[u]);
    } else {
#pragma unroll
      for (int u = 0; u < kUnroll; ++u) {
        const int64_t j = v0 + u * kThreads;
        if (j < nvec) vf(4 * j, KEEP ? ld_keep(p4 + j) : ld_stream(p4 + j));
      }
    }
  }
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < n - tail0) sf(tail0 + threadIdx.x, x[tail0 + threadIdx.x]);
}

#endif  // __CUDACC__

}  // namespace fq
