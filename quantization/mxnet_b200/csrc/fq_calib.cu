// K4/K5: on-device EMA, shared-memory-privatised |x| histograms, block-per-candidate KL search.
//   reference: quantize/convert/convert.py:66-78; quantize/distribution_calibrate.py:31-47,117-171;
//              examples/simulate_quantization.py:310
#include "fq_fused.cuh"

namespace fq {

// ---------------------------------------------------------------------------------------------
// EMA
// ---------------------------------------------------------------------------------------------
__global__ void ema_kernel(float* __restrict__ state, const float* __restrict__ cur, int64_t n, double one_minus_m,
                           float m32, int scalar_cur, int promotion) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float t;
    if (scalar_cur && promotion == FQ_PROMOTION_LEGACY)
      t = (float)(one_minus_m * (double)cur[i]);            // python float * numpy.float32 -> float64
    else
      t = __fmul_rn((float)one_minus_m, cur[i]);            // NEP 50 scalar, or _mul_scalar on an NDArray
    state[i] = __fadd_rn(__fmul_rn(m32, state[i]), t);
  }
}

// ---------------------------------------------------------------------------------------------
// histogram of the clipped non-zero values
// ---------------------------------------------------------------------------------------------
// Block-private histogram `sh` (bins + 1 counters + one trash counter per warp), then one global atomic per
// non-empty bin.
//
// Per element: `setp v > 0` (NaN and -0.0 fail: zeros are dropped, :40), `min(v, max_)` (the upper half of
// ndarray.clip(0, max_), :39), `mul`, `cvt.rzi.u32`, a memory-safety clamp to `bins`, an address LEA, a SELECT
// that sends dropped elements to the warp's trash counter, and one unconditional `red.shared.add.u32` (ptxas:
// ATOMS.POPC.INC, lanes with the same address are merged) -- 8 instructions of straight-line code.  The C++ form
// (`if (v != 0) atomicAdd`) compiled to a divergent branch with its BSSY/BSYNC pair and a rematerialised shared
// base per element (14.7 instructions per element, issue slots 70 % busy); a predicated `red` is turned into
// the same branch by ptxas.
// 0 < v <= max_ keeps trunc(v * sc) inside [0, bins]: fl(bins / (max_ + 1e-5)) * max_ <= bins (1 + 2^-23).
struct HistBins {
  float max_, sc;
  unsigned int bins, base, trash;
  __device__ __forceinline__ void init(unsigned int* sh, float max_in, int bins_in, int promotion) {
    max_ = max_in;
    bins = (unsigned int)bins_in;
    // scales = bins / (max_ + 1e-5)      distribution_calibrate.py:41
    sc = (promotion == FQ_PROMOTION_LEGACY) ? (float)((double)bins_in / ((double)max_in + 1e-5))
                                            : __fdiv_rn((float)bins_in, __fadd_rn(max_in, 1e-5f));
    base = (unsigned int)__cvta_generic_to_shared(sh);
    trash = base + (bins + 1u + (threadIdx.x >> 5)) * 4u;
  }
  __device__ __forceinline__ void add(float v) const {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .f32 t;\n\t.reg .u32 q;\n\t"
        "setp.gt.f32 p, %0, 0f00000000;\n\t"
        "min.f32 t, %0, %1;\n\t"
        "mul.rn.f32 t, t, %2;\n\t"
        "cvt.rzi.u32.f32 q, t;\n\t"
        "min.u32 q, q, %3;\n\t"
        "shl.b32 q, q, 2;\n\t"
        "add.u32 q, q, %4;\n\t"
        "selp.u32 q, q, %5, p;\n\t"
        "red.shared.add.u32 [q], 1;\n\t}" ::"f"(v), "f"(max_), "f"(sc), "r"(bins), "r"(base), "r"(trash)
        : "memory");
  }
  __device__ __forceinline__ void add4(float4 v) const {
    add(v.x);
    add(v.y);
    add(v.z);
    add(v.w);
  }
};

// The reference asserts `np.min(fm) >= 0` on EVERY batch (distribution_calibrate.py:35) and its np.max / np.min
// turn a NaN anywhere into a failed assert.  The histogram kernel silently drops negatives and NaN, so when the
// caller passes a flag word it also keeps a NaN-propagating running minimum (min.NaN.f32: one instruction per
// pair of elements) and raises the flag when that minimum is not >= 0 (-0.0 passes, as it does in NumPy).
template <bool CHECK>
struct MinCheck {
  float m = 0.f;
  __device__ __forceinline__ void see(float v) {
    if (CHECK) asm("min.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(v));
  }
  __device__ __forceinline__ void see4(float4 v) {
    if (CHECK) {
      float a, b;
      asm("min.NaN.f32 %0, %1, %2;" : "=f"(a) : "f"(v.x), "f"(v.y));
      asm("min.NaN.f32 %0, %1, %2;" : "=f"(b) : "f"(v.z), "f"(v.w));
      asm("min.NaN.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
      asm("min.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(a));
    }
  }
  __device__ __forceinline__ void publish(int* flag) const {
    if (CHECK && !(m >= 0.f)) atomicOr(flag, 1);
  }
};

__device__ __forceinline__ void hist_zero(unsigned int* sh, int bins) {
  for (int b = threadIdx.x; b < bins + 1 + 32; b += blockDim.x) sh[b] = 0u;
  __syncthreads();
}
// counts: 64-bit, or 32-bit (`wide` == 0) for callers whose per-bin totals stay below 2^32 -- a data-parallel
// calibration then moves half the bytes in its sum-all-reduce (quantization/mxnet_b200/dist.py CountsRing)
struct CountsOut {
  void* p;
  int wide;
  __host__ __device__ __forceinline__ CountsOut at(int64_t off) const {
    CountsOut o;
    o.p = wide ? (void*)((unsigned long long*)p + off) : (void*)((unsigned int*)p + off);
    o.wide = wide;
    return o;
  }
};
__device__ __forceinline__ void hist_flush(const unsigned int* sh, int bins, CountsOut counts) {
  __syncthreads();
  for (int b = threadIdx.x; b <= bins; b += blockDim.x) {
    const unsigned int c = sh[b];
    if (c) {
      if (counts.wide) atomicAdd((unsigned long long*)counts.p + b, (unsigned long long)c);
      else atomicAdd((unsigned int*)counts.p + b, c);
    }
  }
}

// Launch shape (tools/hist_probe.py on the 27 layer inputs of config 2, 2.56 GB): 256 threads, 3 blocks per SM,
// up to 80 registers -- ptxas then keeps 16 independent 16 B loads in flight per thread (the 4-deep tile loop
// unrolled 4x): 360 us = 7.1 TB/s.  More, smaller blocks (8 per SM at 32 registers: 416 us; 6 per SM: 381 us;
// 4 per SM: 402 us) or fewer (2 per SM: 426-440 us; one 512/1024-thread block: 414-443 us) are all slower.
constexpr int kHistThreads = 256;
constexpr int kHistBlocksPerSM = 3;
constexpr int kHistTileVec = kHistThreads * kUnroll;        // float4 per tile

// Block j of the nblk blocks of a 16 B aligned tensor: tiles of kHistTileVec float4 go round-robin over the
// tensor's blocks, so the resident blocks stream one contiguous window of it.
template <bool CHECK>
__device__ __forceinline__ void hist_stream(const float* __restrict__ x, int64_t n, int j, int nblk, const HistBins& hb,
                                            MinCheck<CHECK>& mc) {
  const float4* p4 = reinterpret_cast<const float4*>(x);
  const int64_t nvec = n >> 2;
  int64_t v0 = (int64_t)j * kHistTileVec + threadIdx.x;
  const int64_t step = (int64_t)nblk * kHistTileVec;
  for (; v0 + (int64_t)(kUnroll - 1) * kHistThreads < nvec; v0 += step) {
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = ld_stream(p4 + v0 + u * kHistThreads);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      hb.add4(v[u]);
      mc.see4(v[u]);
    }
  }
  // the tensor's last, partial tile: only the block whose turn it is still has v0 < nvec
#pragma unroll 1
  for (int u = 0; u < kUnroll; ++u) {
    const int64_t vi = v0 + (int64_t)u * kHistThreads;
    if (vi < nvec) {
      const float4 v = ld_stream(p4 + vi);
      hb.add4(v);
      mc.see4(v);
    }
  }
  if (j == nblk - 1 && threadIdx.x < (n & 3)) {
    const float v = x[(nvec << 2) + threadIdx.x];
    hb.add(v);
    mc.see(v);
  }
}

// Multi-tensor launch: block -> (tensor, block within the tensor) through a table passed by value.
struct HistBatch {
  const float* x[FQ_MAX_BATCH];
  int64_t n[FQ_MAX_BATCH];
  int first_block[FQ_MAX_BATCH + 1];
  int count;
};

template <bool CHECK>
__global__ void __launch_bounds__(kHistThreads, kHistBlocksPerSM)
    hist_multi_kernel(const __grid_constant__ HistBatch tb, const float* __restrict__ maxes, int max_stride,
                      int max_offset, int bins, int promotion, CountsOut counts, int* __restrict__ bad_flags) {
  extern __shared__ unsigned int sh[];     // bins + 1 private counters, 32 trash counters
  int t = 0;
  while (t + 1 < tb.count && (int)blockIdx.x >= tb.first_block[t + 1]) ++t;
  const float max_ = __ldg(maxes + (int64_t)t * max_stride + max_offset);
  HistBins hb;
  hb.init(sh, max_, bins, promotion);
  hist_zero(sh, bins);
  // the reference asserts max_ > 0 (:36); with max_ <= 0 everything would clip to <= 0 and be dropped
  MinCheck<CHECK> mc;
  if (max_ > 0.f)
    hist_stream(tb.x[t], tb.n[t], (int)blockIdx.x - tb.first_block[t], tb.first_block[t + 1] - tb.first_block[t], hb, mc);
  else if (CHECK)
    mc.m = -1.f;                  // `assert max_ > 0` (:36); a NaN max lands here as well
  if (CHECK) mc.publish(bad_flags + t);
  hist_flush(sh, bins, counts.at((int64_t)t * (bins + 1)));
}

// Any alignment: grid-stride scalar loads (views at odd offsets; never the hot path).
__global__ void __launch_bounds__(kThreads) hist_unaligned_kernel(const float* __restrict__ x, int64_t n,
                                                                  const float* __restrict__ max_dev, int bins,
                                                                  int promotion, CountsOut counts,
                                                                  int* __restrict__ bad_flag) {
  extern __shared__ unsigned int sh[];
  const float max_ = __ldg(max_dev);
  HistBins hb;
  hb.init(sh, max_, bins, promotion);
  hist_zero(sh, bins);
  MinCheck<true> mc;
  if (max_ > 0.f) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const float v = x[i];
      hb.add(v);
      mc.see(v);
    }
  } else {
    mc.m = -1.f;
  }
  if (bad_flag != nullptr) mc.publish(bad_flag);
  hist_flush(sh, bins, counts);
}

// counts holds `steps` consecutive batches ([steps, nb]); they are folded in batch order, so that a data-parallel
// run may all-reduce the integer counts of many batches at once and still replay the reference's per-batch
// float32 accumulation exactly.
__global__ void hist_accumulate_kernel(CountsOut counts, float* __restrict__ hist, int nb, int steps, int first,
                                       int* __restrict__ seen_last) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nb) return;
  float h = first ? 0.f : hist[b];
  bool any = false;
  for (int s = 0; s < steps; ++s) {
    unsigned long long c;
    if (counts.wide) {
      unsigned long long* q = (unsigned long long*)counts.p + (int64_t)s * nb + b;
      c = *q;
      *q = 0ull;
    } else {
      unsigned int* q = (unsigned int*)counts.p + (int64_t)s * nb + b;
      c = *q;
      *q = 0u;
    }
    const float f = __ull2float_rn(c);                         // hist.astype("float32")  (:47)
    h = (first && s == 0) ? f : __fadd_rn(h, f);               // last_hist + hist        (:103-104)
    any |= (c != 0ull);
  }
  hist[b] = h;
  if (b == nb - 1 && seen_last != nullptr && any) seen_last[0] = 1;
}

// ---------------------------------------------------------------------------------------------
// KL threshold search: one block per candidate bin count i
// ---------------------------------------------------------------------------------------------
// The three left-to-right sums per candidate (Python's builtin sum: tail/total of P, the sum of Q, the
// divergence) are strictly sequential chains, so the kernel is bound by how many chains are in flight,
// not by bandwidth: small blocks (kKlThreads) and a small footprint (Q and the level buckets in shared
// memory, the histogram read through L1) keep ~12 candidates resident per SM; the tail chain and the
// prefix chain of P run on different warps at the same time.
constexpr int kKlThreads = 64;

template <bool LEGACY>
__global__ void __launch_bounds__(kKlThreads) kl_candidate_kernel(const float* __restrict__ hist_all, int n_data,
                                                                  int levels, int min_bins, int bins,
                                                                  double* __restrict__ div_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int i = min_bins + blockIdx.x;
  const float* __restrict__ H = hist_all + (int64_t)blockIdx.y * n_data;
  double* div = div_all + (int64_t)blockIdx.y * bins;
  double* cand = reinterpret_cast<double*>(smem_raw);                 // [levels]
  double* Q = cand + levels;                                          // [bins]
  __shared__ float s_last, s_total;
  __shared__ double s_tail, s_prefix, s_qsum;
  const int tid = threadIdx.x, nt = blockDim.x;

  // P: tail mass folded into bin i-1, then normalised.  Python's builtin sum is strictly left to
  // right; its accumulator is float32 under NEP 50 and float64 under legacy promotion (:143-145).
  if (tid == 0) {            // sum(data[i:])
    if (LEGACY) {
      double tail = 0.0;
#pragma unroll 8
      for (int j = i; j < n_data; ++j) tail = __dadd_rn(tail, (double)__ldg(H + j));
      s_tail = tail;
    } else {
      float tail = 0.f;
#pragma unroll 8
      for (int j = i; j < n_data; ++j) tail = __fadd_rn(tail, __ldg(H + j));
      s_tail = (double)tail;
    }
  } else if (tid == 32) {    // the first i-1 terms of sum(ref_distribution), which the tail does not touch
    if (LEGACY) {
      double pre = 0.0;
#pragma unroll 8
      for (int j = 0; j < i - 1; ++j) pre = __dadd_rn(pre, (double)__ldg(H + j));
      s_prefix = pre;
    } else {
      float pre = 0.f;
#pragma unroll 8
      for (int j = 0; j < i - 1; ++j) pre = __fadd_rn(pre, __ldg(H + j));
      s_prefix = (double)pre;
    }
  }
  // Q buckets: cand[k] = sum of hist[j] with floor(j*levels/i) == k, float64, ascending j (:149-152)
  for (int k = tid; k < levels; k += nt) {
    const int j0 = (int)(((long long)k * i + levels - 1) / levels);
    const int j1 = (int)(((long long)(k + 1) * i + levels - 1) / levels);
    double c = 0.0;
    for (int j = j0; j < j1 && j < i; ++j) c = __dadd_rn(c, (double)__ldg(H + j));
    cand[k] = c;
  }
  __syncthreads();
  if (tid == 0) {
    if (LEGACY) {
      const float last = (float)__dadd_rn((double)__ldg(H + i - 1), s_tail);
      s_last = last;
      s_total = (float)__dadd_rn(s_prefix, (double)last);   // float32 array /= float64 scalar: the scalar is cast first
    } else {
      const float last = __fadd_rn(__ldg(H + i - 1), (float)s_tail);
      s_last = last;
      s_total = __fadd_rn((float)s_prefix, last);
    }
  }
  __syncthreads();
  const float total = s_total, last = s_last;
  for (int j = tid; j < i; j += nt) {
    const float p = __fdiv_rn(j == i - 1 ? last : __ldg(H + j), total);
    // linear interpolation between the neighbouring buckets (:154-158), no FMA contraction
    const double t = __ddiv_rn((double)((long long)j * levels), (double)i);
    const int fl = (int)t;
    int ce = (int)ceil(t);
    ce = ce > levels - 1 ? levels - 1 : ce;
    double q = __dadd_rn(__dmul_rn(__dsub_rn(cand[ce], cand[fl]), __dsub_rn(t, (double)fl)), cand[fl]);
    q = __dmul_rn(q, (p != 0.f) ? 1.0 : 0.0);              // Q *= (P != 0)   (:159)
    Q[j] = q;
  }
  __syncthreads();
  if (tid == 0) {
    double qs = 0.0;
#pragma unroll 8
    for (int j = 0; j < i; ++j) qs = __dadd_rn(qs, Q[j]);
    s_qsum = qs;
  }
  __syncthreads();
  const double qsum = s_qsum;
  for (int j = tid; j < i; j += nt) {
    const double qn = __ddiv_rn(Q[j], qsum);
    double term = 0.0;                                       // entries with Q == 0 are dropped (:164-165)
    if (qn != 0.0) {
      const double p = (double)__fdiv_rn(j == i - 1 ? last : __ldg(H + j), total);
      term = __dmul_rn(p, log(__ddiv_rn(p, qn)));
    }
    Q[j] = term;
  }
  __syncthreads();
  if (tid == 0) {
    double d = 0.0;
#pragma unroll 8
    for (int j = 0; j < i; ++j) d = __dadd_rn(d, Q[j]);
    div[i] = d;
  }
}

// ---------------------------------------------------------------------------------------------
// KL search, 32 candidates per block: the sequential sums become LANE-PARALLEL chains.
// ---------------------------------------------------------------------------------------------
// In the block-per-candidate kernel above, 40 % of all warp instructions are issued for a single active
// lane (the three left-to-right sums).  Here a block owns 32 consecutive candidates i0 .. i0+31, lane c of
// every warp works for candidate i0+c, and the chains of all 32 run in one warp at once: every chain step is
// one full-width instruction instead of 32 nearly empty ones.  Q (pass 1) and the divergence terms (pass 2)
// are produced by all eight warps in 32-bin chunks into a double-buffered shared-memory ring
// [bin][candidate] and consumed by warp 0; each value is computed with exactly the operations of the
// reference and added in exactly its order, so the results are bit-identical to kl_candidate_kernel.
// Correctly rounded a / b with a reciprocal that is computed once: y = RN(1/b), q0 = RN(a*y),
// r = a - b*q0 (exact in one FMA), q = RN(q0 + r*y).  With a correctly rounded y and the faithful q0 this is
// Markstein's correction step and yields RN(a/b) (checked against exact rational arithmetic in
// tests/test_fast_quotient_math.py, all-ones divisors included); three FMA-class instructions instead of the
// ~35 of an IEEE double division whose divisor never changes (the candidate size i, the sum of Q).
struct DDiv {
  double b, y;
  __device__ __forceinline__ static DDiv make(double b) {
    DDiv d;
    d.b = b;
    d.y = __drcp_rn(b);
    return d;
  }
  __device__ __forceinline__ double div(double a) const {
    const double q0 = __dmul_rn(a, y);
    const double r = __fma_rn(-b, q0, a);
    return __fma_rn(r, y, q0);
  }
};

constexpr int kKlGroup = 32;
constexpr int kKlGroupThreads = 512;      // 16 warps: two resident blocks give 32 warps per SM

template <bool LEGACY>
__global__ void __launch_bounds__(kKlGroupThreads) kl_group_kernel(const float* __restrict__ hist_all, int n_data,
                                                                   int levels, int min_bins, int bins,
                                                                   double* __restrict__ div_all) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const float* __restrict__ hist = hist_all + (int64_t)blockIdx.y * n_data;
  double* div = div_all + (int64_t)blockIdx.y * bins;
  double* cand = reinterpret_cast<double*>(smem_raw);                  // [levels][32]
  double* ring = cand + (size_t)levels * kKlGroup;                     // [2][32 bins][32 candidates]
  float* H = reinterpret_cast<float*>(ring + 2 * 32 * kKlGroup);       // [n_data]
  __shared__ double s_tail[kKlGroup], s_prefix[kKlGroup], s_qsum[kKlGroup];
  __shared__ float s_last[kKlGroup], s_total[kKlGroup];
  const int tid = threadIdx.x, c = tid & 31, w = tid >> 5;
  const int i_lo = min_bins + (int)blockIdx.x * kKlGroup;
  const int i_hi = min(i_lo + kKlGroup - 1, bins - 1);                 // largest candidate of the block
  const bool valid = i_lo + c < bins;
  const int i = valid ? i_lo + c : i_hi;                               // spare lanes shadow the last candidate

  for (int j = tid; j < n_data; j += kKlGroupThreads) H[j] = hist[j];
  __syncthreads();

  // sum(data[i:]) on warp 0 and the first i-1 terms of sum(ref_distribution) on warp 1, lane = candidate;
  // float32 accumulators under NEP 50, float64 under legacy promotion (distribution_calibrate.py:143-145)
  if (w == 0) {
    if (LEGACY) {
      double acc = 0.0;
      for (int j = i_lo; j < n_data; ++j) {
        const double h = (double)H[j];
        if (j >= i) acc = __dadd_rn(acc, h);
      }
      s_tail[c] = acc;
    } else {
      float acc = 0.f;
      for (int j = i_lo; j < n_data; ++j) {
        const float h = H[j];
        if (j >= i) acc = __fadd_rn(acc, h);
      }
      s_tail[c] = (double)acc;
    }
  } else if (w == 1) {
    if (LEGACY) {
      double acc = 0.0;
      for (int j = 0; j < i_hi - 1; ++j) {
        const double h = (double)H[j];
        if (j < i - 1) acc = __dadd_rn(acc, h);
      }
      s_prefix[c] = acc;
    } else {
      float acc = 0.f;
      for (int j = 0; j < i_hi - 1; ++j) {
        const float h = H[j];
        if (j < i - 1) acc = __fadd_rn(acc, h);
      }
      s_prefix[c] = (double)acc;
    }
  }
  // level buckets of every candidate: cand[k][c] = sum of hist[j] with floor(j*levels/i_c) == k (:149-152)
  for (int p = tid; p < levels * kKlGroup; p += kKlGroupThreads) {
    const int cc = p & 31, k = p >> 5;
    const int ii = min(i_lo + cc, bins - 1);
    const int j0 = (int)(((long long)k * ii + levels - 1) / levels);
    const int j1 = (int)(((long long)(k + 1) * ii + levels - 1) / levels);
    double acc = 0.0;
    for (int j = j0; j < j1 && j < ii; ++j) acc = __dadd_rn(acc, (double)H[j]);
    cand[p] = acc;
  }
  __syncthreads();
  if (w == 0) {
    if (LEGACY) {
      const float last = (float)__dadd_rn((double)H[i - 1], s_tail[c]);
      s_last[c] = last;
      s_total[c] = (float)__dadd_rn(s_prefix[c], (double)last);
    } else {
      const float last = __fadd_rn(H[i - 1], (float)s_tail[c]);
      s_last[c] = last;
      s_total[c] = __fadd_rn((float)s_prefix[c], last);
    }
  }
  __syncthreads();
  const float total = s_total[c], last = s_last[c];
  const DDiv by_i = DDiv::make((double)i);

  // Q_j of candidate i (:154-159), the same values as in kl_candidate_kernel
  auto q_of = [&](int j, float& p) -> double {
    p = __fdiv_rn(j == i - 1 ? last : H[j], total);
    const double t = by_i.div((double)((long long)j * levels));
    const int fl = (int)t;
    int ce = (int)ceil(t);
    ce = ce > levels - 1 ? levels - 1 : ce;
    const double cf = cand[fl * kKlGroup + c], cc2 = cand[ce * kKlGroup + c];
    double q = __dadd_rn(__dmul_rn(__dsub_rn(cc2, cf), __dsub_rn(t, (double)fl)), cf);
    return __dmul_rn(q, (p != 0.f) ? 1.0 : 0.0);
  };

  const int n_chunks = (i_hi + 31) / 32;          // bins 0 .. i_hi-1
  for (int pass = 0; pass < 2; ++pass) {
    const DDiv by_qsum = DDiv::make(pass ? s_qsum[c] : 1.0);
    double acc = 0.0;                              // warp 0: the running left-to-right sum of candidate i
    for (int ch = 0; ch <= n_chunks; ++ch) {
      if (ch < n_chunks) {                         // produce chunk ch
        double* buf = ring + (size_t)(ch & 1) * 32 * kKlGroup;
#pragma unroll
        for (int r = 0; r < 32 / (kKlGroupThreads / 32); ++r) {
          const int jj = w + (kKlGroupThreads / 32) * r, j = ch * 32 + jj;
          double v = 0.0;
          if (j < i) {
            float p;
            const double q = q_of(j, p);
            if (pass == 0) {
              v = q;
            } else {                               // entries with Q == 0 are dropped (:164-165)
              const double qn = by_qsum.div(q);
              if (qn != 0.0) {
                const double pd = (double)p;
                v = __dmul_rn(pd, log(__ddiv_rn(pd, qn)));
              }
            }
          }
          buf[jj * kKlGroup + c] = v;
        }
      }
      if (w == 0 && ch > 0) {                      // consume chunk ch-1 (written before the last barrier)
        const double* buf = ring + (size_t)((ch - 1) & 1) * 32 * kKlGroup;
        const int jbase = (ch - 1) * 32;
#pragma unroll 8
        for (int jj = 0; jj < 32; ++jj) {
          const double v = buf[jj * kKlGroup + c];
          if (jbase + jj < i) acc = __dadd_rn(acc, v);
        }
      }
      __syncthreads();
    }
    if (w == 0) {
      if (pass == 0) s_qsum[c] = acc;
      else if (valid) div[i] = acc;
    }
    __syncthreads();
  }
}

// first strict minimum, NaN never wins (:167-169).  margin[l] (optional) = (runner-up - best) / |best|, the relative
// gap to the smallest divergence of any OTHER candidate: D_i is a float64 sum of p*log(p/q) terms whose `log` may
// differ from the reference's libm in the last place (relative 1e-16 per term), so a margin below ~1e-12 means the
// reference's own choice between the two candidates depends on its math library; callers flag margins < 1e-9.
__global__ void kl_argmin_kernel(const double* __restrict__ div_all, int min_bins, int bins, int* __restrict__ best,
                                 double* __restrict__ margin) {
  __shared__ double sv[kThreads], s2[kThreads];
  __shared__ int si[kThreads];
  const double* div = div_all + (int64_t)blockIdx.x * bins;
  double bv = INFINITY, second = INFINITY;
  int bi = min_bins;
  for (int i = min_bins + threadIdx.x; i < bins; i += blockDim.x) {
    const double d = div[i];
    if (d < bv) {
      second = bv;
      bv = d;
      bi = i;
    } else if (d < second) {
      second = d;
    }
  }
  sv[threadIdx.x] = bv;
  s2[threadIdx.x] = second;
  si[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      const double ov = sv[threadIdx.x + o], o2 = s2[threadIdx.x + o];
      const int oi = si[threadIdx.x + o];
      const double mv = sv[threadIdx.x], m2 = s2[threadIdx.x];
      const bool other_wins = ov < mv || (ov == mv && oi < si[threadIdx.x]);
      const double loser = other_wins ? mv : ov;
      s2[threadIdx.x] = fmin(loser, fmin(m2, o2));         // fmin ignores NaN: a NaN divergence never counts
      if (other_wins) {
        sv[threadIdx.x] = ov;
        si[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    best[blockIdx.x] = (sv[0] < INFINITY) ? si[0] : min_bins;
    if (margin != nullptr) {
      // both exactly 0: P == Q term by term for both candidates (a histogram with one occupied bin), and log(1) is
      // 0 in every libm -- an exact tie that the reference's "first strict minimum" resolves the same way
      const bool exact_zero_tie = sv[0] == 0.0 && s2[0] == 0.0;
      margin[blockIdx.x] = (sv[0] < INFINITY && s2[0] < INFINITY && !exact_zero_tie)
                               ? (s2[0] - sv[0]) / fmax(fabs(sv[0]), 1e-300) : INFINITY;
    }
  }
}

__global__ void kl_threshold_kernel(const int* __restrict__ best, const float* __restrict__ fm_max, int bins,
                                    float* __restrict__ input_max, int n) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l < n) input_max[l] = (float)(((double)best[l] + 0.5) * ((double)fm_max[l] / (double)bins));
}

}  // namespace fq

using namespace fq;

extern "C" {

int fq_ema_update(const DLTensor* state_, const DLTensor* cur_, double momentum, int scalar_cur, int promotion,
                  void* stream) {
  View state, cur;
  FQ_TRY(view_of(state_, "fq_ema_update: state", false, &state));
  FQ_TRY(view_of(cur_, "fq_ema_update: cur", false, &cur));
  FQ_REQUIRE(state.is_f32() && cur.is_f32() && state.numel == cur.numel, "fq_ema_update: float32 tensors of equal size");
  FQ_REQUIRE(promotion == FQ_PROMOTION_LEGACY || promotion == FQ_PROMOTION_NEP50, "fq_ema_update: bad promotion");
  if (state.numel == 0) return 0;
  const int grid = (int)((state.numel + 255) / 256 > 1184 ? 1184 : (state.numel + 255) / 256);
  ema_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(state.as<float>(), cur.as<const float>(), state.numel,
                                                     1 - momentum, (float)momentum, scalar_cur, promotion);
  FQ_LAUNCH_CHECK("ema_kernel");
  return 0;
}

// Blocks of one launch shared out over its tensors in proportion to their sizes (at least one each; never more
// than the tensor has tiles).  Fills tb.first_block, returns the grid size.
static int hist_share_blocks(HistBatch* tb, int budget) {
  int64_t total = 0;
  for (int i = 0; i < tb->count; ++i) total += tb->n[i];
  int blocks = 0;
  for (int i = 0; i < tb->count; ++i) {
    tb->first_block[i] = blocks;
    if (tb->n[i] == 0) continue;
    int64_t share = (int64_t)((double)budget * (double)tb->n[i] / (double)total);
    const int64_t tiles = (tb->n[i] + 4LL * kHistTileVec - 1) / (4LL * kHistTileVec);
    if (share > tiles) share = tiles;
    if (share < 1) share = 1;
    blocks += (int)share;
  }
  tb->first_block[tb->count] = blocks;
  return blocks;
}

static int launch_hist_multi(const HistBatch& tb, int blocks, const float* maxes, int max_stride, int max_offset,
                             int bins, int promotion, CountsOut counts, int* flags, cudaStream_t st) {
  const size_t smem = sizeof(unsigned int) * (bins + 1 + 32);
  if (flags != nullptr)
    hist_multi_kernel<true><<<blocks, kHistThreads, smem, st>>>(tb, maxes, max_stride, max_offset, bins, promotion,
                                                                counts, flags);
  else
    hist_multi_kernel<false><<<blocks, kHistThreads, smem, st>>>(tb, maxes, max_stride, max_offset, bins, promotion,
                                                                 counts, nullptr);
  FQ_LAUNCH_CHECK("hist_multi_kernel");
  return 0;
}

// counts tensors: (u)int64 or (u)int32, `n` elements
static bool counts_ok(const char* who, const View& c, int64_t n, CountsOut* out) {
  if ((c.code == kDLInt || c.code == kDLUInt) && (c.bits == 64 || c.bits == 32) && c.numel == n) {
    out->p = c.data;
    out->wide = c.bits == 64;
    return true;
  }
  set_error("%s: counts must be (u)int64 or (u)int32 with %lld elements", who, (long long)n);
  return false;
}

static bool flags_ok(const char* who, const View& f, int64_t n) {
  if (f.null) return true;
  if (f.code == kDLInt && f.bits == 32 && f.numel == n) return true;
  set_error("%s: bad_flags must be int32 [%lld]", who, (long long)n);
  return false;
}

int fq_hist_nonzero(const DLTensor* x_, const DLTensor* max__, int bins, int promotion, const DLTensor* counts_,
                    const DLTensor* bad_flag_, void* stream) {
  View x, mx, counts, flag;
  FQ_TRY(view_of(bad_flag_, "fq_hist_nonzero: bad_flag", true, &flag));
  FQ_TRY(flags_ok("fq_hist_nonzero", flag, 1));
  FQ_TRY(view_of(x_, "fq_hist_nonzero: x", false, &x));
  FQ_TRY(view_of(max__, "fq_hist_nonzero: max_", false, &mx));
  FQ_TRY(view_of(counts_, "fq_hist_nonzero: counts", false, &counts));
  FQ_REQUIRE(x.is_f32() && mx.is_f32() && mx.numel >= 1, "fq_hist_nonzero: x and max_ must be float32");
  FQ_REQUIRE(bins >= 1 && bins <= 8192, "fq_hist_nonzero: bins=%d outside [1, 8192]", bins);
  CountsOut cout_;
  FQ_TRY(counts_ok("fq_hist_nonzero", counts, bins + 1, &cout_));
  FQ_REQUIRE(cout_.wide || x.numel < (1LL << 32), "fq_hist_nonzero: 32-bit counts need fewer than 2^32 elements");
  FQ_REQUIRE(promotion == FQ_PROMOTION_LEGACY || promotion == FQ_PROMOTION_NEP50, "fq_hist_nonzero: bad promotion");
  if (x.numel == 0) return 0;
  const size_t smem = sizeof(unsigned int) * (bins + 1 + 32);
  if (aligned16(x.data)) {
    HistBatch tb = {};
    tb.x[0] = x.as<const float>();
    tb.n[0] = x.numel;
    tb.count = 1;
    const int blocks = hist_share_blocks(&tb, sm_count() * kHistBlocksPerSM);
    FQ_TRY(launch_hist_multi(tb, blocks, mx.as<const float>(), 1, 0, bins, promotion, cout_,
                             flag.null ? nullptr : flag.as<int>(), (cudaStream_t)stream) == 0);
  } else {
    const int64_t b = (x.numel + kThreads - 1) / kThreads;
    const int grid = (int)(b > sm_count() * 8 ? sm_count() * 8 : b);
    hist_unaligned_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(
        x.as<const float>(), x.numel, mx.as<const float>(), bins, promotion, cout_,
        flag.null ? nullptr : flag.as<int>());
    FQ_LAUNCH_CHECK("hist_unaligned_kernel");
  }
  return 0;
}

int fq_hist_nonzero_multi(const DLTensor* const* xs, int n_tensors, const DLTensor* maxes_, int max_stride,
                          int max_offset, int bins, int promotion, const DLTensor* counts_,
                          const DLTensor* bad_flags_, void* stream) {
  const char* who = "fq_hist_nonzero_multi";
  View mx, counts, flags;
  FQ_TRY(view_of(bad_flags_, "fq_hist_nonzero_multi: bad_flags", true, &flags));
  FQ_TRY(flags_ok(who, flags, n_tensors));
  FQ_REQUIRE(xs != nullptr && n_tensors >= 1, "%s: no tensors", who);
  FQ_TRY(view_of(maxes_, "fq_hist_nonzero_multi: maxes", false, &mx));
  FQ_TRY(view_of(counts_, "fq_hist_nonzero_multi: counts", false, &counts));
  FQ_REQUIRE(bins >= 1 && bins <= 8192, "%s: bins=%d outside [1, 8192]", who, bins);
  FQ_REQUIRE(mx.is_f32() && max_stride >= 1 && max_offset >= 0 &&
                 mx.numel >= (int64_t)(n_tensors - 1) * max_stride + max_offset + 1,
             "%s: maxes must be float32 with an entry for every tensor", who);
  CountsOut cout_;
  FQ_TRY(counts_ok(who, counts, (int64_t)n_tensors * (bins + 1), &cout_));
  FQ_REQUIRE(promotion == FQ_PROMOTION_LEGACY || promotion == FQ_PROMOTION_NEP50, "%s: bad promotion", who);
  const int budget = sm_count() * kHistBlocksPerSM;
  for (int base = 0; base < n_tensors; base += FQ_MAX_BATCH) {
    const int cnt = (n_tensors - base < FQ_MAX_BATCH) ? n_tensors - base : FQ_MAX_BATCH;
    HistBatch tb = {};
    for (int i = 0; i < cnt; ++i) {
      View x;
      FQ_TRY(view_of(xs[base + i], "fq_hist_nonzero_multi: x", false, &x));
      FQ_REQUIRE(x.is_f32() && aligned16(x.data), "%s: tensor %d must be float32 and 16-byte aligned", who, base + i);
      FQ_REQUIRE(cout_.wide || x.numel < (1LL << 32), "%s: 32-bit counts need fewer than 2^32 elements per tensor", who);
      tb.x[i] = x.as<const float>();
      tb.n[i] = x.numel;
    }
    tb.count = cnt;
    const int blocks = hist_share_blocks(&tb, budget);
    if (blocks == 0) continue;
    FQ_TRY(launch_hist_multi(tb, blocks, mx.as<const float>() + (int64_t)base * max_stride, max_stride, max_offset, bins,
                             promotion, cout_.at((int64_t)base * (bins + 1)),
                             flags.null ? nullptr : flags.as<int>() + base, (cudaStream_t)stream) == 0);
  }
  return 0;
}

int fq_hist_accumulate_f32(const DLTensor* counts_, const DLTensor* hist_, int first, const DLTensor* seen_last_,
                           void* stream) {
  View counts, hist, seen;
  FQ_TRY(view_of(counts_, "fq_hist_accumulate_f32: counts", false, &counts));
  FQ_TRY(view_of(hist_, "fq_hist_accumulate_f32: hist", false, &hist));
  FQ_TRY(view_of(seen_last_, "fq_hist_accumulate_f32: seen_last", true, &seen));
  FQ_REQUIRE((counts.code == kDLInt || counts.code == kDLUInt) && (counts.bits == 64 || counts.bits == 32),
             "fq_hist_accumulate_f32: counts must be (u)int64 or (u)int32");
  CountsOut cout_;
  cout_.p = counts.data;
  cout_.wide = counts.bits == 64;
  FQ_REQUIRE(hist.is_f32() && hist.numel > 0 && counts.numel >= hist.numel && counts.numel % hist.numel == 0 &&
                 hist.numel <= INT32_MAX && counts.numel / hist.numel <= INT32_MAX,
             "fq_hist_accumulate_f32: hist must be float32 [n] and counts [steps, n]");
  FQ_REQUIRE(seen.null || (seen.code == kDLInt && seen.bits == 32 && seen.numel >= 1), "fq_hist_accumulate_f32: seen_last must be int32");
  const int nb = (int)hist.numel;
  hist_accumulate_kernel<<<(nb + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      cout_, hist.as<float>(), nb, (int)(counts.numel / hist.numel), first,
      seen.null ? nullptr : seen.as<int>());
  FQ_LAUNCH_CHECK("hist_accumulate_kernel");
  return 0;
}

int fq_kl_search(const DLTensor* hist_, int levels, int min_bins, int bins, int promotion, const DLTensor* best_,
                 const DLTensor* divergence_, const DLTensor* margin_, void* stream) {
  const char* who = "fq_kl_search";
  View hist, best, dv, mg;
  FQ_TRY(view_of(margin_, "fq_kl_search: margin", true, &mg));
  FQ_TRY(view_of(hist_, "fq_kl_search: hist", false, &hist));
  FQ_TRY(view_of(best_, "fq_kl_search: best", false, &best));
  FQ_TRY(view_of(divergence_, "fq_kl_search: divergence", false, &dv));
  FQ_REQUIRE(hist.is_f32() && hist_->ndim >= 1, "%s: hist must be float32 [n_data] or [layers, n_data]", who);
  const int n_data = (int)hist_->shape[hist_->ndim - 1];
  const int layers = n_data > 0 ? (int)(hist.numel / n_data) : 0;
  // the reference asserts min_bins >= levels (:133)
  FQ_REQUIRE(levels >= 1 && min_bins >= levels, "%s: min_bins should be greater than levels (%d vs. %d)", who, min_bins, levels);
  FQ_REQUIRE(bins <= 8192 && n_data >= bins && n_data <= bins + 1, "%s: hist length %d must be bins or bins+1 (bins=%d)", who,
             n_data, bins);
  FQ_REQUIRE(best.code == kDLInt && best.bits == 32 && best.numel == layers, "%s: best must be int32 [layers=%d]", who, layers);
  FQ_REQUIRE(dv.code == kDLFloat && dv.bits == 64 && dv.numel == (int64_t)layers * bins,
             "%s: divergence must be float64 [layers, bins] scratch", who);
  FQ_REQUIRE(promotion == FQ_PROMOTION_LEGACY || promotion == FQ_PROMOTION_NEP50, "%s: bad promotion", who);
  FQ_REQUIRE(mg.null || (mg.code == kDLFloat && mg.bits == 64 && mg.numel == layers), "%s: margin must be float64 [layers=%d]",
             who, layers);
  if (layers == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int ncand = bins - min_bins;
  const size_t group_smem = sizeof(double) * ((size_t)levels * kKlGroup + 2 * 32 * kKlGroup) + sizeof(float) * (n_data + 4);
  if (ncand > 0 && group_smem <= 200 * 1024) {
    auto kern = (promotion == FQ_PROMOTION_LEGACY) ? kl_group_kernel<true> : kl_group_kernel<false>;
    FQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)group_smem));
    kern<<<dim3((ncand + kKlGroup - 1) / kKlGroup, layers), kKlGroupThreads, group_smem, st>>>(
        hist.as<const float>(), n_data, levels, min_bins, bins, dv.as<double>());
    FQ_LAUNCH_CHECK("kl_group_kernel");
  } else if (ncand > 0) {        // very many levels: one block per candidate
    const size_t smem = sizeof(double) * (levels + bins);
    auto kern = (promotion == FQ_PROMOTION_LEGACY) ? kl_candidate_kernel<true> : kl_candidate_kernel<false>;
    FQ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3(ncand, layers), kKlThreads, smem, st>>>(hist.as<const float>(), n_data, levels, min_bins, bins,
                                                      dv.as<double>());
    FQ_LAUNCH_CHECK("kl_candidate_kernel");
  }
  kl_argmin_kernel<<<layers, kThreads, 0, st>>>(dv.as<const double>(), min_bins, bins, best.as<int>(),
                                                mg.null ? nullptr : mg.as<double>());
  FQ_LAUNCH_CHECK("kl_argmin_kernel");
  return 0;
}

int fq_kl_threshold(const DLTensor* best_, const DLTensor* fm_max_, int bins, const DLTensor* input_max_,
                    void* stream) {
  View best, mx, im;
  FQ_TRY(view_of(best_, "fq_kl_threshold: best", false, &best));
  FQ_TRY(view_of(fm_max_, "fq_kl_threshold: fm_max", false, &mx));
  FQ_TRY(view_of(input_max_, "fq_kl_threshold: input_max", false, &im));
  FQ_REQUIRE(best.code == kDLInt && best.bits == 32 && mx.is_f32() && im.is_f32(), "fq_kl_threshold: best int32, fm_max/input_max float32");
  FQ_REQUIRE(best.numel == mx.numel && best.numel == im.numel && best.numel > 0, "fq_kl_threshold: sizes differ");
  FQ_REQUIRE(bins >= 1, "fq_kl_threshold: bins must be positive");
  const int n = (int)best.numel;
  kl_threshold_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(best.as<const int>(), mx.as<const float>(), bins,
                                                                        im.as<float>(), n);
  FQ_LAUNCH_CHECK("kl_threshold_kernel");
  return 0;
}

}  // extern "C"
