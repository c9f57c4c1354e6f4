// K2/K3/K6: elementwise fake-quant forward, STE backward, integer export.
// HBM-bound streaming kernels: 128-bit loads/stores with streaming cache hints, kUnroll independent
// loads in flight per thread, one contiguous 16 B aligned slice per block, grid = k * SM count.
//   reference: quantize/convert/ste_func.py:37-44; freeze.py:100-103; nn/quantized_conv.py:54-76
#include "fq_fused.cuh"

namespace fq {

// ---- code (rounded quotient) sinks -----------------------------------------------------------
struct NoCode {
  static constexpr int kBytes = 0;
  __device__ __forceinline__ void put4(int64_t, float4) const {}
  __device__ __forceinline__ void put1(int64_t, float) const {}
};
template <class T>
struct IntCode {
  static constexpr int kBytes = (int)sizeof(T);
  T* p;
  __device__ __forceinline__ void put1(int64_t i, float c) const { p[i] = (T)(int)c; }
  __device__ __forceinline__ void put4(int64_t i, float4 c) const {
    if constexpr (sizeof(T) == 1) {
      uchar4 o = make_uchar4((unsigned char)(T)(int)c.x, (unsigned char)(T)(int)c.y, (unsigned char)(T)(int)c.z,
                             (unsigned char)(T)(int)c.w);
      *reinterpret_cast<uchar4*>(p + i) = o;
    } else if constexpr (sizeof(T) == 2) {
      ushort4 o = make_ushort4((unsigned short)(T)(int)c.x, (unsigned short)(T)(int)c.y, (unsigned short)(T)(int)c.z,
                               (unsigned short)(T)(int)c.w);
      *reinterpret_cast<ushort4*>(p + i) = o;
    } else {
      int4 o = make_int4((int)c.x, (int)c.y, (int)c.z, (int)c.w);
      *reinterpret_cast<int4*>(p + i) = o;
    }
  }
};
struct FloatCode {
  static constexpr int kBytes = 4;
  float* p;
  __device__ __forceinline__ void put1(int64_t i, float c) const { p[i] = c; }
  __device__ __forceinline__ void put4(int64_t i, float4 c) const { st_stream(reinterpret_cast<float4*>(p + i), c); }
};

// ---- y = roundf(clip(x) / d) * s with scalar qparams --------------------------------------------
template <bool CLIP, class Code>
struct ScalarQuant {
  float d, s, lo, hi;
  float* y;
  Code code;
  QDiv q;
  __device__ __forceinline__ void prepare() { q = QDiv::make(d); }
  __device__ __forceinline__ float one(float x, float& c) const {
    if (CLIP) x = clipf(x, lo, hi);
    c = q.code(x);
    return __fmul_rn(c, s);
  }
  __device__ __forceinline__ void vec(int64_t i, float4 v) const {
    if (CLIP) v = make_float4(clipf(v.x, lo, hi), clipf(v.y, lo, hi), clipf(v.z, lo, hi), clipf(v.w, lo, hi));
    const float4 c = q.code4(v);
    st_stream(reinterpret_cast<float4*>(y + i),
              make_float4(__fmul_rn(c.x, s), __fmul_rn(c.y, s), __fmul_rn(c.z, s), __fmul_rn(c.w, s)));
    code.put4(i, c);
  }
  // y only; the codes are returned (8-bit codes leave through a shared-memory stage, see the kernel)
  __device__ __forceinline__ float4 vec_y(int64_t i, float4 v) const {
    if (CLIP) v = make_float4(clipf(v.x, lo, hi), clipf(v.y, lo, hi), clipf(v.z, lo, hi), clipf(v.w, lo, hi));
    const float4 c = q.code4(v);
    st_stream(reinterpret_cast<float4*>(y + i),
              make_float4(__fmul_rn(c.x, s), __fmul_rn(c.y, s), __fmul_rn(c.z, s), __fmul_rn(c.w, s)));
    return c;
  }
  __device__ __forceinline__ void sca(int64_t i, float v) const {
    float c;
    y[i] = one(v, c);
    code.put1(i, c);
  }
};

// One tile of 4 * kThreads * kUnroll elements per block, no loop: like a plain copy kernel the grid
// is as large as the tensor, registers stay low enough for 6+ resident blocks per SM and every
// thread has kUnroll 16 B loads in flight before it touches the data.
// PDL = true: launched as a programmatic dependent of the range kernel (fq_forward_online on small, latency-bound
// tensors).  The kernel then starts while the range kernel's last block is still computing the qparams: the first
// tile of x -- which the range kernel only reads -- is loaded before the wait.  One thread per block waits and
// reads the qparams past L1 (.cg; L1 may still hold this layer's values of the previous step) and shares them
// through shared memory: every warp of every block asking one L2 slice for the same 16 bytes costs microseconds
// even at a few hundred blocks (and halves the throughput at 2^30 elements), so large tensors keep the plain
// launch with L1-cached qparams.
template <bool CLIP, class Code, bool PDL = false>
__global__ void __launch_bounds__(kThreads, PDL ? 4 : 6) forward_scalar_kernel(const float* __restrict__ x, int64_t n,
                                                                     int64_t per_block,
                                                                     const float* __restrict__ qp_dev,
                                                                     ScalarQuant<CLIP, Code> op, int vectorised,
                                                                     int reverse) {
  __shared__ float s_qp[PDL ? 4 : 1];
  auto wait_for_qparams = [&]() {
    if (threadIdx.x == 0) {
      pdl_wait();
      s_qp[0] = __ldcg(qp_dev + FQ_QP_D);
      s_qp[1] = __ldcg(qp_dev + FQ_QP_S);
      s_qp[2] = __ldcg(qp_dev + FQ_QP_LO);
      s_qp[PDL ? 3 : 0] = __ldcg(qp_dev + FQ_QP_HI);
    }
    __syncthreads();
    op.d = s_qp[0];
    op.s = s_qp[1];
    op.lo = s_qp[2];
    op.hi = s_qp[PDL ? 3 : 0];
    op.prepare();
  };
  if constexpr (!PDL) {
    if (qp_dev != nullptr) {
      op.d = __ldg(qp_dev + FQ_QP_D);
      op.s = __ldg(qp_dev + FQ_QP_S);
      op.lo = __ldg(qp_dev + FQ_QP_LO);
      op.hi = __ldg(qp_dev + FQ_QP_HI);
    }
    op.prepare();
  }
  if (vectorised) {
    const int64_t nvec = n >> 2;
    const float4* p4 = reinterpret_cast<const float4*>(x);
    const int64_t ntiles = (nvec + kTileElems / 4 - 1) / (kTileElems / 4);
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const int64_t tile = reverse ? (ntiles - 1 - t) : t;
      const int64_t v0 = tile * (kTileElems / 4) + threadIdx.x;
      if ((tile + 1) * (kTileElems / 4) <= nvec) {        // block-uniform: the barrier below needs that
        float4 v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) v[u] = ld_stream(p4 + v0 + u * kThreads);
        if constexpr (PDL) {
          if (t == blockIdx.x) wait_for_qparams();       // first tile: its loads are already in flight
        }
        if constexpr (Code::kBytes == 1) {
          // 8-bit codes: four 4-byte stores per thread would be the slowest stream of the kernel; the tile's 4096
          // codes are staged in shared memory and leave as one 16-byte streaming store per thread
          __shared__ __align__(16) uchar4 stage[kThreads * kUnroll];
          if (t != (int64_t)blockIdx.x) __syncthreads();               // the previous tile's stage has been read
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) {
            const float4 c = op.vec_y(4 * (v0 + u * kThreads), v[u]);
            stage[threadIdx.x + u * kThreads] = make_uchar4((unsigned char)(int)c.x, (unsigned char)(int)c.y,
                                                            (unsigned char)(int)c.z, (unsigned char)(int)c.w);
          }
          __syncthreads();
          const uint4 packed = reinterpret_cast<const uint4*>(stage)[threadIdx.x];
          __stcs(reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(op.code.p) + tile * kTileElems) + threadIdx.x, packed);
        } else {
#pragma unroll
          for (int u = 0; u < kUnroll; ++u) op.vec(4 * (v0 + u * kThreads), v[u]);
        }
      } else {
        if constexpr (PDL) {
          if (t == blockIdx.x) wait_for_qparams();
        }
        for (int u = 0; u < kUnroll; ++u) {
          const int64_t j = v0 + u * kThreads;
          if (j < nvec) op.vec(4 * j, ld_stream(p4 + j));
        }
      }
    }
    if constexpr (PDL) {
      if ((int64_t)blockIdx.x >= ntiles) wait_for_qparams();     // a block with nothing but the scalar tail
    }
    const int64_t tail0 = nvec << 2;
    if (blockIdx.x == 0 && threadIdx.x < n - tail0) op.sca(tail0 + threadIdx.x, x[tail0 + threadIdx.x]);
  } else {   // some pointer is not 16 B aligned: plain scalar grid-stride loop
    if constexpr (PDL) wait_for_qparams();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      op.sca(i, x[i]);
  }
}

// ---- per-row scale tensor (weights): y = roundf(x / (s_r + 1e-10)) * s_r -------------------------
// Long rows (L >= one tile, so a tile meets at most two rows): one tile per block like the scalar kernel,
// with the two candidate rows' quantisers held in registers.
template <class Code>
__global__ void __launch_bounds__(kThreads, 6) forward_rows_tiles_kernel(const float* __restrict__ x, int64_t n,
                                                                         int64_t L, int64_t rows,
                                                                         const float* __restrict__ scale,
                                                                         float* __restrict__ y, Code code) {
  // `rows` comes from the host: n / L in every thread is a full 64-bit division once n >= 2^32 (~150 instructions
  // against ~190 for the whole tile), which made this kernel issue-bound on 16 GB tensors (6.1 instead of 6.9 TB/s)
  const int64_t nvec = n >> 2;
  const float4* p4 = reinterpret_cast<const float4*>(x);
  for (int64_t tile = blockIdx.x; tile * (kTileElems / 4) < nvec; tile += gridDim.x) {
    const int64_t first = tile * kTileElems;
    const int64_t row0 = first / L;
    const int64_t boundary = (row0 + 1) * L;
    const float s0 = __ldg(scale + row0);
    const float s1 = __ldg(scale + min(row0 + 1, rows - 1));
    const QDiv q0 = QDiv::make(__fadd_rn(s0, 1e-10f)), q1 = QDiv::make(__fadd_rn(s1, 1e-10f));
    const int64_t v0 = tile * (kTileElems / 4) + threadIdx.x;
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t j = v0 + u * kThreads;
      v[u] = j < nvec ? ld_stream(p4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t j = v0 + u * kThreads;
      if (j >= nvec) continue;
      const int64_t i = 4 * j;
      float4 c, o;
      if (i + 3 < boundary) {
        c = q0.code4(v[u]);
        o = make_float4(__fmul_rn(c.x, s0), __fmul_rn(c.y, s0), __fmul_rn(c.z, s0), __fmul_rn(c.w, s0));
      } else if (i >= boundary) {
        c = q1.code4(v[u]);
        o = make_float4(__fmul_rn(c.x, s1), __fmul_rn(c.y, s1), __fmul_rn(c.z, s1), __fmul_rn(c.w, s1));
      } else {       // the float4 straddles the row boundary (L % 4 != 0)
        const float e[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        float cc[4], oo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const bool lo_row = i + t < boundary;
          cc[t] = lo_row ? q0.code(e[t]) : q1.code(e[t]);
          oo[t] = __fmul_rn(cc[t], lo_row ? s0 : s1);
        }
        c = make_float4(cc[0], cc[1], cc[2], cc[3]);
        o = make_float4(oo[0], oo[1], oo[2], oo[3]);
      }
      st_stream(reinterpret_cast<float4*>(y + i), o);
      code.put4(i, c);
    }
  }
  const int64_t tail0 = nvec << 2;
  if (blockIdx.x == 0 && threadIdx.x < n - tail0) {
    const int64_t i = tail0 + threadIdx.x;
    const float s = __ldg(scale + i / L);
    const float c = quant_code(x[i], __fadd_rn(s, 1e-10f));
    y[i] = __fmul_rn(c, s);
    code.put1(i, c);
  }
}

template <class Code>
__global__ void __launch_bounds__(kThreads) forward_rows_kernel(const float* __restrict__ x, int64_t n, int64_t L,
                                                                int64_t per_block, const float* __restrict__ scale,
                                                                float* __restrict__ y, Code code, int vectorised) {
  if (vectorised && L >= 1024) {
    const int64_t begin = (int64_t)blockIdx.x * per_block;
    const int64_t end = min(n, begin + per_block);
    if (begin >= end) return;
    for (int64_t r = begin / L; r * L < end; ++r) {
      ScalarQuant<false, Code> op;
      op.s = __ldg(scale + r);
      op.d = __fadd_rn(op.s, 1e-10f);
      op.prepare();
      op.lo = op.hi = 0.f;
      op.y = y;
      op.code = code;
      for_range<false, false>(
          x, max(begin, r * L), min(end, (r + 1) * L), [&](int64_t i, float4 v) { op.vec(i, v); },
          [&](int64_t i, float v) { op.sca(i, v); });
    }
  } else {   // short rows (depthwise 3x3: L = 9) or misaligned pointers: one element per thread
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const float s = __ldg(scale + i / L);
      const float c = quant_code(x[i], __fadd_rn(s, 1e-10f));
      y[i] = __fmul_rn(c, s);
      code.put1(i, c);
    }
  }
}

// ---- STE backward with the clip mask (extension; the reference's backward is the identity) ------
__global__ void __launch_bounds__(kThreads, 6) ste_mask_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                               const float* __restrict__ qp, float* __restrict__ dx,
                                                               int64_t n, int64_t per_block, int vectorised) {
  const float lo = __ldg(qp + FQ_QP_LO), hi = __ldg(qp + FQ_QP_HI);
  auto m = [&](float g, float v) { return (v >= lo && v <= hi) ? g : 0.f; };
  if (vectorised) {
    const int64_t nvec = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(dy);
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int64_t tile = blockIdx.x; tile * (kTileElems / 4) < nvec; tile += gridDim.x) {
      const int64_t v0 = tile * (kTileElems / 4) + threadIdx.x;
      float4 g[kUnroll / 2], v[kUnroll / 2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {       // two half tiles: 2 x (kUnroll/2) x 2 loads in flight per thread
#pragma unroll
        for (int u = 0; u < kUnroll / 2; ++u) {
          const int64_t j = v0 + (h * (kUnroll / 2) + u) * kThreads;
          if (j < nvec) {
            g[u] = ld_stream(g4 + j);
            v[u] = ld_stream(x4 + j);
          }
        }
#pragma unroll
        for (int u = 0; u < kUnroll / 2; ++u) {
          const int64_t j = v0 + (h * (kUnroll / 2) + u) * kThreads;
          if (j < nvec)
            st_stream(reinterpret_cast<float4*>(dx) + j,
                      make_float4(m(g[u].x, v[u].x), m(g[u].y, v[u].y), m(g[u].z, v[u].z), m(g[u].w, v[u].w)));
        }
      }
    }
    const int64_t tail0 = nvec << 2;
    if (blockIdx.x == 0 && threadIdx.x < n - tail0) dx[tail0 + threadIdx.x] = m(dy[tail0 + threadIdx.x], x[tail0 + threadIdx.x]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
      dx[i] = m(dy[i], x[i]);
  }
}

// ---- MXNet contrib.quantize(out_type=int8), zero centred ---------------------------------------
__global__ void __launch_bounds__(kThreads) int8_export_kernel(const float* __restrict__ w, int64_t n,
                                                               const float* __restrict__ range2,
                                                               signed char* __restrict__ out,
                                                               float* __restrict__ out_range2) {
  const float real = fmaxf(fabsf(__ldg(range2)), fabsf(__ldg(range2 + 1)));
  const float scale = __fdiv_rn(127.0f, real);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = w[i];
    const float mag = fminf(__fadd_rn(__fmul_rn(fabsf(v), scale), 0.5f), 127.0f);
    const float sgn = v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f);
    out[i] = (signed char)(int)__fmul_rn(sgn, mag);      // C cast: truncation
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && out_range2 != nullptr) {
    out_range2[0] = -real;
    out_range2[1] = real;
  }
}

// ---- nn/quantized_conv.py:_quantize / dequantize ------------------------------------------------
__global__ void __launch_bounds__(kThreads) qconv_quantize_kernel(const float* __restrict__ x, int64_t n,
                                                                  const float* __restrict__ range2,
                                                                  int* __restrict__ codes, float* __restrict__ scale_out) {
  const float lo = __ldg(range2), hi = __ldg(range2 + 1);
  const float scale = (hi == -lo) ? __fdiv_rn(hi, 127.0f) : __fdiv_rn(__fsub_rn(hi, lo), 255.0f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    codes[i] = (int)quant_code(clipf(x[i], lo, hi), scale);
  if (blockIdx.x == 0 && threadIdx.x == 0 && scale_out != nullptr) scale_out[0] = scale;
}

__global__ void __launch_bounds__(kThreads) qconv_dequantize_kernel(const int* __restrict__ acc, int64_t n,
                                                                    const float* __restrict__ s_in,
                                                                    const float* __restrict__ s_w,
                                                                    float* __restrict__ y) {
  const float s = __fmul_rn(__ldg(s_in), __ldg(s_w));
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __fmul_rn((float)acc[i], s);
}

static inline int ew_grid(int64_t n) {
  int64_t b = (n + kThreads - 1) / kThreads;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// dispatch on the dtype of the optional `codes` tensor
template <class F>
static int with_code_sink(const char* who, const View& codes, int64_t n, F f) {
  if (codes.null) return f(NoCode());
  if (codes.numel != n) {
    set_error("%s: codes has %lld elements, expected %lld", who, (long long)codes.numel, (long long)n);
    return -1;
  }
  if (codes.code == kDLFloat && codes.bits == 32) return f(FloatCode{codes.as<float>()});
  if (codes.code == kDLInt && codes.bits == 8) return f(IntCode<signed char>{codes.as<signed char>()});
  if (codes.code == kDLUInt && codes.bits == 8) return f(IntCode<unsigned char>{codes.as<unsigned char>()});
  if (codes.code == kDLInt && codes.bits == 16) return f(IntCode<short>{codes.as<short>()});
  if (codes.code == kDLUInt && codes.bits == 16) return f(IntCode<unsigned short>{codes.as<unsigned short>()});
  if (codes.code == kDLInt && codes.bits == 32) return f(IntCode<int>{codes.as<int>()});
  set_error("%s: codes dtype (code %d, %d bits) unsupported; use int8/uint8/int16/uint16/int32/float32", who,
            codes.code, codes.bits);
  return -1;
}

// ---- A/B variant (FQ_FORWARD_BULK=1): the same quantiser fed by the bulk-async copy engine ------------------------
// Persistent blocks; a ring of 3 input tiles is filled with cp.async.bulk (global -> shared, mbarrier complete_tx), the
// threads read their float4s from shared memory, quantise, write the result tile to one of two output buffers
// (fence.proxy.async) and one thread sends it off with cp.async.bulk (shared -> global, bulk_group).  One barrier per
// tile.  Measured against the LDG.128 / STG.128 kernel above (profiles/README.md): kept only as the A/B it is.
constexpr int kBulkStages = 3;
__device__ __forceinline__ uint32_t bulk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(kThreads, 2) forward_scalar_bulk_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                                          int64_t ntiles, const float* __restrict__ qp_dev,
                                                                          float d, float s, float lo, float hi) {
  extern __shared__ __align__(128) unsigned char bulk_smem[];
  float4* in = reinterpret_cast<float4*>(bulk_smem);                                    // [stages][kTileElems / 4]
  float4* out = in + kBulkStages * (kTileElems / 4);                                    // [2][kTileElems / 4]
  __shared__ uint64_t full[kBulkStages];
  constexpr uint32_t kTileBytes = kTileElems * 4;
  if (qp_dev != nullptr) {
    d = __ldg(qp_dev + FQ_QP_D);
    s = __ldg(qp_dev + FQ_QP_S);
    lo = __ldg(qp_dev + FQ_QP_LO);
    hi = __ldg(qp_dev + FQ_QP_HI);
  }
  const QDiv qd = QDiv::make(d);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kBulkStages; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bulk_smem_u32(&full[i])) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto load_tile = [&](int64_t tile, int stage) {       // thread 0 only
    const uint32_t bar = bulk_smem_u32(&full[stage]);
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(bar), "r"(kTileBytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     bulk_smem_u32(in + (size_t)stage * (kTileElems / 4))),
                 "l"(x + tile * kTileElems), "r"(kTileBytes), "r"(bar)
                 : "memory");
  };
  const int64_t first = blockIdx.x, step = gridDim.x;
  if (threadIdx.x == 0)
    for (int i = 0; i < kBulkStages; ++i)
      if (first + i * step < ntiles) load_tile(first + i * step, i);
  int64_t it = 0;
  for (int64_t tile = first; tile < ntiles; tile += step, ++it) {
    const int stage = (int)(it % kBulkStages);
    const uint32_t parity = (uint32_t)(it / kBulkStages) & 1u;
    {
      const uint32_t bar = bulk_smem_u32(&full[stage]);
      asm volatile(
          "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
              bar),
          "r"(parity)
          : "memory");
    }
    const float4* src = in + (size_t)stage * (kTileElems / 4);
    float4* dst = out + (size_t)(it & 1) * (kTileElems / 4);
    float4 v[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) v[u] = src[threadIdx.x + u * kThreads];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const float4 c = qd.code4(make_float4(clipf(v[u].x, lo, hi), clipf(v[u].y, lo, hi), clipf(v[u].z, lo, hi),
                                            clipf(v[u].w, lo, hi)));
      dst[threadIdx.x + u * kThreads] = make_float4(__fmul_rn(c.x, s), __fmul_rn(c.y, s), __fmul_rn(c.z, s), __fmul_rn(c.w, s));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    // the previous tile's store must have finished READING its output buffer before anyone passes the barrier: the
    // next tile's results go into that buffer
    if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y + tile * kTileElems),
                   "r"(bulk_smem_u32(dst)), "r"(kTileBytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      const int64_t nxt = tile + kBulkStages * step;
      if (nxt < ntiles) load_tile(nxt, stage);
    }
  }
  if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

static int forward_scalar_impl(const char* who, const DLTensor* x_, const float* qp_dev, float d, float s, float lo,
                               float hi, bool clip, const DLTensor* y_, const DLTensor* codes_, void* stream,
                               bool reverse = false, bool dependent = false) {
  View x, y, codes;
  FQ_TRY(view_of(x_, who, false, &x));
  FQ_TRY(view_of(y_, who, false, &y));
  FQ_TRY(view_of(codes_, who, true, &codes));
  FQ_REQUIRE(x.is_f32() && y.is_f32(), "%s: x and y must be float32", who);
  FQ_REQUIRE(x.numel == y.numel, "%s: x has %lld elements, y %lld", who, (long long)x.numel, (long long)y.numel);
  if (x.numel == 0) return 0;
  const int64_t n = x.numel;
  const int vec = aligned16(x.data) && aligned16(y.data) && (codes.null || aligned16(codes.data));
  int64_t per_block = 0;
  const int grid = vec ? tile_grid(n, 1 << 30) : ew_grid(n);      // one tile per block
  cudaStream_t st = (cudaStream_t)stream;
  {
    static int bulk = -1;
    if (bulk < 0) {
      const char* env = getenv("FQ_FORWARD_BULK");
      bulk = (env != nullptr && env[0] == '1') ? 1 : 0;
    }
    if (bulk == 1 && clip && !dependent && !reverse && codes.null && vec && n % kTileElems == 0) {
      const size_t smem = (size_t)(kBulkStages + 2) * kTileElems * 4;
      FQ_CUDA(cudaFuncSetAttribute(forward_scalar_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const int64_t ntiles = n / kTileElems;
      const int64_t cap = (int64_t)sm_count() * 2;
      forward_scalar_bulk_kernel<<<(unsigned)(ntiles < cap ? ntiles : cap), kThreads, smem, st>>>(
          x.as<const float>(), y.as<float>(), ntiles, qp_dev, d, s, lo, hi);
      FQ_LAUNCH_CHECK("forward_scalar_bulk_kernel");
      return 0;
    }
  }
  return with_code_sink(who, codes, n, [&](auto sink) -> int {
    using Code = decltype(sink);
    if (clip && dependent) {
      ScalarQuant<true, Code> op{d, s, lo, hi, y.as<float>(), sink, QDiv()};
      FQ_CUDA(launch_dependent(forward_scalar_kernel<true, Code, true>, dim3(grid), dim3(kThreads), 0, st,
                               x.as<const float>(), n, per_block, qp_dev, op, (int)vec, (int)reverse));
    } else if (clip) {
      ScalarQuant<true, Code> op{d, s, lo, hi, y.as<float>(), sink, QDiv()};
      forward_scalar_kernel<true, Code><<<grid, kThreads, 0, st>>>(x.as<const float>(), n, per_block, qp_dev, op, vec,
                                                                   reverse);
    } else {
      ScalarQuant<false, Code> op{d, s, lo, hi, y.as<float>(), sink, QDiv()};
      forward_scalar_kernel<false, Code><<<grid, kThreads, 0, st>>>(x.as<const float>(), n, per_block, qp_dev, op, vec,
                                                                    reverse);
    }
    FQ_LAUNCH_CHECK("forward_scalar_kernel");
    return 0;
  });
}

int launch_forward_scalar_dev(const DLTensor* x, const float* qp_dev, const DLTensor* y, const DLTensor* codes,
                              bool reverse, void* stream) {
  // launched right behind the range kernel that produces qp_dev: as its programmatic dependent when the tensor
  // is small enough to be latency-bound (<= 1024 tiles)
  int64_t n = 1;
  for (int i = 0; x != nullptr && i < x->ndim; ++i) n *= x->shape[i];
  return forward_scalar_impl("fq_forward_online", x, qp_dev, 0, 0, 0, 0, true, y, codes, stream, reverse,
                             n <= 1024 * kTileElems);
}

}  // namespace fq

using namespace fq;

extern "C" {

int fq_forward_scalar(const DLTensor* x, const DLTensor* qparams_, const DLTensor* y, const DLTensor* codes,
                      void* stream) {
  View qp;
  FQ_TRY(view_of(qparams_, "fq_forward_scalar: qparams", false, &qp));
  FQ_REQUIRE(qp.is_f32() && qp.numel == 4, "fq_forward_scalar: qparams must be 4 float32 {d, s, lo, hi}");
  return forward_scalar_impl("fq_forward_scalar", x, qp.as<const float>(), 0, 0, 0, 0, true, y, codes, stream);
}

int fq_forward_scalar_host(const DLTensor* x, float d, float s, float lo, float hi, int use_clip, const DLTensor* y,
                           const DLTensor* codes, void* stream) {
  return forward_scalar_impl("fq_forward_scalar_host", x, nullptr, d, s, lo, hi, use_clip != 0, y, codes, stream);
}

int fq_forward_rows(const DLTensor* x_, int64_t rows, const DLTensor* scale_, const DLTensor* y_,
                    const DLTensor* codes_, void* stream) {
  View x, sc, y, codes;
  FQ_TRY(view_of(x_, "fq_forward_rows: x", false, &x));
  FQ_TRY(view_of(scale_, "fq_forward_rows: scale", false, &sc));
  FQ_TRY(view_of(y_, "fq_forward_rows: y", false, &y));
  FQ_TRY(view_of(codes_, "fq_forward_rows: codes", true, &codes));
  FQ_REQUIRE(x.is_f32() && y.is_f32() && sc.is_f32(), "fq_forward_rows: float32 only");
  FQ_REQUIRE(rows >= 1 && sc.numel == rows, "fq_forward_rows: scale must have rows=%lld elements", (long long)rows);
  FQ_REQUIRE(x.numel == y.numel && x.numel % rows == 0, "fq_forward_rows: numel %lld must match y and divide by rows",
             (long long)x.numel);
  if (x.numel == 0) return 0;
  const int64_t n = x.numel, L = n / rows;
  const int vec = aligned16(x.data) && aligned16(y.data) && (codes.null || aligned16(codes.data));
  int64_t per_block = 0;
  const int grid = (vec && L >= 1024) ? slice_grid(n, sm_count() * 8, &per_block) : ew_grid(n);
  cudaStream_t st = (cudaStream_t)stream;
  return with_code_sink("fq_forward_rows", codes, n, [&](auto sink) -> int {
    if (vec && L >= kTileElems) {
      forward_rows_tiles_kernel<<<tile_grid(n, 1 << 30), kThreads, 0, st>>>(x.as<const float>(), n, L, rows,
                                                                          sc.as<const float>(), y.as<float>(), sink);
      FQ_LAUNCH_CHECK("forward_rows_tiles_kernel");
      return 0;
    }
    forward_rows_kernel<<<grid, kThreads, 0, st>>>(x.as<const float>(), n, L, per_block, sc.as<const float>(),
                                                   y.as<float>(), sink, vec);
    FQ_LAUNCH_CHECK("forward_rows_kernel");
    return 0;
  });
}

int fq_ste_backward(const DLTensor* dy_, const DLTensor* x_, const DLTensor* qparams_, const DLTensor* dx_, int mode,
                    void* stream) {
  View dy, x, qp, dx;
  FQ_TRY(view_of(dy_, "fq_ste_backward: dy", false, &dy));
  FQ_TRY(view_of(dx_, "fq_ste_backward: dx", false, &dx));
  FQ_REQUIRE(dy.is_f32() && dx.is_f32() && dy.numel == dx.numel, "fq_ste_backward: dy/dx must be float32 of equal size");
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == FQ_STE_IDENTITY) {   // ste_func.py:43-44; callers alias instead of calling (0 bytes moved)
    if (dx.data != dy.data && dy.numel > 0)
      FQ_CUDA(cudaMemcpyAsync(dx.data, dy.data, sizeof(float) * dy.numel, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  FQ_REQUIRE(mode == FQ_STE_CLIP_MASK, "fq_ste_backward: bad mode %d", mode);
  FQ_TRY(view_of(x_, "fq_ste_backward: x", false, &x));
  FQ_TRY(view_of(qparams_, "fq_ste_backward: qparams", false, &qp));
  FQ_REQUIRE(x.is_f32() && x.numel == dy.numel && qp.is_f32() && qp.numel == 4,
             "fq_ste_backward: x must match dy and qparams must be 4 float32");
  if (dy.numel == 0) return 0;
  const int vec = aligned16(dy.data) && aligned16(x.data) && aligned16(dx.data);
  int64_t per_block = 0;
  const int grid = vec ? tile_grid(dy.numel, 1 << 30) : ew_grid(dy.numel);
  ste_mask_kernel<<<grid, kThreads, 0, st>>>(dy.as<const float>(), x.as<const float>(), qp.as<const float>(),
                                             dx.as<float>(), dy.numel, per_block, vec);
  FQ_LAUNCH_CHECK("ste_mask_kernel");
  return 0;
}

int fq_quantize_int8_export(const DLTensor* w_, const DLTensor* range2_, const DLTensor* out_,
                            const DLTensor* out_range2_, void* stream) {
  View w, rg, out, org;
  FQ_TRY(view_of(w_, "fq_quantize_int8_export: w", false, &w));
  FQ_TRY(view_of(range2_, "fq_quantize_int8_export: range2", false, &rg));
  FQ_TRY(view_of(out_, "fq_quantize_int8_export: out", false, &out));
  FQ_TRY(view_of(out_range2_, "fq_quantize_int8_export: out_range2", true, &org));
  FQ_REQUIRE(w.is_f32() && rg.is_f32() && rg.numel == 2, "fq_quantize_int8_export: w float32, range2 = 2 float32");
  FQ_REQUIRE(out.code == kDLInt && out.bits == 8 && out.numel == w.numel, "fq_quantize_int8_export: out must be int8 like w");
  FQ_REQUIRE(org.null || (org.is_f32() && org.numel == 2), "fq_quantize_int8_export: out_range2 = 2 float32");
  int8_export_kernel<<<ew_grid(w.numel), kThreads, 0, (cudaStream_t)stream>>>(
      w.as<const float>(), w.numel, rg.as<const float>(), out.as<signed char>(), org.null ? nullptr : org.as<float>());
  FQ_LAUNCH_CHECK("int8_export_kernel");
  return 0;
}

int fq_qconv_quantize(const DLTensor* x_, const DLTensor* range2_, const DLTensor* codes_, const DLTensor* scale_out_,
                      void* stream) {
  View x, rg, codes, so;
  FQ_TRY(view_of(x_, "fq_qconv_quantize: x", false, &x));
  FQ_TRY(view_of(range2_, "fq_qconv_quantize: range2", false, &rg));
  FQ_TRY(view_of(codes_, "fq_qconv_quantize: codes", false, &codes));
  FQ_TRY(view_of(scale_out_, "fq_qconv_quantize: scale_out", true, &so));
  FQ_REQUIRE(x.is_f32() && rg.is_f32() && rg.numel == 2, "fq_qconv_quantize: x float32, range2 = 2 float32");
  FQ_REQUIRE(codes.code == kDLInt && codes.bits == 32 && codes.numel == x.numel, "fq_qconv_quantize: codes must be int32 like x");
  FQ_REQUIRE(so.null || (so.is_f32() && so.numel >= 1), "fq_qconv_quantize: scale_out must be float32");
  qconv_quantize_kernel<<<ew_grid(x.numel), kThreads, 0, (cudaStream_t)stream>>>(
      x.as<const float>(), x.numel, rg.as<const float>(), codes.as<int>(), so.null ? nullptr : so.as<float>());
  FQ_LAUNCH_CHECK("qconv_quantize_kernel");
  return 0;
}

int fq_qconv_dequantize(const DLTensor* acc_, const DLTensor* s_in_, const DLTensor* s_w_, const DLTensor* y_,
                        void* stream) {
  View acc, si, sw, y;
  FQ_TRY(view_of(acc_, "fq_qconv_dequantize: acc", false, &acc));
  FQ_TRY(view_of(s_in_, "fq_qconv_dequantize: s_in", false, &si));
  FQ_TRY(view_of(s_w_, "fq_qconv_dequantize: s_w", false, &sw));
  FQ_TRY(view_of(y_, "fq_qconv_dequantize: y", false, &y));
  FQ_REQUIRE(acc.code == kDLInt && acc.bits == 32 && y.is_f32() && acc.numel == y.numel,
             "fq_qconv_dequantize: acc int32 and y float32 of equal size");
  FQ_REQUIRE(si.is_f32() && sw.is_f32() && si.numel >= 1 && sw.numel >= 1, "fq_qconv_dequantize: scales must be float32");
  if (acc.numel == 0) return 0;
  qconv_dequantize_kernel<<<ew_grid(acc.numel), kThreads, 0, (cudaStream_t)stream>>>(
      acc.as<const int>(), acc.numel, si.as<const float>(), sw.as<const float>(), y.as<float>());
  FQ_LAUNCH_CHECK("qconv_dequantize_kernel");
  return 0;
}

}  // extern "C"
