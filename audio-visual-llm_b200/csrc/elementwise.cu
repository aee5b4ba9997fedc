// Small HBM-bound helpers around the projector GEMMs:
//   pack_weight : W_bf16 = bf16(alpha * W_fp32)   (folds fusion_scale of clip_whisper_model.py:434 into the weights)
//   colsum      : db = alpha * sum over flagged rows of dY   (autograd of the nn.Linear bias,
//                 modality_connector.py:32, with the pad-after-projection row mask of
//                 clip_whisper_model.py:340-345)
#include <cuda_fp16.h>

#include "avc_kernels.h"
#include "avc_ptx.cuh"

namespace avc {

namespace {

__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ src, int64_t src_ld,
                                                          uint8_t* __restrict__ dst, int64_t dst_ld,
                                                          int64_t rows, int64_t cols, float alpha) {
  // one thread = 8 consecutive columns of one row (two float4 loads, one 16-byte store)
  const int64_t groups_per_row = cols >> 3;
  const int64_t total = rows * groups_per_row;
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = g / groups_per_row;
    const int64_t c = (g - r * groups_per_row) << 3;
    const float4 x0 = __ldg(reinterpret_cast<const float4*>(src + r * src_ld + c));
    const float4 x1 = __ldg(reinterpret_cast<const float4*>(src + r * src_ld + c + 4));
    int4 o;
    o.x = static_cast<int>(pack_bf16x2(x0.x * alpha, x0.y * alpha));
    o.y = static_cast<int>(pack_bf16x2(x0.z * alpha, x0.w * alpha));
    o.z = static_cast<int>(pack_bf16x2(x1.x * alpha, x1.y * alpha));
    o.w = static_cast<int>(pack_bf16x2(x1.z * alpha, x1.w * alpha));
    *reinterpret_cast<int4*>(dst + (r * dst_ld + c) * 2) = o;
  }
}

constexpr int CS_THREADS = 256;
constexpr int CS_MAX_CHUNKS = 296;   // row slabs (first-level partial sums)
constexpr int CS_GROUP = 16;         // slabs per second-level group
constexpr int CS_MAX_GROUPS = (CS_MAX_CHUNKS + CS_GROUP - 1) / CS_GROUP;
constexpr int CS_MAX_COLGROUPS = 16;  // grid.y: 2048 columns each
constexpr int CS_HEADER_WORDS = 1024;  // arrival counters (zero between launches), then the partial sums
constexpr int CS_TOP_COUNT = CS_MAX_COLGROUPS * 32;  // [y][group] counters first, then [y], then one
constexpr int CS_FINAL_COUNT = CS_TOP_COUNT + CS_MAX_COLGROUPS;
static_assert(CS_MAX_GROUPS <= 32 && CS_FINAL_COUNT < CS_HEADER_WORDS, "colsum header layout");

__device__ __forceinline__ void acc_bf16x8(float (&s)[8], const int4& v) {
  const uint32_t w[4] = {static_cast<uint32_t>(v.x), static_cast<uint32_t>(v.y),
                         static_cast<uint32_t>(v.z), static_cast<uint32_t>(v.w)};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    s[2 * i + 0] += __uint_as_float(w[i] << 16);
    s[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
  }
}

__device__ __forceinline__ void add8(float (&s)[8], const float* p) {
  const float4 x = __ldcg(reinterpret_cast<const float4*>(p));
  const float4 y = __ldcg(reinterpret_cast<const float4*>(p + 4));
  s[0] += x.x; s[1] += x.y; s[2] += x.z; s[3] += x.w;
  s[4] += y.x; s[5] += y.y; s[6] += y.z; s[7] += y.w;
}
__device__ __forceinline__ void store8(float* p, const float (&s)[8], float alpha) {
  *reinterpret_cast<float4*>(p) = make_float4(alpha * s[0], alpha * s[1], alpha * s[2], alpha * s[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(alpha * s[4], alpha * s[5], alpha * s[6], alpha * s[7]);
}

// `n` CTAs arrive on *counter; true (for every thread of the CTA) in the CTA that arrives last, which also puts the
// counter back to zero for the next launch.  The CTA's global writes before the call are visible to the last CTA
// after it.  Uses no shared memory (the kernel must fit next to a GEMM CTA that owns all of it).
__device__ __forceinline__ bool cta_arrive_last(uint32_t* counter, uint32_t n) {
  __threadfence();
  __syncthreads();
  int last = 0;
  if (threadIdx.x == 0) {
    if (atomicAdd(counter, 1u) == n - 1) {
      atomicExch(counter, 0u);
      last = 1;
    }
  }
  last = __syncthreads_or(last);
  if (last) __threadfence();
  return last != 0;
}

// ONE launch, three levels, fixed summation order at every level (deterministic):
//   1. CTA (slab, y) sums its rows for 2048 columns                       -> level-1 partial [slab][which][col]
//   2. the last CTA of each group of 16 slabs adds the group's partials   -> level-2 partial [group][which][col]
//   3. the last of those CTAs adds the <= 19 group sums, scales, writes out0 / out1
//   4. (data parallel) the last of the grid.y finishing CTAs flags the sums ready for the fused all-reduce
// No shared memory and 256 threads per CTA, so that the kernel runs next to the projector GEMM's CTAs.
__global__ void __launch_bounds__(CS_THREADS, 4) colsum_kernel(const __grid_constant__ ColsumArgs a, int rows_per_chunk,
                                                            int nchunks) {
  const int64_t total_rows = static_cast<int64_t>(a.batch) * a.rows;
  const int chunk = blockIdx.x, y = blockIdx.y;
  const int64_t g0 = static_cast<int64_t>(chunk) * rows_per_chunk;
  int64_t g1 = g0 + rows_per_chunk;
  g1 = g1 < total_rows ? g1 : total_rows;
  const int col = (y * CS_THREADS + threadIdx.x) * 8;
  const bool active = col < a.cols;
  uint32_t* hdr = reinterpret_cast<uint32_t*>(a.workspace);
  float* lvl2 = a.workspace + CS_HEADER_WORDS;
  float* lvl1 = lvl2 + static_cast<int64_t>(CS_MAX_GROUPS) * 2 * a.cols;
  if (active) {
    float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    constexpr int U = 8;  // independent 16-byte loads in flight per thread
    // (b, r) of the first row of this slab; advanced incrementally (no per-row division)
    int b = a.rows > 0 ? static_cast<int>(g0 / a.rows) : 0;
    int r = static_cast<int>(g0 - static_cast<int64_t>(b) * a.rows);
    const uint8_t* base = a.dy + static_cast<int64_t>(col) * 2;
    for (int64_t gb = g0; gb < g1; gb += U) {
      int4 v[U];
      bool f0[U], f1[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        f0[u] = f1[u] = false;
        v[u] = make_int4(0, 0, 0, 0);
        if (gb + u < g1) {
          if (a.row_flags != nullptr) {
            const uint8_t fl = __ldg(a.row_flags + gb + u);
            f0[u] = fl & 1; f1[u] = fl & 2;
          } else {
            f0[u] = r < a.flag_rows0; f1[u] = r < a.flag_rows1;
          }
          if (f0[u] || f1[u])
            v[u] = ld_nc_v4(base + b * a.batch_stride + static_cast<int64_t>(r + a.row_base) * a.row_stride);
          if (++r == a.rows) { r = 0; ++b; }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {  // fixed order: deterministic
        if (f0[u]) acc_bf16x8(s0, v[u]);
        if (f1[u]) acc_bf16x8(s1, v[u]);
      }
    }
    store8(lvl1 + (static_cast<int64_t>(chunk) * 2 + 0) * a.cols + col, s0, 1.f);
    store8(lvl1 + (static_cast<int64_t>(chunk) * 2 + 1) * a.cols + col, s1, 1.f);
  }
  const int group = chunk / CS_GROUP;
  const int ngroups = (nchunks + CS_GROUP - 1) / CS_GROUP;
  const int c0 = group * CS_GROUP;
  const int c1 = c0 + CS_GROUP < nchunks ? c0 + CS_GROUP : nchunks;
  if (!cta_arrive_last(hdr + y * 32 + group, static_cast<uint32_t>(c1 - c0))) return;
  if (active) {
    float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int c = c0; c < c1; ++c) {
      add8(s0, lvl1 + (static_cast<int64_t>(c) * 2 + 0) * a.cols + col);
      add8(s1, lvl1 + (static_cast<int64_t>(c) * 2 + 1) * a.cols + col);
    }
    store8(lvl2 + (static_cast<int64_t>(group) * 2 + 0) * a.cols + col, s0, 1.f);
    store8(lvl2 + (static_cast<int64_t>(group) * 2 + 1) * a.cols + col, s1, 1.f);
  }
  if (!cta_arrive_last(hdr + CS_TOP_COUNT + y, static_cast<uint32_t>(ngroups))) return;
  if (active) {
    float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
    for (int g = 0; g < ngroups; ++g) {
      add8(s0, lvl2 + (static_cast<int64_t>(g) * 2 + 0) * a.cols + col);
      add8(s1, lvl2 + (static_cast<int64_t>(g) * 2 + 1) * a.cols + col);
    }
    if (a.out0 != nullptr) store8(a.out0 + col, s0, a.alpha0);
    if (a.out1 != nullptr) store8(a.out1 + col, s1, a.alpha1);
  }
  if (a.sig_owners <= 0) return;
  __threadfence_system();
  if (!cta_arrive_last(hdr + CS_FINAL_COUNT, gridDim.y)) return;
  if (threadIdx.x < static_cast<unsigned>(a.sig_owners)) {
    __threadfence_system();
    st_release_sys(a.sig_flags[threadIdx.x] + COMM_EXTRA_FLAGS + a.sig_rank, a.sig_epoch);
  }
}

// out[b, i, :] = sum_{t in [row_ptr[i], row_ptr[i+1])} weight[t] * x[b, col[t], :]   (fp32 accumulate)
// TG = taps whose loads are issued back to back: 4 for pooling (6 - 7 taps per output row: 76 -> 84 % of the measured
// HBM rate at [32, 1516, 4096] -> 256 rows), 1 for interpolation / the backward of pooling (1 - 2 taps per row, write
// bound: 32 registers keep 16 blocks per SM resident).
template <int DT, int TG>  // DT = GemmOut code of the element type
__global__ void __launch_bounds__(128, TG == 1 ? 16 : 12) row_resample_kernel(const __grid_constant__ ResampleArgs a) {
  constexpr bool F32 = DT == GEMM_OUT_F32;
  const int i = blockIdx.x, b = blockIdx.y;
  const int t0 = __ldg(a.row_ptr + i), t1 = __ldg(a.row_ptr + i + 1);
  constexpr int EPV = F32 ? 4 : 8;  // elements per 16-byte vector
  const int64_t row_bytes = static_cast<int64_t>(a.hidden) * (F32 ? 4 : 2);
  const uint8_t* xb = a.x + static_cast<int64_t>(b) * a.src_rows * row_bytes;
  uint8_t* ob = a.out + (static_cast<int64_t>(b) * a.dst_rows + i) * row_bytes;
  auto fma_vec = [](float (&acc)[8], float w, const int4& q) {
    if (F32) {
      acc[0] = fmaf(w, __int_as_float(q.x), acc[0]);
      acc[1] = fmaf(w, __int_as_float(q.y), acc[1]);
      acc[2] = fmaf(w, __int_as_float(q.z), acc[2]);
      acc[3] = fmaf(w, __int_as_float(q.w), acc[3]);
    } else {
      const uint32_t u[4] = {static_cast<uint32_t>(q.x), static_cast<uint32_t>(q.y),
                             static_cast<uint32_t>(q.z), static_cast<uint32_t>(q.w)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float lo, hi;
        if (DT == GEMM_OUT_F16) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&u[e]));
          lo = f.x; hi = f.y;
        } else {
          lo = __uint_as_float(u[e] << 16); hi = __uint_as_float(u[e] & 0xffff0000u);
        }
        acc[2 * e + 0] = fmaf(w, lo, acc[2 * e + 0]);
        acc[2 * e + 1] = fmaf(w, hi, acc[2 * e + 1]);
      }
    }
  };
  for (int v = threadIdx.x; v < a.hidden / EPV; v += blockDim.x) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const uint8_t* xv = xb + static_cast<int64_t>(v) * 16;
    // taps in groups of four: the loads of a group are issued back to back (the tap order of the sum is unchanged)
    if (TG == 1) {
      for (int t = t0; t < t1; ++t)
        fma_vec(acc, __ldg(a.weight + t), ld_nc_v4(xv + static_cast<int64_t>(__ldg(a.col + t)) * row_bytes));
    } else
    for (int t = t0; t < t1; t += TG) {
      int4 q[TG];
      float w[TG];
#pragma unroll
      for (int j = 0; j < TG; ++j) {
        const bool on = t + j < t1;
        w[j] = on ? __ldg(a.weight + t + j) : 0.f;
        q[j] = on ? ld_nc_v4(xv + static_cast<int64_t>(__ldg(a.col + t + j)) * row_bytes) : make_int4(0, 0, 0, 0);
      }
#pragma unroll
      for (int j = 0; j < TG; ++j)
        if (t + j < t1) fma_vec(acc, w[j], q[j]);
    }
    int4 o;
    if (F32) {
      o = make_int4(__float_as_int(acc[0]), __float_as_int(acc[1]), __float_as_int(acc[2]), __float_as_int(acc[3]));
    } else if (DT == GEMM_OUT_F16) {
      o = make_int4(static_cast<int>(pack_f16x2(acc[0], acc[1])), static_cast<int>(pack_f16x2(acc[2], acc[3])),
                    static_cast<int>(pack_f16x2(acc[4], acc[5])), static_cast<int>(pack_f16x2(acc[6], acc[7])));
    } else {
      o = make_int4(static_cast<int>(pack_bf16x2(acc[0], acc[1])), static_cast<int>(pack_bf16x2(acc[2], acc[3])),
                    static_cast<int>(pack_bf16x2(acc[4], acc[5])), static_cast<int>(pack_bf16x2(acc[6], acc[7])));
    }
    st_na_v4(ob + static_cast<int64_t>(v) * 16, o);
  }
}

// ---------------------------------------------------------------------------------------------- MLP projector
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// FWD: out = mask(row) * gelu(z)            BWD: out = mask(row) * dh * gelu'(z)      (bf16, 8 elements per thread)
template <bool BWD>
__global__ void __launch_bounds__(256) gelu_kernel(const __grid_constant__ GeluArgs a) {
  const int64_t vec_per_row = a.cols >> 3;
  const int64_t total = a.rows * vec_per_row;
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = g / vec_per_row;
    const int64_t c = (g - r * vec_per_row) << 3;
    const bool on = a.row_flags == nullptr || (__ldg(a.row_flags + r) & a.flag_bit) != 0;
    int4 o = make_int4(0, 0, 0, 0);
    if (on) {
      const int4 zq = ld_nc_v4(a.z + (r * a.z_ld + c) * 2);
      const uint32_t zw[4] = {static_cast<uint32_t>(zq.x), static_cast<uint32_t>(zq.y), static_cast<uint32_t>(zq.z),
                              static_cast<uint32_t>(zq.w)};
      uint32_t ow[4];
      if (BWD) {
        const int4 dq = ld_nc_v4(a.dh + (r * a.dh_ld + c) * 2);
        const uint32_t dw[4] = {static_cast<uint32_t>(dq.x), static_cast<uint32_t>(dq.y),
                                static_cast<uint32_t>(dq.z), static_cast<uint32_t>(dq.w)};
#pragma unroll
        for (int e = 0; e < 4; ++e)
          ow[e] = pack_bf16x2(bf_lo(dw[e]) * gelu_grad_f(bf_lo(zw[e])), bf_hi(dw[e]) * gelu_grad_f(bf_hi(zw[e])));
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) ow[e] = pack_bf16x2(gelu_f(bf_lo(zw[e])), gelu_f(bf_hi(zw[e])));
      }
      o = make_int4(static_cast<int>(ow[0]), static_cast<int>(ow[1]), static_cast<int>(ow[2]), static_cast<int>(ow[3]));
    }
    st_na_v4(a.out + (r * a.out_ld + c) * 2, o);
  }
}

// dst_bf16[c, r] = bf16(alpha * src_f32[r, c]): transposed weight pack for the input-gradient GEMM (dX = dY . W)
__global__ void __launch_bounds__(256) pack_weight_t_kernel(const float* __restrict__ src, int64_t src_ld,
                                                            uint8_t* __restrict__ dst, int64_t dst_ld, int64_t rows,
                                                            int64_t cols, float alpha) {
  __shared__ float tile[32][33];
  const int64_t r0 = static_cast<int64_t>(blockIdx.y) * 32, c0 = static_cast<int64_t>(blockIdx.x) * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int64_t r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < rows && c < cols) ? __ldg(src + r * src_ld + c) * alpha : 0.f;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int64_t c = c0 + i, r = r0 + tx;  // output row = source column
    if (c < cols && r < rows) {
      const float x = tile[tx][i];
      const uint32_t b = pack_bf16x2(x, 0.f);
      reinterpret_cast<uint16_t*>(dst)[c * dst_ld + r] = static_cast<uint16_t>(b & 0xffffu);
    }
  }
}


// dst_bf16[r, c] = bf16(alpha * src[r, c]) for fp32 / fp16 / bf16 sources: the cast of the connector input to the
// compute dtype (modality_connector.py:18-19) and of an fp16 / fp32 upstream gradient to the dW GEMM's operand type.
// One thread = 8 consecutive columns (16-byte store).
template <int SRC>  // GemmOut code of the source element type
__global__ void __launch_bounds__(256) cast_bf16_kernel(const uint8_t* __restrict__ src, int64_t src_ld,
                                                        uint8_t* __restrict__ dst, int64_t dst_ld, int64_t rows,
                                                        int64_t cols, float alpha) {
  const int64_t groups_per_row = cols >> 3;
  const int64_t total = rows * groups_per_row;
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = g / groups_per_row;
    const int64_t c = (g - r * groups_per_row) << 3;
    float x[8];
    if (SRC == GEMM_OUT_F32) {
      const float* sp = reinterpret_cast<const float*>(src) + r * src_ld + c;
      const int4 a = ld_nc_v4(sp), b = ld_nc_v4(sp + 4);
      x[0] = __int_as_float(a.x); x[1] = __int_as_float(a.y); x[2] = __int_as_float(a.z); x[3] = __int_as_float(a.w);
      x[4] = __int_as_float(b.x); x[5] = __int_as_float(b.y); x[6] = __int_as_float(b.z); x[7] = __int_as_float(b.w);
    } else {
      const int4 q = ld_nc_v4(src + (r * src_ld + c) * 2);
      const uint32_t w[4] = {static_cast<uint32_t>(q.x), static_cast<uint32_t>(q.y), static_cast<uint32_t>(q.z),
                             static_cast<uint32_t>(q.w)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (SRC == GEMM_OUT_F16) {
          const __half2 h = *reinterpret_cast<const __half2*>(&w[e]);
          const float2 f = __half22float2(h);
          x[2 * e] = f.x; x[2 * e + 1] = f.y;
        } else {
          x[2 * e] = __uint_as_float(w[e] << 16); x[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
        }
      }
    }
    int4 o;
    o.x = static_cast<int>(pack_bf16x2(x[0] * alpha, x[1] * alpha));
    o.y = static_cast<int>(pack_bf16x2(x[2] * alpha, x[3] * alpha));
    o.z = static_cast<int>(pack_bf16x2(x[4] * alpha, x[5] * alpha));
    o.w = static_cast<int>(pack_bf16x2(x[6] * alpha, x[7] * alpha));
    st_na_v4(dst + (r * dst_ld + c) * 2, o);
  }
}

// Input gradient of the gather (align + stack + concat): every frame (b, t) of one stream receives the sum of the
// column slices of dA rows that read it -- token j reads the stack j / rep, i.e. frames k * (j / rep) .. + k - 1 --
// and zero when no token does (t past the valid length or past the token cap).  One warp per frame, 16-byte vectors.
template <bool OUT_F32>
__global__ void __launch_bounds__(256) gather_bwd_kernel(const __grid_constant__ GatherBwdArgs a) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = static_cast<int64_t>(gridDim.x) * (blockDim.x >> 5);
  const int64_t total = static_cast<int64_t>(a.batch) * a.frames;
  const int nvec = a.dim >> 3;  // 8 bf16 per 16-byte source vector
  for (int64_t f = static_cast<int64_t>(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); f < total; f += warps) {
    const int b = static_cast<int>(f / a.frames);
    const int t = static_cast<int>(f - static_cast<int64_t>(b) * a.frames);
    int len = a.frames;
    if (a.valid != nullptr) {
      const int l = __ldg(a.valid + b);
      len = l < len ? (l < 0 ? 0 : l) : len;
    }
    int64_t row0;
    int ntok;
    if (a.tok_offset != nullptr) {
      row0 = __ldg(a.tok_offset + b);
      ntok = __ldg(a.tok_offset + b + 1) - static_cast<int>(row0);
    } else {
      row0 = static_cast<int64_t>(b) * a.tokens_per_sample;
      ntok = a.tokens_per_sample;
    }
    const int s = t / a.k;                   // stack index
    const int slot = t - s * a.k;            // position inside the stack
    int j0 = s * a.rep, j1 = j0 + a.rep;     // tokens that read this stack
    j1 = j1 < ntok ? j1 : ntok;
    if (t >= len) j1 = j0;                   // zero-padded in the forward: no gradient
    uint8_t* out = a.dst + b * a.batch_stride + t * a.frame_stride;
    for (int v = lane; v < nvec; v += 32) {
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int j = j0; j < j1; ++j) {
        const int4 q = ld_nc_v4(a.da + ((row0 + j) * a.a_row_stride + a.col_off + static_cast<int64_t>(slot) * a.dim +
                                        static_cast<int64_t>(v) * 8) * 2);
        const uint32_t w[4] = {static_cast<uint32_t>(q.x), static_cast<uint32_t>(q.y), static_cast<uint32_t>(q.z),
                               static_cast<uint32_t>(q.w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += __uint_as_float(w[e] << 16);
          acc[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
        }
      }
      if (OUT_F32) {
        float* o = reinterpret_cast<float*>(out) + static_cast<int64_t>(v) * 8;
        st_na_v4(o, make_int4(__float_as_int(acc[0]), __float_as_int(acc[1]), __float_as_int(acc[2]),
                              __float_as_int(acc[3])));
        st_na_v4(o + 4, make_int4(__float_as_int(acc[4]), __float_as_int(acc[5]), __float_as_int(acc[6]),
                                  __float_as_int(acc[7])));
      } else {
        st_na_v4(out + static_cast<int64_t>(v) * 16,
                 make_int4(static_cast<int>(pack_bf16x2(acc[0], acc[1])), static_cast<int>(pack_bf16x2(acc[2], acc[3])),
                           static_cast<int>(pack_bf16x2(acc[4], acc[5])), static_cast<int>(pack_bf16x2(acc[6], acc[7]))));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- trainer step
constexpr int SQ_BLOCKS = 592;  // 4 x 148
constexpr int SQ_THREADS = 256;

// deterministic sum of squares: fixed grid, fixed per-thread stride order, fixed tree
__global__ void __launch_bounds__(SQ_THREADS) sumsq_partial_kernel(const float* __restrict__ x, int64_t n,
                                                                    float* __restrict__ partial) {
  __shared__ float red[SQ_THREADS / 32];
  float s = 0.f;
  const int64_t nv = n >> 2;
  const float4* xv = reinterpret_cast<const float4*>(x);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(SQ_THREADS) + threadIdx.x; i < nv;
       i += static_cast<int64_t>(SQ_BLOCKS) * SQ_THREADS) {
    const float4 v = __ldg(xv + i);
    s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); s = fmaf(v.w, v.w, s);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const float t = x[(nv << 2) + threadIdx.x];
    s = fmaf(t, t, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < SQ_THREADS / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(32) sumsq_final_kernel(const float* __restrict__ partial, float* __restrict__ out,
                                                         int accumulate) {
  // one warp, fixed order: lane l adds partial[l], partial[l + 32], ... then a fixed shuffle tree
  float s = 0.f;
  for (int i = threadIdx.x; i < SQ_BLOCKS; i += 32) s += partial[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) *out = accumulate ? (*out + s) : s;
}

// AdamW (torch.optim.AdamW semantics, decoupled weight decay) on a [rows, cols] fp32 parameter, with the gradient
// pre-scaled by a device scalar (global-norm clip coefficient) and an optional bf16 copy alpha * p written into the
// packed projector operand for the next forward.
__global__ void __launch_bounds__(256) adamw_kernel(const __grid_constant__ AdamWArgs a) {
  const int64_t groups_per_row = a.cols >> 2;
  const int64_t total = a.rows * groups_per_row;
  float gs = a.grad_scale != nullptr ? __ldg(a.grad_scale) : 1.f;
  if (a.clip_sumsq != nullptr && a.max_norm > 0.f)  // torch.nn.utils.clip_grad_norm_: coef = min(1, max_norm / (norm + 1e-6))
    gs *= fminf(1.f, a.max_norm / (sqrtf(__ldg(a.clip_sumsq)) + 1e-6f));
  for (int64_t g = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; g < total;
       g += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = g / groups_per_row;
    const int64_t c = (g - r * groups_per_row) << 2;
    const int64_t i = r * a.cols + c;
    float4 p = *reinterpret_cast<const float4*>(a.param + i);
    const float4 gr = __ldg(reinterpret_cast<const float4*>(a.grad + i));
    float4 m = *reinterpret_cast<const float4*>(a.exp_avg + i);
    float4 v = *reinterpret_cast<const float4*>(a.exp_avg_sq + i);
    float* pp = &p.x; const float* gp = &gr.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float gg = gp[e] * gs;
      pp[e] = pp[e] * a.decay;                                   // p *= 1 - lr * weight_decay
      mp[e] = mp[e] + (gg - mp[e]) * a.one_minus_beta1;          // lerp, as torch does
      vp[e] = vp[e] * a.beta2 + gg * gg * a.one_minus_beta2;
      const float denom = sqrtf(vp[e]) / a.bias_correction2_sqrt + a.eps;
      pp[e] = pp[e] - a.step_size * (mp[e] / denom);
    }
    *reinterpret_cast<float4*>(a.param + i) = p;
    *reinterpret_cast<float4*>(a.exp_avg + i) = m;
    *reinterpret_cast<float4*>(a.exp_avg_sq + i) = v;
    if (a.packed != nullptr) {
      uint2 o;
      o.x = pack_bf16x2(p.x * a.packed_alpha, p.y * a.packed_alpha);
      o.y = pack_bf16x2(p.z * a.packed_alpha, p.w * a.packed_alpha);
      *reinterpret_cast<uint2*>(a.packed + (r * a.packed_ld + c) * 2) = o;
    }
  }
}

}  // namespace

cudaError_t launch_gelu(const GeluArgs& a, bool backward, cudaStream_t stream) {
  if (a.rows <= 0 || a.cols <= 0) return cudaSuccess;
  if (a.cols % 8 != 0 || a.z_ld % 8 != 0 || a.out_ld % 8 != 0 || (backward && a.dh_ld % 8 != 0) ||
      (reinterpret_cast<uintptr_t>(a.z) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.out) & 15) != 0 ||
      (backward && (reinterpret_cast<uintptr_t>(a.dh) & 15) != 0))
    return cudaErrorMisalignedAddress;
  const int64_t total = a.rows * (a.cols >> 3);
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  if (backward) gelu_kernel<true><<<static_cast<int>(grid), 256, 0, stream>>>(a);
  else gelu_kernel<false><<<static_cast<int>(grid), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_pack_weight_t(const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows,
                                 int64_t cols, float alpha, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (rows > 32 * 65535LL) return cudaErrorInvalidValue;
  dim3 grid(static_cast<unsigned>((cols + 31) / 32), static_cast<unsigned>((rows + 31) / 32));
  pack_weight_t_kernel<<<grid, 256, 0, stream>>>(src, src_ld, static_cast<uint8_t*>(dst), dst_ld, rows, cols, alpha);
  return cudaGetLastError();
}

size_t sumsq_workspace_bytes() { return SQ_BLOCKS * sizeof(float); }

cudaError_t launch_sumsq(const float* x, int64_t n, float* out, float* workspace, int accumulate,
                         cudaStream_t stream) {
  if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) return cudaErrorMisalignedAddress;
  sumsq_partial_kernel<<<SQ_BLOCKS, SQ_THREADS, 0, stream>>>(x, n, workspace);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sumsq_final_kernel<<<1, 32, 0, stream>>>(workspace, out, accumulate);
  return cudaGetLastError();
}

cudaError_t launch_adamw(const AdamWArgs& a, cudaStream_t stream) {
  if (a.rows <= 0 || a.cols <= 0) return cudaSuccess;
  if (a.cols % 4 != 0 || (reinterpret_cast<uintptr_t>(a.param) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(a.grad) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.exp_avg) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(a.exp_avg_sq) & 15) != 0)
    return cudaErrorMisalignedAddress;
  if (a.packed != nullptr && ((reinterpret_cast<uintptr_t>(a.packed) & 7) != 0 || a.packed_ld % 4 != 0))
    return cudaErrorMisalignedAddress;
  const int64_t total = a.rows * (a.cols >> 2);
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  adamw_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_row_resample(const ResampleArgs& a, cudaStream_t stream) {
  if (a.batch <= 0 || a.dst_rows <= 0) return cudaSuccess;
  const int epv = a.dtype == GEMM_OUT_F32 ? 4 : 8;
  if (a.hidden % epv != 0 || (reinterpret_cast<uintptr_t>(a.x) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(a.out) & 15) != 0)
    return cudaErrorMisalignedAddress;
  if (a.batch > 65535) return cudaErrorInvalidValue;
  dim3 grid(a.dst_rows, a.batch);
  if (a.src_rows > 2 * a.dst_rows) {  // pooling forward: several taps per output row
    if (a.dtype == GEMM_OUT_F32) row_resample_kernel<GEMM_OUT_F32, 4><<<grid, 128, 0, stream>>>(a);
    else if (a.dtype == GEMM_OUT_F16) row_resample_kernel<GEMM_OUT_F16, 4><<<grid, 128, 0, stream>>>(a);
    else row_resample_kernel<GEMM_OUT_BF16, 4><<<grid, 128, 0, stream>>>(a);
  } else {
    if (a.dtype == GEMM_OUT_F32) row_resample_kernel<GEMM_OUT_F32, 1><<<grid, 128, 0, stream>>>(a);
    else if (a.dtype == GEMM_OUT_F16) row_resample_kernel<GEMM_OUT_F16, 1><<<grid, 128, 0, stream>>>(a);
    else row_resample_kernel<GEMM_OUT_BF16, 1><<<grid, 128, 0, stream>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_pack_weight(const float* src, int64_t src_ld, void* dst, int64_t dst_ld,
                               int64_t rows, int64_t cols, float alpha, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  if (cols % 8 != 0 || src_ld % 4 != 0 || dst_ld % 8 != 0 ||
      (reinterpret_cast<uintptr_t>(src) & 15) != 0 || (reinterpret_cast<uintptr_t>(dst) & 15) != 0)
    return cudaErrorMisalignedAddress;
  const int64_t total = rows * (cols >> 3);
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  pack_weight_kernel<<<static_cast<int>(grid), 256, 0, stream>>>(
      src, src_ld, static_cast<uint8_t*>(dst), dst_ld, rows, cols, alpha);
  return cudaGetLastError();
}


cudaError_t launch_cast_bf16(const void* src, int src_dtype, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows,
                             int64_t cols, float alpha, cudaStream_t stream) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  const int es = src_dtype == GEMM_OUT_F32 ? 4 : 2;
  if (cols % 8 != 0 || (src_ld * es) % 16 != 0 || dst_ld % 8 != 0 || (reinterpret_cast<uintptr_t>(src) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(dst) & 15) != 0)
    return cudaErrorMisalignedAddress;
  const int64_t total = rows * (cols >> 3);
  int64_t grid = (total + 255) / 256;
  if (grid > 148 * 16) grid = 148 * 16;
  const uint8_t* s8 = static_cast<const uint8_t*>(src);
  uint8_t* d8 = static_cast<uint8_t*>(dst);
  switch (src_dtype) {
    case GEMM_OUT_F32:
      cast_bf16_kernel<GEMM_OUT_F32><<<static_cast<int>(grid), 256, 0, stream>>>(s8, src_ld, d8, dst_ld, rows, cols, alpha);
      break;
    case GEMM_OUT_F16:
      cast_bf16_kernel<GEMM_OUT_F16><<<static_cast<int>(grid), 256, 0, stream>>>(s8, src_ld, d8, dst_ld, rows, cols, alpha);
      break;
    case GEMM_OUT_BF16:
      cast_bf16_kernel<GEMM_OUT_BF16><<<static_cast<int>(grid), 256, 0, stream>>>(s8, src_ld, d8, dst_ld, rows, cols, alpha);
      break;
    default:
      return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_gather_bwd(const GatherBwdArgs& a, bool out_f32, cudaStream_t stream) {
  if (a.batch <= 0 || a.frames <= 0) return cudaSuccess;
  const int es = out_f32 ? 4 : 2;
  if (a.dim % 8 != 0 || a.k < 1 || a.rep < 1 || (reinterpret_cast<uintptr_t>(a.da) & 15) != 0 ||
      (reinterpret_cast<uintptr_t>(a.dst) & 15) != 0 || a.a_row_stride % 8 != 0 || a.col_off % 8 != 0 ||
      a.batch_stride % 16 != 0 || a.frame_stride % 16 != 0 || a.frame_stride < static_cast<int64_t>(a.dim) * es)
    return cudaErrorMisalignedAddress;
  const int64_t total = static_cast<int64_t>(a.batch) * a.frames;
  int64_t grid = (total + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  if (out_f32) gather_bwd_kernel<true><<<static_cast<int>(grid), 256, 0, stream>>>(a);
  else gather_bwd_kernel<false><<<static_cast<int>(grid), 256, 0, stream>>>(a);
  return cudaGetLastError();
}

size_t colsum_workspace_bytes(int cols) {
  return (static_cast<size_t>(CS_HEADER_WORDS) +
          static_cast<size_t>(CS_MAX_CHUNKS + CS_MAX_GROUPS) * 2 * static_cast<size_t>(cols)) * sizeof(float);
}
size_t colsum_workspace_header_bytes() { return static_cast<size_t>(CS_HEADER_WORDS) * sizeof(float); }

// CUDA loads kernels lazily, and loading one can wait for running kernels to finish.  colsum_kernel is launched
// while the fused dW + all-reduce GEMM is already spinning on its result, so it must be resident before that.
cudaError_t preload_colsum() {
  cudaFuncAttributes attr;
  return cudaFuncGetAttributes(&attr, colsum_kernel);
}

cudaError_t launch_colsum(const ColsumArgs& a, cudaStream_t stream) {
  if (a.cols <= 0) return cudaSuccess;
  if (a.cols % 8 != 0 || a.row_stride % 16 != 0 || a.batch_stride % 16 != 0 ||
      (reinterpret_cast<uintptr_t>(a.dy) & 15) != 0 || a.workspace == nullptr ||
      (reinterpret_cast<uintptr_t>(a.workspace) & 15) != 0 ||
      (a.out0 != nullptr && (reinterpret_cast<uintptr_t>(a.out0) & 15) != 0) ||
      (a.out1 != nullptr && (reinterpret_cast<uintptr_t>(a.out1) & 15) != 0))
    return cudaErrorMisalignedAddress;
  const int64_t total_rows = static_cast<int64_t>(a.batch) * a.rows;
  int nchunks = static_cast<int>(total_rows < CS_MAX_CHUNKS ? (total_rows > 0 ? total_rows : 1) : CS_MAX_CHUNKS);
  const int rows_per_chunk = static_cast<int>((total_rows + nchunks - 1) / (nchunks > 0 ? nchunks : 1));
  if (rows_per_chunk > 0) nchunks = static_cast<int>((total_rows + rows_per_chunk - 1) / rows_per_chunk);
  if (nchunks < 1) nchunks = 1;
  dim3 grid(nchunks, (a.cols / 8 + CS_THREADS - 1) / CS_THREADS);
  if (grid.y > CS_MAX_COLGROUPS) return cudaErrorInvalidValue;
  colsum_kernel<<<grid, CS_THREADS, 0, stream>>>(a, rows_per_chunk > 0 ? rows_per_chunk : 1, nchunks);
  return cudaGetLastError();
}

}  // namespace avc
