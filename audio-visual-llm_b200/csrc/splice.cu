// Splice: write projected AV rows into the LLM input-embedding sequence at the placeholder
// positions, text-embedding rows everywhere else, and emit the int64 attention mask and the
// int64 -100 label mask.  Backward gathers d(inputs_embeds) rows back into packed dY.
//
// Reference semantics covered (parity mode): `[prompt | AV]` concat + embedding lookup
// (clip_whisper_model.py:448-451, 464-487), all-ones int64 mask (:460), pad -> -100 and
// truncate / right-pad -100 of the labels (:569-570, :586-598).
//
// Grid (ceil(S/32), B).  Each CTA: (1) counts the placeholders left of its 32-position chunk,
// (2) warp 0 turns the chunk's placeholder ballot into packed-row indices with a popc prefix and
// writes the masks, (3) all 8 warps move rows with 128-bit loads/stores, 8 in flight per lane.
#include "avc_kernels.h"
#include "avc_ptx.cuh"

namespace avc {

namespace {

constexpr int SP_THREADS = 256;
constexpr int SP_CHUNK = 32;
constexpr int SP_UNROLL = 8;

__device__ __forceinline__ void copy_row(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                         int row_bytes, int lane) {
  const int nvec = row_bytes >> 4;
  int i = lane;
  for (; i + (SP_UNROLL - 1) * 32 < nvec; i += SP_UNROLL * 32) {
    int4 v[SP_UNROLL];
#pragma unroll
    for (int u = 0; u < SP_UNROLL; ++u) v[u] = ld_nc_v4(src + (static_cast<int64_t>(i + u * 32) << 4));
#pragma unroll
    for (int u = 0; u < SP_UNROLL; ++u) st_na_v4(dst + (static_cast<int64_t>(i + u * 32) << 4), v[u]);
  }
  for (; i < nvec; i += 32) st_na_v4(dst + (static_cast<int64_t>(i) << 4), ld_nc_v4(src + (static_cast<int64_t>(i) << 4)));
}

__device__ __forceinline__ void zero_row(uint8_t* __restrict__ dst, int row_bytes, int lane) {
  const int nvec = row_bytes >> 4;
  const int4 z = make_int4(0, 0, 0, 0);
  for (int i = lane; i < nvec; i += 32) st_na_v4(dst + (static_cast<int64_t>(i) << 4), z);
}

template <bool FWD>
__global__ void __launch_bounds__(SP_THREADS) splice_kernel(const __grid_constant__ SpliceArgs a) {
  __shared__ int s_warp_count[SP_THREADS / 32];
  __shared__ const uint8_t* s_src[SP_CHUNK];
  __shared__ uint8_t* s_dst[SP_CHUNK];

  const int b = blockIdx.y;
  const int p0 = blockIdx.x * SP_CHUNK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t* ids = a.input_ids + static_cast<int64_t>(b) * a.seq;

  // (1) placeholders strictly left of this chunk
  int cnt = 0;
  for (int p = threadIdx.x; p < p0; p += SP_THREADS) cnt += (__ldg(ids + p) == a.placeholder_id) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_warp_count[warp] = cnt;
  __syncthreads();

  // (2) packed-row index per position of the chunk + masks
  if (warp == 0) {
    int base = 0;
#pragma unroll
    for (int w = 0; w < SP_THREADS / 32; ++w) base += s_warp_count[w];
    int64_t row_off;
    int ntok;
    if (a.tok_offset != nullptr) {
      row_off = __ldg(a.tok_offset + b);
      ntok = __ldg(a.tok_offset + b + 1) - static_cast<int>(row_off);
    } else {
      row_off = static_cast<int64_t>(b) * a.tokens_per_sample;
      ntok = a.tokens_per_sample;
    }
    const int p = p0 + lane;
    const bool in = p < a.seq;
    const int64_t id = in ? __ldg(ids + p) : a.pad_id;
    const bool is_ph = in && id == a.placeholder_id;
    const uint32_t ball = __ballot_sync(0xffffffffu, is_ph);
    const int rank = base + __popc(ball & ((1u << lane) - 1u));
    const bool has_row = is_ph && rank < ntok;
    uint8_t* emb_row = a.inputs_embeds + (static_cast<int64_t>(b) * a.seq + p) * a.row_bytes;
    if (FWD) {
      const uint8_t* src = nullptr;
      if (has_row) {
        src = a.y + (row_off + rank) * a.row_bytes;
      } else if (in && !is_ph && a.embed_table != nullptr && id >= 0 && id < a.vocab) {
        src = a.embed_table + id * a.row_bytes;
      }
      s_src[lane] = src;
      s_dst[lane] = (in && !(a.av_in_place && has_row)) ? emb_row : nullptr;
      if (in) {
        if (a.attention_mask != nullptr) {
          int64_t mval = 1;
          if (a.mask_mode == 1) mval = is_ph ? (has_row ? 1 : 0) : (id != a.pad_id ? 1 : 0);
          a.attention_mask[static_cast<int64_t>(b) * a.seq + p] = mval;
        }
        if (a.labels_out != nullptr) {
          int64_t lv = -100;
          if (a.labels_in != nullptr && p < a.label_len)
            lv = __ldg(a.labels_in + static_cast<int64_t>(b) * a.label_len + p);
          else if (a.labels_in == nullptr && a.label_mode == 1)
            lv = id;
          if (lv == a.pad_id) lv = -100;
          if (a.label_mode == 1 && (is_ph || id == a.pad_id)) lv = -100;
          a.labels_out[static_cast<int64_t>(b) * a.seq + p] = lv;
        }
      }
    } else {
      s_src[lane] = has_row ? emb_row : nullptr;
      s_dst[lane] = has_row ? a.dy + (row_off + rank) * a.row_bytes : nullptr;
    }
    // placeholder count must equal the sample's token count
    if (a.status != nullptr && p0 + SP_CHUNK >= a.seq) {
      const int total = base + __popc(ball);
      if (lane == 0 && total != ntok) atomicOr(a.status, 1);
    }
  }
  __syncthreads();

  // (3) move rows
  for (int r = warp; r < SP_CHUNK; r += SP_THREADS / 32) {
    uint8_t* dst = s_dst[r];
    const uint8_t* src = s_src[r];
    if (dst == nullptr) continue;
    if (src != nullptr) copy_row(dst, src, a.row_bytes, lane);
    else if (FWD) zero_row(dst, a.row_bytes, lane);
  }
}

}  // namespace

static cudaError_t check_splice(const SpliceArgs& a) {
  if (a.row_bytes % 16 != 0) return cudaErrorMisalignedAddress;
  if ((reinterpret_cast<uintptr_t>(a.inputs_embeds) & 15) != 0) return cudaErrorMisalignedAddress;
  if (a.batch <= 0 || a.seq <= 0) return cudaErrorInvalidValue;
  if (a.batch > 65535) return cudaErrorInvalidValue;
  return cudaSuccess;
}

cudaError_t launch_splice_fwd(const SpliceArgs& a, int, cudaStream_t stream) {
  cudaError_t e = check_splice(a);
  if (e != cudaSuccess) return e;
  dim3 grid((a.seq + SP_CHUNK - 1) / SP_CHUNK, a.batch);
  splice_kernel<true><<<grid, SP_THREADS, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_splice_bwd(const SpliceArgs& a, int, cudaStream_t stream) {
  cudaError_t e = check_splice(a);
  if (e != cudaSuccess) return e;
  dim3 grid((a.seq + SP_CHUNK - 1) / SP_CHUNK, a.batch);
  splice_kernel<false><<<grid, SP_THREADS, 0, stream>>>(a);
  return cudaGetLastError();
}

}  // namespace avc
