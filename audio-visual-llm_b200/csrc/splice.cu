// Splice: write projected AV rows into the LLM input-embedding sequence at the placeholder
// positions, text-embedding rows everywhere else, and emit the int64 attention mask and the
// int64 -100 label mask.  Backward gathers d(inputs_embeds) rows back into packed dY.
//
// Reference semantics covered (parity mode): `[prompt | AV]` concat + embedding lookup
// (clip_whisper_model.py:448-451, 464-487), all-ones int64 mask (:460), pad -> -100 and
// truncate / right-pad -100 of the labels (:569-570, :586-598).
//
// Persistent CTAs over (sample, 32-position chunk) units.  Per unit: (1) count the placeholders left of the chunk,
// (2) warp 0 turns the chunk's placeholder ballot into packed-row indices with a popc prefix and writes the masks,
// (3) every warp moves its rows with TMA bulk copies (cp.async.bulk global -> smem ring -> global, one elected lane,
// SP_LOOKAHEAD loads in flight per warp): the row data never touches registers.  Zero rows come from a zeroed smem
// block.  Stores stay in flight across units; only the smem-stage reuse is waited for.
#include "avc_kernels.h"
#include "avc_ptx.cuh"

namespace avc {

namespace {

constexpr int SP_WARPS = 4;
constexpr int SP_THREADS = SP_WARPS * 32;
constexpr int SP_CHUNK = 32;
constexpr int SP_STAGES = 4;
constexpr int SP_LOOKAHEAD = 3;   // loads in flight per warp; stage reuse distance = SP_STAGES - SP_LOOKAHEAD stores
constexpr int SP_MAX_PIECE = 4096;  // bytes per bulk copy (rows are cut into pieces of at most this size)

template <bool FWD>
__global__ void __launch_bounds__(SP_THREADS) splice_kernel(const __grid_constant__ SpliceArgs a, int piece_bytes,
                                                            int units_x) {
  extern __shared__ __align__(128) uint8_t sp_smem[];
  __shared__ int s_warp_count[SP_WARPS];
  __shared__ const uint8_t* s_src[SP_CHUNK];
  __shared__ uint8_t* s_dst[SP_CHUNK];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // dynamic smem: [zero block][SP_WARPS rings of SP_STAGES pieces][SP_WARPS x SP_STAGES barriers]
  const uint32_t s_zero = smem_u32(sp_smem);
  const uint32_t s_ring = s_zero + piece_bytes + warp * (SP_STAGES * piece_bytes);
  const uint32_t s_bar = s_zero + piece_bytes + SP_WARPS * SP_STAGES * piece_bytes + warp * (SP_STAGES * 8);
  for (int i = threadIdx.x * 16; i < piece_bytes; i += SP_THREADS * 16) st_shared_v4(s_zero + i, 0u, 0u, 0u, 0u);
  if (lane == 0) {
    for (int st = 0; st < SP_STAGES; ++st) mbar_init(s_bar + 8 * st, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();

  const int pieces = (a.row_bytes + piece_bytes - 1) / piece_bytes;
  const int total_units = units_x * a.batch;
  uint32_t phase_bits = 0;  // bit st = parity of the next load phase to wait for on stage st (lane 0)
  uint32_t nitem = 0;       // items this warp has processed so far (stage = nitem % SP_STAGES)

  for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    const int b = unit / units_x;
    const int p0 = (unit - b * units_x) * SP_CHUNK;
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
      for (int w = 0; w < SP_WARPS; ++w) base += s_warp_count[w];
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

    // (3) move rows: warp w owns rows w, w + SP_WARPS, ...; item t = (row slot t / pieces, piece t % pieces)
    if (lane == 0) {
      const int nitems = (SP_CHUNK / SP_WARPS) * pieces;
      auto item = [&](int t, const uint8_t*& src, uint8_t*& dst, uint32_t& bytes) {
        const int r = warp + (t / pieces) * SP_WARPS;
        const int pc = t % pieces;
        const int off = pc * piece_bytes;
        bytes = static_cast<uint32_t>(a.row_bytes - off < piece_bytes ? a.row_bytes - off : piece_bytes);
        dst = s_dst[r] != nullptr ? s_dst[r] + off : nullptr;
        src = (dst != nullptr && s_src[r] != nullptr) ? s_src[r] + off : nullptr;
      };
      auto issue_load = [&](int t) {
        const uint8_t* src; uint8_t* dst; uint32_t bytes;
        item(t, src, dst, bytes);
        if (src == nullptr) return;
        const uint32_t st = (nitem + static_cast<uint32_t>(t)) % SP_STAGES;
        mbar_arrive_expect_tx(s_bar + 8 * st, bytes);
        bulk_g2s(s_ring + st * piece_bytes, src, bytes, s_bar + 8 * st);
      };
      // every stage the look-ahead loads are about to overwrite was last read by a store group that is at least
      // SP_STAGES - SP_LOOKAHEAD groups old
      bulk_wait_read<SP_STAGES - SP_LOOKAHEAD>();
      for (int t = 0; t < SP_LOOKAHEAD && t < nitems; ++t) issue_load(t);
      for (int t = 0; t < nitems; ++t) {
        const uint8_t* src; uint8_t* dst; uint32_t bytes;
        item(t, src, dst, bytes);
        const uint32_t st = (nitem + static_cast<uint32_t>(t)) % SP_STAGES;
        if (src != nullptr) {
          mbar_wait(s_bar + 8 * st, (phase_bits >> st) & 1u);
          phase_bits ^= 1u << st;
          bulk_s2g(dst, s_ring + st * piece_bytes, bytes);
        } else if (FWD && dst != nullptr) {
          bulk_s2g(dst, s_zero, bytes);  // text position without an embedding row: zeros
        }
        bulk_commit();  // one (possibly empty) group per item keeps the stage-reuse arithmetic uniform
        if (t + SP_LOOKAHEAD < nitems) {
          bulk_wait_read<SP_STAGES - SP_LOOKAHEAD>();
          issue_load(t + SP_LOOKAHEAD);
        }
      }
      nitem += static_cast<uint32_t>(nitems);
    }
    __syncthreads();  // s_src / s_dst / s_warp_count are rewritten by the next unit
  }
  if (lane == 0) bulk_wait_all<0>();
  __syncwarp();
}

// ---- light variant: register-staged 128-bit copies, ~0.5 KB of static shared memory and nothing dynamic, so that its
// CTAs fit next to a projector-GEMM CTA that owns the rest of the SM's shared memory.  Used for the fused step's
// forward, where the GEMM epilogue has already written the AV rows and only the few text rows + masks are left
// (it runs on a side stream UNDER the GEMM).  Grid (ceil(S/32), B).
constexpr int SL_THREADS = 256;
constexpr int SL_CHUNK = 32;
constexpr int SL_UNROLL = 8;

__device__ __forceinline__ void copy_row(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                         int row_bytes, int lane) {
  const int nvec = row_bytes >> 4;
  int i = lane;
  for (; i + (SL_UNROLL - 1) * 32 < nvec; i += SL_UNROLL * 32) {
    int4 v[SL_UNROLL];
#pragma unroll
    for (int u = 0; u < SL_UNROLL; ++u) v[u] = ld_nc_v4(src + (static_cast<int64_t>(i + u * 32) << 4));
#pragma unroll
    for (int u = 0; u < SL_UNROLL; ++u) st_na_v4(dst + (static_cast<int64_t>(i + u * 32) << 4), v[u]);
  }
  for (; i < nvec; i += 32) st_na_v4(dst + (static_cast<int64_t>(i) << 4), ld_nc_v4(src + (static_cast<int64_t>(i) << 4)));
}

__device__ __forceinline__ void zero_row(uint8_t* __restrict__ dst, int row_bytes, int lane) {
  const int nvec = row_bytes >> 4;
  const int4 z = make_int4(0, 0, 0, 0);
  for (int i = lane; i < nvec; i += 32) st_na_v4(dst + (static_cast<int64_t>(i) << 4), z);
}

template <bool FWD>
__global__ void __launch_bounds__(SL_THREADS) splice_light_kernel(const __grid_constant__ SpliceArgs a) {
  __shared__ int s_warp_count[SL_THREADS / 32];
  __shared__ const uint8_t* s_src[SL_CHUNK];
  __shared__ uint8_t* s_dst[SL_CHUNK];

  const int b = blockIdx.y;
  const int p0 = blockIdx.x * SL_CHUNK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t* ids = a.input_ids + static_cast<int64_t>(b) * a.seq;

  // (1) placeholders strictly left of this chunk
  int cnt = 0;
  for (int p = threadIdx.x; p < p0; p += SL_THREADS) cnt += (__ldg(ids + p) == a.placeholder_id) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0) s_warp_count[warp] = cnt;
  __syncthreads();

  // (2) packed-row index per position of the chunk + masks
  if (warp == 0) {
    int base = 0;
#pragma unroll
    for (int w = 0; w < SL_THREADS / 32; ++w) base += s_warp_count[w];
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
    if (a.status != nullptr && p0 + SL_CHUNK >= a.seq) {
      const int total = base + __popc(ball);
      if (lane == 0 && total != ntok) atomicOr(a.status, 1);
    }
  }
  __syncthreads();

  // (3) move rows
  for (int r = warp; r < SL_CHUNK; r += SL_THREADS / 32) {
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
  if (a.y != nullptr && (reinterpret_cast<uintptr_t>(a.y) & 15) != 0) return cudaErrorMisalignedAddress;
  if (a.dy != nullptr && (reinterpret_cast<uintptr_t>(a.dy) & 15) != 0) return cudaErrorMisalignedAddress;
  if (a.embed_table != nullptr && (reinterpret_cast<uintptr_t>(a.embed_table) & 15) != 0)
    return cudaErrorMisalignedAddress;
  if (a.batch <= 0 || a.seq <= 0) return cudaErrorInvalidValue;
  return cudaSuccess;
}

template <bool FWD>
static cudaError_t launch_splice(const SpliceArgs& a, int num_sms, cudaStream_t stream) {
  cudaError_t e = check_splice(a);
  if (e != cudaSuccess) return e;
  int piece = a.row_bytes < SP_MAX_PIECE ? a.row_bytes : SP_MAX_PIECE;
  piece = (piece + 127) & ~127;
  const size_t smem = static_cast<size_t>(piece) * (1 + SP_WARPS * SP_STAGES) + SP_WARPS * SP_STAGES * 8;
  e = cudaFuncSetAttribute(splice_kernel<FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  int ctas_per_sm = static_cast<int>((227 * 1024) / (smem + 1024 + 512));
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 8 ? 8 : ctas_per_sm);
  const int units_x = (a.seq + SP_CHUNK - 1) / SP_CHUNK;
  const int64_t units = static_cast<int64_t>(units_x) * a.batch;
  int64_t grid = static_cast<int64_t>(num_sms > 0 ? num_sms : 148) * ctas_per_sm;
  if (grid > units) grid = units;
  splice_kernel<FWD><<<static_cast<int>(grid), SP_THREADS, smem, stream>>>(a, piece, units_x);
  return cudaGetLastError();
}

cudaError_t launch_splice_fwd(const SpliceArgs& a, int num_sms, cudaStream_t stream) {
  if (a.av_in_place) {  // text rows + masks only: the light kernel co-resides with the GEMM it runs beside
    cudaError_t e = check_splice(a);
    if (e != cudaSuccess) return e;
    if (a.batch > 65535) return cudaErrorInvalidValue;
    dim3 grid((a.seq + SL_CHUNK - 1) / SL_CHUNK, a.batch);
    splice_light_kernel<true><<<grid, SL_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
  }
  return launch_splice<true>(a, num_sms, stream);
}

cudaError_t launch_splice_bwd(const SpliceArgs& a, int num_sms, cudaStream_t stream) {
  return launch_splice<false>(a, num_sms, stream);
}

}  // namespace avc
