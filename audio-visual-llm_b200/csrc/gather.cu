// Gather: temporal align + stride-k frame stacking + feature concat, one pass, HBM-bound.
//
//   A[m, :] = [ audio[b, k_a*j .. k_a*j+k_a-1, :]  ;  video[b, k_v*j .. k_v*j+k_v-1, :] ]   m = (b, j)
//
// Frames at or beyond the sample's valid length are zero.  With k_a = k_v = 1 this is the reference's
// index-by-index alignment + zero pad of the shorter stream (clip_whisper_model.py:424-431); a video
// frame stride of (1+Np)*Dv also folds in the CLS-row select of clip_whisper_model.py:1141-1142.
//
// Data never touches registers: each warp runs a ring of smem stages; one elected lane issues
// cp.async.bulk (TMA, global->smem, mbarrier complete_tx) loads G_LOOKAHEAD items ahead and
// cp.async.bulk (smem->global, bulk_group) stores behind.  Padding comes from a zeroed smem block.
#include "avc_kernels.h"
#include "avc_ptx.cuh"

namespace avc {

namespace {

constexpr int G_WARPS = 4;
constexpr int G_STAGES = 6;
constexpr int G_LOOKAHEAD = 4;  // loads in flight per warp; stage reuse distance = G_STAGES - G_LOOKAHEAD stores

struct Item {
  const uint8_t* src;
  uint8_t* dst;
  int nvalid;   // frames to copy
  int nzero;    // frames to zero-fill
  int mod;
};
constexpr int G_MAX_STAGE_BYTES = 8192;  // a stack wider than this is moved in chunks of whole frames

__device__ __forceinline__ void locate_row(const GatherArgs& a, int64_t m, int& b, int& j) {
  if (a.tok_offset == nullptr) {
    b = static_cast<int>(m / a.tokens_per_sample);
    j = static_cast<int>(m - static_cast<int64_t>(b) * a.tokens_per_sample);
    return;
  }
  int lo = 0, hi = a.batch;  // find b with tok_offset[b] <= m < tok_offset[b+1]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(a.tok_offset + mid) <= m) lo = mid; else hi = mid;
  }
  b = lo;
  j = static_cast<int>(m - __ldg(a.tok_offset + lo));
}

__device__ __forceinline__ int valid_len(const GatherArgs& a, int mod, int b) {
  int len = a.frames[mod];
  if (a.len[mod] != nullptr) {
    const int l = __ldg(a.len[mod] + b);
    len = l < len ? l : len;
    len = len < 0 ? 0 : len;
  }
  return len;
}

// An item = one chunk of one stream's stack of one token: `fpc[mod]` consecutive frames (the whole stack when it fits
// in a stage, i.e. nch[mod] == 1).  Items of a row are numbered [audio chunks | video chunks].
__global__ void __launch_bounds__(G_WARPS * 32) gather_kernel(const __grid_constant__ GatherArgs a, int nmods,
                                                               int mod0, int stage_bytes,
                                                               int zero_bytes, int fpc0, int fpc1, int nch0, int nch1) {
  extern __shared__ __align__(128) uint8_t g_smem[];
  // layout: [zero block][warp rings][barriers]
  const uint32_t s_zero = smem_u32(g_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_ring = s_zero + zero_bytes + warp * (G_STAGES * stage_bytes);
  const uint32_t s_bar = s_zero + zero_bytes + G_WARPS * G_STAGES * stage_bytes +
                         warp * (G_STAGES * 8);

  for (int i = threadIdx.x * 16; i < zero_bytes; i += blockDim.x * 16)
    st_shared_v4(s_zero + i, 0u, 0u, 0u, 0u);
  if (lane == 0) {
    for (int s = 0; s < G_STAGES; ++s) mbar_init(s_bar + 8 * s, 1);
    fence_mbar_init();
  }
  fence_proxy_async_smem();
  __syncthreads();

  if (lane == 0) {
    const int per_row = nch0 + nch1;
    const int64_t total_items = a.total_rows * per_row;
    const int64_t gwarps = static_cast<int64_t>(gridDim.x) * G_WARPS;
    const int64_t gw = static_cast<int64_t>(blockIdx.x) * G_WARPS + warp;
    const int64_t my_items = gw < total_items ? (total_items - gw + gwarps - 1) / gwarps : 0;
    const int64_t seg1_off = (a.src[0] != nullptr) ? static_cast<int64_t>(a.k[0]) * a.frame_bytes[0] : 0;

    auto make_item = [&](int64_t n) {
      Item it;
      const int64_t id = gw + n * gwarps;
      const int64_t m = id / per_row;
      const int sub = static_cast<int>(id - m * per_row);
      it.mod = sub < nch0 ? 0 : 1;
      const int chunk = sub < nch0 ? sub : sub - nch0;
      int b, j;
      locate_row(a, m, b, j);
      const int mod = it.mod;
      const int fpc = mod == 0 ? fpc0 : fpc1;
      const int len = valid_len(a, mod, b);
      const int c0 = chunk * fpc;                                   // first frame of the chunk inside the stack
      const int cf = a.k[mod] - c0 < fpc ? a.k[mod] - c0 : fpc;     // frames in this chunk
      const int64_t f0 = static_cast<int64_t>(j / a.rep[mod]) * a.k[mod] + c0;
      int64_t nv = len - f0;
      nv = nv < 0 ? 0 : (nv > cf ? cf : nv);
      it.nvalid = static_cast<int>(nv);
      it.nzero = cf - it.nvalid;
      it.src = a.src[mod] + b * a.batch_stride[mod] + f0 * a.frame_stride[mod];
      it.dst = a.dst + m * a.dst_row_bytes + (mod == 1 ? seg1_off : 0) + static_cast<int64_t>(c0) * a.frame_bytes[mod];
      if (a.row_flags != nullptr && chunk == 0 && (nmods == 1 || mod == 0)) {
        uint8_t fl = 0;
        if (a.src[0] != nullptr) {
          const int la = (mod == 0) ? len : valid_len(a, 0, b);
          if (static_cast<int64_t>(j / a.rep[0]) * a.k[0] < la) fl |= 1;
        }
        if (a.src[1] != nullptr) {
          const int lv = (mod == 1) ? len : valid_len(a, 1, b);
          if (static_cast<int64_t>(j / a.rep[1]) * a.k[1] < lv) fl |= 2;
        }
        a.row_flags[m] = fl;
      }
      return it;
    };
    auto issue_load = [&](const Item& it, int64_t n) {
      if (it.nvalid == 0) return;
      const int s = static_cast<int>(n % G_STAGES);
      const uint32_t bar = s_bar + 8 * s;
      const uint32_t buf = s_ring + s * stage_bytes;
      const int fb = a.frame_bytes[it.mod];
      mbar_arrive_expect_tx(bar, static_cast<uint32_t>(it.nvalid * fb));
      if (a.frame_stride[it.mod] == fb) {
        bulk_g2s(buf, it.src, static_cast<uint32_t>(it.nvalid * fb), bar);
      } else {
        for (int f = 0; f < it.nvalid; ++f)
          bulk_g2s(buf + f * fb, it.src + f * a.frame_stride[it.mod], fb, bar);
      }
    };

    // loads run G_LOOKAHEAD items ahead of stores; the items in flight are kept in registers
    Item ring[G_LOOKAHEAD];
#pragma unroll
    for (int i = 0; i < G_LOOKAHEAD; ++i) {
      if (i < my_items) {
        ring[i] = make_item(i);
        issue_load(ring[i], i);
      }
    }
    uint32_t phase_bits = 0;  // bit s = parity of the next load phase to wait for on stage s
    for (int64_t n0 = 0; n0 < my_items; n0 += G_LOOKAHEAD) {
#pragma unroll
      for (int i = 0; i < G_LOOKAHEAD; ++i) {
        const int64_t n = n0 + i;
        if (n >= my_items) break;
        const Item it = ring[i];
        const int s = static_cast<int>(n % G_STAGES);
        const int fb = a.frame_bytes[it.mod];
        if (it.nvalid > 0) {
          mbar_wait(s_bar + 8 * s, (phase_bits >> s) & 1u);
          phase_bits ^= 1u << s;
          bulk_s2g(it.dst, s_ring + s * stage_bytes, static_cast<uint32_t>(it.nvalid * fb));
        }
        for (int f = 0; f < it.nzero; ++f)
          bulk_s2g(it.dst + static_cast<int64_t>(it.nvalid + f) * fb, s_zero, fb);
        bulk_commit();
        const int64_t na = n + G_LOOKAHEAD;
        if (na < my_items) {
          // stage (na % G_STAGES) was last read by the store group of item na - G_STAGES = n - (G_STAGES -
          // G_LOOKAHEAD); the G_STAGES - G_LOOKAHEAD groups committed after it may still be reading their stages
          bulk_wait_read<G_STAGES - G_LOOKAHEAD>();
          ring[i] = make_item(na);
          issue_load(ring[i], na);
        }
      }
    }
    bulk_wait_all<0>();
  }
  __syncwarp();
}

}  // namespace

cudaError_t launch_gather(const GatherArgs& a, int num_sms, cudaStream_t stream) {
  if (a.total_rows <= 0) return cudaSuccess;
  const int nmods = (a.src[0] != nullptr ? 1 : 0) + (a.src[1] != nullptr ? 1 : 0);
  if (nmods == 0) return cudaErrorInvalidValue;
  const int mod0 = a.src[0] != nullptr ? 0 : 1;
  int stage_bytes = 0, zero_bytes = 16;
  int fpc[2] = {1, 1}, nch[2] = {0, 0};
  for (int i = 0; i < 2; ++i) {
    if (a.src[i] == nullptr) continue;
    if (a.frame_bytes[i] % 16 != 0 || a.frame_stride[i] % 16 != 0 || a.batch_stride[i] % 16 != 0 ||
        (reinterpret_cast<uintptr_t>(a.src[i]) & 15) != 0)
      return cudaErrorMisalignedAddress;
    // whole stack per item when it fits in a stage, else chunks of whole frames (a frame wider than a stage is its
    // own chunk: the per-CTA shared-memory check below then decides)
    fpc[i] = a.k[i] * a.frame_bytes[i] <= G_MAX_STAGE_BYTES ? a.k[i]
                                                            : (G_MAX_STAGE_BYTES / a.frame_bytes[i] > 0
                                                                   ? G_MAX_STAGE_BYTES / a.frame_bytes[i] : 1);
    nch[i] = (a.k[i] + fpc[i] - 1) / fpc[i];
    const int seg = fpc[i] * a.frame_bytes[i];
    stage_bytes = seg > stage_bytes ? seg : stage_bytes;
    zero_bytes = a.frame_bytes[i] > zero_bytes ? a.frame_bytes[i] : zero_bytes;
  }
  if ((reinterpret_cast<uintptr_t>(a.dst) & 15) != 0 || a.dst_row_bytes % 16 != 0)
    return cudaErrorMisalignedAddress;
  stage_bytes = (stage_bytes + 127) & ~127;
  zero_bytes = (zero_bytes + 127) & ~127;
  const size_t smem = static_cast<size_t>(zero_bytes) + G_WARPS * G_STAGES * stage_bytes +
                      G_WARPS * G_STAGES * 8;
  if (smem > 227 * 1024) return cudaErrorInvalidConfiguration;
  cudaError_t e = cudaFuncSetAttribute(gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  int ctas_per_sm = static_cast<int>((227 * 1024) / (smem + 1024));
  ctas_per_sm = ctas_per_sm < 1 ? 1 : (ctas_per_sm > 4 ? 4 : ctas_per_sm);
  const int64_t total_items = a.total_rows * (nch[0] + nch[1]);
  int64_t grid = static_cast<int64_t>(num_sms) * ctas_per_sm;
  const int64_t need = (total_items + G_WARPS - 1) / G_WARPS;
  if (grid > need) grid = need;
  gather_kernel<<<static_cast<int>(grid), G_WARPS * 32, smem, stream>>>(a, nmods, mod0, stage_bytes, zero_bytes, fpc[0],
                                                                         fpc[1], nch[0], nch[1]);
  return cudaGetLastError();
}

}  // namespace avc
