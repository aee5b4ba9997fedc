// C ABI of the B200-native clip_whisper connector (include/avconnector_b200.h).
// Validates arguments, builds TMA tensor maps, and enqueues the sm_100a kernels on the caller's
// stream.  Nothing here falls back to the CPU; only the avc_comm_* peer-memory helpers allocate or synchronise.
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#include <mutex>

#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include "../../include/avconnector_b200.h"
#include "avc_kernels.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int cuda_fail(cudaError_t e, const char* what) {
  return fail(AVC_ERR_CUDA, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

struct DeviceInfo {
  int device = -1;
  int num_sms = 0;
  int cc_major = 0;
};

// current device's properties, cached per device id
int device_info(DeviceInfo* out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  if (dev < 0 || dev >= 64) return fail(AVC_ERR_UNSUPPORTED, "device index %d out of range", dev);
  std::lock_guard<std::mutex> lock(mu);
  if (cache[dev].device != dev) {
    DeviceInfo d;
    d.device = dev;
    e = cudaDeviceGetAttribute(&d.num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(sm count)");
    e = cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute(cc major)");
    cache[dev] = d;
  }
  *out = cache[dev];
  if (out->cc_major != 10)
    return fail(AVC_ERR_UNSUPPORTED,
                "device %d has compute capability %d.x; this library only runs on sm_100 (B200) "
                "and has no CPU or other-GPU fallback",
                dev, out->cc_major);
  return AVC_OK;
}

PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }();
  return fn;
}

// 3-D tensor map over [batches][rows][cols] (cols contiguous), 128-byte swizzle.
// elem: avc::GemmOut code of the element type (bf16 / fp32 / fp16)
int make_map3d(CUtensorMap* map, const void* ptr, int elem, int64_t cols, int64_t rows,
               int64_t batches, int64_t row_stride_elems, int64_t batch_stride_elems, int box_cols,
               int box_rows, const char* what) {
  auto enc = encode_fn();
  if (enc == nullptr) return fail(AVC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not found");
  const bool fp32 = elem == avc::GEMM_OUT_F32;
  const int64_t es = fp32 ? 4 : 2;
  if (ptr == nullptr) return fail(AVC_ERR_INVALID, "%s: null pointer", what);
  if (reinterpret_cast<uintptr_t>(ptr) & 15) return fail(AVC_ERR_INVALID, "%s: base not 16-byte aligned", what);
  if (cols <= 0 || rows <= 0 || batches <= 0) return fail(AVC_ERR_INVALID, "%s: empty extent", what);
  if ((row_stride_elems * es) % 16 != 0 || row_stride_elems < cols)
    return fail(AVC_ERR_INVALID, "%s: row stride %lld must be >= cols and a multiple of 16 bytes", what,
                static_cast<long long>(row_stride_elems));
  if (batches == 1) batch_stride_elems = rows * row_stride_elems;
  if ((batch_stride_elems * es) % 16 != 0)
    return fail(AVC_ERR_INVALID, "%s: batch stride must be a multiple of 16 bytes", what);
  // Encoded maps are a pure function of (pointer, element type, extents, strides, box): a small per-thread
  // direct-mapped cache spares the driver call when a caller comes back with the same buffers (decode loops, the
  // static-buffer step engine).  SURVEY.md 8(b): "no global state beyond cached TMA descriptors".
  struct Key {
    const void* ptr;
    int64_t cols, rows, batches, rs, bs;
    int elem, box_cols, box_rows;
    bool operator==(const Key& o) const {
      return ptr == o.ptr && cols == o.cols && rows == o.rows && batches == o.batches && rs == o.rs && bs == o.bs &&
             elem == o.elem && box_cols == o.box_cols && box_rows == o.box_rows;
    }
  };
  struct Slot { Key key; CUtensorMap map; bool used; };
  constexpr int kSlots = 64;
  thread_local Slot slots[kSlots] = {};
  const Key key{ptr, cols, rows, batches, row_stride_elems, batch_stride_elems, elem, box_cols, box_rows};
  uint64_t hsh = reinterpret_cast<uintptr_t>(ptr) >> 4;
  hsh = (hsh ^ static_cast<uint64_t>(cols) * 0x9E3779B97F4A7C15ull ^ static_cast<uint64_t>(rows) * 0xC2B2AE3D27D4EB4Full ^
         static_cast<uint64_t>(box_cols) * 0x165667B19E3779F9ull ^ static_cast<uint64_t>(box_rows) * 0x27D4EB2F165667C5ull ^
         static_cast<uint64_t>(batches) * 0x85EBCA77C2B2AE63ull) * 0xD6E8FEB86659FD93ull;
  Slot& slot = slots[(hsh >> 32) % kSlots];
  if (slot.used && slot.key == key) {
    *map = slot.map;
    return AVC_OK;
  }
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows),
                        static_cast<cuuint64_t>(batches)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(row_stride_elems * es),
                           static_cast<cuuint64_t>(batch_stride_elems * es)};
  cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                      : (elem == avc::GEMM_OUT_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                                                                   : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = enc(map, dt, 3,
                   const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(AVC_ERR_CUDA, "%s: cuTensorMapEncodeTiled failed with CUresult %d", what, static_cast<int>(r));
  slot.key = key;
  slot.map = *map;
  slot.used = true;
  return AVC_OK;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace

extern "C" {

// used by the other translation units of the library (avc_mc.cu) to record the thread-local error message
__attribute__((visibility("hidden"))) int avc_set_error_(int code, const char* msg) { return fail(code, "%s", msg); }

int avc_abi_version(void) { return AVC_ABI_VERSION; }

const char* avc_last_error(void) { return g_err; }

int avc_device_check(int device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceCount");
  if (device < 0 || device >= count)
    return fail(AVC_ERR_UNSUPPORTED, "no CUDA device %d (found %d); there is no CPU fallback", device, count);
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  if (major != 10)
    return fail(AVC_ERR_UNSUPPORTED, "device %d is compute capability %d.x, need 10.x (sm_100a)", device, major);
  return AVC_OK;
}

int avc_gather_fwd(const avc_feat* audio, const avc_feat* video, int32_t batch, const int32_t* tok_offset,
                   int32_t tokens_per_sample, int64_t total_rows, void* a_out, int64_t a_row_stride,
                   uint8_t* row_flags, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (batch <= 0) return fail(AVC_ERR_INVALID, "gather: batch must be positive");
  if (a_out == nullptr) return fail(AVC_ERR_INVALID, "gather: null output");
  if (tok_offset == nullptr && tokens_per_sample <= 0)
    return fail(AVC_ERR_INVALID, "gather: tokens_per_sample must be positive when tok_offset is NULL");
  if (tok_offset == nullptr && total_rows != static_cast<int64_t>(batch) * tokens_per_sample)
    return fail(AVC_ERR_INVALID, "gather: total_rows != batch * tokens_per_sample");
  avc::GatherArgs g;
  memset(&g, 0, sizeof(g));
  const avc_feat* f[2] = {audio, video};
  int64_t k_total = 0;
  g.rep[0] = g.rep[1] = 1;
  for (int i = 0; i < 2; ++i) {
    if (f[i] == nullptr || f[i]->ptr == nullptr) continue;
    if (f[i]->repeat < 0) return fail(AVC_ERR_INVALID, "gather: negative repeat");
    g.rep[i] = f[i]->repeat > 0 ? f[i]->repeat : 1;
    if (f[i]->dim <= 0 || f[i]->dim % 8 != 0)
      return fail(AVC_ERR_INVALID, "gather: feature dim %d must be a positive multiple of 8", f[i]->dim);
    if (f[i]->stack < 1 || f[i]->frames < 0) return fail(AVC_ERR_INVALID, "gather: bad stack / frames");
    if (f[i]->frame_stride < f[i]->dim) return fail(AVC_ERR_INVALID, "gather: frame stride < dim");
    g.src[i] = static_cast<const uint8_t*>(f[i]->ptr);
    g.batch_stride[i] = f[i]->batch_stride * 2;
    g.frame_stride[i] = f[i]->frame_stride * 2;
    g.frame_bytes[i] = f[i]->dim * 2;
    g.frames[i] = f[i]->frames;
    g.k[i] = f[i]->stack;
    g.len[i] = f[i]->valid_frames;
    k_total += static_cast<int64_t>(f[i]->stack) * f[i]->dim;
  }
  if (k_total == 0) return fail(AVC_ERR_INVALID, "gather: both modalities absent");
  if (a_row_stride < k_total)
    return fail(AVC_ERR_INVALID, "gather: output row stride %lld < stacked width %lld",
                static_cast<long long>(a_row_stride), static_cast<long long>(k_total));
  g.tok_offset = tok_offset;
  g.tokens_per_sample = tokens_per_sample;
  g.batch = batch;
  g.total_rows = total_rows;
  g.dst = static_cast<uint8_t*>(a_out);
  g.dst_row_bytes = a_row_stride * 2;
  g.row_flags = row_flags;
  cudaError_t e = avc::launch_gather(g, di.num_sms, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "gather launch");
  return AVC_OK;
}

int avc_proj_fwd(int32_t nseg, const avc_mat* a, const avc_mat* w, const avc_mat* y, int32_t y_dtype,
                 const float* bias0, const float* bias1, float bias_scale0, float bias_scale1,
                 const uint8_t* row_flags, int32_t flag_rows0, int32_t flag_rows1, int32_t act, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (y_dtype != AVC_DTYPE_BF16 && y_dtype != AVC_DTYPE_F32 && y_dtype != AVC_DTYPE_F16)
    return fail(AVC_ERR_INVALID, "proj_fwd: y_dtype must be AVC_DTYPE_BF16 (0), AVC_DTYPE_F32 (1) or AVC_DTYPE_F16 (2)");
  const bool y_is_fp32 = y_dtype == AVC_DTYPE_F32;
  if (nseg < 1 || nseg > 2) return fail(AVC_ERR_INVALID, "proj_fwd: nseg must be 1 or 2");
  if (a == nullptr || w == nullptr || y == nullptr) return fail(AVC_ERR_INVALID, "proj_fwd: null matrix");
  if (act != 0 && act != 1) return fail(AVC_ERR_INVALID, "proj_fwd: act must be 0 or 1");
  const int64_t N = y->cols;
  if (N <= 0 || N % 8 != 0) return fail(AVC_ERR_INVALID, "proj_fwd: N must be a positive multiple of 8");
  if (y->rows <= 0 || y->batches <= 0) return fail(AVC_ERR_INVALID, "proj_fwd: empty output");
  avc::GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.nseg = nseg;
  const int cg = avc::gemm_cta_group();
  const int mt = avc::gemm_m_subtiles(cg, avc::GEMM_TN);
  // scatter mode: the operand rows are packed [B*N, K] while Y is a strided [B][N][H] region (e.g. the AV rows of
  // inputs_embeds): M tiles run over the packed rows and the epilogue splits boxes at sample boundaries
  const bool scatter = a[0].batches == 1 && y->batches > 1;
  const int64_t m_rows = scatter ? a[0].rows : y->rows;
  const int64_t m_batches = scatter ? 1 : y->batches;
  if (scatter && row_flags == nullptr && (flag_rows0 < m_rows || flag_rows1 < m_rows) && (bias0 || bias1))
    return fail(AVC_ERR_INVALID, "proj_fwd: analytic row flags are per packed row in scatter mode; pass row_flags");
  g.m_tiles_per_batch = static_cast<int>(ceil_div(m_rows, avc::GEMM_BM * cg * mt));
  g.num_m_blocks = static_cast<int>(m_batches) * g.m_tiles_per_batch;
  if (scatter)
    if (int rc = make_map3d(&g.md_row, y->ptr, y_dtype, N, y->rows, y->batches, y->row_stride,
                            y->batch_stride, y_is_fp32 ? 32 : 64, 1, "proj_fwd Y (row box)"))
      return rc;
  g.scatter_rows = scatter ? static_cast<int>(y->rows) : 0;
  g.scatter_batches = scatter ? static_cast<int>(y->batches) : 0;
  g.bn = avc::pick_gemm_bn(g.num_m_blocks, &N, 1, di.num_sms / cg);
  for (int s = 0; s < nseg; ++s) {
    if (a[s].cols != w[s].cols)
      return fail(AVC_ERR_INVALID, "proj_fwd: segment %d: A has K=%lld but W has K=%lld", s,
                  static_cast<long long>(a[s].cols), static_cast<long long>(w[s].cols));
    if (w[s].rows != N) return fail(AVC_ERR_INVALID, "proj_fwd: segment %d: W rows != N", s);
    if (a[s].batches != y->batches && !(a[s].batches == 1 && a[s].rows == y->batches * y->rows))
      return fail(AVC_ERR_INVALID, "proj_fwd: segment %d: batch mismatch", s);
    if (s > 0 && a[s].batches != a[0].batches) return fail(AVC_ERR_INVALID, "proj_fwd: segments disagree on batching");
    if (a[s].cols % 8 != 0) return fail(AVC_ERR_INVALID, "proj_fwd: K must be a multiple of 8");
    if (int rc = make_map3d(&g.ma[s], a[s].ptr, false, a[s].cols, a[s].rows, a[s].batches, a[s].row_stride,
                            a[s].batch_stride, avc::GEMM_BK, avc::GEMM_BM * mt, "proj_fwd A"))
      return rc;
    if (int rc = make_map3d(&g.mb[s], w[s].ptr, false, w[s].cols, w[s].rows, 1, w[s].row_stride, 0,
                            avc::GEMM_BK, g.bn / cg, "proj_fwd W"))
      return rc;
    g.seg_kblocks[s] = static_cast<int>(ceil_div(a[s].cols, avc::GEMM_BK));
  }
  if (int rc = make_map3d(&g.md[0], y->ptr, y_dtype, N, y->rows, y->batches, y->row_stride,
                          y->batch_stride, y_is_fp32 ? 32 : 64, 32, "proj_fwd Y"))
    return rc;
  g.num_n_blocks = static_cast<int>(ceil_div(N, g.bn));
  g.d_rows = static_cast<int>(m_rows);
  g.d_cols[0] = static_cast<int>(N);
  g.d_cols[1] = static_cast<int>(N);
  g.bias0 = bias0;
  g.bias1 = bias1;
  g.row_flags = row_flags;
  g.flag_rows0 = flag_rows0;
  g.flag_rows1 = flag_rows1;
  g.alpha[0] = g.alpha[1] = 1.f;
  g.bias_scale[0] = bias_scale0;
  g.bias_scale[1] = bias_scale1;
  g.act = act;
  if ((bias0 && (reinterpret_cast<uintptr_t>(bias0) & 15)) || (bias1 && (reinterpret_cast<uintptr_t>(bias1) & 15)))
    return fail(AVC_ERR_INVALID, "proj_fwd: bias pointers must be 16-byte aligned");
  cudaError_t e = avc::launch_gemm(g, avc::GEMM_TN, static_cast<avc::GemmOut>(y_dtype), cg, mt, di.num_sms,
                                   static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "proj_fwd launch");
  return AVC_OK;
}

int avc_proj_bwd_dx(const avc_mat* dy, int32_t nseg, const avc_mat* w_t, const avc_mat* dx, int32_t dx_dtype,
                    void* stream) {
  // dX = dY . W is the forward GEMM with A = dY ([.., H], K-major in H) and B = W^T ([K_in, H], from avc_pack_weight_t)
  if (dy == nullptr || w_t == nullptr || dx == nullptr) return fail(AVC_ERR_INVALID, "proj_bwd_dx: null matrix");
  return avc_proj_fwd(nseg, dy, w_t, dx, dx_dtype, nullptr, nullptr, 0.f, 0.f, nullptr, 0, 0, 0, stream);
}

// avc_comm -> kernel arguments (pointer / range validation; offsets are relative to the local bucket)
static int fill_comm(const avc_comm* c, const float* extra0, int64_t extra0_len, const float* extra1,
                     int64_t extra1_len, avc::CommArgs* out) {
  if (c == nullptr) return fail(AVC_ERR_INVALID, "comm: null descriptor");
  if (c->world < 1 || c->world > avc::COMM_MAX_WORLD || c->rank < 0 || c->rank >= c->world)
    return fail(AVC_ERR_INVALID, "comm: world %d / rank %d out of range (1..%d)", c->world, c->rank,
                avc::COMM_MAX_WORLD);
  if (c->epoch == 0) return fail(AVC_ERR_INVALID, "comm: epochs count from 1");
  if (c->status == nullptr) return fail(AVC_ERR_INVALID, "comm: null status word");
  if (c->bucket_bytes == 0 || c->bucket_bytes % 16 != 0)
    return fail(AVC_ERR_INVALID, "comm: bucket_bytes must be a positive multiple of 16");
  memset(out, 0, sizeof(*out));
  out->world = c->world;
  out->rank = c->rank;
  out->epoch = c->epoch;
  if (c->mc_bucket != nullptr && (reinterpret_cast<uintptr_t>(c->mc_bucket) & 15))
    return fail(AVC_ERR_INVALID, "comm: multicast bucket pointer must be 16-byte aligned");
  out->mc = static_cast<float*>(c->mc_bucket);
  for (int p = 0; p < c->world; ++p) {
    // with a multicast bucket the peers' copies are reached through it: only the local bucket has to be mapped
    const bool need_bucket = c->mc_bucket == nullptr || p == c->rank;
    if ((need_bucket && c->bucket[p] == nullptr) || c->flags[p] == nullptr)
      return fail(AVC_ERR_INVALID, "comm: rank %d's bucket / flag area is not mapped", p);
    if ((reinterpret_cast<uintptr_t>(c->bucket[p]) & 15) || (reinterpret_cast<uintptr_t>(c->flags[p]) & 15))
      return fail(AVC_ERR_INVALID, "comm: bucket / flag pointers must be 16-byte aligned");
    out->data[p] = static_cast<float*>(c->bucket[p]);
    out->flags[p] = static_cast<uint32_t*>(c->flags[p]);
  }
  const float* ex[2] = {extra0, extra1};
  const int64_t len[2] = {extra0_len, extra1_len};
  for (int k = 0; k < 2; ++k) {
    if (ex[k] == nullptr || len[k] == 0) continue;
    const int64_t off = ex[k] - static_cast<const float*>(c->bucket[c->rank]);
    if (off < 0 || off % 4 != 0 || len[k] < 0 || len[k] % 4 != 0 || len[k] > (1 << 30) ||
        static_cast<uint64_t>(off + len[k]) * 4 > c->bucket_bytes)
      return fail(AVC_ERR_INVALID, "comm: extra range %d must start inside the local bucket, 16-byte aligned, and "
                  "hold a multiple of 4 floats", k);
    out->extra_off[k] = off;
    out->extra_len[k] = static_cast<int>(len[k]);
  }
  out->status = c->status;
  out->timeout_ns = c->timeout_ns != 0 ? c->timeout_ns : 20000000000ull;
  return AVC_OK;
}

// Byte layout of the split-reduction workspace for S slices: [arrival counters][partial bias sums][partial dW tiles]
struct SplitLayout {
  size_t counters, bias, tiles, total;
};
static SplitLayout split_layout(int base_items, int64_t H, int64_t ktot, int splits) {
  SplitLayout l;
  // the arrival counters live in a FIXED 64 KB header: a workspace re-used by launches of other shapes must never have
  // partial sums written where a later launch expects zeroed counters (split launches have < 148 base items)
  (void)base_items;
  l.counters = 65536;
  l.bias = (static_cast<size_t>(splits) * 2 * H * sizeof(float) + 255) & ~static_cast<size_t>(255);
  l.tiles = static_cast<size_t>(splits) * H * ktot * sizeof(float);
  l.total = l.counters + l.bias + l.tiles;
  return l;
}

static int proj_bwd_dw_impl(const avc_mat* dy, int32_t dy_row_base, int32_t nseg, const avc_mat* x, const avc_mat* dw,
                            const float* alpha, const avc_bias_grad* bias, const avc::CommArgs* comm,
                            uint64_t bucket_bytes, int32_t max_sms, void* stream, void* workspace = nullptr,
                            size_t workspace_bytes = 0, int32_t* plan_splits = nullptr, size_t* plan_bytes = nullptr) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (max_sms < 0) return fail(AVC_ERR_INVALID, "proj_bwd_dw: negative max_sms");
  if (max_sms >= 2 && max_sms < di.num_sms) di.num_sms = max_sms & ~1;  // whole CTA pairs
  if (nseg < 1 || nseg > 2) return fail(AVC_ERR_INVALID, "proj_bwd_dw: nseg must be 1 or 2");
  if (dy == nullptr || x == nullptr || dw == nullptr) return fail(AVC_ERR_INVALID, "proj_bwd_dw: null matrix");
  const int64_t H = dy->cols;
  if (H <= 0 || H % 8 != 0) return fail(AVC_ERR_INVALID, "proj_bwd_dw: H must be a positive multiple of 8");
  if (dy_row_base < 0 || dy_row_base >= dy->rows) return fail(AVC_ERR_INVALID, "proj_bwd_dw: dy_row_base out of range");
  avc::GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.nseg = nseg;
  if (int rc = make_map3d(&g.ma[0], dy->ptr, false, H, dy->rows, dy->batches, dy->row_stride, dy->batch_stride,
                          64, avc::GEMM_BK, "proj_bwd_dw dY"))
    return rc;
  int64_t red_rows = 0;
  int n_blocks[2] = {0, 0};
  int64_t n_ext[2] = {0, 0};
  for (int s = 0; s < nseg; ++s) n_ext[s] = x[s].cols;
  const int cg = avc::gemm_cta_group();
  const int mt = avc::gemm_m_subtiles(cg, avc::GEMM_NT);
  g.num_m_blocks = static_cast<int>(ceil_div(H, avc::GEMM_BM * cg * mt));
  g.bn = avc::pick_gemm_bn(g.num_m_blocks, n_ext, nseg, di.num_sms / cg);
  for (int s = 0; s < nseg; ++s) {
    if (x[s].batches != dy->batches) return fail(AVC_ERR_INVALID, "proj_bwd_dw: segment %d: batch mismatch", s);
    if (x[s].cols % 8 != 0) return fail(AVC_ERR_INVALID, "proj_bwd_dw: K must be a multiple of 8");
    if (dw[s].rows != H || dw[s].cols != x[s].cols)
      return fail(AVC_ERR_INVALID, "proj_bwd_dw: segment %d: dW must be [H, K]", s);
    if (int rc = make_map3d(&g.mb[s], x[s].ptr, false, x[s].cols, x[s].rows, x[s].batches, x[s].row_stride,
                            x[s].batch_stride, 64, avc::GEMM_BK, "proj_bwd_dw X"))
      return rc;
    if (int rc = make_map3d(&g.md[s], dw[s].ptr, true, dw[s].cols, dw[s].rows, 1, dw[s].row_stride, 0, 32, 32,
                            "proj_bwd_dw dW"))
      return rc;
    red_rows = x[s].rows > red_rows ? x[s].rows : red_rows;
    n_blocks[s] = static_cast<int>(ceil_div(x[s].cols, g.bn));
    g.d_cols[s] = static_cast<int>(x[s].cols);
    g.alpha[s] = alpha != nullptr ? alpha[s] : 1.f;
  }
  g.num_n_blocks = n_blocks[0] + n_blocks[1];
  g.n_blocks_seg0 = n_blocks[0];
  g.red_batches = static_cast<int>(dy->batches);
  g.red_kblocks_per_batch = static_cast<int>(ceil_div(red_rows, avc::GEMM_BK));
  g.a_row_base = dy_row_base;
  g.d_rows = static_cast<int>(H);
  if (bias != nullptr) {
    // db inside the same launch: one 64-wide work item per M block contracts the dY panel with the token-present
    // operand F (bf16 [batches][rows][64]: column 0 / 1 = the row carries an audio / video token)
    const avc_mat* f = bias->present;
    if (f == nullptr || f->ptr == nullptr) return fail(AVC_ERR_INVALID, "proj_bwd_dw_db: null token-present operand");
    if (f->cols != avc::GEMM_BIAS_COLS)
      return fail(AVC_ERR_INVALID, "proj_bwd_dw_db: token-present operand must have %d columns", avc::GEMM_BIAS_COLS);
    if (f->batches != dy->batches || f->rows < red_rows)
      return fail(AVC_ERR_INVALID, "proj_bwd_dw_db: token-present operand must be [batches][>= X rows][%d]",
                  avc::GEMM_BIAS_COLS);
    if (bias->out0 == nullptr && bias->out1 == nullptr) return fail(AVC_ERR_INVALID, "proj_bwd_dw_db: no output");
    if (int rc = make_map3d(&g.mf, f->ptr, 0, f->cols, red_rows, f->batches, f->row_stride, f->batch_stride, 64,
                            avc::GEMM_BK, "proj_bwd_dw_db F"))
      return rc;
    g.bias_items = g.num_m_blocks;
    g.bias_out[0] = bias->out0;
    g.bias_out[1] = bias->out1;
    g.bias_alpha[0] = bias->alpha0;
    g.bias_alpha[1] = bias->alpha1;
  }
  // few tiles: slices of the reduction fill the idle workers (needs a workspace; not together with the fused all-reduce)
  {
    int64_t ktot = 0;
    for (int s = 0; s < nseg; ++s) ktot += x[s].cols;
    const int base_items = g.num_m_blocks * g.num_n_blocks + g.bias_items;
    const int total_kb = g.red_batches * g.red_kblocks_per_batch;
    int splits = comm == nullptr ? avc::gemm_dw_splits(base_items, di.num_sms / cg, total_kb) : 1;
    if (plan_splits != nullptr) {  // planning call: report and return
      *plan_splits = splits;
      if (plan_bytes != nullptr) *plan_bytes = splits > 1 ? split_layout(base_items, H, ktot, splits).total : 0;
      return AVC_OK;
    }
    if (workspace == nullptr) splits = 1;
    while (splits > 1 && split_layout(base_items, H, ktot, splits).total > workspace_bytes) --splits;
    if (splits > 1) {
      if (reinterpret_cast<uintptr_t>(workspace) & 255)
        return fail(AVC_ERR_INVALID, "proj_bwd_dw: the split-reduction workspace must be 256-byte aligned");
      const SplitLayout l = split_layout(base_items, H, ktot, splits);
      uint8_t* wsp = static_cast<uint8_t*>(workspace);
      g.ksplit = splits;
      g.split_count = reinterpret_cast<uint32_t*>(wsp);
      g.ws_bias = reinterpret_cast<float*>(wsp + l.counters);
      g.ws = reinterpret_cast<float*>(wsp + l.counters + l.bias);
      g.ws_split_stride = H * ktot;
      g.ws_ld = static_cast<int>(ktot);
      int64_t col = 0;
      for (int s = 0; s < nseg; ++s) {
        g.ws_seg_col[s] = static_cast<int>(col);
        if (int rc = make_map3d(&g.mws[s], g.ws + col, 1, x[s].cols, H, splits, ktot, H * ktot, 32, 32,
                                "proj_bwd_dw split workspace"))
          return rc;
        g.d_ptr[s] = static_cast<float*>(dw[s].ptr);
        g.d_ld[s] = dw[s].row_stride;
        if (reinterpret_cast<uintptr_t>(dw[s].ptr) & 15 || dw[s].row_stride % 4 != 0)
          return fail(AVC_ERR_INVALID, "proj_bwd_dw: dW must be 16-byte aligned with a row stride that is a multiple of 4");
        col += x[s].cols;
      }
    }
  }
  if (comm != nullptr) {
    if (cg != 2) return fail(AVC_ERR_UNSUPPORTED, "proj_bwd_dw_allreduce needs the CTA-pair GEMM (AVC_GEMM_CTA_GROUP=2)");
    g.comm = *comm;
    for (int s = 0; s < nseg; ++s) {
      // every rank addresses tile (row, col) of segment s at the same offset of its bucket
      if (dw[s].row_stride != dw[s].cols)
        return fail(AVC_ERR_INVALID, "proj_bwd_dw_allreduce: dW segment %d must be contiguous", s);
      const int64_t off = static_cast<const float*>(dw[s].ptr) - comm->data[comm->rank];
      if (off < 0 || off % 4 != 0 || static_cast<uint64_t>(off + dw[s].rows * dw[s].cols) * 4 > bucket_bytes)
        return fail(AVC_ERR_INVALID, "proj_bwd_dw_allreduce: dW segment %d is not inside the local bucket", s);
      g.comm.seg_off[s] = off;
    }
    const int items = avc::gemm_work_items(g, cg, di.num_sms);
    if (items > avc::COMM_MAX_ITEMS)
      return fail(AVC_ERR_UNSUPPORTED, "proj_bwd_dw_allreduce: %d work items exceed the flag area (%d)", items,
                  avc::COMM_MAX_ITEMS);
  }
  cudaError_t e = avc::launch_gemm(g, avc::GEMM_NT, avc::GEMM_OUT_F32, cg, mt, di.num_sms,
                                   static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "proj_bwd_dw launch");
  return AVC_OK;
}

int avc_proj_bwd_dw(const avc_mat* dy, int32_t dy_row_base, int32_t nseg, const avc_mat* x, const avc_mat* dw,
                    const float* alpha, int32_t max_sms, void* stream) {
  return proj_bwd_dw_impl(dy, dy_row_base, nseg, x, dw, alpha, nullptr, nullptr, 0, max_sms, stream);
}

int avc_proj_bwd_dw_db(const avc_mat* dy, int32_t dy_row_base, int32_t nseg, const avc_mat* x, const avc_mat* dw,
                       const float* alpha, const avc_bias_grad* bias, void* workspace, size_t workspace_bytes,
                       int32_t max_sms, void* stream) {
  return proj_bwd_dw_impl(dy, dy_row_base, nseg, x, dw, alpha, bias, nullptr, 0, max_sms, stream, workspace,
                          workspace_bytes);
}

int avc_proj_bwd_dw_plan(const avc_mat* dy, int32_t nseg, const avc_mat* x, int32_t with_bias, int32_t* splits,
                         size_t* workspace_bytes) {
  if (splits == nullptr || workspace_bytes == nullptr) return fail(AVC_ERR_INVALID, "proj_bwd_dw_plan: null output");
  if (dy == nullptr || x == nullptr || nseg < 1 || nseg > 2) return fail(AVC_ERR_INVALID, "proj_bwd_dw_plan: bad operands");
  // the planner only looks at extents: stand-in output descriptors / bias descriptor with the right shapes
  avc_mat dw[2];
  for (int s = 0; s < nseg; ++s) {
    dw[s].ptr = x[s].ptr;  // never dereferenced by the planning path
    dw[s].rows = dy->cols; dw[s].cols = x[s].cols; dw[s].row_stride = x[s].cols; dw[s].batches = 1; dw[s].batch_stride = 0;
  }
  avc_mat f = *dy;
  f.cols = avc::GEMM_BIAS_COLS; f.row_stride = avc::GEMM_BIAS_COLS; f.rows = x[0].rows;
  f.batch_stride = f.rows * f.row_stride;
  float dummy = 0.f;
  avc_bias_grad b = {&f, &dummy, nullptr, 1.f, 1.f};
  return proj_bwd_dw_impl(dy, 0, nseg, x, dw, nullptr, with_bias ? &b : nullptr, nullptr, 0, 0, nullptr, nullptr, 0,
                          splits, workspace_bytes);
}

int avc_proj_bwd_dw_db_allreduce(const avc_mat* dy, int32_t dy_row_base, int32_t nseg, const avc_mat* x,
                                 const avc_mat* dw, const float* alpha, const avc_bias_grad* bias,
                                 const avc_comm* comm, int32_t max_sms, void* stream) {
  if (bias == nullptr) return fail(AVC_ERR_INVALID, "proj_bwd_dw_db_allreduce: null bias descriptor");
  if (dy == nullptr) return fail(AVC_ERR_INVALID, "proj_bwd_dw_db_allreduce: null dY");
  avc::CommArgs c;
  // the bias gradients are the launch's extra ranges: produced by its own bias items, reduced by its comm warps
  if (int rc = fill_comm(comm, bias->out0, bias->out0 ? dy->cols : 0, bias->out1, bias->out1 ? dy->cols : 0, &c))
    return rc;
  return proj_bwd_dw_impl(dy, dy_row_base, nseg, x, dw, alpha, bias, &c, comm->bucket_bytes, max_sms, stream);
}

int avc_proj_bwd_dw_allreduce(const avc_mat* dy, int32_t dy_row_base, int32_t nseg, const avc_mat* x,
                              const avc_mat* dw, const float* alpha, const avc_comm* comm, const float* extra0,
                              int64_t extra0_len, const float* extra1, int64_t extra1_len, int32_t max_sms,
                              void* stream) {
  avc::CommArgs c;
  if (int rc = fill_comm(comm, extra0, extra0_len, extra1, extra1_len, &c)) return rc;
  return proj_bwd_dw_impl(dy, dy_row_base, nseg, x, dw, alpha, nullptr, &c, comm->bucket_bytes, max_sms, stream);
}

int avc_comm_signal_extra(const avc_comm* comm, int64_t extra0_len, int64_t extra1_len, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  avc::CommArgs c;
  // only the lengths matter here (which ranks own a chunk); offsets are not dereferenced
  if (int rc = fill_comm(comm, nullptr, 0, nullptr, 0, &c)) return rc;
  if (extra0_len < 0 || extra1_len < 0 || extra0_len > (1 << 30) || extra1_len > (1 << 30))
    return fail(AVC_ERR_INVALID, "comm_signal_extra: bad range length");
  c.extra_len[0] = static_cast<int>(extra0_len);
  c.extra_len[1] = static_cast<int>(extra1_len);
  cudaError_t e = avc::launch_comm_signal_extra(c, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "comm_signal_extra launch");
  return AVC_OK;
}

int avc_debug_gemm_profile(void* device_buf) {
  avc::set_gemm_profile_buffer(static_cast<unsigned long long*>(device_buf));
  return AVC_OK;
}

size_t avc_comm_flag_bytes(void) { return static_cast<size_t>(avc::COMM_FLAG_WORDS) * 4; }

int avc_comm_alloc(size_t bytes, void** ptr) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (ptr == nullptr || bytes == 0) return fail(AVC_ERR_INVALID, "comm_alloc: bad argument");
  void* p = nullptr;
  cudaError_t e = avc::preload_comm_kernels();
  if (e == cudaSuccess) e = avc::preload_colsum();
  if (e != cudaSuccess) return cuda_fail(e, "comm_alloc kernel preload");
  e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return cuda_fail(e, "comm_alloc cudaMalloc");
  e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    return cuda_fail(e, "comm_alloc cudaMemset");
  }
  *ptr = p;
  return AVC_OK;
}

int avc_comm_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) return cuda_fail(e, "comm_free");
  return AVC_OK;
}

int avc_comm_export(const void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");
  if (ptr == nullptr || handle64 == nullptr) return fail(AVC_ERR_INVALID, "comm_export: null pointer");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, const_cast<void*>(ptr));
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcGetMemHandle");
  memcpy(handle64, &h, sizeof(h));
  return AVC_OK;
}

int avc_comm_open(const void* handle64, void** ptr) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (ptr == nullptr || handle64 == nullptr) return fail(AVC_ERR_INVALID, "comm_open: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  // maps the exporting process's allocation and enables peer access from the current device to its GPU
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcOpenMemHandle");
  *ptr = p;
  return AVC_OK;
}

int avc_comm_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) return cuda_fail(e, "cudaIpcCloseMemHandle");
  return AVC_OK;
}

size_t avc_colsum_workspace_bytes(int32_t cols) { return avc::colsum_workspace_bytes(cols); }

size_t avc_colsum_workspace_header_bytes(void) { return avc::colsum_workspace_header_bytes(); }

static int colsum_impl(const avc_mat* dy, int32_t dy_row_base, int32_t sum_rows, const uint8_t* row_flags,
                       int32_t flag_rows0, int32_t flag_rows1, float alpha0, float alpha1, float* out0, float* out1,
                       void* workspace, const avc_comm* comm, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (dy == nullptr || dy->ptr == nullptr) return fail(AVC_ERR_INVALID, "colsum: null dY");
  if (workspace == nullptr) return fail(AVC_ERR_INVALID, "colsum: null workspace");
  avc::ColsumArgs c;
  memset(&c, 0, sizeof(c));
  c.dy = static_cast<const uint8_t*>(dy->ptr);
  c.row_stride = dy->row_stride * 2;
  c.batch_stride = dy->batch_stride * 2;
  c.batch = static_cast<int>(dy->batches);
  if (dy_row_base < 0 || sum_rows < 0 || dy_row_base + sum_rows > dy->rows)
    return fail(AVC_ERR_INVALID, "colsum: rows [%d, %d) outside the %lld rows of dY", dy_row_base,
                dy_row_base + sum_rows, static_cast<long long>(dy->rows));
  c.rows = sum_rows;
  c.row_base = dy_row_base;
  c.cols = static_cast<int>(dy->cols);
  c.row_flags = row_flags;
  c.flag_rows0 = flag_rows0;
  c.flag_rows1 = flag_rows1;
  c.alpha0 = alpha0;
  c.alpha1 = alpha1;
  c.out0 = out0;
  c.out1 = out1;
  c.workspace = static_cast<float*>(workspace);
  if (comm != nullptr) {
    avc::CommArgs k;
    if (int rc = fill_comm(comm, out0, out0 ? dy->cols : 0, out1, out1 ? dy->cols : 0, &k)) return rc;
    // the ranks that own a 128-float chunk of the extra ranges wait for this flag (chunk e belongs to rank e % world)
    const int nch = ((k.extra_len[0] + 127) >> 7) + ((k.extra_len[1] + 127) >> 7);
    c.sig_owners = nch < k.world ? nch : k.world;
    c.sig_rank = k.rank;
    c.sig_epoch = k.epoch;
    for (int p = 0; p < k.world; ++p) c.sig_flags[p] = k.flags[p];
  }
  cudaError_t e = avc::launch_colsum(c, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "colsum launch");
  return AVC_OK;
}

int avc_colsum(const avc_mat* dy, int32_t dy_row_base, int32_t sum_rows, const uint8_t* row_flags, int32_t flag_rows0,
               int32_t flag_rows1, float alpha0, float alpha1, float* out0, float* out1, void* workspace,
               void* stream) {
  return colsum_impl(dy, dy_row_base, sum_rows, row_flags, flag_rows0, flag_rows1, alpha0, alpha1, out0, out1,
                     workspace, nullptr, stream);
}

int avc_colsum_comm(const avc_mat* dy, int32_t dy_row_base, int32_t sum_rows, const uint8_t* row_flags,
                    int32_t flag_rows0, int32_t flag_rows1, float alpha0, float alpha1, float* out0, float* out1,
                    void* workspace, const avc_comm* comm, void* stream) {
  if (comm == nullptr) return fail(AVC_ERR_INVALID, "colsum_comm: null comm descriptor");
  return colsum_impl(dy, dy_row_base, sum_rows, row_flags, flag_rows0, flag_rows1, alpha0, alpha1, out0, out1,
                     workspace, comm, stream);
}

int avc_pack_weight(const float* src, int64_t src_ld, void* dst_bf16, int64_t dst_ld, int64_t rows, int64_t cols,
                    float alpha, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (src == nullptr || dst_bf16 == nullptr) return fail(AVC_ERR_INVALID, "pack_weight: null pointer");
  cudaError_t e = avc::launch_pack_weight(src, src_ld, dst_bf16, dst_ld, rows, cols, alpha,
                                          static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "pack_weight launch");
  return AVC_OK;
}

int avc_cast_bf16(const void* src, int32_t src_dtype, int64_t src_ld, void* dst_bf16, int64_t dst_ld, int64_t rows,
                  int64_t cols, float alpha, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (src == nullptr || dst_bf16 == nullptr) return fail(AVC_ERR_INVALID, "cast_bf16: null pointer");
  if (src_dtype != AVC_DTYPE_BF16 && src_dtype != AVC_DTYPE_F32 && src_dtype != AVC_DTYPE_F16)
    return fail(AVC_ERR_INVALID, "cast_bf16: src_dtype must be an AVC_DTYPE_* code");
  cudaError_t e = avc::launch_cast_bf16(src, src_dtype, src_ld, dst_bf16, dst_ld, rows, cols, alpha,
                                        static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "cast_bf16 launch");
  return AVC_OK;
}

int avc_gather_bwd(const void* da, int64_t a_row_stride, int64_t col_off, const avc_feat* feat, int32_t out_dtype,
                   int32_t batch, const int32_t* tok_offset, int32_t tokens_per_sample, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (da == nullptr || feat == nullptr || feat->ptr == nullptr) return fail(AVC_ERR_INVALID, "gather_bwd: null pointer");
  if (out_dtype != AVC_DTYPE_BF16 && out_dtype != AVC_DTYPE_F32)
    return fail(AVC_ERR_INVALID, "gather_bwd: out_dtype must be AVC_DTYPE_BF16 or AVC_DTYPE_F32");
  if (batch <= 0) return fail(AVC_ERR_INVALID, "gather_bwd: batch must be positive");
  if (feat->dim <= 0 || feat->dim % 8 != 0) return fail(AVC_ERR_INVALID, "gather_bwd: dim must be a positive multiple of 8");
  if (feat->stack < 1 || feat->frames < 0 || feat->repeat < 0) return fail(AVC_ERR_INVALID, "gather_bwd: bad stack / frames / repeat");
  if (tok_offset == nullptr && tokens_per_sample <= 0)
    return fail(AVC_ERR_INVALID, "gather_bwd: tokens_per_sample must be positive when tok_offset is NULL");
  if (col_off < 0 || a_row_stride < col_off + static_cast<int64_t>(feat->stack) * feat->dim)
    return fail(AVC_ERR_INVALID, "gather_bwd: segment [col_off, col_off + stack * dim) exceeds the dA row");
  const int es = out_dtype == AVC_DTYPE_F32 ? 4 : 2;
  avc::GatherBwdArgs g;
  memset(&g, 0, sizeof(g));
  g.da = static_cast<const uint8_t*>(da);
  g.a_row_stride = a_row_stride;
  g.col_off = col_off;
  g.dst = static_cast<uint8_t*>(const_cast<void*>(feat->ptr));
  g.batch_stride = feat->batch_stride * es;
  g.frame_stride = feat->frame_stride * es;
  g.batch = batch;
  g.frames = feat->frames;
  g.dim = feat->dim;
  g.k = feat->stack;
  g.rep = feat->repeat > 0 ? feat->repeat : 1;
  g.valid = feat->valid_frames;
  g.tok_offset = tok_offset;
  g.tokens_per_sample = tokens_per_sample;
  cudaError_t e = avc::launch_gather_bwd(g, out_dtype == AVC_DTYPE_F32, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "gather_bwd launch");
  return AVC_OK;
}

static int fill_splice(const avc_splice* s, avc::SpliceArgs* k) {
  if (s == nullptr || s->input_ids == nullptr) return fail(AVC_ERR_INVALID, "splice: null input_ids");
  if (s->hidden <= 0 || s->hidden % 8 != 0) return fail(AVC_ERR_INVALID, "splice: hidden must be a positive multiple of 8");
  if (s->batch <= 0 || s->seq <= 0) return fail(AVC_ERR_INVALID, "splice: empty batch / seq");
  if (s->tok_offset == nullptr && s->tokens_per_sample < 0)
    return fail(AVC_ERR_INVALID, "splice: negative tokens_per_sample");
  memset(k, 0, sizeof(*k));
  k->input_ids = s->input_ids;
  k->placeholder_id = s->placeholder_id;
  k->pad_id = s->pad_id;
  k->batch = s->batch;
  k->seq = s->seq;
  if (s->elem_size != 0 && s->elem_size != 2 && s->elem_size != 4)
    return fail(AVC_ERR_INVALID, "splice: elem_size must be 2 (bf16) or 4 (fp32)");
  k->row_bytes = s->hidden * (s->elem_size == 4 ? 4 : 2);
  k->tok_offset = s->tok_offset;
  k->tokens_per_sample = s->tokens_per_sample;
  k->embed_table = static_cast<const uint8_t*>(s->embed_table);
  k->vocab = s->vocab;
  k->attention_mask = s->attention_mask;
  k->mask_mode = s->mask_mode;
  k->labels_in = s->labels_in;
  k->label_len = s->label_len;
  k->labels_out = s->labels_out;
  k->label_mode = s->label_mode;
  k->status = s->status;
  k->av_in_place = s->av_rows_in_place;
  return AVC_OK;
}

int avc_splice_fwd(const avc_splice* s, const void* y, void* inputs_embeds, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  avc::SpliceArgs k;
  if (int rc = fill_splice(s, &k)) return rc;
  if (inputs_embeds == nullptr) return fail(AVC_ERR_INVALID, "splice_fwd: null inputs_embeds");
  k.y = static_cast<const uint8_t*>(y);
  k.inputs_embeds = static_cast<uint8_t*>(inputs_embeds);
  cudaError_t e = avc::launch_splice_fwd(k, di.num_sms, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "splice_fwd launch");
  return AVC_OK;
}

int avc_splice_bwd(const avc_splice* s, const void* d_inputs_embeds, void* dy, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  avc::SpliceArgs k;
  if (int rc = fill_splice(s, &k)) return rc;
  if (d_inputs_embeds == nullptr || dy == nullptr) return fail(AVC_ERR_INVALID, "splice_bwd: null pointer");
  k.inputs_embeds = const_cast<uint8_t*>(static_cast<const uint8_t*>(d_inputs_embeds));
  k.dy = static_cast<uint8_t*>(dy);
  k.attention_mask = nullptr;
  k.labels_out = nullptr;
  cudaError_t e = avc::launch_splice_bwd(k, di.num_sms, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "splice_bwd launch");
  return AVC_OK;
}

int avc_row_resample(const void* x, void* out, int32_t dtype, int32_t batch, int32_t src_rows, int32_t dst_rows,
                     int32_t hidden, const int32_t* row_ptr, const int32_t* col_idx, const float* weight,
                     void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (x == nullptr || out == nullptr || row_ptr == nullptr || col_idx == nullptr || weight == nullptr)
    return fail(AVC_ERR_INVALID, "row_resample: null pointer");
  if (dtype != AVC_DTYPE_BF16 && dtype != AVC_DTYPE_F32 && dtype != AVC_DTYPE_F16)
    return fail(AVC_ERR_INVALID, "row_resample: dtype must be an AVC_DTYPE_* code");
  if (batch < 0 || src_rows <= 0 || dst_rows < 0 || hidden <= 0) return fail(AVC_ERR_INVALID, "row_resample: bad extents");
  avc::ResampleArgs r;
  memset(&r, 0, sizeof(r));
  r.x = static_cast<const uint8_t*>(x);
  r.out = static_cast<uint8_t*>(out);
  r.dtype = dtype;
  r.batch = batch;
  r.src_rows = src_rows;
  r.dst_rows = dst_rows;
  r.hidden = hidden;
  r.row_ptr = row_ptr;
  r.col = col_idx;
  r.weight = weight;
  cudaError_t e = avc::launch_row_resample(r, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "row_resample launch");
  return AVC_OK;
}

static int run_gelu(const void* dh, int64_t dh_ld, const void* z, int64_t z_ld, void* out, int64_t out_ld, int64_t rows,
                    int64_t cols, const uint8_t* row_flags, int32_t flag_bit, bool bwd, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (z == nullptr || out == nullptr || (bwd && dh == nullptr)) return fail(AVC_ERR_INVALID, "gelu: null pointer");
  avc::GeluArgs g;
  memset(&g, 0, sizeof(g));
  g.z = static_cast<const uint8_t*>(z);
  g.dh = static_cast<const uint8_t*>(dh);
  g.out = static_cast<uint8_t*>(out);
  g.rows = rows; g.cols = cols; g.z_ld = z_ld; g.dh_ld = dh_ld; g.out_ld = out_ld;
  g.row_flags = row_flags;
  g.flag_bit = flag_bit;
  cudaError_t e = avc::launch_gelu(g, bwd, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, bwd ? "gelu_bwd launch" : "gelu_fwd launch");
  return AVC_OK;
}

int avc_gelu_fwd(const void* z, int64_t z_ld, void* out, int64_t out_ld, int64_t rows, int64_t cols,
                 const uint8_t* row_flags, int32_t flag_bit, void* stream) {
  return run_gelu(nullptr, 0, z, z_ld, out, out_ld, rows, cols, row_flags, flag_bit, false, stream);
}

int avc_gelu_bwd(const void* dh, int64_t dh_ld, const void* z, int64_t z_ld, void* out, int64_t out_ld, int64_t rows,
                 int64_t cols, const uint8_t* row_flags, int32_t flag_bit, void* stream) {
  return run_gelu(dh, dh_ld, z, z_ld, out, out_ld, rows, cols, row_flags, flag_bit, true, stream);
}

int avc_pack_weight_t(const float* src, int64_t src_ld, void* dst_bf16, int64_t dst_ld, int64_t rows, int64_t cols,
                      float alpha, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (src == nullptr || dst_bf16 == nullptr) return fail(AVC_ERR_INVALID, "pack_weight_t: null pointer");
  cudaError_t e = avc::launch_pack_weight_t(src, src_ld, dst_bf16, dst_ld, rows, cols, alpha,
                                            static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "pack_weight_t launch");
  return AVC_OK;
}

size_t avc_sumsq_workspace_bytes(void) { return avc::sumsq_workspace_bytes(); }

int avc_sumsq(const float* x, int64_t n, float* out, void* workspace, int32_t accumulate, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (x == nullptr || out == nullptr || workspace == nullptr || n < 0) return fail(AVC_ERR_INVALID, "sumsq: bad argument");
  cudaError_t e = avc::launch_sumsq(x, n, out, static_cast<float*>(workspace), accumulate,
                                    static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "sumsq launch");
  return AVC_OK;
}

int avc_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t rows, int64_t cols,
                   float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                   const float* grad_scale, const float* clip_sumsq, float max_norm, void* packed_bf16,
                   int64_t packed_ld, float packed_alpha, void* stream) {
  DeviceInfo di;
  if (int rc = device_info(&di)) return rc;
  if (param == nullptr || grad == nullptr || exp_avg == nullptr || exp_avg_sq == nullptr)
    return fail(AVC_ERR_INVALID, "adamw_step: null pointer");
  if (step < 1) return fail(AVC_ERR_INVALID, "adamw_step: step counts from 1");
  avc::AdamWArgs a;
  memset(&a, 0, sizeof(a));
  a.param = param;
  a.grad = grad;
  a.exp_avg = exp_avg;
  a.exp_avg_sq = exp_avg_sq;
  a.rows = rows;
  a.cols = cols;
  // scalar prefactors in double, as torch's _single_tensor_adamw computes them on the host
  const double bc1 = 1.0 - pow(static_cast<double>(beta1), step);
  const double bc2 = 1.0 - pow(static_cast<double>(beta2), step);
  a.decay = static_cast<float>(1.0 - static_cast<double>(lr) * weight_decay);
  a.one_minus_beta1 = static_cast<float>(1.0 - static_cast<double>(beta1));
  a.beta2 = beta2;
  a.one_minus_beta2 = static_cast<float>(1.0 - static_cast<double>(beta2));
  a.step_size = static_cast<float>(static_cast<double>(lr) / bc1);
  a.bias_correction2_sqrt = static_cast<float>(sqrt(bc2));
  a.eps = eps;
  a.grad_scale = grad_scale;
  a.clip_sumsq = clip_sumsq;
  a.max_norm = max_norm;
  a.packed = static_cast<uint8_t*>(packed_bf16);
  a.packed_ld = packed_ld;
  a.packed_alpha = packed_alpha;
  cudaError_t e = avc::launch_adamw(a, static_cast<cudaStream_t>(stream));
  if (e != cudaSuccess) return cuda_fail(e, "adamw_step launch");
  return AVC_OK;
}

}  // extern "C"
