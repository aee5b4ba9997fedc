/* avconnector_b200.h -- C ABI of the B200-native clip_whisper multimodal connector.
 *
 * The reference (rishabhjain16/audio-visual-llm) has no FFI: its connector is plain PyTorch
 * (src/clip_whisper/models/modality_connector.py, clip_whisper_model.py).  This header is the
 * boundary a maintainer would bind instead of those eager ops (INTEGRATION.md shows the ctypes stub).
 * Each entry point names the reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; nothing is allocated or freed
 *     (except by the avc_comm_alloc / avc_comm_free peer-memory helpers)
 *   - GEMM operands (features, packed weights, dY) are bf16; projected rows / inputs_embeds are bf16, fp16 or fp32
 *     (AVC_DTYPE_*); masks, ids and labels are int64 (reference dtype)
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it
 *   - return 0 on success, non-zero on error; avc_last_error() gives the message (thread-local)
 *   - there is no CPU fallback: a device that is not sm_100 is an error
 */
#ifndef AVCONNECTOR_B200_H_
#define AVCONNECTOR_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define AVC_API __attribute__((visibility("default")))
#else
#define AVC_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define AVC_ABI_VERSION 3  /* 2: colsum workspace header contract, avc_comm_* / avc_mc_* data-parallel entry points
                              3: output dtype codes (fp16), db inside the dW launch (avc_proj_bwd_dw_db*),
                                 avc_proj_bwd_dx, avc_cast_bf16, avc_gather_bwd */

/* element types of matrices the kernels write (operands of the tensor-core GEMMs are always bf16) */
#define AVC_DTYPE_BF16 0
#define AVC_DTYPE_F32 1
#define AVC_DTYPE_F16 2   /* the reference's use_fp16 mode: fp16 LLM / connector output (clip_whisper_model.py:164) */

enum avc_status {
  AVC_OK = 0,
  AVC_ERR_INVALID = 1,      /* bad argument (shape, alignment, null) */
  AVC_ERR_CUDA = 2,         /* CUDA runtime / driver error */
  AVC_ERR_UNSUPPORTED = 3   /* not an sm_100 device, or a size outside kernel limits */
};

AVC_API int avc_abi_version(void);
AVC_API const char* avc_last_error(void);
/* AVC_OK iff `device` is a compute-capability 10.x GPU.  Replaces the reference's
 * `"cuda" if torch.cuda.is_available() else "cpu"` fallback (clip_whisper_model.py:91). */
AVC_API int avc_device_check(int device);

/* One modality's tower output: bf16 [batch, frames, dim]. */
typedef struct avc_feat {
  const void* ptr;             /* NULL: modality absent */
  int64_t batch_stride;        /* elements between samples */
  int64_t frame_stride;        /* elements between frames: dim if dense; (1+Np)*dim selects the CLS row
                                  of a CLIP last_hidden_state (clip_whisper_model.py:1141-1142) */
  int32_t frames;              /* T */
  int32_t dim;                 /* D, multiple of 8 */
  int32_t stack;               /* k >= 1 frames stacked per token (k = 1: reference behaviour) */
  int32_t repeat;              /* r >= 1 (0 = 1): token j reads the stack starting at frame (j / r) * k, i.e. every
                                  stacked token is used r times -- pairs a slower stream with a faster one
                                  (25 fps video against 50 Hz audio at stride 1: r = 2) */
  const int32_t* valid_frames; /* [batch] per-sample valid frame count, or NULL (= frames) */
} avc_feat;

/* A strided bf16 / fp32 matrix batch: [batches][rows][cols], cols contiguous. */
typedef struct avc_mat {
  void* ptr;
  int64_t rows;          /* rows per batch entry */
  int64_t cols;
  int64_t row_stride;    /* elements */
  int64_t batches;       /* 1 for a plain matrix */
  int64_t batch_stride;  /* elements */
} avc_mat;

/* ---- gather: temporal align + stride-k stack + concat (HBM-bound, TMA bulk copies) ---------------
 * A[m, :] = [audio[b, ka*ja .. ka*ja+ka-1, :] ; video[b, kv*jv .. kv*jv+kv-1, :]], ja = j / ra, jv = j / rv,
 * m = tok_offset[b] + j,
 * frames past the valid length are zero.  row_flags[m] bit0/bit1 = row has an audio / video token.
 * Replaces _pad_or_truncate + index alignment (clip_whisper_model.py:320-374, 424-431), moved in
 * front of the projection (exact: padding rows are zero and their bias is masked via row_flags). */
AVC_API int avc_gather_fwd(const avc_feat* audio, const avc_feat* video, int32_t batch,
                   const int32_t* tok_offset /* [batch+1] or NULL */, int32_t tokens_per_sample,
                   int64_t total_rows, void* a_out, int64_t a_row_stride, uint8_t* row_flags,
                   void* stream);

/* ---- projector forward (tensor-core bound, tcgen05 + TMEM + TMA) -------------------------------
 * Y[b, r, :] = act( sum_s A_s[b, r, :] . W_s^T  + flag0(b,r) * bias_scale0 * bias0 + flag1(b,r) * bias_scale1 * bias1 )
 * flags come from row_flags (packed rows) or, if NULL, flag_i = (r < flag_rows_i).
 * Scatter output: when every A_s is packed ([B*N, K_s], batches = 1) and y has batches = B, rows = N, packed row
 * m = b*N + j is written to y[b, j, :] -- y may be the AV region of inputs_embeds (ptr = &emb[0, P, 0],
 * batch_stride = S*H), which fuses the splice of the projected rows into the GEMM epilogue.
 * Replaces SimpleModalityConnector.forward x2 + weighted sum (modality_connector.py:43-44,
 * clip_whisper_model.py:434) through W = [fs*Wa | (1-fs)*Wv], bias0 = fs*ba, bias1 = (1-fs)*bv. */
AVC_API int avc_proj_fwd(int32_t nseg, const avc_mat* a /* [nseg] bf16 */, const avc_mat* w /* [nseg] bf16 [N, K_s] */,
                 const avc_mat* y, int32_t y_dtype /* AVC_DTYPE_* */, const float* bias0, const float* bias1,
                 float bias_scale0, float bias_scale1, const uint8_t* row_flags, int32_t flag_rows0,
                 int32_t flag_rows1, int32_t act /* 0 none, 1 GELU(erf) */, void* stream);

/* Debug / profiling hook: when `device_buf` (3 x 148 x 8 uint64, zero it first) is non-NULL every later projector GEMM
 * launch records per-CTA SM-cycle counters into it: [cta][0] producer waiting for a free smem stage, [1] MMA issuer
 * waiting for operands, [2] MMA issuer waiting for a free accumulator, [3] first epilogue warp waiting for a full
 * accumulator, [4] its epilogue time, [5] its work items, [6] producer loop cycles; the second block holds the epilogue's
 * per-phase cycles, the third %globaltimer stamps (kernel start, epilogue / comm-warp milestones, end).  NULL: off. */
AVC_API int avc_debug_gemm_profile(void* device_buf);

/* ---- projector backward: weight gradient ------------------------------------------------------
 * dW_s[h, k] = alpha_s * sum_{b, r < x.rows} dY[b, dy_row_base + r, h] * X_s[b, r, k]   (fp32 out)
 * Replaces autograd of nn.Linear (dW = dY^T X); dX is not produced (towers are frozen,
 * clip_whisper_model.py:1096,1136). */
AVC_API int avc_proj_bwd_dw(const avc_mat* dy /* bf16 [b][rows][H] */, int32_t dy_row_base, int32_t nseg,
                    const avc_mat* x /* [nseg] bf16 */, const avc_mat* dw /* [nseg] fp32 [H, K_s] */,
                    const float* alpha /* [nseg] */, int32_t max_sms /* 0 = all; < SM count leaves SMs free for a
                    concurrent collective (gradient all-reduce overlap) */, void* stream);

/* ---- projector backward: weight AND bias gradients in one launch ------------------------------------------------
 * avc_proj_bwd_dw plus, from extra 64-wide work items of the same persistent tile schedule (they re-use the dY panels
 * the weight tiles stream and fill the idle workers of the schedule's last round):
 *   out_i[h] = alpha_i * sum_{b, r} dY[b, dy_row_base + r, h] * present[b, r, i]          i = 0 (audio), 1 (video)
 * `present` is a bf16 [batches][rows >= x rows][64] matrix whose column i is 1 where row r of sample b carries a
 * token of stream i and 0 elsewhere (all ones for dense streams; the zero-padded rows of the shorter stream are what
 * clip_whisper_model.py:340-345 pads AFTER the projection, so they must not see that stream's bias).  Columns 2..63
 * are ignored.  Replaces autograd's db = sum dY of nn.Linear (modality_connector.py:32); deterministic.
 *
 * Split reduction.  When the launch has fewer work items than the GPU has CTA pairs (short K_s: 8 x 5 tiles at 1280 ->
 * 4096), every item is cut into S <= 8 slices of the row reduction that run on the idle pairs; slices store partial tiles
 * in `workspace` and the last slice to finish a 32-row slab adds the partials IN SLICE ORDER (deterministic) into dW /
 * db.  avc_proj_bwd_dw_plan reports S and the bytes needed for the given operands; zero-fill the workspace once after
 * allocating it (256-byte aligned; every launch leaves its arrival counters zero).  workspace == NULL, or bias == NULL
 * for dW alone, are allowed; a workspace that is too small lowers S. */
typedef struct avc_bias_grad {
  const avc_mat* present;  /* bf16 [batches][rows][64] */
  float* out0;             /* [H] fp32 or NULL */
  float* out1;             /* [H] fp32 or NULL */
  float alpha0, alpha1;
} avc_bias_grad;
AVC_API int avc_proj_bwd_dw_db(const avc_mat* dy, int32_t dy_row_base, int32_t nseg, const avc_mat* x,
                               const avc_mat* dw, const float* alpha, const avc_bias_grad* bias /* or NULL */,
                               void* workspace /* or NULL */, size_t workspace_bytes, int32_t max_sms, void* stream);
AVC_API int avc_proj_bwd_dw_plan(const avc_mat* dy, int32_t nseg, const avc_mat* x, int32_t with_bias,
                                 int32_t* splits, size_t* workspace_bytes);

/* ---- projector backward: input gradient (unfrozen towers, freeze_encoders=False: clip_whisper_model.py:1096,1136)
 * dX[b, r, :] = sum_s dY[b, r, :] . W_s   computed as the forward GEMM with B = W_s^T from avc_pack_weight_t
 * (w_t[s]: bf16 [K_in, H_s]); dx is bf16 / fp16 / fp32 (dx_dtype = AVC_DTYPE_*). */
AVC_API int avc_proj_bwd_dx(const avc_mat* dy /* [nseg] bf16 */, int32_t nseg, const avc_mat* w_t /* [nseg] */,
                            const avc_mat* dx, int32_t dx_dtype, void* stream);

/* ---- gather backward: input gradients of the align / stack / concat step ------------------------------------------
 * d_feat[b, t, :] = sum over the tokens j that read frame t (j / repeat == t / stack, j < tokens of sample b) of
 *                   dA[tok_offset[b] + j, col_off + (t % stack) * dim : + dim],   zero for frames no token reads
 * (frames at or past valid_frames[b], or past the token cap).  `feat` describes the OUTPUT d_feat here (ptr, strides,
 * frames, dim, stack, repeat, valid_frames as in avc_gather_fwd); out_dtype = AVC_DTYPE_BF16 or AVC_DTYPE_F32. */
AVC_API int avc_gather_bwd(const void* da /* bf16 [total_rows, a_row_stride] */, int64_t a_row_stride, int64_t col_off,
                           const avc_feat* feat, int32_t out_dtype, int32_t batch, const int32_t* tok_offset,
                           int32_t tokens_per_sample, void* stream);

/* ---- data parallel: the same weight-gradient GEMM with the gradient all-reduce fused into it ------------------
 * One process per GPU; every rank keeps its projector gradients in ONE flat fp32 bucket allocated with
 * avc_comm_alloc and mapped into every other rank's process (avc_comm_export / avc_comm_open: CUDA IPC, peer access
 * over NVLink).  avc_proj_bwd_dw_allreduce computes this rank's dW tiles into its bucket exactly like
 * avc_proj_bwd_dw and, IN THE SAME KERNEL, extra warps of the GEMM CTAs sum every finished tile over the ranks
 * (peer loads, fixed rank order: deterministic and bit-identical on every rank) and store the sum into every rank's
 * bucket (peer stores) while the tensor cores work on later tiles.  Up to two extra flat ranges of the bucket (the
 * bias gradients, produced by avc_colsum on another stream) are reduced by the same launch once every rank has
 * called avc_comm_signal_extra for this epoch.  When the launch completes the local bucket holds the SUM over ranks
 * (fold 1 / world into alpha).  Every rank must issue the same sequence of calls with the same `epoch` (strictly
 * increasing from 1).  The reference has no distributed code; the insertion point is between loss.backward() and
 * clip_grad_norm_ (clip_whisper_trainer.py:454-458).
 * A peer that does not show up within timeout_ns sets *status to 1 (and the kernel finishes) instead of hanging. */
#define AVC_COMM_MAX_WORLD 8
typedef struct avc_comm {
  int32_t world;                     /* 1 .. AVC_COMM_MAX_WORLD */
  int32_t rank;
  uint32_t epoch;
  uint32_t reserved;
  void* bucket[AVC_COMM_MAX_WORLD];  /* base of every rank's gradient bucket as mapped in THIS process */
  void* flags[AVC_COMM_MAX_WORLD];   /* every rank's flag area: avc_comm_flag_bytes() bytes from avc_comm_alloc */
  int32_t* status;                   /* local device int32, zero-initialised */
  uint64_t timeout_ns;               /* 0 = 20 s */
  uint64_t bucket_bytes;             /* size of every rank's bucket (ranges passed to the calls are checked against it) */
  void* mc_bucket;                   /* optional (NULL: peer loads / stores): multicast address of all ranks' buckets from
                                        avc_mc_bucket_alloc; the launch then reduces with multimem.ld_reduce (summed in
                                        the NVSwitch) and multimem.st, and bucket[p] is only needed for p == rank */
} avc_comm;
AVC_API size_t avc_comm_flag_bytes(void);
/* cudaMalloc + zero-fill on the current device (IPC-exportable, unlike pooled allocator blocks). */
AVC_API int avc_comm_alloc(size_t bytes, void** ptr);
AVC_API int avc_comm_free(void* ptr);
/* 64-byte cudaIpcMemHandle_t of an avc_comm_alloc block, to be sent to the other ranks' processes ... */
AVC_API int avc_comm_export(const void* ptr, void* handle64);
/* ... which map it (enabling peer access from the current device) and unmap it again. */
AVC_API int avc_comm_open(const void* handle64, void** ptr);
AVC_API int avc_comm_close(void* ptr);
/* dw[s].ptr, extra0, extra1 must lie inside comm->bucket[comm->rank]; dw[s] must be contiguous (row_stride == cols). */
AVC_API int avc_proj_bwd_dw_allreduce(const avc_mat* dy, int32_t dy_row_base, int32_t nseg, const avc_mat* x,
                                      const avc_mat* dw, const float* alpha, const avc_comm* comm,
                                      const float* extra0, int64_t extra0_len, const float* extra1,
                                      int64_t extra1_len, int32_t max_sms, void* stream);
/* The same with the bias gradients produced by the launch itself (avc_proj_bwd_dw_db): bias->out0 / out1 are the
 * extra ranges (inside the local bucket), flagged ready by the epilogue of the last bias work item. */
AVC_API int avc_proj_bwd_dw_db_allreduce(const avc_mat* dy, int32_t dy_row_base, int32_t nseg, const avc_mat* x,
                                         const avc_mat* dw, const float* alpha, const avc_bias_grad* bias,
                                         const avc_comm* comm, int32_t max_sms, void* stream);
/* Enqueue after the kernel that wrote this rank's extra ranges (same stream): flags them ready for comm->epoch. */
AVC_API int avc_comm_signal_extra(const avc_comm* comm, int64_t extra0_len, int64_t extra1_len, void* stream);

/* ---- multicast bucket (NVSwitch "NVLS"): optional transport of the fused all-reduce --------------------------------
 * Every rank's gradient bucket is bound at the same offset of ONE multicast object, so that a multimem.ld_reduce on
 * mc_ptr + x returns the sum of all ranks' words at x (added inside the switch) and a multimem.st writes all of them.
 * Setup (one process per GPU; the fd travels between processes over an AF_UNIX socket with SCM_RIGHTS):
 *   all ranks   avc_mc_padded_bytes(world, bytes) -> size rounded to the multicast granularity (512 MiB on B200)
 *   rank 0      avc_mc_create -> object handle + POSIX fd          other ranks  avc_mc_import(fd)
 *   all ranks   avc_mc_add_device (current device); BARRIER; avc_mc_bucket_alloc (cuMemCreate + bind + map, zero-filled);
 *               BARRIER before first use.  avc_comm.bucket[rank] = ptr, avc_comm.mc_bucket = mc_ptr.
 * avc_mc_supported: 1 iff the device and driver support multicast objects (NVSwitch systems). */
typedef struct avc_mc_bucket {
  void* ptr;             /* this rank's bucket (ordinary device address) */
  void* mc_ptr;          /* multicast address of every rank's bucket (multimem.* instructions only) */
  uint64_t bytes;        /* padded size */
  uint64_t mem_handle;   /* driver handles, released by avc_mc_bucket_free */
  uint64_t mc_handle;
} avc_mc_bucket;
AVC_API int avc_mc_supported(int32_t device, int32_t* supported);
AVC_API int avc_mc_padded_bytes(int32_t world, uint64_t min_bytes, uint64_t* padded_bytes);
AVC_API int avc_mc_create(int32_t world, uint64_t padded_bytes, uint64_t* mc_handle, int32_t* fd);
AVC_API int avc_mc_import(int32_t fd, uint64_t* mc_handle);
AVC_API int avc_mc_add_device(uint64_t mc_handle);
AVC_API int avc_mc_bucket_alloc(uint64_t mc_handle, uint64_t padded_bytes, avc_mc_bucket* out);
AVC_API int avc_mc_bucket_free(avc_mc_bucket* b);

/* ---- bias gradient: deterministic column sum over flagged rows (one launch, fixed summation tree) -------------
 * out_i[c] = alpha_i * sum_{b, r < sum_rows : flag_i(b, r)} dY[b, dy_row_base + r, c]      (db = sum dY)
 * flags: row_flags[b * sum_rows + r] bit i, or (r < flag_rows_i) when row_flags is NULL.
 * workspace: avc_colsum_workspace_bytes(cols) bytes, 16-byte aligned; its first avc_colsum_workspace_header_bytes()
 * bytes (arrival counters) must be ZERO before the first call -- every call leaves them zero again.  Calls that share
 * a workspace must be ordered on one stream.  The kernel uses no shared memory, so it can run next to the projector
 * GEMM on another stream.
 * avc_colsum_comm: the same, and when the sums are final it flags them ready for comm->epoch's fused all-reduce
 * (what avc_comm_signal_extra would do after it), out0 / out1 being the extra ranges of that launch. */
AVC_API size_t avc_colsum_workspace_bytes(int32_t cols);
AVC_API size_t avc_colsum_workspace_header_bytes(void);
AVC_API int avc_colsum(const avc_mat* dy /* bf16 */, int32_t dy_row_base, int32_t sum_rows,
               const uint8_t* row_flags, int32_t flag_rows0, int32_t flag_rows1, float alpha0,
               float alpha1, float* out0, float* out1, void* workspace, void* stream);
AVC_API int avc_colsum_comm(const avc_mat* dy /* bf16 */, int32_t dy_row_base, int32_t sum_rows,
               const uint8_t* row_flags, int32_t flag_rows0, int32_t flag_rows1, float alpha0,
               float alpha1, float* out0, float* out1, void* workspace, const avc_comm* comm, void* stream);

/* ---- weight pack: W_bf16 = bf16(alpha * W_fp32) (folds fusion_scale into the projector) -------- */
AVC_API int avc_pack_weight(const float* src, int64_t src_ld, void* dst_bf16, int64_t dst_ld, int64_t rows,
                    int64_t cols, float alpha, void* stream);
/* ---- operand cast: dst_bf16[r, c] = bf16(alpha * src[r, c]) for src of type AVC_DTYPE_* (the base class's cast of the
 * connector input to the module dtype, modality_connector.py:18-19; also fp16 / fp32 dY -> the bf16 GEMM operand). */
AVC_API int avc_cast_bf16(const void* src, int32_t src_dtype, int64_t src_ld, void* dst_bf16, int64_t dst_ld,
                          int64_t rows, int64_t cols, float alpha, void* stream);

/* ---- splice: scatter projected rows + text embeddings into inputs_embeds, emit masks ------------
 * inputs_embeds[b, p] = Y[tok_offset[b] + rank(p)]  where input_ids[b, p] == placeholder_id
 *                     = embed_table[input_ids[b, p]] elsewhere (zeros if no table / id out of range)
 * attention_mask (int64): mask_mode 0 -> all ones (clip_whisper_model.py:460); 1 -> valid tokens only
 * labels_out (int64):  label_mode 0 -> reference eval rule: pad -> -100, truncate / right-pad -100
 *                      (clip_whisper_model.py:569-570, 586-598); 1 -> also -100 on placeholders
 * *status (device int32) gets bit0 if a sample's placeholder count != its token count.
 * Replaces _embed_prompt + torch.cat + torch.ones (clip_whisper_model.py:448-451, 460, 464-487). */
typedef struct avc_splice {
  const int64_t* input_ids;   /* [batch, seq] */
  int64_t placeholder_id;
  int64_t pad_id;
  int32_t batch;
  int32_t seq;
  int32_t hidden;             /* H, multiple of 8 */
  int32_t tokens_per_sample;  /* used when tok_offset is NULL */
  const int32_t* tok_offset;  /* [batch+1] or NULL */
  const void* embed_table;    /* [vocab, H] (elem_size bytes per element) or NULL */
  int64_t vocab;
  int64_t* attention_mask;    /* [batch, seq] or NULL */
  int32_t mask_mode;
  int32_t label_mode;
  const int64_t* labels_in;   /* [batch, label_len] or NULL */
  int32_t label_len;
  int32_t elem_size;          /* bytes per element of y / embed_table / inputs_embeds: 2 (bf16, fp16) or 4 (fp32); 0 = 2 */
  int64_t* labels_out;        /* [batch, seq] or NULL */
  int32_t* status;            /* device int32 or NULL */
  int32_t av_rows_in_place;   /* fwd: 1 = the placeholder rows of inputs_embeds were already written by
                                 avc_proj_fwd (scatter output): only text rows and masks are produced */
  int32_t reserved;
} avc_splice;

AVC_API int avc_splice_fwd(const avc_splice* s, const void* y /* bf16 [M, H] */, void* inputs_embeds,
                   void* stream);
/* dY[tok_offset[b] + rank(p)] = d_inputs_embeds[b, p] at placeholder positions (autograd of the cat). */
AVC_API int avc_splice_bwd(const avc_splice* s, const void* d_inputs_embeds, void* dy /* bf16 [M, H] */,
                   void* stream);

/* ---- row resample: out[b, i, :] = sum_{t in CSR row i} weight[t] * x[b, col_idx[t], :] ------------
 * With the CSR matrix of adaptive average pooling (windows [floor(i*S/L), ceil((i+1)*S/L)), weight 1/len) or
 * of linear interpolation with align_corners (two taps) this replaces _adaptive_projection
 * (clip_whisper_model.py:621-676); with the transposed matrix it is that op's backward.
 * dtype: AVC_DTYPE_* of x and out (accumulation is fp32 either way). */
AVC_API int avc_row_resample(const void* x, void* out, int32_t dtype, int32_t batch, int32_t src_rows,
                             int32_t dst_rows, int32_t hidden, const int32_t* row_ptr, const int32_t* col_idx,
                             const float* weight, void* stream);

/* ---- MLP projector helpers (Linear -> GELU -> Linear; north_star extension, erf GELU as modality_connector.py:60)
 * avc_gelu_fwd:  out[r, :] = flag(r) ? gelu(z[r, :]) : 0          avc_gelu_bwd: out[r, :] = flag(r) ? dh * gelu'(z) : 0
 *   bf16 matrices (rows x cols, leading dimensions in elements); flag(r) = row_flags[r] & flag_bit (all on if NULL).
 *   Forward-only inference can use the GELU epilogue of avc_proj_fwd (act = 1) instead.
 * avc_pack_weight_t: dst_bf16[c, r] = bf16(alpha * src_f32[r, c]) -- the transposed operand of the input-gradient
 *   GEMM  dX = dY . W  (avc_proj_fwd with A = dY, W = this). */
AVC_API int avc_gelu_fwd(const void* z, int64_t z_ld, void* out, int64_t out_ld, int64_t rows, int64_t cols,
                         const uint8_t* row_flags, int32_t flag_bit, void* stream);
AVC_API int avc_gelu_bwd(const void* dh, int64_t dh_ld, const void* z, int64_t z_ld, void* out, int64_t out_ld,
                         int64_t rows, int64_t cols, const uint8_t* row_flags, int32_t flag_bit, void* stream);
AVC_API int avc_pack_weight_t(const float* src, int64_t src_ld, void* dst_bf16, int64_t dst_ld, int64_t rows,
                              int64_t cols, float alpha, void* stream);

/* ---- trainer step for the projector parameters ----------------------------------------------------
 * Replaces, for the connector parameters, clip_grad_norm_ + AdamW.step of the reference trainer
 * (clip_whisper_trainer.py:171-207 param groups / AdamW(betas=(0.9, 0.95), eps=1e-8), :453-464 step).
 * avc_sumsq: *out (+)= sum x[i]^2, deterministic (fixed grid and reduction order); the caller adds the other
 *   parameters' contribution and turns the global norm into a clip coefficient kept ON THE DEVICE.
 * avc_adamw_step: torch.optim.AdamW semantics on one [rows, cols] fp32 tensor; grad is multiplied by *grad_scale
 *   first, and by min(1, max_norm / (sqrt(*clip_sumsq) + 1e-6)) when clip_sumsq != NULL and max_norm > 0
 *   (torch.nn.utils.clip_grad_norm_ with the squared global norm kept on the device); optionally writes bf16(packed_alpha * param) into the packed projector operand (row stride
 *   packed_ld elements), which replaces the per-step avc_pack_weight. */
AVC_API size_t avc_sumsq_workspace_bytes(void);
AVC_API int avc_sumsq(const float* x, int64_t n, float* out, void* workspace, int32_t accumulate, void* stream);
AVC_API int avc_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t rows,
                           int64_t cols, float lr, float beta1, float beta2, float eps, float weight_decay,
                           int32_t step, const float* grad_scale, const float* clip_sumsq, float max_norm,
                           void* packed_bf16, int64_t packed_ld, float packed_alpha, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AVCONNECTOR_B200_H_ */
