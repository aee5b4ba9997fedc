// Internal launch interface between the C-ABI layer (avc_capi.cu) and the sm_100a kernels.
#pragma once
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

namespace avc {

// ------------------------------------------------------------------ projector GEMM (tcgen05)
constexpr int GEMM_BM = 128;  // UMMA M (TMEM lanes)
constexpr int GEMM_BN = 256;  // UMMA N (TMEM columns per accumulator stage)
constexpr int GEMM_BK = 64;   // contraction elements per smem stage (= one 128-byte swizzle row)

enum GemmMode : int {
  GEMM_TN = 0,  // D[m,n] = sum_k A[m,k] * B[n,k]      (both operands K-major)   -> forward, dX = dY . W
  GEMM_NT = 1,  // D[m,n] = sum_r A[r,m] * B[r,n]      (both operands MN-major)  -> dW = dY^T * X
};
// element type the epilogue writes (the accumulator is always fp32)
enum GemmOut : int {
  GEMM_OUT_BF16 = 0,
  GEMM_OUT_F32 = 1,
  GEMM_OUT_F16 = 2,  // the reference's use_fp16 mode (clip_whisper_model.py:164)
};
constexpr int GEMM_BIAS_COLS = 64;  // width of the bias-gradient work items (one MN-major swizzle atom)

// ------------------------------------------------------------------ fused dW GEMM + gradient all-reduce over peer memory
// Data-parallel training: every rank's dW GEMM writes its partial weight gradient into its own copy of the flat
// gradient bucket; the buckets of all ranks are mapped into every process (CUDA IPC over NVLink).  Work item w of the
// persistent GEMM schedule (a tile or tail sub-tile, the same list on every rank) is owned by rank w % world.  Extra
// "comm" warps of the GEMM CTAs wait until all ranks have flagged item w, sum the `world` partial tiles in rank
// order with peer loads and store the result into every rank's bucket with peer stores -- while the tensor cores
// work on later tiles.  Flags are epoch numbers (one epoch per launch), so nothing is reset between steps.
constexpr int COMM_MAX_WORLD = 8;
constexpr int COMM_MAX_ITEMS = 2048;                       // work items per launch the flag area has room for
constexpr int COMM_ITEM_FLAGS = 0;                         // [item][src rank] epoch: src's partial of item is in its bucket
constexpr int COMM_EXTRA_FLAGS = COMM_MAX_ITEMS * COMM_MAX_WORLD;          // [src rank]: src's extra ranges (bias sums) are ready
constexpr int COMM_DONE_FLAGS = COMM_EXTRA_FLAGS + COMM_MAX_WORLD;         // [src rank]: src has broadcast everything it owns
constexpr int COMM_ITEM_COUNT = COMM_DONE_FLAGS + COMM_MAX_WORLD;          // local: epilogue-warp arrivals per item
constexpr int COMM_DONE_COUNT = COMM_ITEM_COUNT + COMM_MAX_ITEMS;          // local: comm warps that have finished
constexpr int COMM_EXTRA_COUNT = COMM_DONE_COUNT + 1;                      // local: epilogue warps of finished bias items
constexpr int COMM_FLAG_WORDS = COMM_DONE_COUNT + 32;
constexpr int COMM_ERR_TIMEOUT = 1;

struct CommArgs {
  int world;                        // 0: plain GEMM (no comm warps)
  int rank;
  uint32_t epoch;                   // same strictly increasing number on every rank, one per launch
  float* data[COMM_MAX_WORLD];      // gradient bucket of every rank as mapped in this process (data[rank] is local)
  float* mc;                        // optional: multicast address of all buckets (NVSwitch); then data[p != rank] is unused
                                    //   and the tiles are reduced with multimem.ld_reduce / multimem.st
  uint32_t* flags[COMM_MAX_WORLD];  // flag area (COMM_FLAG_WORDS words) of every rank
  int64_t seg_off[2];               // float offset of output segment s (dW_s, row-major, ld = d_cols[s]) in the bucket
  int64_t extra_off[2];             // up to two flat float ranges reduced as well once every rank flagged them
  int extra_len[2];                 //   (the bias gradients; multiples of 4 floats, 16-byte aligned)
  int32_t* status;                  // local device word: COMM_ERR_* on failure
  uint64_t timeout_ns;              // give up (status = timeout) instead of hanging when a peer never shows up
  uint32_t poll_ns;                 // set by launch_gemm: sleep between two polls of a flag
};

struct GemmArgs {
  CUtensorMap ma[2];  // TN: A operand per K segment.        NT: ma[0] = dY  ([batch][rows][m])
  CUtensorMap mb[2];  // TN: B operand (weights) per K seg.  NT: X per output segment ([batch][rows][n])
  CUtensorMap md[2];  // TN: md[0] = output.                 NT: output per segment
  CUtensorMap md_row; // TN scatter output: same tensor as md[0] with a one-row box (sample-straddling boxes)
  CUtensorMap mf;     // NT, bias_items > 0: token-present operand F [batch][rows][GEMM_BIAS_COLS] bf16,
                      //   F[b, r, i] = 1 if row r of sample b carries a token of stream i (i = 0 audio, 1 video), else 0
  int num_m_blocks;
  int num_n_blocks;
  int bn;  // N tile of this launch: 64 | 128 | 192 | 256 (B tensor-map box rows / atoms must match)
  int full_tiles;  // set by launch_gemm: tiles processed whole; the remaining ones are cut into
  int tail_split;  // `tail_split` sub-tiles of width bn / tail_split (tail of the persistent schedule)
  int group_m;     // set by launch_gemm: M blocks per rasterisation group (L2 reuse of the B panels)
  int l2_hints;    // set by launch_gemm: TMA L2 eviction hints on / off
  // TN
  int m_tiles_per_batch;  // an M tile never straddles two batch entries
  int nseg;               // K segments
  int seg_kblocks[2];
  // NT
  int red_batches;            // batches reduced over
  int red_kblocks_per_batch;  // ceil(rows_per_batch / GEMM_BK)
  int a_row_base;             // first dY row of the token run inside each batch entry
  int n_blocks_seg0;          // N tiles that belong to output segment 0
  // output extents (stores that start outside are skipped; partial boxes are clipped by TMA)
  int d_rows;     // TN: rows per batch entry.  NT: rows of D
  int d_cols[2];  // TN: d_cols[0] = N.         NT: columns of D per segment
  // epilogue
  const float* bias0;        // [N] fp32, added where flag bit0 is set (nullptr: none)
  const float* bias1;        // [N] fp32, added where flag bit1 is set (nullptr: none)
  const uint8_t* row_flags;  // packed-row mode: flags[batch * d_rows + row]; nullptr: analytic
  int flag_rows0;            // analytic: bit0 = row < flag_rows0
  int flag_rows1;            // analytic: bit1 = row < flag_rows1
  int scatter_rows;          // TN: > 0: A rows are packed [batches * scatter_rows] while the output map is
  int scatter_batches;       //     [scatter_batches][scatter_rows][N] (epilogue writes into a strided region)
  float alpha[2];            // NT: output scale per segment
  float bias_scale[2];       // TN: bias multipliers (fusion_scale folded into the bias)
  int act;                   // 0 = identity, 1 = GELU (erf form)
  // NT: bias gradients inside the same launch.  bias_items = num_m_blocks appends one GEMM_BIAS_COLS-wide work item
  // per M block that contracts the dY panel with F:  bias_out[i][h] = bias_alpha[i] * sum_{b, r} dY[b, r, h] * F[b, r, i]
  // (db of the nn.Linear bias with the pad-after-projection row mask; same fixed reduction order as dW: deterministic).
  // With comm.world >= 1 the epilogue of the last bias item flags the comm extras (the bias ranges) ready.
  int bias_items;
  float* bias_out[2];        // [d_rows] fp32 each, may be null
  float bias_alpha[2];
  // NT, few tiles: split the reduction of every work item into ksplit slices (see gemm_dw_splits).  Partial tiles go to
  // the workspace ws = [ksplit][d_rows][ws_ld] fp32 through mws[seg] ([ksplit][d_rows][d_cols[seg]] at column
  // ws_seg_col[seg]); partial bias sums to ws_bias = [ksplit][2][d_rows]; split_count = zeroed, self-resetting arrival
  // counters [(tiles + bias items)][8].  The final sums are written through d_ptr / d_ld (the dW segments).
  int ksplit;
  CUtensorMap mws[2];
  float* ws;
  int64_t ws_split_stride;   // elements between two slices of the workspace
  int ws_ld;                 // row pitch of the workspace (elements)
  int ws_seg_col[2];
  float* ws_bias;
  uint32_t* split_count;
  float* d_ptr[2];
  int64_t d_ld[2];
  CommArgs comm;             // NT only: world >= 1 fuses the gradient all-reduce into the launch (0: plain GEMM)
  // optional per-CTA cycle counters (debug, avc_debug_gemm_profile): [cta][8] = {producer wait-empty, MMA wait-full,
  // MMA wait-tempty, epilogue wait-tfull, epilogue body, tiles, kernel cycles, 0}
  unsigned long long* prof;
};
void set_gemm_profile_buffer(unsigned long long* buf);

// N tile that minimises (waves x per-tile time) for `m_blocks` x ceil(n_s / bn) tiles on `num_sms` persistent CTAs.
int pick_gemm_bn(int m_blocks, const int64_t* n_extent, int nseg, int num_workers);
// 2 (default): CTA pairs with tcgen05.mma.cta_group::2 (256-row tiles); 1: single-CTA tiles (AVC_GEMM_CTA_GROUP)
int gemm_cta_group();
// M sub-tiles per CTA of the pair kernel: 2 = every CTA stages 256 rows of A -> 512 x bn pair tiles (default for
// the dW GEMM), 1 = 256 x bn pair tiles (default for the forward).  AVC_GEMM_MT[_TN|_NT] override.
int gemm_m_subtiles(int cta_group, GemmMode mode);
// num_m_blocks counts blocks of cta_group * m_subtiles * GEMM_BM rows; the TN tensor-map boxes must hold
// m_subtiles * GEMM_BM rows of A and bn / cta_group rows of B
cudaError_t launch_gemm(const GemmArgs& args, GemmMode mode, GemmOut out, int cta_group, int m_subtiles,
                        int num_sms, cudaStream_t stream);
// Reduction slices per work item that minimise the dW launch's makespan for `base_items` (tiles + bias items) on
// `workers` persistent workers with `total_kb` K iterations each: ceil(items * S / workers) / S + a per-slice overhead,
// S <= 8, at least 8 K iterations per slice.  1 when the items already fill the workers.
int gemm_dw_splits(int base_items, int workers, int total_kb);
// work items (tiles + tail sub-tiles) the launch above will schedule; the fused all-reduce needs <= COMM_MAX_ITEMS
int gemm_work_items(const GemmArgs& args, int cta_group, int num_sms);
// flags every owner that this rank's extra ranges (bias gradients) hold their partial sums for `epoch`
cudaError_t launch_comm_signal_extra(const CommArgs& comm, cudaStream_t stream);
// Force-load the kernels that get launched while a fused GEMM is already waiting for them (lazy module loading can
// block behind running kernels): colsum and the signal kernel.
cudaError_t preload_comm_kernels();
cudaError_t preload_colsum();

// ------------------------------------------------------------------ gather (align + stack + concat)
struct GatherArgs {
  const uint8_t* src[2];      // audio, video feature bases (bytes); nullptr: modality absent
  int64_t batch_stride[2];    // bytes
  int64_t frame_stride[2];    // bytes
  int frame_bytes[2];         // D * sizeof(bf16)
  int frames[2];              // T per modality (upper bound on valid frames)
  int k[2];                   // frames stacked per token
  int rep[2];                 // each stacked token is used rep times (token j -> stack j / rep)
  const int32_t* len[2];      // optional per-sample valid frame counts (device), nullptr: frames[i]
  const int32_t* tok_offset;  // [B+1] packed row offsets (device), nullptr: uniform tokens_per_sample
  int tokens_per_sample;      // uniform mode
  int batch;
  int64_t total_rows;         // M
  uint8_t* dst;               // A [M, K] bytes
  int64_t dst_row_bytes;      // K * 2
  uint8_t* row_flags;         // [M] (bit0 audio token present, bit1 video token present), may be null
};
cudaError_t launch_gather(const GatherArgs& args, int num_sms, cudaStream_t stream);

// ------------------------------------------------------------------ splice (scatter + masks)
struct SpliceArgs {
  const int64_t* input_ids;   // [B, S]
  int64_t placeholder_id;
  int64_t pad_id;
  int batch, seq;             // B, S
  int row_bytes;              // H * 2
  const int32_t* tok_offset;  // [B+1] or nullptr (uniform)
  int tokens_per_sample;
  // forward: Y rows + embedding table -> inputs_embeds
  const uint8_t* y;           // [M, H] projected rows (fwd) ; nullptr in bwd
  const uint8_t* embed_table; // [V, H] bf16, may be null (text rows zero-filled)
  int64_t vocab;
  uint8_t* inputs_embeds;     // [B, S, H] (fwd: out; bwd: grad in)
  uint8_t* dy;                // bwd: [M, H] out
  // masks
  int64_t* attention_mask;    // [B, S] out (fwd) may be null
  int mask_mode;              // 0 = all ones (reference), 1 = valid tokens only
  const int64_t* labels_in;   // [B, L] may be null
  int label_len;              // L
  int64_t* labels_out;        // [B, S] may be null
  int label_mode;             // 0 = reference (pad->-100, truncate / right-pad -100), 1 = causal-LM
  int32_t* status;            // device int: set to nonzero on malformed input (placeholder count)
  int av_in_place;            // fwd: placeholder rows were already written by the GEMM epilogue: leave them
};
cudaError_t launch_splice_fwd(const SpliceArgs& args, int num_sms, cudaStream_t stream);
cudaError_t launch_splice_bwd(const SpliceArgs& args, int num_sms, cudaStream_t stream);

// ------------------------------------------------------------------ small elementwise / reductions
// dst_bf16[r, c] = bf16(alpha * src_f32[r, c]),  src ld = src_ld, dst ld = dst_ld (elements)
cudaError_t launch_pack_weight(const float* src, int64_t src_ld, void* dst, int64_t dst_ld,
                               int64_t rows, int64_t cols, float alpha, cudaStream_t stream);

// dst_bf16[r, c] = bf16(alpha * src[r, c]); src_dtype is a GemmOut code (bf16 / fp32 / fp16); leading dims in elements
cudaError_t launch_cast_bf16(const void* src, int src_dtype, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows,
                             int64_t cols, float alpha, cudaStream_t stream);

// Input gradient of the gather for ONE stream (see gather.cu for the forward map): d_feat[b, t, :] = sum of
// dA[row(b, j), col_off + (t % k) * dim ...] over the tokens j of sample b with j / rep == t / k; zero otherwise.
struct GatherBwdArgs {
  const uint8_t* da;          // bf16 [total_rows, a_row_stride]
  int64_t a_row_stride;       // elements
  int64_t col_off;            // first column of this stream's segment (elements)
  uint8_t* dst;               // d_feat [batch][frames][dim], bf16 or fp32
  int64_t batch_stride;       // bytes
  int64_t frame_stride;       // bytes
  int batch, frames, dim;
  int k, rep;
  const int32_t* valid;       // [batch] or nullptr
  const int32_t* tok_offset;  // [batch + 1] or nullptr (uniform tokens_per_sample)
  int tokens_per_sample;
};
cudaError_t launch_gather_bwd(const GatherBwdArgs& args, bool out_f32, cudaStream_t stream);

// Column sums of dY over flagged rows, deterministic two-pass:
//   out0[c] = alpha0 * sum_{r : flag bit0} dY[r, c],  out1[c] = alpha1 * sum_{r : flag bit1} dY[r, c]
struct ColsumArgs {
  const uint8_t* dy;         // [batch][rows][H] bf16
  int64_t row_stride;        // bytes
  int64_t batch_stride;      // bytes
  int batch, rows, cols;     // rows summed per batch entry
  int row_base;              // first of those rows inside each batch entry (e.g. the prompt length)
  const uint8_t* row_flags;  // [batch*rows] or nullptr (analytic)
  int flag_rows0, flag_rows1;
  float alpha0, alpha1;
  float* out0;               // [cols] may be null
  float* out1;               // [cols] may be null
  float* workspace;          // colsum_workspace_bytes(); the first colsum_workspace_header_bytes() are zero
                             // before the first launch (the kernel leaves them zero)
  // data parallel: when the sums are final, flag them ready for this epoch's fused all-reduce at the
  // first sig_owners ranks (the ranks that own a chunk of the extra ranges); 0: no signal
  int sig_owners;
  int sig_rank;
  uint32_t sig_epoch;
  uint32_t* sig_flags[COMM_MAX_WORLD];
};
size_t colsum_workspace_bytes(int cols);
size_t colsum_workspace_header_bytes();
cudaError_t launch_colsum(const ColsumArgs& args, cudaStream_t stream);

// Row resampling (sparse row mixing): out[b, i, :] = sum_t weight[t] * x[b, col[t], :], t in CSR row i.
// Train-time length adaptation of the reference (adaptive avg-pool / linear interpolation,
// clip_whisper_model.py:621-676) and its transpose for the backward.
struct ResampleArgs {
  const uint8_t* x;   // [batch, src_rows, hidden]
  uint8_t* out;       // [batch, dst_rows, hidden]
  int dtype;          // GemmOut code: bf16, fp32 or fp16 (accumulation is fp32)
  int batch, src_rows, dst_rows, hidden;
  const int32_t* row_ptr;  // [dst_rows + 1]
  const int32_t* col;      // [nnz] source row index
  const float* weight;     // [nnz]
};
cudaError_t launch_row_resample(const ResampleArgs& args, cudaStream_t stream);

// GELU (erf form, modality_connector.py:60) between the two layers of the MLP projector, with the per-row
// "token present" mask of the fused connector:  fwd out = m * gelu(z);  bwd out = m * dh * gelu'(z).  bf16.
struct GeluArgs {
  const uint8_t* z;
  const uint8_t* dh;   // bwd only
  uint8_t* out;
  int64_t rows, cols, z_ld, dh_ld, out_ld;  // leading dimensions in elements
  const uint8_t* row_flags;  // [rows] or null (all rows on)
  int flag_bit;              // mask applied to row_flags[r]
};
cudaError_t launch_gelu(const GeluArgs& args, bool backward, cudaStream_t stream);
// dst_bf16[c, r] = bf16(alpha * src_f32[r, c])
cudaError_t launch_pack_weight_t(const float* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t rows,
                                 int64_t cols, float alpha, cudaStream_t stream);

// Trainer step for the projector parameters (clip_whisper_trainer.py:171-207, 453-464): deterministic
// sum of squares of the gradient bucket (for the global-norm clip) and AdamW.
size_t sumsq_workspace_bytes();
cudaError_t launch_sumsq(const float* x, int64_t n, float* out, float* workspace, int accumulate,
                         cudaStream_t stream);
struct AdamWArgs {
  float* param;            // [rows, cols] fp32, contiguous
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  int64_t rows, cols;
  float decay;             // 1 - lr * weight_decay
  float one_minus_beta1, beta2, one_minus_beta2;
  float step_size;         // lr / (1 - beta1^t)
  float bias_correction2_sqrt;  // sqrt(1 - beta2^t)
  float eps;
  const float* grad_scale; // device scalar multiplied into the gradient, may be null
  const float* clip_sumsq; // device scalar: squared global gradient norm; with max_norm > 0 the gradient is also
  float max_norm;          //   multiplied by min(1, max_norm / (sqrt(*clip_sumsq) + 1e-6)) (clip_grad_norm_)
  uint8_t* packed;         // optional bf16 [rows, packed_ld] destination for packed_alpha * param
  int64_t packed_ld;
  float packed_alpha;
};
cudaError_t launch_adamw(const AdamWArgs& args, cudaStream_t stream);

}  // namespace avc
