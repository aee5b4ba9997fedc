// Projector GEMMs for the clip_whisper connector on sm_100a.
//
//   forward  (GEMM_TN):  Y[b, r, n] = act( sum_seg  A_seg[b, r, :] . W_seg[n, :]  + flag0*bias0[n] + flag1*bias1[n] )
//   backward (GEMM_NT):  dW_seg[h, k] = alpha_seg * sum_{b, r} dY[b, r, h] * X_seg[b, r, k]
//                        db_i[h]      = alpha_i   * sum_{b, r} dY[b, r, h] * F[b, r, i]      ("bias items", same launch)
//
// This replaces the two nn.Linear calls + pad + weighted sum of the reference
// (modality_connector.py:43-44, clip_whisper_model.py:424-434) and their autograd dW.
//
// Structure: persistent CTAs, 6 warps (10 with the fused gradient all-reduce):
//   warp 0      TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      tcgen05.mma issuer (one thread), owns the TMEM allocation (512 columns = 2 accumulators)
//   warps 2..5  epilogue: tcgen05.ld (TMEM -> registers) -> bias/GELU/scale -> swizzled smem -> TMA store
//   warps 6..9  (COMM != 0, data-parallel dW only) all-reduce of finished gradient tiles over peer-mapped memory:
//               wait for every rank's "tile ready" flag, peer loads, sum in rank order, peer stores to every rank
// The two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
//
// CG = 1: one CTA per SM computes a 128 x bn tile (4 stages of 16 KB A + 32 KB B).
// CG = 2: a CTA pair (cluster of 2, one TPC) computes a 256 x bn tile with tcgen05.mma.cta_group::2:
//         each CTA stages its own 128 rows of A and HALF of B (6 stages of 16 KB + 16 KB), the leader CTA
//         issues the MMAs for both, each CTA's TMEM receives its 128 rows, each CTA runs its own epilogue.
//         Shared-memory traffic per FLOP drops by a third, which is what bounds the single-CTA kernel.
#include <cstdlib>

#include "avc_kernels.h"
#include "avc_ptx.cuh"

namespace avc {

namespace {

constexpr int BM = GEMM_BM, BN = GEMM_BN, BK = GEMM_BK;
constexpr int UMMA_K = 16;
constexpr int kAccStages = 2;
constexpr int kTmemCols = 512;
constexpr int A_STAGE_BYTES = BM * BK * 2;  // 16 KB
constexpr int EPI_WARPS = 4;
constexpr int EPI_BUF_BYTES = 32 * 128;                   // 32 rows x 128 B per warp per buffer
constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_BUF_BYTES;  // 32 KB
constexpr int BAR_BYTES = 256;
constexpr int COMM_WARPS = 4;      // fused all-reduce: warps that move finished gradient tiles between the GPUs
constexpr int COMM_UNIT_ROWS = 8;  // rows of a tile one comm warp reduces per step of its work list
constexpr int threads_for(bool comm) { return 32 * (2 + EPI_WARPS + (comm ? COMM_WARPS : 0)); }
constexpr int MN_ATOM_BYTES = BK * 128;  // one 64-wide MN-major atom column: BK rows x 128 B
constexpr int kMaxStages = 6;

// MT = M sub-tiles per CTA.  MT = 2 (CTA pairs only): each CTA stages 256 rows of A and the pair computes a
// 512 x bn tile as two M = 256 UMMAs per K step that share the B operand; the two accumulators fill all 512 TMEM
// columns (no accumulator double buffering).  Operand bytes staged per FLOP drop by a quarter (the B half is
// written once and read by both UMMAs), which is what bounds the 256 x 256 kernel.
template <int CG, int MT>
struct Cfg {
  static constexpr int kStages = CG == 2 ? (MT == 2 ? 4 : 6) : 4;
  static constexpr int A_BYTES = A_STAGE_BYTES * MT;        // 16 KB / 32 KB
  static constexpr int B_STAGE_BYTES = (BN / CG) * BK * 2;  // 32 KB / 16 KB
  static constexpr int STAGE_BYTES = A_BYTES + B_STAGE_BYTES;
  static constexpr int kAcc = MT == 2 ? 1 : 2;              // accumulator stages in TMEM
  static constexpr int SMEM_USED = kStages * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
  static constexpr int SMEM_ALLOC = SMEM_USED + 1024;  // slack for manual 1024-B alignment
  static_assert(SMEM_ALLOC <= 232448, "exceeds 227 KB of dynamic shared memory");
  static_assert(kStages <= kMaxStages, "barrier area too small");
};
static_assert(kAccStages * BN <= kTmemCols, "accumulators exceed TMEM");
static_assert((2 * kMaxStages + 2 * kAccStages) * 8 + 8 <= BAR_BYTES, "barrier area too small");
static_assert(2 * BN <= kTmemCols, "two M sub-tile accumulators exceed TMEM");

__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// ---------------------------------------------------------------------------------- fused all-reduce (comm warps)
// Wait until every rank's flag word f[0 .. world) has reached this launch's epoch.  Bounded: a peer that never shows
// up sets the status word (the host raises) instead of hanging the box.
__device__ __forceinline__ bool comm_wait(const uint32_t* f, const CommArgs& cm, int lane) {
  uint64_t t0 = 0;
  uint32_t spins = 0;
  for (;;) {
    // Relaxed polls; once every flag has arrived, an ACQUIRE load of the same words orders the data loads behind the
    // flags.  Not a fence: fence.acq_rel.sys (MEMBAR.SYS) also waits until this warp's earlier peer stores are
    // acknowledged system-wide, which measured ~4 us per wait (ncu: stall_membar dominated the comm warps).
    const uint32_t v = lane < cm.world ? ld_relaxed_sys_u32(f + lane) : cm.epoch;
    if (__all_sync(0xffffffffu, static_cast<int32_t>(v - cm.epoch) >= 0)) {
      if (lane < cm.world) (void)ld_acquire_sys(f + lane);
      __syncwarp();
      return true;
    }
    if (cm.poll_ns != 0) __nanosleep(cm.poll_ns);  // measured: 0 .. 20 us makes no difference to the launch time
    if ((++spins & 0x7fu) == 0) {
      int bail = 0;
      if (lane == 0) {
        const uint64_t now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > cm.timeout_ns || *reinterpret_cast<volatile int32_t*>(cm.status) != 0) {
          atomicCAS(cm.status, 0, COMM_ERR_TIMEOUT);
          bail = 1;
        }
      }
      if (__shfl_sync(0xffffffffu, bail, 0)) return false;
    }
  }
}

// out[p][base + r * ld + 4 c] (every rank p) = sum over ranks q (in rank order) of in[q][same] for r < rows,
// c < nvec: 16 / W peer loads of 16 bytes in flight per lane, then the adds, then W peer stores.
template <int W, bool RUNTIME_WORLD>
__device__ __forceinline__ void comm_reduce_rows(const CommArgs& cm, int64_t base, int ld, int rows, int nvec,
                                                 int lane) {
  constexpr int C = 16 / W;
  const int total = rows * nvec;
  const int world = RUNTIME_WORLD ? cm.world : W;
  for (int v0 = 0; v0 < total; v0 += 32 * C) {
    float4 x[C][W];
    int64_t off[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int v = v0 + c * 32 + lane;
      off[c] = -1;
      if (v < total) {
        const int r = v / nvec;
        off[c] = base + static_cast<int64_t>(r) * ld + 4 * (v - r * nvec);
#pragma unroll
        for (int p = 0; p < W; ++p)
          if (p < world) x[c][p] = ld_peer_v4(cm.data[p] + off[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (off[c] < 0) continue;
      float4 s = x[c][0];
#pragma unroll
      for (int p = 1; p < W; ++p)
        if (p < world) { s.x += x[c][p].x; s.y += x[c][p].y; s.z += x[c][p].z; s.w += x[c][p].w; }
#pragma unroll
      for (int p = 0; p < W; ++p)
        if (p < world) st_peer_v4(cm.data[p] + off[c], s);
    }
  }
}

// The same reduction through a multicast mapping of all ranks' buckets: one multimem.ld_reduce per vector returns the sum
// over the ranks (added inside the NVSwitch), one multimem.st writes it to every rank.  Per GPU and direction this moves
// (1 + 1 / world) x the bucket instead of 2 (world - 1) / world x with peer loads and stores.
// 8 vectors of 16 bytes per lane (4 KB per warp) in flight.  16 -- for every round, or only for the schedule's last,
// exposed round -- measured no faster at 2 GPUs (0.940 vs 0.938 ms / step) and costs 70 registers per thread.
__device__ __forceinline__ void comm_reduce_rows_mc(const CommArgs& cm, int64_t base, int ld, int rows, int nvec,
                                                    int lane) {
  constexpr int C = 8;
  const int total = rows * nvec;
  for (int v0 = 0; v0 < total; v0 += 32 * C) {
    float4 x[C];
    int64_t off[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const int v = v0 + c * 32 + lane;
      off[c] = -1;
      if (v < total) {
        const int r = v / nvec;
        off[c] = base + static_cast<int64_t>(r) * ld + 4 * (v - r * nvec);
        x[c] = multimem_ld_reduce_add_v4(cm.mc + off[c]);
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c)
      if (off[c] >= 0) multimem_st_v4(cm.mc + off[c], x[c]);
  }
}

// COMM template parameter of the kernel: 0 = plain GEMM; 1 / 2 / 4 / 8 = fused all-reduce specialised for that world
// size; -1 = fused all-reduce for any world size <= COMM_MAX_WORLD (peer loops predicated at run time).
// COMM == COMM_MC: any world size, data through the multicast mapping (a separate instantiation: the kernel's register
// count must stay low enough for the bias-sum kernel's CTAs to fit next to the GEMM CTAs, see launch_cg).
constexpr int COMM_MC = -2;
template <int COMM>
__device__ __forceinline__ void comm_reduce(const CommArgs& cm, int64_t base, int ld, int rows, int nvec, int lane) {
  if constexpr (COMM == COMM_MC) comm_reduce_rows_mc(cm, base, ld, rows, nvec, lane);
  else if constexpr (COMM > 0) comm_reduce_rows<COMM, false>(cm, base, ld, rows, nvec, lane);
  else comm_reduce_rows<COMM_MAX_WORLD, true>(cm, base, ld, rows, nvec, lane);
}

// ACT (forward only): 0 = identity, 1 = erf GELU in the epilogue.  A template parameter, not a run-time branch: ptxas
// if-converts the branch, and ~1600 predicated-off GELU instructions per 64-column chunk then sit in every epilogue
// warp's instruction stream (measured: the epilogue of a 512 x 256 tile took 12 us instead of 3).
template <int MODE, int OUT, int CG, int MT, int COMM, int ACT = 0>
__global__ void __launch_bounds__(threads_for(COMM != 0), 1) gemm_kernel(const __grid_constant__ GemmArgs args) {
  using C = Cfg<CG, MT>;
  constexpr bool OUT_F32 = OUT == GEMM_OUT_F32;
  constexpr int kStages = C::kStages;
  constexpr int STAGE_BYTES = C::STAGE_BYTES;
  constexpr int A_BYTES = C::A_BYTES;
  constexpr int kAcc = C::kAcc;
  constexpr int TILE_M = BM * CG * MT;  // rows of one work item
  static_assert(MT == 1 || CG == 2, "MT = 2 needs the CTA-pair kernel");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_epi = smem_base + kStages * STAGE_BYTES;
  const uint32_t s_bar = s_epi + EPI_BYTES;
  auto full_bar = [&](uint32_t s) { return s_bar + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return s_bar + 8u * (kStages + s); };
  auto tfull_bar = [&](uint32_t a) { return s_bar + 8u * (2 * kStages + a); };
  auto tempty_bar = [&](uint32_t a) { return s_bar + 8u * (2 * kStages + kAccStages + a); };
  const uint32_t s_tmem_slot = s_bar + 8u * (2 * kStages + 2 * kAccStages);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;  // position inside the CTA pair
  const bool leader = rank == 0;
  const int worker = CG == 2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = CG == 2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (args.prof != nullptr && threadIdx.x == 0) args.prof[2 * 148 * 8 + blockIdx.x * 8 + 0] = globaltimer_ns();
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&args.ma[0]);
    tma_prefetch_desc(&args.mb[0]);
    tma_prefetch_desc(&args.md[0]);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(full_bar(s), 1);
        mbar_init(empty_bar(s), 1);
      }
      for (int a = 0; a < kAcc; ++a) {
        mbar_init(tfull_bar(a), 1);
        mbar_init(tempty_bar(a), EPI_WARPS * CG);  // epilogue warps of every CTA of the group
      }
      fence_mbar_init();
    }
    __syncwarp();
    if (CG == 2) {
      tmem_alloc_cg2(s_tmem_slot, kTmemCols);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(s_tmem_slot, kTmemCols);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();  // peer barriers are initialised before any remote arrive / multicast
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(s_tmem_slot));

  const int num_tiles = args.num_m_blocks * args.num_n_blocks;  // m blocks of CG*128 rows
  const int bn = args.bn;        // runtime N tile (multiple of 64, <= BN)
  const int bn_cta = bn / CG;    // B rows / columns this CTA stages (TMA box); sub-tiles use a prefix of it
  const int nb_boxes = (bn_cta + 63) / 64;  // NT: 64-wide MN atoms this CTA loads for B
  // Work list: `full_tiles` whole tiles (a multiple of the worker count), then the leftover tiles of the last,
  // partial round cut into `tail_split` narrower sub-tiles so that the tail spreads over all workers.
  const int full_tiles = args.full_tiles;
  const int tail_split = args.tail_split;
  const int total_work = full_tiles + (num_tiles - full_tiles) * tail_split;
  // dW only: after the tiles, one GEMM_BIAS_COLS-wide "bias item" per M block contracts the same dY panel with the
  // token-present operand F instead of X -> columns 0 / 1 of its accumulator are the bias gradients of that M block.
  // They are the shortest items of the list and sit at its end, where the last (partial) round has idle workers.
  // dW with few tiles (fewer than workers): every item is cut into `ksplit` slices of the reduction (K iterations
  // [s * total_kb / ksplit, (s + 1) * total_kb / ksplit)); the slices of an item are neighbours in the work list, their
  // partial results go to a workspace and the LAST epilogue warp to arrive per 32-row slab adds them in slice order
  // (deterministic) into the real output.  ksplit == 1 everywhere else.
  const int ksplit = MODE == GEMM_NT ? args.ksplit : 1;
  const int base_work = total_work + (MODE == GEMM_NT ? args.bias_items : 0);
  const int all_work = base_work * ksplit;
  // returns true for a bias item; w is a BASE item index (work item / ksplit)
  auto decode = [&](int w, int& m_blk, int& n_blk, int& n_off, int& width) -> bool {
    int tile = w;
    n_off = 0;
    width = bn;
    if (MODE == GEMM_NT && w >= total_work) {
      m_blk = w - total_work;
      n_blk = 0;
      width = GEMM_BIAS_COLS;
      return true;
    }
    if (w >= full_tiles) {
      const int u = w - full_tiles;
      tile = full_tiles + u / tail_split;
      width = bn / tail_split;
      n_off = (u % tail_split) * width;
    }
    // grouped rasterisation: `group_m` M blocks share each sweep over N, so the tiles of one round touch
    // ~group_m A panels and ~workers/group_m B panels instead of a few A panels and every B panel
    const int gm = args.group_m;
    const int per_group = gm * args.num_n_blocks;
    const int g = tile / per_group;
    const int first_m = g * gm;
    const int rows_in_group = (args.num_m_blocks - first_m) < gm ? (args.num_m_blocks - first_m) : gm;
    const int r = tile - g * per_group;
    m_blk = first_m + r % rows_in_group;
    n_blk = r / rows_in_group;
    return false;
  };
  // bytes this CTA's TMA loads deliver per stage (OOB parts of a box are zero-filled and still counted)
  const uint32_t cta_tx = A_BYTES + (MODE == GEMM_TN ? static_cast<uint32_t>(bn_cta) * (BK * 2)
                                                            : static_cast<uint32_t>(nb_boxes) * MN_ATOM_BYTES);
  const int total_kb = (MODE == GEMM_TN)
                           ? (args.seg_kblocks[0] + (args.nseg > 1 ? args.seg_kblocks[1] : 0))
                           : (args.red_batches * args.red_kblocks_per_batch);

  if (warp == 0) {
    // ======================================================================= TMA producer
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      // forward: activations stream through once per N sweep, the weights are re-read by every M block
      const uint64_t pol_a = args.l2_hints ? (MODE == GEMM_TN ? kEvictFirst : kEvictNormal) : kEvictNormal;
      const uint64_t pol_b = args.l2_hints ? (MODE == GEMM_TN ? kEvictLast : kEvictNormal) : kEvictNormal;
      // Operand panels of one work item (what does not change along its K loop)
      struct Item {
        int w;
        int c1a, c2a;        // TN: A row / batch coordinate.   NT: first M column of this CTA (c1a)
        int c1b;             // TN: B (weight) row.              NT: first N column of this CTA
        const CUtensorMap* mapb;
        int nbx;             // NT: 64-wide B atoms to load
        uint32_t tx;         // bytes this CTA's loads deliver per stage
      };
      auto locate = [&](Item& t, int w) {
        int m_blk, n_blk, n_off, width;
        const bool bias = decode(w, m_blk, n_blk, n_off, width);
        const int w_cta = width / CG;  // B rows / columns of this CTA that the MMA reads
        t.w = w;
        t.tx = cta_tx;
        if (MODE == GEMM_TN) {
          t.c2a = m_blk / args.m_tiles_per_batch;
          t.c1a = (m_blk % args.m_tiles_per_batch) * TILE_M + static_cast<int>(rank) * (BM * MT);
          t.c1b = n_blk * bn + n_off + static_cast<int>(rank) * w_cta;
          t.mapb = nullptr;
          t.nbx = 1;
        } else {
          const int seg = (!bias && n_blk >= args.n_blocks_seg0) ? 1 : 0;
          t.c1a = m_blk * TILE_M + static_cast<int>(rank) * (BM * MT);
          t.c2a = 0;
          t.c1b = bias ? static_cast<int>(rank) * w_cta
                       : (n_blk - (seg ? args.n_blocks_seg0 : 0)) * bn + n_off + static_cast<int>(rank) * w_cta;
          // bias item: the B operand is one 64-wide atom of F (columns past GEMM_BIAS_COLS are zero-filled by TMA)
          t.mapb = bias ? &args.mf : &args.mb[seg];
          t.nbx = bias ? 1 : nb_boxes;
          if (bias) t.tx = A_BYTES + MN_ATOM_BYTES;
        }
      };
      // visits the TMA boxes of K iteration `it` of item t: f(map, smem offset inside the stage, c0, c1, c2, policy)
      auto boxes = [&](const Item& t, int it, auto&& f) {
        if (MODE == GEMM_TN) {
          const int seg = it < args.seg_kblocks[0] ? 0 : 1;
          const int kb = it - (seg ? args.seg_kblocks[0] : 0);
          f(&args.ma[seg], 0u, kb * BK, t.c1a, t.c2a, pol_a);
          f(&args.mb[seg], static_cast<uint32_t>(A_BYTES), kb * BK, t.c1b, 0, pol_b);
        } else {
          const int bb = it / args.red_kblocks_per_batch;
          const int row = (it - bb * args.red_kblocks_per_batch) * BK;
#pragma unroll
          for (int i = 0; i < BM * MT / 64; ++i)
            f(&args.ma[0], static_cast<uint32_t>(i * MN_ATOM_BYTES), t.c1a + i * 64, args.a_row_base + row, bb, pol_a);
          for (int i = 0; i < t.nbx; ++i)
            f(t.mapb, static_cast<uint32_t>(A_BYTES + i * MN_ATOM_BYTES), t.c1b + i * 64, row, bb, pol_b);
        }
      };
      // (An L2 prefetch cursor running a few K iterations ahead of the loads -- cp.async.bulk.prefetch.tensor, by every
      // CTA or only by the first CTA of a round that touches a panel -- measured 15 - 40 % SLOWER: the prefetches occupy
      // the TMA unit like loads do.  profiles/README.md, round 2.)
      Item cur;
      long long t_wait = 0;
      const long long t_begin = clock64();
      for (int w = worker; w < all_work; w += num_workers) {
        locate(cur, w / ksplit);
        const int sl = w % ksplit;
        const int it_end = (sl + 1) * total_kb / ksplit;
        for (int it = sl * total_kb / ksplit; it < it_end; ++it) {
          const long long t0 = clock64();
          mbar_wait(empty_bar(stage), phase ^ 1u);
          t_wait += clock64() - t0;
          if (leader) mbar_arrive_expect_tx(full_bar(stage), cur.tx * CG);
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t bar = full_bar(stage);
          boxes(cur, it, [&](const CUtensorMap* m, uint32_t off, int c0, int c1, int c2, uint64_t pol) {
            if (CG == 2) tma_load_3d_cg2(sa + off, m, bar & kPeerBitMask, c0, c1, c2, pol);
            else tma_load_3d_hint(sa + off, m, bar, c0, c1, c2, pol);
          });
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
      if (args.prof != nullptr) {
        args.prof[blockIdx.x * 8 + 0] = static_cast<unsigned long long>(t_wait);
        args.prof[blockIdx.x * 8 + 6] = static_cast<unsigned long long>(clock64() - t_begin);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ======================================================================= MMA issuer (leader CTA only)
    if (lane == 0 && leader) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      long long t_full = 0, t_tempty = 0;
      for (int w = worker; w < all_work; w += num_workers) {
        int m_blk, n_blk, n_off, width;
        decode(w / ksplit, m_blk, n_blk, n_off, width);
        const uint32_t idesc = make_idesc_bf16(BM * CG, width, MODE == GEMM_NT ? 1 : 0, MODE == GEMM_NT ? 1 : 0);
        const long long t0 = clock64();
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        t_tempty += clock64() - t0;
        tc_fence_after();
        const int it_begin = (w % ksplit) * total_kb / ksplit;
        const int it_end = (w % ksplit + 1) * total_kb / ksplit;
        for (int it = it_begin; it < it_end; ++it) {
          const long long t1 = clock64();
          mbar_wait(full_bar(stage), phase);
          t_full += clock64() - t1;
          tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          const uint32_t sb = sa + A_BYTES;
#pragma unroll
          for (int h = 0; h < MT; ++h) {
            // MT = 2: rows [128 h, 128 h + 128) of each CTA's A stage -> accumulator h (TMEM columns 256 h ..)
            const uint32_t d_tmem = tmem_base + (MT == 2 ? h * BN : acc * BN);
            const uint32_t sah = sa + h * A_STAGE_BYTES;
#pragma unroll
            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
              uint64_t da, db;
              if (MODE == GEMM_TN) {
                // K-major: 8-row x 128-B swizzle atoms stacked along M/N (SBO = 1024 B); stepping
                // UMMA_K = 16 elements inside the 128-B row is +32 B on the start address.
                da = make_smem_desc_sw128(sah + kk * (UMMA_K * 2), 0, 1024);
                db = make_smem_desc_sw128(sb + kk * (UMMA_K * 2), 0, 1024);
              } else {
                // MN-major: each contraction row is one 128-B swizzle row of 64 M/N elements;
                // 8 rows per atom (SBO = 1024 B), next 64 M/N elements at LBO = BK*128 B;
                // UMMA_K = 16 contraction rows = +2048 B.
                da = make_smem_desc_sw128(sah + kk * (UMMA_K * 128), MN_ATOM_BYTES, 1024);
                db = make_smem_desc_sw128(sb + kk * (UMMA_K * 128), MN_ATOM_BYTES, 1024);
              }
              const uint32_t accum = (it != it_begin || kk != 0) ? 1u : 0u;
              if (CG == 2) umma_f16_cg2(d_tmem, da, db, idesc, accum);
              else umma_f16(d_tmem, da, db, idesc, accum);
            }
          }
          // frees this smem stage (in both CTAs of a pair) once the MMAs above retire
          if (CG == 2) umma_commit_cg2(empty_bar(stage), 3);
          else umma_commit(empty_bar(stage));
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        // accumulator complete -> epilogue (of both CTAs)
        if (CG == 2) umma_commit_cg2(tfull_bar(acc), 3);
        else umma_commit(tfull_bar(acc));
        if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
      }
      if (args.prof != nullptr) {
        args.prof[blockIdx.x * 8 + 1] = static_cast<unsigned long long>(t_full);
        args.prof[blockIdx.x * 8 + 2] = static_cast<unsigned long long>(t_tempty);
        args.prof[2 * 148 * 8 + blockIdx.x * 8 + 5] = globaltimer_ns();
      }
    }
    __syncwarp();
  } else if (COMM == 0 || warp < 2 + EPI_WARPS) {
    // ======================================================================= epilogue warps
    const int q = warp & 3;    // TMEM lane quarter this warp may access
    const int ew = warp - 2;   // staging buffer owner index
    constexpr int COLS = OUT_F32 ? 32 : 64;  // columns per 128-byte staging row
    uint32_t acc = 0, acc_phase = 0, buf = 0;
    long long t_tfull = 0, t_body = 0, n_tiles = 0;
    long long t_ph[5] = {0, 0, 0, 0, 0};  // per chunk: wait staging buffer, TMEM load, math + st.shared, proxy fence, TMA store
    for (int w = worker; w < all_work; w += num_workers) {
      int m_blk, n_blk, n_off, width;
      const int wb = w / ksplit;         // base item
      const int sl = w - wb * ksplit;    // reduction slice of this work item
      const bool bias = decode(wb, m_blk, n_blk, n_off, width);
      const int nchunk = bias ? 0 : width / COLS;
      int tile_row0, out_batch, out_col0, seg = 0;
      if (MODE == GEMM_TN) {
        out_batch = m_blk / args.m_tiles_per_batch;
        tile_row0 = (m_blk % args.m_tiles_per_batch) * TILE_M + static_cast<int>(rank) * (BM * MT) + q * 32;
        out_col0 = n_blk * bn + n_off;
      } else {
        out_batch = 0;
        tile_row0 = m_blk * TILE_M + static_cast<int>(rank) * (BM * MT) + q * 32;
        seg = (!bias && n_blk >= args.n_blocks_seg0) ? 1 : 0;
        out_col0 = (n_blk - (seg ? args.n_blocks_seg0 : 0)) * bn + n_off;
      }
      const float alpha = args.alpha[seg];
      const int ncols = args.d_cols[seg];

      // Bias slice of this item's columns, spread over the lanes (float4 group G = column / 4 lives in lane G % 32, slot
      // G / 32) and loaded BEFORE the wait for the accumulator: the L2 latency of these loads (~800 cycles under load,
      // paid twice per 64-column chunk when they sat inside the chunk loop) disappears behind the main loop.
      float4 bias_reg[2][2];
      if (MODE == GEMM_TN) {
#pragma unroll
        for (int s2 = 0; s2 < 2; ++s2) {
          const float* bp = s2 == 0 ? args.bias0 : args.bias1;
#pragma unroll
          for (int slot = 0; slot < 2; ++slot) {
            const int n = out_col0 + (slot * 32 + lane) * 4;
            bias_reg[s2][slot] = (bp != nullptr && (slot * 32 + lane) * 4 < width && n < ncols)
                                     ? __ldg(reinterpret_cast<const float4*>(bp + n))
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
      const long long te0 = clock64();
      mbar_wait(tfull_bar(acc), acc_phase);
      tc_fence_after();
      const long long te1 = clock64();
      t_tfull += te1 - te0;
      ++n_tiles;

#pragma unroll 1
      for (int h = 0; h < MT; ++h) {  // MT = 2: accumulator h holds rows [128 h, 128 h + 128) of this CTA
      const int out_row0 = tile_row0 + h * BM;
      const int my_row = out_row0 + lane;
      float f0 = 0.f, f1 = 0.f;
      if (MODE == GEMM_TN) {
        if (args.row_flags != nullptr) {
          if (my_row < args.d_rows) {
            const uint8_t fl = args.row_flags[static_cast<int64_t>(out_batch) * args.d_rows + my_row];
            f0 = (fl & 1) ? args.bias_scale[0] : 0.f;
            f1 = (fl & 2) ? args.bias_scale[1] : 0.f;
          }
        } else {
          f0 = my_row < args.flag_rows0 ? args.bias_scale[0] : 0.f;
          f1 = my_row < args.flag_rows1 ? args.bias_scale[1] : 0.f;
        }
      }
      const bool has_b0 = MODE == GEMM_TN && args.bias0 != nullptr;
      const bool has_b1 = MODE == GEMM_TN && args.bias1 != nullptr;
      const uint32_t t_row = tmem_base + (MT == 2 ? h * BN : acc * BN) + (static_cast<uint32_t>(q * 32) << 16);
      if (MODE == GEMM_NT && bias) {
        // bias item: accumulator column i of this thread's row h is sum_r dY[r, h] * F[r, i]
        uint32_t v0[32];
        tmem_ld_32x32(t_row, v0);
        tmem_ld_wait();
        if (my_row < args.d_rows) {
          if (ksplit > 1) {  // partial sums of this slice: [slice][stream][row]
            float* pb = args.ws_bias + static_cast<int64_t>(sl) * 2 * args.d_rows + my_row;
            pb[0] = args.bias_alpha[0] * __uint_as_float(v0[0]);
            pb[args.d_rows] = args.bias_alpha[1] * __uint_as_float(v0[1]);
          } else {
            if (args.bias_out[0] != nullptr) args.bias_out[0][my_row] = args.bias_alpha[0] * __uint_as_float(v0[0]);
            if (args.bias_out[1] != nullptr) args.bias_out[1][my_row] = args.bias_alpha[1] * __uint_as_float(v0[1]);
          }
        }
      }

#pragma unroll 1
      for (int c = 0; c < nchunk; ++c) {
        const int col = out_col0 + c * COLS;
        // the store issued from this staging buffer two chunks ago must have finished reading it
        const long long tc0 = clock64();
        if (lane == 0 || (MODE == GEMM_TN && args.scatter_rows > 0)) bulk_wait_read<1>();
        __syncwarp();
        const long long tc1 = clock64();
        t_ph[0] += tc1 - tc0;
        const uint32_t sbuf = s_epi + (ew * 2 + buf) * EPI_BUF_BYTES;
        const uint32_t srow = sbuf + lane * 128;
        // COLS accumulator columns of this thread's row -> registers
        uint32_t v[COLS];
        {
          uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
          tmem_ld_32x32(t_row + c * COLS, v0);
          if (!OUT_F32) {
            uint32_t(&v1)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[COLS - 32]);
            tmem_ld_32x32(t_row + c * COLS + 32, v1);
          }
          tmem_ld_wait();
        }
        const long long tc2 = clock64();
        t_ph[1] += tc2 - tc1;
        float x[COLS];
#pragma unroll
        for (int i = 0; i < COLS; ++i) x[i] = __uint_as_float(v[i]);
        if (MODE == GEMM_TN) {
          // Bias: x += f0 * bias0[n] + f1 * bias1[n] with the per-row multipliers f0 / f1 (0 on rows whose token lacks
          // that stream, so no per-lane branches: every predicate below is warp-uniform).  The bias values come from the
          // lane-distributed copy loaded at the top of the work item (columns past N hold zeros).
          constexpr int GPC = COLS / 4;  // float4 groups per chunk
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            if (s2 == 0 ? !has_b0 : !has_b1) continue;  // warp-uniform
            const float f = s2 == 0 ? f0 : f1;
            const int slot = (c * GPC) >> 5;             // warp-uniform: the chunk's groups share one slot
            const float4 mine = slot == 0 ? bias_reg[s2][0] : bias_reg[s2][1];
            const int lane0 = (c * GPC) & 31;
#pragma unroll
            for (int g = 0; g < GPC; ++g) {
              const float bx = __shfl_sync(0xffffffffu, mine.x, lane0 + g);
              const float by = __shfl_sync(0xffffffffu, mine.y, lane0 + g);
              const float bz = __shfl_sync(0xffffffffu, mine.z, lane0 + g);
              const float bw = __shfl_sync(0xffffffffu, mine.w, lane0 + g);
              x[4 * g + 0] = fmaf(f, bx, x[4 * g + 0]); x[4 * g + 1] = fmaf(f, by, x[4 * g + 1]);
              x[4 * g + 2] = fmaf(f, bz, x[4 * g + 2]); x[4 * g + 3] = fmaf(f, bw, x[4 * g + 3]);
            }
          }
          if (ACT == 1) {
#pragma unroll
            for (int i = 0; i < COLS; ++i) x[i] = gelu_erf(x[i]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < COLS; ++i) x[i] *= alpha;
        }
        // one 128-byte staging row per thread, 16-byte chunks XOR-swizzled (matches SWIZZLE_128B)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t dst = srow + ((j ^ (lane & 7)) << 4);
          if (OUT_F32) {
            st_shared_v4(dst, __float_as_uint(x[4 * j + 0]), __float_as_uint(x[4 * j + 1]),
                         __float_as_uint(x[4 * j + 2]), __float_as_uint(x[4 * j + 3]));
          } else {
            const int o = (8 * j) % COLS;  // (COLS == 64 here)
            if (OUT == GEMM_OUT_F16)
              st_shared_v4(dst, pack_f16x2(x[o + 0], x[o + 1]), pack_f16x2(x[o + 2], x[o + 3]),
                           pack_f16x2(x[o + 4], x[o + 5]), pack_f16x2(x[o + 6], x[o + 7]));
            else
              st_shared_v4(dst, pack_bf16x2(x[o + 0], x[o + 1]), pack_bf16x2(x[o + 2], x[o + 3]),
                           pack_bf16x2(x[o + 4], x[o + 5]), pack_bf16x2(x[o + 6], x[o + 7]));
          }
        }
        const long long tc3 = clock64();
        t_ph[2] += tc3 - tc2;
        fence_proxy_async_smem();
        __syncwarp();
        const long long tc4 = clock64();
        t_ph[3] += tc4 - tc3;
        if (MODE == GEMM_TN && args.scatter_rows > 0) {
          // packed rows m = b * scatter_rows + j go to row j of batch entry b of the output map (the AV region of
          // inputs_embeds).  A 32-row box is one TMA store (rows past the sample's end are clipped); the rows of a
          // straddling box that belong to the NEXT sample follow one by one, a 128-byte TMA store per lane (TMA stores
          // may run past the end of a dimension but must not start at a negative coordinate).  A tile never spans more
          // than two samples per 32-row box as long as a sample has >= 32 fused tokens; shorter samples take the
          // row-by-row path for every row past the first sample.
          const int sr = args.scatter_rows;
          const int b0 = out_row0 / sr;
          const int j0 = out_row0 - b0 * sr;
          if (out_row0 < args.d_rows && col < ncols) {
            // the rows of sample b0: one box store (rows past the sample's end are clipped by TMA) ...
            if (lane == 0) tma_store_3d(&args.md[seg], sbuf, col, j0, b0);
            // ... and, when the box straddles the boundary, the rows that belong to the next sample one by one
            if (j0 + lane >= sr && b0 != args.scatter_batches - 1 && my_row < args.d_rows) {
              const int bb = my_row / sr;
              tma_store_3d(&args.md_row, srow, col, my_row - bb * sr, bb);
            }
          }
          bulk_commit();  // every lane keeps its own (possibly empty) bulk-group sequence in scatter mode
        } else if (lane == 0) {
          if (out_row0 < args.d_rows && col < ncols) {
            if (MODE == GEMM_NT && ksplit > 1) tma_store_3d(&args.mws[seg], sbuf, col, out_row0, sl);  // partial tile
            else tma_store_3d(&args.md[seg], sbuf, col, out_row0, out_batch);
          }
          bulk_commit();
        }
        buf ^= 1u;
        t_ph[4] += clock64() - tc4;
      }
      }  // h
      // all TMEM reads of this accumulator are complete (tmem_ld_wait above): hand it back to the issuer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
        else mbar_arrive(tempty_bar(acc));
      }
      if (++acc == kAcc) { acc = 0; acc_phase ^= 1u; }
      t_body += clock64() - te1;
      if (MODE == GEMM_NT && COMM == 0 && ksplit > 1) {
        // Split reduction: this warp's slab of the item (rows tile_row0 + h * 128 .. + 32 for every h) is complete once
        // the same warp slot of all `ksplit` slices has stored its partial.  The last to arrive adds the partials in
        // slice order -- the same order whoever comes last -- and writes the result.
        int last = 0;
        __threadfence();  // bias partials are plain stores of every lane: ordered before lane 0's arrival below
        __syncwarp();
        if (lane == 0) {
          bulk_wait_all<0>();           // this warp's TMA stores of the partial tile have landed
          fence_proxy_async_all();
          __threadfence();
          uint32_t* cnt = args.split_count + static_cast<int64_t>(wb) * (EPI_WARPS * CG) + ew + EPI_WARPS * rank;
          if (atomicAdd(cnt, 1u) == static_cast<uint32_t>(ksplit - 1)) {
            atomicExch(cnt, 0u);
            __threadfence();
            last = 1;
          }
        }
        last = __shfl_sync(0xffffffffu, last, 0);
        if (last) {
          if (bias) {
#pragma unroll 1
            for (int h = 0; h < MT; ++h) {
              const int row = tile_row0 + h * BM + lane;
              if (row < args.d_rows) {
                float s0 = 0.f, s1 = 0.f;
                for (int s2 = 0; s2 < ksplit; ++s2) {
                  const float* pb = args.ws_bias + static_cast<int64_t>(s2) * 2 * args.d_rows + row;
                  s0 += __ldcg(pb);
                  s1 += __ldcg(pb + args.d_rows);
                }
                if (args.bias_out[0] != nullptr) args.bias_out[0][row] = s0;
                if (args.bias_out[1] != nullptr) args.bias_out[1][row] = s1;
              }
            }
          } else {
            // 8 rows x 2 vectors per lane and slice in flight (the partial tiles come from L2 / HBM at ~1 us a round trip)
            constexpr int RB = 8;
            const float* wsb = args.ws + args.ws_seg_col[seg] + out_col0;
            float* dst = args.d_ptr[seg] + out_col0;
#pragma unroll 1
            for (int h = 0; h < MT; ++h) {
#pragma unroll 1
              for (int r0 = 0; r0 < 32; r0 += RB) {
                float4 acc4[RB][2];
#pragma unroll
                for (int i = 0; i < RB; ++i) acc4[i][0] = acc4[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                const int row0 = tile_row0 + h * BM + r0;
                // unconditional loads (rows / columns past the edge re-read the last valid one and are dropped at the
                // store): nothing keeps the compiler from issuing the 16 loads of a slice back to back
                int roff[RB], coff[2];
#pragma unroll
                for (int i = 0; i < RB; ++i) roff[i] = (row0 + i < args.d_rows ? row0 + i : args.d_rows - 1) * args.ws_ld;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  int c4 = 4 * (lane + 32 * j);
                  if (c4 >= width) c4 = width - 4;
                  if (out_col0 + c4 >= ncols) c4 = ncols - out_col0 - 4;
                  coff[j] = c4;
                }
#pragma unroll 1
                for (int s2 = 0; s2 < ksplit; ++s2) {
                  const float* ps = wsb + static_cast<int64_t>(s2) * args.ws_split_stride;
                  float4 t[RB][2];
#pragma unroll
                  for (int i = 0; i < RB; ++i) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) t[i][j] = __ldcg(reinterpret_cast<const float4*>(ps + roff[i] + coff[j]));
                  }
#pragma unroll
                  for (int i = 0; i < RB; ++i) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                      acc4[i][j].x += t[i][j].x; acc4[i][j].y += t[i][j].y;
                      acc4[i][j].z += t[i][j].z; acc4[i][j].w += t[i][j].w;
                    }
                  }
                }
#pragma unroll
                for (int i = 0; i < RB; ++i) {
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    const int c4 = 4 * (lane + 32 * j);
                    if (row0 + i < args.d_rows && c4 < width && out_col0 + c4 < ncols)
                      *reinterpret_cast<float4*>(dst + static_cast<int64_t>(row0 + i) * args.d_ld[seg] + c4) = acc4[i][j];
                  }
                }
              }
            }
          }
        }
      }
      if (COMM != 0 && bias) {
        // The bias gradients of this rank are complete once every epilogue warp of every bias item has stored its
        // rows (plain stores): the last of them flags the extra ranges ready at every rank that owns a chunk of them.
        if (lane == 0) {
          const CommArgs& cm = args.comm;
          __threadfence();
          uint32_t* lf = cm.flags[cm.rank];
          const uint32_t expect = static_cast<uint32_t>(EPI_WARPS * CG * args.bias_items);
          if (atomicAdd(lf + COMM_EXTRA_COUNT, 1u) == expect - 1) {
            atomicExch(lf + COMM_EXTRA_COUNT, 0u);
            __threadfence_system();
            const int nch = ((cm.extra_len[0] + 127) >> 7) + ((cm.extra_len[1] + 127) >> 7);
            for (int p = 0; p < cm.world && p < nch; ++p)
              st_relaxed_sys_u32(cm.flags[p] + COMM_EXTRA_FLAGS + cm.rank, cm.epoch);
          }
        }
        __syncwarp();
      } else if (COMM != 0) {
        // This warp's part of work item w is in the local bucket once its TMA stores have completed.  The last of the
        // item's EPI_WARPS x CG epilogue warps tells the item's owner rank that this rank's partial tile is ready.
        if (lane == 0) {
          const CommArgs& cm = args.comm;
          bulk_wait_all<0>();
          fence_proxy_async_all();
          __threadfence();  // gpu scope is enough up to the arrival counter; the last arriver fences at system scope
          uint32_t* lf = cm.flags[cm.rank];
          if (atomicAdd(lf + COMM_ITEM_COUNT + w, 1u) == EPI_WARPS * CG - 1) {
            atomicExch(lf + COMM_ITEM_COUNT + w, 0u);
            __threadfence_system();  // acquires the other warps' arrivals, releases the flag store below
            st_relaxed_sys_u32(cm.flags[w % cm.world] + COMM_ITEM_FLAGS + w * COMM_MAX_WORLD + cm.rank, cm.epoch);
          }
        }
        __syncwarp();
      }
    }
    if (lane == 0 || (MODE == GEMM_TN && args.scatter_rows > 0)) bulk_wait_all<0>();
    __syncwarp();
    if (args.prof != nullptr && warp == 2 && lane == 0) {
      args.prof[blockIdx.x * 8 + 3] = static_cast<unsigned long long>(t_tfull);
      args.prof[blockIdx.x * 8 + 4] = static_cast<unsigned long long>(t_body);
      args.prof[blockIdx.x * 8 + 5] = static_cast<unsigned long long>(n_tiles);
      for (int i = 0; i < 5; ++i) args.prof[148 * 8 + blockIdx.x * 8 + i] = static_cast<unsigned long long>(t_ph[i]);
      args.prof[2 * 148 * 8 + blockIdx.x * 8 + 1] = globaltimer_ns();
    }
  } else if (COMM != 0) {
    // ======================================================================= comm warps (fused gradient all-reduce)
    const CommArgs& cm = args.comm;
    const int g = static_cast<int>(blockIdx.x) * COMM_WARPS + (warp - 2 - EPI_WARPS);
    const int G = static_cast<int>(gridDim.x) * COMM_WARPS;
    uint32_t* lf = cm.flags[cm.rank];
    constexpr int RG = TILE_M / COMM_UNIT_ROWS;  // row groups (units) per work item
    bool ok = true;
    int cur = -1;
    const bool stamp = args.prof != nullptr && warp == 2 + EPI_WARPS && lane == 0;
    // The schedule's rounds (num_workers items each) complete one after the other, the items of one round together.
    // Per round, the units of the items this rank owns are dealt to the GPU's comm warps in CONTIGUOUS runs, so a warp
    // waits for (and pays the system-scope acquire of) one or two items per round instead of one per unit.
    for (int r0 = 0; ok && r0 < total_work; r0 += num_workers) {
      const int r1 = r0 + num_workers < total_work ? r0 + num_workers : total_work;
      const int first = r0 + (cm.rank - r0 % cm.world + cm.world) % cm.world;  // first owned item of the round
      const int n_own = first < r1 ? (r1 - first + cm.world - 1) / cm.world : 0;
      const int units = n_own * RG;
      const int per = (units + G - 1) / G;
      const int l1 = (g + 1) * per < units ? (g + 1) * per : units;
      for (int l = g * per; l < l1; ++l) {
        const int i = l / RG;
        const int w = first + i * cm.world;
        if (w != cur) {
          ok = comm_wait(lf + COMM_ITEM_FLAGS + w * COMM_MAX_WORLD, cm, lane);
          cur = w;
          if (!ok) break;
          if (stamp && r1 == total_work) args.prof[2 * 148 * 8 + blockIdx.x * 8 + 2] = globaltimer_ns();
        }
        int m_blk, n_blk, n_off, width;
        decode(w, m_blk, n_blk, n_off, width);
        const int seg = (n_blk >= args.n_blocks_seg0) ? 1 : 0;
        const int ld = args.d_cols[seg];
        const int col0 = (n_blk - (seg ? args.n_blocks_seg0 : 0)) * bn + n_off;
        const int row0 = m_blk * TILE_M + (l - i * RG) * COMM_UNIT_ROWS;
        const int rows = min(COMM_UNIT_ROWS, args.d_rows - row0);
        const int cols = min(width, ld - col0);
        if (rows <= 0 || cols <= 0) continue;
        comm_reduce<COMM>(cm, cm.seg_off[seg] + static_cast<int64_t>(row0) * ld + col0, ld, rows, cols >> 2, lane);
      }
    }
    // extra flat ranges (bias gradients, produced by another kernel): 128-float chunks, chunk e owned by rank
    // e % world, the owner's chunks dealt to its comm warps from the back of the warp list
    {
      const int nch0 = (cm.extra_len[0] + 127) >> 7, nch1 = (cm.extra_len[1] + 127) >> 7;
      bool waited = false;
      for (int e = cm.rank; ok && e < nch0 + nch1; e += cm.world) {
        if ((e / cm.world) % G != G - 1 - g) continue;
        if (!waited) {
          ok = comm_wait(lf + COMM_EXTRA_FLAGS, cm, lane);
          waited = true;
          if (!ok) break;
        }
        const int k = e >= nch0 ? 1 : 0;
        const int ch = e - (k ? nch0 : 0);
        comm_reduce<COMM>(cm, cm.extra_off[k] + 128 * ch, 0, 1, min(32, (cm.extra_len[k] - 128 * ch) >> 2), lane);
      }
    }
    // This rank is done once all its comm warps are; the last one tells every rank and then waits until every rank
    // has finished writing into this rank's bucket -- the launch does not complete before the bucket is final.
    __syncwarp();
    if (stamp) args.prof[2 * 148 * 8 + blockIdx.x * 8 + 3] = globaltimer_ns();
    int last = 0;
    if (lane == 0) {
      __threadfence_system();
      if (stamp) args.prof[2 * 148 * 8 + blockIdx.x * 8 + 6] = globaltimer_ns();
      if (atomicAdd(lf + COMM_DONE_COUNT, 1u) == static_cast<uint32_t>(G - 1)) {
        atomicExch(lf + COMM_DONE_COUNT, 0u);
        __threadfence_system();
        for (int p = 0; p < cm.world; ++p) st_relaxed_sys_u32(cm.flags[p] + COMM_DONE_FLAGS + cm.rank, cm.epoch);
        last = 1;
      }
    }
    if (__shfl_sync(0xffffffffu, last, 0)) comm_wait(lf + COMM_DONE_FLAGS, cm, lane);
    if (stamp) args.prof[2 * 148 * 8 + blockIdx.x * 8 + 4] = globaltimer_ns();
  }

  tc_fence_before();
  __syncthreads();
  if (args.prof != nullptr && threadIdx.x == 0) args.prof[2 * 148 * 8 + blockIdx.x * 8 + 7] = globaltimer_ns();
  if (CG == 2) cluster_sync_all();  // the peer may still signal this CTA's barriers / read its smem until here
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, kTmemCols);
    else tmem_dealloc(tmem_base, kTmemCols);
  }
}

unsigned long long* g_prof = nullptr;

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e != nullptr ? atoi(e) : dflt;
}

// Fills in the schedule fields of `args` (tail split, rasterisation, L2 hints); returns the worker count.
template <int CG>
int plan_schedule(GemmArgs& args, GemmMode mode, int num_sms) {
  const int num_tiles = args.num_m_blocks * args.num_n_blocks;
  int max_workers = num_sms / CG;
  const int cap = env_int("AVC_GEMM_MAX_WORKERS", 0);  // test knob: exercise multi-round schedules on small shapes
  if (cap > 0 && cap < max_workers) max_workers = cap;
  // every work item -- tiles and (dW) bias items -- gets its own worker while there are workers left
  const int items = (num_tiles + (mode == GEMM_NT ? args.bias_items : 0)) * (args.ksplit > 1 ? args.ksplit : 1);
  const int workers = items < max_workers ? items : max_workers;
  // Tail of the persistent schedule: cut the leftover tiles of the last (partial) round into narrower
  // sub-tiles when that shortens the round.  Sub-tile width >= 64 (one bf16 epilogue chunk).
  args.full_tiles = num_tiles;
  args.tail_split = 1;
  args.l2_hints = env_int("AVC_GEMM_L2_HINTS", 1);
  // forward: M-major order streams each activation panel once while the (L2-resident, evict-last) weights are
  // re-read; dW: 8 x ~9 tile groups per round cut the panel re-reads (measured DRAM bytes: profiles/)
  args.group_m = env_int("AVC_GEMM_GROUP_M", mode == GEMM_TN ? 1 : 8);
  if (args.group_m < 1) args.group_m = 1;
  if (args.group_m > args.num_m_blocks) args.group_m = args.num_m_blocks;
  const int leftover = num_tiles % workers;
  if (leftover != 0 && env_int("AVC_GEMM_TAIL_SPLIT", 1) != 0) {
    int best = 1;
    double best_cost = 1e30;
    for (int split = 1; split <= 4; split *= 2) {
      const int width = args.bn / split;
      if (width < 64 || width % 64 != 0 || (width / CG) % 16 != 0) break;
      const int rounds = (leftover * split + workers - 1) / workers;
      const double cost = rounds * (width + 128.0);
      if (cost < best_cost * 0.97) { best_cost = cost; best = split; }
    }
    if (best > 1) {
      args.full_tiles = num_tiles - leftover;
      args.tail_split = best;
    }
  }
  return workers;
}

template <int CG, int MT>
cudaError_t launch_cg(const GemmArgs& args_in, GemmMode mode, GemmOut out, int num_sms, cudaStream_t stream) {
  GemmArgs args = args_in;
  args.prof = g_prof;
  const int workers = plan_schedule<CG>(args, mode, num_sms);
  const bool comm = args.comm.world > 0;
  const bool out_fp32 = out == GEMM_OUT_F32;
  if (comm && !(mode == GEMM_NT && out_fp32 && CG == 2)) return cudaErrorInvalidValue;
  if (mode != GEMM_NT && args.bias_items != 0) return cudaErrorInvalidValue;
  if (mode == GEMM_NT && out == GEMM_OUT_F16) return cudaErrorInvalidValue;
  if (args.bias_items != 0 && args.bias_items != args.num_m_blocks) return cudaErrorInvalidValue;
  if (args.ksplit < 1) args.ksplit = 1;
  if (args.ksplit > 1) {
    if (mode != GEMM_NT || !out_fp32 || comm || args.ws == nullptr || args.split_count == nullptr) return cudaErrorInvalidValue;
    args.full_tiles = args.num_m_blocks * args.num_n_blocks;  // no narrow tail sub-tiles together with reduction slices
    args.tail_split = 1;
  }
  args.comm.poll_ns = static_cast<uint32_t>(env_int("AVC_COMM_POLL_NS", 200));
  auto run = [&](auto kern) -> cudaError_t {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<CG, MT>::SMEM_ALLOC);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(workers * CG);
    cfg.blockDim = dim3(threads_for(comm));
    cfg.dynamicSmemBytes = Cfg<CG, MT>::SMEM_ALLOC;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args);
  };
  if constexpr (CG == 2) {
    if (comm) {
      if (args.comm.mc != nullptr) return run(gemm_kernel<GEMM_NT, GEMM_OUT_F32, CG, MT, COMM_MC>);
      switch (args.comm.world) {
        case 1: return run(gemm_kernel<GEMM_NT, GEMM_OUT_F32, CG, MT, 1>);
        case 2: return run(gemm_kernel<GEMM_NT, GEMM_OUT_F32, CG, MT, 2>);
        case 4: return run(gemm_kernel<GEMM_NT, GEMM_OUT_F32, CG, MT, 4>);
        case 8: return run(gemm_kernel<GEMM_NT, GEMM_OUT_F32, CG, MT, 8>);
        default: return run(gemm_kernel<GEMM_NT, GEMM_OUT_F32, CG, MT, -1>);
      }
    }
  }
  if (mode == GEMM_TN) {
    if (args.act == 1) {
      if (out == GEMM_OUT_F32) return run(gemm_kernel<GEMM_TN, GEMM_OUT_F32, CG, MT, 0, 1>);
      if (out == GEMM_OUT_F16) return run(gemm_kernel<GEMM_TN, GEMM_OUT_F16, CG, MT, 0, 1>);
      return run(gemm_kernel<GEMM_TN, GEMM_OUT_BF16, CG, MT, 0, 1>);
    }
    if (out == GEMM_OUT_F32) return run(gemm_kernel<GEMM_TN, GEMM_OUT_F32, CG, MT, 0>);
    if (out == GEMM_OUT_F16) return run(gemm_kernel<GEMM_TN, GEMM_OUT_F16, CG, MT, 0>);
    return run(gemm_kernel<GEMM_TN, GEMM_OUT_BF16, CG, MT, 0>);
  }
  return out_fp32 ? run(gemm_kernel<GEMM_NT, GEMM_OUT_F32, CG, MT, 0>)
                  : run(gemm_kernel<GEMM_NT, GEMM_OUT_BF16, CG, MT, 0>);
}

__global__ void comm_signal_extra_kernel(const __grid_constant__ CommArgs cm) {
  // the kernel that produced the extra ranges ran before this one on the same stream
  __threadfence_system();
  const int nch = ((cm.extra_len[0] + 127) >> 7) + ((cm.extra_len[1] + 127) >> 7);
  const int p = threadIdx.x;  // every rank that owns at least one chunk waits for this flag
  if (p < cm.world && p < nch) st_release_sys(cm.flags[p] + COMM_EXTRA_FLAGS + cm.rank, cm.epoch);
}

}  // namespace

void set_gemm_profile_buffer(unsigned long long* buf) { g_prof = buf; }

int gemm_cta_group() {
  const int v = env_int("AVC_GEMM_CTA_GROUP", 2);
  return v == 1 ? 1 : 2;
}

int gemm_m_subtiles(int cta_group, GemmMode mode) {
  if (cta_group != 2) return 1;
  // Measured on B200 at the cfg2 shapes (profiles/README.md): the forward (47 x 16 pair tiles of 256 rows) loses
  // more to wave quantisation and the exposed epilogue with 512-row tiles than it gains from the lower operand
  // traffic; dW (long reduction, few tiles) gains 10 %.
  const int dflt = mode == GEMM_NT ? 2 : 1;
  const int v = env_int(mode == GEMM_NT ? "AVC_GEMM_MT_NT" : "AVC_GEMM_MT_TN", env_int("AVC_GEMM_MT", dflt));
  return v == 2 ? 2 : 1;
}

int pick_gemm_bn(int m_blocks, const int64_t* n_extent, int nseg, int num_workers) {
  const int forced = env_int("AVC_GEMM_BN", 0);
  if (forced == 64 || forced == 128 || forced == 192 || forced == 256) return forced;
  // Persistent workers run ceil(tiles / workers) rounds; a tile's time grows with bn plus a fixed part (the
  // A-operand shared-memory traffic, pipeline fill/drain, epilogue tail).  Measured on B200: narrower tiles
  // lose far more to operand traffic than they win back from wave quantisation, hence the large constant.
  int best = 256;
  double best_cost = 1e30;
  const int cand[2] = {256, 128};
  for (int c = 0; c < 2; ++c) {
    const int bn = cand[c];
    int64_t tiles = 0;
    for (int s = 0; s < nseg; ++s) tiles += static_cast<int64_t>(m_blocks) * ((n_extent[s] + bn - 1) / bn);
    const int64_t rounds = (tiles + num_workers - 1) / num_workers;
    const double cost = static_cast<double>(rounds) * (bn + 128);
    if (cost < best_cost * 0.98) { best_cost = cost; best = bn; }
  }
  return best;
}

int gemm_dw_splits(int base_items, int workers, int total_kb) {
  const int forced = env_int("AVC_GEMM_KSPLIT", 0);
  if (base_items <= 0 || workers <= 0) return 1;
  int best = 1;
  double best_cost = 1e30;
  // only when the items do not fill one round of the workers: with two or more rounds the schedule is already dense and
  // the partial tiles would cost more than the rounding they remove
  for (int sp = 1; sp <= (base_items < workers ? 8 : 1); ++sp) {
    if (sp > 1 && total_kb / sp < 8) break;
    const int rounds = (base_items * sp + workers - 1) / workers;
    const double cost = static_cast<double>(rounds) / sp + 0.04 * (sp - 1);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = sp; }
  }
  if (forced >= 1 && forced <= 8 && (forced == 1 || total_kb / forced >= 1)) best = forced;
  return best;
}

int gemm_work_items(const GemmArgs& args_in, int cta_group, int num_sms) {
  GemmArgs args = args_in;
  if (cta_group == 2) plan_schedule<2>(args, GEMM_NT, num_sms);
  else plan_schedule<1>(args, GEMM_NT, num_sms);
  const int num_tiles = args.num_m_blocks * args.num_n_blocks;
  return args.full_tiles + (num_tiles - args.full_tiles) * args.tail_split;
}

cudaError_t preload_comm_kernels() {
  cudaFuncAttributes attr;
  return cudaFuncGetAttributes(&attr, comm_signal_extra_kernel);
}

cudaError_t launch_comm_signal_extra(const CommArgs& comm, cudaStream_t stream) {
  comm_signal_extra_kernel<<<1, 32, 0, stream>>>(comm);
  return cudaGetLastError();
}

cudaError_t launch_gemm(const GemmArgs& args, GemmMode mode, GemmOut out, int cta_group, int m_subtiles,
                        int num_sms, cudaStream_t stream) {
  const int num_tiles = args.num_m_blocks * args.num_n_blocks;
  if (num_tiles <= 0) return cudaSuccess;
  if (args.bn < 64 || args.bn > BN || args.bn % 64 != 0) return cudaErrorInvalidValue;
  if (cta_group == 2 && m_subtiles == 2) return launch_cg<2, 2>(args, mode, out, num_sms, stream);
  if (cta_group == 2) return launch_cg<2, 1>(args, mode, out, num_sms, stream);
  return launch_cg<1, 1>(args, mode, out, num_sms, stream);
}

}  // namespace avc
