// Persistent, warp-specialised bf16 GEMM for sm_100a:
//   TMA (128B-swizzled tiles) -> mbarrier ring -> tcgen05.mma (cta_group::1: 128 x BN x 16, or cta_group::2: a CTA pair
//   on a 256 x 256 tile) -> double-buffered TMEM accumulators -> fused epilogue -> per-warp TMA store / fp32 red.add.
//
//   D[b][m, n] = epilogue( sum_k A[b][m, k] * B[b][n, k] )
//
// Both operands may be K-major (rows = M or N, contiguous K) or MN-major (rows = K, contiguous
// M or N) -- the token-mixing GEMMs of the vision-MLP blocks consume [B, N, C] activations
// directly as MN-major operands, so no permute/transposed copy is ever materialised
// (reference: the Conv1d(k=1)-over-tokens trick, models_pytorch/mlp_mixer.py:34,37).
//
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (both run warp-convergent loops of a
// few dozen instructions per k-block and issue under elect.sync), warps 2..17 = epilogue.  A warp may only read the TMEM
// lane quarter (warp_idx % 4), so four warps share each 32-row quarter and own one 64-column chunk of the tile each.
// Every epilogue warp is an independent pipeline: it owns one or two 2 KB staging buffers (32 rows x 32 columns,
// SWIZZLE_64B), issues its own TMA stores, and fetches the auxiliary operand (residual: register prefetch one step ahead;
// saved gelu' / gate: per-warp TMA into the idle staging buffer, L2-prefetched by the producer one tile ahead) -- no
// block-level barrier anywhere in the steady state.  DESIGN.md section 4.1 has the measurements behind each choice.
#pragma once
#include "ptx.cuh"

namespace vmlp {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;       // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_EPI_WARPS = 16;
constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
constexpr int GEMM_STAGE_A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KB
constexpr int GEMM_WARP_STAGING = 32 * 32 * 2;              // 32 rows x 32 bf16 cols = 2 KB (64-byte rows, SWIZZLE_64B)

enum GemmEpilogue : int {
  EPI_STORE = 0,    // D = acc (+bias)                                       -> bf16 via TMA store
  EPI_GELU = 1,     // z = acc + bias; D = gelu'(z) (kept for backward), D2 = gelu(z) -> two bf16 TMA stores
  EPI_RESID = 2,    // D = (acc + bias) * colscale + aux                     -> bf16 (aux = residual)
  EPI_DGELU = 3,    // D = acc * aux                                         -> bf16 (aux = gelu'(z) saved by EPI_GELU)
  EPI_ATOMIC = 4,   // out_f32[m, n] += acc                                  (split-K weight gradients)
  EPI_MUL = 5,      // D = (acc + bias) * aux                                -> bf16 (gMLP spatial gate)
  EPI_GELU_ONLY = 6,  // D = gelu(acc + bias)                                -> bf16 (no pre-activation saved)
  EPI_RESID_DUAL = 7, // EPI_RESID plus D2 = acc + bias (the un-scaled branch output, needed for d(layer-scale))
  EPI_MUL_DUAL = 8    // EPI_MUL plus D2 = acc + bias (the gate value, needed for d(gated operand))
};
__host__ __device__ constexpr bool epi_is_resid(int e) { return e == EPI_RESID || e == EPI_RESID_DUAL; }
__host__ __device__ constexpr bool epi_is_mul(int e) { return e == EPI_MUL || e == EPI_MUL_DUAL; }
__host__ __device__ constexpr bool epi_is_dual(int e) { return e == EPI_GELU || e == EPI_RESID_DUAL || e == EPI_MUL_DUAL; }

struct GemmParams {
  int M, N;              // logical rows / cols of one output matrix (bounds for aux loads / atomics)
  int tiles_m, tiles_n;  // ceil(M/128), ceil(N/BN)
  int batch;             // output batches (1 when K spans the batch)
  int split_k;           // >= 1
  int k_blocks;          // total 64-wide K blocks (summed over the batch when kbatch != 0)
  int kpb;               // K blocks per batch element (== k_blocks unless kbatch != 0)
  int last_ksteps;       // valid 16-wide k-steps in the last block of each kpb group (1..4)
  int a_mn, b_mn;        // operand major-ness (0 = K-major, 1 = MN-major)
  int a_batched, b_batched;
  int kbatch;            // 1: contraction runs over (batch, k) -- token-mixing weight gradients
  int bias_mode;         // 0 none, 1 per output column (n), 2 per output row (m)
  const __nv_bfloat16* bias;
  const __nv_bfloat16* colscale;  // optional per-column scale (ResMLP layer-scale), EPI_RESID only
  const __nv_bfloat16* aux;       // residual / pre-activation / gate operand
  long long aux_ld, aux_bs;       // row stride, batch stride (elements)
  float* out_f32;                 // EPI_ATOMIC destination
  long long out_ld;
  int out_trans;                  // EPI_ATOMIC: element (m, n) is added at out_f32[n * out_ld + m] (the transposed gradient)
  float* red_out;                 // optional fp32 accumulator of the bf16-rounded D: per column (red_mode 1) or per row (2)
  int red_mode;                   //   = bias gradient of the layer whose d(pre-activation) this GEMM produces
  __nv_bfloat16* d2;              // second output of the *_DUAL / GELU epilogues (direct stores)
  long long d2_ld, d2_bs;
  float inv_tiles_n, inv_tiles_m, inv_batch;   // reciprocals for the division-free tile decode (host checks tiles < 2^22)
};

// q = n / d, r = n % d for 0 <= n < 2^22 through the float reciprocal and one correction step (a hardware integer
// division is ~40 instructions; decode_tile runs up to five times per tile per epilogue warp).
__device__ __forceinline__ void fast_divmod(int n, int d, float inv, int& q, int& r) {
  q = __float2int_rz(static_cast<float>(n) * inv);
  r = n - q * d;
  if (r < 0) { r += d; --q; }
  else if (r >= d) { r -= d; ++q; }
}

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) owns a 256 x BN tile;
// each CTA stages its own 128 rows of A and HALF of the B tile, so the per-SM L2->SMEM traffic per MMA drops by a
// third and two more pipeline stages fit.
template <int BN, int EPI, int CG = 1>
struct GemmSmem {
  static constexpr bool DUAL = (EPI == EPI_GELU || EPI == EPI_RESID_DUAL || EPI == EPI_MUL_DUAL);
  // single-output epilogues with an aux operand fetch it by TMA into a second staging buffer and transform it in place
  // (where the aux tensor is as large as the output; the small residual of EPI_RESID keeps the register path and a 4th/6th stage)
  static constexpr bool TMA_AUX = (EPI == EPI_DGELU || EPI == EPI_MUL);
  static constexpr bool TWO_BUF = DUAL || TMA_AUX;
  // dual-output epilogues need a second staging buffer per warp; they are epilogue-bound, so fewer stages are enough
  // the main loop is TMA-latency bound (refill latency ~2800 cycles vs 512 cycles of MMA per stage): every byte of shared
  // memory not needed by the epilogue staging goes to pipeline stages
  static constexpr int STAGES = (CG == 2) ? (TWO_BUF ? 5 : (EPI == EPI_ATOMIC) ? 7 : 6) : (TWO_BUF ? 3 : 4);
  static constexpr int STAGE_B_BYTES = (BN / CG) * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = GEMM_STAGE_A_BYTES + STAGE_B_BYTES;
  static constexpr int PIPE_BYTES = STAGES * STAGE_BYTES;
  static constexpr int STAGING_OFF = PIPE_BYTES;
  static constexpr int STAGING_BYTES = (EPI == EPI_ATOMIC) ? 0 : GEMM_EPI_WARPS * GEMM_WARP_STAGING * (TWO_BUF ? 2 : 1);
  static constexpr int BAR_OFF = STAGING_OFF + STAGING_BYTES;
  static constexpr int AUX_BAR_OFF = BAR_OFF + 192;    // 2 per epilogue warp (TMA_AUX only)
  static constexpr int TOTAL = BAR_OFF + 512 + 1024;   // barriers + slack for 1024B alignment
};

struct TileCoord {
  int m0, n0, b_idx, s_idx;
};
// m0 is THIS CTA's first row (for CG = 2 the pair's tile is 256 rows and cta_rank selects the half)
template <int BN, int CG>
__device__ __forceinline__ TileCoord decode_tile(const GemmParams& p, int tile, int cta_rank) {
  TileCoord c;
  int t, ni, mi;
  fast_divmod(tile, p.tiles_n, p.inv_tiles_n, t, ni);
  fast_divmod(t, p.tiles_m, p.inv_tiles_m, t, mi);
  fast_divmod(t, p.batch, p.inv_batch, c.s_idx, c.b_idx);
  c.n0 = ni * BN;
  c.m0 = mi * (GEMM_BM * CG) + cta_rank * GEMM_BM;
  return c;
}

template <int BN, int EPI, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_sm100(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmD, const __grid_constant__ CUtensorMap tmD2,
                const __grid_constant__ CUtensorMap tmPf, const GemmParams p) {
  using L = GemmSmem<BN, EPI, CG>;
  const int cta_rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;
  const bool is_leader = (cta_rank == 0);
  const int cluster_id = blockIdx.x / CG, num_clusters = gridDim.x / CG;
  constexpr int STAGES = L::STAGES;
  constexpr int NCHUNK = (BN + 63) / 64;   // BN = 208 (token weight gradients, N = 196): the last chunk is 16 columns wide
  constexpr bool HAS_AUX = (epi_is_resid(EPI) || EPI == EPI_DGELU || epi_is_mul(EPI));
  constexpr bool DUAL = epi_is_dual(EPI);
  constexpr bool TMA_AUX = L::TMA_AUX;
  constexpr uint32_t TMEM_COLS = (2 * BN <= 256) ? 256u : 512u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem + L::STAGING_OFF;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + L::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint64_t* aux_bar_base = reinterpret_cast<uint64_t*>(smem + L::AUX_BAR_OFF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (EPI != EPI_ATOMIC) tma_prefetch_desc(&tmD);
    if (DUAL || TMA_AUX) tma_prefetch_desc(&tmD2);   // second output, or (TMA_AUX) the aux operand's map
    if (TMA_AUX) tma_prefetch_desc(&tmPf);           // the aux operand again, one [128 x BN] box per CTA tile (L2 prefetch)
    if (TMA_AUX)
      for (int i = 0; i < 2 * GEMM_EPI_WARPS; ++i) mbar_init(&aux_bar_base[i], 1);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], GEMM_EPI_WARPS * CG);   // the leader's barrier collects both CTAs' epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    // 2 accumulator stages x BN fp32 columns, rounded up to a power of two (256 / 512); for CG = 2 one warp of EACH CTA takes part
    if (CG == 2) { tmem_alloc_2cta(tmem_slot, TMEM_COLS); tmem_relinquish_2cta(); }
    else { tmem_alloc(tmem_slot, TMEM_COLS); tmem_relinquish(); }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_per_split = p.tiles_m * p.tiles_n * p.batch;
  const int total_tiles = tiles_per_split * p.split_k;
  const int kb_per_split = (p.k_blocks + p.split_k - 1) / p.split_k;

  if (warp == 0) {
    // ===================================================================== TMA producer
    // The whole warp walks the pipeline (warp-uniform control flow); one elected lane issues the TMA instructions.
    // The loop body is kept to a few dozen instructions: this warp shares its scheduler with four epilogue warps, and
    // with the previous body (two integer divisions + descriptor rebuilds per k-block, ~130 instructions at IPC ~0.2)
    // it needed more cycles per k-block than the 512 the MMAs of that k-block take (ncu, GELU GEMM: 1165 cycles).
    const uint32_t sbase = smem_u32(smem);
    const uint32_t fb_local = smem_u32(full_bar);
    const uint32_t fb_sig = (CG == 2) ? leader_cta_addr(fb_local) : fb_local;   // where the TMA bytes are signalled
    const uint64_t mapA = reinterpret_cast<uint64_t>(&tmA), mapB = reinterpret_cast<uint64_t>(&tmB);
    const int a_bmask = p.a_batched ? -1 : 0, b_bmask = p.b_batched ? -1 : 0;
    uint32_t stage = 0, phase = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const TileCoord tc = decode_tile<BN, CG>(p, tile, cta_rank);
      const int kb0 = tc.s_idx * kb_per_split;
      const int kb1 = min(kb0 + kb_per_split, p.k_blocks);
      const int nb0 = tc.n0 + cta_rank * (BN / CG);      // this CTA's slice of the B tile
      // running K coordinates: kin = k-block inside its group of kpb blocks, kbat = batch element the group belongs to
      int kbat = tc.b_idx, kin = kb0;
      if (p.kbatch) { kbat = kb0 / p.kpb; kin = kb0 - kbat * p.kpb; }
      // TMA_AUX epilogues stream an aux tensor as large as the output through per-warp 2 KB TMA loads that are issued only
      // one sub-chunk ahead (all the shared memory the staging can get), i.e. with the full HBM latency exposed.  The
      // producer is a whole tile ahead of the epilogue: one L2 prefetch of this CTA's [128 x BN] aux box per tile turns
      // those loads into L2 hits.
      if (TMA_AUX && elect_one_sync()) tma_prefetch_l2_3d(&tmPf, tc.n0, tc.m0, tc.b_idx);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait<64>(&empty_bar[stage], phase ^ 1);   // a long wait in steady state: back off, leave the issue slots to the epilogue warps
        if (elect_one_sync()) {
          const uint32_t sa = sbase + stage * L::STAGE_BYTES;
          const uint32_t sb = sa + GEMM_STAGE_A_BYTES;
          const uint32_t bar = fb_sig + stage * 8;
          // CG = 2: both CTAs' TMA loads complete_tx on the LEADER's full barrier, which expects both stages' bytes
          if (is_leader) mbar_arrive_expect_tx_u32(fb_local + stage * 8, L::STAGE_BYTES * CG);
          const int kc = kin * GEMM_BK;
          const int ba = kbat & a_bmask, bb = kbat & b_bmask;
          if (!p.a_mn) {
            tma_load_3d_u32<CG>(sa, mapA, bar, kc, tc.m0, ba);                            // box (64 k, 128 m)
          } else {
#pragma unroll
            for (int a = 0; a < GEMM_BM / 64; ++a)                                        // box (64 m, 64 k) per MN atom
              tma_load_3d_u32<CG>(sa + a * (GEMM_BK * 128), mapA, bar, tc.m0 + a * 64, kc, ba);
          }
          if (!p.b_mn) {
            tma_load_3d_u32<CG>(sb, mapB, bar, kc, nb0, bb);                              // box (64 k, BN / CG n)
          } else {
#pragma unroll
            for (int a = 0; a < BN / CG / 64; ++a)
              tma_load_3d_u32<CG>(sb + a * (GEMM_BK * 128), mapB, bar, nb0 + a * 64, kc, bb);
          }
        }   // elected lane
        __syncwarp();
        if (++kin == p.kpb) { kin = 0; ++kbat; }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // Whole warp, warp-uniform control flow; one elected lane issues the MMAs and commits.
    if (is_leader) {
      const uint32_t idesc = umma_idesc_bf16(GEMM_BM * CG, BN, p.a_mn, p.b_mn);
      // Shared-memory descriptors (umma_smem_desc_sw128): the high word is a constant, the low word is
      // (address >> 4) | (LBO >> 4) << 16, so stepping a stage or a k-step is one integer add on the low word.
      // K-major : 8-row groups 1024 B apart (SBO); one 128B swizzle atom along K (LBO unused).
      // MN-major: 8-k groups 1024 B apart (SBO); 64-wide MN atoms BK*128 B apart (LBO).
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);
      const uint32_t sbase = smem_u32(smem);
      const uint32_t a_lo0 = (sbase >> 4) | ((p.a_mn ? (GEMM_BK * 128u) >> 4 : 0u) << 16);
      const uint32_t b_lo0 = ((sbase + GEMM_STAGE_A_BYTES) >> 4) | ((p.b_mn ? (GEMM_BK * 128u) >> 4 : 0u) << 16);
      const uint32_t a_kinc = (p.a_mn ? 16u * 128u : 32u) >> 4, b_kinc = (p.b_mn ? 16u * 128u : 32u) >> 4;
      uint32_t stage = 0, phase = 0;
      int tc = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tc) {
        const int s_idx = decode_tile<BN, CG>(p, tile, 0).s_idx;
        const int kb0 = s_idx * kb_per_split;
        const int kb1 = min(kb0 + kb_per_split, p.k_blocks);
        int kin = p.kbatch ? (kb0 % p.kpb) : kb0;
        const int as = tc & 1;
        mbar_wait<32>(&tmem_empty[as], ((tc >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * (L::STAGE_BYTES >> 4);
          const uint32_t b_lo = b_lo0 + stage * (L::STAGE_BYTES >> 4);
          const int ksteps = (kin == p.kpb - 1) ? p.last_ksteps : (GEMM_BK / 16);
          if (elect_one_sync()) {
            umma_bf16_lo<CG>(d_tmem, a_lo, b_lo, DESC_HI, idesc, kb > kb0 ? 1u : 0u);
#pragma unroll
            for (int k = 1; k < GEMM_BK / 16; ++k)
              if (k < ksteps) umma_bf16_lo<CG>(d_tmem, a_lo + k * a_kinc, b_lo + k * b_kinc, DESC_HI, idesc, 1u);
            // frees the smem slot (in both CTAs for CG = 2) when these MMAs retire; last k-block: accumulator complete
            if (CG == 2) {
              umma_commit_2cta_mc(&empty_bar[stage]);
              if (kb == kb1 - 1) umma_commit_2cta_mc(&tmem_full[as]);
            } else {
              umma_commit(&empty_bar[stage]);
              if (kb == kb1 - 1) umma_commit(&tmem_full[as]);
            }
          }   // elected lane
          __syncwarp();
          if (++kin == p.kpb) kin = 0;
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..17)
    // 16 warps = 4 per scheduler: the epilogue math (GELU / GELU') is latency-bound with fewer.  Warp (q, c) owns the
    // 32 accumulator rows of TMEM lane quarter q = warp % 4 and the 64-column chunk c = (warp - 2) / 4 of every tile, walks
    // it in four 16-column steps (tcgen05.ld.x16 keeps the live registers < 113), stages each 32-column half of the bf16
    // result(s) in its private 2 KB 64B-swizzled buffer(s) and issues its own TMA store(s).
    const int ew = warp - 2;                // 0..15
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int c = ew >> 2;                  // 64-column chunk owned by this warp
    const bool has_chunk = c < NCHUNK;
    const int row = q * 32 + lane;          // accumulator row owned by this thread
    uint8_t* st0 = staging + ew * GEMM_WARP_STAGING * (L::TWO_BUF ? 2 : 1);
    uint8_t* st1 = st0 + GEMM_WARP_STAGING;
    uint64_t* aux_bar = aux_bar_base + ew * 2;
    const bool store_lane = elect_one_sync() != 0;   // this lane owns the warp's TMA-store bulk groups for the whole kernel

    // ---- auxiliary-operand prefetch (registers): 16 columns of this thread's row = 2 x 16 B, one step ahead
    uint4 aux_nxt[2];
    auto load_aux = [&](int tile, int step, uint4 (&dst)[2]) {
      dst[0] = make_uint4(0, 0, 0, 0);
      dst[1] = make_uint4(0, 0, 0, 0);
      if (!HAS_AUX || TMA_AUX || !has_chunk || tile >= total_tiles) return;
      const TileCoord t = decode_tile<BN, CG>(p, tile, cta_rank);
      const int grow = t.m0 + row;
      const int col = t.n0 + c * 64 + step * 16;
      if (grow >= p.M) return;
      const __nv_bfloat16* src = p.aux + (long long)t.b_idx * p.aux_bs + (long long)grow * p.aux_ld + col;
      if (col < p.N) dst[0] = ldg_v4(src);
      if (col + 8 < p.N) dst[1] = ldg_v4(src + 8);
    };
    if (HAS_AUX && !TMA_AUX) load_aux(cluster_id, 0, aux_nxt);
    // ---- TMA path: the aux sub-chunk (32 rows x 32 cols) of this warp's NEXT work item lands in the idle staging buffer
    // a work item (this warp's 32 rows x 64 columns of a tile) is dead when it lies entirely outside the output: ragged N
    // (columns) or ragged M (token GEMMs: M = 784 = 6.125 tiles -> three of the last tile's four row quarters are padding)
    auto item_live = [&](const TileCoord& t) {
      return has_chunk && (t.n0 + c * 64 < p.N) && (t.m0 + q * 32 < p.M);
    };
    auto chunk_is_live = [&](int tile) { return item_live(decode_tile<BN, CG>(p, tile, cta_rank)); };
    auto next_live_tile = [&](int tile) {
      while (tile < total_tiles && !chunk_is_live(tile)) tile += num_clusters;
      return tile;
    };
    auto issue_aux = [&](int tile, int half, uint32_t subidx) {
      if (tile >= total_tiles || !store_lane) return;
      const TileCoord t = decode_tile<BN, CG>(p, tile, cta_rank);
      const uint32_t b = subidx & 1;
      mbar_arrive_expect_tx(&aux_bar[b], GEMM_WARP_STAGING);
      tma_load_3d(st0 + b * GEMM_WARP_STAGING, &tmD2, &aux_bar[b], t.n0 + c * 64 + half * 32, t.m0 + q * 32, t.b_idx);
    };
    uint32_t sub = 0;                      // sub-chunks done by this warp: selects buffer (sub & 1) and barrier parity
    if (TMA_AUX) issue_aux(next_live_tile(cluster_id), 0, 0);

    int tcnt = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tcnt) {
      const TileCoord tc = decode_tile<BN, CG>(p, tile, cta_rank);
      const int grow = tc.m0 + row;
      const bool row_ok = grow < p.M;
      const int as = tcnt & 1;
      mbar_wait(&tmem_full[as], (tcnt >> 1) & 1);
      tc_fence_after();
      float rbias = 0.f;
      if (p.bias_mode == 2 && row_ok) rbias = __bfloat162float(p.bias[grow]);
      const int col0 = tc.n0 + c * 64;
      const bool chunk_live = item_live(tc);               // dead work items are skipped entirely
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        uint4 aux_cur[2];
        if (HAS_AUX && !TMA_AUX) {
          aux_cur[0] = aux_nxt[0];
          aux_cur[1] = aux_nxt[1];
          if (st < 3) load_aux(tile, st + 1, aux_nxt);
          else load_aux(tile + num_clusters, 0, aux_nxt);
        }
        uint8_t* bufp = st0 + (TMA_AUX ? (sub & 1) * GEMM_WARP_STAGING : 0);   // staging buffer of this sub-chunk
        if (chunk_live && EPI != EPI_ATOMIC && (st & 1) == 0) {
          if (TMA_AUX) {
            mbar_wait(&aux_bar[sub & 1], (sub >> 1) & 1);       // aux sub-chunk has landed in bufp
          } else {
            // the 2 KB staging buffer(s) must have been drained by the previous 32-column TMA store of this warp
            if (store_lane) tma_store_wait_read<0>();
            __syncwarp();
          }
        }
        uint32_t v[16];
        if (chunk_live) {
          tmem_ld_32x32b_x16(tmem_base + as * BN + c * 64 + st * 16 + (uint32_t(q * 32) << 16), v);
          tmem_ld_wait();
        }
        if (st == 3) {
          // last TMEM read of this accumulator stage by this warp: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (is_leader) mbar_arrive(&tmem_empty[as]);
            else mbar_arrive_remote(&tmem_empty[as], 0);     // the MMA issuer lives in the leader CTA
          }
        }
        if (!chunk_live) continue;
        const int cols = col0 + st * 16;
        if (EPI == EPI_ATOMIC) {
          if (row_ok) {
            if (!p.out_trans) {
              float* orow = p.out_f32 + (long long)grow * p.out_ld;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cols + j < p.N) red_add_f32(orow + cols + j, __uint_as_float(v[j]));
            } else {          // lanes = consecutive rows m: each RED of the warp covers 32 consecutive floats
              float* ocol = p.out_f32 + (long long)cols * p.out_ld + grow;
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (cols + j < p.N) red_add_f32(ocol + (long long)j * p.out_ld, __uint_as_float(v[j]));
            }
          }
          continue;
        }
        float f[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(v[j]) + rbias;
        if (p.bias_mode == 1) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (cols + g * 8 < p.N) {
              const uint4 bv = *reinterpret_cast<const uint4*>(p.bias + cols + g * 8);
              const uint32_t w[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[g * 8 + 2 * e] += bf16lo(w[e]);
                f[g * 8 + 2 * e + 1] += bf16hi(w[e]);
              }
            }
          }
        }
        uint32_t o[8], o2[8];
        if ((epi_is_resid(EPI) || epi_is_mul(EPI)) && DUAL) {
#pragma unroll
          for (int e = 0; e < 8; ++e) o2[e] = pack_bf16x2(f[2 * e], f[2 * e + 1]);   // branch output before scale / gate
        }
        if (epi_is_resid(EPI) && p.colscale != nullptr) {
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            if (cols + g * 8 < p.N) {
              const uint4 sv = *reinterpret_cast<const uint4*>(p.colscale + cols + g * 8);
              const uint32_t w[4] = {sv.x, sv.y, sv.z, sv.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                f[g * 8 + 2 * e] *= bf16lo(w[e]);
                f[g * 8 + 2 * e + 1] *= bf16hi(w[e]);
              }
            }
          }
        }
        if (HAS_AUX) {
          if (TMA_AUX) {
#pragma unroll
            for (int g = 0; g < 2; ++g)
              aux_cur[g] = ld_shared_v4(smem_u32(bufp) + lane * 64 + ((((st & 1) * 2 + g) ^ ((lane >> 1) & 3)) * 16));
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const uint32_t w[4] = {aux_cur[g].x, aux_cur[g].y, aux_cur[g].z, aux_cur[g].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float x0 = f[g * 8 + 2 * e], x1 = f[g * 8 + 2 * e + 1];
              const float a0 = bf16lo(w[e]), a1 = bf16hi(w[e]);
              if (epi_is_resid(EPI)) { x0 += a0; x1 += a1; }
              if (epi_is_mul(EPI)) { x0 *= a0; x1 *= a1; }
              if (EPI == EPI_DGELU) { x0 *= a0; x1 *= a1; }
              o[g * 4 + e] = pack_bf16x2(x0, x1);
            }
          }
        } else if (EPI == EPI_GELU || EPI == EPI_GELU_ONLY) {
          // packed fp32x2 path (bias add included; f[] above is dead code for these epilogues).  One erf/exp evaluation
          // yields both gelu(z) and gelu'(z); backward then only multiplies by the saved gelu'.
          f32x2 z[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) z[e] = pack2(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          if (p.bias_mode == 2) {
            const f32x2 rb = pack2(rbias, rbias);
#pragma unroll
            for (int e = 0; e < 8; ++e) z[e] = add2(z[e], rb);
          } else if (p.bias_mode == 1) {
#pragma unroll
            for (int g = 0; g < 2; ++g) {
              if (cols + g * 8 < p.N) {
                const uint4 bv = *reinterpret_cast<const uint4*>(p.bias + cols + g * 8);
                const uint32_t w[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) z[g * 4 + e] = add2(z[g * 4 + e], pack2(bf16lo(w[e]), bf16hi(w[e])));
              }
            }
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            f32x2 gl, dg;
            gelu_erf_pair<EPI == EPI_GELU>(z[e], gl, dg);
            if (EPI == EPI_GELU) {
              o[e] = pack_bf16x2_f2(dg);
              o2[e] = pack_bf16x2_f2(gl);
            } else {
              o[e] = pack_bf16x2_f2(gl);
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] = pack_bf16x2(f[2 * e], f[2 * e + 1]);
        }
        if (p.red_mode == 2) {
          // bias gradient fused into the epilogue (token mixing: bias index = output row): sum of the bf16-rounded values
          // this thread stores (N is even, so validity is per bf16 pair).
          if (row_ok) {
            float rs = 0.f;
            const int nv = min(8, max(0, (p.N - cols) >> 1));
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (e < nv) rs += bf16lo(o[e]) + bf16hi(o[e]);
            red_add_f32(p.red_out + grow, rs);
          }
        }
        // staging tile = [32 rows][64 B]; 16-byte chunk c of row r lives at c ^ ((r >> 1) & 3) (matches SWIZZLE_64B)
        const uint32_t srow = lane * 64;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          const int chunk = ((st & 1) * 2 + g) ^ ((lane >> 1) & 3);
          st_shared_v4(smem_u32(bufp) + srow + chunk * 16, make_uint4(o[g * 4], o[g * 4 + 1], o[g * 4 + 2], o[g * 4 + 3]));
          if (DUAL)
            st_shared_v4(smem_u32(st1) + srow + chunk * 16,
                         make_uint4(o2[g * 4], o2[g * 4 + 1], o2[g * 4 + 2], o2[g * 4 + 3]));
        }
        if (st & 1) {
          // a 32-column half of the chunk is complete: hand it to the TMA engine
          fence_proxy_async_smem();
          __syncwarp();
          if (store_lane) {
            const int colh = col0 + (st >> 1) * 32;
            if (colh < p.N) {
              tma_store_3d(&tmD, bufp, colh, tc.m0 + q * 32, tc.b_idx);
              if (DUAL) tma_store_3d(&tmD2, st1, colh, tc.m0 + q * 32, tc.b_idx);
            }
            tma_store_commit();
            // TMA_AUX: the OTHER buffer's store is the previous bulk group -> once it is drained, refill it with the aux
            // operand of this warp's next sub-chunk (second half of this chunk, or first half of the next live tile)
            if (TMA_AUX) tma_store_wait_read<1>();
          }
          if (p.red_mode == 1) {
            // bias gradient fused into the epilogue (channel mixing: bias index = output column).  The finished 32 x 32
            // sub-chunk sits in the staging buffer: lane l sums column pair (l & 15) over 16 rows -- one conflict-free
            // LDS.32 + 2 unpacks + 2 adds per row, ~3 instructions per element where the register butterfly this
            // replaces (31 SHFL + 62 FSEL per 16 columns) cost ~10 and made the DGELU GEMMs issue-bound.
            const int pr = lane & 15, hh = lane >> 4;
            const int kq = pr >> 2;
            // row r = hh * 16 + (i ^ hh): the upper half-warp walks rows of the opposite parity, i.e. the other 16 banks
            const uint32_t base_even = smem_u32(bufp) + hh * 1024 + hh * 64 + (pr & 3) * 4;
            const uint32_t base_odd = smem_u32(bufp) + hh * 1024 - hh * 64 + (pr & 3) * 4;
            const uint32_t kx[4] = {uint32_t(kq) * 16, uint32_t(kq ^ 1) * 16, uint32_t(kq ^ 2) * 16, uint32_t(kq ^ 3) * 16};
            const int nrows = min(32, p.M - (tc.m0 + q * 32));      // > 0: the work item is live
            float s0 = 0.f, s1 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const uint32_t addr = ((i & 1) ? base_odd : base_even) + i * 64 + kx[(i >> 1) & 3];
              uint32_t w;
              asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(addr) : "memory");
              if (nrows == 32 || hh * 16 + (i ^ hh) < nrows) {
                s0 += bf16lo(w);
                s1 += bf16hi(w);
              }
            }
            s0 += __shfl_xor_sync(0xffffffffu, s0, 16);
            s1 += __shfl_xor_sync(0xffffffffu, s1, 16);
            const int colp = col0 + (st >> 1) * 32 + 2 * pr;
            if (hh == 0 && colp < p.N) {
              red_add_f32(p.red_out + colp, s0);
              red_add_f32(p.red_out + colp + 1, s1);
            }
          }
          if (TMA_AUX) {
            if (st == 1) issue_aux(tile, 1, sub + 1);
            else issue_aux(next_live_tile(tile + num_clusters), 0, sub + 1);
            ++sub;
          }
        }
      }
    }
    if (EPI != EPI_ATOMIC && store_lane) tma_store_wait_all<0>();
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // neither CTA may exit while its peer still uses it
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2cta(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace vmlp
