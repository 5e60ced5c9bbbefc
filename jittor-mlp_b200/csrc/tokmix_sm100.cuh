// Fused token-mixing MLP of the MLP-Mixer block for sm_100a (reference: models_pytorch/mlp_mixer.py:16-27,34,37 --
// FeedForward(num_patches, 4 * num_patches, dense = Conv1d(k=1)) applied along the token axis of [B, N, C]).
//
//   forward :  U[b]  = X[b] + W2 * gelu(W1 * Xh[b] + b1) + b2          Xh = LN1(X) [B, N, C],  W1 [Ds, N],  W2 [N, Ds]
//   backward:  dXh[b] = W1^T * ( (W2^T * dU[b]) .* gelu'(W1 * Xh[b] + b1) )     (+ dZ^T saved for the weight gradient)
//
// The hidden tensor Z/H [B, Ds, C] never makes an HBM round trip between the two GEMMs.  Work item of one CTA = one image
// and 128 channels ("tile"); a CTA pair (cluster of 2, tcgen05 cta_group::2, M = 256) shares every weight chunk: each
// CTA loads half of the chunk's rows, so the L2 -> SMEM weight traffic per SM is half of what two independent CTAs need.
// Everything is computed TRANSPOSED (channels = M = TMEM lanes):
//
//   G1:  Z^T[c, m]   = sum_n Xh^T[c, n] * W1[m, n]       A = the [N, 128c] activation tile as it lies in memory (MN-major),
//                                                         B = 64 rows of W1 (K-major), accumulator 64 TMEM columns
//   epi: H^T[c, m]   = gelu(Z^T + b1[m])  -> bf16 -> SMEM tile [128c x 64m] (K-major A operand of G2) (+ TMA store, fwd)
//   G2:  U^T[c, n]  += sum_m H^T[c, m] * W2[n, m]         B = [N rows x 64 m] of W2 (K-major), accumulator NT columns
//
// and in backward  G1 (Z recomputed), G2: dH^T = dU^T * W2 chunk, epi: dZ^T = dH^T .* gelu'(Z^T), G3: dXh^T += dZ^T * W1 chunk.
// Hidden chunks of 64 divide Ds = 784 with 16 left over (UMMA N = 16 for the tail chunk): no padded MMA work along Ds;
// the token axis N = 196 is padded to NT = 208 by TMA zero fill (13 k-steps of 16).
//
// Warp roles, forward (384 threads): warp 0 = TMA producer of the weight rings and the activation tile, warp 1 = issuer of
// G1 + TMEM owner, warp 3 = issuer of G2, warp 2 = forwarder of the epilogue's hand-offs + TMA stores (hidden tile,
// residual-in / output-out tile), warps 4..11 = epilogue in two ping-pong groups of 4 (TMEM lane quarter = warp % 4, all
// 64 columns of every other chunk).  Backward (384 threads): warp 0 = producer, warp 1 = the MMA issuer (G1, G2 and, two
// chunks behind, G3), warp 2 = TMA store of the dZ tile, warps 2..3 = its column sums (d b1), warps 4..11 = epilogue
// (lane quarter = warp % 4, 32 of the chunk's 64 columns).  All hand-offs between roles are mbarriers (multicast
// tcgen05.commit towards both CTAs, remote arrives towards the leader); inside an epilogue group, hardware named barriers.
//
// What the versions taught (clock64 timeline of CTA 0, tools/tokmix_trace.py; ncu stall reasons and source page):
//  * every mbarrier operation costs a warp 100-200 cycles even when it succeeds at once: ONE issuer warp that waits on
//    five barriers and issues 17 MMAs per chunk needs ~1900 cycles for 832 cycles of tensor work, and 16 epilogue warps
//    that each wait / load / arrive / compute / wait / write / fence / arrive in lock step need ~1100 cycles of pure
//    hand-off latency per chunk -> forward: two issuer warps, one polling warp per group + named barriers, one forwarding
//    warp, two groups on alternate chunks;
//  * the GELU is a 14-deep (gelu': ~20-deep) dependent chain per element pair; with 640 threads ptxas had 96 registers
//    and serialised the pairs (ncu: fixed-latency "wait" the top stall, FMA pipe 41 % busy, issue slots 50 % used)
//    -> 8 epilogue warps with 168 registers instead of 16 with 96: same lanes, several pairs in flight per warp;
//  * erf-GELU with a single MUFU per element (ptx.cuh: gelu_rcp16_xn) -- rcp + ex2 per element is 1024 MUFU cycles per
//    chunk and SM, more than the chunk's 832 MMA cycles.
#pragma once
#include "ptx.cuh"

namespace vmlp {

constexpr int TM_CH = 64;                              // hidden chunk (UMMA N of G1, K of G2 per chunk)
// Registers are allocated to warps in groups of four: 20 warps leave 96 registers per thread, 21..24 warps leave 80.
constexpr int TM_FWD_EPI0 = 4;                                   // forward: 4 service warps + 8 epilogue warps
constexpr int TM_FWD_EPI_WARPS = 8;                              // 384 threads -> up to 168 registers each (see the epilogue)
constexpr int TM_FWD_THREADS = 32 * (TM_FWD_EPI0 + TM_FWD_EPI_WARPS);
constexpr int TM_NB_START = 1, TM_NB_ZE = 3, TM_NB_HW = 5, TM_NB_DZ = 7;          // named (hardware) barrier ids (+ group), see the epilogue
constexpr int TM_BWD_EPI0 = 4;                                   // backward: 4 service warps + 8 epilogue warps
constexpr int TM_BWD_EPI_WARPS = 8;
constexpr int TM_BWD_THREADS = 32 * (TM_BWD_EPI0 + TM_BWD_EPI_WARPS);
constexpr int TM_HTILE = 128 * TM_CH * 2;              // 16 KB: [128 channels x 64 hidden] bf16, K-major SWIZZLE_128B
constexpr int TM_MAX_DS = 1024;
constexpr int TM_BAR_BYTES = 1024;

struct TokParams {
  int B, N, NT, C, Ds;       // NT = N rounded up to 16 (UMMA K-steps along tokens, UMMA N of the output GEMM)
  int tiles_c;               // ceil(C / 128)
  int n_tiles;               // B * tiles_c
  int n_pairs;               // ceil(n_tiles / 2)
  int n_chunks;              // ceil(Ds / 64)
  int last_n1;               // hidden columns of the last chunk, rounded up to 16
  int ka;                    // 64-wide K atoms along the token axis: ceil(NT / 64) (the last one may be partly used)
  int nhb;                   // hidden-tile buffers in shared memory: 2, or 1 when two do not fit (backward, NT = 208)
  int wa_stage;              // bytes of one [32 rows x NT k] weight stage (G1 / G2-of-backward B operand): ka * 4 KB
  int wb_stage;              // bytes of one [NT/2 rows x 64 k] weight stage (output-GEMM B operand)
  int s_wa, s_wb;            // ring depths
  float inv_tiles_c;
  const __nv_bfloat16* b1;   // [Ds]
  const __nv_bfloat16* b2;   // [N]         (forward)
  const __nv_bfloat16* resid;// [B, N, C]   (forward: x)
  __nv_bfloat16* out;        // [B, N, C]   forward: u; backward: dXh
  float* db1;                // [Ds] fp32   backward: += sum over (b, c) of dZ (the hidden-bias gradient)
  long long* trace;          // optional timeline of CTA 0 (clock64 stamps, bring-up only): [role 0..3][chunk < 64][8 events]
  int flags;                 // profiling experiments only (VMLP_TM_FLAGS): 1 no GELU math, 2 no hidden-tile TMA store,
                             // 4 no column sums, 8 no hidden-tile SMEM write, 16 no TMEM load, 32 no weight TMA loads
                             // -- results are then WRONG; 64 chunk rotation off (results unchanged)
};

// shared-memory descriptor high word (SBO = 1024 B between 8-row groups, version 1, SWIZZLE_128B); K-major swizzled
// operands carry LBO = 1 (16 bytes), the canonical ((8,n),2):((8,SBO),1) form.  Every operand here is SWIZZLE_128B: a
// SWIZZLE_32B tail atom for the 16 left-over token columns (NT = 208 = 3 x 64 + 16) produced wrong products and, behind
// three full atoms, a launch failure on the first two bring-up runs -- the tail is a fourth 64-wide atom instead, of
// which only the first k-step is issued.
constexpr uint32_t TM_DESC_HI_SW128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t TM_LBO_K = 1u << 16;

__device__ __forceinline__ void tm_arrive_leader(uint64_t* bar, bool is_leader) {
  if (is_leader) mbar_arrive(bar);
  else mbar_arrive_remote(bar, 0);
}
__device__ __forceinline__ uint32_t ldg_u16(const __nv_bfloat16* p) {
  unsigned short v;
  asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_u16(__nv_bfloat16* p, uint32_t v) {
  asm volatile("st.global.u16 [%0], %1;" ::"l"(p), "h"(static_cast<unsigned short>(v)) : "memory");
}
__device__ __forceinline__ uint32_t bf16_bits(float f) {
  __nv_bfloat16 h = __float2bfloat16_rn(f);
  return *reinterpret_cast<unsigned short*>(&h);
}

// TMA loads of one [32 rows x NT k] weight stage (K-major): ka SWIZZLE_128B atoms of [32 rows x 64 k]
__device__ __forceinline__ void tm_load_wa(uint32_t dst, uint64_t map128, uint32_t bar, int row, const TokParams& p) {
  for (int a = 0; a < p.ka; ++a) tma_load_3d_u32<2>(dst + a * 4096, map128, bar, a * 64, row, 0);
}
// G1-type MMA: D[tmem] = A (MN-major activation tile, k-steps over the token axis) * B (weight stage); n1 = UMMA N
__device__ __forceinline__ void tm_mma_over_tokens(uint32_t d_tmem, uint32_t act_addr, uint32_t w_addr, uint32_t idesc,
                                                   const TokParams& p) {
  const uint32_t a_lbo = (static_cast<uint32_t>(p.NT) * 128u) >> 4;      // distance between the two 64-channel atoms
  uint32_t a_lo = (act_addr >> 4) | (a_lbo << 16);
  const uint32_t b_base = (w_addr >> 4) | TM_LBO_K;
  const int ksteps = p.NT >> 4;
  for (int ks = 0; ks < ksteps; ++ks) {
    umma_bf16_lo<2>(d_tmem, a_lo, b_base + (ks >> 2) * 256 + (ks & 3) * 2, TM_DESC_HI_SW128, idesc, ks ? 1u : 0u);
    a_lo += 2048u >> 4;                                                   // 16 token rows of 128 B
  }
}
// G2/G3-type MMA: D[tmem] (+)= A (hidden tile in SMEM, K-major) * B ([NT/2 rows x 64 k] weight stage), ksteps of 16
__device__ __forceinline__ void tm_mma_over_hidden(uint32_t d_tmem, uint32_t h_addr, uint32_t w_addr, uint32_t idesc,
                                                   int ksteps, bool first) {
  const uint32_t a_base = (h_addr >> 4) | TM_LBO_K, b_base = (w_addr >> 4) | TM_LBO_K;
  for (int ks = 0; ks < ksteps; ++ks)
    umma_bf16_lo<2>(d_tmem, a_base + ks * 2, b_base + ks * 2, TM_DESC_HI_SW128, idesc, (first && ks == 0) ? 0u : 1u);
}

// Hidden chunks are visited in a per-cluster ROTATED order (chunk index = position + cluster_id, mod n_chunks): the sum
// over chunks is order-free, and 74 CTA pairs no longer stream the same 30 KB of W1 / W2 out of the same L2 lines at the
// same moment.  VMLP_TM_FLAGS bit 64 switches the rotation off (A/B measurements).
__device__ __forceinline__ int tm_chunk(int pos, int rot, int nc) {
  const int j = pos + rot;
  return j < nc ? j : j - nc;
}

struct TokTile {
  int b, c0;
  bool valid;
};
// bring-up timeline: one elected/first lane of a role in CTA 0 stamps clock64() at event `ev` of global chunk `g`
__device__ __forceinline__ void tm_stamp(const TokParams& p, int role, int g, int ev) {
  if (p.trace != nullptr && blockIdx.x == 0 && g < 64) p.trace[(role * 64 + g) * 8 + ev] = clock64();
}
__device__ __forceinline__ TokTile tm_tile(const TokParams& p, int pair, int cta_rank) {
  TokTile t;
  const int tile = 2 * pair + cta_rank;
  t.valid = tile < p.n_tiles;
  int b = __float2int_rz(static_cast<float>(tile) * p.inv_tiles_c);
  int ct = tile - b * p.tiles_c;
  if (ct < 0) { ct += p.tiles_c; --b; }
  else if (ct >= p.tiles_c) { ct -= p.tiles_c; ++b; }
  t.b = t.valid ? b : p.B;          // a dead CTA (odd tile count) loads zero-filled boxes and stores nothing
  t.c0 = ct * 128;
  return t;
}

// non-suspending probe of an mbarrier phase (the suspending try_wait is used once the probe has failed)
__device__ __forceinline__ uint32_t mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
// wait for two barriers: both probes are in flight together (one ~150-cycle latency instead of two in sequence)
template <int SLEEP_NS>
__device__ __forceinline__ void mbar_wait2(uint64_t* a, uint32_t pa, uint64_t* b, uint32_t pb) {
  const uint32_t oa = mbar_test(a, pa), ob = mbar_test(b, pb);
  if (!oa) mbar_wait<SLEEP_NS>(a, pa);
  if (!ob) mbar_wait<SLEEP_NS>(b, pb);
}
// tcgen05.wait::ld that also names the 16 registers an earlier (prefetching) tcgen05.ld targets: their first use must not
// be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// One epilogue thread's 16 packed bf16 values -> its row of the [128 x 64] SWIZZLE_128B hidden tile
__device__ __forceinline__ void tm_store_hidden_row(uint32_t tile_addr, int row, int cq, const uint32_t (&o)[8]) {
  const uint32_t base = tile_addr + (row >> 3) * 1024 + (row & 7) * 128;
  const int sw = row & 7;
  st_shared_v4(base + (((2 * cq) ^ sw) << 4), make_uint4(o[0], o[1], o[2], o[3]));
  st_shared_v4(base + (((2 * cq + 1) ^ sw) << 4), make_uint4(o[4], o[5], o[6], o[7]));
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}

// ------------------------------------------------------------------------------------------------------------------
// Forward.  SMEM: [barriers 1 KB][Xh^T tile NT*256][residual/output tile NT*256][W1 ring][W2 ring][H tiles][b1][b2]
// TMEM: Z double buffer at columns [0, 128), U accumulator at [128, 128 + NT).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TM_FWD_THREADS, 1)
tokmix_fwd_sm100(const __grid_constant__ CUtensorMap tmX,      // Xh   [B, N, C]   box (64 c, NT rows)
                 const __grid_constant__ CUtensorMap tmW1,     // W1   [Ds, NT]    box (64 k, 32 rows) SWIZZLE_128B
                 const __grid_constant__ CUtensorMap tmW2,     // W2   [N, Ds]     box (64 k, NT/2 rows)
                 const __grid_constant__ CUtensorMap tmH,      // H^T  [B, C, Ds]  box (64 m, 128 c)  (saved for backward)
                 const __grid_constant__ CUtensorMap tmR,      // x    [B, N, C]   box (128 c, NT rows), no swizzle (residual in)
                 const __grid_constant__ CUtensorMap tmU,      // u    [B, N, C]   box (128 c, NT rows), no swizzle (output)
                 const TokParams p, const int save_hidden) {
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const bool is_leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* xt_full = bars + 0;     uint64_t* xt_empty = bars + 1;
  uint64_t* u_full = bars + 2;      uint64_t* u_empty = bars + 3;
  uint64_t* z_full = bars + 4;      uint64_t* z_empty = bars + 6;      // [2] each
  uint64_t* h_full = bars + 8;      uint64_t* h_free = bars + 11;      // [3] each; h_free: G2 and the TMA store have read the tile
  uint64_t* wa_full = bars + 16;    uint64_t* wa_empty = bars + 24;    // up to 8 stages each
  uint64_t* wb_full = bars + 32;    uint64_t* wb_empty = bars + 40;
  uint64_t* ro_full = bars + 48;    uint64_t* ro_done = bars + 49;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 56);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_xt = s_base + TM_BAR_BYTES;
  const uint32_t s_ro = s_xt + p.NT * 256;          // [NT tokens][128 channels] bf16, row-major: residual in, output out
  const uint32_t s_wa = s_ro + p.NT * 256;
  const uint32_t s_wb = s_wa + p.s_wa * p.wa_stage;
  const uint32_t s_h = s_wb + p.s_wb * p.wb_stage;
  const uint32_t s_b1 = s_h + p.nhb * TM_HTILE;     // fp32 biases
  const uint32_t s_b2 = s_b1 + p.n_chunks * TM_CH * 4;
  float* sb1 = reinterpret_cast<float*>(smem + (s_b1 - s_base));
  float* sb2 = reinterpret_cast<float*>(smem + (s_b2 - s_base));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW1); tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmH); tma_prefetch_desc(&tmR); tma_prefetch_desc(&tmU);
    mbar_init(xt_full, 1); mbar_init(xt_empty, 1);
    mbar_init(ro_full, 1); mbar_init(ro_done, TM_FWD_EPI_WARPS);
    mbar_init(u_full, 1);  mbar_init(u_empty, 2 * TM_FWD_EPI_WARPS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&z_full[i], 1);  mbar_init(&z_empty[i], TM_FWD_EPI_WARPS);   // 4 warps of the owning group x 2 CTAs
    }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&h_full[i], 2);  mbar_init(&h_free[i], 2);       // one forwarded arrival per CTA; G2 commit + TMA store
    }
    for (int i = 0; i < 8; ++i) {
      mbar_init(&wa_full[i], 1); mbar_init(&wa_empty[i], 1);
      mbar_init(&wb_full[i], 1); mbar_init(&wb_empty[i], 1);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.n_chunks * TM_CH; i += TM_FWD_THREADS) sb1[i] = i < p.Ds ? __bfloat162float(p.b1[i]) : 0.f;
  for (int i = threadIdx.x; i < p.NT; i += TM_FWD_THREADS) sb2[i] = i < p.N ? __bfloat162float(p.b2[i]) : 0.f;
  if (warp == 1) { tmem_alloc_2cta(tmem_slot, 512); tmem_relinquish_2cta(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int NC = p.n_chunks;
  const int rot = (p.flags & 64) ? 0 : cluster_id % NC;
  int my_items = 0;
  for (int pair = cluster_id; pair < p.n_pairs; pair += num_clusters) ++my_items;
  const int total = my_items * NC;                  // hidden chunks this cluster processes

  if (warp == 0) {
    // ================================================================ TMA producer (both CTAs; bytes signalled on the leader)
    const uint64_t mX = reinterpret_cast<uint64_t>(&tmX), mW1 = reinterpret_cast<uint64_t>(&tmW1),
                   mW2 = reinterpret_cast<uint64_t>(&tmW2);
    const uint32_t b_xt = leader_cta_addr(smem_u32(xt_full));
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
    auto load_wa = [&](int pos) {
      const int j = tm_chunk(pos, rot, NC);
      const int n1 = (j == NC - 1) ? p.last_n1 : TM_CH;
      mbar_wait<128>(&wa_empty[sa], pa ^ 1);
      if (elect_one_sync()) {
        if (p.flags & 32) { if (is_leader) mbar_arrive(&wa_full[sa]); }
        else {
          if (is_leader) mbar_arrive_expect_tx(&wa_full[sa], 2 * p.wa_stage);
          tm_load_wa(s_wa + sa * p.wa_stage, mW1, leader_cta_addr(smem_u32(&wa_full[sa])),
                     j * TM_CH + cta_rank * (n1 >> 1), p);
        }
      }
      __syncwarp();
      if (++sa == (uint32_t)p.s_wa) { sa = 0; pa ^= 1; }
    };
    auto load_wb = [&](int pos) {
      const int j = tm_chunk(pos, rot, NC);
      mbar_wait<128>(&wb_empty[sb], pb ^ 1);
      if (elect_one_sync()) {
        if (p.flags & 32) { if (is_leader) mbar_arrive(&wb_full[sb]); }
        else {
          if (is_leader) mbar_arrive_expect_tx(&wb_full[sb], 2 * p.wb_stage);
          tma_load_3d_u32<2>(s_wb + sb * p.wb_stage, mW2, leader_cta_addr(smem_u32(&wb_full[sb])), j * TM_CH,
                             cta_rank * (p.NT >> 1), 0);
        }
      }
      __syncwarp();
      if (++sb == (uint32_t)p.s_wb) { sb = 0; pb ^= 1; }
    };
    // within an item W1(j) is requested two chunks ahead of W2(j - 2) (the order the issuer warps consume them in); the
    // last two W2 chunks of an item go out before the producer waits for the activation buffer of the next one
    int it = 0;
    for (int pair = cluster_id; pair < p.n_pairs; pair += num_clusters, ++it) {
      const TokTile t = tm_tile(p, pair, cta_rank);
      mbar_wait<128>(xt_empty, (it & 1) ^ 1);
      if (elect_one_sync()) {
        if (is_leader) mbar_arrive_expect_tx(xt_full, 2 * p.NT * 256);
        tma_load_3d_u32<2>(s_xt, mX, b_xt, t.c0, 0, t.b);
        tma_load_3d_u32<2>(s_xt + p.NT * 128, mX, b_xt, t.c0 + 64, 0, t.b);
      }
      __syncwarp();
      for (int j = 0; j < NC; ++j) {
        load_wa(j);
        if (j >= 2) load_wb(j - 2);
      }
      if (NC >= 2) load_wb(NC - 2);
      load_wb(NC - 1);
    }
  } else if (warp == 1) {
    // ================================================================ issuer of G1: Z^T = Xh^T * W1chunk^T (leader CTA)
    if (is_leader) {
      uint32_t sa = 0, pa = 0;
      int j1 = 0, it1 = 0;
      for (int t1 = 0; t1 < total; ++t1) {
        const int n1 = (tm_chunk(j1, rot, NC) == NC - 1) ? p.last_n1 : TM_CH;
        const int zb = t1 & 1;
        if (lane == 0) tm_stamp(p, 0, t1, 0);
        if (j1 == 0) mbar_wait<32>(xt_full, it1 & 1);
        mbar_wait2<32>(&z_empty[zb], ((t1 >> 1) & 1) ^ 1, &wa_full[sa], pa);
        if (lane == 0) tm_stamp(p, 0, t1, 1);
        tc_fence_after();
        if (elect_one_sync()) {
          tm_mma_over_tokens(tmem_base + zb * TM_CH, s_xt, s_wa + sa * p.wa_stage, umma_idesc_bf16(256, n1, 1, 0), p);
          umma_commit_2cta_mc(&wa_empty[sa]);
          umma_commit_2cta_mc(&z_full[zb]);
          if (j1 == NC - 1) umma_commit_2cta_mc(xt_empty);
        }
        __syncwarp();
        if (lane == 0) tm_stamp(p, 0, t1, 2);
        if (++sa == (uint32_t)p.s_wa) { sa = 0; pa ^= 1; }
        if (++j1 == NC) { j1 = 0; ++it1; }
      }
    }
  } else if (warp == 3) {
    // ================================================================ issuer of G2: U^T (+)= H^T * W2chunk^T (leader CTA)
    if (is_leader) {
      const uint32_t idesc_g2 = umma_idesc_bf16(256, p.NT, 0, 0);
      uint32_t sb = 0, pb = 0;
      int j2 = 0, it2 = 0;
      int hb = 0;                 // hidden-tile buffer of chunk t2 = t2 % nhb, hph = (t2 / nhb) & 1
      uint32_t hph = 0;
      for (int t2 = 0; t2 < total; ++t2) {
        if (lane == 0) tm_stamp(p, 0, t2, 4);
        if (j2 == 0) mbar_wait<32>(u_empty, (it2 & 1) ^ 1);
        mbar_wait2<32>(&h_full[hb], hph, &wb_full[sb], pb);
        if (lane == 0) tm_stamp(p, 0, t2, 5);
        tc_fence_after();
        if (elect_one_sync()) {
          const int ksteps = (tm_chunk(j2, rot, NC) == NC - 1) ? (p.last_n1 >> 4) : (TM_CH >> 4);
          tm_mma_over_hidden(tmem_base + 2 * TM_CH, s_h + hb * TM_HTILE, s_wb + sb * p.wb_stage, idesc_g2, ksteps, j2 == 0);
          umma_commit_2cta_mc(&wb_empty[sb]);
          umma_commit_2cta_mc(&h_free[hb]);
          if (j2 == NC - 1) umma_commit_2cta_mc(u_full);
        }
        __syncwarp();
        if (lane == 0) tm_stamp(p, 0, t2, 6);
        if (++sb == (uint32_t)p.s_wb) { sb = 0; pb ^= 1; }
        if (++j2 == NC) { j2 = 0; ++it2; }
        if (++hb == p.nhb) { hb = 0; hph ^= 1; }
      }
    }
  } else if (warp == 2) {
    // ================================================================ TMA stores (each CTA its own tiles): the saved hidden
    // tile of every chunk, and per item the residual-in / output-out tile: once the epilogue warps have turned the
    // residual into the output in place (ro_done) it is stored and the NEXT item's residual is loaded behind it.  (While
    // this warp waits for ro_done the epilogue is in its output phase and writes no hidden tile; the first hidden tiles of
    // the next item are stored a little late, which the double buffer absorbs.)
    auto load_residual = [&](int pair) {
      const TokTile t = tm_tile(p, pair, cta_rank);
      mbar_arrive_expect_tx(ro_full, p.NT * 256);
      tma_load_3d_u32<1>(s_ro, reinterpret_cast<uint64_t>(&tmR), smem_u32(ro_full), t.c0, 0, t.b);
    };
    if (cluster_id < p.n_pairs && elect_one_sync()) load_residual(cluster_id);
    __syncwarp();
    // "H(g) has been written by the 4 warps of group g & 1" is forwarded as ONE arrival per CTA to the leader's h_full; the
    // tile is stored (saved activation) and the buffer released once the store has READ it -- checked one chunk later
    // (wait_read<1>), so that this warp never sits in a TMA wait while the next hand-off is due.  With three hidden-tile
    // buffers the late release costs nothing.  At the end of an item the output tile.
    int pend_hb = -1;                                      // buffer whose store was committed last and is not yet released
    auto release_pending = [&](int keep) {                 // keep = stores that may still be reading (0 or 1)
      if (pend_hb >= 0 && elect_one_sync()) {
        if (keep) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
        mbar_arrive(&h_free[pend_hb]);
      }
      __syncwarp();
      pend_hb = -1;
    };
    int hbf = 0;
    auto forward_hw = [&](int g) {
      const int gi = g & 1;
      const int item = g / NC, j = g - item * NC;
      const TokTile t = tm_tile(p, cluster_id + item * num_clusters, cta_rank);
      named_bar_sync(TM_NB_HW + gi, 32 * (TM_FWD_EPI_WARPS / 2 + 1));
      if (lane == 0) tm_stamp(p, 2, g, 0);
      const bool store = save_hidden && t.valid && !(p.flags & 2);
      if (elect_one_sync()) {
        tm_arrive_leader(&h_full[hbf], is_leader);
        if (store) {
          tma_store_3d(&tmH, smem + (s_h - s_base) + hbf * TM_HTILE, tm_chunk(j, rot, NC) * TM_CH, t.c0, t.b);
          tma_store_commit();
        }
      }
      __syncwarp();
      // The release of buffer g - 1 comes AFTER this warp has passed the named barrier of chunk g, also when nothing is
      // stored: a group can therefore never signal "written" for chunk g + 2 before its signal for chunk g has been
      // consumed (a named barrier must not collect two generations of arrivals).
      release_pending(store ? 1 : 0);                      // the PREVIOUS store has long finished reading
      pend_hb = hbf;
      if (++hbf == p.nhb) hbf = 0;
      if (lane == 0) tm_stamp(p, 2, g, 1);
      if (j == NC - 1) {                                   // last chunk of an item: its output tile follows (the hidden
        mbar_wait<64>(ro_done, item & 1);                  // buffer of this chunk stays pending like any other one)
        if (elect_one_sync()) {
          if (t.valid) {
            tma_store_3d(&tmU, smem + (s_ro - s_base), t.c0, 0, t.b);     // rows >= N and channels >= C are clipped
            tma_store_commit();
            tma_store_wait_read<0>();
          }
          if (item + 1 < my_items) load_residual(cluster_id + (item + 1) * num_clusters);
        }
        __syncwarp();
      }
    };
    for (int g = 0; g < total; ++g) forward_hw(g);
    release_pending(0);
    if (elect_one_sync()) tma_store_wait_all<0>();
    __syncwarp();
  } else if (warp >= TM_FWD_EPI0) {
    // ================================================================ epilogue warps (software-pipelined over chunks)
    // Two groups of 4 warps work on ALTERNATE chunks (group = chunk parity = Z buffer = hidden-tile buffer): while one group
    // runs the GELU of chunk g, the other one is in the hand-off phases of chunk g + 1 (barrier, TMEM load, tile write,
    // fence).  A warp owns TMEM lane quarter warp % 4 and ALL 64 columns of its chunk: 8 fat warps instead of 16 thin ones,
    // because the GELU is a 14-deep dependent chain per element pair and at 96 registers per thread (640 threads) ptxas
    // serialised the pairs -- ncu: "wait" (fixed-latency dependency) the top stall, 2.4 warps per issue, the FMA pipe 41 %
    // busy; with 384 threads it has 168 registers and overlaps the chains of many pairs.
    const int q = warp & 3;                               // TMEM lane quarter
    const int gi = (warp - TM_FWD_EPI0) >> 2;             // group: chunks with g % 2 == gi
    const int cq = gi;                                    // token-group phase of this warp in the output epilogue
    const int row = q * 32 + lane;                        // channel within the tile
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int ngrp = p.NT >> 4;
    const bool tr = (warp - TM_FWD_EPI0) % 4 == 0 && lane == 0;
    const bool poller = ((warp - TM_FWD_EPI0) & 3) == 0;   // the one warp of the group that polls the mbarriers
    int it = 0;                                           // item whose output epilogue this warp does next
    int hb = gi % p.nhb;                                  // hidden-tile buffer of chunk g = g % nhb, phase (g / nhb) & 1
    uint32_t hph = (gi / p.nhb) & 1;
    for (int g = gi; ; g += 2) {
      // ---- output epilogues of every item that ends before chunk g (or all remaining ones once g runs out)
      const int item_of_g = g < total ? g / NC : my_items;
      for (; it < item_of_g; ++it) {
        // ---- output: U[b, n, ch] = U^T[ch, n] + b2[n] + x[b, n, ch].  This thread owns one channel (TMEM lane) and gets 16
        // tokens per tcgen05.ld; the [token][channel] transposition goes through the shared residual/output tile: 2-byte
        // accesses at [n * 256 + ch * 2] -- the 32 lanes of a warp touch 64 consecutive bytes, conflict-free -- updated in
        // place, then ONE TMA store per tile (issued by warp 3) writes it out and clips rows >= N / channels >= C.
        mbar_wait2<0>(ro_full, it & 1, u_full, it & 1);
        tc_fence_after();
        const uint32_t ro_col = s_ro + row * 2;
  #pragma unroll 1
        for (int grp = cq; grp < ngrp + 2; grp += 2) {
          const bool has = grp < ngrp;
          const bool last = grp + 2 >= ngrp;                // this warp's last visit (possibly an empty one)
          uint32_t v[16];
          if (has) {
            tmem_ld_32x32b_x16(tmem_base + 2 * TM_CH + grp * 16 + lane_addr, v);
            tmem_ld_wait();                                 // (also completes a pending prefetch of the next item's chunk 0)
          }
          if (last) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tm_arrive_leader(u_empty, is_leader);
          }
          if (has) {
            const uint32_t a0 = ro_col + grp * 16 * 256;
  #pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 bv = lds_f4(s_b2 + (grp * 16 + 4 * i4) * 4);
              const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int i = 4 * i4 + e;
                unsigned short r;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(r) : "r"(a0 + i * 256) : "memory");
                const float f = __uint_as_float(v[i]) + bb[e] + __uint_as_float(static_cast<uint32_t>(r) << 16);
                asm volatile("st.shared.u16 [%0], %1;" ::"r"(a0 + i * 256), "h"(static_cast<unsigned short>(bf16_bits(f))) : "memory");
              }
            }
          }
          if (last) break;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(ro_done);
      }
      if (g >= total) break;
      // ---- chunk g
      {
        const int pos = g - item_of_g * NC;
        const int j = tm_chunk(pos, rot, NC);
        const int n1 = (j == NC - 1) ? p.last_n1 : TM_CH;
        if (tr) tm_stamp(p, 1, g, 0);
        // ONE warp of the group polls the mbarriers (Z(g) complete; hidden buffer free: G2 and the TMA store of its previous
        // user done), the named barrier releases the other seven: an mbarrier operation costs a warp 100-200 cycles
        if (poller) mbar_wait2<0>(&z_full[gi], (g >> 1) & 1, &h_free[hb], hph ^ 1);
        named_bar_sync(TM_NB_START + gi, 32 * (TM_FWD_EPI_WARPS / 2));
        tc_fence_after();
        if (tr) tm_stamp(p, 1, g, 1);
        // all 64 accumulator columns of the chunk first (the Z buffer goes back to the G1 issuer at once), then the math
        uint32_t v[4][16];
        if (!(p.flags & 16)) {
          if (n1 > 16) tmem_ld_x16_pair_wait(tmem_base + gi * TM_CH + lane_addr, tmem_base + gi * TM_CH + 16 + lane_addr, v[0], v[1]);
          else tmem_ld_x16_wait(tmem_base + gi * TM_CH + lane_addr, v[0]);
          if (n1 > 32) tmem_ld_x16_pair_wait(tmem_base + gi * TM_CH + 32 + lane_addr, tmem_base + gi * TM_CH + 48 + lane_addr, v[2], v[3]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) tm_arrive_leader(&z_empty[gi], is_leader);           // "Z(g) consumed": the G1 issuer may overwrite it
        if (tr) tm_stamp(p, 1, g, 2);
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
          if (16 * hh >= n1) continue;
          uint32_t o[8];
          if (p.flags & 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e) o[e] = pack_bf16x2(__uint_as_float(v[hh][2 * e]), __uint_as_float(v[hh][2 * e + 1]));
          } else {
            f32x2 zz[8], gl[8];                                     // 16 columns = 8 pairs in flight (gelu_rcp16_xn)
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 bv = lds_f4(s_b1 + (j * TM_CH + 16 * hh + 4 * e4) * 4);
              zz[2 * e4] = pack2(__uint_as_float(v[hh][4 * e4]) + bv.x, __uint_as_float(v[hh][4 * e4 + 1]) + bv.y);
              zz[2 * e4 + 1] = pack2(__uint_as_float(v[hh][4 * e4 + 2]) + bv.z, __uint_as_float(v[hh][4 * e4 + 3]) + bv.w);
            }
            gelu_rcp16_xn<8>(zz, gl);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) o[qq] = pack_bf16x2_f2(gl[qq]);
          }
          if (!(p.flags & 8)) tm_store_hidden_row(s_h + hb * TM_HTILE, row, hh, o);
        }
        if (tr) tm_stamp(p, 1, g, 3);
        fence_proxy_async_smem();
        named_bar_arrive(TM_NB_HW + gi, 32 * (TM_FWD_EPI_WARPS / 2 + 1));   // warp 2 forwards "H(g) written", stores the tile
        if (tr) tm_stamp(p, 1, g, 4);
        hb += 2; if (hb >= p.nhb) { hb -= p.nhb; hph ^= 1; }                // next chunk of this group is g + 2
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Backward (data-gradient chain).  SMEM: [barriers][Xh^T tile][dU^T tile][W1 ring][W2^T ring][W1^T ring][dZ tiles 2 x 16 KB][b1]
// TMEM: Z double buffer [0, 128), dH double buffer [128, 256), dXh accumulator [256, 256 + NT).
// Per hidden chunk: G1 Z^T = Xh^T * W1chunk^T (recomputed, never stored), G2 dH^T = dU^T * W2[:, chunk],
// epilogue dZ^T = dH^T .* gelu'(Z^T + b1) -> bf16 SMEM tile (+ TMA store into dZ^T [B, C, Ds] for the weight gradient),
// G3 dXh^T += dZ^T * W1[chunk, :].  The weight operands are all K-major copies prepared once per step by
// vmlp_tokmix_prepare: W1 [Ds, Np], W2^T [Ds, Np], W1^T [N, Ds].
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TM_BWD_THREADS, 1)
tokmix_bwd_sm100(const __grid_constant__ CUtensorMap tmX,      // Xh    [B, N, C]   box (64 c, NT rows)
                 const __grid_constant__ CUtensorMap tmDU,     // dU    [B, N, C]   box (64 c, NT rows)
                 const __grid_constant__ CUtensorMap tmW1,     // W1    [Ds, NT]    box (64 k, 32 rows)
                 const __grid_constant__ CUtensorMap tmW2T,    // W2^T  [Ds, NT]    box (64 k, 32 rows)
                 const __grid_constant__ CUtensorMap tmW1T,    // W1^T  [N, Ds]     box (64 k, NT/2 rows)
                 const __grid_constant__ CUtensorMap tmDZ,     // dZ^T  [B, C, Ds]  box (64 m, 128 c)
                 const TokParams p) {
  const int cta_rank = static_cast<int>(cluster_ctarank());
  const bool is_leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* in_full = bars + 0;     uint64_t* in_empty = bars + 1;
  uint64_t* dx_full = bars + 2;     uint64_t* dx_empty = bars + 3;
  uint64_t* zd_full = bars + 4;     uint64_t* zd_empty = bars + 6;     // [2] each
  uint64_t* dz_full = bars + 8;     uint64_t* dz_empty = bars + 10;
  uint64_t* dz_done = bars + 12;    uint64_t* dzs_empty = bars + 14;
  uint64_t* w1_full = bars + 16;    uint64_t* w1_empty = bars + 20;    // up to 4 stages each
  uint64_t* w2_full = bars + 24;    uint64_t* w2_empty = bars + 28;
  uint64_t* w3_full = bars + 32;    uint64_t* w3_empty = bars + 36;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 48);
  const uint32_t s_base = smem_u32(smem);
  const uint32_t s_xt = s_base + TM_BAR_BYTES;
  const uint32_t s_dut = s_xt + p.NT * 256;
  const uint32_t s_w1 = s_dut + p.NT * 256;
  const uint32_t s_w2 = s_w1 + p.s_wa * p.wa_stage;
  const uint32_t s_w3 = s_w2 + p.s_wa * p.wa_stage;
  const uint32_t s_dz = s_w3 + p.s_wb * p.wb_stage;
  float* sb1 = reinterpret_cast<float*>(smem + (s_dz - s_base) + p.nhb * TM_HTILE);
  float* sdb = sb1 + p.n_chunks * TM_CH;          // partial sums of d b1, one array per helper warp, flushed once at the end
  const uint32_t s_db = s_dz + p.nhb * TM_HTILE + p.n_chunks * TM_CH * 4;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmDU); tma_prefetch_desc(&tmW1);
    tma_prefetch_desc(&tmW2T); tma_prefetch_desc(&tmW1T); tma_prefetch_desc(&tmDZ);
    mbar_init(in_full, 1); mbar_init(in_empty, 1);
    mbar_init(dx_full, 1); mbar_init(dx_empty, 2 * TM_BWD_EPI_WARPS);
    const int per_chunk = p.nhb == 2 ? TM_BWD_EPI_WARPS / 2 : TM_BWD_EPI_WARPS;   // epilogue warps that work on one chunk
    for (int i = 0; i < 2; ++i) {
      mbar_init(&zd_full[i], 1);  mbar_init(&zd_empty[i], 2 * per_chunk);
      mbar_init(&dz_full[i], 2 * per_chunk);  mbar_init(&dz_empty[i], 1);
      mbar_init(&dz_done[i], TM_BWD_EPI_WARPS);      mbar_init(&dzs_empty[i], 2);   // store warp + column-sum warp
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&w1_full[i], 1); mbar_init(&w1_empty[i], 1);
      mbar_init(&w2_full[i], 1); mbar_init(&w2_empty[i], 1);
      mbar_init(&w3_full[i], 1); mbar_init(&w3_empty[i], 1);
    }
    fence_barrier_init();
  }
  for (int i = threadIdx.x; i < p.n_chunks * TM_CH; i += TM_BWD_THREADS) {
    sb1[i] = i < p.Ds ? __bfloat162float(p.b1[i]) : 0.f;
    sdb[i] = 0.f;
    sdb[p.n_chunks * TM_CH + i] = 0.f;
  }
  if (warp == 1) { tmem_alloc_2cta(tmem_slot, 512); tmem_relinquish_2cta(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int NC = p.n_chunks;
  const int rot = (p.flags & 64) ? 0 : cluster_id % NC;

  if (warp == 0) {
    // ================================================================ TMA producer
    const uint64_t mX = reinterpret_cast<uint64_t>(&tmX), mDU = reinterpret_cast<uint64_t>(&tmDU),
                   mW1 = reinterpret_cast<uint64_t>(&tmW1), mW2T = reinterpret_cast<uint64_t>(&tmW2T),
                   mW1T = reinterpret_cast<uint64_t>(&tmW1T);
    const uint32_t b_in = leader_cta_addr(smem_u32(in_full));
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
    auto load_w12 = [&](int pos) {
      const int j = tm_chunk(pos, rot, NC);
      const int n1 = (j == NC - 1) ? p.last_n1 : TM_CH;
      const int row = j * TM_CH + cta_rank * (n1 >> 1);
      mbar_wait<128>(&w1_empty[sa], pa ^ 1);
      if (elect_one_sync()) {
        if (p.flags & 32) { if (is_leader) mbar_arrive(&w1_full[sa]); }
        else {
          if (is_leader) mbar_arrive_expect_tx(&w1_full[sa], 2 * p.wa_stage);
          tm_load_wa(s_w1 + sa * p.wa_stage, mW1, leader_cta_addr(smem_u32(&w1_full[sa])), row, p);
        }
      }
      __syncwarp();
      mbar_wait<128>(&w2_empty[sa], pa ^ 1);
      if (elect_one_sync()) {
        if (p.flags & 32) { if (is_leader) mbar_arrive(&w2_full[sa]); }
        else {
          if (is_leader) mbar_arrive_expect_tx(&w2_full[sa], 2 * p.wa_stage);
          tm_load_wa(s_w2 + sa * p.wa_stage, mW2T, leader_cta_addr(smem_u32(&w2_full[sa])), row, p);
        }
      }
      __syncwarp();
      if (++sa == (uint32_t)p.s_wa) { sa = 0; pa ^= 1; }
    };
    auto load_w3 = [&](int pos) {
      const int j = tm_chunk(pos, rot, NC);
      mbar_wait<128>(&w3_empty[sb], pb ^ 1);
      if (elect_one_sync()) {
        if (p.flags & 32) { if (is_leader) mbar_arrive(&w3_full[sb]); }
        else {
          if (is_leader) mbar_arrive_expect_tx(&w3_full[sb], 2 * p.wb_stage);
          tma_load_3d_u32<2>(s_w3 + sb * p.wb_stage, mW1T, leader_cta_addr(smem_u32(&w3_full[sb])), j * TM_CH,
                             cta_rank * (p.NT >> 1), 0);
        }
      }
      __syncwarp();
      if (++sb == (uint32_t)p.s_wb) { sb = 0; pb ^= 1; }
    };
    int it = 0;
    for (int pair = cluster_id; pair < p.n_pairs; pair += num_clusters, ++it) {
      const TokTile t = tm_tile(p, pair, cta_rank);
      mbar_wait<128>(in_empty, (it & 1) ^ 1);
      if (elect_one_sync()) {
        if (is_leader) mbar_arrive_expect_tx(in_full, 4 * p.NT * 256);
        tma_load_3d_u32<2>(s_xt, mX, b_in, t.c0, 0, t.b);
        tma_load_3d_u32<2>(s_xt + p.NT * 128, mX, b_in, t.c0 + 64, 0, t.b);
        tma_load_3d_u32<2>(s_dut, mDU, b_in, t.c0, 0, t.b);
        tma_load_3d_u32<2>(s_dut + p.NT * 128, mDU, b_in, t.c0 + 64, 0, t.b);
      }
      __syncwarp();
      for (int j = 0; j < NC; ++j) {
        load_w12(j);
        if (j >= 2) load_w3(j - 2);
      }
      if (NC >= 2) load_w3(NC - 2);
      load_w3(NC - 1);
    }
  } else if (warp == 1) {
    // ================================================================ MMA issuer (leader CTA)
    if (is_leader) {
      const uint32_t idesc_g3 = umma_idesc_bf16(256, p.NT, 0, 0);
      int my_items = 0;
      for (int pair = cluster_id; pair < p.n_pairs; pair += num_clusters) ++my_items;
      const int total = my_items * NC;
      uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
      int j1 = 0, it1 = 0, j3 = 0, it3 = 0;
      int t1 = 0, t3 = 0;
      auto do_g12 = [&]() {       // ---- Z^T = Xh^T * W1chunk^T and dH^T = dU^T * W2[:, chunk]
        const int n1 = (tm_chunk(j1, rot, NC) == NC - 1) ? p.last_n1 : TM_CH;
        if (lane == 0) tm_stamp(p, 0, t1, 0);
        if (j1 == 0) mbar_wait<32>(in_full, it1 & 1);
        const int zb = t1 & 1;
        mbar_wait<32>(&zd_empty[zb], ((t1 >> 1) & 1) ^ 1);
        if (lane == 0) tm_stamp(p, 0, t1, 1);
        mbar_wait<32>(&w1_full[sa], pa);
        tc_fence_after();
        if (elect_one_sync()) {
          tm_mma_over_tokens(tmem_base + zb * TM_CH, s_xt, s_w1 + sa * p.wa_stage, umma_idesc_bf16(256, n1, 1, 0), p);
          umma_commit_2cta_mc(&w1_empty[sa]);
        }
        __syncwarp();
        mbar_wait<32>(&w2_full[sa], pa);
        tc_fence_after();
        if (elect_one_sync()) {
          tm_mma_over_tokens(tmem_base + 2 * TM_CH + zb * TM_CH, s_dut, s_w2 + sa * p.wa_stage,
                             umma_idesc_bf16(256, n1, 1, 0), p);
          umma_commit_2cta_mc(&w2_empty[sa]);
          umma_commit_2cta_mc(&zd_full[zb]);
          if (j1 == NC - 1) umma_commit_2cta_mc(in_empty);
        }
        __syncwarp();
        if (lane == 0) tm_stamp(p, 0, t1, 2);
        if (++sa == (uint32_t)p.s_wa) { sa = 0; pa ^= 1; }
        if (++j1 == NC) { j1 = 0; ++it1; }
        ++t1;
      };
      auto do_g3 = [&]() {        // ---- dXh^T (+)= dZ^T * W1[chunk, :]
        const int hb = p.nhb == 2 ? (t3 & 1) : 0;
        if (lane == 0) tm_stamp(p, 0, t3, 4);
        mbar_wait<32>(&dz_full[hb], (p.nhb == 2 ? (t3 >> 1) : t3) & 1);
        if (lane == 0) tm_stamp(p, 0, t3, 5);
        mbar_wait<32>(&w3_full[sb], pb);
        if (j3 == 0) mbar_wait<32>(dx_empty, (it3 & 1) ^ 1);
        tc_fence_after();
        if (elect_one_sync()) {
          const int ksteps = (tm_chunk(j3, rot, NC) == NC - 1) ? (p.last_n1 >> 4) : (TM_CH >> 4);
          tm_mma_over_hidden(tmem_base + 4 * TM_CH, s_dz + hb * TM_HTILE, s_w3 + sb * p.wb_stage, idesc_g3, ksteps, j3 == 0);
          umma_commit_2cta_mc(&w3_empty[sb]);
          umma_commit_2cta_mc(&dz_empty[hb]);
          if (j3 == NC - 1) umma_commit_2cta_mc(dx_full);
        }
        __syncwarp();
        if (lane == 0) tm_stamp(p, 0, t3, 6);
        if (++sb == (uint32_t)p.s_wb) { sb = 0; pb ^= 1; }
        if (++j3 == NC) { j3 = 0; ++it3; }
        ++t3;
      };
      for (int t = 0; t < total + 2; ++t) {
        const bool g1 = t < total, g3 = t >= 2;
        // Fixed order: G1 / G2 of chunk t, then G3 of chunk t - 2 (first only at an item boundary, where the next item's
        // activation tiles may still be in flight, and when nothing else is left).  Experiment behind VMLP_TM_FLAGS=128:
        // G3 first whenever its dZ tile is already complete (the timeline shows G3 issued ~860 cycles after its tile was
        // there, the issuer being busy with the 26 MMAs of chunk t, and the epilogue waiting ~330 cycles per chunk for it)
        // -- measured SLOWER (290 vs 283 us): what G3 gains, G1 / G2 of the chunk after next lose.
        bool g3_first = g3 && (!g1 || j1 < 2);
        if (g3 && !g3_first && (p.flags & 128)) {
          const int hb3 = p.nhb == 2 ? (t3 & 1) : 0;
          g3_first = mbar_test(&dz_full[hb3], (p.nhb == 2 ? (t3 >> 1) : t3) & 1) != 0;
          g3_first = __shfl_sync(0xffffffffu, g3_first, 0);           // warp-uniform decision
        }
        if (g3 && g3_first) do_g3();
        if (g1) do_g12();
        if (g3 && !g3_first) do_g3();
      }
    }
  } else if (warp == 2 || warp == 3) {
    // ================================================================ helper warps: warp 2 stores the dZ^T tile by TMA;
    // both sum the tile's columns over their 64 channel rows (d b1[m] = sum over (b, c) of dZ): lane l owns hidden
    // columns 2l, 2l+1 of the chunk and reads one 32-bit word per row (the 16-byte chunk index is un-swizzled per row).
    // warp 2 owns channel rows 0..63 of the tile, warp 3 rows 64..127; each has its own partial-sum array
    const int r0 = (warp - 2) * 64;
    const uint32_t kc = lane & 7;                      // 16-byte chunk = eight hidden columns
    const int rs = lane >> 3;                          // row phase: rows r0 + rs + 4 i
    const int nb_threads = 32 * ((p.nhb == 2 ? TM_BWD_EPI_WARPS / 2 : TM_BWD_EPI_WARPS) + 2);
    int g = 0;
    for (int pair = cluster_id; pair < p.n_pairs; pair += num_clusters) {
      const TokTile t = tm_tile(p, pair, cta_rank);
      for (int pos = 0; pos < NC; ++pos, ++g) {
        const int j = tm_chunk(pos, rot, NC);
        const int hb = p.nhb == 2 ? (g & 1) : 0;
        // "tile written" comes over a hardware named barrier (the epilogue warps of the chunk arrive, the 2 helper warps
        // sync): an mbarrier poll with back-off woke the helpers ~1100 cycles late (clock64 timeline), without back-off it
        // steals issue slots from the math warps
        named_bar_sync(TM_NB_DZ + hb, nb_threads);
        if (warp == 2 && lane == 0) tm_stamp(p, 2, g, 0);
        if (warp == 2 && elect_one_sync()) {
          if (t.valid && !(p.flags & 2)) {
            tma_store_3d(&tmDZ, smem + (s_dz - s_base) + hb * TM_HTILE, j * TM_CH, t.c0, t.b);
            tma_store_commit();
          }
        }
        __syncwarp();
        if (t.valid && !(p.flags & 4)) {
          // one LDS.128 per row, eight independent fp32 accumulators, two shuffle steps fold the four row phases
          const uint32_t tb = s_dz + hb * TM_HTILE;
          float acc[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 8
          for (int i = 0; i < 16; ++i) {
            const int r = r0 + rs + 4 * i;
            const uint4 w = ld_shared_v4(tb + r * 128 + ((kc ^ (r & 7)) << 4));
            acc[0] += bf16lo(w.x); acc[1] += bf16hi(w.x); acc[2] += bf16lo(w.y); acc[3] += bf16hi(w.y);
            acc[4] += bf16lo(w.z); acc[5] += bf16hi(w.z); acc[6] += bf16lo(w.w); acc[7] += bf16hi(w.w);
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
            acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
          }
          if (rs == 0) {           // columns beyond Ds hold stale or zero data and are never flushed
            // each (helper warp, lane) owns 8 columns of the partial-sum array: a plain read-modify-write.  (Shared memory
            // has no native fp32 add: atomicAdd / red.shared compile to an ATOMS.CAS spin loop -- the clock64 timeline showed
            // 3100 cycles per tile here, with the epilogue's next write waiting behind it.)
            const uint32_t a = s_db + (((warp - 2) * p.n_chunks + j) * TM_CH + kc * 8) * 4;
            float4 s0 = lds_f4(a), s1 = lds_f4(a + 16);
            s0.x += acc[0]; s0.y += acc[1]; s0.z += acc[2]; s0.w += acc[3];
            s1.x += acc[4]; s1.y += acc[5]; s1.z += acc[6]; s1.w += acc[7];
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(s0.x), "f"(s0.y), "f"(s0.z), "f"(s0.w) : "memory");
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a + 16), "f"(s1.x), "f"(s1.y), "f"(s1.z), "f"(s1.w) : "memory");
          }
        }
        __syncwarp();
        if (warp == 2 && lane == 0) tm_stamp(p, 2, g, 1);
        if (elect_one_sync()) {
          if (warp == 2) tma_store_wait_read<0>();
          mbar_arrive(&dzs_empty[hb]);
        }
        __syncwarp();
        if (warp == 2 && lane == 0) tm_stamp(p, 2, g, 2);
      }
    }
    if (warp == 2 && elect_one_sync()) tma_store_wait_all<0>();
    __syncwarp();
  } else if (warp >= TM_BWD_EPI0) {
    // ================================================================ epilogue warps: 8 warps, each owns TMEM lane quarter
    // warp % 4 and 32 of the chunk's 64 columns (two 16-column slices).  Eight fat warps instead of sixteen thin ones: gelu'
    // is a ~20-deep dependent chain per element pair, and with 640 threads (96 registers) ptxas serialised the pairs; with
    // 384 threads (168 registers) it overlaps them, and every chunk costs half as many mbarrier operations.
    if (p.nhb == 2) {
      // ---------------------------------------------------------------- ping-pong form (two dZ buffers): two groups of 4
      // warps on ALTERNATE chunks (group = chunk parity = Z / dH buffer = dZ buffer), a warp does all 64 columns of its chunk
      // in two halves of 32.  While one group computes, the other one is in its load / write / fence / barrier phases, and
      // neither G3 nor the helper warps of a chunk are in the way of the NEXT chunk's write.
      const int q = warp & 3;
      const int gi = (warp - TM_BWD_EPI0) >> 2;
      const int row = q * 32 + lane;
      const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
      const int ngrp = p.NT >> 4;
      const bool tr = (warp - TM_BWD_EPI0) % 4 == 0 && lane == 0;
      int my_items = 0;
      for (int pair = cluster_id; pair < p.n_pairs; pair += num_clusters) ++my_items;
      const int total = my_items * NC;
      int it = 0;
      for (int g = gi; ; g += 2) {
        const int item_of_g = g < total ? g / NC : my_items;
        for (; it < item_of_g; ++it) {
          // ---- output of item `it`: dXh[b, n, ch] = dXh^T[ch, n]; the two warps of a lane quarter take alternate token groups
          const TokTile t = tm_tile(p, cluster_id + it * num_clusters, cta_rank);
          const int ch = t.c0 + row;
          const bool ch_ok = t.valid && ch < p.C;
          __nv_bfloat16* obase = p.out + ((long long)t.b * p.N) * p.C + ch;
          mbar_wait(dx_full, it & 1);
          tc_fence_after();
#pragma unroll 1
          for (int grp = gi; grp < ngrp + 2; grp += 2) {
            const bool has = grp < ngrp;
            const bool last = grp + 2 >= ngrp;
            uint32_t v[16];
            if (has) tmem_ld_x16_wait(tmem_base + 4 * TM_CH + grp * 16 + lane_addr, v);
            if (last) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) tm_arrive_leader(dx_empty, is_leader);
            }
            if (has && ch_ok) {
              __nv_bfloat16* po = obase + (long long)(grp * 16) * p.C;
#pragma unroll
              for (int i = 0; i < 16; ++i, po += p.C)
                if (grp * 16 + i < p.N) stg_u16(po, bf16_bits(__uint_as_float(v[i])));
            }
            if (last) break;
          }
        }
        if (g >= total) break;
        const int j = tm_chunk(g - item_of_g * NC, rot, NC);
        const int n1 = (j == NC - 1) ? p.last_n1 : TM_CH;
        const TokTile t = tm_tile(p, cluster_id + item_of_g * num_clusters, cta_rank);
        (void)t;
        if (tr) tm_stamp(p, 1, g, 0);
        mbar_wait(&zd_full[gi], (g >> 1) & 1);
        tc_fence_after();
        if (tr) tm_stamp(p, 1, g, 1);
        uint32_t o[4][8];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const bool live0 = half * 32 < n1, live1 = half * 32 + 16 < n1;
          uint32_t vz[2][16], vh[2][16];
          if (!(p.flags & 16)) {
            const uint32_t az = tmem_base + gi * TM_CH + half * 32 + lane_addr, ah = az + 2 * TM_CH;
            if (live1) { tmem_ld_x16_pair_wait(az, az + 16, vz[0], vz[1]); tmem_ld_x16_pair_wait(ah, ah + 16, vh[0], vh[1]); }
            else if (live0) tmem_ld_x16_pair_wait(az, ah, vz[0], vh[0]);
          }
          if (half == 1) {                                           // Z / dH of this chunk are in registers: G1 / G2 may refill
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tm_arrive_leader(&zd_empty[gi], is_leader);
            if (tr) tm_stamp(p, 1, g, 2);
          }
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            if (!(hh ? live1 : live0)) continue;
            uint32_t (&oo)[8] = o[2 * half + hh];
            if (p.flags & 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                oo[e] = pack_bf16x2(__uint_as_float(vz[hh][2 * e]) + __uint_as_float(vh[hh][2 * e]),
                                    __uint_as_float(vz[hh][2 * e + 1]) + __uint_as_float(vh[hh][2 * e + 1]));
            } else {
              const float4* bp = reinterpret_cast<const float4*>(sb1 + j * TM_CH + half * 32 + 16 * hh);
#pragma unroll
              for (int e4 = 0; e4 < 4; ++e4) {
                const float4 bv = bp[e4];
                f32x2 gl, dg;
                gelu_erf_pair<true>(pack2(__uint_as_float(vz[hh][4 * e4]) + bv.x, __uint_as_float(vz[hh][4 * e4 + 1]) + bv.y), gl, dg);
                oo[2 * e4] = pack_bf16x2_f2(mul2(dg, pack2(__uint_as_float(vh[hh][4 * e4]), __uint_as_float(vh[hh][4 * e4 + 1]))));
                gelu_erf_pair<true>(pack2(__uint_as_float(vz[hh][4 * e4 + 2]) + bv.z, __uint_as_float(vz[hh][4 * e4 + 3]) + bv.w), gl, dg);
                oo[2 * e4 + 1] = pack_bf16x2_f2(mul2(dg, pack2(__uint_as_float(vh[hh][4 * e4 + 2]), __uint_as_float(vh[hh][4 * e4 + 3]))));
              }
            }
          }
        }
        const uint32_t hph = ((g >> 1) & 1) ^ 1;
        if (tr) tm_stamp(p, 1, g, 3);
        mbar_wait(&dz_empty[gi], hph);                               // G3 of chunk g - 2 has read this buffer
        mbar_wait(&dzs_empty[gi], hph);                              // ... and so have its TMA store and column sums
        if (tr) tm_stamp(p, 1, g, 4);
        if (!(p.flags & 8)) {
#pragma unroll
          for (int sl = 0; sl < 4; ++sl)
            if (sl * 16 < n1) tm_store_hidden_row(s_dz + gi * TM_HTILE, row, sl, o[sl]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tm_arrive_leader(&dz_full[gi], is_leader);
        named_bar_arrive(TM_NB_DZ + gi, 32 * (TM_BWD_EPI_WARPS / 2 + 2));
        if (tr) tm_stamp(p, 1, g, 5);
      }
    } else {
    const int q = warp & 3;
    const int cq = (warp - TM_BWD_EPI0) >> 2;             // columns [32 cq, 32 cq + 32) of the chunk
    const int row = q * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
    const int ngrp = p.NT >> 4;
    int g = 0, it = 0;
    for (int pair = cluster_id; pair < p.n_pairs; pair += num_clusters, ++it) {
      const TokTile t = tm_tile(p, pair, cta_rank);
      const int ch = t.c0 + row;
      const bool ch_ok = t.valid && ch < p.C;
      for (int pos = 0; pos < NC; ++pos, ++g) {
        const int j = tm_chunk(pos, rot, NC);
        const int zb = g & 1;
        const int n1 = (j == NC - 1) ? p.last_n1 : TM_CH;
        const bool live0 = cq * 32 < n1, live1 = cq * 32 + 16 < n1;
        const bool tr = warp == TM_BWD_EPI0 && lane == 0;
        if (tr) tm_stamp(p, 1, g, 0);
        mbar_wait(&zd_full[zb], (g >> 1) & 1);
        tc_fence_after();
        if (tr) tm_stamp(p, 1, g, 1);
        uint32_t vz[2][16], vh[2][16];
        if (!(p.flags & 16)) {
          const uint32_t az = tmem_base + zb * TM_CH + cq * 32 + lane_addr, ah = az + 2 * TM_CH;
          if (live1) { tmem_ld_x16_pair_wait(az, az + 16, vz[0], vz[1]); tmem_ld_x16_pair_wait(ah, ah + 16, vh[0], vh[1]); }
          else if (live0) tmem_ld_x16_pair_wait(az, ah, vz[0], vh[0]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) tm_arrive_leader(&zd_empty[zb], is_leader);
        if (tr) tm_stamp(p, 1, g, 2);
        uint32_t o[2][8];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          if (!(hh ? live1 : live0)) continue;
          if (p.flags & 1) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
              o[hh][e] = pack_bf16x2(__uint_as_float(vz[hh][2 * e]) + __uint_as_float(vh[hh][2 * e]),
                                     __uint_as_float(vz[hh][2 * e + 1]) + __uint_as_float(vh[hh][2 * e + 1]));
          } else {
            const float4* bp = reinterpret_cast<const float4*>(sb1 + j * TM_CH + cq * 32 + 16 * hh);
#pragma unroll
            for (int e4 = 0; e4 < 4; ++e4) {
              const float4 bv = bp[e4];
              f32x2 gl, dg;
              gelu_erf_pair<true>(pack2(__uint_as_float(vz[hh][4 * e4]) + bv.x, __uint_as_float(vz[hh][4 * e4 + 1]) + bv.y), gl, dg);
              o[hh][2 * e4] = pack_bf16x2_f2(mul2(dg, pack2(__uint_as_float(vh[hh][4 * e4]), __uint_as_float(vh[hh][4 * e4 + 1]))));
              gelu_erf_pair<true>(pack2(__uint_as_float(vz[hh][4 * e4 + 2]) + bv.z, __uint_as_float(vz[hh][4 * e4 + 3]) + bv.w), gl, dg);
              o[hh][2 * e4 + 1] = pack_bf16x2_f2(mul2(dg, pack2(__uint_as_float(vh[hh][4 * e4 + 2]), __uint_as_float(vh[hh][4 * e4 + 3]))));
            }
          }
        }
        const int hb = p.nhb == 2 ? zb : 0;                        // dZ tile buffer (single-buffered when two do not fit)
        const uint32_t hph = ((p.nhb == 2 ? (g >> 1) : g) & 1) ^ 1;
        if (tr) tm_stamp(p, 1, g, 3);
        mbar_wait(&dz_empty[hb], hph);                             // G3 of the previous user of this buffer has read it
        mbar_wait(&dzs_empty[hb], hph);                            // ... and so have its TMA store and column sums
        if (tr) tm_stamp(p, 1, g, 4);
        if (!(p.flags & 8)) {
          if (live0) tm_store_hidden_row(s_dz + hb * TM_HTILE, row, 2 * cq, o[0]);
          if (live1) tm_store_hidden_row(s_dz + hb * TM_HTILE, row, 2 * cq + 1, o[1]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) tm_arrive_leader(&dz_full[hb], is_leader);
        named_bar_arrive(TM_NB_DZ + hb, 32 * (TM_BWD_EPI_WARPS + 2));
        if (tr) tm_stamp(p, 1, g, 5);
      }
      // ---- output: dXh[b, n, ch] = dXh^T[ch, n]
      __nv_bfloat16* obase = p.out + ((long long)t.b * p.N) * p.C + ch;
      mbar_wait(dx_full, it & 1);
      tc_fence_after();
#pragma unroll 1
      for (int grp = cq; grp < ngrp + 2; grp += 2) {
        const bool has = grp < ngrp;
        const bool last = grp + 2 >= ngrp;
        uint32_t v[16];
        if (has) {
          tmem_ld_32x32b_x16(tmem_base + 4 * TM_CH + grp * 16 + lane_addr, v);
          tmem_ld_wait();
        }
        if (last) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) tm_arrive_leader(dx_empty, is_leader);
        }
        if (has && ch_ok) {
          __nv_bfloat16* po = obase + (long long)(grp * 16) * p.C;
          if (grp * 16 + 16 <= p.N) {
#pragma unroll
            for (int i = 0; i < 16; ++i, po += p.C) stg_u16(po, bf16_bits(__uint_as_float(v[i])));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i, po += p.C)
              if (grp * 16 + i < p.N) stg_u16(po, bf16_bits(__uint_as_float(v[i])));
          }
        }
        if (last) break;
      }
    }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
  if (p.db1 != nullptr)
    for (int i = threadIdx.x; i < p.Ds; i += TM_BWD_THREADS) red_add_f32(p.db1 + i, sdb[i] + sdb[p.n_chunks * TM_CH + i]);
}

// W [rows, cols] -> padded copy [rows, ld] (zero fill) and/or transposed copy [cols, ldt] (zero fill): the K-major weight
// operands of the fused kernels (a few hundred KB per block, once per step)
__global__ void tokmix_prepare_kernel(const __nv_bfloat16* __restrict__ w, int rows, int cols, __nv_bfloat16* __restrict__ pad,
                                      int ld, __nv_bfloat16* __restrict__ tr, int ldt) {
  const long long total_p = pad ? (long long)rows * ld : 0;
  const long long total_t = tr ? (long long)cols * ldt : 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total_p + total_t; i += (long long)gridDim.x * blockDim.x) {
    if (i < total_p) {
      const int r = (int)(i / ld), c = (int)(i % ld);
      pad[i] = c < cols ? w[(long long)r * cols + c] : __float2bfloat16(0.f);
    } else {
      const long long k = i - total_p;
      const int c = (int)(k / ldt), r = (int)(k % ldt);
      tr[k] = r < rows ? w[(long long)r * cols + c] : __float2bfloat16(0.f);
    }
  }
}

}  // namespace vmlp
