// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor),
// tcgen05 (alloc / mma / commit / ld / fences).  No library code, no CUTLASS.
// Everything here is device-side glue used by gemm_sm100.cuh.
#pragma once
#include <cuda.h>          // CUtensorMap (type only; the driver entry point is fetched at run time)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdio>

namespace vmlp {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One elected lane of a fully converged warp.  Issuing tcgen05 / TMA instructions (uniform-datapath SASS: UTCHMMA,
// UTMALDG, UTMASTG, UTCBAR) under this predicate lets the compiler keep their operands in uniform registers; issuing
// them under `lane == 0` instead costs an ELECT / R2UR.BROADCAST / BRA.U.ANY loop per instruction (~250 cycles).
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "@px mov.s32 %0, 1;\n"
      "}\n"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with an explicit suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint
// expires) instead of spinning -- a spinning epilogue warp costs the math warps of its scheduler their issue slots
// (ncu: 22 % of all executed instructions of the GELU GEMM were try_wait spins before this).
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must end the kernel (CUDA error on the host) instead of hanging the GPU box.
// 2^31 cycles is > 1 s at any B200 clock.  Before trapping, lane 0 of the waiting warp records (block, warp, barrier
// address, parity) in a host-mapped buffer (vmlp_debug_read): a trap discards the device printf buffer, host memory
// survives the dead context.
__device__ unsigned int* g_vmlp_dbg = nullptr;      // host-mapped: [0] = record count, then 4 words per record
__device__ __noinline__ void mbar_timeout(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0 && g_vmlp_dbg != nullptr) {
    const unsigned int slot = atomicAdd_system(g_vmlp_dbg, 1u);
    if (slot < 200) {
      volatile unsigned int* r = g_vmlp_dbg + 4 + 4 * slot;
      r[0] = blockIdx.x; r[1] = threadIdx.x >> 5; r[2] = smem_u32(bar); r[3] = parity;
    }
    __threadfence_system();
  }
  const long long t1 = clock64();
  while (clock64() - t1 < (1ll << 27)) {}      // let the other waiters of the same deadlock record theirs
  __trap();
}
// SLEEP_NS > 0: back off with nanosleep between polls.  The hardware-suspended try_wait comes back on every update of
// the barrier word (each of 32 arrivals, each TMA transaction), so a service warp that waits most of the time (TMA
// producer, MMA issuer, store warp) otherwise polls continuously: in the first fused token-mixing kernel those three
// warps executed 29 % of all instructions of the SM, taking issue slots from the 16 epilogue warps that bound the kernel.
template <int SLEEP_NS = 0>
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (SLEEP_NS > 0) __nanosleep(SLEEP_NS);
    if ((++polls & 255u) == 0 && clock64() - t0 > (1ll << 31)) mbar_timeout(bar, parity);
  }
}

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 3-D tiled load, global -> shared, completion on an mbarrier (bytes).
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 4-D tiled load (channels-last image tiles with halo: coordinates may be negative / past the edge, the engine zero-fills).
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                            int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of a tile (no shared memory, no completion tracking): lets the producer run a whole output tile ahead of
// the SMEM ring, so the later cp.async.bulk.tensor of the same box is an L2 hit instead of an HBM round trip.
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* map, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 3-D tiled store, shared -> global (bulk async group).
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* smem_src, int c0, int c1,
                                             int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------- tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; bf16 inputs, fp32 accumulate; issued by ONE thread.
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// All previously issued MMAs of this thread arrive on `bar` when complete
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns.
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// Two 16-column loads and the wait in ONE asm statement: the destination registers of tcgen05.ld are written
// asynchronously until tcgen05.wait::ld, and a register spill (or copy) the compiler might place between separate
// statements would read them too early -- seen as garbage gradients in a register-starved build of the backward kernel.
__device__ __forceinline__ void tmem_ld_x16_pair_wait(uint32_t ta, uint32_t tb, uint32_t (&a)[16], uint32_t (&b)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
        "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15]),
        "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]),
        "=r"(b[8]), "=r"(b[9]), "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
      : "r"(ta), "r"(tb)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16_wait(uint32_t ta, uint32_t (&a)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]),
        "=r"(a[8]), "=r"(a[9]), "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
      : "r"(ta)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 UMMA), 128-byte swizzle.
//   bits [0,14)  start address >> 4          bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4     bits [46,48) version = 1 (Blackwell)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes,
                                                         uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for tcgen05.mma kind::f16, bf16 x bf16 -> fp32.
//   [4,6) c_format=1 (F32)  [7,10) a_format=1 (BF16)  [10,13) b_format=1 (BF16)
//   [15] a_major (0 = K, 1 = MN)  [16] b_major  [17,23) N>>3  [24,29) M>>4
__host__ __device__ inline uint32_t umma_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 1u << 7;
  d |= 1u << 10;
  d |= static_cast<uint32_t>(a_mn_major & 1) << 15;
  d |= static_cast<uint32_t>(b_mn_major & 1) << 16;
  d |= static_cast<uint32_t>(n >> 3) << 17;
  d |= static_cast<uint32_t>(m >> 4) << 24;
  return d;
}

// ----------------------------------------------------------------------------- misc
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// non-blocking arrival on a named barrier (the warps that only produce; the consumer warp calls named_bar_sync)
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ldg_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float bf16lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// erf-GELU with the exact-form semantics of torch nn.GELU() (mlp_mixer.py:21): gelu(z) = z * Phi(z),
// Phi via the Abramowitz-Stegun 7.1.26 erfc form (|err| <= 1.5e-7 plus MUFU rcp/ex2 rounding ~1e-7; both far
// below the bf16 quantum of the stored result).  2 MUFU + ~12 FMA-pipe ops per element.
//   Phi(z) = h            (z <  0)      h = 0.5 * poly(t) * exp(-z^2/2),  t = 1 / (1 + p |z| / sqrt2)
//          = 1 - h        (z >= 0)
template <bool WITH_GRAD>
__device__ __forceinline__ float gelu_erf_t(float z, float& dgelu) {
  const float az = fabsf(z);
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, az, 1.0f));
  float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  p = fmaf(p, t, 0.5f * 1.421413741f);
  p = fmaf(p, t, 0.5f * -0.284496736f);
  p = fmaf(p, t, 0.5f * 0.254829592f);
  const float e = ex2_approx(az * az * (-0.5f * 1.4426950408889634f));   // exp(-z^2/2)
  const float h = p * t * e;
  const float cdf = (z < 0.f) ? h : 1.0f - h;
  if (WITH_GRAD) dgelu = fmaf(z * 0.3989422804014327f, e, cdf);           // Phi(z) + z * phi(z)
  return z * cdf;
}
__device__ __forceinline__ float gelu_erf(float z) {
  float unused;
  return gelu_erf_t<false>(z, unused);
}
__device__ __forceinline__ float dgelu_erf(float z) {
  float d;
  gelu_erf_t<true>(z, d);
  return d;
}


// ----------------------------------------------------------------------------- packed fp32x2 math (sm_100: FFMA2 / FMUL2 / FADD2)
// Two fp32 lanes per instruction: the GELU epilogues are issue-bound (ncu: 68 % issue-active, 23 thread instructions
// per element), so every packed op is an issue slot saved.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// gelu_erf_t for a pair of pre-activations (same formula, same constants): the Horner steps, the products and the
// final gelu / gelu' combine run packed; |z|, the two MUFUs and the z < 0 select stay per element.
// 23 instructions per pair (11 packed + 4 MUFU + 8 scalar) against 36 scalar ones.
template <bool WITH_GRAD>
__device__ __forceinline__ void gelu_erf_pair(f32x2 z, f32x2& gelu, f32x2& dgelu) {
  float z0, z1;
  unpack2(z, z0, z1);
  const float t0 = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, fabsf(z0), 1.0f));
  const float t1 = rcp_approx(fmaf(0.3275911f * 0.70710678118654752f, fabsf(z1), 1.0f));
  const f32x2 t = pack2(t0, t1);
  f32x2 p = fma2(pack2(0.5f * 1.061405429f, 0.5f * 1.061405429f), t, pack2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
  p = fma2(p, t, pack2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
  p = fma2(p, t, pack2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
  p = fma2(p, t, pack2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
  p = mul2(p, t);
  // exp(-z^2/2) = 2^(-(z * s)^2), s = sqrt(log2(e) / 2)
  const f32x2 w = mul2(z, pack2(0.84932180028801907f, 0.84932180028801907f));
  float w0, w1;
  unpack2(mul2(w, w), w0, w1);
  const float e0 = ex2_approx(-w0), e1 = ex2_approx(-w1);
  const f32x2 e = pack2(e0, e1);
  float h0, h1;
  unpack2(mul2(p, e), h0, h1);
  const float c0 = (z0 < 0.f) ? h0 : 1.0f - h0;
  const float c1 = (z1 < 0.f) ? h1 : 1.0f - h1;
  const f32x2 cdf = pack2(c0, c1);
  gelu = mul2(z, cdf);
  if (WITH_GRAD) dgelu = fma2(mul2(z, pack2(0.3989422804014327f, 0.3989422804014327f)), e, cdf);   // Phi(z) + z * phi(z)
}

// erf-GELU with ONE MUFU per element (the epilogues of the fused token-mixing kernels are MUFU-bound with the rcp + ex2 of
// gelu_erf_pair: 16 lanes/clk/SM).  Abramowitz-Stegun 7.1.28:  erfc(x) = P(x)^-16,  P = 1 + a1 x + ... + a6 x^6, x >= 0,
// |err| <= 3e-7.  With a = |z| and the 1/sqrt2 of x = a/sqrt2 folded into the coefficients:
//     r = 1 / P(a),  h = 0.5 r^16 = 1 - Phi(a),   gelu(z) = 0.5 z + a (0.5 - h)                       (exact identity in h)
//     phi(a) = -dh/da = 8 P'(a) r^17,             gelu'(z) = 0.5 + sgn(z) (0.5 + r^16 (8 a P'(a) r - 0.5))
// i.e. the density comes from the DERIVATIVE of the same rational form -- no exp.  Against fp64 erf/exp over |z| <= 9:
// |gelu err| <= 9e-7, |gelu' err| <= 1.5e-6 (tools/ check in DESIGN.md), far below the bf16 quantum of what is stored.
__device__ __forceinline__ f32x2 abs2(f32x2 v) {
  f32x2 r;
  asm("and.b64 %0, %1, 0x7fffffff7fffffff;" : "=l"(r) : "l"(v));
  return r;
}
#define VMLP_P2(c) pack2((c), (c))
template <bool WITH_GRAD>
__device__ __forceinline__ void gelu_rcp16_pair(f32x2 z, f32x2& gelu, f32x2& dgelu) {
  constexpr float B1 = 0.0705230784f * 0.70710678118654752f, B2 = 0.0422820123f * 0.5f,
                  B3 = 0.0092705272f * 0.35355339059327376f, B4 = 0.0001520143f * 0.25f,
                  B5 = 0.0002765672f * 0.17677669529663688f, B6 = 0.0000430638f * 0.125f;
  const f32x2 a = abs2(z);
  f32x2 pp = fma2(VMLP_P2(B6), a, VMLP_P2(B5));
  pp = fma2(pp, a, VMLP_P2(B4));
  pp = fma2(pp, a, VMLP_P2(B3));
  pp = fma2(pp, a, VMLP_P2(B2));
  pp = fma2(pp, a, VMLP_P2(B1));
  pp = fma2(pp, a, VMLP_P2(1.0f));
  float p0, p1;
  unpack2(pp, p0, p1);
  const f32x2 r = pack2(rcp_approx(p0), rcp_approx(p1));
  f32x2 r16 = mul2(r, r);
  r16 = mul2(r16, r16);
  r16 = mul2(r16, r16);
  r16 = mul2(r16, r16);
  if (!WITH_GRAD) {
    const f32x2 g = fma2(r16, VMLP_P2(-0.5f), VMLP_P2(0.5f));            // Phi(a) - 0.5
    gelu = fma2(a, g, mul2(z, VMLP_P2(0.5f)));
  } else {
    f32x2 q = fma2(VMLP_P2(6.0f * B6), a, VMLP_P2(5.0f * B5));
    q = fma2(q, a, VMLP_P2(4.0f * B4));
    q = fma2(q, a, VMLP_P2(3.0f * B3));
    q = fma2(q, a, VMLP_P2(2.0f * B2));
    q = fma2(q, a, VMLP_P2(B1));                                          // P'(a)
    const f32x2 t = mul2(mul2(a, q), r);
    const f32x2 u = fma2(t, VMLP_P2(8.0f), VMLP_P2(-0.5f));
    const f32x2 w = fma2(r16, u, VMLP_P2(0.5f));                          // Phi(a) + a phi(a) - 0.5  (>= 0)
    f32x2 ws;
    uint32_t lo, hi;                                                      // copysign(w, z) per half: (z & sign) | w, immLut 0xEA
    asm("lop3.b32 %0, %1, 0x80000000, %2, 0xea;" : "=r"(lo) : "r"((uint32_t)z), "r"((uint32_t)w));
    asm("lop3.b32 %0, %1, 0x80000000, %2, 0xea;" : "=r"(hi) : "r"((uint32_t)(z >> 32)), "r"((uint32_t)(w >> 32)));
    ws = static_cast<f32x2>(lo) | (static_cast<f32x2>(hi) << 32);
    dgelu = add2(ws, VMLP_P2(0.5f));
    gelu = z;   // not needed by the callers of the gradient form
  }
}

// volatile twins: the compiler keeps volatile asm statements in program order, which is how gelu_rcp16_x4 imposes its
// interleaved schedule
__device__ __forceinline__ f32x2 fma2v(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2v(f32x2 a, f32x2 b) {
  f32x2 r;
  asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float rcp_approx_v(float x) {
  float r;
  asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// NP independent element pairs through the forward form of gelu_rcp16_pair with every step issued for all four before
// the next one (NP = 8 in the token kernel): the chain is 14 dependent operations deep (6 Horner steps, the reciprocal, 4 squarings, 2 more), and left
// to itself ptxas serialises the pairs to stay inside the register budget -- the epilogue warps then sit in fixed-latency
// dependency stalls (ncu "wait": 2.4 warps per issue) instead of filling the FMA pipe.
template <int NP>
__device__ __forceinline__ void gelu_rcp16_xn(const f32x2 (&z)[NP], f32x2 (&gelu)[NP]) {
  constexpr float B1 = 0.0705230784f * 0.70710678118654752f, B2 = 0.0422820123f * 0.5f,
                  B3 = 0.0092705272f * 0.35355339059327376f, B4 = 0.0001520143f * 0.25f,
                  B5 = 0.0002765672f * 0.17677669529663688f, B6 = 0.0000430638f * 0.125f;
  f32x2 a[NP], pp[NP];
#pragma unroll
  for (int i = 0; i < NP; ++i) a[i] = abs2(z[i]);
#pragma unroll
  for (int i = 0; i < NP; ++i) pp[i] = fma2v(VMLP_P2(B6), a[i], VMLP_P2(B5));
#pragma unroll
  for (int i = 0; i < NP; ++i) pp[i] = fma2v(pp[i], a[i], VMLP_P2(B4));
#pragma unroll
  for (int i = 0; i < NP; ++i) pp[i] = fma2v(pp[i], a[i], VMLP_P2(B3));
#pragma unroll
  for (int i = 0; i < NP; ++i) pp[i] = fma2v(pp[i], a[i], VMLP_P2(B2));
#pragma unroll
  for (int i = 0; i < NP; ++i) pp[i] = fma2v(pp[i], a[i], VMLP_P2(B1));
#pragma unroll
  for (int i = 0; i < NP; ++i) pp[i] = fma2v(pp[i], a[i], VMLP_P2(1.0f));
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    float p0, p1;
    unpack2(pp[i], p0, p1);
    pp[i] = pack2(rcp_approx_v(p0), rcp_approx_v(p1));
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int i = 0; i < NP; ++i) pp[i] = mul2v(pp[i], pp[i]);
#pragma unroll
  for (int i = 0; i < NP; ++i) pp[i] = fma2v(pp[i], VMLP_P2(-0.5f), VMLP_P2(0.5f));      // Phi(a) - 0.5
#pragma unroll
  for (int i = 0; i < NP; ++i) gelu[i] = fma2v(a[i], pp[i], mul2v(z[i], VMLP_P2(0.5f)));
}

__device__ __forceinline__ uint32_t pack_bf16x2_f2(f32x2 v) {
  float lo, hi;
  unpack2(v, lo, hi);
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
}  // namespace vmlp

// ============================================================================= cta_group::2 (CTA pair) variants
namespace vmlp {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// Address of the same shared-memory object in the even (leader) CTA of the pair: clear the CTA-rank bit of the
// shared::cluster address (valid for 2-CTA clusters; CUTLASS calls this Sm100MmaPeerBitMask).
__device__ __forceinline__ uint32_t leader_cta_addr(uint32_t smem_addr) { return smem_addr & 0xFEFFFFFFu; }

// TMA load whose completion is signalled on the LEADER CTA's mbarrier (data lands in the issuing CTA's smem).
__device__ __forceinline__ void tma_load_3d_2cta(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1,
                                                 int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
      "[%2];"
      ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_cta_addr(smem_u32(bar))), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// arrive on the mbarrier at the same offset in CTA `cta` of the cluster.  No `.release.cluster`: that form compiles to
// MEMBAR.ALL.GPU + ERRBAR in front of the arrive (8 % of the CTA-pair GELU GEMM's stall samples).  The only thing this
// arrive publishes is "my tcgen05.ld reads of the accumulator are done", which tcgen05.wait::ld + fence::before_thread_sync
// already order.
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2cta() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A * B with M = 256 split over the CTA pair; issued by ONE thread of the leader CTA.
__device__ __forceinline__ void umma_bf16_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// completion of all prior MMAs arrives on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2cta_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// ----------------------------------------------------------------------------- lean main-loop forms
// Address-taking variants for the producer / MMA loops: operands are plain 32-bit shared-window addresses and
// descriptor words that the caller advances with one integer add per stage / k-step.
__device__ __forceinline__ void mbar_arrive_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// CG = 2: `bar` is the leader CTA's barrier (leader_cta_addr), data lands in the issuing CTA's shared memory.
template <int CG>
__device__ __forceinline__ void tma_load_3d_u32(uint32_t smem_dst, uint64_t map, uint32_t bar, int c0, int c1, int c2) {
  if (CG == 2)
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], "
        "[%2];" ::"r"(smem_dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// tcgen05.mma with the two shared-memory descriptors given as (low word, shared high word).
template <int CG>
__device__ __forceinline__ void umma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                             uint32_t idesc, uint32_t accumulate) {
  if (CG == 2)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 da, {%1, %3};\n"
        "mov.b64 db, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

}  // namespace vmlp
