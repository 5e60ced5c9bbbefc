// Host side of libvmlp_b200.so: TMA descriptor encoding, kernel launchers, and the block-level
// orchestration (which GEMM consumes which operand in which major-ness).  C ABI in include/vmlp_b200.h.
#include "../../include/vmlp_b200.h"
#include "gemm_sm100.cuh"
#include "rowwise.cuh"
#include "spatial.cuh"
#include "spatial2.cuh"
#include "tokmix_sm100.cuh"
#include "aux_sm100.cuh"

#include <cstdarg>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>

using namespace vmlp;

namespace {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};   // kernels launched by this library (bench.py reports it)
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_OK(expr)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess) return fail(VMLP_ELAUNCH, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct DeviceInfo {
  int sms = 0, cc_major = 0, cc_minor = 0;
  bool ok = false;
  unsigned int* dbg_host = nullptr;     // host-mapped record buffer of the bounded mbarrier waits (ptx.cuh)
};
constexpr int DBG_WORDS = 4 + 4 * 200;
const DeviceInfo& device_info() {
  static DeviceInfo info[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) {
    static DeviceInfo bad;
    return bad;
  }
  static std::once_flag once[64];
  DeviceInfo& d = info[dev];
  std::call_once(once[dev], [&] {
    cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&d.cc_major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&d.cc_minor, cudaDevAttrComputeCapabilityMinor, dev);
    d.ok = d.sms > 0;
    void* h = nullptr;
    if (d.ok && cudaHostAlloc(&h, DBG_WORDS * sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess) {
      memset(h, 0, DBG_WORDS * sizeof(unsigned int));
      void* dptr = nullptr;
      if (cudaHostGetDevicePointer(&dptr, h, 0) == cudaSuccess &&
          cudaMemcpyToSymbol(g_vmlp_dbg, &dptr, sizeof(dptr)) == cudaSuccess)
        d.dbg_host = static_cast<unsigned int*>(h);
    }
  });
  return d;
}

// ---- process-wide knobs, read once (getenv is not cheap and not re-entrant against setenv)
struct EnvKnobs {
  int force_cta_group = 0;   // VMLP_FORCE_CTA_GROUP = 1 | 2 (A/B measurements)
  bool debug = false;        // VMLP_DEBUG
  bool tokmix = true;        // VMLP_TOKMIX = 0 keeps the unfused token-mixing GEMM sequence
};
const EnvKnobs& env_knobs() {
  static const EnvKnobs k = [] {
    EnvKnobs e;
    if (const char* v = getenv("VMLP_FORCE_CTA_GROUP")) e.force_cta_group = atoi(v) == 2 ? 2 : 1;
    e.debug = getenv("VMLP_DEBUG") != nullptr;
    if (const char* v = getenv("VMLP_TOKMIX")) e.tokmix = atoi(v) != 0;
    return e;
  }();
  return k;
}

// Per-device one-time opt-in to > 48 KB of dynamic shared memory (cudaFuncSetAttribute applies to the CURRENT device
// only, so the flag is kept per device; atomics make concurrent first calls from several host threads benign -- at
// worst the attribute is set twice to the same value).
template <typename K>
int smem_optin(K kern, int bytes, std::atomic<int>* done_bytes /* [64], zero-initialised */) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (done_bytes[dev].load(std::memory_order_acquire) < bytes) {
    CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done_bytes[dev].store(bytes, std::memory_order_release);
  }
  return VMLP_OK;
}
// Resident blocks per SM of a persistent kernel, cached per device.
template <typename K>
int occupancy_cached(K kern, int threads, int smem, std::atomic<int>* occ /* [64] */, int* out) {
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  int v = occ[dev].load(std::memory_order_acquire);
  if (v == 0) {
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, threads, smem));
    if (v < 1) v = 1;
    occ[dev].store(v, std::memory_order_release);
  }
  *out = v;
  return VMLP_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// 3-D bf16 tensor map: dims (inner, rows, batch), 128B swizzle, zero OOB fill.
int make_map(CUtensorMap* m, const void* ptr, int64_t inner, int64_t rows, int64_t batch, int64_t ld,
             int64_t bstride, int box_inner, int box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(VMLP_ELAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
  if (!aligned16(ptr) || (ld % 8) != 0 || (batch > 1 && (bstride % 8) != 0))
    return fail(VMLP_EALIGN, "operand %p ld %lld bs %lld must be 16-byte aligned", ptr, (long long)ld,
                (long long)bstride);
  if (batch <= 1 || bstride == 0) { batch = 1; bstride = rows * ld; }
  cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bstride * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(VMLP_EINVAL, "cuTensorMapEncodeTiled failed (%d): inner %lld rows %lld batch %lld ld %lld", (int)r,
                (long long)inner, (long long)rows, (long long)batch, (long long)ld);
  return VMLP_OK;
}

// 4-D bf16 tensor map over a channels-last image tensor [B, H, W, C]: dims (C, W, H, B), no swizzle, zero OOB fill.
int make_map_nhwc(CUtensorMap* m, const void* ptr, int B, int H, int W, int C, int box_c, int box_w, int box_h) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return fail(VMLP_ELAUNCH, "cuTensorMapEncodeTiled entry point unavailable");
  if (!aligned16(ptr) || (C % 8) != 0) return fail(VMLP_EALIGN, "image tensor %p C=%d must be 16-byte aligned rows", ptr, C);
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VMLP_EINVAL, "cuTensorMapEncodeTiled (NHWC) failed (%d): B %d H %d W %d C %d", (int)r, B, H, W, C);
  return VMLP_OK;
}

template <int BN, int EPI, int CG>
int launch_gemm_t(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td, const CUtensorMap& td2,
                  const CUtensorMap& tpf, const GemmParams& p, int grid, cudaStream_t st) {
  auto kern = gemm_bf16_sm100<BN, EPI, CG>;
  using SM = GemmSmem<BN, EPI, CG>;
  static std::atomic<int> optin[64];
  if (int rc = smem_optin(kern, SM::TOTAL, optin)) return rc;
  if (CG == 1) {
    kern<<<grid, GEMM_THREADS, SM::TOTAL, st>>>(ta, tb, td, td2, tpf, p);
  } else {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = SM::TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;   // CTA pair: tcgen05 cta_group::2
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, td, td2, tpf, p));
  }
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

template <int BN, int CG>
int launch_gemm_bn(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& td,
                   const CUtensorMap& td2, const CUtensorMap& tpf, const GemmParams& p, int grid, cudaStream_t st) {
  switch (epi) {
    case EPI_STORE: return launch_gemm_t<BN, EPI_STORE, CG>(ta, tb, td, td2, tpf, p, grid, st);
    case EPI_GELU: return launch_gemm_t<BN, EPI_GELU, CG>(ta, tb, td, td2, tpf, p, grid, st);
    case EPI_RESID: return launch_gemm_t<BN, EPI_RESID, CG>(ta, tb, td, td2, tpf, p, grid, st);
    case EPI_DGELU: return launch_gemm_t<BN, EPI_DGELU, CG>(ta, tb, td, td2, tpf, p, grid, st);
    case EPI_ATOMIC: return launch_gemm_t<BN, EPI_ATOMIC, CG>(ta, tb, td, td2, tpf, p, grid, st);
    case EPI_MUL: return launch_gemm_t<BN, EPI_MUL, CG>(ta, tb, td, td2, tpf, p, grid, st);
    case EPI_GELU_ONLY: return launch_gemm_t<BN, EPI_GELU_ONLY, CG>(ta, tb, td, td2, tpf, p, grid, st);
    case EPI_RESID_DUAL: return launch_gemm_t<BN, EPI_RESID_DUAL, CG>(ta, tb, td, td2, tpf, p, grid, st);
    case EPI_MUL_DUAL: return launch_gemm_t<BN, EPI_MUL_DUAL, CG>(ta, tb, td, td2, tpf, p, grid, st);
  }
  return fail(VMLP_EINVAL, "unknown epilogue %d", epi);
}

int gemm_impl(const vmlp_gemm_args& g, cudaStream_t st) {
  const DeviceInfo& dv = device_info();
  if (!dv.ok || dv.cc_major != 10)
    return fail(VMLP_EARCH, "device compute capability %d.%d is not sm_100 (no fallback)", dv.cc_major, dv.cc_minor);
  if (g.M <= 0 || g.N <= 0 || g.K <= 0 || g.batch <= 0) return fail(VMLP_EINVAL, "empty GEMM %d %d %d %d", g.M, g.N, g.K, g.batch);
  if (!g.A.ptr || !g.B.ptr) return fail(VMLP_EINVAL, "null operand");
  const int epi = g.epilogue;
  if (epi != EPI_ATOMIC && (g.N % 8) != 0) return fail(VMLP_EINVAL, "N=%d must be a multiple of 8", g.N);
  int bn = g.block_n;
  if (bn == 0) {
    bn = (g.N <= 128) ? 128 : 256;
    // weight gradients whose N is a token count (196 -> 13 x 16): a 208-column tile instead of 256 (K-major B only)
    if (epi == EPI_ATOMIC && g.B.major == 0 && g.N > 128 && g.N <= 208) bn = 208;
  }
  if (bn != 128 && bn != 256 && bn != 208) return fail(VMLP_EINVAL, "block_n must be 128, 208 or 256");
  if (bn == 208 && (epi != EPI_ATOMIC || g.B.major != 0 || g.cta_group == 2))
    return fail(VMLP_EINVAL, "block_n 208 exists for EPI_ATOMIC with a K-major B operand on single CTAs only");

  // CTA-pair kernel (cta_group::2, 256 x 256 tiles) for the large un-batched GEMMs: channel-mixing fwd/dgrad/wgrad
  const bool one_output = (g.batch == 1 || g.contract_batch);
  int cg = g.cta_group;
  // ... and wherever a 256-row pair tile wastes no more rows than two 128-row tiles would (M = 196 token GEMMs: measured
  // 88 -> 80 us dgrad, 100 -> 91 us wgrad; M = 784 is 4 x 256 = 1024 vs 7 x 128 = 896 rows and stays on single CTAs)
  const long long pad1 = (g.M + 127) / 128 * 128, pad2 = (g.M + 255) / 256 * 256;
  if (cg == 0)
    cg = (bn == 256 && ((one_output && (g.M % 256 == 0 || g.M >= 4096) && g.M >= 512) || (pad1 == pad2 && g.M > 128))) ? 2 : 1;
  if (const int f = env_knobs().force_cta_group) cg = (f == 2 && bn == 256) ? 2 : 1;
  if (cg != 1 && cg != 2) return fail(VMLP_EINVAL, "cta_group must be 0, 1 or 2");
  if (cg == 2 && bn != 256) return fail(VMLP_EINVAL, "cta_group 2 needs block_n 256");

  // operand views
  const bool a_mn = g.A.major != 0, b_mn = g.B.major != 0;
  if (!a_mn ? (g.A.rows != g.M || g.A.cols != g.K) : (g.A.rows != g.K || g.A.cols != g.M))
    return fail(VMLP_EINVAL, "A view %lldx%lld does not match M=%d K=%d major=%d", (long long)g.A.rows,
                (long long)g.A.cols, g.M, g.K, g.A.major);
  if (!b_mn ? (g.B.rows != g.N || g.B.cols != g.K) : (g.B.rows != g.K || g.B.cols != g.N))
    return fail(VMLP_EINVAL, "B view %lldx%lld does not match N=%d K=%d major=%d", (long long)g.B.rows,
                (long long)g.B.cols, g.N, g.K, g.B.major);
  const bool a_batched = g.A.batch_stride != 0 && g.batch > 1;
  const bool b_batched = g.B.batch_stride != 0 && g.batch > 1;

  CUtensorMap ta, tb, td, td2, tpf;
  memset(&td, 0, sizeof(td));
  memset(&td2, 0, sizeof(td2));
  memset(&tpf, 0, sizeof(tpf));
  int rc;
  // K-major: inner = K, box (64, 128 | BN).  MN-major: inner = M/N, box (64, 64) per swizzle atom.
  rc = make_map(&ta, g.A.ptr, g.A.cols, g.A.rows, a_batched ? g.batch : 1, g.A.ld, g.A.batch_stride, 64,
                a_mn ? GEMM_BK : GEMM_BM);
  if (rc) return rc;
  rc = make_map(&tb, g.B.ptr, g.B.cols, g.B.rows, b_batched ? g.batch : 1, g.B.ld, g.B.batch_stride, 64,
                b_mn ? GEMM_BK : bn / cg);
  if (rc) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = g.M;
  p.N = g.N;
  p.tiles_m = (g.M + GEMM_BM * cg - 1) / (GEMM_BM * cg);
  p.tiles_n = (g.N + bn - 1) / bn;
  p.kbatch = g.contract_batch ? 1 : 0;
  p.batch = p.kbatch ? 1 : g.batch;
  p.kpb = (g.K + GEMM_BK - 1) / GEMM_BK;
  p.k_blocks = p.kpb * (p.kbatch ? g.batch : 1);
  const int krem = g.K - (p.kpb - 1) * GEMM_BK;
  p.last_ksteps = (krem + 15) / 16;
  p.a_mn = a_mn;
  p.b_mn = b_mn;
  p.a_batched = a_batched;
  p.b_batched = b_batched;
  p.bias_mode = g.bias ? g.bias_mode : 0;
  p.bias = static_cast<const __nv_bfloat16*>(g.bias);
  p.colscale = static_cast<const __nv_bfloat16*>(g.colscale);
  p.aux = static_cast<const __nv_bfloat16*>(g.aux);
  p.aux_ld = g.aux_ld;
  p.aux_bs = g.aux_bs;
  p.out_f32 = g.out_f32;
  p.out_ld = g.out_ld;
  p.out_trans = (epi == EPI_ATOMIC && g.out_trans) ? 1 : 0;
  p.red_out = g.red_out;
  p.red_mode = g.red_out ? g.red_mode : 0;
  if (p.red_mode < 0 || p.red_mode > 2 || (p.red_mode && epi == EPI_ATOMIC)) return fail(VMLP_EINVAL, "bad red_mode");
  if (p.bias_mode == 1 && !aligned16(g.bias)) return fail(VMLP_EALIGN, "bias must be 16-byte aligned");
  if (p.colscale && !aligned16(p.colscale)) return fail(VMLP_EALIGN, "colscale must be 16-byte aligned");
  const bool needs_aux = (epi_is_resid(epi) || epi == EPI_DGELU || epi_is_mul(epi));
  if (needs_aux) {
    if (!g.aux) return fail(VMLP_EINVAL, "epilogue %d needs aux", epi);
    if (!aligned16(g.aux) || (g.aux_ld % 8) || (g.aux_bs % 8)) return fail(VMLP_EALIGN, "aux alignment");
    if (g.batch > 1 && !g.contract_batch && g.aux_bs == 0) return fail(VMLP_EINVAL, "batched output needs a batched aux operand");
  } else {
    p.aux = nullptr;
  }

  const int base_tiles = p.tiles_m * p.tiles_n * p.batch;
  int split = 1;
  if (epi == EPI_ATOMIC) {
    if (!g.out_f32) return fail(VMLP_EINVAL, "EPI_ATOMIC needs out_f32");
    split = g.split_k > 0 ? g.split_k : ((dv.sms / cg) / base_tiles);
    if (split < 1) split = 1;
    if (split > p.k_blocks) split = p.k_blocks;
    const int per = (p.k_blocks + split - 1) / split;
    split = (p.k_blocks + per - 1) / per;     // no empty split
  } else {
    if (!g.D) return fail(VMLP_EINVAL, "null output");
    rc = make_map(&td, g.D, g.N, g.M, p.batch, g.d_ld, g.d_bs, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);   // per-warp store box: 32 cols x 32 rows
    if (rc) return rc;
    if (epi == EPI_DGELU || epi == EPI_MUL) {
      // single-output aux epilogues: the aux operand is fetched by TMA (32 x 32 boxes); its map rides in the D2 slot
      rc = make_map(&td2, g.aux, g.N, g.M, (g.aux_bs != 0 ? p.batch : 1), g.aux_ld, g.aux_bs, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
      // ... and is L2-prefetched by the producer, one [128 rows x BN cols] box per CTA tile
      rc = make_map(&tpf, g.aux, g.N, g.M, (g.aux_bs != 0 ? p.batch : 1), g.aux_ld, g.aux_bs, bn, GEMM_BM, CU_TENSOR_MAP_SWIZZLE_NONE);
      if (rc) return rc;
    }
    if (epi_is_dual(epi)) {
      if (!g.D2) return fail(VMLP_EINVAL, "dual-output epilogue needs D2");
      rc = make_map(&td2, g.D2, g.N, g.M, p.batch, g.d2_ld, g.d2_bs, 32, 32, CU_TENSOR_MAP_SWIZZLE_64B);
      if (rc) return rc;
    }
  }
  p.split_k = split;
  const long long total = (long long)base_tiles * split;
  if (total >= (1ll << 22)) return fail(VMLP_EINVAL, "%lld output tiles exceed the 2^22 limit of the tile decode", total);
  p.inv_tiles_n = 1.0f / (float)p.tiles_n;
  p.inv_tiles_m = 1.0f / (float)p.tiles_m;
  p.inv_batch = 1.0f / (float)p.batch;
  const long long clusters = dv.sms / cg;
  const int grid = (int)(total < clusters ? total : clusters) * cg;
  if (cg == 2) return launch_gemm_bn<256, 2>(epi, ta, tb, td, td2, tpf, p, grid, st);
  if (bn == 208) return launch_gemm_t<208, EPI_ATOMIC, 1>(ta, tb, td, td2, tpf, p, grid, st);
  if (bn == 256) return launch_gemm_bn<256, 1>(epi, ta, tb, td, td2, tpf, p, grid, st);
  return launch_gemm_bn<128, 1>(epi, ta, tb, td, td2, tpf, p, grid, st);
}

int rw_grid(long long rows) {
  const DeviceInfo& dv = device_info();
  long long blocks = (rows + RW_WARPS - 1) / RW_WARPS;
  const long long cap = (long long)(dv.sms > 0 ? dv.sms : 148) * 8;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}
int vpl_for(int C) { return (C / 8 + 31) / 32; }

#define DISPATCH_VPL(C, ...)                                                     \
  do {                                                                           \
    const int vpl__ = vpl_for(C);                                                \
    if (vpl__ <= 1) { constexpr int VPL = 1; __VA_ARGS__; }                      \
    else if (vpl__ <= 2) { constexpr int VPL = 2; __VA_ARGS__; }                 \
    else if (vpl__ <= 3) { constexpr int VPL = 3; __VA_ARGS__; }                 \
    else if (vpl__ <= 4) { constexpr int VPL = 4; __VA_ARGS__; }                 \
    else if (vpl__ <= 6) { constexpr int VPL = 6; __VA_ARGS__; }                 \
    else if (vpl__ <= 8) { constexpr int VPL = 8; __VA_ARGS__; }                 \
    else if (vpl__ <= 16) { constexpr int VPL = 16; __VA_ARGS__; }               \
    else return fail(VMLP_EINVAL, "row length %d too large (max 4096)", C);      \
  } while (0)

typedef const __nv_bfloat16* cbf;
typedef __nv_bfloat16* bf;

// ------------------------------------------------------------------------------------------- GEMM arg helpers
vmlp_operand opnd(const void* p, int64_t rows, int64_t cols, int64_t ld, int64_t bs, int major) {
  vmlp_operand o;
  o.ptr = p; o.rows = rows; o.cols = cols; o.ld = ld; o.batch_stride = bs; o.major = major;
  return o;
}
vmlp_gemm_args gemm_args(int M, int N, int K, int batch, vmlp_operand A, vmlp_operand B, int epi) {
  vmlp_gemm_args g;
  memset(&g, 0, sizeof(g));
  g.M = M; g.N = N; g.K = K; g.batch = batch; g.A = A; g.B = B; g.epilogue = epi;
  return g;
}

// LayerNorm backward (+ residual-gradient add, + dgamma / dbeta).  add_colsum / out_rowsum (both optional, C <= 1536
// and add != null only): the Mixer block backward folds its two output-bias gradients into this pass (rowwise.cuh).
int layernorm_bwd_impl(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, const float* mean,
                       const float* rstd, const void* gamma, const void* add, int64_t add_ld, void* dx,
                       int64_t dx_ld, float* dgamma, float* dbeta, int64_t rows, int32_t C, float* add_colsum,
                       float* out_rowsum, int row_period, vmlp_stream_t stream) {
  if (!dy || !x || !mean || !rstd || !gamma || !dx || !dgamma || !dbeta || rows <= 0 || (C % 8))
    return fail(VMLP_EINVAL, "layernorm_bwd args");
  if (!aligned16(dy) || !aligned16(x) || !aligned16(dx) || !aligned16(gamma) || (add && !aligned16(add)) ||
      (dy_ld % 8) || (x_ld % 8) || (dx_ld % 8) || (add_ld % 8))
    return fail(VMLP_EALIGN, "layernorm_bwd alignment");
  const bool extra = add_colsum != nullptr || out_rowsum != nullptr;
  if (extra && (C > 1536 || !add || row_period <= 0)) return fail(VMLP_EINVAL, "layernorm_bwd fused sums need C <= 1536 and add");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DeviceInfo& dv = device_info();
  long long blocks = (rows + RW_WARPS - 1) / RW_WARPS;
  const int grid = (int)(blocks < dv.sms * 6 ? blocks : dv.sms * 6);
  if (C <= 1536) {
    // warp-private column partials: 8 warps x 16 (24) x NV floats (<= 96 KB) needs the dynamic-smem opt-in
    if (extra) {
      DISPATCH_VPL(C, {
        auto kern = layernorm_bwd_kernel<VPL, 1, 1>;
        const size_t sh = (size_t)RW_WARPS * 24 * VPL * 32 * sizeof(float);
        static std::atomic<int> optin[64];
        if (int rc2 = smem_optin(kern, (int)sh, optin)) return rc2;
        kern<<<grid, RW_THREADS, sh, st>>>((cbf)dy, dy_ld, (cbf)x, x_ld, mean, rstd, (cbf)gamma, (cbf)add, add_ld, (bf)dx,
                                           dx_ld, dgamma, dbeta, rows, C, add_colsum, out_rowsum, row_period);
      });
    } else if (C <= 128) {
      // narrow rows: 4 (C <= 64) or 2 rows per warp; same 8 x 16 x 32 floats of warp-private partials
      const size_t sh = (size_t)RW_WARPS * 16 * 32 * sizeof(float);
      const long long blocks2 = (rows + RW_WARPS * (C <= 64 ? 4 : 2) - 1) / (RW_WARPS * (C <= 64 ? 4 : 2));
      const int grid2 = (int)(blocks2 < dv.sms * 6 ? blocks2 : dv.sms * 6);
      if (C <= 64)
        layernorm_bwd_kernel<1, 1, 0, 8><<<grid2, RW_THREADS, sh, st>>>((cbf)dy, dy_ld, (cbf)x, x_ld, mean, rstd, (cbf)gamma, (cbf)add, add_ld,
                                                                      (bf)dx, dx_ld, dgamma, dbeta, rows, C, nullptr, nullptr, 1);
      else
        layernorm_bwd_kernel<1, 1, 0, 16><<<grid2, RW_THREADS, sh, st>>>((cbf)dy, dy_ld, (cbf)x, x_ld, mean, rstd, (cbf)gamma, (cbf)add, add_ld,
                                                                       (bf)dx, dx_ld, dgamma, dbeta, rows, C, nullptr, nullptr, 1);
    } else {
      DISPATCH_VPL(C, {
        auto kern = layernorm_bwd_kernel<VPL, 1, 0>;
        const size_t sh = (size_t)RW_WARPS * 16 * VPL * 32 * sizeof(float);
        static std::atomic<int> optin[64];
        if (int rc2 = smem_optin(kern, (int)sh, optin)) return rc2;
        kern<<<grid, RW_THREADS, sh, st>>>((cbf)dy, dy_ld, (cbf)x, x_ld, mean, rstd, (cbf)gamma, (cbf)add, add_ld, (bf)dx,
                                           dx_ld, dgamma, dbeta, rows, C, nullptr, nullptr, 1);
      });
    }
  } else {
    DISPATCH_VPL(C, (layernorm_bwd_kernel<VPL, 0, 0><<<grid, RW_THREADS, 16 * VPL * 32 * sizeof(float), st>>>(
                        (cbf)dy, dy_ld, (cbf)x, x_ld, mean, rstd, (cbf)gamma, (cbf)add, add_ld, (bf)dx, dx_ld, dgamma,
                        dbeta, rows, C, nullptr, nullptr, 1)));
  }
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

// ------------------------------------------------------------------------------------------- fused token-mixing MLP
constexpr int TM_SMEM_MAX = 227 * 1024;
long long* g_tokmix_trace = nullptr;     // bring-up only: device buffer for the clock64 timeline of CTA 0 (vmlp_tokmix_set_trace)

// Fills the shape-derived fields; returns the dynamic shared memory the kernel needs (0 = shape not supported).
int tokmix_plan(TokParams& p, int B, int N, int C, int Ds, bool backward) {
  memset(&p, 0, sizeof(p));
  if (B <= 0 || N <= 0 || C <= 0 || Ds <= 0 || (C % 8) || (Ds % 8) || N > 256 || Ds > TM_MAX_DS) return 0;
  p.B = B; p.N = N; p.C = C; p.Ds = Ds;
  p.NT = (N + 15) & ~15;
  p.tiles_c = (C + 127) / 128;
  p.n_tiles = B * p.tiles_c;
  p.n_pairs = (p.n_tiles + 1) / 2;
  p.n_chunks = (Ds + TM_CH - 1) / TM_CH;
  p.last_n1 = (Ds - (p.n_chunks - 1) * TM_CH + 15) & ~15;
  p.ka = (p.NT + 63) / 64;
  p.wa_stage = p.ka * 4096;
  p.wb_stage = p.NT * 64;
  p.inv_tiles_c = 1.0f / (float)p.tiles_c;
  if ((long long)p.n_tiles >= (1ll << 22)) return 0;
  // shared memory: barriers + alignment slack, two activation-sized tiles (forward: Xh + residual/output; backward: Xh + dU),
  // weight rings, hidden tile(s), fp32 bias (+ d-bias sums).  Measured: without any weight traffic the kernels are as fast
  // as with it (the L2 -> SMEM stream is not the limiter), so a second hidden-tile buffer comes before ring depth.
  const int force_depth = [] { const char* e = getenv("VMLP_TM_DEPTH"); return e ? atoi(e) : 0; }();
  const int force_nhb = [] { const char* e = getenv("VMLP_TM_NHB"); return e ? atoi(e) : 0; }();
  const int bias_bytes = p.n_chunks * TM_CH * 4 * (backward ? 3 : 1) + (backward ? 0 : p.NT * 4);   // b1 (+ 2 x d b1 sums | b2)
  auto finish = [&](int nhb, int s_wa, int s_wb) -> int {
    const int bytes = TM_BAR_BYTES + 1024 + nhb * TM_HTILE + p.NT * 256 * 2 + bias_bytes +
                      s_wa * p.wa_stage * (backward ? 2 : 1) + s_wb * p.wb_stage;
    if (bytes > TM_SMEM_MAX) return 0;
    p.s_wa = s_wa; p.s_wb = s_wb; p.nhb = nhb;
    // never less than half an SM's shared memory: one CTA per SM, so the 512-column TMEM allocation of a CTA pair can
    // never wait for a co-resident CTA of another pair (allocation order across two SMs could deadlock)
    return bytes > 120 * 1024 ? bytes : 120 * 1024;
  };
  // backward with TWO dZ buffers = ping-pong epilogue groups (paid for with a one-stage W1^T ring).  Measured SLOWER at the
  // Mixer shapes (385 vs 275 us): one math warp per scheduler at a time cannot fill the FMA / MUFU pipes with gelu' chains,
  // two (lock step) do better.  Kept selectable for experiments: VMLP_TM_NHB=2.
  if (backward && !force_depth && force_nhb == 2)
    if (int r = finish(2, 2, 1)) return r;
  // forward: three hidden-tile buffers (a group never waits for G2 of its own previous chunk), at least two
  for (int nhb = backward ? 1 : 3; nhb >= (backward ? 1 : 2); --nhb) {
    if (force_nhb && nhb != force_nhb) continue;
    for (int depth = 4; depth >= 2; --depth) {
      if (force_depth && depth != force_depth) continue;
      if (int r = finish(nhb, depth, depth)) return r;
    }
  }
  return 0;
}

int tokmix_launch_cfg(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, const TokParams& p, int smem, int threads, cudaStream_t st) {
  const DeviceInfo& dv = device_info();
  memset(&cfg, 0, sizeof(cfg));
  const int clusters = p.n_pairs < dv.sms / 2 ? p.n_pairs : dv.sms / 2;
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return VMLP_OK;
}

int tokmix_fwd_impl(const void* xhat, const void* x, const void* w1_pad, int Np, const void* w2, const void* b1,
                    const void* b2, void* u, void* hT, int B, int N, int C, int Ds, cudaStream_t st) {
  const DeviceInfo& dv = device_info();
  if (!dv.ok || dv.cc_major != 10) return fail(VMLP_EARCH, "device is not sm_100 (no fallback)");
  TokParams p;
  const int smem = tokmix_plan(p, B, N, C, Ds, false);
  if (!smem) return fail(VMLP_EINVAL, "tokmix_fwd: unsupported shape B %d N %d C %d Ds %d", B, N, C, Ds);
  if (!xhat || !x || !w1_pad || !w2 || !b1 || !b2 || !u) return fail(VMLP_EINVAL, "tokmix_fwd null pointer");
  if (Np < p.NT || (Np % 8)) return fail(VMLP_EINVAL, "tokmix_fwd: padded weight pitch %d < ceil16(N) = %d", Np, p.NT);
  p.b1 = (cbf)b1; p.b2 = (cbf)b2; p.resid = (cbf)x; p.out = (bf)u;
  if (const char* e = getenv("VMLP_TM_FLAGS")) p.flags = atoi(e);     // profiling experiments only
  if (!aligned16(x) || !aligned16(u)) return fail(VMLP_EALIGN, "tokmix_fwd: x / u must be 16-byte aligned");
  p.trace = g_tokmix_trace;
  CUtensorMap tX, tW1, tW2, tH, tR, tU;
  int rc;
  if ((rc = make_map(&tX, xhat, C, N, B, C, (long long)N * C, 64, p.NT))) return rc;
  if ((rc = make_map(&tW1, w1_pad, p.NT, Ds, 1, Np, 0, 64, 32))) return rc;
  if ((rc = make_map(&tW2, w2, Ds, N, 1, Ds, 0, 64, p.NT / 2))) return rc;
  if (hT) { if ((rc = make_map(&tH, hT, Ds, C, B, Ds, (long long)C * Ds, 64, 128))) return rc; }
  else tH = tW2;
  // residual in / output out: [NT tokens x 128 channels] row-major tiles (256-byte rows, no swizzle); loads zero-fill and
  // stores clip tokens >= N and channels >= C
  if ((rc = make_map(&tR, x, C, N, B, C, (long long)N * C, 128, p.NT, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  if ((rc = make_map(&tU, u, C, N, B, C, (long long)N * C, 128, p.NT, CU_TENSOR_MAP_SWIZZLE_NONE))) return rc;
  static std::atomic<int> optin[64];
  if ((rc = smem_optin(tokmix_fwd_sm100, smem, optin))) return rc;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  tokmix_launch_cfg(cfg, attr, p, smem, TM_FWD_THREADS, st);
  CUDA_OK(cudaLaunchKernelEx(&cfg, tokmix_fwd_sm100, tX, tW1, tW2, tH, tR, tU, p, hT ? 1 : 0));
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int tokmix_bwd_impl(const void* xhat, const void* du, const void* w1_pad, const void* w2T_pad, int Np, const void* w1T,
                    const void* b1, void* dxhat, void* dzT, float* db1, int B, int N, int C, int Ds, cudaStream_t st) {
  const DeviceInfo& dv = device_info();
  if (!dv.ok || dv.cc_major != 10) return fail(VMLP_EARCH, "device is not sm_100 (no fallback)");
  TokParams p;
  const int smem = tokmix_plan(p, B, N, C, Ds, true);
  if (!smem) return fail(VMLP_EINVAL, "tokmix_bwd: unsupported shape B %d N %d C %d Ds %d", B, N, C, Ds);
  if (!xhat || !du || !w1_pad || !w2T_pad || !w1T || !b1 || !dxhat || !dzT) return fail(VMLP_EINVAL, "tokmix_bwd null pointer");
  if (Np < p.NT || (Np % 8)) return fail(VMLP_EINVAL, "tokmix_bwd: padded weight pitch %d < ceil16(N) = %d", Np, p.NT);
  p.b1 = (cbf)b1; p.out = (bf)dxhat; p.db1 = db1;
  if (const char* e = getenv("VMLP_TM_FLAGS")) p.flags = atoi(e);     // profiling experiments only
  p.trace = g_tokmix_trace;
  CUtensorMap tX, tDU, tW1, tW2T, tW1T, tDZ;
  int rc;
  if ((rc = make_map(&tX, xhat, C, N, B, C, (long long)N * C, 64, p.NT))) return rc;
  if ((rc = make_map(&tDU, du, C, N, B, C, (long long)N * C, 64, p.NT))) return rc;
  if ((rc = make_map(&tW1, w1_pad, p.NT, Ds, 1, Np, 0, 64, 32))) return rc;
  if ((rc = make_map(&tW2T, w2T_pad, p.NT, Ds, 1, Np, 0, 64, 32))) return rc;
  if ((rc = make_map(&tW1T, w1T, Ds, N, 1, Ds, 0, 64, p.NT / 2))) return rc;
  if ((rc = make_map(&tDZ, dzT, Ds, C, B, Ds, (long long)C * Ds, 64, 128))) return rc;
  static std::atomic<int> optin[64];
  if ((rc = smem_optin(tokmix_bwd_sm100, smem, optin))) return rc;
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  tokmix_launch_cfg(cfg, attr, p, smem, TM_BWD_THREADS, st);
  CUDA_OK(cudaLaunchKernelEx(&cfg, tokmix_bwd_sm100, tX, tDU, tW1, tW2T, tW1T, tDZ, p));
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

// w [rows, cols] -> pad [rows, ld] and / or tr [cols, ldt] (either may be null)
int tokmix_prepare_impl(const void* w, int rows, int cols, void* pad, int ld, void* tr, int ldt, cudaStream_t st) {
  if (!w || rows <= 0 || cols <= 0 || (pad && ld < cols) || (tr && ldt < rows)) return fail(VMLP_EINVAL, "tokmix_prepare args");
  const long long n = (pad ? (long long)rows * ld : 0) + (tr ? (long long)cols * ldt : 0);
  if (n == 0) return VMLP_OK;
  long long blocks = (n + 255) / 256;
  if (blocks > 4096) blocks = 4096;
  tokmix_prepare_kernel<<<(int)blocks, 256, 0, st>>>((cbf)w, rows, cols, (bf)pad, ld, (bf)tr, ldt);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

}  // namespace

// ============================================================================================ C ABI
extern "C" {

int vmlp_abi_version(void) { return VMLP_ABI_VERSION; }
#ifndef VMLP_SRC_HASH
#define VMLP_SRC_HASH "unknown"
#endif
const char* vmlp_source_hash(void) { return VMLP_SRC_HASH; }
int vmlp_abi_struct_bytes(int32_t which) {
  switch (which) {
    case 0: return (int)sizeof(vmlp_operand);
    case 1: return (int)sizeof(vmlp_gemm_args);
    case 2: return (int)sizeof(vmlp_mixer_params);
    case 3: return (int)sizeof(vmlp_mixer_saved);
    case 4: return (int)sizeof(vmlp_hire_dims);
    case 5: return (int)sizeof(vmlp_optim_chunk);
    case 6: return (int)sizeof(vmlp_optim_hyper);
  }
  return -1;
}
const char* vmlp_last_error(void) { return g_err; }
int vmlp_device_check(void) {
  const DeviceInfo& dv = device_info();
  if (!dv.ok) return fail(VMLP_ELAUNCH, "no CUDA device");
  if (dv.cc_major != 10) return fail(VMLP_EARCH, "compute capability %d.%d is not sm_100", dv.cc_major, dv.cc_minor);
  return VMLP_OK;
}
int vmlp_sm_count(void) { return device_info().sms; }
int vmlp_debug_read(uint32_t* out, int32_t n_words) {
  const DeviceInfo& dv = device_info();
  if (!dv.dbg_host || !out || n_words <= 0) return 0;
  const int n = n_words < DBG_WORDS ? n_words : DBG_WORDS;
  for (int i = 0; i < n; ++i) out[i] = reinterpret_cast<volatile unsigned int*>(dv.dbg_host)[i];
  return n;
}
int64_t vmlp_launch_count(void) { return g_launches.load(); }

int vmlp_gemm_bf16(const vmlp_gemm_args* args, vmlp_stream_t stream) {
  if (!args) return fail(VMLP_EINVAL, "null args");
  return gemm_impl(*args, static_cast<cudaStream_t>(stream));
}

int vmlp_layernorm_fwd(const void* x, int64_t x_ld, const void* gamma, const void* beta, void* y, int64_t y_ld,
                       float* mean, float* rstd, int64_t rows, int32_t C, float eps, vmlp_stream_t stream) {
  if (!x || !y || !gamma || !beta || rows <= 0 || C <= 0 || (C % 8)) return fail(VMLP_EINVAL, "layernorm_fwd args");
  if (!aligned16(x) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta) || (x_ld % 8) || (y_ld % 8))
    return fail(VMLP_EALIGN, "layernorm_fwd alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (C <= 64)         // narrow rows: 4 (2) rows per warp
    layernorm_fwd_kernel<1, 8><<<rw_grid((rows + 3) / 4), RW_THREADS, 0, st>>>((cbf)x, x_ld, (cbf)gamma, (cbf)beta, (bf)y, y_ld, mean, rstd, rows, C, eps);
  else if (C <= 128)
    layernorm_fwd_kernel<1, 16><<<rw_grid((rows + 1) / 2), RW_THREADS, 0, st>>>((cbf)x, x_ld, (cbf)gamma, (cbf)beta, (bf)y, y_ld, mean, rstd, rows, C, eps);
  else
  DISPATCH_VPL(C, (layernorm_fwd_kernel<VPL><<<rw_grid(rows), RW_THREADS, 0, st>>>(
                      (cbf)x, x_ld, (cbf)gamma, (cbf)beta, (bf)y, y_ld, mean, rstd, rows, C, eps)));
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_layernorm_bwd(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, const float* mean,
                       const float* rstd, const void* gamma, const void* add, int64_t add_ld, void* dx,
                       int64_t dx_ld, float* dgamma, float* dbeta, int64_t rows, int32_t C, vmlp_stream_t stream) {
  return layernorm_bwd_impl(dy, dy_ld, x, x_ld, mean, rstd, gamma, add, add_ld, dx, dx_ld, dgamma, dbeta, rows, C,
                            nullptr, nullptr, 1, stream);
}

int vmlp_affine_fwd(const void* x, const void* alpha, const void* beta, void* y, int64_t rows, int32_t C,
                    vmlp_stream_t stream) {
  if (!x || !y || !alpha || !beta || rows <= 0 || (C % 8)) return fail(VMLP_EINVAL, "affine_fwd args");
  if (!aligned16(x) || !aligned16(y) || !aligned16(alpha) || !aligned16(beta)) return fail(VMLP_EALIGN, "affine_fwd alignment");
  const long long nvec = rows * (C / 8);
  long long blocks = (nvec + RW_THREADS - 1) / RW_THREADS;
  const long long cap = (long long)device_info().sms * 16;
  affine_fwd_kernel<<<(int)(blocks < cap ? blocks : cap), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (cbf)x, (cbf)alpha, (cbf)beta, (bf)y, nvec, C / 8);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_affine_bwd(const void* dy, const void* x, const void* alpha, const void* add, void* dx, float* dalpha,
                    float* dbeta, int64_t rows, int32_t C, vmlp_stream_t stream) {
  if (!dy || !x || !alpha || !dx || !dalpha || !dbeta || rows <= 0 || (C % 8)) return fail(VMLP_EINVAL, "affine_bwd args");
  if (!aligned16(dy) || !aligned16(x) || !aligned16(dx) || !aligned16(alpha) || (add && !aligned16(add)))
    return fail(VMLP_EALIGN, "affine_bwd alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DeviceInfo& dv = device_info();
  long long blocks = (rows + RW_WARPS - 1) / RW_WARPS;
  const int grid = (int)(blocks < dv.sms * 4 ? blocks : dv.sms * 4);
  DISPATCH_VPL(C, (affine_bwd_kernel<VPL><<<grid, RW_THREADS, VPL * 256 * sizeof(float), st>>>(
                      (cbf)dy, (cbf)x, (cbf)alpha, (cbf)add, (bf)dx, dalpha, dbeta, rows, C)));
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_colsum(const void* a, int64_t a_ld, const void* b, int64_t b_ld, float* out, int64_t rows, int32_t C,
                vmlp_stream_t stream) {
  if (!a || !out || rows <= 0 || (C % 8)) return fail(VMLP_EINVAL, "colsum args");
  if (!aligned16(a) || (a_ld % 8) || (b && (!aligned16(b) || (b_ld % 8)))) return fail(VMLP_EALIGN, "colsum alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const DeviceInfo& dv = device_info();
  if (C <= 8 * RW_THREADS) {
    const int rpb = RW_THREADS / (C / 8);
    // >= 16 rows per thread, so that the block's setup and its C global atomics stay small next to its streaming work
    long long gx = (rows + 16LL * rpb - 1) / (16LL * rpb);
    const long long cap = (long long)dv.sms * 6;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    colsum_flat_kernel<0><<<(unsigned)gx, RW_THREADS, (size_t)rpb * C * sizeof(float), st>>>((cbf)a, a_ld, (cbf)b, b_ld, out,
                                                                                          nullptr, rows, C);
    CUDA_OK(cudaGetLastError());
    ++g_launches;
    return VMLP_OK;
  }
  const int slabs = (C + 255) / 256;
  long long gx = (rows + RW_WARPS - 1) / RW_WARPS;
  const long long cap = (long long)(dv.sms * 8 + slabs - 1) / slabs;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  colsum_kernel<<<dim3((unsigned)gx, (unsigned)slabs), RW_THREADS, 0, st>>>((cbf)a, a_ld, (cbf)b, b_ld, out, rows, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_colsum2(const void* a, const void* b, float* out_a, float* out_ab, int64_t rows, int32_t C,
                 vmlp_stream_t stream) {
  if (!a || !b || !out_a || !out_ab || rows <= 0 || (C % 8) || C > 8 * RW_THREADS) return fail(VMLP_EINVAL, "colsum2 args");
  if (!aligned16(a) || !aligned16(b)) return fail(VMLP_EALIGN, "colsum2 alignment");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int rpb = RW_THREADS / (C / 8);
  long long gx = (rows + 16LL * rpb - 1) / (16LL * rpb);
  const long long cap = (long long)device_info().sms * 6;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  colsum_flat_kernel<1><<<(unsigned)gx, RW_THREADS, 2 * (size_t)rpb * C * sizeof(float), st>>>((cbf)a, C, (cbf)b, C, out_a, out_ab,
                                                                                               rows, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_rowsum_batched(const void* a, float* out, int64_t batch, int32_t rows_per_batch, int32_t C,
                        vmlp_stream_t stream) {
  if (!a || !out || batch <= 0 || rows_per_batch <= 0 || (C % 8)) return fail(VMLP_EINVAL, "rowsum args");
  if (!aligned16(a)) return fail(VMLP_EALIGN, "rowsum alignment");
  const long long rows = batch * rows_per_batch;
  rowsum_batched_kernel<<<rw_grid(rows), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>((cbf)a, out, rows,
                                                                                            rows_per_batch, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_cast_f32_to_bf16(const float* src, void* dst, int64_t n, vmlp_stream_t stream) {
  if (!src || !dst || n <= 0) return fail(VMLP_EINVAL, "cast args");
  long long blocks = (n + 255) / 256;
  const long long cap = (long long)device_info().sms * 16;
  cast_f32_bf16_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>(src, (bf)dst, n);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_add_bf16(const void* a, const void* b, void* dst, int64_t n, vmlp_stream_t stream) {
  if (!a || !b || !dst || n <= 0 || (n % 8)) return fail(VMLP_EINVAL, "add args");
  if (!aligned16(a) || !aligned16(b) || !aligned16(dst)) return fail(VMLP_EALIGN, "add alignment");
  const long long nvec = n / 8;
  long long blocks = (nvec + 255) / 256;
  const long long cap = (long long)device_info().sms * 16;
  add_bf16_kernel<<<(int)(blocks < cap ? blocks : cap), 256, 0, static_cast<cudaStream_t>(stream)>>>((cbf)a, (cbf)b, (bf)dst, nvec);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_pad_rows(const void* src, void* dst, int32_t rows, int32_t cols, int32_t ld_dst, vmlp_stream_t stream) {
  if (!src || !dst || rows <= 0 || cols <= 0 || ld_dst < cols || (ld_dst % 8)) return fail(VMLP_EINVAL, "pad_rows args");
  const long long n = (long long)rows * ld_dst;
  pad_rows_kernel<<<(int)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>((cbf)src, (bf)dst, rows, cols, ld_dst);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
// grid for kernels that put the batch on blockIdx.y and walk a per-sample vector index with an int grid-stride loop
static dim3 sample_grid(long long per_sample_vec, int B) {
  long long gx = (per_sample_vec + RW_THREADS - 1) / RW_THREADS;
  const long long cap = ((long long)device_info().sms * 16 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)B);
}
static int ew_grid(long long total_vec) {
  long long blocks = (total_vec + RW_THREADS - 1) / RW_THREADS;
  const long long cap = (long long)device_info().sms * 16;
  return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}
int vmlp_mul_colvec(const void* a, int64_t a_ld, const void* v, void* out, int64_t out_ld, int64_t rows, int32_t C,
                    vmlp_stream_t stream) {
  if (!a || !v || !out || rows <= 0 || (C % 8)) return fail(VMLP_EINVAL, "mul_colvec args");
  if (!aligned16(a) || !aligned16(v) || !aligned16(out) || (a_ld % 8) || (out_ld % 8)) return fail(VMLP_EALIGN, "mul_colvec alignment");
  ew_kernel<0><<<ew_grid(rows * (C / 8)), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (cbf)a, a_ld, (cbf)v, 0, nullptr, 0, nullptr, 0, (bf)out, out_ld, nullptr, 0, rows, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_dgelu_mul(const void* a, int64_t a_ld, const void* z, int64_t z_ld, void* out, int64_t out_ld, int64_t rows,
                   int32_t C, vmlp_stream_t stream) {
  if (!a || !z || !out || rows <= 0 || (C % 8)) return fail(VMLP_EINVAL, "dgelu_mul args");
  if (!aligned16(a) || !aligned16(z) || !aligned16(out) || (a_ld % 8) || (z_ld % 8) || (out_ld % 8)) return fail(VMLP_EALIGN, "dgelu_mul alignment");
  ew_kernel<1><<<ew_grid(rows * (C / 8)), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (cbf)a, a_ld, (cbf)z, z_ld, nullptr, 0, nullptr, 0, (bf)out, out_ld, nullptr, 0, rows, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_gate_bwd(const void* dg, int64_t dg_ld, const void* vt, int64_t vt_ld, const void* zp_u, int64_t zp_ld,
                  const void* u, int64_t u_ld, void* out, int64_t out_ld, void* out2, int64_t out2_ld, int64_t rows,
                  int32_t C, vmlp_stream_t stream) {
  if (!dg || !vt || !zp_u || !u || !out || !out2 || rows <= 0 || (C % 8)) return fail(VMLP_EINVAL, "gate_bwd args");
  if (!aligned16(dg) || !aligned16(vt) || !aligned16(zp_u) || !aligned16(u) || !aligned16(out) || !aligned16(out2) ||
      (dg_ld % 8) || (vt_ld % 8) || (zp_ld % 8) || (u_ld % 8) || (out_ld % 8) || (out2_ld % 8))
    return fail(VMLP_EALIGN, "gate_bwd alignment");
  ew_kernel<2><<<ew_grid(rows * (C / 8)), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (cbf)dg, dg_ld, (cbf)vt, vt_ld, (cbf)zp_u, zp_ld, (cbf)u, u_ld, (bf)out, out_ld, (bf)out2, out2_ld, rows, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

// ============================================================================================ spatial operators
int vmlp_shift_nhwc(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t mode,
                    int32_t ngroups, const int32_t* start, const int32_t* dh, const int32_t* dw, vmlp_stream_t stream) {
  if (!in || !out || !start || !dh || !dw || B <= 0 || H <= 0 || W <= 0 || (C % 8) || ngroups < 1 || ngroups > 8 ||
      mode < 0 || mode > 2)
    return fail(VMLP_EINVAL, "shift_nhwc args");
  if (!aligned16(in) || !aligned16(out)) return fail(VMLP_EALIGN, "shift_nhwc alignment");
  ShiftTable t;
  memset(&t, 0, sizeof(t));
  t.ngroups = ngroups;
  for (int g = 0; g < ngroups; ++g) { t.start[g] = start[g]; t.dh[g] = dh[g]; t.dw[g] = dw[g]; }
  t.start[ngroups] = start[ngroups];
  if ((long long)H * W * (C / 8) >= (1 << 22) || B > 65535) return fail(VMLP_EINVAL, "shift_nhwc: sample too large");
  const dim3 grid = sample_grid((long long)H * W * (C / 8), B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (mode == 0) shift_nhwc_kernel<0><<<grid, RW_THREADS, 0, st>>>((cbf)in, (bf)out, B, H, W, C, t);
  else if (mode == 1) shift_nhwc_kernel<1><<<grid, RW_THREADS, 0, st>>>((cbf)in, (bf)out, B, H, W, C, t);
  else shift_nhwc_kernel<2><<<grid, RW_THREADS, 0, st>>>((cbf)in, (bf)out, B, H, W, C, t);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

static dim3 gn_grid(int B, long long per_sample_vec) {
  long long gx = (per_sample_vec + RW_THREADS * 4 - 1) / (RW_THREADS * 4);
  const long long cap = ((long long)device_info().sms * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)B);
}
int vmlp_gn_stats(const void* x, float* acc, int32_t B, int64_t P, int32_t C, vmlp_stream_t stream) {
  if (!x || !acc || B <= 0 || P <= 0 || (C % 8)) return fail(VMLP_EINVAL, "gn_stats args");
  if (!aligned16(x)) return fail(VMLP_EALIGN, "gn_stats alignment");
  const long long psv = P * (C / 8);
  gn_stats_kernel<<<gn_grid(B, psv), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>((cbf)x, acc, psv);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_gn_apply(const void* x, const float* acc, const void* gamma, const void* beta, void* y, int32_t B, int64_t P,
                  int32_t C, float eps, int32_t gelu, vmlp_stream_t stream) {
  if (!x || !acc || !gamma || !beta || !y || B <= 0 || P <= 0 || (C % 8)) return fail(VMLP_EINVAL, "gn_apply args");
  if (!aligned16(x) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta)) return fail(VMLP_EALIGN, "gn_apply alignment");
  const long long psv = P * (C / 8), total = psv * B;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (psv >= (1 << 22) || B > 65535) return fail(VMLP_EINVAL, "gn_apply: sample too large");
  if (gelu) gn_apply_kernel<1><<<sample_grid(psv, B), RW_THREADS, 0, st>>>((cbf)x, acc, (cbf)gamma, (cbf)beta, (bf)y, psv, C, eps, total);
  else gn_apply_kernel<0><<<sample_grid(psv, B), RW_THREADS, 0, st>>>((cbf)x, acc, (cbf)gamma, (cbf)beta, (bf)y, psv, C, eps, total);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_gn_bwd(const void* dy, const void* x, const float* acc, const void* gamma, const void* beta, void* dn,
                float* acc2, float* dgamma, float* dbeta, void* dx, int32_t B, int64_t P, int32_t C, float eps,
                int32_t gelu, vmlp_stream_t stream) {
  if (!dy || !x || !acc || !gamma || !beta || !dn || !acc2 || !dgamma || !dbeta || !dx || B <= 0 || P <= 0 || (C % 8))
    return fail(VMLP_EINVAL, "gn_bwd args");
  if (!aligned16(dy) || !aligned16(x) || !aligned16(dn) || !aligned16(dx) || !aligned16(gamma) || !aligned16(beta))
    return fail(VMLP_EALIGN, "gn_bwd alignment");
  const long long psv = P * (C / 8), total = psv * B;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t sh = 2 * (size_t)C * sizeof(float);
  if (C / 8 > RW_THREADS) return fail(VMLP_EINVAL, "gn_bwd: C=%d exceeds %d channels", C, 8 * RW_THREADS);
  if (gelu) gn_bwd_reduce_kernel<1><<<gn_grid(B, psv), RW_THREADS, sh, st>>>((cbf)dy, (cbf)x, acc, (cbf)gamma, (cbf)beta, (bf)dn, acc2, dgamma, dbeta, psv, C, eps);
  else gn_bwd_reduce_kernel<0><<<gn_grid(B, psv), RW_THREADS, sh, st>>>((cbf)dy, (cbf)x, acc, (cbf)gamma, (cbf)beta, (bf)dn, acc2, dgamma, dbeta, psv, C, eps);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  if (psv >= (1 << 22) || B > 65535) return fail(VMLP_EINVAL, "gn_bwd: sample too large");
  gn_bwd_apply_kernel<<<sample_grid(psv, B), RW_THREADS, 0, st>>>((cbf)dn, (cbf)x, acc, acc2, (cbf)gamma, (bf)dx, psv, C, eps, total);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_chan_lin(const void* p, const void* q, const void* z, const float* A, const float* Bq, const float* Cc,
                  void* out, int64_t rows, int32_t C, vmlp_stream_t stream) {
  if (!p || !A || !Cc || !out || rows <= 0 || (C % 8) || (q && !Bq)) return fail(VMLP_EINVAL, "chan_lin args");
  if (!aligned16(p) || !aligned16(out) || (q && !aligned16(q)) || (z && !aligned16(z))) return fail(VMLP_EALIGN, "chan_lin alignment");
  const long long total = rows * (C / 8);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // grid * 256 must be a multiple of C/8 (see the kernel): round up to a multiple of nvec / gcd(nvec, 256)
  int grid = ew_grid(total);
  {
    const int nvec = C / 8;
    int g = nvec, r = RW_THREADS;
    while (r) { const int t = g % r; g = r; r = t; }
    const int step = nvec / g;
    grid = ((grid + step - 1) / step) * step;
  }
  if (q && z) chan_lin_kernel<1, 1><<<grid, RW_THREADS, 0, st>>>((cbf)p, (cbf)q, (cbf)z, A, Bq, Cc, (bf)out, total, C);
  else if (q) chan_lin_kernel<1, 0><<<grid, RW_THREADS, 0, st>>>((cbf)p, (cbf)q, (cbf)z, A, Bq, Cc, (bf)out, total, C);
  else if (z) chan_lin_kernel<0, 1><<<grid, RW_THREADS, 0, st>>>((cbf)p, (cbf)q, (cbf)z, A, Bq, Cc, (bf)out, total, C);
  else chan_lin_kernel<0, 0><<<grid, RW_THREADS, 0, st>>>((cbf)p, (cbf)q, (cbf)z, A, Bq, Cc, (bf)out, total, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_bn_fwd_coef(const float* s1, const float* s2, const void* gamma, const void* beta, float* A, float* Cc,
                     float* mean, float* rstd, float* running_mean, float* running_var, int64_t R, float eps,
                     float momentum, int32_t C, vmlp_stream_t stream) {
  if (!s1 || !s2 || !gamma || !beta || !A || !Cc || !mean || !rstd || R <= 0 || C <= 0) return fail(VMLP_EINVAL, "bn_fwd_coef args");
  bn_fwd_coef_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      s1, s2, (cbf)gamma, (cbf)beta, A, Cc, mean, rstd, running_mean, running_var, (float)R, eps, momentum, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_bn_bwd_coef(const float* sdy, const float* sdya, const void* gamma, const float* mean, const float* rstd,
                     float* A, float* Bq, float* Cc, float* dgamma, float* dbeta, int64_t R, int32_t C,
                     vmlp_stream_t stream) {
  if (!sdy || !sdya || !gamma || !mean || !rstd || !A || !Bq || !Cc || !dgamma || !dbeta || R <= 0 || C <= 0)
    return fail(VMLP_EINVAL, "bn_bwd_coef args");
  bn_bwd_coef_kernel<<<(C + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(
      sdy, sdya, (cbf)gamma, mean, rstd, A, Bq, Cc, dgamma, dbeta, (float)R, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

static int s2v2_check(const void* a, const void* b, int B, int H, int W, int C) {
  if (!a || !b || B <= 0 || H <= 0 || W <= 0 || (C % 8) || C > 2048) return fail(VMLP_EINVAL, "s2v2 args");
  if (!aligned16(a) || !aligned16(b)) return fail(VMLP_EALIGN, "s2v2 alignment");
  return VMLP_OK;
}
static dim3 s2v2_reduce_grid(int B, int H, int W, int C) {
  const int plane = RW_THREADS / (C / 8);
  long long gx = ((long long)H * W + plane - 1) / plane;
  const long long cap = ((long long)device_info().sms * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  return dim3((unsigned)gx, (unsigned)B);
}
int vmlp_s2v2_sum(const void* t, float* a_f32, int32_t B, int32_t H, int32_t W, int32_t C, int32_t plain, vmlp_stream_t stream) {
  int rc = s2v2_check(t, a_f32, B, H, W, C);
  if (rc) return rc;
  s2v2_reduce_kernel<0><<<s2v2_reduce_grid(B, H, W, C), RW_THREADS, RW_THREADS * 8 * sizeof(float),
                          static_cast<cudaStream_t>(stream)>>>((cbf)t, nullptr, a_f32, H, W, C, plain);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_s2v2_combine(const void* t, const void* hat, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                      int32_t plain, vmlp_stream_t stream) {
  int rc = s2v2_check(t, out, B, H, W, C);
  if (rc) return rc;
  if (!hat || !aligned16(hat)) return fail(VMLP_EALIGN, "s2v2 hat");
  s2v2_combine_kernel<<<s2v2_reduce_grid(B, H, W, C), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (cbf)t, (cbf)hat, (bf)out, B, H, W, C, plain);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_s2v2_combine_bwd(const void* t, const void* hat, const void* dout, float* dbar_f32, void* dhat, void* dt,
                          int32_t B, int32_t H, int32_t W, int32_t C, int32_t plain, vmlp_stream_t stream) {
  int rc = s2v2_check(t, dhat, B, H, W, C);
  if (rc) return rc;
  if (!hat || !dout || !dbar_f32 || !dhat || !aligned16(hat) || !aligned16(dout) || !aligned16(dhat) || (dt && !aligned16(dt)))
    return fail(VMLP_EALIGN, "s2v2 bwd");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  s2v2_reduce_kernel<1><<<s2v2_reduce_grid(B, H, W, C), RW_THREADS, RW_THREADS * 24 * sizeof(float), st>>>(
      (cbf)t, (cbf)dout, dbar_f32, H, W, C, plain);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  const long long nv = (long long)B * (C / 8);
  s2v2_softmax_bwd_kernel<<<(int)((nv + 127) / 128), 128, 0, st>>>((cbf)hat, dbar_f32, (bf)dhat, B, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  if (dt) {     // dt == NULL: the caller finishes with vmlp_s2v2_dt_fused once d(a) is known
    s2v2_dt_kernel<0><<<s2v2_reduce_grid(B, H, W, C), RW_THREADS, 0, st>>>((cbf)dout, (cbf)hat, nullptr, (bf)dt, B, H, W, C, plain);
    CUDA_OK(cudaGetLastError());
    ++g_launches;
  }
  return VMLP_OK;
}
int vmlp_s2v2_sum_bwd(const void* da, void* dt, int32_t B, int32_t H, int32_t W, int32_t C, int32_t plain, vmlp_stream_t stream) {
  int rc = s2v2_check(da, dt, B, H, W, C);
  if (rc) return rc;
  s2v2_dt_kernel<1><<<s2v2_reduce_grid(B, H, W, C), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      nullptr, nullptr, (cbf)da, (bf)dt, B, H, W, C, plain);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_s2v2_dt_fused(const void* dout, const void* hat, const void* da, void* dt, int32_t B, int32_t H, int32_t W,
                       int32_t C, int32_t plain, vmlp_stream_t stream) {
  int rc = s2v2_check(dout, dt, B, H, W, C);
  if (rc) return rc;
  if (!hat || !da || !aligned16(hat) || !aligned16(da)) return fail(VMLP_EALIGN, "s2v2 dt_fused");
  s2v2_dt_kernel<2><<<s2v2_reduce_grid(B, H, W, C), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (cbf)dout, (cbf)hat, (cbf)da, (bf)dt, B, H, W, C, plain);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_permute5(const void* in, void* out, const int32_t dims[5], const int64_t in_strides[4],
                  const int64_t out_strides[4], int32_t accumulate, vmlp_stream_t stream) {
  if (!in || !out || !dims || !in_strides || !out_strides) return fail(VMLP_EINVAL, "permute5 args");
  if (!aligned16(in) || !aligned16(out)) return fail(VMLP_EALIGN, "permute5 alignment");
  for (int i = 0; i < 5; ++i)
    if (dims[i] <= 0) return fail(VMLP_EINVAL, "permute5: dims[%d] = %d", i, dims[i]);
  if (dims[4] % 8) return fail(VMLP_EINVAL, "permute5: inner run %d is not a multiple of 8 elements", dims[4]);
  Permute5 p;
  p.n1 = dims[1]; p.n2 = dims[2]; p.n3 = dims[3]; p.nv = dims[4] / 8;
  const long long slab = (long long)p.n1 * p.n2 * p.n3 * p.nv;
  if (slab >= (1 << 22)) return fail(VMLP_EINVAL, "permute5: dims[1..4] too large (%lld vectors per dims[0] step)", slab);
  for (int i = 0; i < 4; ++i) {
    if (in_strides[i] % 8 || out_strides[i] % 8) return fail(VMLP_EALIGN, "permute5: strides must be multiples of 8");
    p.is[i] = in_strides[i]; p.os[i] = out_strides[i];
  }
  p.total = slab * dims[0];
  long long blocks = (p.total + 255) / 256;
  const long long cap = (long long)device_info().sms * 16;
  if (blocks > cap) blocks = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (accumulate) permute5_kernel<true><<<(unsigned)blocks, 256, 0, st>>>((cbf)in, (bf)out, p);
  else permute5_kernel<false><<<(unsigned)blocks, 256, 0, st>>>((cbf)in, (bf)out, p);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

int vmlp_token_mean(const void* x, void* out, int32_t B, int32_t P, int32_t C, vmlp_stream_t stream) {
  if (!x || !out || B <= 0 || P <= 0 || C <= 0 || (C % 8)) return fail(VMLP_EINVAL, "token_mean args");
  if (!aligned16(x) || !aligned16(out)) return fail(VMLP_EALIGN, "token_mean alignment");
  token_mean_kernel<<<dim3((C / 8 + 31) / 32, B), 256, 0, static_cast<cudaStream_t>(stream)>>>((cbf)x, (bf)out, P, C, 1.0f / (float)P);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_token_mean_bwd(const void* g, void* dx, int32_t B, int32_t P, int32_t C, vmlp_stream_t stream) {
  if (!g || !dx || B <= 0 || P <= 0 || C <= 0 || (C % 8)) return fail(VMLP_EINVAL, "token_mean_bwd args");
  if (!aligned16(g) || !aligned16(dx)) return fail(VMLP_EALIGN, "token_mean_bwd alignment");
  long long gx = ((long long)P * (C / 8) + 255) / 256;
  const long long cap = ((long long)device_info().sms * 8 + B - 1) / B;
  if (gx > cap) gx = cap;
  if (gx < 1) gx = 1;
  token_mean_bwd_kernel<<<dim3((unsigned)gx, B), 256, 0, static_cast<cudaStream_t>(stream)>>>((cbf)g, (bf)dx, P, C, 1.0f / (float)P);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

static_assert(sizeof(vmlp_optim_chunk) == sizeof(OptimChunk), "optimizer table entry layout");
static_assert(sizeof(vmlp_optim_hyper) == sizeof(OptimHyper), "optimizer hyper-parameter layout");
int vmlp_optim_step(const vmlp_optim_chunk* table_dev, int32_t n_chunks, float* master, float* mom, float* var,
                    const vmlp_optim_hyper* hyper, vmlp_stream_t stream) {
  if (!table_dev || n_chunks <= 0 || !master || !mom || !hyper) return fail(VMLP_EINVAL, "optim_step args");
  if (hyper->kind != 0 && hyper->kind != 1) return fail(VMLP_EINVAL, "optim_step: kind %d", hyper->kind);
  if (hyper->kind == 0 && !var) return fail(VMLP_EINVAL, "optim_step: AdamW needs the second-moment buffer");
  if (!aligned16(master) || !aligned16(mom) || (var && !aligned16(var))) return fail(VMLP_EALIGN, "optim_step state");
  OptimHyper h;
  memcpy(&h, hyper, sizeof(h));
  optim_step_kernel<<<n_chunks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const OptimChunk*>(table_dev), master, mom, var, h);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

static int hire_dims(const vmlp_hire_dims* d, HireDims& o) {
  if (!d || d->B <= 0 || d->H <= 0 || d->W <= 0 || (d->C % 8) || d->h <= 0 || d->w <= 0) return fail(VMLP_EINVAL, "hire dims");
  o.B = d->B; o.H = d->H; o.W = d->W; o.C = d->C;
  o.nh = d->h; o.Hp = d->H + (d->h - d->H % d->h); o.Gh = o.Hp / d->h; o.step_h = d->step_h;
  o.nw = d->w; o.Wp = d->W + (d->w - d->W % d->w); o.Gw = o.Wp / d->w; o.step_w = d->step_w;
  if ((long long)o.Hp * o.Wp * (o.C / 8) * 2 >= (1 << 22) || o.B > 65535) return fail(VMLP_EINVAL, "hire: sample too large");
  return VMLP_OK;
}
int vmlp_hire_build(const void* x, void* zh, void* zw, const vmlp_hire_dims* d, vmlp_stream_t stream) {
  HireDims hd;
  int rc = hire_dims(d, hd);
  if (rc) return rc;
  if (!x || !zh || !zw || !aligned16(x) || !aligned16(zh) || !aligned16(zw)) return fail(VMLP_EALIGN, "hire_build pointers");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nv = hd.C / 8;
  hire_build_kernel<0><<<sample_grid((long long)hd.Gh * hd.W * hd.nh * nv, hd.B), RW_THREADS, 0, st>>>((cbf)x, (bf)zh, hd);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  hire_build_kernel<1><<<sample_grid((long long)hd.H * hd.Gw * hd.nw * nv, hd.B), RW_THREADS, 0, st>>>((cbf)x, (bf)zw, hd);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_hire_build_adj(const void* dzh, const void* dzw, void* dx, const vmlp_hire_dims* d, vmlp_stream_t stream) {
  HireDims hd;
  int rc = hire_dims(d, hd);
  if (rc) return rc;
  if (!dzh || !dzw || !dx || !aligned16(dzh) || !aligned16(dzw) || !aligned16(dx)) return fail(VMLP_EALIGN, "hire_build_adj pointers");
  hire_build_adj_kernel<<<sample_grid((long long)hd.H * hd.W * (hd.C / 8), hd.B), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (cbf)dzh, (cbf)dzw, (bf)dx, hd);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_hire_combine(const void* base, const void* oh, const void* ow, void* out, const vmlp_hire_dims* d,
                      vmlp_stream_t stream) {
  HireDims hd;
  int rc = hire_dims(d, hd);
  if (rc) return rc;
  if (!base || !oh || !ow || !out || !aligned16(base) || !aligned16(oh) || !aligned16(ow) || !aligned16(out))
    return fail(VMLP_EALIGN, "hire_combine pointers");
  hire_combine_kernel<<<sample_grid((long long)hd.H * hd.W * (hd.C / 8), hd.B), RW_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      (cbf)base, (cbf)oh, (cbf)ow, (bf)out, hd);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
int vmlp_hire_restore_adj(const void* dout, void* dzh, void* dzw, const vmlp_hire_dims* d, vmlp_stream_t stream) {
  HireDims hd;
  int rc = hire_dims(d, hd);
  if (rc) return rc;
  if (!dout || !dzh || !dzw || !aligned16(dout) || !aligned16(dzh) || !aligned16(dzw)) return fail(VMLP_EALIGN, "hire_restore_adj pointers");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long nv = hd.C / 8;
  hire_restore_adj_kernel<0><<<sample_grid((long long)hd.Gh * hd.W * hd.nh * nv, hd.B), RW_THREADS, 0, st>>>((cbf)dout, (bf)dzh, hd);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  hire_restore_adj_kernel<1><<<sample_grid((long long)hd.H * hd.Gw * hd.nw * nv, hd.B), RW_THREADS, 0, st>>>((cbf)dout, (bf)dzw, hd);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

}  // extern "C" (templates need C++ linkage)
template <int K, int FLIP, int EPI>
static int dwconv_launch(const void* x, const void* w, const void* bias, void* o1, void* o2, int B, int H, int W, int C,
                         cudaStream_t st) {
  auto kern = dwconv_kernel<K, FLIP, EPI>;
  static std::atomic<int> optin[64];
  if (int rc0 = smem_optin(kern, DwSmem<K>::BYTES, optin)) return rc0;
  const long long tiles = (long long)B * ((H + DW_TH - 1) / DW_TH) * ((W + DW_TW - 1) / DW_TW);
  if (tiles >= (1 << 22)) return fail(VMLP_EINVAL, "dwconv: too many tiles");
  const int cb = (C + DW_CH - 1) / DW_CH;
  // persistent blocks: the grid must not exceed the resident capacity (asked from the runtime: registers, shared memory
  // and the L1 carve-out decide), or the few blocks of a second wave double the kernel time
  static std::atomic<int> occ_cache[64];
  int occ = 1;
  if (int rc0 = occupancy_cached(kern, DW_THREADS, DwSmem<K>::BYTES, occ_cache, &occ)) return rc0;
  if (env_knobs().debug) fprintf(stderr, "vmlp: dwconv<%d,%d,%d> %d blocks/SM\n", K, FLIP, EPI, occ);
  long long gx = ((long long)device_info().sms * occ) / cb;
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  CUtensorMap tmx;
  int rc = make_map_nhwc(&tmx, x, B, H, W, C, DW_CH, DwSmem<K>::IN_W, DwSmem<K>::IN_H);
  if (rc) return rc;
  kern<<<dim3((unsigned)gx, (unsigned)cb), DW_THREADS, DwSmem<K>::BYTES, st>>>(tmx, (cbf)w, (cbf)bias, (bf)o1, (bf)o2, B, H, W, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
template <int K>
static int dwconv_wgrad_launch(const void* x, const void* dz, float* dw, int B, int H, int W, int C, cudaStream_t st) {
  auto kern = dwconv_wgrad_kernel<K>;
  static std::atomic<int> optin[64];
  if (int rc0 = smem_optin(kern, DwSmem<K>::BYTES_WGRAD, optin)) return rc0;
  const long long tiles = (long long)B * ((H + DW_TH - 1) / DW_TH) * ((W + DW_TW - 1) / DW_TW);
  if (tiles >= (1 << 22)) return fail(VMLP_EINVAL, "dwconv: too many tiles");
  const int cb = (C + DW_CH - 1) / DW_CH;
  static std::atomic<int> occ_cache[64];
  int occ = 1;
  if (int rc0 = occupancy_cached(kern, 32 * K, DwSmem<K>::BYTES_WGRAD, occ_cache, &occ)) return rc0;
  if (env_knobs().debug) fprintf(stderr, "vmlp: dwconv_wgrad<%d> %d blocks/SM\n", K, occ);
  long long gx = ((long long)device_info().sms * occ) / cb;
  if (gx < 1) gx = 1;
  if (gx > tiles) gx = tiles;
  CUtensorMap tmx, tmdz;
  int rc = make_map_nhwc(&tmx, x, B, H, W, C, DW_CH, DwSmem<K>::IN_W, DwSmem<K>::IN_H);
  if (rc) return rc;
  rc = make_map_nhwc(&tmdz, dz, B, H, W, C, DW_CH, DW_TW, DW_TH);
  if (rc) return rc;
  kern<<<dim3((unsigned)gx, (unsigned)cb), 32 * K, DwSmem<K>::BYTES_WGRAD, st>>>(tmx, tmdz, dw, B, H, W, C);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}
extern "C" {
static int dwconv_check(const void* a, const void* b, const void* c, int B, int H, int W, int C, int K) {
  if (!a || !b || !c || B <= 0 || H <= 0 || W <= 0 || (C % 8)) return fail(VMLP_EINVAL, "dwconv args");
  if (K != 3 && K != 5 && K != 7 && K != 9) return fail(VMLP_EINVAL, "dwconv kernel_size %d not in {3,5,7,9}", K);
  if (!aligned16(a) || !aligned16(c)) return fail(VMLP_EALIGN, "dwconv alignment");
  return VMLP_OK;
}
#define DW_DISPATCH(K, CALL3, CALL5, CALL7, CALL9) ((K) == 3 ? (CALL3) : (K) == 5 ? (CALL5) : (K) == 7 ? (CALL7) : (CALL9))
int vmlp_dwconv_fwd(const void* x, const void* weight, const void* bias, void* z, void* a, int32_t B, int32_t H,
                    int32_t W, int32_t C, int32_t K, vmlp_stream_t stream) {
  int rc = dwconv_check(x, weight, z, B, H, W, C, K);
  if (rc) return rc;
  if (!bias || !a) return fail(VMLP_EINVAL, "dwconv_fwd needs bias and both outputs");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return DW_DISPATCH(K, (dwconv_launch<3, 0, 1>(x, weight, bias, z, a, B, H, W, C, st)),
                     (dwconv_launch<5, 0, 1>(x, weight, bias, z, a, B, H, W, C, st)),
                     (dwconv_launch<7, 0, 1>(x, weight, bias, z, a, B, H, W, C, st)),
                     (dwconv_launch<9, 0, 1>(x, weight, bias, z, a, B, H, W, C, st)));
}
int vmlp_dwconv_fwd_plain(const void* x, const void* weight, const void* bias, void* y, int32_t B, int32_t H, int32_t W,
                          int32_t C, int32_t K, vmlp_stream_t stream) {
  int rc = dwconv_check(x, weight, y, B, H, W, C, K);
  if (rc) return rc;
  if (!bias) return fail(VMLP_EINVAL, "dwconv_fwd_plain needs a bias");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return DW_DISPATCH(K, (dwconv_launch<3, 0, 2>(x, weight, bias, y, nullptr, B, H, W, C, st)),
                     (dwconv_launch<5, 0, 2>(x, weight, bias, y, nullptr, B, H, W, C, st)),
                     (dwconv_launch<7, 0, 2>(x, weight, bias, y, nullptr, B, H, W, C, st)),
                     (dwconv_launch<9, 0, 2>(x, weight, bias, y, nullptr, B, H, W, C, st)));
}
int vmlp_dwconv_dgrad(const void* dz, const void* weight, void* dx, int32_t B, int32_t H, int32_t W, int32_t C,
                      int32_t K, vmlp_stream_t stream) {
  int rc = dwconv_check(dz, weight, dx, B, H, W, C, K);
  if (rc) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return DW_DISPATCH(K, (dwconv_launch<3, 1, 0>(dz, weight, nullptr, dx, nullptr, B, H, W, C, st)),
                     (dwconv_launch<5, 1, 0>(dz, weight, nullptr, dx, nullptr, B, H, W, C, st)),
                     (dwconv_launch<7, 1, 0>(dz, weight, nullptr, dx, nullptr, B, H, W, C, st)),
                     (dwconv_launch<9, 1, 0>(dz, weight, nullptr, dx, nullptr, B, H, W, C, st)));
}
int vmlp_dwconv_wgrad(const void* x, const void* dz, float* dw, int32_t B, int32_t H, int32_t W, int32_t C, int32_t K,
                      vmlp_stream_t stream) {
  int rc = dwconv_check(x, dz, dz, B, H, W, C, K);
  if (rc) return rc;
  if (!dw) return fail(VMLP_EINVAL, "dwconv_wgrad null dw");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  return DW_DISPATCH(K, (dwconv_wgrad_launch<3>(x, dz, dw, B, H, W, C, st)), (dwconv_wgrad_launch<5>(x, dz, dw, B, H, W, C, st)),
                     (dwconv_wgrad_launch<7>(x, dz, dw, B, H, W, C, st)), (dwconv_wgrad_launch<9>(x, dz, dw, B, H, W, C, st)));
}

int vmlp_patchify(const void* src, void* dst, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t P, int32_t forward,
                  vmlp_stream_t stream) {
  if (!src || !dst || B <= 0 || Cin <= 0 || P <= 0 || (P % 8) || (H % P) || (W % P)) return fail(VMLP_EINVAL, "patchify args");
  if (!aligned16(src) || !aligned16(dst)) return fail(VMLP_EALIGN, "patchify alignment");
  const long long per = (long long)Cin * H * W / 8;
  if (per >= (1 << 22) || B > 65535) return fail(VMLP_EINVAL, "patchify: sample too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (forward) patchify_kernel<1><<<sample_grid(per, B), RW_THREADS, 0, st>>>((cbf)src, (bf)dst, B, Cin, H, W, P);
  else patchify_kernel<0><<<sample_grid(per, B), RW_THREADS, 0, st>>>((cbf)src, (bf)dst, B, Cin, H, W, P);
  CUDA_OK(cudaGetLastError());
  ++g_launches;
  return VMLP_OK;
}

// ============================================================================================ fused token-mixing MLP
int vmlp_tokmix_set_trace(void* buf) { g_tokmix_trace = static_cast<long long*>(buf); return VMLP_OK; }
int vmlp_tokmix_supported(int32_t B, int32_t N, int32_t C, int32_t Ds, int32_t backward) {
  TokParams p;
  return tokmix_plan(p, B, N, C, Ds, backward != 0) ? 1 : 0;
}
int vmlp_tokmix_prepare(const void* w, int32_t rows, int32_t cols, void* pad, int32_t ld, void* tr, int32_t ldt,
                        vmlp_stream_t stream) {
  return tokmix_prepare_impl(w, rows, cols, pad, ld, tr, ldt, static_cast<cudaStream_t>(stream));
}
int vmlp_tokmix_fwd(const void* xhat, const void* x, const void* w1_pad, int32_t Np, const void* w2, const void* b1,
                    const void* b2, void* u, void* hT, int32_t B, int32_t N, int32_t C, int32_t Ds, vmlp_stream_t stream) {
  return tokmix_fwd_impl(xhat, x, w1_pad, Np, w2, b1, b2, u, hT, B, N, C, Ds, static_cast<cudaStream_t>(stream));
}
int vmlp_tokmix_bwd(const void* xhat, const void* du, const void* w1_pad, const void* w2T_pad, int32_t Np,
                    const void* w1T, const void* b1, void* dxhat, void* dzT, float* db1, int32_t B, int32_t N, int32_t C,
                    int32_t Ds, vmlp_stream_t stream) {
  return tokmix_bwd_impl(xhat, du, w1_pad, w2T_pad, Np, w1T, b1, dxhat, dzT, db1, B, N, C, Ds,
                         static_cast<cudaStream_t>(stream));
}

// ============================================================================================ MLP-Mixer block
static int mixer_check(const vmlp_mixer_params* p) {
  if (!p) return fail(VMLP_EINVAL, "null params");
  if (p->B <= 0 || p->N <= 0 || p->C <= 0 || p->Ds <= 0 || p->Dc <= 0) return fail(VMLP_EINVAL, "mixer dims");
  if ((p->C % 8) || (p->Ds % 8) || (p->Dc % 8)) return fail(VMLP_EINVAL, "mixer: C, Ds, Dc must be multiples of 8");
  return VMLP_OK;
}
static inline int pad8(int n) { return (n + 7) & ~7; }
static inline int pad16(int n) { return (n + 15) & ~15; }   // pitch of the zero-padded token weights (k-steps of 16)
// The token-mixing half runs as the fused on-chip kernels (tokmix_sm100.cuh) whenever both directions support the
// shape; VMLP_TOKMIX=0 (read once) keeps the unfused GEMM sequence for A/B measurements.
static bool mixer_token_fused(const vmlp_mixer_params* p) {
  if (!env_knobs().tokmix) return false;
  TokParams t;
  return tokmix_plan(t, p->B, p->N, p->C, p->Ds, false) && tokmix_plan(t, p->B, p->N, p->C, p->Ds, true);
}
int vmlp_mixer_token_fused(const vmlp_mixer_params* p) { return (p && mixer_check(p) == VMLP_OK && mixer_token_fused(p)) ? 1 : 0; }

int vmlp_mixer_block_fwd(const vmlp_mixer_params* p, const void* x, void* y, const vmlp_mixer_saved* s,
                         vmlp_stream_t stream) {
  int rc = mixer_check(p);
  if (rc) return rc;
  if (!x || !y || !s) return fail(VMLP_EINVAL, "mixer fwd null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = p->B, N = p->N, C = p->C, Ds = p->Ds, Dc = p->Dc;
  const long long R = (long long)B * N;
  const int Np = pad16(N);
  float* mean1 = s->stats; float* rstd1 = mean1 + R; float* mean2 = rstd1 + R; float* rstd2 = mean2 + R;

  // ---- token mixing: u = x + W2t * gelu(W1t * LN1(x) + b1t) + b2t  (contraction over tokens, per image)
  rc = vmlp_layernorm_fwd(x, C, p->ln1_w, p->ln1_b, s->xhat1, C, mean1, rstd1, R, C, p->eps, stream);
  if (rc) return rc;
  {
    const long long n = (long long)Ds * Np;
    pad_rows_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>((cbf)p->w1t, (bf)s->w1t_pad, Ds, N, Np);
    CUDA_OK(cudaGetLastError());
    ++g_launches;
  }
  if (mixer_token_fused(p)) {
    // one kernel: U = X + W2t gelu(W1t Xhat1 + b1t) + b2t; the hidden tensor stays on chip, H^T [B, C, Ds] is saved
    if (!s->h1) return fail(VMLP_EINVAL, "mixer fwd: h1 buffer missing");
    rc = tokmix_fwd_impl(s->xhat1, x, s->w1t_pad, Np, p->w2t, p->b1t, p->b2t, s->u, s->h1, B, N, C, Ds, st);
    if (rc) return rc;
  } else {
    {  // Z1[b] [Ds, C] = W1t [Ds, N] * Xhat1[b] [N, C] ; H1 = gelu(Z1)
      vmlp_gemm_args g = gemm_args(Ds, C, N, B, opnd(s->w1t_pad, Ds, N, Np, 0, 0),
                                   opnd(s->xhat1, N, C, C, (long long)N * C, 1), VMLP_EPI_GELU);
      g.D = s->z1; g.d_ld = C; g.d_bs = (long long)Ds * C;
      g.D2 = s->h1; g.d2_ld = C; g.d2_bs = (long long)Ds * C;
      g.bias = p->b1t; g.bias_mode = 2;
      rc = gemm_impl(g, st);
      if (rc) return rc;
    }
    {  // U[b] [N, C] = W2t [N, Ds] * H1[b] [Ds, C] + b2t[n] + X[b]
      vmlp_gemm_args g = gemm_args(N, C, Ds, B, opnd(p->w2t, N, Ds, Ds, 0, 0),
                                   opnd(s->h1, Ds, C, C, (long long)Ds * C, 1), VMLP_EPI_RESID);
      g.D = s->u; g.d_ld = C; g.d_bs = (long long)N * C;
      g.bias = p->b2t; g.bias_mode = 2;
      g.aux = x; g.aux_ld = C; g.aux_bs = (long long)N * C;
      rc = gemm_impl(g, st);
      if (rc) return rc;
    }
  }
  // ---- channel mixing: y = u + gelu(LN2(u) W1c^T + b1c) W2c^T + b2c   (rows = B*N tokens)
  rc = vmlp_layernorm_fwd(s->u, C, p->ln2_w, p->ln2_b, s->xhat2, C, mean2, rstd2, R, C, p->eps, stream);
  if (rc) return rc;
  {
    vmlp_gemm_args g = gemm_args((int)R, Dc, C, 1, opnd(s->xhat2, R, C, C, 0, 0), opnd(p->w1c, Dc, C, C, 0, 0),
                                 VMLP_EPI_GELU);
    g.D = s->z2; g.d_ld = Dc; g.D2 = s->h2; g.d2_ld = Dc;
    g.bias = p->b1c; g.bias_mode = 1;
    rc = gemm_impl(g, st);
    if (rc) return rc;
  }
  {
    vmlp_gemm_args g = gemm_args((int)R, C, Dc, 1, opnd(s->h2, R, Dc, Dc, 0, 0), opnd(p->w2c, C, Dc, Dc, 0, 0),
                                 VMLP_EPI_RESID);
    g.D = y; g.d_ld = C;
    g.bias = p->b2c; g.bias_mode = 1;
    g.aux = s->u; g.aux_ld = C;
    rc = gemm_impl(g, st);
    if (rc) return rc;
  }
  return VMLP_OK;
}

int64_t vmlp_mixer_grad_elems(const vmlp_mixer_params* p) {
  if (!p) return 0;
  const int64_t N = p->N, C = p->C, Ds = p->Ds, Dc = p->Dc;
  return 2 * C + Ds * N + Ds + N * Ds + N + 2 * C + Dc * C + Dc + C * Dc + C;
}
int64_t vmlp_mixer_bwd_workspace_elems(const vmlp_mixer_params* p) {
  if (!p) return 0;
  const int64_t R = (int64_t)p->B * p->N, C = p->C;
  const int64_t tok = (int64_t)p->Ds * C, chn = (int64_t)p->N * p->Dc;
  const int64_t hid = (int64_t)p->B * (tok > chn ? tok : chn);
  // dZ (max of both halves) + dXhat + dU + the transposed token weights of the fused backward (W2t^T [Ds, Np], W1t^T [N, Ds])
  return hid + 2 * R * C + (int64_t)p->Ds * pad16(p->N) + (int64_t)pad16(p->N) * p->Ds;
}

int vmlp_mixer_block_bwd(const vmlp_mixer_params* p, const void* x, const void* dy, void* dx,
                         const vmlp_mixer_saved* s, float* grads, void* workspace, int64_t workspace_elems,
                         vmlp_stream_t stream) {
  int rc = mixer_check(p);
  if (rc) return rc;
  if (!x || !dy || !dx || !s || !grads || !workspace) return fail(VMLP_EINVAL, "mixer bwd null pointer");
  if (workspace_elems < vmlp_mixer_bwd_workspace_elems(p)) return fail(VMLP_EWORKSPACE, "mixer bwd workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int B = p->B, N = p->N, C = p->C, Ds = p->Ds, Dc = p->Dc;
  const long long R = (long long)B * N;
  const int Np = pad16(N);
  float* mean1 = s->stats; float* rstd1 = mean1 + R; float* mean2 = rstd1 + R; float* rstd2 = mean2 + R;
  // fp32 gradient accumulators, in vmlp_mixer_params order
  float* g_ln1w = grads;            float* g_ln1b = g_ln1w + C;
  float* g_w1t = g_ln1b + C;        float* g_b1t = g_w1t + (long long)Ds * N;
  float* g_w2t = g_b1t + Ds;        float* g_b2t = g_w2t + (long long)N * Ds;
  float* g_ln2w = g_b2t + N;        float* g_ln2b = g_ln2w + C;
  float* g_w1c = g_ln2b + C;        float* g_b1c = g_w1c + (long long)Dc * C;
  float* g_w2c = g_b1c + Dc;        float* g_b2c = g_w2c + (long long)C * Dc;
  // bf16 workspace
  bf ws = (bf)workspace;
  const long long wtr = (long long)Ds * Np;
  const long long hid = vmlp_mixer_bwd_workspace_elems(p) - 2 * R * C - 2 * wtr;
  bf dZ = ws; bf dXh = ws + hid; bf dU = dXh + R * C; bf w2T_pad = dU + R * C; bf w1T = w2T_pad + wtr;

  // ================= channel half:  y = u + FF(LN2(u))
  {  // dZ2 = (dY * W2c) .* gelu'(Z2)            [R, Dc];  W2c [C, Dc] is the MN-major B operand
    vmlp_gemm_args g = gemm_args((int)R, Dc, C, 1, opnd(dy, R, C, C, 0, 0), opnd(p->w2c, C, Dc, Dc, 0, 1), VMLP_EPI_DGELU);
    g.D = dZ; g.d_ld = Dc; g.aux = s->z2; g.aux_ld = Dc;
    g.red_out = g_b1c; g.red_mode = 1;            // db1c = column sums of dZ2, fused into the epilogue
    if ((rc = gemm_impl(g, st))) return rc;
  }
  {  // dW2c [C, Dc] += dY^T * H2     (contraction over the R token rows: both operands MN-major)
    vmlp_gemm_args g = gemm_args(C, Dc, (int)R, 1, opnd(dy, R, C, C, 0, 1), opnd(s->h2, R, Dc, Dc, 0, 1), VMLP_EPI_ATOMIC);
    g.out_f32 = g_w2c; g.out_ld = Dc;
    if ((rc = gemm_impl(g, st))) return rc;
  }
  const bool fused_sums = C <= 1536;   // db2c = colsum(dY) and db2t = per-token sums of dU ride in the LN2 backward pass
  if (!fused_sums && (rc = vmlp_colsum(dy, C, nullptr, 0, g_b2c, R, C, stream))) return rc;
  {  // dXhat2 = dZ2 * W1c                        [R, C];  W1c [Dc, C] MN-major B
    vmlp_gemm_args g = gemm_args((int)R, C, Dc, 1, opnd(dZ, R, Dc, Dc, 0, 0), opnd(p->w1c, Dc, C, C, 0, 1), VMLP_EPI_STORE);
    g.D = dXh; g.d_ld = C;
    if ((rc = gemm_impl(g, st))) return rc;
  }
  {  // dW1c [Dc, C] += dZ2^T * Xhat2
    vmlp_gemm_args g = gemm_args(Dc, C, (int)R, 1, opnd(dZ, R, Dc, Dc, 0, 1), opnd(s->xhat2, R, C, C, 0, 1), VMLP_EPI_ATOMIC);
    g.out_f32 = g_w1c; g.out_ld = C;
    if ((rc = gemm_impl(g, st))) return rc;
  }
  // dU = dY + LN2'(dXhat2)
  if ((rc = layernorm_bwd_impl(dXh, C, s->u, C, mean2, rstd2, p->ln2_w, dy, C, dU, C, g_ln2w, g_ln2b, R, C,
                               fused_sums ? g_b2c : nullptr, fused_sums ? g_b2t : nullptr, N, stream))) return rc;

  // ================= token half:  u = x + TokenFF(LN1(x))      (the padded W1t copy of the forward pass is reused)
  if (mixer_token_fused(p)) {
    // K-major transposed weight copies, then ONE kernel for the data-gradient chain (Z recomputed on chip):
    // dXhat1 = W1t^T ((W2t^T dU) .* gelu'(W1t Xhat1 + b1t)), dZ^T [B, C, Ds] saved for dW1t, db1t accumulated on the way
    if ((rc = tokmix_prepare_impl(p->w2t, N, Ds, nullptr, 0, w2T_pad, Np, st))) return rc;
    if ((rc = tokmix_prepare_impl(p->w1t, Ds, N, nullptr, 0, w1T, Ds, st))) return rc;
    if ((rc = tokmix_bwd_impl(s->xhat1, dU, s->w1t_pad, w2T_pad, Np, w1T, p->b1t, dXh, dZ, g_b1t, B, N, C, Ds, st))) return rc;
    {  // dW2t^T [Ds, N] += sum_b H^T[b]^T [Ds, C] * dU[b]^T [C, N], added transposed into dW2t [N, Ds]: the same
       // (M = Ds in 128-row tiles, N = 196 in one 208-column tile) shape as dW1t below -- 1.2x padded MMA work
       // instead of the 1.7x of a [256-row x 4 x 256-column] tiling of [196, 784]
      vmlp_gemm_args g = gemm_args(Ds, N, C, B, opnd(s->h1, C, Ds, Ds, (long long)C * Ds, 1), opnd(dU, N, C, C, (long long)N * C, 0), VMLP_EPI_ATOMIC);
      g.contract_batch = 1; g.out_f32 = g_w2t; g.out_ld = Ds; g.out_trans = 1;
      if ((rc = gemm_impl(g, st))) return rc;
    }
    if (!fused_sums && (rc = vmlp_rowsum_batched(dU, g_b2t, B, N, C, stream))) return rc;
    {  // dW1t [Ds, N] += sum_b dZ^T[b]^T [Ds, C] * Xhat1[b]^T [C, N] (A operand MN-major)
      vmlp_gemm_args g = gemm_args(Ds, N, C, B, opnd(dZ, C, Ds, Ds, (long long)C * Ds, 1), opnd(s->xhat1, N, C, C, (long long)N * C, 0), VMLP_EPI_ATOMIC);
      g.contract_batch = 1; g.out_f32 = g_w1t; g.out_ld = N;
      if ((rc = gemm_impl(g, st))) return rc;
    }
  } else {
    {  // dZ1[b] [Ds, C] = (W2t^T [Ds, N] * dU[b] [N, C]) .* gelu'(Z1[b]);  W2t [N, Ds] is the MN-major A operand
      vmlp_gemm_args g = gemm_args(Ds, C, N, B, opnd(p->w2t, N, Ds, Ds, 0, 1), opnd(dU, N, C, C, (long long)N * C, 1), VMLP_EPI_DGELU);
      g.D = dZ; g.d_ld = C; g.d_bs = (long long)Ds * C;
      g.aux = s->z1; g.aux_ld = C; g.aux_bs = (long long)Ds * C;
      g.red_out = g_b1t; g.red_mode = 2;            // db1t[m] = sum over (batch, channels) of dZ1: per output row
      if ((rc = gemm_impl(g, st))) return rc;
    }
    {  // dW2t [N, Ds] += sum_b dU[b] [N, C] * H1[b]^T [C, Ds]    (contraction over batch and channels)
      vmlp_gemm_args g = gemm_args(N, Ds, C, B, opnd(dU, N, C, C, (long long)N * C, 0), opnd(s->h1, Ds, C, C, (long long)Ds * C, 0), VMLP_EPI_ATOMIC);
      g.contract_batch = 1; g.out_f32 = g_w2t; g.out_ld = Ds;
      if ((rc = gemm_impl(g, st))) return rc;
    }
    if (!fused_sums && (rc = vmlp_rowsum_batched(dU, g_b2t, B, N, C, stream))) return rc;
    {  // dXhat1[b] [N, C] = W1t^T [N, Ds] * dZ1[b] [Ds, C];  padded W1t [Ds, Np] as MN-major A
      vmlp_gemm_args g = gemm_args(N, C, Ds, B, opnd(s->w1t_pad, Ds, N, Np, 0, 1), opnd(dZ, Ds, C, C, (long long)Ds * C, 1), VMLP_EPI_STORE);
      g.D = dXh; g.d_ld = C; g.d_bs = (long long)N * C;
      if ((rc = gemm_impl(g, st))) return rc;
    }
    {  // dW1t [Ds, N] += sum_b dZ1[b] [Ds, C] * Xhat1[b]^T [C, N]
      vmlp_gemm_args g = gemm_args(Ds, N, C, B, opnd(dZ, Ds, C, C, (long long)Ds * C, 0), opnd(s->xhat1, N, C, C, (long long)N * C, 0), VMLP_EPI_ATOMIC);
      g.contract_batch = 1; g.out_f32 = g_w1t; g.out_ld = N;
      if ((rc = gemm_impl(g, st))) return rc;
    }
  }
  // dX = dU + LN1'(dXhat1)
  if ((rc = vmlp_layernorm_bwd(dXh, C, x, C, mean1, rstd1, p->ln1_w, dU, C, dx, C, g_ln1w, g_ln1b, R, C, stream))) return rc;
  return VMLP_OK;
}

}  // extern "C"
