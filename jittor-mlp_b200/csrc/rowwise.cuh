// HBM-bound row-wise kernels around the GEMMs: LayerNorm fwd/bwd, per-channel affine,
// column / batched-row reductions (bias gradients), pad / cast helpers.
// One warp owns one row at a time; every global access is a 16-byte vector; column
// partials stay in registers across the grid-stride loop and are flushed once per block.
#pragma once
#include "ptx.cuh"

namespace vmlp {

constexpr int RW_THREADS = 256;
constexpr int RW_WARPS = RW_THREADS / 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// 128-bit shared-memory read-modify-write of four fp32 partial sums
__device__ __forceinline__ void add_f4(float* p, float a, float b, float c, float d) {
  float4 v = *reinterpret_cast<float4*>(p);
  v.x += a; v.y += b; v.z += c; v.w += d;
  *reinterpret_cast<float4*>(p) = v;
}
// sum over aligned groups of G lanes (G = 8, 16, 32): narrow rows pack 32 / G rows into one warp
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  f[0] = bf16lo(u.x); f[1] = bf16hi(u.x); f[2] = bf16lo(u.y); f[3] = bf16hi(u.y);
  f[4] = bf16lo(u.z); f[5] = bf16hi(u.z); f[6] = bf16lo(u.w); f[7] = bf16hi(u.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                    pack_bf16x2(f[6], f[7]));
}

// Division-free index helpers.  A 64-bit div/mod per 16-byte vector costs ~100 instructions and turned the elementwise
// kernels instruction-bound; FastDiv (float reciprocal + one correction, exact for 0 <= n < 2^22) and VecIter (one
// division per THREAD, then incremental row/col updates along the grid-stride) replace them.
struct FastDiv {
  int d;
  float inv;
  __host__ __device__ explicit FastDiv(int d_) : d(d_), inv(1.0f / static_cast<float>(d_)) {}
  __device__ __forceinline__ void divmod(int n, int& q, int& r) const {
    q = __float2int_rz(static_cast<float>(n) * inv);
    r = n - q * d;
    if (r < 0) { r += d; --q; }
    else if (r >= d) { r -= d; ++q; }
  }
};
struct VecIter {   // walks a [rows, nvec] grid of vectors: element index = row * nvec + col
  long long row, drow;
  int col, dcol, nvec;
  __device__ __forceinline__ explicit VecIter(int nvec_) : nvec(nvec_) {
    const long long start = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    row = start / nvec; col = static_cast<int>(start % nvec);
    drow = stride / nvec; dcol = static_cast<int>(stride % nvec);
  }
  __device__ __forceinline__ void next() {
    row += drow; col += dcol;
    if (col >= nvec) { col -= nvec; ++row; }
  }
};

// Flush per-lane column partials (VPL vectors x 8 columns, lane-strided) of all warps of the block
// into global fp32 with one red.add per column per block.
template <int VPL>
__device__ __forceinline__ void flush_col_partials(float (&acc)[VPL][8], float* out, int nvec, float* sh) {
  // sh: [RW_WARPS][VPL*32*8] floats is too large for VPL=8 (64 KB); reduce warp by warp instead.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int w = 0; w < RW_WARPS; ++w) {
    __syncthreads();
    if (warp == w) {
#pragma unroll
      for (int v = 0; v < VPL; ++v)
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int idx = (v * 32 + lane) * 8 + e;
          sh[idx] = (w == 0 ? 0.f : sh[idx]) + acc[v][e];
        }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nvec * 8; i += RW_THREADS) red_add_f32(out + i, sh[i]);
}

// --------------------------------------------------------------------------- LayerNorm forward
// y = (x - mean) * rstd * gamma + beta over the last axis (biased variance, eps inside sqrt):
// torch.nn.LayerNorm semantics used by PreNormResidual (models_pytorch/mlp_mixer.py:6-13).
// G = lanes per row.  Rows of <= 128 channels (Hire-MLP stage 0/1: C = 64 / 128) would leave 24 / 16 lanes of a
// one-row-per-warp layout idle (measured 121 us for 205 MB, 4x off the HBM floor), so 32 / G rows share a warp.
template <int VPL, int G = 32>
__global__ void __launch_bounds__(RW_THREADS)
layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, long long x_ld, const __nv_bfloat16* __restrict__ gamma,
                     const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ y, long long y_ld,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, long long rows, int C,
                     float eps) {
  static_assert(G == 32 || VPL == 1, "row packing needs one vector per lane");
  constexpr int RPW = 32 / G;
  const int lane = threadIdx.x & 31;
  const int sl = lane & (G - 1), sub = lane / G;
  const int nvec = C >> 3;
  const long long gw = (long long)blockIdx.x * RW_WARPS + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * RW_WARPS;
  for (long long r0 = gw * RPW; r0 < rows; r0 += nw * RPW) {
    const long long r = r0 + sub;
    const bool row_ok = r < rows;
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = i * 32 + sl;
#pragma unroll
      for (int e = 0; e < 8; ++e) v[i][e] = 0.f;
      if (row_ok && vi < nvec) {
        unpack8(ldg_nc_v4(x + r * x_ld + vi * 8), v[i]);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += v[i][e];
      }
    }
    const float mean = group_sum<G>(s) / C;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      if (i * 32 + sl < nvec) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { const float d = v[i][e] - mean; ss += d * d; }
      }
    }
    const float rstd = rsqrtf(group_sum<G>(ss) / C + eps);
    if (sl == 0 && row_ok) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = i * 32 + sl;
      if (row_ok && vi < nvec) {
        float g[8], b[8], o[8];
        unpack8(*reinterpret_cast<const uint4*>(gamma + vi * 8), g);
        unpack8(*reinterpret_cast<const uint4*>(beta + vi * 8), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (v[i][e] - mean) * rstd * g[e] + b[e];
        *reinterpret_cast<uint4*>(y + r * y_ld + vi * 8) = pack8(o);
      }
    }
  }
}

// --------------------------------------------------------------------------- LayerNorm backward
// dx = add + rstd * (g - mean_C(g) - xhat * mean_C(g * xhat)),  g = dy * gamma,  xhat = (x - mean) * rstd
// dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy   (fp32 accumulators)
// HBM-bound (3 reads + 1 write per element): rows stay packed (bf16) in registers so that 3-4 blocks fit per
// SM and enough loads are in flight; the per-column partials of a block live in shared memory
// ([e][vector] layout => conflict-free red.shared) and are flushed with one global red.add per column.
// PRIV = 1: every warp owns a private [2][8][NV] fp32 slice of shared memory and updates it with plain ld/add/st
// (conflict-free, no atomic-unit serialisation); PRIV = 0 (rows longer than 1536): one slice per block, red.shared.
// EXTRA = 1 (Mixer block backward, PRIV only): the same pass also produces the two bias gradients that otherwise cost a
// pass each -- add_colsum[c] += sum_r add[r, c] (the channel-MLP output bias: `add` is the block's dY) and
// out_rowsum[r % row_period] += sum_c dx[r, c] (the token-MLP output bias: dx is the token half's dY).
// G < 32 (PRIV only): 32 / G narrow rows per warp, one private shared-memory slice per row slot.
template <int VPL, int PRIV, int EXTRA, int G = 32>
__global__ void __launch_bounds__(RW_THREADS, (VPL <= 4) ? 3 : 1)
layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy, long long dy_ld, const __nv_bfloat16* __restrict__ x,
                     long long x_ld, const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                     const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ add,
                     long long add_ld, __nv_bfloat16* __restrict__ dx, long long dx_ld,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows, int C,
                     float* __restrict__ add_colsum, float* __restrict__ out_rowsum, int row_period) {
  static_assert(G == 32 || (VPL == 1 && PRIV == 1 && EXTRA == 0), "row packing: one vector per lane, private slices");
  extern __shared__ float sh[];           // [PRIV ? warps * rows-per-warp : 1][2 + EXTRA][8][NV]
  constexpr int RPW = 32 / G;
  constexpr int NV = (G < 32) ? G : VPL * 32;
  constexpr int SLICES = PRIV ? RW_WARPS * RPW : 1;
  constexpr int SL = (2 + EXTRA) * 8 * NV;      // floats per slice
  const int lane = threadIdx.x & 31;
  const int sl = lane & (G - 1), sub = lane / G;
  float* sh_g = sh + (PRIV ? ((threadIdx.x >> 5) * RPW + sub) * SL : 0);
  float* sh_b = sh_g + 8 * NV;
  float* sh_a = sh_b + 8 * NV;                  // EXTRA only
  const int nvec = C >> 3;
  for (int i = threadIdx.x; i < SLICES * SL; i += RW_THREADS) sh[i] = 0.f;
  __syncthreads();
  const long long gw = (long long)blockIdx.x * RW_WARPS + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * RW_WARPS;
  for (long long r0 = gw * RPW; r0 < rows; r0 += nw * RPW) {
    const long long r = r0 + sub;
    const bool row_ok = r < rows;
    float mean = 0.f, rstd = 0.f;
    if (row_ok) { mean = mean_in[r]; rstd = rstd_in[r]; }
    uint4 rd[VPL], rx[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = i * 32 + sl;
      rd[i] = make_uint4(0, 0, 0, 0);
      rx[i] = make_uint4(0, 0, 0, 0);
      if (row_ok && vi < nvec) {
        rd[i] = ldg_nc_v4(dy + r * dy_ld + vi * 8);
        rx[i] = ldg_nc_v4(x + r * x_ld + vi * 8);
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = i * 32 + sl;
      if (row_ok && vi < nvec) {
        float d[8], xv[8], gm[8], pg[8];
        unpack8(rd[i], d);
        unpack8(rx[i], xv);
        unpack8(*reinterpret_cast<const uint4*>(gamma + vi * 8), gm);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[e] - mean) * rstd;
          const float g = d[e] * gm[e];
          s1 += g;
          s2 += g * xh;
          if (PRIV) {
            pg[e] = d[e] * xh;
          } else {
            atomicAdd(&sh_g[e * NV + vi], d[e] * xh);
            atomicAdd(&sh_b[e * NV + vi], d[e]);
          }
        }
        if (PRIV) {
          // warp-private slice, [half][vector][4] layout: 128-bit read-modify-writes (conflict-free: lane l touches the
          // 16 bytes at l * 16), 8 shared-memory instructions per 8 elements where scalar updates needed 32
          add_f4(sh_g + vi * 4, pg[0], pg[1], pg[2], pg[3]);
          add_f4(sh_g + (NV + vi) * 4, pg[4], pg[5], pg[6], pg[7]);
          add_f4(sh_b + vi * 4, d[0], d[1], d[2], d[3]);
          add_f4(sh_b + (NV + vi) * 4, d[4], d[5], d[6], d[7]);
        }
      }
    }
    s1 = group_sum<G>(s1) / C;
    s2 = group_sum<G>(s2) / C;
    float osum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = i * 32 + sl;
      if (row_ok && vi < nvec) {
        float d[8], xv[8], gm[8], o[8], a[8];
        unpack8(rd[i], d);
        unpack8(rx[i], xv);
        unpack8(*reinterpret_cast<const uint4*>(gamma + vi * 8), gm);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = 0.f;
        if (add) unpack8(ldg_nc_v4(add + r * add_ld + vi * 8), a);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xh = (xv[e] - mean) * rstd;
          o[e] = a[e] + rstd * (d[e] * gm[e] - s1 - xh * s2);
          if (EXTRA) osum += o[e];
        }
        if (EXTRA) {
          add_f4(sh_a + vi * 4, a[0], a[1], a[2], a[3]);
          add_f4(sh_a + (NV + vi) * 4, a[4], a[5], a[6], a[7]);
        }
        *reinterpret_cast<uint4*>(dx + r * dx_ld + vi * 8) = pack8(o);
      }
    }
    if (EXTRA && out_rowsum != nullptr) {
      osum = warp_sum(osum);
      if (lane == 0) red_add_f32(out_rowsum + (int)(r % row_period), osum);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nvec * 8; i += RW_THREADS) {
    const int vi = i >> 3, e = i & 7;
    float sg = 0.f, sb = 0.f, sa = 0.f;
    // PRIV slices use the [half][vector][4] layout, the shared (atomic) slice the [e][vector] one
    const int off = PRIV ? (((e >> 2) * NV + vi) * 4 + (e & 3)) : (e * NV + vi);
#pragma unroll
    for (int w = 0; w < SLICES; ++w) {
      sg += sh[w * SL + off];
      sb += sh[w * SL + 8 * NV + off];
      if (EXTRA) sa += sh[w * SL + 16 * NV + off];
    }
    red_add_f32(dgamma + i, sg);
    red_add_f32(dbeta + i, sb);
    if (EXTRA && add_colsum != nullptr) red_add_f32(add_colsum + i, sa);
  }
}

// --------------------------------------------------------------------------- per-channel affine (ResMLP Aff)
// y = x * alpha + beta (models_pytorch/res_mlp.py:11-19)
__global__ void __launch_bounds__(RW_THREADS)
affine_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ alpha,
                  const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ y, long long nvec_total,
                  int nvec_row) {
  const long long rows = nvec_total / nvec_row;
  for (VecIter it(nvec_row); it.row < rows; it.next()) {
    const long long i = it.row * nvec_row + it.col;
    const int c = it.col * 8;
    float xv[8], a[8], b[8], o[8];
    unpack8(ldg_nc_v4(x + i * 8), xv);
    unpack8(*reinterpret_cast<const uint4*>(alpha + c), a);
    unpack8(*reinterpret_cast<const uint4*>(beta + c), b);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = xv[e] * a[e] + b[e];
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(o);
  }
}
// dx = dy * alpha (+ add); dalpha += sum dy * x ; dbeta += sum dy
template <int VPL>
__global__ void __launch_bounds__(RW_THREADS)
affine_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                  const __nv_bfloat16* __restrict__ alpha, const __nv_bfloat16* __restrict__ add,
                  __nv_bfloat16* __restrict__ dx, float* __restrict__ dalpha, float* __restrict__ dbeta,
                  long long rows, int C) {
  extern __shared__ float sh[];
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  const long long gw = (long long)blockIdx.x * RW_WARPS + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * RW_WARPS;
  float aa[VPL][8], ab[VPL][8], al[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = i * 32 + lane;
#pragma unroll
    for (int e = 0; e < 8; ++e) { aa[i][e] = 0.f; ab[i][e] = 0.f; al[i][e] = 0.f; }
    if (vi < nvec) unpack8(*reinterpret_cast<const uint4*>(alpha + vi * 8), al[i]);
  }
  for (long long r = gw; r < rows; r += nw) {
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int vi = i * 32 + lane;
      if (vi < nvec) {
        float d[8], xv[8], a[8], o[8];
        unpack8(ldg_nc_v4(dy + r * C + vi * 8), d);
        unpack8(ldg_nc_v4(x + r * C + vi * 8), xv);
#pragma unroll
        for (int e = 0; e < 8; ++e) a[e] = 0.f;
        if (add) unpack8(ldg_nc_v4(add + r * C + vi * 8), a);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          aa[i][e] += d[e] * xv[e];
          ab[i][e] += d[e];
          o[e] = d[e] * al[i][e] + a[e];
        }
        *reinterpret_cast<uint4*>(dx + r * C + vi * 8) = pack8(o);
      }
    }
  }
  flush_col_partials<VPL>(aa, dalpha, nvec, sh);
  flush_col_partials<VPL>(ab, dbeta, nvec, sh);
}

// --------------------------------------------------------------------------- column sums  out[c] += sum_r (a[r,c] (* b[r,c]))
// grid = (row groups, 256-column slabs): a warp reads 512 contiguous bytes per row of its slab and keeps
// 8 fp32 partials per lane; one smem reduction + one global red.add per column per block.
__global__ void __launch_bounds__(RW_THREADS)
colsum_kernel(const __nv_bfloat16* __restrict__ a, long long a_ld, const __nv_bfloat16* __restrict__ b,
              long long b_ld, float* __restrict__ out, long long rows, int C) {
  __shared__ float sh[RW_WARPS][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int col = blockIdx.y * 256 + lane * 8;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (col < C) {
    for (long long r = (long long)blockIdx.x * RW_WARPS + warp; r < rows; r += (long long)gridDim.x * RW_WARPS) {
      float d[8];
      unpack8(ldg_nc_v4(a + r * a_ld + col), d);
      if (b) {
        float m[8];
        unpack8(ldg_nc_v4(b + r * b_ld + col), m);
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] *= m[e];
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += d[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sh[warp][lane * 8 + e] = acc[e];
  __syncthreads();
  const int c = threadIdx.x;     // 256 threads <-> 256 columns of the slab
  if (blockIdx.y * 256 + c < C) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < RW_WARPS; ++w) s += sh[w][c];
    red_add_f32(out + blockIdx.y * 256 + c, s);
  }
}

// Rows of C <= 2048 channels (any row stride): thread t owns channel vector (t % nvec) of row slot (t / nvec), so a
// block streams whole rows fully coalesced whatever C is (the slab kernel above keeps 20 of 32 lanes idle at C = 96) and
// keeps four 16-byte loads per operand in flight.  Block reduction: per-slot partials in shared memory, summed by one
// thread per column (shared-memory float atomics serialise; with ~10 us of work per block they dominated).
// DUAL = 1: out[c] += sum_r a[r, c] AND out2[c] += sum_r a[r, c] * b[r, c] in the same pass -- the two BatchNorm
// statistics (sum, sum of squares with b == a; sum dy, sum dy * a in backward) read each tensor once instead of twice.
template <int DUAL>
__global__ void __launch_bounds__(RW_THREADS)
colsum_flat_kernel(const __nv_bfloat16* __restrict__ a, long long a_ld, const __nv_bfloat16* __restrict__ b,
                   long long b_ld, float* __restrict__ out, float* __restrict__ out2, long long rows, int C) {
  extern __shared__ float shc[];          // [(1 + DUAL)][rpb][C]
  const int nvec = C >> 3;
  const int rpb = RW_THREADS / nvec;      // rows per block iteration
  const int cv = threadIdx.x % nvec, ro = threadIdx.x / nvec;
  const bool same = DUAL && (a == b);
  float acc[8], acc2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { acc[e] = 0.f; acc2[e] = 0.f; }
  auto accumulate = [&](const uint4& va, const uint4& vb) {
    float d[8], m[8];
    unpack8(va, d);
    if (b) unpack8(vb, m);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (DUAL) { acc[e] += d[e]; acc2[e] += d[e] * m[e]; }
      else acc[e] += b ? d[e] * m[e] : d[e];
    }
  };
  if (ro < rpb) {
    const long long step = (long long)gridDim.x * rpb;
    long long r = (long long)blockIdx.x * rpb + ro;
    const __nv_bfloat16* pa = a + cv * 8;
    const __nv_bfloat16* pb = b ? b + cv * 8 : nullptr;
    for (; r + 3 * step < rows; r += 4 * step) {
      uint4 va[4], vb[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        va[u] = ldg_nc_v4(pa + (r + u * step) * a_ld);
        vb[u] = va[u];
        if (b && !same) vb[u] = ldg_nc_v4(pb + (r + u * step) * b_ld);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) accumulate(va[u], vb[u]);
    }
    for (; r < rows; r += step) {
      const uint4 va = ldg_nc_v4(pa + r * a_ld);
      uint4 vb = va;
      if (b && !same) vb = ldg_nc_v4(pb + r * b_ld);
      accumulate(va, vb);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      shc[ro * C + cv * 8 + e] = acc[e];
      if (DUAL) shc[(rpb + ro) * C + cv * 8 + e] = acc2[e];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C; i += RW_THREADS) {
    float s1 = 0.f, s2 = 0.f;
    for (int k = 0; k < rpb; ++k) {
      s1 += shc[k * C + i];
      if (DUAL) s2 += shc[(rpb + k) * C + i];
    }
    red_add_f32(out + i, s1);
    if (DUAL) red_add_f32(out2 + i, s2);
  }
}

// --------------------------------------------------------------------------- batched row sums
// out[m] += sum_{b, c} a[b, m, c]   (token-mixing bias gradients: reductions over B*C)
__global__ void __launch_bounds__(RW_THREADS)
rowsum_batched_kernel(const __nv_bfloat16* __restrict__ a, float* __restrict__ out, long long rows_total,
                      int rows_per_batch, int C) {
  const int lane = threadIdx.x & 31;
  const int nvec = C >> 3;
  const long long gw = (long long)blockIdx.x * RW_WARPS + (threadIdx.x >> 5);
  const long long nw = (long long)gridDim.x * RW_WARPS;
  for (long long r = gw; r < rows_total; r += nw) {
    float s = 0.f;
    for (int vi = lane; vi < nvec; vi += 32) {
      float d[8];
      unpack8(ldg_nc_v4(a + r * C + vi * 8), d);
#pragma unroll
      for (int e = 0; e < 8; ++e) s += d[e];
    }
    s = warp_sum(s);
    if (lane == 0) red_add_f32(out + (r % rows_per_batch), s);
  }
}

// --------------------------------------------------------------------------- small helpers
// dst[r, 0:cols] = src[r, 0:cols], dst[r, cols:ld_dst] = 0  (makes the row pitch a multiple of 16 B for TMA)
__global__ void pad_rows_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int rows,
                                int cols, int ld_dst) {
  const long long n = (long long)rows * ld_dst;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = static_cast<int>(i / ld_dst), c = static_cast<int>(i % ld_dst);
    dst[i] = (c < cols) ? src[(long long)r * cols + c] : __float2bfloat16(0.f);
  }
}
__global__ void cast_f32_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = __float2bfloat16(src[i]);
}
// dst[i] = a[i] + b[i]
__global__ void add_bf16_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                __nv_bfloat16* __restrict__ dst, long long nvec) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    float x[8], y[8];
    unpack8(ldg_nc_v4(a + i * 8), x);
    unpack8(ldg_nc_v4(b + i * 8), y);
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] += y[e];
    *reinterpret_cast<uint4*>(dst + i * 8) = pack8(x);
  }
}

// --------------------------------------------------------------------------- strided elementwise helpers
// All take [rows, C] views with explicit row strides (elements); C % 8 == 0.
// MODE 0: out = a * colvec[c]                                  (ResMLP: dF = dY * gamma, res_mlp.py:54,56)
// MODE 1: out = a * b          (b = gelu'(z) saved by the forward epilogue: d(pre-activation) = d(activation) * gelu')
// MODE 2: out = a * b * c3 ; out2 = a * d4   (gMLP gate backward: dZp_u = dG * vt * gelu'(Zp_u), dVt = dG * u)
template <int MODE>
__global__ void __launch_bounds__(RW_THREADS)
ew_kernel(const __nv_bfloat16* __restrict__ a, long long a_ld, const __nv_bfloat16* __restrict__ b, long long b_ld,
          const __nv_bfloat16* __restrict__ c3, long long c3_ld, const __nv_bfloat16* __restrict__ d4, long long d4_ld,
          __nv_bfloat16* __restrict__ out, long long out_ld, __nv_bfloat16* __restrict__ out2, long long out2_ld,
          long long rows, int C) {
  const int nvr = C >> 3;
  for (VecIter it(nvr); it.row < rows; it.next()) {
    const long long r = it.row;
    const int col = it.col * 8;
    float x[8], y[8], o[8];
    unpack8(ldg_nc_v4(a + r * a_ld + col), x);
    if (MODE == 0) {
      unpack8(*reinterpret_cast<const uint4*>(b + col), y);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = x[e] * y[e];
    } else if (MODE == 1) {
      unpack8(ldg_nc_v4(b + r * b_ld + col), y);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = x[e] * y[e];
    } else {
      float z[8], u[8], o2[8];
      unpack8(ldg_nc_v4(b + r * b_ld + col), y);
      unpack8(ldg_nc_v4(c3 + r * c3_ld + col), z);
      unpack8(ldg_nc_v4(d4 + r * d4_ld + col), u);
#pragma unroll
      for (int e = 0; e < 8; ++e) { o[e] = x[e] * y[e] * z[e]; o2[e] = x[e] * u[e]; }
      *reinterpret_cast<uint4*>(out2 + r * out2_ld + col) = pack8(o2);
    }
    *reinterpret_cast<uint4*>(out + r * out_ld + col) = pack8(o);
  }
}

}  // namespace vmlp
