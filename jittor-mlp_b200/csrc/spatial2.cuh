// Hire-MLP region rearrangement (hire_mlp.py:53-152) and ConvMixer depthwise convolution (conv_mixer.py:24)
// on channels-last tensors.  Companion of spatial.cuh.
#pragma once
#include "spatial.cuh"

namespace vmlp {

// ============================================================================================ Hire-MLP
// x: [B, H, W, C] (LayerNorm output).  Circular padding to Hp = H + (h - H % h), Wp = W + (w - W % w)
// (hire_mlp.py:134-136: a FULL extra region when already divisible), roll by `step` (CrossRegion, :44-51),
// then the strided region gather  z[b, c*h + i, g, w] = x_p[b, c, i*G + g, w]  (InnerRegionH, :64-73).
// Here the gathered feature axis is ordered [i][c] (the host permutes the 1x1-conv weights to match), so each
// GEMM row is n contiguous C-vectors fetched from n different tokens -- pure index arithmetic in the loads,
// no pad / roll / rearrange tensor is ever materialised.  Columns (rows) that exist only because of the padding
// are never read by the cropped output and are skipped.
struct HireDims {
  int B, H, W, C;
  int nh, Gh, Hp, step_h;   // H branch: nh regions' worth of tokens per row, Gh = Hp / nh
  int nw, Gw, Wp, step_w;   // W branch
};
__device__ __forceinline__ int pmod(int a, int m) { a %= m; return a < 0 ? a + m : a; }

// DIR 0 (H branch): Z[((b*Gh + g)*W + w)*nh*C + i*C + c] = x[b, ((i*Gh + g - step) mod Hp) mod H, w, c]
// DIR 1 (W branch): Z[((b*H + r)*Gw + g)*nw*C + j*C + c] = x[b, r, ((j*Gw + g - step) mod Wp) mod W, c]
template <int DIR>
__global__ void __launch_bounds__(RW_THREADS)
hire_build_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ z, const HireDims d) {
  const int nvec = d.C >> 3;
  const int n = DIR ? d.nw : d.nh, G = DIR ? d.Gw : d.Gh;
  const int L = DIR ? d.H : d.W;
  const int per = G * L * n * nvec;                 // vectors per sample (batch on blockIdx.y, division-free decode)
  const FastDiv dv(nvec), dn(n), dl(DIR ? G : d.W);
  const long long b = blockIdx.y;
  for (int idx = blockIdx.x * RW_THREADS + threadIdx.x; idx < per; idx += gridDim.x * RW_THREADS) {
    int t, cv, t2, i, hi, lo;
    dv.divmod(idx, t, cv);
    dn.divmod(t, t2, i);
    dl.divmod(t2, hi, lo);
    int r, w;
    if (DIR == 0) { w = lo; r = pmod(i * G + hi - d.step_h, d.Hp) % d.H; }        // t2 = g * W + w
    else { r = hi; w = pmod(i * G + lo - d.step_w, d.Wp) % d.W; }                 // t2 = r * G + g
    *reinterpret_cast<uint4*>(z + (b * per + idx) * 8) = ldg_nc_v4(x + ((b * d.H + r) * d.W + w) * d.C + cv * 8);
  }
}
// adjoint of both builds: dx[b, r, w, c] = sum over padded copies of r of dZh[...] + sum over copies of w of dZw[...]
__global__ void __launch_bounds__(RW_THREADS)
hire_build_adj_kernel(const __nv_bfloat16* __restrict__ dzh, const __nv_bfloat16* __restrict__ dzw,
                      __nv_bfloat16* __restrict__ dx, const HireDims d) {
  const int nvec = d.C >> 3;
  const int per = d.H * d.W * nvec;
  const FastDiv dv(nvec), dw_(d.W);
  const long long b = blockIdx.y;
  for (int i0 = blockIdx.x * RW_THREADS + threadIdx.x; i0 < per; i0 += gridDim.x * RW_THREADS) {
    int pos, cv, r, w;
    dv.divmod(i0, pos, cv);
    dw_.divmod(pos, r, w);
    const int c0 = cv * 8;
    const long long idx = b * per + i0;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int rp = r; rp < d.Hp; rp += d.H) {
      const int rr = pmod(rp + d.step_h, d.Hp);
      const int i = rr / d.Gh, g = rr % d.Gh;
      float v[8];
      unpack8(ldg_nc_v4(dzh + (((b * d.Gh + g) * d.W + w) * d.nh + i) * d.C + c0), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
    for (int wp = w; wp < d.Wp; wp += d.W) {
      const int ww = pmod(wp + d.step_w, d.Wp);
      const int j = ww / d.Gw, g = ww % d.Gw;
      float v[8];
      unpack8(ldg_nc_v4(dzw + (((b * d.H + r) * d.Gw + g) * d.nw + j) * d.C + c0), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += v[e];
    }
    *reinterpret_cast<uint4*>(dx + idx * 8) = pack8(acc);
  }
}
// out[b, r, w, c] = base + restore_H(Oh) + restore_W(Ow), cropped to H x W (hire_mlp.py:143-151):
//   restore_H: T[c, i*G + g, w] = Oh[(b, g, w), i*C + c]; X_h[r] = T[(r + step) mod Hp]
__global__ void __launch_bounds__(RW_THREADS)
hire_combine_kernel(const __nv_bfloat16* __restrict__ base, const __nv_bfloat16* __restrict__ oh,
                    const __nv_bfloat16* __restrict__ ow, __nv_bfloat16* __restrict__ out, const HireDims d) {
  const int nvec = d.C >> 3;
  const int per = d.H * d.W * nvec;
  const FastDiv dv(nvec), dw_(d.W);
  const long long b = blockIdx.y;
  for (int i0 = blockIdx.x * RW_THREADS + threadIdx.x; i0 < per; i0 += gridDim.x * RW_THREADS) {
    int pos, cv, r, w;
    dv.divmod(i0, pos, cv);
    dw_.divmod(pos, r, w);
    const int c0 = cv * 8;
    const long long idx = b * per + i0;
    float a[8], v[8];
    unpack8(ldg_nc_v4(base + idx * 8), a);
    const int rr = pmod(r + d.step_h, d.Hp);
    unpack8(ldg_nc_v4(oh + (((b * d.Gh + rr % d.Gh) * d.W + w) * d.nh + rr / d.Gh) * d.C + c0), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += v[e];
    const int ww = pmod(w + d.step_w, d.Wp);
    unpack8(ldg_nc_v4(ow + (((b * d.H + r) * d.Gw + ww % d.Gw) * d.nw + ww / d.Gw) * d.C + c0), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] += v[e];
    *reinterpret_cast<uint4*>(out + idx * 8) = pack8(a);
  }
}
// adjoint of the restore: dO[(.., g, ..), i*C + c] = dout at the token this entry lands on, 0 if it is cropped away
template <int DIR>
__global__ void __launch_bounds__(RW_THREADS)
hire_restore_adj_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ dz, const HireDims d) {
  const int nvec = d.C >> 3;
  const int n = DIR ? d.nw : d.nh, G = DIR ? d.Gw : d.Gh;
  const int L = DIR ? d.H : d.W;
  const int per = G * L * n * nvec;
  const FastDiv dv(nvec), dn(n), dl(DIR ? G : d.W);
  const long long b = blockIdx.y;
  for (int idx = blockIdx.x * RW_THREADS + threadIdx.x; idx < per; idx += gridDim.x * RW_THREADS) {
    int t, cv, t2, i, hi, lo;
    dv.divmod(idx, t, cv);
    dn.divmod(t, t2, i);
    dl.divmod(t2, hi, lo);
    int r, w;
    bool live;
    if (DIR == 0) { w = lo; r = pmod(i * G + hi - d.step_h, d.Hp); live = r < d.H; }
    else { r = hi; w = pmod(i * G + lo - d.step_w, d.Wp); live = w < d.W; }
    uint4 v = make_uint4(0, 0, 0, 0);
    if (live) v = ldg_nc_v4(dout + ((b * d.H + r) * d.W + w) * d.C + cv * 8);
    *reinterpret_cast<uint4*>(dz + (b * per + idx) * 8) = v;
  }
}

// ============================================================================================ ConvMixer depthwise conv
// nn.Conv2d(dim, dim, k, groups=dim, padding="same") on [B, H, W, C] (conv_mixer.py:24): a shared-memory stencil.
// Block = 16 x 8 output pixels x 64 channels, 128 threads: lane = channel PAIR (one 32-bit shared-memory word holds
// both bf16 channels, all arithmetic is packed fp32x2: FFMA2), warp = four output rows.  The warp slides over its
// 4 + K - 1 input rows once; each input row is loaded into a register window (TW + K - 1 pairs) and feeds the up to
// four output rows it overlaps, so one window load serves up to 4 * K * 8 packed FMAs.  FLIP = 1 evaluates the same
// stencil with the kernel rotated by 180 degrees, which is the input gradient.  EPI = 1 adds the bias and writes
// gelu'(z) (for backward) and gelu(z).
// (Round-1 version: scalar bf16 loads, one channel per thread, 2 output rows: 15.8 TFLOP/s fp32, 1.25 ms per layer.)
constexpr int DW_TH = 16;                 // tile rows
constexpr int DW_TW = 8;                  // tile columns
constexpr int DW_CH = 64;                 // channels per block (32 pairs = one warp's lanes)
constexpr int DW_ROWS = 4;                // output rows per warp
constexpr int DW_THREADS = 32 * (DW_TH / DW_ROWS);
template <int K>
struct DwSmem {
  static constexpr int IN_H = DW_TH + K - 1;
  static constexpr int IN_W = DW_TW + K - 1;
  static constexpr int IN_ELEMS = IN_H * IN_W * DW_CH;          // bf16
  static constexpr int W_ELEMS = K * K * DW_CH;                 // float ([tap][pair] float2)
  static constexpr int DZ_ELEMS = DW_TH * DW_TW * DW_CH;        // bf16 (wgrad only)
  static constexpr int BYTES = 2 * IN_ELEMS * 2 + W_ELEMS * 4 + 128 + 16;              // two input buffers + weights + align + barriers
  static constexpr int BYTES_WGRAD = 2 * (IN_ELEMS * 2 + DZ_ELEMS * 2) + 128 + 16;    // two (x, dz) buffer pairs
};

// Tile loads are single TMA instructions: a 4-D map (C, W, H, B) with box (64, COLS + K - 1, ROWS + K - 1, 1) lands the
// halo tile as [row][col][64 channels] in shared memory; coordinates outside the image (the "same" padding) and
// channels past C are zero-filled by the engine.  One elected thread issues the load of the NEXT tile into the other
// buffer before the block computes the current one.  (Per-thread 16-byte loads with their address arithmetic were 20 %
// of all executed instructions and, before double buffering, 40 % of the warp time as long-scoreboard stalls.)
__device__ __forceinline__ f32x2 dw_lds_pair(const __nv_bfloat16* p) {     // two adjacent bf16 channels -> packed fp32x2
  const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
  return pack2(bf16lo(w), bf16hi(w));
}

template <int K, int FLIP, int EPI>
__global__ void __launch_bounds__(DW_THREADS)
dwconv_kernel(const __grid_constant__ CUtensorMap tmX, const __nv_bfloat16* __restrict__ wgt,
              const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out2,
              int B, int H, int W, int C) {
  extern __shared__ uint8_t dw_smem_raw[];
  uint8_t* dw_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dw_smem_raw) + 127) & ~uintptr_t(127));
  constexpr int IN_W = DwSmem<K>::IN_W, P = K / 2;
  __nv_bfloat16* s_buf = reinterpret_cast<__nv_bfloat16*>(dw_smem);               // 2 x IN_ELEMS
  f32x2* s_w = reinterpret_cast<f32x2*>(dw_smem + 2 * DwSmem<K>::IN_ELEMS * 2);    // [tap][pair]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dw_smem + 2 * DwSmem<K>::IN_ELEMS * 2 + DwSmem<K>::W_ELEMS * 4);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int tiles_w = (W + DW_TW - 1) / DW_TW, tiles_h = (H + DW_TH - 1) / DW_TH;
  const int c0 = blockIdx.y * DW_CH;
  const int pr = threadIdx.x & 31, rg = threadIdx.x >> 5;      // channel pair, row group (warp-uniform)
  const int c = c0 + 2 * pr;
  for (int i = threadIdx.x; i < K * K * (DW_CH / 2); i += blockDim.x) {
    const int pp = i % (DW_CH / 2), tap = i / (DW_CH / 2);
    const int src = FLIP ? (K * K - 1 - tap) : tap;
    const int cc = c0 + 2 * pp;
    const float wa = (cc < C) ? __bfloat162float(wgt[(long long)cc * K * K + src]) : 0.f;
    const float wb = (cc + 1 < C) ? __bfloat162float(wgt[(long long)(cc + 1) * K * K + src]) : 0.f;
    s_w[i] = pack2(wa, wb);
  }
  f32x2 bv = pack2(0.f, 0.f);
  if (EPI && c < C) bv = pack2(__bfloat162float(bias[c]), __bfloat162float(bias[c + 1]));
  const FastDiv dtw(tiles_w), dth(tiles_h);
  const int ntiles = B * tiles_h * tiles_w;
  auto issue = [&](int tile, int buf) {                    // one thread: TMA load of a halo tile into buffer `buf`
    int t1, tw_i, b, th_i;
    dtw.divmod(tile, t1, tw_i);
    dth.divmod(t1, b, th_i);
    mbar_arrive_expect_tx(&bars[buf], DwSmem<K>::IN_ELEMS * 2);
    tma_load_4d(s_buf + buf * DwSmem<K>::IN_ELEMS, &tmX, &bars[buf], c0, tw_i * DW_TW - P, th_i * DW_TH - P, b);
  };
  if (threadIdx.x == 0 && static_cast<int>(blockIdx.x) < ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    int t1, tw_i, b, th_i;
    dtw.divmod(tile, t1, tw_i);
    dth.divmod(t1, b, th_i);
    const int h0 = th_i * DW_TH, w0 = tw_i * DW_TW;
    const __nv_bfloat16* s_in = s_buf + (it & 1) * DwSmem<K>::IN_ELEMS;
    __syncthreads();                 // every warp is done with the other buffer (previous tile) and with s_w setup
    if (threadIdx.x == 0 && tile + static_cast<int>(gridDim.x) < ntiles) issue(tile + gridDim.x, (it + 1) & 1);
    mbar_wait(&bars[it & 1], (it >> 1) & 1);     // this tile has landed
    f32x2 acc[DW_ROWS][DW_TW];
#pragma unroll
    for (int r = 0; r < DW_ROWS; ++r)
#pragma unroll
      for (int q = 0; q < DW_TW; ++q) acc[r][q] = bv;
    const __nv_bfloat16* rowp = s_in + (rg * DW_ROWS) * IN_W * DW_CH + 2 * pr;
#pragma unroll 1
    for (int ir = 0; ir < DW_ROWS + K - 1; ++ir) {
      f32x2 win[IN_W];
#pragma unroll
      for (int q = 0; q < IN_W; ++q) win[q] = dw_lds_pair(rowp + (ir * IN_W + q) * DW_CH);
#pragma unroll
      for (int r = 0; r < DW_ROWS; ++r) {
        const int i = ir - r;                  // tap row of output row r that reads input row ir (warp-uniform)
        if (i >= 0 && i < K) {
#pragma unroll
          for (int j = 0; j < K; ++j) {
            const f32x2 wv = s_w[(i * K + j) * (DW_CH / 2) + pr];
#pragma unroll
            for (int q = 0; q < DW_TW; ++q) acc[r][q] = fma2(wv, win[q + j], acc[r][q]);
          }
        }
      }
    }
    if (c < C) {
#pragma unroll
      for (int r = 0; r < DW_ROWS; ++r) {
        const int hh = h0 + rg * DW_ROWS + r;
        if (hh >= H) continue;
        const long long orow = (((long long)b * H + hh) * W + w0) * C + c;
#pragma unroll
        for (int q = 0; q < DW_TW; ++q) {
          if (w0 + q >= W) continue;
          const long long o = orow + (long long)q * C;
          if (EPI == 1) {       // out = gelu'(z) (kept for backward), out2 = gelu(z); EPI == 2: bias only, plain store
            f32x2 gl, dg;
            gelu_erf_pair<true>(acc[r][q], gl, dg);
            *reinterpret_cast<uint32_t*>(out + o) = pack_bf16x2_f2(dg);
            *reinterpret_cast<uint32_t*>(out2 + o) = pack_bf16x2_f2(gl);
          } else {
            *reinterpret_cast<uint32_t*>(out + o) = pack_bf16x2_f2(acc[r][q]);
          }
        }
      }
    }
  }
}

// dW[c][i][j] += sum_{b,h,w} dz[b,h,w,c] * x[b, h+i-P, w+j-P, c]   (fp32 atomics once per block)
// K warps: warp = tap row i, lane = channel pair; K packed accumulators per thread live in registers across all tiles
// of the block.  Per output row the warp reads the dz row (8 pairs) and the x row it pairs with (8 + K - 1 pairs).
template <int K>
__global__ void __launch_bounds__(32 * K)
dwconv_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDz,
                    float* __restrict__ dw, int B, int H, int W, int C) {
  extern __shared__ uint8_t dw_smem_raw[];
  uint8_t* dw_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dw_smem_raw) + 127) & ~uintptr_t(127));
  constexpr int IN_W = DwSmem<K>::IN_W, P = K / 2;
  constexpr int STAGE = DwSmem<K>::IN_ELEMS + DwSmem<K>::DZ_ELEMS;            // bf16 elements per (x, dz) buffer pair
  __nv_bfloat16* s_buf = reinterpret_cast<__nv_bfloat16*>(dw_smem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(dw_smem + 2 * STAGE * 2);
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmX);
    tma_prefetch_desc(&tmDz);
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int tiles_w = (W + DW_TW - 1) / DW_TW, tiles_h = (H + DW_TH - 1) / DW_TH;
  const int c0 = blockIdx.y * DW_CH;
  const int pr = threadIdx.x & 31, ti = threadIdx.x >> 5;      // channel pair, tap row (warp-uniform)
  f32x2 acc[K];
#pragma unroll
  for (int j = 0; j < K; ++j) acc[j] = pack2(0.f, 0.f);
  const FastDiv dtw(tiles_w), dth(tiles_h);
  const int ntiles = B * tiles_h * tiles_w;
  auto issue = [&](int tile, int buf) {                    // one thread: x halo tile + dz tile of the same origin
    int t1, tw_i, b, th_i;
    dtw.divmod(tile, t1, tw_i);
    dth.divmod(t1, b, th_i);
    mbar_arrive_expect_tx(&bars[buf], STAGE * 2);
    tma_load_4d(s_buf + buf * STAGE, &tmX, &bars[buf], c0, tw_i * DW_TW - P, th_i * DW_TH - P, b);
    tma_load_4d(s_buf + buf * STAGE + DwSmem<K>::IN_ELEMS, &tmDz, &bars[buf], c0, tw_i * DW_TW, th_i * DW_TH, b);
  };
  if (threadIdx.x == 0 && static_cast<int>(blockIdx.x) < ntiles) issue(blockIdx.x, 0);
  int it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
    const __nv_bfloat16* s_in = s_buf + (it & 1) * STAGE;
    const __nv_bfloat16* s_dz = s_in + DwSmem<K>::IN_ELEMS;
    __syncthreads();
    if (threadIdx.x == 0 && tile + static_cast<int>(gridDim.x) < ntiles) issue(tile + gridDim.x, (it + 1) & 1);
    mbar_wait(&bars[it & 1], (it >> 1) & 1);
#pragma unroll 2
    for (int r = 0; r < DW_TH; ++r) {
      f32x2 g[DW_TW], win[IN_W];
#pragma unroll
      for (int q = 0; q < DW_TW; ++q) g[q] = dw_lds_pair(s_dz + (r * DW_TW + q) * DW_CH + 2 * pr);
#pragma unroll
      for (int q = 0; q < IN_W; ++q) win[q] = dw_lds_pair(s_in + ((r + ti) * IN_W + q) * DW_CH + 2 * pr);
#pragma unroll
      for (int j = 0; j < K; ++j)
#pragma unroll
        for (int q = 0; q < DW_TW; ++q) acc[j] = fma2(g[q], win[q + j], acc[j]);
    }
  }
  const int c = c0 + 2 * pr;
  if (c < C) {
#pragma unroll
    for (int j = 0; j < K; ++j) {
      float a0, a1;
      unpack2(acc[j], a0, a1);
      red_add_f32(dw + (long long)c * K * K + ti * K + j, a0);
      red_add_f32(dw + (long long)(c + 1) * K * K + ti * K + j, a1);
    }
  }
}

}  // namespace vmlp

// ============================================================================================ patch embedding (stem)
// Conv2d(Cin, C, kernel = stride = P) == GEMM over non-overlapping patches (mlp_mixer.py:58-60,68-71):
//   rows[(b, ph, pw), (ci, i, j)] = x[b, ci, ph*P + i, pw*P + j]        (x NCHW; P % 8 == 0 so each j-run is 16-byte vectors)
// FWD = 1 gathers the patch rows; FWD = 0 is the exact inverse (scatter of d(rows) back to d(x), one-to-one).
namespace vmlp {
template <int FWD>
__global__ void __launch_bounds__(RW_THREADS)
patchify_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int B, int Cin, int H, int W, int P) {
  const int nph = H / P, npw = W / P, vp = P / 8;
  const int per = nph * npw * Cin * P * vp;
  const FastDiv d1(vp), d2(P), d3(Cin), d4(npw);
  const long long b = blockIdx.y;
  (void)B;
  for (int i0 = blockIdx.x * RW_THREADS + threadIdx.x; i0 < per; i0 += gridDim.x * RW_THREADS) {
    int t, jv, t2, i, t3, ci, ph, pw;
    d1.divmod(i0, t, jv);
    d2.divmod(t, t2, i);
    d3.divmod(t2, t3, ci);
    d4.divmod(t3, ph, pw);
    const long long idx = b * per + i0;
    const long long xoff = ((b * Cin + ci) * H + ph * P + i) * W + pw * P + jv * 8;
    if (FWD) *reinterpret_cast<uint4*>(dst + idx * 8) = ldg_nc_v4(src + xoff);
    else *reinterpret_cast<uint4*>(dst + xoff) = ldg_nc_v4(src + idx * 8);
  }
}
}  // namespace vmlp
