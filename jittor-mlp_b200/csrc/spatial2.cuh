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
// Block = 8x8 output pixels x 64 channels; thread = (channel, 2 output rows); each tap row is slid over a register
// window so every shared-memory load feeds K FMAs.  FLIP = 1 evaluates the same stencil with the kernel rotated by
// 180 degrees, which is the input gradient.  EPI = 1 adds the bias and writes gelu'(z) (for backward) and gelu(z).
constexpr int DW_TILE = 8;
constexpr int DW_CH = 64;
template <int K>
struct DwSmem {
  static constexpr int IN_W = DW_TILE + K - 1;
  static constexpr int IN_ELEMS = IN_W * IN_W * DW_CH;          // bf16
  static constexpr int W_ELEMS = K * K * DW_CH;                 // float
  static constexpr int BYTES = IN_ELEMS * 2 + W_ELEMS * 4 + DW_TILE * DW_TILE * DW_CH * 2;
};

template <int K>
__device__ __forceinline__ void dw_load_tile(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* s_in, long long b, int h0,
                                             int w0, int c0, int H, int W, int C) {
  constexpr int IN_W = DW_TILE + K - 1, P = K / 2;
  for (int v = threadIdx.x; v < IN_W * IN_W * (DW_CH / 8); v += blockDim.x) {
    const int cv = v % (DW_CH / 8), pos = v / (DW_CH / 8);
    const int hh = h0 - P + pos / IN_W, ww = w0 - P + pos % IN_W;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (hh >= 0 && hh < H && ww >= 0 && ww < W && c0 + cv * 8 < C)
      val = ldg_nc_v4(x + ((b * H + hh) * W + ww) * C + c0 + cv * 8);
    *reinterpret_cast<uint4*>(s_in + pos * DW_CH + cv * 8) = val;
  }
}

template <int K, int FLIP, int EPI>
__global__ void __launch_bounds__(256)
dwconv_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ wgt,
              const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, __nv_bfloat16* __restrict__ out2,
              int B, int H, int W, int C) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  constexpr int IN_W = DW_TILE + K - 1;
  __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(dw_smem);
  float* s_w = reinterpret_cast<float*>(dw_smem + DwSmem<K>::IN_ELEMS * 2);
  const int tiles_w = (W + DW_TILE - 1) / DW_TILE, tiles_h = (H + DW_TILE - 1) / DW_TILE;
  const int c0 = blockIdx.y * DW_CH;
  const int tx = threadIdx.x % DW_CH, ty = threadIdx.x / DW_CH;
  const int c = c0 + tx;
  for (int i = threadIdx.x; i < K * K * DW_CH; i += blockDim.x) {
    const int cc = i % DW_CH, tap = i / DW_CH;
    const int src = FLIP ? (K * K - 1 - tap) : tap;
    s_w[i] = (c0 + cc < C) ? __bfloat162float(wgt[(long long)(c0 + cc) * K * K + src]) : 0.f;
  }
  const float bv = (EPI && c < C) ? __bfloat162float(bias[c]) : 0.f;
  for (long long tile = blockIdx.x; tile < (long long)B * tiles_h * tiles_w; tile += gridDim.x) {
    const long long b = tile / (tiles_h * tiles_w);
    const int h0 = static_cast<int>((tile / tiles_w) % tiles_h) * DW_TILE, w0 = static_cast<int>(tile % tiles_w) * DW_TILE;
    __syncthreads();
    dw_load_tile<K>(x, s_in, b, h0, w0, c0, H, W, C);
    __syncthreads();
    float acc[2][DW_TILE];
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int q = 0; q < DW_TILE; ++q) acc[r][q] = bv;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int orow = ty * 2 + r;
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float win[IN_W];
#pragma unroll
        for (int q = 0; q < IN_W; ++q) win[q] = __bfloat162float(s_in[((orow + i) * IN_W + q) * DW_CH + tx]);
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const float wv = s_w[(i * K + j) * DW_CH + tx];
#pragma unroll
          for (int q = 0; q < DW_TILE; ++q) acc[r][q] = fmaf(wv, win[q + j], acc[r][q]);
        }
      }
    }
    if (c < C) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int hh = h0 + ty * 2 + r;
        if (hh >= H) continue;
#pragma unroll
        for (int q = 0; q < DW_TILE; ++q) {
          const int ww = w0 + q;
          if (ww >= W) continue;
          const long long o = ((b * H + hh) * W + ww) * C + c;
          if (EPI) {            // out = gelu'(z) (kept for backward), out2 = gelu(z)
            float d;
            const float gz = gelu_erf_t<true>(acc[r][q], d);
            out[o] = __float2bfloat16(d);
            out2[o] = __float2bfloat16(gz);
          } else {
            out[o] = __float2bfloat16(acc[r][q]);
          }
        }
      }
    }
  }
}

// dW[c][i][j] += sum_{b,h,w} dz[b,h,w,c] * x[b, h+i-P, w+j-P, c]   (fp32 atomics once per block)
// thread = (channel, tap rows i == ty mod 4): up to ceil(K/4) x K accumulators live in registers across all tiles.
template <int K>
__global__ void __launch_bounds__(256)
dwconv_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dz, float* __restrict__ dw,
                    int B, int H, int W, int C) {
  extern __shared__ __align__(16) uint8_t dw_smem[];
  constexpr int IN_W = DW_TILE + K - 1, TR = (K + 3) / 4;
  __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(dw_smem);
  __nv_bfloat16* s_dz = reinterpret_cast<__nv_bfloat16*>(dw_smem + DwSmem<K>::IN_ELEMS * 2 + DwSmem<K>::W_ELEMS * 4);
  const int tiles_w = (W + DW_TILE - 1) / DW_TILE, tiles_h = (H + DW_TILE - 1) / DW_TILE;
  const int c0 = blockIdx.y * DW_CH;
  const int tx = threadIdx.x % DW_CH, ty = threadIdx.x / DW_CH;
  float acc[TR][K];
#pragma unroll
  for (int a = 0; a < TR; ++a)
#pragma unroll
    for (int j = 0; j < K; ++j) acc[a][j] = 0.f;
  for (long long tile = blockIdx.x; tile < (long long)B * tiles_h * tiles_w; tile += gridDim.x) {
    const long long b = tile / (tiles_h * tiles_w);
    const int h0 = static_cast<int>((tile / tiles_w) % tiles_h) * DW_TILE, w0 = static_cast<int>(tile % tiles_w) * DW_TILE;
    __syncthreads();
    dw_load_tile<K>(x, s_in, b, h0, w0, c0, H, W, C);
    for (int v = threadIdx.x; v < DW_TILE * DW_TILE * (DW_CH / 8); v += blockDim.x) {
      const int cv = v % (DW_CH / 8), pos = v / (DW_CH / 8);
      const int hh = h0 + pos / DW_TILE, ww = w0 + pos % DW_TILE;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (hh < H && ww < W && c0 + cv * 8 < C) val = ldg_nc_v4(dz + ((b * H + hh) * W + ww) * C + c0 + cv * 8);
      *reinterpret_cast<uint4*>(s_dz + pos * DW_CH + cv * 8) = val;
    }
    __syncthreads();
#pragma unroll
    for (int a = 0; a < TR; ++a) {
      const int i = ty + 4 * a;
      if (i < K) {
#pragma unroll
        for (int r = 0; r < DW_TILE; ++r) {
          float g[DW_TILE], win[IN_W];
#pragma unroll
          for (int q = 0; q < DW_TILE; ++q) g[q] = __bfloat162float(s_dz[(r * DW_TILE + q) * DW_CH + tx]);
#pragma unroll
          for (int q = 0; q < IN_W; ++q) win[q] = __bfloat162float(s_in[((r + i) * IN_W + q) * DW_CH + tx]);
#pragma unroll
          for (int j = 0; j < K; ++j)
#pragma unroll
            for (int q = 0; q < DW_TILE; ++q) acc[a][j] = fmaf(g[q], win[q + j], acc[a][j]);
        }
      }
    }
  }
  if (c0 + tx < C) {
#pragma unroll
    for (int a = 0; a < TR; ++a) {
      const int i = ty + 4 * a;
      if (i < K) {
#pragma unroll
        for (int j = 0; j < K; ++j) red_add_f32(dw + (long long)(c0 + tx) * K * K + i * K + j, acc[a][j]);
      }
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
