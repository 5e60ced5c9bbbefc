// Spatial (NHWC) kernels of the shift family: channel-group token shifts as load-time index offsets,
// GroupNorm(1, C) whole-sample statistics, per-channel affine with fp32 coefficients (BatchNorm apply /
// backward), the S2-MLPv2 split-attention reduce / combine, Hire-MLP region gather, and ConvMixer's
// depthwise stencil.  All tensors are channels-last [B, H, W, C] bf16 (C % 8 == 0); every global access is a
// 16-byte vector of 8 channels; a vector that straddles two channel groups falls back to per-element offsets.
#pragma once
#include "rowwise.cuh"

namespace vmlp {

// Channel group g covers channels [start[g], start[g+1]) and reads its input at (h + dh[g], w + dw[g]).
struct ShiftTable {
  int ngroups;
  int start[9];
  int dh[8];
  int dw[8];
};

// MODE 0: zero padding  out[h, w] = in[h + dh, w + dw] (0 outside)        -- AS-MLP Shift (shift_cuda.py:44-72);
//                        its adjoint is the same gather with negated offsets (shift_cuda.py:75-103).
// MODE 1: clamp-to-edge out[h, w] = in[clamp(h + dh), clamp(w + dw)]      -- S2-MLP spatial_shift, intended semantics
//                        (s2_mlp_v1.py:19-25, s2_mlp_v2.py:15-29).
// MODE 2: adjoint of MODE 1 for |dh|, |dw| <= 1: gin[j] = gout[j - d] (if inside) + [j is the clamped edge] gout[j].
template <int MODE>
__device__ __forceinline__ float shift_fetch(const __nv_bfloat16* __restrict__ in, long long img_base, int h, int w,
                                             int c, int H, int W, int C, int dh, int dw) {
  if (MODE == 0) {
    const int hs = h + dh, ws = w + dw;
    if (hs < 0 || hs >= H || ws < 0 || ws >= W) return 0.f;
    return __bfloat162float(in[img_base + ((long long)hs * W + ws) * C + c]);
  } else if (MODE == 1) {
    const int hs = min(max(h + dh, 0), H - 1), ws = min(max(w + dw, 0), W - 1);
    return __bfloat162float(in[img_base + ((long long)hs * W + ws) * C + c]);
  } else {
    // forward was out[i] = in[clamp(i + d)]; adjoint source for position j: i = j - d, plus the edge self term
    float v = 0.f;
    const int hs = h - dh, ws = w - dw;
    if (hs >= 0 && hs < H && ws >= 0 && ws < W) v = __bfloat162float(in[img_base + ((long long)hs * W + ws) * C + c]);
    const bool edge = (dh < 0 && h == 0) || (dh > 0 && h == H - 1) || (dw < 0 && w == 0) || (dw > 0 && w == W - 1);
    if (edge) v += __bfloat162float(in[img_base + ((long long)h * W + w) * C + c]);
    // the opposite edge is read by nobody in the forward pass except through the (already counted) shifted term
    return v;
  }
}

// One source vector of channel group g for output position (h, w): MODE 0 zero-fills outside, MODE 1 clamps,
// MODE 2 is the clamp adjoint (shifted neighbour, plus the position itself on the clamped edge).
template <int MODE>
__device__ __forceinline__ void shift_vec8(const __nv_bfloat16* __restrict__ in, long long img, int h, int w, int c0,
                                           int H, int W, int C, int dh, int dw, float (&o)[8]) {
#pragma unroll
  for (int e = 0; e < 8; ++e) o[e] = 0.f;
  if (MODE == 0) {
    const int hs = h + dh, ws = w + dw;
    // L1-allocating load: neighbouring positions re-read the other half of each 32-byte sector (group width 40 B)
    if (hs >= 0 && hs < H && ws >= 0 && ws < W) unpack8(ldg_v4(in + img + ((long long)hs * W + ws) * C + c0), o);
  } else if (MODE == 1) {
    const int hs = min(max(h + dh, 0), H - 1), ws = min(max(w + dw, 0), W - 1);
    unpack8(ldg_v4(in + img + ((long long)hs * W + ws) * C + c0), o);
  } else {
    const int hs = h - dh, ws = w - dw;
    if (hs >= 0 && hs < H && ws >= 0 && ws < W) unpack8(ldg_v4(in + img + ((long long)hs * W + ws) * C + c0), o);
    const bool edge = (dh < 0 && h == 0) || (dh > 0 && h == H - 1) || (dw < 0 && w == 0) || (dw > 0 && w == W - 1);
    if (edge) {
      float s[8];
      unpack8(ldg_v4(in + img + ((long long)h * W + w) * C + c0), s);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += s[e];
    }
  }
}

// A 16-byte vector inside one channel group moves as a whole; a vector straddling two ADJACENT groups (AS-MLP: groups of
// ceil(C / 5) = 20 channels, 2 of every 12 vectors) is stitched from the two groups' source vectors.  The per-element
// path (8 scalar loads + 8 group searches, ~200 instructions, taken by every warp) made the first version
// instruction-bound at 3.6x the HBM floor; it is left for vectors that span three or more groups.
template <int MODE>
__global__ void __launch_bounds__(RW_THREADS)
shift_nhwc_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int B, int H, int W, int C,
                  const ShiftTable tab) {
  __shared__ int s_start[9], s_dh[8], s_dw[8];
  if (threadIdx.x < 9) s_start[threadIdx.x] = tab.start[threadIdx.x];
  if (threadIdx.x < 8) { s_dh[threadIdx.x] = tab.dh[threadIdx.x]; s_dw[threadIdx.x] = tab.dw[threadIdx.x]; }
  __syncthreads();
  const int ng = tab.ngroups;
  auto group_of = [&](int c) {
    int g = 0;
    for (int i = 1; i < ng; ++i)
      if (c >= s_start[i]) g = i;
    return g;
  };
  const int nvec = C >> 3;
  const FastDiv dv(nvec), dw_(W);
  const long long b = blockIdx.y;
  const long long img = b * (long long)H * W * C;
  const int per_img = H * W * nvec;
  (void)B;
  for (int i = blockIdx.x * RW_THREADS + threadIdx.x; i < per_img; i += gridDim.x * RW_THREADS) {
    int pos, cv, h, w;
    dv.divmod(i, pos, cv);
    dw_.divmod(pos, h, w);
    const int c0 = cv * 8;
    const int g0 = group_of(c0), g1 = group_of(c0 + 7);
    __nv_bfloat16* dst = out + img + ((long long)h * W + w) * C + c0;
    float o[8];
    if (g0 == g1) {
      if (MODE == 0 || MODE == 1) {
        // pure copy: no unpack / repack
        int hs = h + s_dh[g0], ws = w + s_dw[g0];
        bool inside = true;
        if (MODE == 0) inside = (hs >= 0 && hs < H && ws >= 0 && ws < W);
        else { hs = min(max(hs, 0), H - 1); ws = min(max(ws, 0), W - 1); }
        *reinterpret_cast<uint4*>(dst) = inside ? ldg_v4(in + img + ((long long)hs * W + ws) * C + c0) : make_uint4(0, 0, 0, 0);
      } else {
        shift_vec8<MODE>(in, img, h, w, c0, H, W, C, s_dh[g0], s_dw[g0], o);
        *reinterpret_cast<uint4*>(dst) = pack8(o);
      }
    } else if (g1 == g0 + 1) {
      float o1[8];
      shift_vec8<MODE>(in, img, h, w, c0, H, W, C, s_dh[g0], s_dw[g0], o);
      shift_vec8<MODE>(in, img, h, w, c0, H, W, C, s_dh[g1], s_dw[g1], o1);
      const int split = s_start[g1] - c0;     // first element of the vector that belongs to group g1 (1..7)
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (e < split) ? o[e] : o1[e];
      *reinterpret_cast<uint4*>(dst) = pack8(o);
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int g = group_of(c0 + e);
        o[e] = shift_fetch<MODE>(in, img, h, w, c0 + e, H, W, C, s_dh[g], s_dw[g]);
      }
      *reinterpret_cast<uint4*>(dst) = pack8(o);
    }
  }
}

// --------------------------------------------------------------------------- GroupNorm(1, C): whole-sample statistics
// acc[b] = (sum, sum of squares) over the sample's P*C contiguous elements (fp32 red.add; caller zero-fills).
// Reference: MyNorm = nn.GroupNorm(1, dim), as_mlp.py:343-344 (biased variance, eps 1e-5).
__global__ void __launch_bounds__(RW_THREADS)
gn_stats_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ acc, long long per_sample_vec) {
  __shared__ float sh[2][RW_WARPS];
  const long long b = blockIdx.y;
  const __nv_bfloat16* xs = x + b * per_sample_vec * 8;
  float s = 0.f, ss = 0.f;
  for (long long i = (long long)blockIdx.x * RW_THREADS + threadIdx.x; i < per_sample_vec;
       i += (long long)gridDim.x * RW_THREADS) {
    float v[8];
    unpack8(ldg_nc_v4(xs + i * 8), v);
#pragma unroll
    for (int e = 0; e < 8; ++e) { s += v[e]; ss += v[e] * v[e]; }
  }
  s = warp_sum(s);
  ss = warp_sum(ss);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int wv = 0; wv < RW_WARPS; ++wv) { a += sh[0][wv]; c += sh[1][wv]; }
    red_add_f32(acc + 2 * b, a);
    red_add_f32(acc + 2 * b + 1, c);
  }
}
__device__ __forceinline__ void gn_mean_rstd(const float* acc, long long b, float n, float eps, float& mean, float& rstd) {
  mean = acc[2 * b] / n;
  const float var = fmaxf(acc[2 * b + 1] / n - mean * mean, 0.f);
  rstd = rsqrtf(var + eps);
}
// y = gn(x) (GELU == 0) or gelu(gn(x)) (GELU == 1)
template <int GELU>
__global__ void __launch_bounds__(RW_THREADS)
gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ acc,
                const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                __nv_bfloat16* __restrict__ y, long long per_sample_vec, int C, float eps, long long total_vec) {
  const int nvec = C >> 3;
  const FastDiv dv(nvec);
  const float n = static_cast<float>(per_sample_vec * 8);
  const long long b = blockIdx.y;
  float mean, rstd;
  gn_mean_rstd(acc, b, n, eps, mean, rstd);
  (void)total_vec;
  for (int j = blockIdx.x * RW_THREADS + threadIdx.x; j < per_sample_vec; j += gridDim.x * RW_THREADS) {
    int pos, cv;
    dv.divmod(j, pos, cv);
    const int c0 = cv * 8;
    const long long i = b * per_sample_vec + j;
    float v[8], g[8], bt[8], o[8];
    unpack8(ldg_nc_v4(x + i * 8), v);
    unpack8(*reinterpret_cast<const uint4*>(gamma + c0), g);
    unpack8(*reinterpret_cast<const uint4*>(beta + c0), bt);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float nv = (v[e] - mean) * rstd * g[e] + bt[e];
      o[e] = GELU ? gelu_erf(nv) : nv;
    }
    *reinterpret_cast<uint4*>(y + i * 8) = pack8(o);
  }
}
// backward pass A: dn = dy (* gelu'(n)),  n = xhat * gamma + beta  (recomputed from x);
//   per sample:  acc2[b] += (sum g, sum g * xhat),  g = dn * gamma
//   per channel: dgamma += sum dn * xhat, dbeta += sum dn
// Thread t of a block owns channel vector (t % nvec) of position (t / nvec) and strides over positions, so its 16
// per-channel partial sums live in registers for the whole kernel and reach shared memory once (the first version did
// two shared-memory atomics per ELEMENT on 2C addresses: 220 us for the 308 MB of AS-MLP-T stage 0, 21 % of the step).
template <int GELU>
__global__ void __launch_bounds__(RW_THREADS)
gn_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                     const float* __restrict__ acc, const __nv_bfloat16* __restrict__ gamma,
                     const __nv_bfloat16* __restrict__ beta, __nv_bfloat16* __restrict__ dn, float* __restrict__ acc2,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long per_sample_vec, int C, float eps) {
  extern __shared__ float shc[];          // [2][C] column partials
  __shared__ float shs[2][RW_WARPS];
  const int nvec = C >> 3;                // host guarantees nvec <= RW_THREADS
  for (int i = threadIdx.x; i < 2 * C; i += RW_THREADS) shc[i] = 0.f;
  __syncthreads();
  const long long b = blockIdx.y;
  const float n = static_cast<float>(per_sample_vec * 8);
  float mean, rstd;
  gn_mean_rstd(acc, b, n, eps, mean, rstd);
  const long long base = b * per_sample_vec;
  const int ppb = RW_THREADS / nvec;      // positions per block iteration
  const int cv = threadIdx.x % nvec, po = threadIdx.x / nvec;
  const bool active = po < ppb;
  const int c0 = cv * 8;
  const int P = static_cast<int>(per_sample_vec / nvec);
  float s1 = 0.f, s2 = 0.f;
  float ag[8], ab[8], g[8], bt[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) { ag[e] = 0.f; ab[e] = 0.f; bt[e] = 0.f; }
  unpack8(*reinterpret_cast<const uint4*>(gamma + c0), g);
  if (GELU) unpack8(*reinterpret_cast<const uint4*>(beta + c0), bt);
  if (active) {
    for (int pos = blockIdx.x * ppb + po; pos < P; pos += gridDim.x * ppb) {
      const long long i = base + (long long)pos * nvec + cv;
      float d[8], v[8], o[8];
      unpack8(ldg_nc_v4(dy + i * 8), d);
      unpack8(ldg_nc_v4(x + i * 8), v);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float xh = (v[e] - mean) * rstd;
        float dv = d[e];
        if (GELU) dv *= dgelu_erf(xh * g[e] + bt[e]);
        o[e] = dv;
        const float gg = dv * g[e];
        s1 += gg;
        s2 += gg * xh;
        ag[e] += dv * xh;
        ab[e] += dv;
      }
      *reinterpret_cast<uint4*>(dn + i * 8) = pack8(o);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      atomicAdd(&shc[c0 + e], ag[e]);
      atomicAdd(&shc[C + c0 + e], ab[e]);
    }
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { shs[0][warp] = s1; shs[1][warp] = s2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
#pragma unroll
    for (int wv = 0; wv < RW_WARPS; ++wv) { a += shs[0][wv]; c += shs[1][wv]; }
    red_add_f32(acc2 + 2 * b, a);
    red_add_f32(acc2 + 2 * b + 1, c);
  }
  for (int i = threadIdx.x; i < C; i += RW_THREADS) {
    red_add_f32(dgamma + i, shc[i]);
    red_add_f32(dbeta + i, shc[C + i]);
  }
}
// backward pass B: dx = rstd * (dn * gamma - S1/n - xhat * S2/n)
__global__ void __launch_bounds__(RW_THREADS)
gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ dn, const __nv_bfloat16* __restrict__ x,
                    const float* __restrict__ acc, const float* __restrict__ acc2,
                    const __nv_bfloat16* __restrict__ gamma, __nv_bfloat16* __restrict__ dx, long long per_sample_vec,
                    int C, float eps, long long total_vec) {
  const int nvec = C >> 3;
  const FastDiv dv(nvec);
  const float n = static_cast<float>(per_sample_vec * 8);
  const long long b = blockIdx.y;
  float mean, rstd;
  gn_mean_rstd(acc, b, n, eps, mean, rstd);
  const float m1 = acc2[2 * b] / n, m2 = acc2[2 * b + 1] / n;
  (void)total_vec;
  for (int j = blockIdx.x * RW_THREADS + threadIdx.x; j < per_sample_vec; j += gridDim.x * RW_THREADS) {
    int pos, cv;
    dv.divmod(j, pos, cv);
    const int c0 = cv * 8;
    const long long i = b * per_sample_vec + j;
    float d[8], v[8], g[8], o[8];
    unpack8(ldg_nc_v4(dn + i * 8), d);
    unpack8(ldg_nc_v4(x + i * 8), v);
    unpack8(*reinterpret_cast<const uint4*>(gamma + c0), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float xh = (v[e] - mean) * rstd;
      o[e] = rstd * (d[e] * g[e] - m1 - xh * m2);
    }
    *reinterpret_cast<uint4*>(dx + i * 8) = pack8(o);
  }
}

// --------------------------------------------------------------------------- per-channel linear combination (fp32 coefficients)
//   out = (A[c] * p + Bq[c] * q + Cc[c]) (* gelu'(z))
// BatchNorm apply (+ residual: q = x, Bq = 1), BatchNorm backward (p = dy, q = activation), optionally fused with
// the backward of the GELU that precedes the BatchNorm (conv_mixer.py:23-32).
template <int HAS_Q, int DGELU>
__global__ void __launch_bounds__(RW_THREADS)
chan_lin_kernel(const __nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ q,
                const __nv_bfloat16* __restrict__ z, const float* __restrict__ A, const float* __restrict__ Bq,
                const float* __restrict__ Cc, __nv_bfloat16* __restrict__ out, long long total_vec, int C) {
  // the host sizes the grid so that (gridDim.x * RW_THREADS) % (C / 8) == 0: every thread then stays on ONE channel
  // vector for its whole grid-stride loop and keeps its 8 (x3) fp32 coefficients in registers
  const int nvec = C >> 3;
  const long long start = (long long)blockIdx.x * RW_THREADS + threadIdx.x;
  const int c0 = static_cast<int>(start % nvec) * 8;
  float a[8], bq[8], cc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    a[e] = A[c0 + e];
    cc[e] = Cc[c0 + e];
    bq[e] = HAS_Q ? Bq[c0 + e] : 0.f;
  }
  for (long long i = start; i < total_vec; i += (long long)gridDim.x * RW_THREADS) {
    float pv[8], qv[8], zv[8], o[8];
    unpack8(ldg_nc_v4(p + i * 8), pv);
    if (HAS_Q) unpack8(ldg_nc_v4(q + i * 8), qv);
    if (DGELU) unpack8(ldg_nc_v4(z + i * 8), zv);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = fmaf(a[e], pv[e], cc[e]);
      if (HAS_Q) v = fmaf(bq[e], qv[e], v);
      if (DGELU) v *= zv[e];        // z holds the saved gelu'(pre-activation)
      o[e] = v;
    }
    *reinterpret_cast<uint4*>(out + i * 8) = pack8(o);
  }
}

// BatchNorm2d (training) coefficient kernel, one thread per channel (conv_mixer.py:20,27,31):
//   forward : mean = s1/R, var = s2/R - mean^2 (biased); A = gamma*rstd, C = beta - mean*A;
//             running_mean/var momentum update with the UNBIASED variance; save mean, rstd.
//   backward: given sum(dy) and sum(dy*a):  dgamma = rstd*(sum(dy*a) - mean*sum(dy)), dbeta = sum(dy)
//             dx = A*dy + B*a + C with A = gamma*rstd, B = -gamma*rstd^2*dgamma/R, C = -A*dbeta/R - B*mean
__global__ void bn_fwd_coef_kernel(const float* __restrict__ s1, const float* __restrict__ s2,
                                   const __nv_bfloat16* __restrict__ gamma, const __nv_bfloat16* __restrict__ beta,
                                   float* __restrict__ A, float* __restrict__ Cc, float* __restrict__ mean_out,
                                   float* __restrict__ rstd_out, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float R, float eps, float momentum, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = s1[c] / R;
  const float var = fmaxf(s2[c] / R - mean * mean, 0.f);
  const float rstd = rsqrtf(var + eps);
  const float a = __bfloat162float(gamma[c]) * rstd;
  A[c] = a;
  Cc[c] = __bfloat162float(beta[c]) - mean * a;
  mean_out[c] = mean;
  rstd_out[c] = rstd;
  if (running_mean) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (R / fmaxf(R - 1.f, 1.f));
  }
}
__global__ void bn_bwd_coef_kernel(const float* __restrict__ sdy, const float* __restrict__ sdya,
                                   const __nv_bfloat16* __restrict__ gamma, const float* __restrict__ mean,
                                   const float* __restrict__ rstd, float* __restrict__ A, float* __restrict__ Bq,
                                   float* __restrict__ Cc, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                   float R, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float g = __bfloat162float(gamma[c]);
  const float dg = rstd[c] * (sdya[c] - mean[c] * sdy[c]);
  const float a = g * rstd[c];
  const float b = -g * rstd[c] * rstd[c] * dg / R;
  A[c] = a;
  Bq[c] = b;
  Cc[c] = -a * sdy[c] / R - b * mean[c];
  dgamma[c] += dg;
  dbeta[c] += sdy[c];
}

// --------------------------------------------------------------------------- S2-MLPv2 split attention (s2_mlp_v2.py:31-69)
// t: [B, H, W, 3C] = mlp1 output;  x_k = shift_k(t[..., kC:(k+1)C]) for k = 0 (plan 1), 1 (plan 2), 2 (identity).
// The shifts are never materialised: every kernel below reads t at clamp(position + offset_k(channel quarter)).
// `plain` != 0: no shifts at all -- the same SplitAttention as Vision Permutator uses it on its stacked H / W / C branch
// outputs (vip.py:37-57), which the branch kernels write side by side into one [B, H, W, 3C] buffer.
__device__ __forceinline__ void s2_plan_offset(int k, int quarter, int& dh, int& dw) {
  // `x[:,1:] = x[:,:-1]` => out[i] = in[i-1] => offset -1 (s2_mlp_v2.py:15-29)
  // plan 1: quarters (0, 1) move along h by (-1, +1), quarters (2, 3) along w; plan 2 swaps the axes; k = 2 is identity
  const int sgn = (quarter & 1) ? 1 : -1;
  const bool on_h = ((quarter >> 1) == 0) != (k == 1);
  dh = (k < 2 && on_h) ? sgn : 0;
  dw = (k < 2 && !on_h) ? sgn : 0;
}
// value of x_k[b, h, w, c0 .. c0+7] (forward gather, clamp) as 8 floats
__device__ __forceinline__ void s2_gather8(const __nv_bfloat16* __restrict__ t, long long img, int h, int w, int k,
                                           int c0, int H, int W, int C, float (&o)[8], int plain = 0) {
  const int qs = C >> 2;
  const int q0 = min(c0 / qs, 3), q1 = min((c0 + 7) / qs, 3);
  if (q0 == q1) {
    int dh, dw;
    s2_plan_offset(plain ? 2 : k, q0, dh, dw);
    const int hs = min(max(h + dh, 0), H - 1), ws = min(max(w + dw, 0), W - 1);
    unpack8(ldg_nc_v4(t + img + ((long long)hs * W + ws) * 3 * C + k * C + c0), o);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int dh, dw;
      s2_plan_offset(plain ? 2 : k, min((c0 + e) / qs, 3), dh, dw);
      const int hs = min(max(h + dh, 0), H - 1), ws = min(max(w + dw, 0), W - 1);
      o[e] = __bfloat162float(t[img + ((long long)hs * W + ws) * 3 * C + k * C + c0 + e]);
    }
  }
}
// adjoint gather of a [B, H, W, C] field g for branch k: sum of g over the output positions that read (h, w).
// A vector that lies inside one channel quarter (always, when C/4 is a multiple of 8) costs at most two 16-byte loads:
// the shifted neighbour and, on the clamped edge, the position itself.
__device__ __forceinline__ void s2_adjoint8(const __nv_bfloat16* __restrict__ g, long long img, int h, int w, int k,
                                            int c0, int H, int W, int C, float (&o)[8], int plain = 0) {
  const int qs = C >> 2;
  const int q0 = min(c0 / qs, 3), q1 = min((c0 + 7) / qs, 3);
  if (q0 == q1) {
    int dh, dw;
    s2_plan_offset(plain ? 2 : k, q0, dh, dw);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
    const int hs = h - dh, ws = w - dw;
    if (hs >= 0 && hs < H && ws >= 0 && ws < W) unpack8(ldg_nc_v4(g + img + ((long long)hs * W + ws) * C + c0), o);
    const bool edge = (dh < 0 && h == 0) || (dh > 0 && h == H - 1) || (dw < 0 && w == 0) || (dw > 0 && w == W - 1);
    if (edge) {
      float s[8];
      unpack8(ldg_nc_v4(g + img + ((long long)h * W + w) * C + c0), s);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += s[e];
    }
    return;
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    int dh, dw;
    s2_plan_offset(plain ? 2 : k, min((c0 + e) / qs, 3), dh, dw);
    o[e] = shift_fetch<2>(g, img, h, w, c0 + e, H, W, C, dh, dw);
  }
}
// number of output positions that read input position (h, w) under offset (dh, dw) with clamping (0, 1 or 2)
__device__ __forceinline__ float s2_read_count(int h, int w, int H, int W, int dh, int dw) {
  float n = 0.f;
  const int hs = h - dh, ws = w - dw;
  if (hs >= 0 && hs < H && ws >= 0 && ws < W) n += 1.f;
  if ((dh < 0 && h == 0) || (dh > 0 && h == H - 1) || (dw < 0 && w == 0) || (dw > 0 && w == W - 1)) n += 1.f;
  return n;
}

// MODE 0: out[b, c]    += sum_pos (x_0 + x_1 + x_2)                 (a of SplitAttention, s2_mlp_v2.py:44)
// MODE 1: out[b, k, c] += sum_pos g[b, pos, c] * x_k[b, pos, c]       (d(bar_a) in backward)
template <int MODE>
__global__ void __launch_bounds__(RW_THREADS)
s2v2_reduce_kernel(const __nv_bfloat16* __restrict__ t, const __nv_bfloat16* __restrict__ g, float* __restrict__ out,
                   int H, int W, int C, int plain) {
  extern __shared__ float sh[];     // [plane][nvec*8*(MODE?3:1)]
  const int nvec = C >> 3;
  const int plane = RW_THREADS / nvec;
  const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
  const int b = blockIdx.y;
  const long long img = (long long)b * H * W * 3 * C, gimg = (long long)b * H * W * C;
  constexpr int NK = MODE ? 3 : 1;
  float acc[NK][8];
#pragma unroll
  for (int k = 0; k < NK; ++k)
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[k][e] = 0.f;
  if (pl < plane) {
    const FastDiv dw_(W);
#pragma unroll 2
    for (int pos = blockIdx.x * plane + pl; pos < H * W; pos += gridDim.x * plane) {
      int h, w;
      dw_.divmod(pos, h, w);
      float gv[8];
      if (MODE) unpack8(ldg_nc_v4(g + gimg + (long long)pos * C + v * 8), gv);
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        float xv[8];
        s2_gather8(t, img, h, w, k, v * 8, H, W, C, xv, plain);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          if (MODE) acc[k][e] += gv[e] * xv[e];
          else acc[0][e] += xv[e];
        }
      }
    }
  }
  const int width = nvec * 8 * NK;
  if (pl < plane) {
#pragma unroll
    for (int k = 0; k < NK; ++k)
#pragma unroll
      for (int e = 0; e < 8; ++e) sh[pl * width + k * nvec * 8 + v * 8 + e] = acc[k][e];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < width; i += RW_THREADS) {
    float s = 0.f;
    for (int p2 = 0; p2 < plane; ++p2) s += sh[p2 * width + i];
    red_add_f32(out + (long long)b * width + i, s);     // MODE 1 layout [b][k][c] == [b][k*C + c]
  }
}
__device__ __forceinline__ void s2_softmax3(const __nv_bfloat16* __restrict__ hat, long long b, int c0, int C,
                                            float (&bar)[3][8]) {
  float l[3][8];
#pragma unroll
  for (int k = 0; k < 3; ++k) unpack8(*reinterpret_cast<const uint4*>(hat + b * 3 * C + k * C + c0), l[k]);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float m = fmaxf(l[0][e], fmaxf(l[1][e], l[2][e]));
    const float e0 = __expf(l[0][e] - m), e1 = __expf(l[1][e] - m), e2 = __expf(l[2][e] - m);
    const float inv = 1.f / (e0 + e1 + e2);
    bar[0][e] = e0 * inv; bar[1][e] = e1 * inv; bar[2][e] = e2 * inv;
  }
}
// out[b, pos, c] = sum_k softmax_k(hat[b, :, c]) * x_k[b, pos, c]      (s2_mlp_v2.py:45-51)
// Thread t owns channel vector (t % nvec) and walks positions (t / nvec) + k * plane: the softmax over k of its 8
// channels (24 exp) is evaluated once per thread, not once per position.
__global__ void __launch_bounds__(RW_THREADS)
s2v2_combine_kernel(const __nv_bfloat16* __restrict__ t, const __nv_bfloat16* __restrict__ hat,
                    __nv_bfloat16* __restrict__ out, int B, int H, int W, int C, int plain) {
  const int nvec = C >> 3;
  const int plane = RW_THREADS / nvec;
  const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
  const FastDiv dw_(W);
  const long long b = blockIdx.y;
  (void)B;
  if (pl >= plane) return;
  const int c0 = v * 8;
  float bar[3][8];
  s2_softmax3(hat, b, c0, C, bar);
  const long long img = b * H * W * 3 * C;
#pragma unroll 2
  for (int pos = blockIdx.x * plane + pl; pos < H * W; pos += gridDim.x * plane) {
    int h, w;
    dw_.divmod(pos, h, w);
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = 0.f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float xv[8];
      s2_gather8(t, img, h, w, k, c0, H, W, C, xv, plain);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += bar[k][e] * xv[e];
    }
    *reinterpret_cast<uint4*>(out + (b * H * W + pos) * C + c0) = pack8(o);
  }
}
// dhat[b, k, c] = bar_k * (dbar_k - sum_j bar_j dbar_j)                 (softmax over k backward)
__global__ void s2v2_softmax_bwd_kernel(const __nv_bfloat16* __restrict__ hat, const float* __restrict__ dbar,
                                        __nv_bfloat16* __restrict__ dhat, int B, int C) {
  const int nvec = C >> 3;
  const long long total = (long long)B * nvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / nvec;
    const int c0 = static_cast<int>(i % nvec) * 8;
    float bar[3][8], o[3][8];
    s2_softmax3(hat, b, c0, C, bar);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d0 = dbar[b * 3 * C + c0 + e], d1 = dbar[b * 3 * C + C + c0 + e], d2 = dbar[b * 3 * C + 2 * C + c0 + e];
      const float dot = bar[0][e] * d0 + bar[1][e] * d1 + bar[2][e] * d2;
      o[0][e] = bar[0][e] * (d0 - dot); o[1][e] = bar[1][e] * (d1 - dot); o[2][e] = bar[2][e] * (d2 - dot);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) *reinterpret_cast<uint4*>(dhat + b * 3 * C + k * C + c0) = pack8(o[k]);
  }
}
// dt[b, pos, kC + c] = bar_k[b, c] * adjoint_k(dout)[pos, c]   (MODE 0: combine backward, hat given)
//                    = da[b, c] * read_count_k(pos, c)         (MODE 1: sum backward, da given as bf16 [B, C])
//                    = both terms added                        (MODE 2: the whole split-attention backward in one write;
//                      as two autograd nodes it cost a second 3C-wide write plus a 3-tensor add pass)
template <int MODE>
__global__ void __launch_bounds__(RW_THREADS)
s2v2_dt_kernel(const __nv_bfloat16* __restrict__ src, const __nv_bfloat16* __restrict__ hat,
               const __nv_bfloat16* __restrict__ da, __nv_bfloat16* __restrict__ dt, int B, int H, int W, int C,
               int plain) {
  const int nvec = C >> 3;
  const int qs = C >> 2;
  const int plane = RW_THREADS / nvec;
  const int v = threadIdx.x % nvec, pl = threadIdx.x / nvec;
  const FastDiv dw_(W);
  const long long b = blockIdx.y;
  (void)B;
  if (pl >= plane) return;
  const int c0 = v * 8;
  const int q0 = min(c0 / qs, 3), q1 = min((c0 + 7) / qs, 3);
  float bar[3][8];
  if (MODE != 1) s2_softmax3(hat, b, c0, C, bar);
  float dav[8];
  if (MODE != 0) unpack8(*reinterpret_cast<const uint4*>(da + b * C + c0), dav);
  const long long simg = b * H * W * C;
#pragma unroll 2
  for (int pos = blockIdx.x * plane + pl; pos < H * W; pos += gridDim.x * plane) {
    int h, w;
    dw_.divmod(pos, h, w);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = 0.f;
      if (MODE != 1) {
        float gv[8];
        if (k < 2) s2_adjoint8(src, simg, h, w, k, c0, H, W, C, gv, plain);
        else unpack8(ldg_nc_v4(src + simg + (long long)pos * C + c0), gv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = bar[k][e] * gv[e];
      }
      if (MODE != 0) {
        if (k == 2 || q0 == q1) {
          int dh = 0, dw = 0;
          if (k < 2) s2_plan_offset(plain ? 2 : k, q0, dh, dw);
          const float cnt = (k < 2) ? s2_read_count(h, w, H, W, dh, dw) : 1.f;
#pragma unroll
          for (int e = 0; e < 8; ++e) o[e] += dav[e] * cnt;
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            int dh, dw;
            s2_plan_offset(plain ? 2 : k, min((c0 + e) / qs, 3), dh, dw);
            o[e] += dav[e] * s2_read_count(h, w, H, W, dh, dw);
          }
        }
      }
      *reinterpret_cast<uint4*>(dt + (b * H * W + pos) * 3 * C + k * C + c0) = pack8(o);
    }
  }
}

}  // namespace vmlp
