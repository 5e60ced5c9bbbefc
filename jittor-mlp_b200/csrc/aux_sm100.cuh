// Two HBM-bound helpers around the block path:
//   * permute5: a strided 5-D copy with a contiguous inner run -- Vision Permutator's `b h w (c s) -> b w c (h s)` /
//     `-> b h c (w s)` rearrangements and their inverses (vip.py:68-76), optionally accumulating into the destination;
//   * optim_step: multi-tensor AdamW / SGD-momentum step over bf16 parameters with fp32 master weights and moments
//     (the reference has no optimizer, compare.py:141-145 only interchanges state_dicts).
// Both are 1-read-1-write streaming kernels: bound = HBM, 16-byte vectors, grid = multiple of the SM count.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rowwise.cuh"

namespace vmlp {

struct Permute5 {
  int n1, n2, n3, nv;               // extents of dims 1..3 and 16-byte vectors per inner run (dim 0 extent = total / rest)
  long long is[4], os[4];           // element strides of dims 0..3 in the source / destination (inner run: stride 1)
  long long total;                  // n0 * n1 * n2 * n3 * nv
};

// The linear index walks the DESTINATION in (d0, d1, d2, d3, v) order: when `os` is the row-major stride set of that order
// the writes are fully coalesced and the reads are whole 32-byte sectors as long as the inner run is >= 16 elements.
template <bool ACC>
__global__ void __launch_bounds__(256)
permute5_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, const Permute5 p) {
  const FastDiv dv(p.nv), d3(p.n3), d2(p.n2), d1(p.n1);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.total; i += (long long)gridDim.x * blockDim.x) {
    // 64-bit quotient chain only for the outermost split; everything below fits 32 bits per image-sized slab
    const long long slab = (long long)p.n1 * p.n2 * p.n3 * p.nv;
    const long long i0 = i / slab;
    int r = static_cast<int>(i - i0 * slab);
    int q, v, c3, c2, c1;
    dv.divmod(r, q, v); r = q;
    d3.divmod(r, q, c3); r = q;
    d2.divmod(r, q, c2); r = q;
    d1.divmod(r, q, c1);
    (void)q;
    const long long si = i0 * p.is[0] + c1 * p.is[1] + c2 * p.is[2] + c3 * p.is[3] + v * 8;
    const long long di = i0 * p.os[0] + c1 * p.os[1] + c2 * p.os[2] + c3 * p.os[3] + v * 8;
    uint4 val = ldg_nc_v4(in + si);
    if (ACC) {
      float a[8], b[8];
      unpack8(val, a);
      unpack8(*reinterpret_cast<const uint4*>(out + di), b);
#pragma unroll
      for (int e = 0; e < 8; ++e) a[e] += b[e];
      val = pack8(a);
    }
    *reinterpret_cast<uint4*>(out + di) = val;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
struct OptimChunk {                 // one run of <= OPT_CHUNK elements of one parameter tensor
  __nv_bfloat16* param;
  const __nv_bfloat16* grad;
  long long state_off;              // offset of the run in the flat fp32 state buffers
  int n;
  int step;                         // > 0: this parameter's own step count t (bias corrections 1 - beta^t per tensor, like
                                    // torch.optim.AdamW); 0: use the corrections in OptimHyper
};
struct OptimHyper {
  int kind;                         // 0 = AdamW (decoupled weight decay), 1 = SGD with momentum (torch.optim.SGD)
  int first_step;                   // SGD: the momentum buffer starts as the first gradient
  float lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2_sqrt, grad_scale, momentum;
};
constexpr int OPT_CHUNK = 32768;

__device__ __forceinline__ float optim_update(const OptimHyper& h, float g, float& w, float& m, float& v) {
  g *= h.grad_scale;
  if (h.kind == 0) {
    w *= 1.f - h.lr * h.weight_decay;
    m = h.beta1 * m + (1.f - h.beta1) * g;
    v = h.beta2 * v + (1.f - h.beta2) * g * g;
    w -= (h.lr / h.bias_c1) * m / (sqrtf(v) / h.bias_c2_sqrt + h.eps);
  } else {
    g += h.weight_decay * w;
    m = h.first_step ? g : h.momentum * m + g;
    w -= h.lr * m;
  }
  return w;
}

// one block per chunk; 16-byte bf16 vectors when both pointers allow it, scalar otherwise (views into flat buckets may
// start on any 2-byte boundary)
__global__ void __launch_bounds__(256)
optim_step_kernel(const OptimChunk* __restrict__ table, float* __restrict__ master, float* __restrict__ mom,
                  float* __restrict__ var, const OptimHyper h_in) {
  const OptimChunk c = table[blockIdx.x];
  OptimHyper h = h_in;
  if (h.kind == 0 && c.step > 0) {
    h.bias_c1 = 1.f - exp2f(static_cast<float>(c.step) * log2f(h.beta1));
    h.bias_c2_sqrt = sqrtf(1.f - exp2f(static_cast<float>(c.step) * log2f(h.beta2)));
  }
  float* w = master + c.state_off;
  float* m = mom + c.state_off;
  float* v = (h.kind == 0) ? var + c.state_off : nullptr;
  const bool vec = ((reinterpret_cast<uintptr_t>(c.param) | reinterpret_cast<uintptr_t>(c.grad)) & 15) == 0 &&
                   (c.state_off & 3) == 0;
  const int nvec = vec ? (c.n >> 3) : 0;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float g[8], o[8];
    unpack8(ldg_nc_v4(c.grad + i * 8), g);
    float4 w0 = reinterpret_cast<float4*>(w)[2 * i], w1 = reinterpret_cast<float4*>(w)[2 * i + 1];
    float4 m0 = reinterpret_cast<float4*>(m)[2 * i], m1 = reinterpret_cast<float4*>(m)[2 * i + 1];
    float4 v0 = make_float4(0, 0, 0, 0), v1 = v0;
    if (v) { v0 = reinterpret_cast<float4*>(v)[2 * i]; v1 = reinterpret_cast<float4*>(v)[2 * i + 1]; }
    o[0] = optim_update(h, g[0], w0.x, m0.x, v0.x); o[1] = optim_update(h, g[1], w0.y, m0.y, v0.y);
    o[2] = optim_update(h, g[2], w0.z, m0.z, v0.z); o[3] = optim_update(h, g[3], w0.w, m0.w, v0.w);
    o[4] = optim_update(h, g[4], w1.x, m1.x, v1.x); o[5] = optim_update(h, g[5], w1.y, m1.y, v1.y);
    o[6] = optim_update(h, g[6], w1.z, m1.z, v1.z); o[7] = optim_update(h, g[7], w1.w, m1.w, v1.w);
    reinterpret_cast<float4*>(w)[2 * i] = w0; reinterpret_cast<float4*>(w)[2 * i + 1] = w1;
    reinterpret_cast<float4*>(m)[2 * i] = m0; reinterpret_cast<float4*>(m)[2 * i + 1] = m1;
    if (v) { reinterpret_cast<float4*>(v)[2 * i] = v0; reinterpret_cast<float4*>(v)[2 * i + 1] = v1; }
    *reinterpret_cast<uint4*>(c.param + i * 8) = pack8(o);
  }
  for (int i = nvec * 8 + threadIdx.x; i < c.n; i += blockDim.x) {
    float wi = w[i], mi = m[i], vi = v ? v[i] : 0.f;
    const float o = optim_update(h, __bfloat162float(c.grad[i]), wi, mi, vi);
    w[i] = wi; m[i] = mi;
    if (v) v[i] = vi;
    c.param[i] = __float2bfloat16(o);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Token / position mean of the classification heads (`x.mean(dim=1)` mlp_mixer.py:75, `Reduce('b h w c -> b c', 'mean')`
// hire_mlp.py:219, AdaptiveAvgPool2d(1) as_mlp.py:437-439): out[b, c] = (1 / P) sum_p x[b, p, c], fp32 accumulation.
// Block = 32 channel vectors x 8 position phases; a warp reads 512 contiguous bytes per position.
__global__ void __launch_bounds__(256)
token_mean_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out, int P, int C, float inv) {
  __shared__ float sh[8][32][9];
  const int nvec = C >> 3;
  const int vx = threadIdx.x & 31, ph = threadIdx.x >> 5;
  const int v = blockIdx.x * 32 + vx;
  const long long img = (long long)blockIdx.y * P * C;
  float acc[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (v < nvec) {
    for (int pos = ph; pos < P; pos += 8) {
      float f[8];
      unpack8(ldg_nc_v4(x + img + (long long)pos * C + v * 8), f);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] += f[e];
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) sh[ph][vx][e] = acc[e];
  __syncthreads();
  if (ph == 0 && v < nvec) {
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += sh[k][vx][e];
      o[e] = t * inv;
    }
    *reinterpret_cast<uint4*>(out + (long long)blockIdx.y * C + v * 8) = pack8(o);
  }
}
// dx[b, p, c] = g[b, c] / P
__global__ void __launch_bounds__(256)
token_mean_bwd_kernel(const __nv_bfloat16* __restrict__ g, __nv_bfloat16* __restrict__ dx, int P, int C, float inv) {
  const int nvec = C >> 3;
  const long long per_img = (long long)P * nvec;
  const long long b = blockIdx.y;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_img; i += (long long)gridDim.x * blockDim.x) {
    const int v = static_cast<int>(i % nvec);
    float f[8];
    unpack8(*reinterpret_cast<const uint4*>(g + b * C + v * 8), f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] *= inv;
    *reinterpret_cast<uint4*>(dx + b * P * C + i * 8) = pack8(f);
  }
}

}  // namespace vmlp
