/*
 * vmlp_b200.h -- C ABI of the B200-native vision-MLP block library (libvmlp_b200.so).
 *
 * Drop-in boundary for the per-block bodies of liuruiyang98/Jittor-MLP `models_pytorch`.
 * The reference has exactly one custom-operator boundary, `_shift.forward/backward`
 * (models_pytorch/utils/shift_cuda.py:106-162): a torch.autograd.Function that receives
 * contiguous device tensors, allocates its output on the Python side, and launches a raw
 * kernel on `torch.cuda.current_stream()` with `data_ptr()` arguments.  Every entry point
 * below follows that same convention, generalised to whole blocks:
 *
 *   - plain pointers, sizes and a stream handle; no torch / C++ types in any signature;
 *   - all device memory (inputs, outputs, saved activations, workspace) is caller-owned;
 *     the library never allocates, frees or retains a pointer past return;
 *   - asynchronous on the given stream, no host synchronisation, re-entrant;
 *   - every function returns 0 on success or a negative VMLP_E* code; nothing falls back
 *     to a CPU or library path.
 *
 * Element type is bf16 (device `__nv_bfloat16`, passed as `void*`) unless a parameter is
 * declared `float*`.  All row strides ("ld") and batch strides are in ELEMENTS and must be
 * multiples of 8 (16 bytes) because tiles are moved by TMA.
 */
#ifndef VMLP_B200_H_
#define VMLP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vmlp_stream_t; /* cudaStream_t */

enum {
  VMLP_OK = 0,
  VMLP_EINVAL = -1,      /* bad shape / null pointer / unsupported hyper-parameter */
  VMLP_EALIGN = -2,      /* pointer or stride not 16-byte aligned                  */
  VMLP_EARCH = -3,       /* device is not sm_100 (no fallback path exists)         */
  VMLP_ELAUNCH = -4,     /* CUDA launch / driver error (see vmlp_last_error)       */
  VMLP_EWORKSPACE = -5   /* caller-provided workspace too small                     */
};

/* Library / device introspection.  VMLP_ABI_VERSION changes whenever a struct or a signature below changes; a binding
 * checks it together with vmlp_abi_struct_bytes (0 operand, 1 gemm_args, 2 mixer_params, 3 mixer_saved, 4 hire_dims)
 * before the first call.  vmlp_source_hash: digest of the sources the library was built from (stale-build detection). */
#define VMLP_ABI_VERSION 3
int vmlp_abi_version(void);
const char* vmlp_source_hash(void);
int vmlp_abi_struct_bytes(int32_t which);
const char* vmlp_last_error(void);          /* thread-local message for the last failing call */
int vmlp_device_check(void);                /* VMLP_OK iff current device is compute capability 10.x */
int vmlp_sm_count(void);
/* Diagnostics of a kernel that ended in a bounded-wait trap (protocol bug): out[0] = number of records, then from
 * out[4] on 4 words per record (block, warp, shared-memory address of the mbarrier, parity).  Host memory: readable
 * after the CUDA context has died.  Returns the number of words copied. */
int vmlp_debug_read(uint32_t* out, int32_t n_words);
int64_t vmlp_launch_count(void);             /* kernels launched by this library in this process */

/* --------------------------------------------------------------------------------------------
 * Generic fused GEMM (the building block of every mixing MLP):
 *     D[b][m, n] = epilogue( sum_k A[b][m, k] * B[b][n, k] )
 * A matrix operand is described by a dense 2-D view plus an optional batch:
 *     major = 0  "K-major":  rows = M (or N), cols = K   (a row-major weight `[out, in]`)
 *     major = 1  "MN-major": rows = K,        cols = M/N (a `[tokens, channels]` activation used
 *                                                          as the contraction-over-tokens operand)
 * batch_stride == 0 means the operand is shared by all batches.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  const void* ptr;      /* bf16 */
  int64_t rows, cols;   /* logical extent of the 2-D view */
  int64_t ld;           /* row stride, elements */
  int64_t batch_stride; /* elements; 0 = shared */
  int32_t major;        /* 0 = K-major, 1 = MN-major */
} vmlp_operand;

enum {
  VMLP_EPI_STORE = 0,     /* D = acc (+bias)                                  */
  VMLP_EPI_GELU = 1,      /* z = acc + bias ; D = gelu_erf'(z) ; D2 = gelu_erf(z) */
  VMLP_EPI_RESID = 2,     /* D = (acc + bias) * colscale + aux                */
  VMLP_EPI_DGELU = 3,     /* D = acc * aux  (aux = the gelu_erf'(z) saved by EPI_GELU) */
  VMLP_EPI_ATOMIC = 4,    /* out_f32 += acc   (split-K, weight gradients)     */
  VMLP_EPI_MUL = 5,       /* D = (acc + bias) * aux                           */
  VMLP_EPI_GELU_ONLY = 6, /* D = gelu_erf(acc + bias)                         */
  VMLP_EPI_RESID_DUAL = 7,/* VMLP_EPI_RESID and D2 = acc + bias               */
  VMLP_EPI_MUL_DUAL = 8   /* VMLP_EPI_MUL and D2 = acc + bias                 */
};

typedef struct {
  int32_t M, N, K;         /* per-batch logical GEMM extents */
  int32_t batch;           /* number of batches (>= 1) */
  int32_t contract_batch;  /* 1: the contraction also runs over the batch (one output matrix) */
  vmlp_operand A, B;
  int32_t epilogue;        /* VMLP_EPI_* */
  void* D;  int64_t d_ld, d_bs;    /* bf16 output [batch][M, N] */
  void* D2; int64_t d2_ld, d2_bs;  /* second output (VMLP_EPI_GELU) */
  const void* bias; int32_t bias_mode; /* 0 none, 1 per column n, 2 per row m */
  const void* colscale;                /* optional per-column scale (VMLP_EPI_RESID) */
  const void* aux; int64_t aux_ld, aux_bs;
  float* out_f32; int64_t out_ld;      /* VMLP_EPI_ATOMIC destination [M, N] (caller zero-fills) */
  int32_t split_k;                     /* 0 = choose automatically */
  int32_t block_n;                     /* 0 = choose automatically (128, 256; 208 for VMLP_EPI_ATOMIC with 128 < N <= 208) */
  int32_t cta_group;                   /* 0 = automatic, 1 = one CTA per 128-row tile, 2 = CTA pair per 256-row tile */
  float* red_out; int32_t red_mode;    /* optional: += sums of the stored D per column (1) or per row (2): the bias
                                          gradient when D is d(pre-activation); caller zero-fills */
  int32_t out_trans;                   /* VMLP_EPI_ATOMIC: add element (m, n) at out_f32[n * out_ld + m] */
} vmlp_gemm_args;

int vmlp_gemm_bf16(const vmlp_gemm_args* args, vmlp_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Row-wise operators (HBM-bound glue between the GEMMs)
 * ------------------------------------------------------------------------------------------ */
/* torch.nn.LayerNorm(C) over the last axis -- PreNormResidual.norm, models_pytorch/mlp_mixer.py:10-13 */
int vmlp_layernorm_fwd(const void* x, int64_t x_ld, const void* gamma, const void* beta, void* y, int64_t y_ld,
                       float* mean, float* rstd, int64_t rows, int32_t C, float eps, vmlp_stream_t stream);
/* dx = add + LN'(dy); dgamma/dbeta are fp32 accumulators (+=) */
int vmlp_layernorm_bwd(const void* dy, int64_t dy_ld, const void* x, int64_t x_ld, const float* mean,
                       const float* rstd, const void* gamma, const void* add, int64_t add_ld, void* dx,
                       int64_t dx_ld, float* dgamma, float* dbeta, int64_t rows, int32_t C, vmlp_stream_t stream);
/* Aff: y = x * alpha + beta -- models_pytorch/res_mlp.py:11-19 */
int vmlp_affine_fwd(const void* x, const void* alpha, const void* beta, void* y, int64_t rows, int32_t C,
                    vmlp_stream_t stream);
int vmlp_affine_bwd(const void* dy, const void* x, const void* alpha, const void* add, void* dx, float* dalpha,
                    float* dbeta, int64_t rows, int32_t C, vmlp_stream_t stream);
/* out[c] += sum_r a[r, c] * (b ? b[r, c] : 1) */
int vmlp_colsum(const void* a, int64_t a_ld, const void* b, int64_t b_ld, float* out, int64_t rows, int32_t C,
                vmlp_stream_t stream);
/* one pass, contiguous [rows, C] tensors (C <= 2048): out_a[c] += sum_r a[r, c];  out_ab[c] += sum_r a[r, c] * b[r, c]
   (b may alias a).  The two batch statistics of nn.BatchNorm2d in training mode (conv_mixer.py:20,27,31) and of its
   backward in one read of each tensor. */
int vmlp_colsum2(const void* a, const void* b, float* out_a, float* out_ab, int64_t rows, int32_t C,
                 vmlp_stream_t stream);
/* out[m] += sum_{b, c} a[b, m, c] */
int vmlp_rowsum_batched(const void* a, float* out, int64_t batch, int32_t rows_per_batch, int32_t C,
                        vmlp_stream_t stream);
int vmlp_cast_f32_to_bf16(const float* src, void* dst, int64_t n, vmlp_stream_t stream);
/* dst[r, 0:cols] = src[r, 0:cols], zero elsewhere: gives a [rows, cols] weight a 16-byte row pitch (ld_dst % 8 == 0) */
int vmlp_pad_rows(const void* src, void* dst, int32_t rows, int32_t cols, int32_t ld_dst, vmlp_stream_t stream);
int vmlp_add_bf16(const void* a, const void* b, void* dst, int64_t n, vmlp_stream_t stream);
/* Strided elementwise helpers over [rows, C] views (row strides in elements, C % 8 == 0):
 *   vmlp_mul_colvec : out = a * v[c]                   (ResMLP layer-scale backward, res_mlp.py:54,56)
 *   vmlp_dgelu_mul  : out = a * g            (g = saved gelu_erf'(z))
 *   vmlp_gate_bwd   : out = dg * vt * g_u ; out2 = dg * u                 (gMLP SGU gate, g_mlp.py:21) */
int vmlp_mul_colvec(const void* a, int64_t a_ld, const void* v, void* out, int64_t out_ld, int64_t rows, int32_t C,
                    vmlp_stream_t stream);
int vmlp_dgelu_mul(const void* a, int64_t a_ld, const void* z, int64_t z_ld, void* out, int64_t out_ld, int64_t rows,
                   int32_t C, vmlp_stream_t stream);
int vmlp_gate_bwd(const void* dg, int64_t dg_ld, const void* vt, int64_t vt_ld, const void* zp_u, int64_t zp_ld,
                  const void* u, int64_t u_ld, void* out, int64_t out_ld, void* out2, int64_t out2_ld, int64_t rows,
                  int32_t C, vmlp_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Spatial operators of the shift family.  Tensors are channels-last [B, H, W, C] bf16, C % 8 == 0.
 * ------------------------------------------------------------------------------------------ */
/* Channel-group token shift: group g = channels [start[g], start[g+1]) reads (h + dh[g], w + dw[g]).
 *   mode 0 zero padding   -- AS-MLP Shift forward; its backward is the same call with negated offsets
 *                            (models_pytorch/utils/shift_cuda.py:44-103)
 *   mode 1 clamp-to-edge  -- S2-MLP spatial_shift, intended semantics (s2_mlp_v1.py:19-25, s2_mlp_v2.py:15-29)
 *   mode 2 adjoint of mode 1 (|offset| <= 1)
 * start has ngroups + 1 entries; ngroups <= 8. */
int vmlp_shift_nhwc(const void* in, void* out, int32_t B, int32_t H, int32_t W, int32_t C, int32_t mode,
                    int32_t ngroups, const int32_t* start, const int32_t* dh, const int32_t* dw, vmlp_stream_t stream);
/* GroupNorm(1, C) (as_mlp.py:343-344): statistics over each sample's P*C elements.
 * acc: fp32 [B][2] (sum, sum of squares), zero-filled by the caller before vmlp_gn_stats. */
int vmlp_gn_stats(const void* x, float* acc, int32_t B, int64_t P, int32_t C, vmlp_stream_t stream);
int vmlp_gn_apply(const void* x, const float* acc, const void* gamma, const void* beta, void* y, int32_t B, int64_t P,
                  int32_t C, float eps, int32_t gelu, vmlp_stream_t stream);
/* dx of y = [gelu](gn(x)); dn: bf16 scratch [B*P*C]; acc2: fp32 [B][2] zero-filled; dgamma/dbeta fp32 += */
int vmlp_gn_bwd(const void* dy, const void* x, const float* acc, const void* gamma, const void* beta, void* dn,
                float* acc2, float* dgamma, float* dbeta, void* dx, int32_t B, int64_t P, int32_t C, float eps,
                int32_t gelu, vmlp_stream_t stream);
/* out = (A[c]*p + Bq[c]*q + Cc[c]) * (z ? gelu_erf'(z) : 1); q, Bq, z optional; A/Bq/Cc fp32 [C] */
int vmlp_chan_lin(const void* p, const void* q, const void* z, const float* A, const float* Bq, const float* Cc,
                  void* out, int64_t rows, int32_t C, vmlp_stream_t stream);
/* BatchNorm2d training-mode coefficients (conv_mixer.py:20,27,31): s1 = sum(a), s2 = sum(a*a) over R rows */
int vmlp_bn_fwd_coef(const float* s1, const float* s2, const void* gamma, const void* beta, float* A, float* Cc,
                     float* mean, float* rstd, float* running_mean, float* running_var, int64_t R, float eps,
                     float momentum, int32_t C, vmlp_stream_t stream);
int vmlp_bn_bwd_coef(const float* sdy, const float* sdya, const void* gamma, const float* mean, const float* rstd,
                     float* A, float* Bq, float* Cc, float* dgamma, float* dbeta, int64_t R, int32_t C,
                     vmlp_stream_t stream);

/* Split attention of S2-MLPv2 (s2_mlp_v2.py:31-69) and Vision Permutator (vip.py:37-57).  t: [B,H,W,3C]; branch k =
 * t[..., kC:(k+1)C].  plain = 0: S2-MLPv2 -- branch 0 / 1 are read through spatial_shift1 / spatial_shift2 as load-time
 * clamp offsets, branch 2 unshifted.  plain = 1: ViP -- three unshifted branches (the stacked H / W / C permute-MLP outputs).
 *   sum      : a_f32[b,c] += sum_pos (x_0 + x_1 + x_2)                        (caller zero-fills a_f32)
 *   combine  : out[b,pos,c] = sum_k softmax_k(hat[b,:,c]) * x_k[b,pos,c]      (hat: bf16 [B,3C])
 *   combine_bwd : dt (bf16 [B,H,W,3C]) and dhat (bf16 [B,3C]) from dout; dbar_f32 is a zero-filled fp32 [B,3C] scratch
 *   sum_bwd  : dt from da (bf16 [B,C]) */
int vmlp_s2v2_sum(const void* t, float* a_f32, int32_t B, int32_t H, int32_t W, int32_t C, int32_t plain,
                  vmlp_stream_t stream);
int vmlp_s2v2_combine(const void* t, const void* hat, void* out, int32_t B, int32_t H, int32_t W, int32_t C,
                      int32_t plain, vmlp_stream_t stream);
int vmlp_s2v2_combine_bwd(const void* t, const void* hat, const void* dout, float* dbar_f32, void* dhat, void* dt,
                          int32_t B, int32_t H, int32_t W, int32_t C, int32_t plain, vmlp_stream_t stream);
int vmlp_s2v2_sum_bwd(const void* da, void* dt, int32_t B, int32_t H, int32_t W, int32_t C, int32_t plain,
                      vmlp_stream_t stream);
/* the whole split-attention backward w.r.t. t in one write: dt = softmax_k(hat) * adjoint_k(dout) + da * read_count_k.
   Pair with vmlp_s2v2_combine_bwd(..., dt = NULL, ...), which then only produces dbar / dhat. */
int vmlp_s2v2_dt_fused(const void* dout, const void* hat, const void* da, void* dt, int32_t B, int32_t H, int32_t W,
                       int32_t C, int32_t plain, vmlp_stream_t stream);

/* Strided 5-D copy with a contiguous inner run: out[i0*os0 + i1*os1 + i2*os2 + i3*os3 + j] (+)= in[i0*is0 + ... + j],
 * 0 <= j < dims[4].  Vision Permutator's `b h w (c s) -> b w c (h s)` / `b h w (c s) -> b h c (w s)` rearrangements and
 * their inverses (vip.py:68-70,73-75) are this copy with the inner run = one segment; the destination may be a channel
 * slot of a wider buffer (row stride in os).  dims[4] and every stride must be multiples of 8 elements, both pointers
 * 16-byte aligned, dims[1]*dims[2]*dims[3]*dims[4]/8 < 2^22.  accumulate != 0: out += in (bf16, fp32 add). */
int vmlp_permute5(const void* in, void* out, const int32_t dims[5], const int64_t in_strides[4],
                  const int64_t out_strides[4], int32_t accumulate, vmlp_stream_t stream);

/* Position mean of the classification heads (x.mean(dim=1), mlp_mixer.py:75; Reduce('b h w c -> b c', 'mean'),
 * hire_mlp.py:219; AdaptiveAvgPool2d(1), as_mlp.py:437-439): out[b,c] = mean over P positions of x[b,p,c] (fp32
 * accumulation); backward dx[b,p,c] = g[b,c] / P.  C % 8 == 0. */
int vmlp_token_mean(const void* x, void* out, int32_t B, int32_t P, int32_t C, vmlp_stream_t stream);
int vmlp_token_mean_bwd(const void* g, void* dx, int32_t B, int32_t P, int32_t C, vmlp_stream_t stream);

/* Multi-tensor optimizer step (the reference has no optimizer -- compare.py:141-145 only
 * interchanges state_dicts -- so the semantics are torch.optim.AdamW / torch.optim.SGD(momentum) on fp32 master weights).
 * table: DEVICE array of n_chunks entries, each a run of <= 32768 elements of one bf16 parameter tensor and its bf16
 * gradient; state_off = offset of the run in the flat fp32 buffers master / mom / var (var unused for SGD).
 * One launch updates master, moments and the bf16 parameter copy of every run. */
typedef struct {
  void* param;           /* bf16, updated in place */
  const void* grad;      /* bf16 */
  int64_t state_off;
  int32_t n;
  int32_t step;          /* > 0: step count t of THIS parameter (AdamW bias corrections 1 - beta^t are per tensor in
                            torch.optim: a parameter that skipped a step lags behind); 0: use vmlp_optim_hyper's */
} vmlp_optim_chunk;
typedef struct {
  int32_t kind;          /* 0 = AdamW, 1 = SGD with momentum */
  int32_t first_step;    /* SGD: momentum buffer := gradient on the first step (torch.optim.SGD) */
  float lr, beta1, beta2, eps, weight_decay;
  float bias_c1;         /* 1 - beta1^t       } used by chunks whose own `step` is 0 */
  float bias_c2_sqrt;    /* sqrt(1 - beta2^t) } */
  float grad_scale;      /* gradients are multiplied by this first (1/world for SUM all-reduces, loss-scale inverse) */
  float momentum;
} vmlp_optim_hyper;
int vmlp_optim_step(const vmlp_optim_chunk* table_dev, int32_t n_chunks, float* master, float* mom, float* var,
                    const vmlp_optim_hyper* hyper, vmlp_stream_t stream);

/* Hire-MLP region rearrangement (hire_mlp.py:44-152) as load-time index arithmetic on channels-last tensors.
 * Padding is circular with Hp = H + (h - H % h), Wp = W + (w - W % w); step = cross_region_step or 0.
 *   build       : zh [B, Hp/h, W, h*C], zw [B, H, Wp/w, w*C]  (feature axis ordered [region index][channel])
 *   build_adj   : dx = adjoint of both gathers (sums the circular-padding duplicates)
 *   combine     : out = base + restore_H(oh) + restore_W(ow), cropped to H x W
 *   restore_adj : dzh, dzw from dout (entries that land in the cropped padding get 0) */
typedef struct {
  int32_t B, H, W, C;
  int32_t h, w;              /* region sizes */
  int32_t step_h, step_w;    /* roll amounts (0 when this block does not cross regions) */
} vmlp_hire_dims;
int vmlp_hire_build(const void* x, void* zh, void* zw, const vmlp_hire_dims* d, vmlp_stream_t stream);
int vmlp_hire_build_adj(const void* dzh, const void* dzw, void* dx, const vmlp_hire_dims* d, vmlp_stream_t stream);
int vmlp_hire_combine(const void* base, const void* oh, const void* ow, void* out, const vmlp_hire_dims* d,
                      vmlp_stream_t stream);
int vmlp_hire_restore_adj(const void* dout, void* dzh, void* dzw, const vmlp_hire_dims* d, vmlp_stream_t stream);

/* ConvMixer depthwise k x k convolution, padding "same", channels-last (conv_mixer.py:24); K in {3, 5, 7, 9}.
 * weight: bf16 [C, 1, K, K].  fwd writes z = gelu_erf'(conv + bias) (for backward) and a = gelu_erf(conv + bias);
 * dgrad is the 180-degree-rotated stencil;
 * wgrad accumulates fp32 dW [C, K, K] (caller zero-fills). */
int vmlp_dwconv_fwd(const void* x, const void* weight, const void* bias, void* z, void* a, int32_t B, int32_t H,
                    int32_t W, int32_t C, int32_t K, vmlp_stream_t stream);
/* depthwise conv + bias WITHOUT activation (sparse_mlp.py:88-90: Conv2d(d, d, 3, padding 1, groups = d) inside a
 * BatchNorm pre-norm residual); backward = vmlp_dwconv_dgrad / _wgrad with dz = dy. */
int vmlp_dwconv_fwd_plain(const void* x, const void* weight, const void* bias, void* y, int32_t B, int32_t H, int32_t W,
                          int32_t C, int32_t K, vmlp_stream_t stream);
int vmlp_dwconv_dgrad(const void* dz, const void* weight, void* dx, int32_t B, int32_t H, int32_t W, int32_t C,
                      int32_t K, vmlp_stream_t stream);
int vmlp_dwconv_wgrad(const void* x, const void* dz, float* dw, int32_t B, int32_t H, int32_t W, int32_t C, int32_t K,
                      vmlp_stream_t stream);

/* Patch-embedding stem: Conv2d(Cin, C, kernel = stride = P) as a GEMM over non-overlapping patches
 * (mlp_mixer.py:58-60,68-71).  x: NCHW bf16; rows: [B * (H/P) * (W/P), Cin * P * P] with column order (ci, i, j) = the
 * Conv2d weight's own [C, Cin*P*P] layout.  P % 8 == 0.  forward != 0 gathers rows from x; forward == 0 scatters
 * d(rows) back into d(x) (exact inverse). */
int vmlp_patchify(const void* src, void* dst, int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t P, int32_t forward,
                  vmlp_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * Fused token-mixing MLP (models_pytorch/mlp_mixer.py:16-27,34,37 -- FeedForward with Conv1d(k=1) over the token axis):
 *     forward :  u[b]   = x[b] + W2 gelu_erf(W1 xhat[b] + b1) + b2        xhat, x, u : [B, N, C];  W1 [Ds, N], W2 [N, Ds]
 *     backward:  dxhat[b] = W1^T ((W2^T du[b]) .* gelu_erf'(W1 xhat[b] + b1))
 * One kernel each: the hidden tensor [B, Ds, C] stays in TMEM / shared memory between the two contractions (forward)
 * and is recomputed on chip in backward.  Saved / produced for the weight gradients, TRANSPOSED ([B, C, Ds], the layout
 * the epilogue holds them in):  hT = gelu(..)^T (forward, optional), dzT = d(pre-activation)^T (backward).
 * Weight operands are K-major copies made by vmlp_tokmix_prepare (w [rows, cols] -> pad [rows, ld] zero-padded and / or
 * tr [cols, ldt] transposed, zero-padded; either may be NULL):
 *     w1_pad  = pad(W1)  [Ds, Np]      w2T_pad = tr(W2) [Ds, Np]      w1T = tr(W1) [N, Ds]        Np >= ceil16(N)
 * db1 (fp32 [Ds], optional) += sum over (b, c) of dz.  vmlp_tokmix_supported: N <= 256, Ds <= 1024, C % 8 == Ds % 8 == 0
 * and the tiles of that shape fit in shared memory; otherwise the caller composes vmlp_gemm_bf16 calls.
 * ------------------------------------------------------------------------------------------ */
int vmlp_tokmix_supported(int32_t B, int32_t N, int32_t C, int32_t Ds, int32_t backward);
/* bring-up: int64 device buffer [4][64][8] that the forward kernel's CTA 0 fills with clock64 stamps (NULL = off) */
int vmlp_tokmix_set_trace(void* buf);
int vmlp_tokmix_prepare(const void* w, int32_t rows, int32_t cols, void* pad, int32_t ld, void* tr, int32_t ldt,
                        vmlp_stream_t stream);
int vmlp_tokmix_fwd(const void* xhat, const void* x, const void* w1_pad, int32_t Np, const void* w2, const void* b1,
                    const void* b2, void* u, void* hT, int32_t B, int32_t N, int32_t C, int32_t Ds, vmlp_stream_t stream);
int vmlp_tokmix_bwd(const void* xhat, const void* du, const void* w1_pad, const void* w2T_pad, int32_t Np,
                    const void* w1T, const void* b1, void* dxhat, void* dzT, float* db1, int32_t B, int32_t N, int32_t C,
                    int32_t Ds, vmlp_stream_t stream);

/* --------------------------------------------------------------------------------------------
 * MLP-Mixer block: models_pytorch/mlp_mixer.py:35-40
 *     u = x + TokenFF(LN1(x))   (FeedForward with Conv1d(k=1) over tokens, :16-27,:37)
 *     y = u + ChanFF(LN2(u))    (FeedForward with Linear over channels,    :38)
 * Shapes: x, y, u : [B, N, C];  token weights W1t [Ds, N], W2t [N, Ds] (Conv1d weight[:, :, 0]);
 * channel weights W1c [Dc, C], W2c [C, Dc].  C % 64 == 0, Dc % 64 == 0, Ds % 8 == 0.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t B, N, C, Ds, Dc;
  float eps;
  /* parameters (bf16) */
  const void *ln1_w, *ln1_b, *w1t, *b1t, *w2t, *b2t;
  const void *ln2_w, *ln2_b, *w1c, *b1c, *w2c, *b2c;
} vmlp_mixer_params;

/* Activations the forward pass keeps for backward (caller-allocated, bf16 unless noted):
 *   xhat1 [B,N,C]  z1,h1 [B,Ds,C]  u [B,N,C]  xhat2 [B,N,C]  z2,h2 [B*N,Dc]
 *   When vmlp_mixer_token_fused(p) != 0 the token half runs as vmlp_tokmix_fwd/_bwd: z1 is not used (may be NULL) and
 *   h1 holds the transposed hidden activation [B, C, Ds] (same element count).
 *   stats: fp32 [4][B*N] = mean1, rstd1, mean2, rstd2
 *   w1t_pad: bf16 [Ds, ceil16(N)] scratch for the zero-padded copy of W1t (16-byte pitch, whole k-steps of 16) */
typedef struct {
  void *xhat1, *z1, *h1, *u, *xhat2, *z2, *h2;
  float* stats;
  void* w1t_pad;
} vmlp_mixer_saved;

int vmlp_mixer_token_fused(const vmlp_mixer_params* p);
int vmlp_mixer_block_fwd(const vmlp_mixer_params* p, const void* x, void* y, const vmlp_mixer_saved* s,
                         vmlp_stream_t stream);

/* Backward.  grads_f32 is one flat fp32 accumulator (caller zero-fills) laid out in the order of
 * vmlp_mixer_params: ln1_w[C] ln1_b[C] w1t[Ds*N] b1t[Ds] w2t[N*Ds] b2t[N] ln2_w[C] ln2_b[C]
 * w1c[Dc*C] b1c[Dc] w2c[C*Dc] b2c[C]  (vmlp_mixer_grad_elems gives the total).
 * workspace: bf16, vmlp_mixer_bwd_workspace_elems(...) elements. */
int64_t vmlp_mixer_grad_elems(const vmlp_mixer_params* p);
int64_t vmlp_mixer_bwd_workspace_elems(const vmlp_mixer_params* p);
int vmlp_mixer_block_bwd(const vmlp_mixer_params* p, const void* x, const void* dy, void* dx,
                         const vmlp_mixer_saved* s, float* grads_f32, void* workspace, int64_t workspace_elems,
                         vmlp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VMLP_B200_H_ */
