// CUDA-core kernels of the Conformer hot path (everything that is not a tensor-core contraction):
// fused (double) LayerNorm, CMVN + first subsampling conv, depthwise-conv + norm + Swish,
// row soft-max, CTC greedy decode. All are HBM-bandwidth bound: coalesced 16-byte accesses,
// warp-shuffle reductions, no re-reads.
#include <cstdlib>
#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace ppasr {

void count_launch();

// ------------------------------------------------------------------------------------------------
// LayerNorm over D=256 channels, one warp per row (8 channels per lane).
//   single: y_bf16 = LN(x; g1, b1)                      (rows with t >= lens[b] are written as 0)
//   double: x <- LN(x; g1, b1) (fp32, in place), y_bf16 = LN(x; g2, b2)
// Reference: the five LayerNorms of ConformerEncoderLayer (ppasr/model_utils/conformer/encoder.py:327-336,
// 380-429): norm_final of layer i is fused with norm_ff_macaron of layer i+1 (or after_norm).
// ------------------------------------------------------------------------------------------------
template <int D>
__global__ void __launch_bounds__(256) layernorm_kernel(float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                        const float* __restrict__ g1, const float* __restrict__ b1,
                                                        const float* __restrict__ g2, const float* __restrict__ b2,
                                                        const int* __restrict__ lens, int T, int M, float eps) {
  constexpr int PER = D / 32;
  static_assert(PER % 4 == 0, "D must be a multiple of 128");
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  float v[PER];
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * D + lane * PER);
#pragma unroll
  for (int i = 0; i < PER / 4; ++i) {
    float4 t = src[i];
    v[4 * i] = t.x, v[4 * i + 1] = t.y, v[4 * i + 2] = t.z, v[4 * i + 3] = t.w;
  }
  auto norm = [&](const float* g, const float* b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / D);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float d = v[i] - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = (v[i] - mean) * rstd * __ldg(g + lane * PER + i) + __ldg(b + lane * PER + i);
  };
  norm(g1, b1);
  if (g2 != nullptr) {
    float4* dst = reinterpret_cast<float4*>(x + (size_t)row * D + lane * PER);
#pragma unroll
    for (int i = 0; i < PER / 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    norm(g2, b2);
  }
  bool zero = false;
  if (lens != nullptr) {
    const int b = row / T;
    zero = (row - b * T) >= __ldg(lens + b);
  }
  uint32_t pk[PER / 2];
#pragma unroll
  for (int i = 0; i < PER / 2; ++i) pk[i] = zero ? 0u : pack_bf16x2(v[2 * i], v[2 * i + 1]);
  uint4* dy = reinterpret_cast<uint4*>(y + (size_t)row * D + lane * PER);
#pragma unroll
  for (int i = 0; i < PER / 8; ++i) dy[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
}

cudaError_t launch_layernorm(float* x, __nv_bfloat16* y, const float* g1, const float* b1, const float* g2,
                             const float* b2, const int* lens, int T, int M, int D, float eps, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  const int grid = (M + 7) / 8;
  if (D == 256)
    layernorm_kernel<256><<<grid, 256, 0, st>>>(x, y, g1, b1, g2, b2, lens, T, M, eps);
  else if (D == 512)
    layernorm_kernel<512><<<grid, 256, 0, st>>>(x, y, g1, b1, g2, b2, lens, T, M, eps);
  else
    return cudaErrorInvalidValue;
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// GlobalCMVN + Conv2d(1 -> C, 3x3, stride 2) + ReLU, written as four stride-phase images so that the
// second stride-2 conv becomes nine unit-stride shifted GEMM taps (see gemm.cuh, CONV mode).
//   reference: ppasr/model_utils/utils/cmvn.py:29-32, ppasr/model_utils/conformer/subsampling.py:84-85,110
//   in : feats fp32 [B, T, F]
//   out: P[ph][b][th][fh][c] bf16, ph = (t1&1)*2 + (f1&1), th = t1>>1 in [0,Th), fh = f1>>1 in [0,FH)
//        (slots with t1 >= T1 or f1 >= F1 are written as 0)
// One block per (b, th), both time phases, all frequency slots.
// ------------------------------------------------------------------------------------------------
// Thread = (channel octet, frequency group): 8 output channels per thread so every store is 16 bytes and a
// warp writes 512 contiguous bytes; the 3x3 input window is a shared-memory broadcast. A block covers
// CONV1_THB consecutive `th` so the 72 weight registers and the input window are set up once per 8x more work.
constexpr int CONV1_THB = 8;
template <int C>
__global__ void __launch_bounds__(256) conv1_subsample_kernel(const float* __restrict__ feats,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ istd,
                                                              const float* __restrict__ w,     // [C, 9]
                                                              const float* __restrict__ bias,  // [C]
                                                              __nv_bfloat16* __restrict__ out, int B, int T, int F,
                                                              int T1, int F1, int Th, int FH) {
  static_assert(C == 256, "one warp spans the 256 channels as 32 octets");
  extern __shared__ float xs[];  // [4*CONV1_THB + 1][F] normalised input rows 4*th0 ..
  const int tiles = (Th + CONV1_THB - 1) / CONV1_THB;
  const int b = blockIdx.x / tiles;
  const int th0 = (blockIdx.x - b * tiles) * CONV1_THB;
  const int nrows = 4 * CONV1_THB + 1;
  for (int i = threadIdx.x; i < nrows * F; i += 256) {
    const int r = i / F, f = i - r * F;
    const int t = 4 * th0 + r;
    float v = 0.f;
    if (t < T) v = (feats[((size_t)b * T + t) * F + f] - __ldg(mean + f)) * __ldg(istd + f);
    xs[i] = v;
  }
  const int co = (threadIdx.x & 31) * 8;  // first of this thread's 8 channels
  const int fg = threadIdx.x >> 5;        // 0..7
  float wr[8][9];
  float bs[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    bs[c] = __ldg(bias + co + c);
#pragma unroll
    for (int i = 0; i < 9; ++i) wr[c][i] = __ldg(w + (co + c) * 9 + i);
  }
  __syncthreads();
  const size_t phase_stride = (size_t)B * Th * FH * C;
#pragma unroll 1
  for (int dt = 0; dt < 2 * CONV1_THB; ++dt) {
    const int th = th0 + (dt >> 1), pt = dt & 1;
    if (th >= Th) break;
    const int t1 = 2 * th + pt;
    const float* r0 = xs + (2 * dt) * F;  // input row 2*t1 - 4*th0 = 2*dt
#pragma unroll 1
    for (int f1 = fg; f1 < 2 * FH; f1 += 8) {
      float acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.f;
      if (t1 < T1 && f1 < F1) {
        float xv[9];
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) xv[kh * 3 + kw] = r0[kh * F + 2 * f1 + kw];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float a = bs[c];
#pragma unroll
          for (int i = 0; i < 9; ++i) a = fmaf(wr[c][i], xv[i], a);
          acc[c] = fmaxf(a, 0.f);
        }
      }
      const int ph = pt * 2 + (f1 & 1);
      uint4 o = make_uint4(pack_bf16x2(acc[0], acc[1]), pack_bf16x2(acc[2], acc[3]), pack_bf16x2(acc[4], acc[5]),
                           pack_bf16x2(acc[6], acc[7]));
      *reinterpret_cast<uint4*>(out + ph * phase_stride + (((size_t)b * Th + th) * FH + (f1 >> 1)) * C + co) = o;
    }
  }
}

cudaError_t launch_conv1_subsample(const float* feats, const float* mean, const float* istd, const float* w,
                                   const float* bias, __nv_bfloat16* out, int B, int T, int F, int C, int T1, int F1,
                                   int Th, int FH, cudaStream_t st) {
  if (C != 256) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(4 * CONV1_THB + 1) * F * sizeof(float);
  const int tiles = (Th + CONV1_THB - 1) / CONV1_THB;
  conv1_subsample_kernel<256><<<B * tiles, 256, smem, st>>>(feats, mean, istd, w, bias, out, B, T, F, T1, F1, Th, FH);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Depthwise Conv1d(k) over time + (LayerNorm over channels | folded BatchNorm) + Swish.
//   reference: ppasr/model_utils/conformer/convolution.py:125-131
//   in : g bf16 [B, Tin, C] (GLU output);  out: bf16 [B, Tout, C]
//   out[t, c] = swish(norm(bias[c] + sum_j w[c, j] * in[t - lpad + j, c]))
//   rows with index < 0 read pad_left[c] (causal offline: the value GLU produces for a zero input
//   row, because the reference left-pads *before* pointwise_conv1, convolution.py:108-110),
//   rows >= Tin read 0 (symmetric padding of the depthwise conv itself, convolution.py:47-54).
//   Chunk mode passes Tin = lorder + Tout with lpad = 0 ("valid" convolution over [cache ; chunk]).
// One block per (b, TT-frame tile), C * SPLIT threads: thread = (channel, 1/SPLIT of the tile's rows) for the conv -- the
// K-1 halo rows are shared through smem, so splitting the rows costs no extra global traffic -- and warp = row for the norm.
// The kernel is a latency chain (ncu r2_small: 1061 instructions per warp at 15 cycles each, 40 % issue utilisation), so
// the work per thread is kept short and every constant (taps, bias, gamma, beta) is loaded before the grid dependency
// resolves, underneath the predecessor's tail.
// ------------------------------------------------------------------------------------------------
template <int C, int TT, int K, int SPLIT>
__global__ void __launch_bounds__(C * SPLIT) dwconv_norm_swish_kernel(const __nv_bfloat16* __restrict__ g,
                                                              const float* __restrict__ w,  // [C, K]
                                                              const float* __restrict__ bias,
                                                              const float* __restrict__ pad_left,  // [C] or null
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int use_layer_norm,
                                                              __nv_bfloat16* __restrict__ out, int Tin, int Tout,
                                                              int lpad, float eps, const int* __restrict__ zero_lens) {
  extern __shared__ uint8_t dsm[];
  constexpr int ROWS = TT + K - 1;
  constexpr int NT = C * SPLIT;
  constexpr int RPT = TT / SPLIT;  // conv rows per thread
  constexpr int PER = C / 32;
  static_assert(TT % SPLIT == 0 && NT / 32 == TT, "one warp per output row in the norm phase");
  __nv_bfloat16* sin = reinterpret_cast<__nv_bfloat16*>(dsm);                 // [ROWS][C]
  float* sout = reinterpret_cast<float*>(dsm + (size_t)ROWS * C * 2);         // [TT][C]
  const int tiles = (Tout + TT - 1) / TT;
  const int b = blockIdx.x / tiles;
  const int t0 = (blockIdx.x - b * tiles) * TT;
  const int c = threadIdx.x & (C - 1);
  const int part = threadIdx.x / C;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- constants (weights of the model: safe to read before griddepcontrol.wait) ----
  float wk[K];
#pragma unroll
  for (int j = 0; j < K; ++j) wk[j] = __ldg(w + c * K + j);
  const float bs = __ldg(bias + c);
  const float sc = use_layer_norm ? 1.f : __ldg(gamma + c);
  const float sh = use_layer_norm ? 0.f : __ldg(beta + c);
  float gm[PER], bt[PER];
  if (use_layer_norm) {
#pragma unroll
    for (int i = 0; i < PER; i += 4) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(gamma + lane * PER + i));
      const float4 d = __ldg(reinterpret_cast<const float4*>(beta + lane * PER + i));
      gm[i] = a.x, gm[i + 1] = a.y, gm[i + 2] = a.z, gm[i + 3] = a.w;
      bt[i] = d.x, bt[i + 1] = d.y, bt[i + 2] = d.z, bt[i + 3] = d.w;
    }
  }
  pdl_wait();
  pdl_launch_dependents();
  const int zlen = zero_lens != nullptr ? __ldg(zero_lens + b) : 0x7fffffff;
  // coalesced load of the input window rows [t0 - lpad, t0 - lpad + ROWS)
  {
    constexpr int NV = (ROWS * (C / 8) + NT - 1) / NT;
    uint4 v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = threadIdx.x + k * NT;
      const int r = i / (C / 8), seg = i - r * (C / 8);
      const int ti = t0 - lpad + r;
      v[k] = make_uint4(0, 0, 0, 0);
      if (r < ROWS) {
        if (ti >= 0 && ti < Tin) {
          v[k] = *reinterpret_cast<const uint4*>(g + ((size_t)b * Tin + ti) * C + seg * 8);
        } else if (ti < 0 && pad_left != nullptr) {
          const float4 p0 = __ldg(reinterpret_cast<const float4*>(pad_left + seg * 8));
          const float4 p1 = __ldg(reinterpret_cast<const float4*>(pad_left + seg * 8) + 1);
          v[k] = make_uint4(pack_bf16x2(p0.x, p0.y), pack_bf16x2(p0.z, p0.w), pack_bf16x2(p1.x, p1.y),
                            pack_bf16x2(p1.z, p1.w));
        }
      }
    }
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int i = threadIdx.x + k * NT;
      const int r = i / (C / 8), seg = i - r * (C / 8);
      if (r < ROWS) *reinterpret_cast<uint4*>(sin + (size_t)r * C + seg * 8) = v[k];
    }
  }
  __syncthreads();
  // depthwise conv: thread = (channel, rows [part*RPT, (part+1)*RPT)), sliding K-wide register window over time
  {
    const int r0 = part * RPT;
    float win[K];
#pragma unroll
    for (int j = 0; j < K - 1; ++j) win[j] = __bfloat162float(sin[(size_t)(r0 + j) * C + c]);
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      win[(u + K - 1) % K] = __bfloat162float(sin[(size_t)(r0 + u + K - 1) * C + c]);
      float acc = bs;
#pragma unroll
      for (int j = 0; j < K; ++j) acc = fmaf(wk[j], win[(u + j) % K], acc);
      sout[(r0 + u) * C + c] = acc * sc + sh;
    }
  }
  __syncthreads();
  // norm + swish, warp = row
  const int t = warp;
  if (t0 + t >= Tout) return;
  float v[PER];
#pragma unroll
  for (int i = 0; i < PER; i += 4) {
    const float4 a = *reinterpret_cast<const float4*>(sout + t * C + lane * PER + i);
    v[i] = a.x, v[i + 1] = a.y, v[i + 2] = a.z, v[i + 3] = a.w;
  }
  if (use_layer_norm) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) s += v[i];
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      const float d = v[i] - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
#pragma unroll
    for (int i = 0; i < PER; ++i) v[i] = (v[i] - mean) * rstd * gm[i] + bt[i];
  }
  const bool zero = (t0 + t) >= zlen;
  uint32_t pk[PER / 2];
#pragma unroll
  for (int i = 0; i < PER / 2; ++i) pk[i] = zero ? 0u : pack_bf16x2(swish_rcp(v[2 * i]), swish_rcp(v[2 * i + 1]));
  uint4* dst = reinterpret_cast<uint4*>(out + ((size_t)b * Tout + t0 + t) * C + lane * PER);
#pragma unroll
  for (int i = 0; i < PER / 8; ++i) dst[i] = make_uint4(pk[4 * i], pk[4 * i + 1], pk[4 * i + 2], pk[4 * i + 3]);
}

cudaError_t launch_dwconv_norm_swish(const __nv_bfloat16* g, const float* w, const float* bias, const float* pad_left,
                                     const float* gamma, const float* beta, int use_layer_norm, __nv_bfloat16* out,
                                     int B, int Tin, int Tout, int C, int K, int lpad, float eps, const int* zero_lens,
                                     cudaStream_t st) {
  if (C != 256) return cudaErrorInvalidValue;
  // 1 (default): 256 threads per 8-row tile; 2: 512 threads per 16-row tile (two row halves per channel). Both put one warp on
  // each output row in the norm phase. C2 live: 14.0 / 14.7 us per launch (16.9 before: 16-row tiles, 256 threads, constants
  // loaded after the grid dependency); env PPASR_B200_DW_SPLIT.
  static const int split = [] {
    const char* e = std::getenv("PPASR_B200_DW_SPLIT");
    return e ? std::atoi(e) : 1;
  }();
#define PPASR_DW_LAUNCH(KK, TT, SPLIT)                                                                        \
  {                                                                                                           \
    const int tiles = (Tout + TT - 1) / TT;                                                                   \
    const size_t smem = (size_t)(TT + KK - 1) * C * 2 + (size_t)TT * C * 4;                                   \
    auto kern = dwconv_norm_swish_kernel<256, TT, KK, SPLIT>;                                                 \
    cudaError_t le = launch_pdl(kern, dim3(B * tiles), dim3(256 * SPLIT), smem, st, g, w, bias, pad_left, gamma, beta, \
                                use_layer_norm, out, Tin, Tout, lpad, eps, zero_lens);                        \
    if (le != cudaSuccess) return le;                                                                         \
  }
#define PPASR_DW_LAUNCH_K(KK) \
  if (split == 2) PPASR_DW_LAUNCH(KK, 16, 2) else PPASR_DW_LAUNCH(KK, 8, 1)
  if (K == 15) { PPASR_DW_LAUNCH_K(15) } else if (K == 31) { PPASR_DW_LAUNCH_K(31) } else if (K == 7) { PPASR_DW_LAUNCH_K(7) } else return cudaErrorInvalidValue;
#undef PPASR_DW_LAUNCH_K
#undef PPASR_DW_LAUNCH
  count_launch();
  return cudaGetLastError();
}

// pad_left[c] = bf16(GLU(bias_a[c], bias_g[c])) computed with exactly the epilogue's arithmetic
// (interleaved bias layout: [2c] = a, [2c+1] = gate).
__global__ void glu_pad_kernel(const float* __restrict__ bias_il, float* __restrict__ pad, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) pad[c] = __bfloat162float(__float2bfloat16_rn(bias_il[2 * c] * sigmoid_f(bias_il[2 * c + 1])));
}
cudaError_t launch_glu_pad(const float* bias_il, float* pad, int C, cudaStream_t st) {
  glu_pad_kernel<<<(C + 127) / 128, 128, 0, st>>>(bias_il, pad, C);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Streaming conv cache: one block per stream. ycat[b] = [cache[b] ; y[b]] ; cache[b] <- tail(ycat[b], lorder)
// ------------------------------------------------------------------------------------------------
__global__ void conv_cache_concat_kernel(__nv_bfloat16* __restrict__ cache, const __nv_bfloat16* __restrict__ y,
                                         __nv_bfloat16* __restrict__ ycat, int T, int lorder, int C,
                                         const int* __restrict__ slots) {
  extern __shared__ uint4 s_old[];  // lorder * C / 8
  const int b = blockIdx.x;
  const int vec = C / 8;
  uint4* cb = reinterpret_cast<uint4*>(cache + (size_t)(slots ? slots[b] : b) * lorder * C);
  const uint4* yb = reinterpret_cast<const uint4*>(y + (size_t)b * T * C);
  uint4* ob = reinterpret_cast<uint4*>(ycat + (size_t)b * (lorder + T) * C);
  for (int i = threadIdx.x; i < lorder * vec; i += blockDim.x) s_old[i] = cb[i];
  __syncthreads();
  for (int i = threadIdx.x; i < (lorder + T) * vec; i += blockDim.x) {
    const int r = i / vec;
    const uint4 v = (r < lorder) ? s_old[i] : yb[i - lorder * vec];
    ob[i] = v;
    const int cr = r - T;  // row of the new cache = ycat row T + cr
    if (cr >= 0) cb[cr * vec + (i - r * vec)] = v;
  }
}
cudaError_t launch_conv_cache_concat(__nv_bfloat16* cache, const __nv_bfloat16* y, __nv_bfloat16* ycat, int B, int T,
                                     int lorder, int C, cudaStream_t st, const int* slots) {
  if (B <= 0) return cudaSuccess;
  conv_cache_concat_kernel<<<B, 256, (size_t)lorder * C * 2, st>>>(cache, y, ycat, T, lorder, C, slots);
  count_launch();
  return cudaGetLastError();
}

// att_cache export: out[h, j, 0:64] = k[h, k0 + j, :], out[h, j, 64:128] = v[h, :, k0 + j]  (fp32, stream 0 of the batch)
__global__ void export_att_cache_kernel(const __nv_bfloat16* __restrict__ kk, const __nv_bfloat16* __restrict__ vt,
                                        float* __restrict__ out, int H, int Tcap, int Tcapp, int k0, int t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= H * t * 128) return;
  const int d = i & 127, j = (i >> 7) % t, h = (i >> 7) / t;
  float v;
  if (d < 64)
    v = __bfloat162float(kk[((size_t)h * Tcap + k0 + j) * 64 + d]);
  else
    v = __bfloat162float(vt[((size_t)h * 64 + (d - 64)) * Tcapp + k0 + j]);
  out[i] = v;
}
cudaError_t launch_export_att_cache(const __nv_bfloat16* kk, const __nv_bfloat16* vt, float* out, int H, int Tcap,
                                    int Tcapp, int k0, int t, cudaStream_t st) {
  const int n = H * t * 128;
  if (n <= 0) return cudaSuccess;
  export_att_cache_kernel<<<(n + 255) / 256, 256, 0, st>>>(kk, vt, out, H, Tcap, Tcapp, k0, t);
  count_launch();
  return cudaGetLastError();
}
// cnn_cache export: out[c, j] = cache[j, c]   (reference layout [1, C, lorder])
__global__ void export_cnn_cache_kernel(const __nv_bfloat16* __restrict__ cache, float* __restrict__ out, int lorder,
                                        int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= lorder * C) return;
  const int j = i % lorder, c = i / lorder;
  out[i] = __bfloat162float(cache[(size_t)j * C + c]);
}
cudaError_t launch_export_cnn_cache(const __nv_bfloat16* cache, float* out, int lorder, int C, cudaStream_t st) {
  export_cnn_cache_kernel<<<(lorder * C + 255) / 256, 256, 0, st>>>(cache, out, lorder, C);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Row soft-max of the CTC logits: probs[row, 0:V] (dense) = softmax(logits[row, 0:V]) (ld = ldl)
//   reference: ppasr/model_utils/loss/ctc.py:62-70
// One block per row; values stay in registers between the passes.
// ------------------------------------------------------------------------------------------------
template <int THREADS, int MAXPER>
__global__ void __launch_bounds__(THREADS) softmax_rows_kernel(const float* __restrict__ logits, int ldl,
                                                               float* __restrict__ probs, int V) {
  __shared__ float red[THREADS / 32];
  __shared__ float bcast;
  const size_t row = blockIdx.x;
  const float* src = logits + row * ldl;
  float v[MAXPER];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAXPER; ++i) {
    const int j = threadIdx.x + i * THREADS;
    v[i] = (j < V) ? src[j] : -INFINITY;
    m = fmaxf(m, v[i]);
  }
  m = warp_max(m);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < THREADS / 32) ? red[threadIdx.x] : -INFINITY;
    t = warp_max(t);
    if (threadIdx.x == 0) bcast = t;
  }
  __syncthreads();
  m = bcast;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXPER; ++i) {
    v[i] = expf(v[i] - m);  // exp(-inf) = 0 for the tail
    s += v[i];
  }
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = (threadIdx.x < THREADS / 32) ? red[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) bcast = t;
  }
  __syncthreads();
  const float inv = 1.0f / bcast;
  float* dst = probs + row * V;
#pragma unroll
  for (int i = 0; i < MAXPER; ++i) {
    const int j = threadIdx.x + i * THREADS;
    if (j < V) dst[j] = v[i] * inv;
  }
}

cudaError_t launch_softmax_rows(const float* logits, int ldl, float* probs, int M, int V, cudaStream_t st) {
  if (M <= 0) return cudaSuccess;
  if (V <= 256 * 8)
    softmax_rows_kernel<256, 8><<<M, 256, 0, st>>>(logits, ldl, probs, V);
  else if (V <= 256 * 20)
    softmax_rows_kernel<256, 20><<<M, 256, 0, st>>>(logits, ldl, probs, V);
  else if (V <= 512 * 32)
    softmax_rows_kernel<512, 32><<<M, 512, 0, st>>>(logits, ldl, probs, V);
  else
    return cudaErrorInvalidValue;
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// CTC greedy decode
//   reference: ppasr/decoders/ctc_greedy_decoder.py:21-31 (argmax -> collapse repeats -> drop blank;
//   score = mean of the max-probabilities of the non-blank frames)
// ------------------------------------------------------------------------------------------------
// (a) frame arg-max over a dense probability matrix probs[rows, V] (first maximum wins, like np.argmax).
//     One warp per frame, 4 frames per block; the row is streamed once with 16-byte loads where the
//     row start allows (rows are only 4-byte aligned because V is odd in general).
__global__ void __launch_bounds__(128) argmax_rows_kernel(const float* __restrict__ probs, int V, int rows,
                                                          int* __restrict__ idx, float* __restrict__ maxp) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* src = probs + (size_t)row * V;
  float bm = -INFINITY;
  int bi = 0x7fffffff;
  auto upd = [&](float v, int j) {
    if (v > bm || (v == bm && j < bi)) {
      bm = v;
      bi = j;
    }
  };
  // peel to 16-byte alignment
  const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
  const int head = mis ? min(4 - mis, V) : 0;
  if (lane < head) upd(src[lane], lane);
  const float4* v4 = reinterpret_cast<const float4*>(src + head);
  const int n4 = (V - head) >> 2;
  int i = lane;
  for (; i + 96 < n4; i += 128) {  // 4 independent 16-byte loads in flight per lane
    const float4 a = __ldcs(v4 + i), b = __ldcs(v4 + i + 32), c = __ldcs(v4 + i + 64), d = __ldcs(v4 + i + 96);
    int j = head + 4 * i;
    upd(a.x, j), upd(a.y, j + 1), upd(a.z, j + 2), upd(a.w, j + 3);
    j += 128;
    upd(b.x, j), upd(b.y, j + 1), upd(b.z, j + 2), upd(b.w, j + 3);
    j += 128;
    upd(c.x, j), upd(c.y, j + 1), upd(c.z, j + 2), upd(c.w, j + 3);
    j += 128;
    upd(d.x, j), upd(d.y, j + 1), upd(d.z, j + 2), upd(d.w, j + 3);
  }
  for (; i < n4; i += 32) {
    const float4 a = __ldcs(v4 + i);
    const int j = head + 4 * i;
    upd(a.x, j), upd(a.y, j + 1), upd(a.z, j + 2), upd(a.w, j + 3);
  }
  const int tail0 = head + 4 * n4;
  if (tail0 + lane < V) upd(src[tail0 + lane], tail0 + lane);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, bm, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (om > bm || (om == bm && oi < bi)) {
      bm = om;
      bi = oi;
    }
  }
  if (lane == 0) {
    idx[row] = bi;
    maxp[row] = bm;
  }
}

cudaError_t launch_argmax_rows(const float* probs, int V, int rows, int* idx, float* maxp, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  argmax_rows_kernel<<<(rows + 3) / 4, 128, 0, st>>>(probs, V, rows, idx, maxp);
  count_launch();
  return cudaGetLastError();
}

// (b) fused-head variant: combine the per-tile partial statistics written by EpiCtcStats.
//     idx = arg-max of the logits (first maximum), maxp = softmax probability of it = 1 / sum exp(l - max).
__global__ void ctc_stats_finalize_kernel(const float* __restrict__ pmax, const int* __restrict__ parg,
                                          const float* __restrict__ psum, int parts, int rows, int* __restrict__ idx,
                                          float* __restrict__ maxp) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float bm = -INFINITY;
  int bi = 0x7fffffff;
  for (int p = lane; p < parts; p += 32) {
    const float m = pmax[(size_t)row * parts + p];
    const int a = parg[(size_t)row * parts + p];
    if (m > bm || (m == bm && a < bi)) {
      bm = m;
      bi = a;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, bm, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (om > bm || (om == bm && oi < bi)) {
      bm = om;
      bi = oi;
    }
  }
  float s = 0.f;
  for (int p = lane; p < parts; p += 32) {
    const float m = pmax[(size_t)row * parts + p];
    if (m > -INFINITY) s += psum[(size_t)row * parts + p] * expf(m - bm);
  }
  s = warp_sum(s);
  if (lane == 0) {
    idx[row] = bi;
    maxp[row] = 1.0f / s;
  }
}

cudaError_t launch_ctc_stats_finalize(const float* pmax, const int* parg, const float* psum, int parts, int rows,
                                      int* idx, float* maxp, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  cudaError_t le = launch_pdl(ctc_stats_finalize_kernel, dim3((rows + 3) / 4), dim3(128), (size_t)0, st, pmax, parg, psum, parts, rows, idx, maxp);
  if (le != cudaSuccess) return le;
  count_launch();
  return cudaGetLastError();
}

// (c) blank-collapse: one warp per utterance. ids_out[b, 0:n] = collapsed non-blank ids, out_len[b] = n,
//     score[b] = (sequential fp32 sum of maxp over non-blank frames) / count   (0 if none); the
//     sequential order reproduces Python's sum() over np.float32 values bit-for-bit.
//     prev_id (nullable): last frame id of the previous chunk for streaming collapse (-1 = none).
__global__ void ctc_collapse_kernel(const int* __restrict__ idx, const float* __restrict__ maxp, int B, int T,
                                    const int* __restrict__ frame_lens, int blank, int* __restrict__ ids_out,
                                    int ld_out, int* __restrict__ out_len, float* __restrict__ score,
                                    float* __restrict__ score_sum, int* __restrict__ score_cnt) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= B) return;
  const int n_frames = frame_lens ? min(T, frame_lens[b]) : T;
  const int* id = idx + (size_t)b * T;
  int n_out = 0;
  for (int t0 = 0; t0 < n_frames; t0 += 32) {
    const int t = t0 + lane;
    int cur = blank, prev = -1;
    if (t < n_frames) {
      cur = id[t];
      prev = (t > 0) ? id[t - 1] : -1;
    }
    const bool keep = (t < n_frames) && (cur != blank) && (cur != prev);
    const unsigned mask = __ballot_sync(0xffffffffu, keep);
    if (keep) ids_out[(size_t)b * ld_out + n_out + __popc(mask & ((1u << lane) - 1))] = cur;
    n_out += __popc(mask);
  }
  // sequential (in frame order) fp32 sum, computed redundantly by every lane from warp-wide loads
  float s = 0.f;
  int cnt = 0;
  const float* mp = maxp + (size_t)b * T;
  for (int t0 = 0; t0 < n_frames; t0 += 32) {
    const int t = t0 + lane;
    const float pv = (t < n_frames) ? mp[t] : 0.f;
    const bool nb = (t < n_frames) && (id[t] != blank);
    const unsigned nbmask = __ballot_sync(0xffffffffu, nb);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float v = __shfl_sync(0xffffffffu, pv, k);
      if ((nbmask >> k) & 1u) {
        s += v;
        ++cnt;
      }
    }
  }
  if (lane == 0) {
    out_len[b] = n_out;
    score[b] = cnt > 0 ? s / (float)cnt : 0.f;
    if (score_sum) score_sum[b] = s;
    if (score_cnt) score_cnt[b] = cnt;
  }
}

cudaError_t launch_ctc_collapse(const int* idx, const float* maxp, int B, int T, const int* frame_lens, int blank,
                                int* ids_out, int ld_out, int* out_len, float* score, float* score_sum,
                                int* score_cnt, cudaStream_t st) {
  if (B <= 0) return cudaSuccess;
  cudaError_t le = launch_pdl(ctc_collapse_kernel, dim3((B + 3) / 4), dim3(128), (size_t)0, st, idx, maxp, B, T, frame_lens, blank,
                              ids_out, ld_out, out_len, score, score_sum, score_cnt);
  if (le != cudaSuccess) return le;
  count_launch();
  return cudaGetLastError();
}

}  // namespace ppasr

// ---------------------------------------------------------------------------------------------------------------
// Squeezeformer time-reduction layer, depthwise part (reference: ppasr/model_utils/squeezeformer/time_reduction.py
// :61-84 TimeReductionLayer1D k=5 s=2 pad=3, :183-206 TimeReductionLayerStream k=1 s=2 pad=0). The input is zeroed at
// the pad frames first (masked_fill(xs, mask_pad.equal(0), 0)); the pointwise conv that follows is a GEMM.
// ---------------------------------------------------------------------------------------------------------------
namespace ppasr {

void count_launch();

__global__ void time_reduce_dw_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                      const int* __restrict__ lens, __nv_bfloat16* __restrict__ out, int T, int Tr, int C, int K,
                                      int pad) {
  pdl_wait();
  pdl_launch_dependents();
  const int b = blockIdx.y;
  const int tr = blockIdx.x * blockDim.y + threadIdx.y;
  if (tr >= Tr) return;
  const int len = lens ? lens[b] : T;
  for (int c = threadIdx.x * 4; c < C; c += blockDim.x * 4) {
    float4 acc = *reinterpret_cast<const float4*>(bias + c);
    for (int k = 0; k < K; ++k) {
      const int t = 2 * tr + k - pad;
      if (t < 0 || t >= T || t >= len) continue;
      const float4 v = *reinterpret_cast<const float4*>(x + ((size_t)b * T + t) * C + c);
      acc.x = fmaf(w[(c + 0) * K + k], v.x, acc.x);
      acc.y = fmaf(w[(c + 1) * K + k], v.y, acc.y);
      acc.z = fmaf(w[(c + 2) * K + k], v.z, acc.z);
      acc.w = fmaf(w[(c + 3) * K + k], v.w, acc.w);
    }
    *reinterpret_cast<uint2*>(out + ((size_t)b * Tr + tr) * C + c) = make_uint2(pack_bf16x2(acc.x, acc.y), pack_bf16x2(acc.z, acc.w));
  }
}

cudaError_t launch_time_reduce_dw(const float* x, const float* w, const float* bias, const int* lens, __nv_bfloat16* out,
                                  int B, int T, int Tr, int C, int K, int pad, cudaStream_t st) {
  if (C % 4 != 0 || B <= 0 || Tr <= 0) return cudaErrorInvalidValue;
  dim3 block(64, 4);
  dim3 grid((Tr + 3) / 4, B);
  cudaError_t le = launch_pdl(time_reduce_dw_kernel, grid, block, (size_t)0, st, x, w, bias, lens, out, T, Tr, C, K, pad);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

__global__ void halve_lens_kernel(const int* __restrict__ a, int* __restrict__ o, int B) {
  pdl_wait();
  pdl_launch_dependents();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) o[i] = (a[i] + 1) / 2;
}

cudaError_t launch_halve_lens(const int* lens_in, int* lens_out, int B, cudaStream_t st) {
  cudaError_t le = launch_pdl(halve_lens_kernel, dim3((B + 127) / 128), dim3(128), (size_t)0, st, lens_in, lens_out, B);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace ppasr
