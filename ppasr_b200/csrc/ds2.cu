// DeepSpeech2 kernels for sm_100a (reference: ppasr/model_utils/deepspeech2/conv.py:8-21, encoder.py:37-55,86-93).
//
//  * ds2_conv_kernel       GlobalCMVN + Conv2d(1->32,k3,s2)+ReLU + Conv2d(32->32,k3,s2)+ReLU -> bf16 [B*T', Kpad]
//                          (feature index c*F2 + f like x.transpose([0,2,1,3]).reshape, conv.py:19), one CTA per (b, t').
//                          0.35 MFLOP per output frame: CUDA cores, everything staged in shared memory.
//  * rnn_layer_kernel      the recurrent half of paddle.nn.LSTM / GRU: a persistent, weight-stationary kernel. The input
//                          half  x W_ih^T + b_ih  of all time steps is one tcgen05 GEMM (gemm.cuh) done beforehand; the
//                          recurrence is a chain of T' dependent [B,H] x [H,G*H] products that no GEMM shape can hide, so
//                          each CTA keeps its 64 (LSTM) / 48 (GRU) rows of W_hh in shared memory for the whole sequence
//                          (H/16 CTAs per direction, both directions of a bidirectional layer run concurrently), h_t is
//                          exchanged through L2 and the CTAs of a direction meet at a software grid barrier once per step
//                          (cooperative launch guarantees co-residency). Cell state c lives in shared memory.
//  * row_layernorm_kernel  LayerNorm over 1024 / 2048 features, fp32 -> bf16 operand of the next GEMM.
#include <cooperative_groups.h>

#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace ppasr {

void count_launch();

// ----------------------------------------------------------------------------------------------------------------
constexpr int DS2_C = 32;

__global__ void __launch_bounds__(256)
ds2_conv_kernel(const float* __restrict__ feats, const float* __restrict__ mean, const float* __restrict__ istd,
                const float* __restrict__ w1, const float* __restrict__ b1, const float* __restrict__ w2t,
                const float* __restrict__ b2, __nv_bfloat16* __restrict__ out, int T, int F, int Tp, int F1, int F2, int Kpad) {
  extern __shared__ float sm[];
  float* xin = sm;                       // [7][F]
  float* c1 = xin + 7 * F;               // [3][F1][32]
  float* w2s = c1 + 3 * F1 * DS2_C;      // [288][32]  (cin*9 + k major, cout minor)
  const int b = blockIdx.y, tp = blockIdx.x;
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_launch_dependents();
  for (int i = tid; i < 7 * F; i += blockDim.x) {
    const int r = i / F, f = i - r * F;
    const int t = 4 * tp + r;
    xin[i] = t < T ? (feats[((size_t)b * T + t) * F + f] - mean[f]) * istd[f] : 0.f;
  }
  for (int i = tid; i < 288 * DS2_C; i += blockDim.x) w2s[i] = w2t[i];
  __syncthreads();
  // conv1 rows 2t'..2t'+2
  for (int i = tid; i < 3 * F1 * DS2_C; i += blockDim.x) {
    const int c = i & 31;
    const int f1 = (i >> 5) % F1;
    const int r = (i >> 5) / F1;
    float acc = b1[c];
#pragma unroll
    for (int ki = 0; ki < 3; ++ki)
#pragma unroll
      for (int kj = 0; kj < 3; ++kj) acc = fmaf(w1[c * 9 + ki * 3 + kj], xin[(2 * r + ki) * F + 2 * f1 + kj], acc);
    c1[i] = fmaxf(acc, 0.f);
  }
  __syncthreads();
  // conv2: thread = (cout = lane, f' = warp, warp + 8, ...)
  const int co = tid & 31;
  __nv_bfloat16* orow = out + ((size_t)b * Tp + tp) * Kpad;
  for (int f2 = tid >> 5; f2 < F2; f2 += (blockDim.x >> 5)) {
    float acc = b2[co];
    for (int kh = 0; kh < 3; ++kh)
      for (int kw = 0; kw < 3; ++kw) {
        const float* src = c1 + (kh * F1 + 2 * f2 + kw) * DS2_C;
        const float* wk = w2s + (kh * 3 + kw) * DS2_C + co;
#pragma unroll 8
        for (int ci = 0; ci < DS2_C; ++ci) acc = fmaf(wk[ci * 9 * DS2_C], src[ci], acc);
      }
    orow[co * F2 + f2] = __float2bfloat16_rn(fmaxf(acc, 0.f));
  }
  for (int i = DS2_C * F2 + tid; i < Kpad; i += blockDim.x) orow[i] = __float2bfloat16_rn(0.f);
}

cudaError_t launch_ds2_conv(const float* feats, const float* mean, const float* istd, const float* w1, const float* b1,
                            const float* w2t, const float* b2, __nv_bfloat16* out, int B, int T, int F, int Tp, int Kpad,
                            cudaStream_t st) {
  const int F1 = (F - 1) / 2, F2 = (F1 - 1) / 2;
  const size_t smem = (size_t)(7 * F + 3 * F1 * DS2_C + 288 * DS2_C) * 4;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(ds2_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (smem > 100 * 1024) return cudaErrorInvalidValue;
  cudaError_t le = launch_pdl(ds2_conv_kernel, dim3(Tp, B), dim3(256), smem, st, feats, mean, istd, w1, b1, w2t, b2, out, T,
                              F, Tp, F1, F2, Kpad);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
struct RnnParams {
  const float* xg;            // [B*T, nd*G*H] = x W_ih^T + b_ih, column d*G*H + g*H + u
  const __nv_bfloat16* whh;   // [nd][G*H][H]
  const float* bhh;           // [nd][G*H]
  const int* lens;            // [B] valid steps
  const float* h_init;        // [nd][B][H] or null (zeros)
  const float* c_init;        // [nd][B][H] or null
  float* h_final;             // [nd][B][H] or null
  float* c_final;             // [nd][B][H] or null (LSTM)
  float* out;                 // [B*T, nd*H]
  float* hbuf;                // [2][nd][B][H] exchange buffer
  unsigned* bar;              // [nd] zeroed before the launch
  int B, T, H, nd, gru;
};

constexpr int RNN_U = 16;       // hidden units per CTA
constexpr int RNN_THREADS = 256;
constexpr int RNN_BT = 8;       // batch tile held in shared memory
constexpr int RNN_HP = 1024;    // pitch of the h tile in shared memory (H <= 1024)

// recurrent mat-vec for NB batch slots: warp w owns rows w, w+8, ... in groups of 4 (h values are loaded once per group);
// lanes split K in bf16 pairs; fp32 accumulation; gs[row][slot] = W_hh[row] . h[slot] + b_hh[row]
template <int NB>
DEVINL void rnn_matvec(const __nv_bfloat16* ws, const float* hs, float* gs, const float* bs, int H, int rows, int warp, int lane) {
  for (int r0 = warp * 4; r0 < rows; r0 += (RNN_THREADS / 32) * 4) {
    float acc[4][NB];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) acc[i][bb] = 0.f;
    const __nv_bfloat162* wr = reinterpret_cast<const __nv_bfloat162*>(ws + (size_t)r0 * H);
    const int hp2 = H / 2;
#pragma unroll 4
    for (int k2 = lane; k2 < hp2; k2 += 32) {
      float2 wv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) wv[i] = __bfloat1622float2(wr[i * hp2 + k2]);
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) {
        const float2 hv = *reinterpret_cast<const float2*>(hs + bb * RNN_HP + 2 * k2);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][bb] = fmaf(wv[i].y, hv.y, fmaf(wv[i].x, hv.x, acc[i][bb]));
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int bb = 0; bb < NB; ++bb) {
        const float v = warp_sum(acc[i][bb]);
        if (lane == 0) gs[(r0 + i) * RNN_BT + bb] = v + bs[r0 + i];
      }
  }
}

template <int G>
__global__ void __launch_bounds__(RNN_THREADS, 1) rnn_layer_kernel(const RnnParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int H = p.H, B = p.B, T = p.T;
  const int nchunk = H / RNN_U;
  const int d = blockIdx.x / nchunk;
  const int u0 = (blockIdx.x - d * nchunk) * RNN_U;
  constexpr int ROWS = G * RNN_U;
  __nv_bfloat16* ws = reinterpret_cast<__nv_bfloat16*>(smem_raw);                  // [ROWS][H]
  float* hs = reinterpret_cast<float*>(smem_raw + (size_t)ROWS * H * 2);           // [RNN_BT][RNN_HP]
  float* gs = hs + RNN_BT * RNN_HP;                                                 // [ROWS][RNN_BT]
  float* cs = gs + ROWS * RNN_BT;                                                   // [RNN_U][B]   (LSTM cell state)
  float* bs = cs + RNN_U * B;                                                       // [ROWS] recurrent bias
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // stationary weights: row g*U + u  <-  W_hh[d][g*H + u0 + u][:]
  for (int i = tid; i < ROWS * (H / 8); i += RNN_THREADS) {
    const int r = i / (H / 8), k8 = i - r * (H / 8);
    const int g = r / RNN_U, u = r - g * RNN_U;
    reinterpret_cast<uint4*>(ws)[i] =
        *reinterpret_cast<const uint4*>(p.whh + ((size_t)d * G * H + (size_t)g * H + u0 + u) * H + k8 * 8);
  }
  for (int i = tid; i < ROWS; i += RNN_THREADS) {
    const int g = i / RNN_U, u = i - g * RNN_U;
    bs[i] = p.bhh[(size_t)d * G * H + (size_t)g * H + u0 + u];
  }
  for (int i = tid; i < RNN_U * B; i += RNN_THREADS) {
    const int u = i / B, b = i - u * B;
    cs[i] = (p.c_init && G == 4) ? p.c_init[((size_t)d * B + b) * H + u0 + u] : 0.f;
  }
  // h_{-1} of this CTA's units -> exchange buffer 0
  for (int i = tid; i < RNN_U * B; i += RNN_THREADS) {
    const int u = i / B, b = i - u * B;
    p.hbuf[((size_t)(0 * p.nd + d) * B + b) * H + u0 + u] = p.h_init ? p.h_init[((size_t)d * B + b) * H + u0 + u] : 0.f;
  }
  __syncthreads();
  unsigned arrived = 0;
  auto dir_barrier = [&]() {
    __syncthreads();
    if (tid == 0) {
      // release our h writes (ordered before by the CTA barrier), then wait for every CTA of this direction
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.bar + d) : "memory");
      arrived += nchunk;
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p.bar + d) : "memory");
      } while (v < arrived);
    }
    __syncthreads();
  };
  dir_barrier();

  const int ldx = p.nd * G * H;
  for (int s = 0; s < T; ++s) {
    const int t = d == 1 ? T - 1 - s : s;
    const float* hcur = p.hbuf + (size_t)((s & 1) * p.nd + d) * B * H;
    float* hnext = p.hbuf + (size_t)(((s + 1) & 1) * p.nd + d) * B * H;
    for (int b0 = 0; b0 < B; b0 += RNN_BT) {
      const int nb = min(RNN_BT, B - b0);
      // gate pre-activations of this step (independent of h): issue the loads before the mat-vec
      const int pu = tid % RNN_U, pb = tid / RNN_U;  // point-wise owner: unit pu, batch slot pb
      float xgv[G];
      const bool powner = pb < nb;
      if (powner) {
        const float* xr = p.xg + ((size_t)(b0 + pb) * T + t) * ldx + (size_t)d * G * H + u0 + pu;
#pragma unroll
        for (int g = 0; g < G; ++g) xgv[g] = __ldcg(xr + (size_t)g * H);
      }
      for (int i = tid; i < nb * (H / 4); i += RNN_THREADS) {
        const int bb = i / (H / 4), k4 = i - bb * (H / 4);
        reinterpret_cast<float4*>(hs)[bb * (RNN_HP / 4) + k4] =
            __ldcg(reinterpret_cast<const float4*>(hcur + (size_t)(b0 + bb) * H) + k4);
      }
      // slots beyond nb are read by the next larger mat-vec instantiation (nb = 3, 5, 6, 7): keep them finite
      for (int i = nb * (H / 4) + tid; i < (nb <= 2 ? nb : nb <= 4 ? 4 : RNN_BT) * (H / 4); i += RNN_THREADS) {
        const int bb = i / (H / 4), k4 = i - bb * (H / 4);
        reinterpret_cast<float4*>(hs)[bb * (RNN_HP / 4) + k4] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      __syncthreads();
      if (nb == 1) rnn_matvec<1>(ws, hs, gs, bs, H, ROWS, warp, lane);
      else if (nb == 2) rnn_matvec<2>(ws, hs, gs, bs, H, ROWS, warp, lane);
      else if (nb <= 4) rnn_matvec<4>(ws, hs, gs, bs, H, ROWS, warp, lane);
      else rnn_matvec<8>(ws, hs, gs, bs, H, ROWS, warp, lane);
      __syncthreads();
      if (powner) {
        const int b = b0 + pb;
        const bool valid = t < __ldg(p.lens + b);
        const float hold = hs[pb * RNN_HP + u0 + pu];
        float hnew;
        if (G == 4) {
          // paddle.nn.LSTMCell: chunks (i, f, g, o)
          const float ig = sigmoid_precise(xgv[0] + gs[(0 * RNN_U + pu) * RNN_BT + pb]);
          const float fg = sigmoid_precise(xgv[1] + gs[(1 * RNN_U + pu) * RNN_BT + pb]);
          const float gg = tanhf(xgv[2] + gs[(2 * RNN_U + pu) * RNN_BT + pb]);
          const float og = sigmoid_precise(xgv[3] + gs[(3 * RNN_U + pu) * RNN_BT + pb]);
          const float cold = cs[pu * B + b];
          const float cnew = fg * cold + ig * gg;
          hnew = og * tanhf(cnew);
          if (valid) cs[pu * B + b] = cnew;
        } else {
          // paddle.nn.GRUCell: chunks (r, z, c); the reset gate multiplies (W_hc h + b_hc)
          const float rg = sigmoid_precise(xgv[0] + gs[(0 * RNN_U + pu) * RNN_BT + pb]);
          const float zg = sigmoid_precise(xgv[1] + gs[(1 * RNN_U + pu) * RNN_BT + pb]);
          const float cand = tanhf(xgv[2] + rg * gs[(2 * RNN_U + pu) * RNN_BT + pb]);
          hnew = (hold - cand) * zg + cand;
        }
        hnext[(size_t)b * H + u0 + pu] = valid ? hnew : hold;
        p.out[((size_t)b * T + t) * (p.nd * H) + (size_t)d * H + u0 + pu] = valid ? hnew : 0.f;
      }
      __syncthreads();
    }
    dir_barrier();
  }
  // final states (encoder.py:95-102)
  const float* hlast = p.hbuf + (size_t)((T & 1) * p.nd + d) * B * H;
  for (int i = tid; i < RNN_U * B; i += RNN_THREADS) {
    const int u = i / B, b = i - u * B;
    if (p.h_final) p.h_final[((size_t)d * B + b) * H + u0 + u] = __ldcg(hlast + (size_t)b * H + u0 + u);
    if (p.c_final && G == 4) p.c_final[((size_t)d * B + b) * H + u0 + u] = cs[u * B + b];
  }
}

cudaError_t launch_rnn_layer(const float* xg, const __nv_bfloat16* whh, const float* bhh, const int* lens, const float* h_init,
                             const float* c_init, float* h_final, float* c_final, float* out, float* hbuf, unsigned* bar,
                             int B, int T, int H, int nd, int gru, int num_sms, cudaStream_t st) {
  const int G = gru ? 3 : 4;
  if (H % 64 != 0 || B <= 0 || T <= 0 || nd < 1 || nd > 2) return cudaErrorInvalidValue;
  const int grid = nd * (H / RNN_U);
  if (grid > num_sms) return cudaErrorInvalidConfiguration;  // all CTAs must be co-resident (grid barrier)
  const size_t smem = (size_t)G * RNN_U * H * 2 + (size_t)RNN_BT * RNN_HP * 4 + (size_t)G * RNN_U * RNN_BT * 4 +
                      (size_t)RNN_U * B * 4 + (size_t)G * RNN_U * 4;
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(bar, 0, sizeof(unsigned) * 2, st);
  if (e != cudaSuccess) return e;
  RnnParams p;
  p.xg = xg, p.whh = whh, p.bhh = bhh, p.lens = lens, p.h_init = h_init, p.c_init = c_init, p.h_final = h_final;
  p.c_final = c_final, p.out = out, p.hbuf = hbuf, p.bar = bar, p.B = B, p.T = T, p.H = H, p.nd = nd, p.gru = gru;
  void* args[] = {&p};
  const void* fn = gru ? reinterpret_cast<const void*>(rnn_layer_kernel<3>) : reinterpret_cast<const void*>(rnn_layer_kernel<4>);
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(RNN_THREADS), args, smem, st);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
}

// ----------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
row_layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ b,
                     __nv_bfloat16* __restrict__ y, int N, float eps) {
  __shared__ float red[2][8];
  pdl_wait();
  pdl_launch_dependents();
  const size_t row = blockIdx.x;
  const float* xr = x + row * N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float v[8];
  float s = 0.f;
  int n = 0;
  for (int i = tid; i < N; i += 256) s += (v[n++] = xr[i]);
  s = warp_sum(s);
  if (lane == 0) red[0][warp] = s;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) tot += red[0][w];
  const float mean = tot / (float)N;
  float q = 0.f;
  for (int j = 0; j < n; ++j) q += (v[j] - mean) * (v[j] - mean);
  q = warp_sum(q);
  if (lane == 0) red[1][warp] = q;
  __syncthreads();
  float qt = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) qt += red[1][w];
  const float rstd = rsqrtf(qt / (float)N + eps);
  n = 0;
  for (int i = tid; i < N; i += 256) y[row * N + i] = __float2bfloat16_rn((v[n++] - mean) * rstd * g[i] + b[i]);
}

cudaError_t launch_row_layernorm(const float* x, const float* g, const float* b, __nv_bfloat16* y, int M, int N, float eps,
                                 cudaStream_t st) {
  if (N > 2048 || N <= 0 || M <= 0) return cudaErrorInvalidValue;
  cudaError_t le = launch_pdl(row_layernorm_kernel, dim3(M), dim3(256), (size_t)0, st, x, g, b, y, N, eps);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace ppasr
