// Fused subsampling front end for sm_100a: GlobalCMVN + Conv2d(1->256, 3x3, s2) + ReLU + Conv2d(256->256, 3x3, s2) + ReLU
// in ONE persistent tcgen05 kernel -- the first conv's output never exists in memory.
//
// Reference: ppasr/model_utils/utils/cmvn.py:29-32, ppasr/model_utils/conformer/subsampling.py:84-87,110-111.
//
// GEMM view of the second conv (as in gemm.cuh CONV mode): rows = padded raster R = (b*Th + t')*FH + f',
// N = 256 output channels, K = 9 taps x 256 input channels (36 k-blocks of 64). The A operand of k-block (tap, kc),
//   A[r][c] = relu(b1[c] + sum_{i,j} w1[c][i][j] * xn[b, 4t' + 2kh + i, 4f' + 2kw + j]),   c = 64 kc + 0..63,
// is *computed* by 16 producer warps (thread = (4 rows, 4 channels): 36 normalised inputs in registers, weights from
// smem, fp32 accumulate, one bf16 rounding -- the same arithmetic as the stand-alone conv1 kernel) and stored
// directly in the 128B-swizzled K-major layout the UMMA descriptor expects; the B operand (conv2 weights) streams in by
// TMA into its own ring. The normalised
// inputs of a tile (8 raster lines x 7 input rows x 84 columns) sit in a double-buffered smem patch: the next tile's
// raw features are prefetched with cp.async while the current tile is computed.
//   warp 0: TMA (B)   warp 1: MMA issuer   warp 2: TMEM alloc   warps 4..7: epilogue   warps 8..23: A producers
// Compared with conv1 -> phase images -> conv2 this removes 2 x 318 MB of HBM traffic and one launch; the CUDA-core
// work of conv1 (now recomputed per tap: 9/4 x) runs concurrently with the tensor-core MMAs of the previous stage.
#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace ppasr {

void count_launch();

constexpr int CF_BLOCK_M = 128;
constexpr int CF_BLOCK_N = 256;
constexpr int CF_A_STAGES = 4;   // A tiles written by the producer warps
constexpr int CF_B_STAGES = 3;   // conv2 weight tiles streamed by TMA
constexpr int CF_THREADS = 768;  // 4 control/idle warps + 4 epilogue warps + 16 producer warps
constexpr int CF_PRODUCERS = 512;
constexpr int CF_A_BYTES = CF_BLOCK_M * 64 * 2;   // 16 KB
constexpr int CF_B_BYTES = CF_BLOCK_N * 64 * 2;   // 32 KB
constexpr int CF_PATCH_LINES = 8;   // a 128-row tile touches at most 8 raster lines of FH = 20 rows
constexpr int CF_PATCH_COLS = 84;   // 4*19 + 2*2 + 2 = 82 is the last column a (discarded) pad row can touch
constexpr int CF_PATCH_ELEMS = CF_PATCH_LINES * 7 * CF_PATCH_COLS;  // [8 lines][7 input rows][84 cols] fp32
constexpr int CF_SMEM_B = CF_A_STAGES * CF_A_BYTES;
constexpr int CF_SMEM_W1 = CF_SMEM_B + CF_B_STAGES * CF_B_BYTES;  // [10][256] floats: w1[k][c] (k < 9) and b1[c] (k = 9)
constexpr int CF_SMEM_BIAS2 = CF_SMEM_W1 + 10 * 256 * 4;
constexpr int CF_SMEM_CMVN = CF_SMEM_BIAS2 + 256 * 4;             // mean[84], istd[84] (zero beyond F)
constexpr int CF_SMEM_PATCH = CF_SMEM_CMVN + 1024;                // two patches (current tile / prefetch of the next)
constexpr int CF_SMEM_BAR = CF_SMEM_PATCH + 2 * CF_PATCH_ELEMS * 4;
constexpr int CF_SMEM_TOTAL = CF_SMEM_BAR + 256 + 1024;

struct ConvFrontParams {
  const float* feats;  // [B, T, F]
  const float* mean;   // [F]
  const float* istd;   // [F]
  const float* w1;     // [256, 9]
  const float* b1;     // [256]
  const float* b2;     // [256]
  __nv_bfloat16* out;  // [B*Tout*Fout, 256]
  int B, T, F;
  int Th, FH;          // raster: Th time rows per utterance, FH = pitch (20)
  int Tout, Fout;      // valid outputs per utterance (T', 19)
  int M;               // B * Th * FH raster rows
  int num_tiles;
};

DEVINL void cp_async_f32_zfill(float* dst_smem, const float* src, bool valid) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(smem_u32(dst_smem)), "l"(src), "r"(valid ? 4 : 0) : "memory");
}

__global__ void __launch_bounds__(CF_THREADS, 1)
conv_front_kernel(const __grid_constant__ CUtensorMap tmap_w2, const ConvFrontParams p) {
  extern __shared__ uint8_t smem_raw[];
  // offset arithmetic (not an integer round trip) keeps the pointers in the shared address space: LDS/STS, not LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + CF_SMEM_B;
  float* s_w1 = reinterpret_cast<float*>(smem + CF_SMEM_W1);
  float* s_b2 = reinterpret_cast<float*>(smem + CF_SMEM_BIAS2);
  float* s_mean = reinterpret_cast<float*>(smem + CF_SMEM_CMVN);
  float* s_istd = s_mean + CF_PATCH_COLS;
  float* s_patch = reinterpret_cast<float*>(smem + CF_SMEM_PATCH);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + CF_SMEM_BAR);
  uint64_t* a_empty = a_full + CF_A_STAGES;
  uint64_t* b_full = a_empty + CF_A_STAGES;
  uint64_t* b_empty = b_full + CF_B_STAGES;
  uint64_t* tmem_full_bar = b_empty + CF_B_STAGES;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && elect_one()) tma_prefetch_desc(&tmap_w2);
  if (warp_idx == 1 && elect_one()) {
    for (int i = 0; i < CF_A_STAGES; ++i) {
      mbar_init(&a_full[i], CF_PRODUCERS / 32);  // one arrival per producer warp
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < CF_B_STAGES; ++i) {
      mbar_init(&b_full[i], 1);
      mbar_init(&b_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full_bar[i], 1);
      mbar_init(&tmem_empty_bar[i], 128);
    }
    fence_barrier_init();
    fence_proxy_async_smem();
  }
  if (warp_idx == 2) tmem_alloc<512>(tmem_ptr_smem);
  // conv1 weights / biases -> smem (constants: may be read before the grid dependency resolves)
  for (int i = threadIdx.x; i < 256 * 10; i += CF_THREADS) {
    const int c = i / 10, k = i - c * 10;
    s_w1[k * 256 + c] = k < 9 ? __ldg(p.w1 + c * 9 + k) : __ldg(p.b1 + c);
  }
  for (int i = threadIdx.x; i < 256; i += CF_THREADS) s_b2[i] = __ldg(p.b2 + i);
  for (int i = threadIdx.x; i < CF_PATCH_COLS; i += CF_THREADS) {
    s_mean[i] = i < p.F ? __ldg(p.mean + i) : 0.f;
    s_istd[i] = i < p.F ? __ldg(p.istd + i) : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_launch_dependents();

  if (warp_idx == 0) {
    // ===================== TMA producer of the conv2 weights =====================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < 36; ++kb) {
          mbar_wait(&b_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&b_full[stage], CF_B_BYTES);
          tma_load_2d(smem_b + stage * CF_B_BYTES, &tmap_w2, &b_full[stage], kb * 64, 0);
          if (++stage == CF_B_STAGES) stage = 0, phase ^= 1;
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_bf16(CF_BLOCK_M, CF_BLOCK_N);
      int sa = 0, sb = 0;
      uint32_t pa = 0, pb = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * CF_BLOCK_N;
        for (int kb = 0; kb < 36; ++kb) {
          mbar_wait(&b_full[sb], pb);
          mbar_wait(&a_full[sa], pa);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem_a + sa * CF_A_BYTES);
          const uint32_t b_addr = smem_u32(smem_b + sb * CF_B_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_d, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc,
                      (kb | k) != 0 ? 1u : 0u);
          umma_commit(&a_empty[sa]);
          umma_commit(&b_empty[sb]);
          if (++sa == CF_A_STAGES) sa = 0, pa ^= 1;
          if (++sb == CF_B_STAGES) sb = 0, pb ^= 1;
        }
        umma_commit(&tmem_full_bar[as]);
        if (++as == 2) as = 0, aphase ^= 1;
      }
    }
  } else if (warp_idx >= 4 && warp_idx < 8) {
    // ===================== epilogue: +bias2, ReLU, bf16, compact raster =====================
    const int quad = warp_idx - 4;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int row = tile * CF_BLOCK_M + quad * 32 + lane;
      bool valid = false;
      size_t orow = 0;
      if (row < p.M) {
        const int b = row / (p.Th * p.FH);
        const int rem = row - b * (p.Th * p.FH);
        const int t = rem / p.FH;
        const int f = rem - t * p.FH;
        valid = (t < p.Tout) && (f < p.Fout);
        orow = ((size_t)b * p.Tout + t) * p.Fout + f;
      }
      mbar_wait(&tmem_full_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * CF_BLOCK_N;
#pragma unroll 1
      for (int c = 0; c < CF_BLOCK_N / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (valid) {
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = *reinterpret_cast<const float4*>(s_b2 + c * 32 + 4 * j);
            pk[2 * j] = pack_bf16x2(fmaxf(__uint_as_float(r[4 * j]) + bv.x, 0.f), fmaxf(__uint_as_float(r[4 * j + 1]) + bv.y, 0.f));
            pk[2 * j + 1] =
                pack_bf16x2(fmaxf(__uint_as_float(r[4 * j + 2]) + bv.z, 0.f), fmaxf(__uint_as_float(r[4 * j + 3]) + bv.w, 0.f));
          }
          uint4* dst = reinterpret_cast<uint4*>(p.out + orow * 256 + c * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) dst[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[as]);
      if (++as == 2) as = 0, aphase ^= 1;
    }
  } else if (warp_idx >= 8) {
    // ===================== A producers: conv1 + ReLU computed per tap from a normalised input patch ===========
    // thread = (4 consecutive rows, 4 channels): per k-block 40 weights (10 LDS.128; the 16 channel-group lanes of a
    // half warp read 256 contiguous bytes) and 36 inputs (registers, reused over the four 64-channel blocks of a tap)
    // feed 144 FMAs in 16 independent chains; 16 warps hide the LDS / FMA latencies.
    const int pt = threadIdx.x - 256;                 // 0..511
    const int cg = pt & 15;                           // 4-channel group inside the 64-wide k-block
    const int r0 = (pt >> 4) * 4;                     // first of this thread's 4 tile rows (same raster line: FH % 4 == 0)
    const int lines_total = p.B * p.Th;
    // raw inputs of one tile -> patch buffer (cp.async, zero fill outside the utterance): line l needs rows 4t'..4t'+6
    auto prefetch_patch = [&](int tile, float* dst) {
      const int line0 = (tile * CF_BLOCK_M) / p.FH;
      for (int idx = pt; idx < CF_PATCH_ELEMS; idx += CF_PRODUCERS) {
        const int l = idx / (7 * CF_PATCH_COLS);
        const int rem = idx - l * (7 * CF_PATCH_COLS);
        const int i = rem / CF_PATCH_COLS;
        const int col = rem - i * CF_PATCH_COLS;
        const int gl = line0 + l;
        const int b = gl / p.Th;
        const int ti = 4 * (gl - b * p.Th) + i;
        const bool ok = tile < p.num_tiles && gl < lines_total && col < p.F && ti < p.T;
        cp_async_f32_zfill(dst + idx, ok ? p.feats + ((size_t)b * p.T + ti) * p.F + col : p.feats, ok);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int stage = 0;
    uint32_t phase = 0;
    int buf = 0;
    prefetch_patch(blockIdx.x, s_patch);
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, buf ^= 1) {
      float* patch = s_patch + buf * CF_PATCH_ELEMS;
      // own copies landed -> GlobalCMVN in place on the elements this thread copied -> visible to all after the barrier
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      for (int idx = pt; idx < CF_PATCH_ELEMS; idx += CF_PRODUCERS) {
        const int col = idx % CF_PATCH_COLS;
        patch[idx] = (patch[idx] - s_mean[col]) * s_istd[col];
      }
      named_bar_sync(1, CF_PRODUCERS);  // also: everyone is done with the other buffer (previous tile)
      prefetch_patch(tile + gridDim.x, s_patch + (buf ^ 1) * CF_PATCH_ELEMS);
      const int line0 = (tile * CF_BLOCK_M) / p.FH;
      const int rrow = tile * CF_BLOCK_M + r0;        // raster row of the thread's first row
      const int l = rrow / p.FH - line0;
      const int fq = rrow - (line0 + l) * p.FH;       // f' of the first row (multiple of 4)
      const float* pl = patch + l * (7 * CF_PATCH_COLS) + 4 * fq;
      for (int tap = 0; tap < 9; ++tap) {
        const int kh = tap / 3, kw = tap - kh * 3;
        // 3x3 input windows of conv1 output positions (2t'+kh, 2(f'+rr)+kw), rr = 0..3
        float xv[4][9];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr)
#pragma unroll
          for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) xv[rr][i * 3 + j] = pl[(2 * kh + i) * CF_PATCH_COLS + 4 * rr + 2 * kw + j];
        for (int kc = 0; kc < 4; ++kc) {
          const float* wk = s_w1 + kc * 64 + cg * 4;
          float acc[4][4];
          {
            const float4 b0 = *reinterpret_cast<const float4*>(wk + 9 * 256);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) acc[rr][0] = b0.x, acc[rr][1] = b0.y, acc[rr][2] = b0.z, acc[rr][3] = b0.w;
          }
#pragma unroll
          for (int k = 0; k < 9; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(wk + k * 256);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
              acc[rr][0] = fmaf(w.x, xv[rr][k], acc[rr][0]);
              acc[rr][1] = fmaf(w.y, xv[rr][k], acc[rr][1]);
              acc[rr][2] = fmaf(w.z, xv[rr][k], acc[rr][2]);
              acc[rr][3] = fmaf(w.w, xv[rr][k], acc[rr][3]);
            }
          }
          mbar_wait(&a_empty[stage], phase ^ 1);
          uint8_t* abase = smem_a + stage * CF_A_BYTES + (cg & 1) * 8;
#pragma unroll
          for (int rr = 0; rr < 4; ++rr) {
            const int r = r0 + rr;
            const uint32_t lo = pack_bf16x2(fmaxf(acc[rr][0], 0.f), fmaxf(acc[rr][1], 0.f));
            const uint32_t hi = pack_bf16x2(fmaxf(acc[rr][2], 0.f), fmaxf(acc[rr][3], 0.f));
            *reinterpret_cast<uint2*>(abase + r * 128 + (((cg >> 1) ^ (r & 7)) << 4)) = make_uint2(lo, hi);
          }
          fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&a_full[stage]);
          if (++stage == CF_A_STAGES) stage = 0, phase ^= 1;
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

cudaError_t launch_conv_front(const CUtensorMap& tmap_w2, const float* feats, const float* mean, const float* istd,
                              const float* w1, const float* b1, const float* b2, __nv_bfloat16* out, int B, int T, int F,
                              int Th, int FH, int Tout, int Fout, int num_sms, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv_front_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CF_SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  ConvFrontParams p;
  p.feats = feats, p.mean = mean, p.istd = istd, p.w1 = w1, p.b1 = b1, p.b2 = b2, p.out = out;
  p.B = B, p.T = T, p.F = F, p.Th = Th, p.FH = FH, p.Tout = Tout, p.Fout = Fout;
  p.M = B * Th * FH;
  p.num_tiles = (p.M + CF_BLOCK_M - 1) / CF_BLOCK_M;
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  cudaError_t le = launch_pdl(conv_front_kernel, dim3(grid), dim3(CF_THREADS), (size_t)CF_SMEM_TOTAL, st, tmap_w2, p);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace ppasr
