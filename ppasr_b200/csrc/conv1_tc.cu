// GlobalCMVN + Conv2d(1 -> 256, 3x3, stride 2) + ReLU on the tensor cores (sm_100a), writing the four stride-phase
// images the conv2 GEMM reads (DESIGN.md §3 `phase`).
//
// Reference: ppasr/model_utils/utils/cmvn.py:29-32, ppasr/model_utils/conformer/subsampling.py:84-85,110.
//
// Why a GEMM for a K = 9 convolution: the CUDA-core kernel (kernels.cu::conv1_subsample_kernel) spends ~12 issue slots
// per output element (91 M warp instructions per C2 step, 141 us, 55 % issue utilisation -- ncu r2_conv) while the
// output it has to write (326 MB of bf16) needs ~50 us of HBM time. Here one output row (a conv1 position) is a GEMM
// row, the 256 channels are N, and the 3x3 window is K; the CUDA cores only build the 128-byte im2col row per position
// and convert the accumulators, ~1.2 issue slots per element.
//
// fp32 accuracy from tf32 tensor cores: x = x_hi + x_lo and w = w_hi + w_lo with x_hi, w_hi rounded to tf32
// (10 mantissa bits; x_lo = x - x_hi is exact in fp32 and keeps 10 more bits after its own rounding). The K dimension
// carries  [x_hi | x_lo | x_hi | 1 | 1 | 0 0 0] . [w_hi | w_hi | w_lo | b_hi | b_lo | 0 0 0]  (9 + 9 + 9 + 2 + 3 = 32),
// i.e. every product except x_lo * w_lo (2^-22 relative) and the bias, accumulated in fp32 by the tensor core. The
// result agrees with the fp32 FMA chain of the CUDA-core kernel to ~1e-6 relative, far inside one bf16 ulp.
//
// Tile = 128 consecutive raster rows r = (b*Th + th)*FH + f2 of ONE phase image (ph = 2*(t1&1) + (f1&1),
// t1 = 2 th + pt, f1 = 2 f2 + pf) x 256 channels = 64 KB contiguous in global memory.
//   warps 0-3 : A producers (thread = row): 9 inputs -> normalise -> hi/lo split -> one swizzled 128-byte K-major row
//   warps 4-7 : epilogue: TMEM -> ReLU + bf16 (cvt.rn.relu.bf16x2.f32) -> 128B-swizzled staging -> TMA tensor store
//   warp  8   : MMA issuer (4 x tcgen05.mma kind::tf32 128x256x8 per tile), TMEM alloc (2 x 256 columns)
// Everything is double-buffered (A tile, accumulator, output staging) so the kernel runs at the rate the 64 KB tile
// stores drain to HBM. Rows outside the image (t1 >= T1, f1 >= F1, r >= Mr) are zero K-rows -> relu(0) = 0, as before.
#include "kernels.h"
#include "launch.h"
#include "ptx.cuh"

namespace ppasr {

void count_launch();

constexpr int C1_BLOCK_M = 128;
constexpr int C1_N = 256;
constexpr int C1_THREADS = 288;
constexpr int C1_B_BYTES = C1_N * 128;          // 32 KB: [256 channels][32 tf32]
constexpr int C1_A_BYTES = C1_BLOCK_M * 128;    // 16 KB: [128 rows][32 tf32]
constexpr int C1_OUT_BYTES = C1_BLOCK_M * C1_N * 2;  // 64 KB: 4 boxes of [128 rows][64 bf16]
constexpr int C1_SMEM_A = C1_B_BYTES;
constexpr int C1_SMEM_OUT = C1_SMEM_A + 2 * C1_A_BYTES;
constexpr int C1_SMEM_BAR = C1_SMEM_OUT + 2 * C1_OUT_BYTES;
constexpr int C1_SMEM_CMVN = C1_SMEM_BAR + 128;   // mean[256], istd[256]
constexpr int C1_SMEM_TOTAL = C1_SMEM_CMVN + 2048 + 1024;

struct Conv1TcParams {
  const float* feats;  // [B, T, F]
  const float* mean;   // [F]
  const float* istd;   // [F]
  const float* w;      // [256, 9]
  const float* bias;   // [256]
  int B, T, F, T1, F1, Th, FH;
  int Mr;              // B * Th * FH raster rows per phase image
  int mtiles;          // ceil(Mr / 128)
  int num_tiles;       // 4 * mtiles
};

DEVINL uint32_t tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// {lo half: a, hi half: b}, each max(x, 0) rounded to nearest-even bf16
DEVINL uint32_t relu_pack_bf16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
DEVINL void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::tf32, A and B K-major tf32, fp32 accumulate (cute/arch/mma_sm100_desc.hpp: F16F32Format::TF32 = 2)
__host__ __device__ constexpr uint32_t umma_idesc_tf32(uint32_t M, uint32_t N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}
DEVINL void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
DEVINL void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
DEVINL void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
DEVINL void bulk_wait_group_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(C1_THREADS, 1)
conv1_tc_kernel(const __grid_constant__ CUtensorMap tmap_phase, const Conv1TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* smem_b = smem;
  uint8_t* smem_a = smem + C1_SMEM_A;
  uint8_t* smem_out = smem + C1_SMEM_OUT;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + C1_SMEM_BAR);
  uint64_t* a_empty = a_full + 2;
  uint64_t* acc_full = a_empty + 2;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* s_mean = reinterpret_cast<float*>(smem + C1_SMEM_CMVN);
  float* s_istd = s_mean + 256;

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp_idx == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_phase);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 4);     // one arrival per producer warp
      mbar_init(&a_empty[i], 1);    // tcgen05.commit
      mbar_init(&acc_full[i], 1);   // tcgen05.commit
      mbar_init(&acc_empty[i], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp_idx == 8) tmem_alloc<512>(tmem_ptr_smem);
  // B operand (constants, may be read before the grid dependency resolves): [w_hi | w_hi | w_lo | b_hi b_lo 0 0 0]
  for (int idx = threadIdx.x; idx < C1_N * 8; idx += C1_THREADS) {
    const int c = idx >> 3, j = idx & 7;
    uint32_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = 4 * j + e;
      uint32_t o = 0;
      if (k < 27) {
        const float w = __ldg(p.w + c * 9 + (k % 9));
        const uint32_t hi = tf32_rna(w);
        o = k < 18 ? hi : tf32_rna(w - __uint_as_float(hi));
      } else if (k < 29) {
        const float b = __ldg(p.bias + c);
        const uint32_t hi = tf32_rna(b);
        o = k == 27 ? hi : tf32_rna(b - __uint_as_float(hi));
      }
      v[e] = o;
    }
    *reinterpret_cast<uint4*>(smem_b + c * 128 + ((j ^ (c & 7)) << 4)) = make_uint4(v[0], v[1], v[2], v[3]);
  }
  for (int i = threadIdx.x; i < p.F; i += C1_THREADS) s_mean[i] = __ldg(p.mean + i), s_istd[i] = __ldg(p.istd + i);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  pdl_wait();
  pdl_launch_dependents();

  if (warp_idx < 4) {
    // ===================== A producers: one im2col row per thread =====================
    // The raw inputs of the next two tiles are already in flight (registers) while the current row is converted: the loads
    // miss to L2 / HBM underneath 3+ TB/s of phase-image writes, and one tile period (~1.5 us) does not cover that latency.
    const int row = threadIdx.x;  // 0..127
    const int sw = row & 7;
    float x0[9], x1[9], x2[9];
    int c0 = -1, c1 = -1, c2 = -1;  // first input column of the window (2 * f1), -1 = row outside the image
    auto fetch = [&](int tile, float (&x)[9], int& col) {
      col = -1;
#pragma unroll
      for (int i = 0; i < 9; ++i) x[i] = 0.f;
      if (tile >= p.num_tiles) return;
      const int ph = tile & 3, mt = tile >> 2;
      const int r = mt * C1_BLOCK_M + row;
      const int bl = r / p.FH;
      const int f2 = r - bl * p.FH;
      const int b = bl / p.Th;
      const int th = bl - b * p.Th;
      const int t1 = 2 * th + (ph >> 1), f1 = 2 * f2 + (ph & 1);
      if (r < p.Mr && t1 < p.T1 && f1 < p.F1) {
        col = 2 * f1;
        const float* src = p.feats + ((size_t)b * p.T + 2 * t1) * p.F + col;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) x[kh * 3 + kw] = __ldg(src + kh * p.F + kw);
      }
    };
    fetch(blockIdx.x, x0, c0);
    fetch(blockIdx.x + gridDim.x, x1, c1);
    int s = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      fetch(tile + 2 * gridDim.x, x2, c2);
      const bool valid = c0 >= 0;
      uint32_t hi[9], lo[9];
#pragma unroll
      for (int i = 0; i < 9; ++i) hi[i] = 0u, lo[i] = 0u;
      if (valid) {
        float m[3], sd[3];
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) m[kw] = s_mean[c0 + kw], sd[kw] = s_istd[c0 + kw];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const float x = (x0[i] - m[i % 3]) * sd[i % 3];
          const uint32_t h = tf32_rna(x);
          hi[i] = h;
          lo[i] = tf32_rna(x - __uint_as_float(h));
        }
      }
      const uint32_t one = valid ? 0x3F800000u : 0u;
      mbar_wait(&a_empty[s], phase ^ 1);
      uint8_t* arow = smem_a + s * C1_A_BYTES + row * 128;
      // k: 0-8 x_hi, 9-17 x_lo, 18-26 x_hi, 27-28 one, 29-31 zero
      *reinterpret_cast<uint4*>(arow + ((0 ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(arow + ((1 ^ sw) << 4)) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      *reinterpret_cast<uint4*>(arow + ((2 ^ sw) << 4)) = make_uint4(hi[8], lo[0], lo[1], lo[2]);
      *reinterpret_cast<uint4*>(arow + ((3 ^ sw) << 4)) = make_uint4(lo[3], lo[4], lo[5], lo[6]);
      *reinterpret_cast<uint4*>(arow + ((4 ^ sw) << 4)) = make_uint4(lo[7], lo[8], hi[0], hi[1]);
      *reinterpret_cast<uint4*>(arow + ((5 ^ sw) << 4)) = make_uint4(hi[2], hi[3], hi[4], hi[5]);
      *reinterpret_cast<uint4*>(arow + ((6 ^ sw) << 4)) = make_uint4(hi[6], hi[7], hi[8], one);
      *reinterpret_cast<uint4*>(arow + ((7 ^ sw) << 4)) = make_uint4(one, 0u, 0u, 0u);
      fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_full[s]);
      if (++s == 2) s = 0, phase ^= 1;
#pragma unroll
      for (int i = 0; i < 9; ++i) x0[i] = x1[i], x1[i] = x2[i];
      c0 = c1, c1 = c2;
    }
  } else if (warp_idx < 8) {
    // ===================== epilogue: ReLU + bf16 -> swizzled staging -> TMA store =====================
    const int quad = warp_idx - 4;
    const int row = quad * 32 + lane;
    const int sw = row & 7;
    const bool issuer = threadIdx.x == 128;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int ph = tile & 3, mt = tile >> 2;
      uint8_t* obuf = smem_out + as * C1_OUT_BYTES;
      if (issuer) bulk_wait_group_read<1>();  // the store issued two tiles ago has finished reading this buffer
      named_bar_sync(1, 128);
      mbar_wait(&acc_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + as * C1_N;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(taddr + c * 32, v);
        tmem_ld_wait();
        uint8_t* orow = obuf + (c >> 1) * (C1_BLOCK_M * 128) + row * 128;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 o;
          o.x = relu_pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
          o.y = relu_pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
          o.z = relu_pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
          o.w = relu_pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
          *reinterpret_cast<uint4*>(orow + ((((c & 1) * 4 + j) ^ sw) << 4)) = o;
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();  // staging writes -> visible to the TMA store (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[as]);
      named_bar_sync(1, 128);
      if (issuer) {
#pragma unroll
        for (int g = 0; g < 4; ++g) tma_store_3d(&tmap_phase, obuf + g * (C1_BLOCK_M * 128), g * 64, mt * C1_BLOCK_M, ph);
        bulk_commit_group();
      }
      if (++as == 2) as = 0, aphase ^= 1;
    }
    if (issuer) bulk_wait_group_all();
  } else {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      constexpr uint32_t idesc = umma_idesc_tf32(C1_BLOCK_M, C1_N);
      const uint32_t b_addr = smem_u32(smem_b);
      int s = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&acc_empty[s], phase ^ 1);
        mbar_wait(&a_full[s], phase);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem_a + s * C1_A_BYTES);
        const uint32_t tmem_d = tmem_base + s * C1_N;
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_tf32(tmem_d, umma_desc_k_sw128(a_addr + k * 32), umma_desc_k_sw128(b_addr + k * 32), idesc, k != 0 ? 1u : 0u);
        umma_commit(&a_empty[s]);
        umma_commit(&acc_full[s]);
        if (++s == 2) s = 0, phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp_idx == 8) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

cudaError_t launch_conv1_tc(const CUtensorMap& tmap_phase, const float* feats, const float* mean, const float* istd,
                            const float* w, const float* bias, int B, int T, int F, int T1, int F1, int Th, int FH,
                            int num_sms, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (F > 256) return cudaErrorInvalidValue;  // s_mean / s_istd staging
  Conv1TcParams p;
  p.feats = feats, p.mean = mean, p.istd = istd, p.w = w, p.bias = bias;
  p.B = B, p.T = T, p.F = F, p.T1 = T1, p.F1 = F1, p.Th = Th, p.FH = FH;
  p.Mr = B * Th * FH;
  p.mtiles = (p.Mr + C1_BLOCK_M - 1) / C1_BLOCK_M;
  p.num_tiles = 4 * p.mtiles;
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  cudaError_t le = launch_pdl(conv1_tc_kernel, dim3(grid), dim3(C1_THREADS), (size_t)C1_SMEM_TOTAL, st, tmap_phase, p);
  count_launch();
  return le != cudaSuccess ? le : cudaGetLastError();
}

}  // namespace ppasr
