// CTC prefix beam search on the GPU (optionally with a character n-gram scorer).
//
// Replaces the C++ library behind ppasr/decoders/swig_wrapper.py:35-121 / beam_search_decoder.py:45-96
// (`paddlespeech_ctcdecoders`: ctc_beam_search_decoder.cpp, decoder_utils.cpp, path_trie.cpp -- not vendored
// in the reference, restated in oracle/decoders_oracle.py).
//
// Two kernels:
//  (1) ctc_prune_cta_kernel (V >= 512; ctc_prune_kernel = the same algorithm with one WARP per frame, small vocabularies)
//      -- the HBM-bound scan of the posterior [B,T,V]: one CTA of four warps per frame. Pass 1 streams the row
//      with 16-byte loads keeping each lane's two largest values; the 64 kept values are sorted in registers (warp
//      bitonic network) and a double-precision prefix sum over them gives a threshold tau that is a lower bound of
//      everything that can be selected (decoder_utils.cpp get_pruned_log_probs: descending probability, ties lower
//      index first, smallest head whose cumulative probability reaches cutoff_prob, capped at cutoff_top_n). Pass 2
//      re-reads the row (L2 hits) and compacts the elements >= tau; the <= 64 survivors are sorted once by
//      (probability desc, index asc) in registers and cut with the same prefix sum. Output per frame: n, ids[n],
//      log(p + FLT_MIN)[n]. Every frame is independent, so the scan runs at full grid width; the sequential part
//      never touches the [B,T,V] tensor.
//  (2) ctc_prefix_beam_kernel<NT> -- one CTA of NT threads per utterance (128 for beams <= 32, 256 / 512 above)
//      walks the frames over the compact candidate lists. The prefix trie of the reference is replaced by an
//      equivalent flat form: a beam entry carries a 64-bit hash of its id string (its identity) and an id into a
//      (parent id, char) table used only to read the string back; an extension (prefix i, char c) merges into beam
//      entry j iff hash[j] == H(hash[i], c) (looked up in a shared-memory hash table of the beam), otherwise it is a
//      fresh prefix (a revived trie node is reset to -inf in path_trie.cpp, i.e. indistinguishable from a fresh one).
//      Top-`beam` selection per frame: every live candidate becomes one 64-bit key (score desc | last char asc | slot
//      asc, all distinct); beams <= 32 keep the best keys by a tournament of register sorts (each warp sorts 64-key
//      chunks with a shuffle network and keeps each chunk's best 2^ceil(log2 beam)), wider beams by one shared-memory
//      bitonic sort -- prefix_compare's order, beam sorted best first as in the reference. State (beam + id table) lives in global memory so the same kernel serves the streaming decode_chunk API.
#include <float.h>

#include "kernels.h"
#include "ptx.cuh"

#include <cstdlib>
#include <mutex>

namespace ppasr {

void count_launch();

constexpr int BEAM_MAXB = BEAM_MAX_BEAM;   // max beam size (512)
constexpr int BEAM_MAXC = BEAM_MAX_TOPN;   // max cutoff_top_n (64)

// ------------------------------------------------------------------------------------------------
// (1) pruning scan
// ------------------------------------------------------------------------------------------------
// Selection order everywhere: probability descending, index ascending on ties.
DEVINL bool prune_before(float pa, int ia, float pb, int ib) { return pa > pb || (pa == pb && ia < ib); }

constexpr int PRUNE_CAND_MAX = 256;  // compacted candidates per row; more (mass ties) falls back to full-row rounds
constexpr int PRUNE_WARPS = 4;

// probability (>= 0, or -inf = empty) and index -> one key whose ASCENDING order is (probability desc, index asc)
DEVINL unsigned long long prune_key(float p, int idx) {
  const unsigned u = (p == -INFINITY) ? 0u : (__float_as_uint(p) | 0x80000000u);  // monotone for p >= 0; empty sorts last
  return ((unsigned long long)(~u) << 32) | (unsigned)idx;
}
DEVINL float prune_key_prob(unsigned long long k) {
  const unsigned u = ~(unsigned)(k >> 32);
  return u == 0u ? -INFINITY : __uint_as_float(u & 0x7fffffffu);
}

// 64 values, element e = r * 32 + lane in register a_r: bitonic network, result DESCENDING in e
DEVINL void warp_sort64_desc(float& a0, float& a1, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {
        const float hi = fmaxf(a0, a1), lo = fminf(a0, a1);
        a0 = hi, a1 = lo;
      } else {
        const float p0 = __shfl_xor_sync(0xffffffffu, a0, j);
        const float p1 = __shfl_xor_sync(0xffffffffu, a1, j);
        const bool lower = (lane & j) == 0;
        const bool desc0 = (lane & k) == 0;          // e = lane      : (e & k) == 0 (k <= 32), always for k = 64
        const bool desc1 = ((32 + lane) & k) == 0;   // e = 32 + lane
        a0 = (desc0 == lower) ? fmaxf(a0, p0) : fminf(a0, p0);
        a1 = (desc1 == lower) ? fmaxf(a1, p1) : fminf(a1, p1);
      }
    }
  }
}
// same network on 64-bit keys, result ASCENDING in e
DEVINL void warp_sort64_keys(unsigned long long& a0, unsigned long long& a1, int lane) {
#pragma unroll
  for (int k = 2; k <= 64; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j == 32) {
        const unsigned long long lo = a0 < a1 ? a0 : a1, hi = a0 < a1 ? a1 : a0;
        a0 = lo, a1 = hi;
      } else {
        const unsigned long long p0 = __shfl_xor_sync(0xffffffffu, a0, j);
        const unsigned long long p1 = __shfl_xor_sync(0xffffffffu, a1, j);
        const bool lower = (lane & j) == 0;
        const bool asc0 = (lane & k) == 0;
        const bool asc1 = ((32 + lane) & k) == 0;
        a0 = (asc0 == lower) ? (a0 < p0 ? a0 : p0) : (a0 < p0 ? p0 : a0);
        a1 = (asc1 == lower) ? (a1 < p1 ? a1 : p1) : (a1 < p1 ? p1 : a1);
      }
    }
  }
}
// 32 keys, one per lane, ascending in lane
DEVINL void warp_sort32_keys(unsigned long long& a, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const unsigned long long p = __shfl_xor_sync(0xffffffffu, a, j);
      const bool lower = (lane & j) == 0;
      const bool asc = (lane & k) == 0 || k == 32;
      a = (asc == lower) ? (a < p ? a : p) : (a < p ? p : a);
    }
  }
}
// inclusive prefix sum in double over element order e = r * 32 + lane (values in x0, x1)
DEVINL void warp_scan64(double& x0, double& x1, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double y0 = __shfl_up_sync(0xffffffffu, x0, o);
    const double y1 = __shfl_up_sync(0xffffffffu, x1, o);
    if (lane >= o) x0 += y0, x1 += y1;
  }
  x1 += __shfl_sync(0xffffffffu, x0, 31);
}

__global__ void __launch_bounds__(PRUNE_WARPS * 32, 12) ctc_prune_kernel(const float* __restrict__ probs, int V, int rows,
                                                                      float cutoff_prob, int top_n, int* __restrict__ cnt,
                                                                      int* __restrict__ cid, float* __restrict__ clp) {
  __shared__ float s_cp[PRUNE_WARPS][PRUNE_CAND_MAX];
  __shared__ int s_ci[PRUNE_WARPS][PRUNE_CAND_MAX];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * PRUNE_WARPS + warp;
  if (row >= rows) return;
  float* cp = s_cp[warp];
  int* ci = s_ci[warp];
  const float* src = probs + (size_t)row * V;
  const bool prune = (cutoff_prob < 1.0f) || (top_n < V);
  const int limit = min(prune ? min(top_n, V) : V, BEAM_MAXC);
  const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
  const int head = mis ? min(4 - mis, V) : 0;
  const float4* v4 = reinterpret_cast<const float4*>(src + head);
  const int n4 = (V - head) >> 2;
  const int tail0 = head + 4 * n4;
  int* oid = cid + (size_t)row * BEAM_MAXC;
  float* olp = clp + (size_t)row * BEAM_MAXC;
  // ---- pass 1: stream the row once keeping each lane's two largest values (2 FMNMX per element) ----
  float t1 = -INFINITY, t2 = -INFINITY;
  auto upd = [&](float x) {
    t2 = fmaxf(t2, fminf(t1, x));
    t1 = fmaxf(t1, x);
  };
  auto upd4 = [&](const float4& a) { upd(a.x), upd(a.y), upd(a.z), upd(a.w); };
  if (lane < head) upd(src[lane]);
  if (tail0 + lane < V) upd(src[tail0 + lane]);
  int i = lane;
  for (; i + 224 < n4; i += 256) {
    const float4 a0 = __ldg(v4 + i), a1 = __ldg(v4 + i + 32), a2 = __ldg(v4 + i + 64), a3 = __ldg(v4 + i + 96);
    const float4 a4 = __ldg(v4 + i + 128), a5 = __ldg(v4 + i + 160), a6 = __ldg(v4 + i + 192), a7 = __ldg(v4 + i + 224);
    upd4(a0), upd4(a1), upd4(a2), upd4(a3), upd4(a4), upd4(a5), upd4(a6), upd4(a7);
  }
  for (; i < n4; i += 32) upd4(__ldg(v4 + i));
  // ---- threshold: the 64 per-lane maxima in descending order. They are row elements, so the k-th of them is a lower
  //      bound of the k-th largest of the row and their running sum a lower bound of the true cumulative mass: at the
  //      first k where it reaches cutoff_prob (or k = limit) everything that can be selected is >= tau = that value.
  float tau = -INFINITY;
  if (prune) {
    float a0 = t1, a1 = t2;
    warp_sort64_desc(a0, a1, lane);
    double c0 = a0 > -INFINITY ? (double)a0 : 0.0, c1 = a1 > -INFINITY ? (double)a1 : 0.0;
    warp_scan64(c0, c1, lane);
    const double thr = (double)cutoff_prob * 1.0001;  // margin: lower-bound sum vs the exact sum used below
    const bool stop0 = (cutoff_prob < 1.0f && c0 >= thr) || (lane + 1 >= limit) || a0 == -INFINITY;
    const bool stop1 = (cutoff_prob < 1.0f && c1 >= thr) || (lane + 33 >= limit) || a1 == -INFINITY;
    const unsigned b0 = __ballot_sync(0xffffffffu, stop0);
    const unsigned b1 = __ballot_sync(0xffffffffu, stop1);
    if (b0) {
      tau = __shfl_sync(0xffffffffu, a0, __ffs(b0) - 1);
    } else if (b1) {
      tau = __shfl_sync(0xffffffffu, a1, __ffs(b1) - 1);
    } else {
      tau = __shfl_sync(0xffffffffu, a1, 31);
    }
    // an exhausted list (-inf reached before the stop) means every finite element is a candidate
  }
  // ---- pass 2: compact elements >= tau. Only the lanes that hold a candidate do any work: the slot comes from a
  //      shared-memory counter (order irrelevant: the survivors are sorted by (probability, index) below) ----
  __shared__ int s_cnt[PRUNE_WARPS];
  if (lane == 0) s_cnt[warp] = 0;
  __syncwarp();
  auto emit = [&](float ev, int idx) {
    if (ev >= tau && ev > -INFINITY) {
      const int pos = atomicAdd(&s_cnt[warp], 1);
      if (pos < PRUNE_CAND_MAX) {
        cp[pos] = ev;
        ci[pos] = idx;
      }
    }
  };
  if (lane < head) emit(src[lane], lane);
  if (tail0 + lane < V) emit(src[tail0 + lane], tail0 + lane);
  for (int q0 = 0; q0 < n4; q0 += 128) {
    float4 x[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int q = q0 + u * 32 + lane;
      x[u] = (q < n4) ? __ldg(v4 + q) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float mx = fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w));
      if (mx >= tau && mx > -INFINITY) {
        const int base = head + 4 * (q0 + u * 32 + lane);
        emit(x[u].x, base);
        emit(x[u].y, base + 1);
        emit(x[u].z, base + 2);
        emit(x[u].w, base + 3);
      }
    }
  }
  __syncwarp();
  const int c = s_cnt[warp];
  __syncwarp();
  if (c <= 64) {
    // ---- the common case: sort the survivors once, cut with an exact (double) running sum ----
    unsigned long long k0 = lane < c ? prune_key(cp[lane], ci[lane]) : ~0ull;
    unsigned long long k1 = 32 + lane < c ? prune_key(cp[32 + lane], ci[32 + lane]) : ~0ull;
    if (c <= 32) {
      warp_sort32_keys(k0, lane);
    } else {
      warp_sort64_keys(k0, k1, lane);
    }
    const float p0 = k0 != ~0ull ? prune_key_prob(k0) : -INFINITY;
    const float p1 = k1 != ~0ull ? prune_key_prob(k1) : -INFINITY;
    double c0 = p0 > -INFINITY ? (double)p0 : 0.0, c1 = p1 > -INFINITY ? (double)p1 : 0.0;
    warp_scan64(c0, c1, lane);
    int n = min(c, limit);
    if (cutoff_prob < 1.0f) {
      const unsigned b0 = __ballot_sync(0xffffffffu, lane < c && c0 >= (double)cutoff_prob);
      const unsigned b1 = __ballot_sync(0xffffffffu, 32 + lane < c && c1 >= (double)cutoff_prob);
      if (b0) n = min(n, __ffs(b0));
      else if (b1) n = min(n, 32 + __ffs(b1));
    }
    if (lane < n) {
      oid[lane] = (int)(unsigned)k0;
      olp[lane] = logf(p0 + FLT_MIN);
    }
    if (32 + lane < n) {
      oid[32 + lane] = (int)(unsigned)k1;
      olp[32 + lane] = logf(p1 + FLT_MIN);
    }
    if (lane == 0) cnt[row] = n;
    return;
  }
  // ---- many ties at tau (e.g. uniform rows): arg-max rounds with a Kahan-compensated running sum ----
  float last_p = INFINITY;
  int last_i = -1;
  float cum = 0.f, cum_c = 0.f;
  int n = 0;
  const bool compact = c <= PRUNE_CAND_MAX;
  while (n < limit) {
    float bm = -INFINITY;
    int bi = 0x7fffffff;
    if (compact) {
      for (int j = lane; j < c; j += 32) {
        const float pv = cp[j];
        const int iv = ci[j];
        if (prune_before(last_p, last_i, pv, iv) && prune_before(pv, iv, bm, bi)) {
          bm = pv;
          bi = iv;
        }
      }
    } else {
      for (int j = lane; j < V; j += 32) {
        const float pv = __ldg(src + j);
        if (prune_before(last_p, last_i, pv, j) && prune_before(pv, j, bm, bi)) {
          bm = pv;
          bi = j;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, bm, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (prune_before(om, oi, bm, bi)) {
        bm = om;
        bi = oi;
      }
    }
    if (bi == 0x7fffffff) break;
    if (lane == 0) {
      oid[n] = bi;
      olp[n] = logf(bm + FLT_MIN);
    }
    last_p = bm;
    last_i = bi;
    ++n;
    const float yk = bm - cum_c;
    const float tk = cum + yk;
    cum_c = (tk - cum) - yk;
    cum = tk;
    if (cutoff_prob < 1.0f && cum >= cutoff_prob) break;
  }
  if (lane == 0) cnt[row] = n;
}

// ------------------------------------------------------------------------------------------------
// CTA-per-row variant (default for V >= 512): the four warps of a CTA share ONE row, a quarter each. Same algorithm, but
//  * the rows in flight are the resident CTAs, not the resident warps: 12 CTAs x 148 SMs x 17 KB = 30 MB instead of
//    120 MB, so the second pass really hits L2 (the warp-per-row kernel re-read 80 % of the posterior from DRAM: ncu
//    241 MB for 134 MB) while 48 warps per SM keep the loads in flight;
//  * every thread's share of the row is requested in batches of 5 float4 (two DRAM round trips per 4233-wide row);
//  * threshold: the 128 per-thread maxima (row elements, hence lower bounds as before) -- sorted per warp in registers,
//    merged pairwise with bitonic merges (2 x 6 + 6 shuffle stages), prefix-summed in double precision by warp 0.
// ------------------------------------------------------------------------------------------------
DEVINL void warp_sort32_desc_f(float& a, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const float p = __shfl_xor_sync(0xffffffffu, a, j);
      const bool lower = (lane & j) == 0;
      const bool desc = (lane & k) == 0 || k == 32;
      a = (desc == lower) ? fmaxf(a, p) : fminf(a, p);
    }
  }
}
// (a0, a1) = a bitonic sequence of 64 (element e = r * 32 + lane) -> sorted DESCENDING
DEVINL void warp_bitonic_merge64_desc(float& a0, float& a1, int lane) {
  {
    const float hi = fmaxf(a0, a1), lo = fminf(a0, a1);
    a0 = hi, a1 = lo;
  }
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const float p0 = __shfl_xor_sync(0xffffffffu, a0, j);
    const float p1 = __shfl_xor_sync(0xffffffffu, a1, j);
    const bool lower = (lane & j) == 0;
    a0 = lower ? fmaxf(a0, p0) : fminf(a0, p0);
    a1 = lower ? fmaxf(a1, p1) : fminf(a1, p1);
  }
}

__global__ void __launch_bounds__(128, 12) ctc_prune_cta_kernel(const float* __restrict__ probs, int V, int rows,
                                                                 float cutoff_prob, int top_n, int* __restrict__ cnt,
                                                                 int* __restrict__ cid, float* __restrict__ clp) {
  __shared__ float s_top[128];        // per-thread maxima, sorted descending per warp
  __shared__ float s_m[2][64];        // pairwise merges
  __shared__ float s_cp[PRUNE_CAND_MAX];
  __shared__ int s_ci[PRUNE_CAND_MAX];
  __shared__ int s_c;
  __shared__ float s_tau;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = blockIdx.x;
  const float* src = probs + (size_t)row * V;
  const bool prune = (cutoff_prob < 1.0f) || (top_n < V);
  const int limit = min(prune ? min(top_n, V) : V, BEAM_MAXC);
  const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3);
  const int head = mis ? min(4 - mis, V) : 0;
  const float4* v4 = reinterpret_cast<const float4*>(src + head);
  const int n4 = (V - head) >> 2;
  const int tail0 = head + 4 * n4;
  int* oid = cid + (size_t)row * BEAM_MAXC;
  float* olp = clp + (size_t)row * BEAM_MAXC;
  if (tid == 0) s_c = 0;
  // ---- pass 1: per-thread maximum; the whole row is requested at once (<= 5 + 4 float4 per thread for V <= 4608) ----
  float m = -INFINITY;
  if (tid < head) m = fmaxf(m, src[tid]);
  if (tail0 + tid < V) m = fmaxf(m, src[tail0 + tid]);
  for (int q0 = 0; q0 < n4; q0 += 640) {
    float4 x[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int q = q0 + u * 128 + tid;
      x[u] = (q < n4) ? __ldg(v4 + q) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) m = fmaxf(m, fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w)));
  }
  // ---- threshold from the 128 maxima ----
  float tau = -INFINITY;
  if (prune) {
    warp_sort32_desc_f(m, lane);
    s_top[tid] = m;
    __syncthreads();
    if ((warp & 1) == 0) {  // warps 0 and 2: merge their run with the next warp's (reversed) -> sorted 64
      float a0 = s_top[warp * 32 + lane], a1 = s_top[(warp + 1) * 32 + 31 - lane];
      warp_bitonic_merge64_desc(a0, a1, lane);
      s_m[warp >> 1][lane] = a0;
      s_m[warp >> 1][32 + lane] = a1;
    }
    __syncthreads();
    if (warp == 0) {
      // top 64 of the union of two sorted runs of 64: max(X[e], Y[63 - e]) is bitonic
      float a0 = fmaxf(s_m[0][lane], s_m[1][63 - lane]);
      float a1 = fmaxf(s_m[0][32 + lane], s_m[1][31 - lane]);
      warp_bitonic_merge64_desc(a0, a1, lane);
      double c0 = a0 > -INFINITY ? (double)a0 : 0.0, c1 = a1 > -INFINITY ? (double)a1 : 0.0;
      warp_scan64(c0, c1, lane);
      const double thr = (double)cutoff_prob * 1.0001;  // margin: lower-bound sum vs the exact sum used below
      const bool stop0 = (cutoff_prob < 1.0f && c0 >= thr) || (lane + 1 >= limit) || a0 == -INFINITY;
      const bool stop1 = (cutoff_prob < 1.0f && c1 >= thr) || (lane + 33 >= limit) || a1 == -INFINITY;
      const unsigned b0 = __ballot_sync(0xffffffffu, stop0);
      const unsigned b1 = __ballot_sync(0xffffffffu, stop1);
      float t;
      if (b0) t = __shfl_sync(0xffffffffu, a0, __ffs(b0) - 1);
      else if (b1) t = __shfl_sync(0xffffffffu, a1, __ffs(b1) - 1);
      else t = __shfl_sync(0xffffffffu, a1, 31);
      if (lane == 0) s_tau = t;
    }
    __syncthreads();
    tau = s_tau;
  } else {
    __syncthreads();
  }
  // ---- pass 2 (L2): compact the elements >= tau ----
  auto emit = [&](float ev, int idx) {
    if (ev >= tau && ev > -INFINITY) {
      const int pos = atomicAdd(&s_c, 1);
      if (pos < PRUNE_CAND_MAX) {
        s_cp[pos] = ev;
        s_ci[pos] = idx;
      }
    }
  };
  if (tid < head) emit(src[tid], tid);
  if (tail0 + tid < V) emit(src[tail0 + tid], tail0 + tid);
  for (int q0 = 0; q0 < n4; q0 += 640) {
    float4 x[5];
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const int q = q0 + u * 128 + tid;
      x[u] = (q < n4) ? __ldg(v4 + q) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      const float mx = fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w));
      if (mx >= tau && mx > -INFINITY) {
        const int base = head + 4 * (q0 + u * 128 + tid);
        emit(x[u].x, base);
        emit(x[u].y, base + 1);
        emit(x[u].z, base + 2);
        emit(x[u].w, base + 3);
      }
    }
  }
  __syncthreads();
  if (warp != 0) return;
  const int c = s_c;
  if (c <= 64) {
    unsigned long long k0 = lane < c ? prune_key(s_cp[lane], s_ci[lane]) : ~0ull;
    unsigned long long k1 = 32 + lane < c ? prune_key(s_cp[32 + lane], s_ci[32 + lane]) : ~0ull;
    if (c <= 32) {
      warp_sort32_keys(k0, lane);
    } else {
      warp_sort64_keys(k0, k1, lane);
    }
    const float p0 = k0 != ~0ull ? prune_key_prob(k0) : -INFINITY;
    const float p1 = k1 != ~0ull ? prune_key_prob(k1) : -INFINITY;
    double c0 = p0 > -INFINITY ? (double)p0 : 0.0, c1 = p1 > -INFINITY ? (double)p1 : 0.0;
    warp_scan64(c0, c1, lane);
    int n = min(c, limit);
    if (cutoff_prob < 1.0f) {
      const unsigned b0 = __ballot_sync(0xffffffffu, lane < c && c0 >= (double)cutoff_prob);
      const unsigned b1 = __ballot_sync(0xffffffffu, 32 + lane < c && c1 >= (double)cutoff_prob);
      if (b0) n = min(n, __ffs(b0));
      else if (b1) n = min(n, 32 + __ffs(b1));
    }
    if (lane < n) {
      oid[lane] = (int)(unsigned)k0;
      olp[lane] = logf(p0 + FLT_MIN);
    }
    if (32 + lane < n) {
      oid[32 + lane] = (int)(unsigned)k1;
      olp[32 + lane] = logf(p1 + FLT_MIN);
    }
    if (lane == 0) cnt[row] = n;
    return;
  }
  // ---- many ties at tau (e.g. uniform rows): arg-max rounds with a Kahan-compensated running sum (warp 0) ----
  float last_p = INFINITY;
  int last_i = -1;
  float cum = 0.f, cum_c = 0.f;
  int n = 0;
  const bool compact = c <= PRUNE_CAND_MAX;
  while (n < limit) {
    float bm = -INFINITY;
    int bi = 0x7fffffff;
    if (compact) {
      for (int j = lane; j < c; j += 32) {
        const float pv = s_cp[j];
        const int iv = s_ci[j];
        if (prune_before(last_p, last_i, pv, iv) && prune_before(pv, iv, bm, bi)) {
          bm = pv;
          bi = iv;
        }
      }
    } else {
      for (int j = lane; j < V; j += 32) {
        const float pv = __ldg(src + j);
        if (prune_before(last_p, last_i, pv, j) && prune_before(pv, j, bm, bi)) {
          bm = pv;
          bi = j;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, bm, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (prune_before(om, oi, bm, bi)) {
        bm = om;
        bi = oi;
      }
    }
    if (bi == 0x7fffffff) break;
    if (lane == 0) {
      oid[n] = bi;
      olp[n] = logf(bm + FLT_MIN);
    }
    last_p = bm;
    last_i = bi;
    ++n;
    const float yk = bm - cum_c;
    const float tk = cum + yk;
    cum_c = (tk - cum) - yk;
    cum = tk;
    if (cutoff_prob < 1.0f && cum >= cutoff_prob) break;
  }
  if (lane == 0) cnt[row] = n;
}

// Occupancy: the scan is latency bound per warp (serial phases: stream, sort, re-read, sort), so what saturates HBM is the
// number of rows in flight -- measured on [64,748,4233] (scripts/gpu_prune_sweep.py): 2 / 4 / 8 / 9+ resident CTAs per SM ->
// 25 % / 41 % / 50 % / 61 % of the copy bandwidth. The kernel is therefore built for 12 CTAs (48 warps) per SM (<= 40
// registers, 8 KB static shared memory). PPASR_B200_PRUNE_CTAS caps the residency through a dynamic shared-memory
// reservation (experiments only; default: no cap).
static int prune_ctas_per_sm() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("PPASR_B200_PRUNE_CTAS");
    v = e ? atoi(e) : 16;
    if (v < 1 || v > 16) v = 16;
  }
  return v;
}

cudaError_t launch_ctc_prune(const float* probs, int V, int rows, float cutoff_prob, int top_n, int* cnt, int* cid,
                             float* clp, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  const int per_sm = prune_ctas_per_sm();
  const int dyn = per_sm >= 12 ? 0 : (225 * 1024) / per_sm - 10 * 1024;  // static 8 KB + 1 KB system reservation per CTA
  static std::once_flag once;
  static cudaError_t cfg_err = cudaSuccess;
  std::call_once(once, [] { cfg_err = cudaFuncSetAttribute(ctc_prune_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 216 * 1024); });
  if (cfg_err != cudaSuccess) return cfg_err;
  static int use_cta = -1;
  if (use_cta < 0) {
    const char* e = getenv("PPASR_B200_PRUNE_KERNEL");  // "warp" selects the warp-per-row kernel for A/B runs
    use_cta = (e != nullptr && e[0] == 'w') ? 0 : 1;
  }
  if (use_cta && V >= 512)
    ctc_prune_cta_kernel<<<rows, 128, 0, st>>>(probs, V, rows, cutoff_prob, top_n, cnt, cid, clp);
  else
    ctc_prune_kernel<<<(rows + PRUNE_WARPS - 1) / PRUNE_WARPS, PRUNE_WARPS * 32, dyn > 0 ? dyn : 0, st>>>(probs, V, rows, cutoff_prob,
                                                                                                   top_n, cnt, cid, clp);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// (2) prefix search
// ------------------------------------------------------------------------------------------------
DEVINL float lse2(float a, float b) {  // decoder_utils.h log_sum_exp
  if (a == -INFINITY) return b;
  if (b == -INFINITY) return a;
  const float m = fmaxf(a, b);
  return logf(expf(a - m) + expf(b - m)) + m;
}

// ---- external scorer: back-off n-gram query (scorer.cpp get_log_cond_prob over make_ngram(prefix), character based) ----
// home slot = murmur3 finaliser of the packed token ids (ppasr_b200/decoders/ngram_lm.py::lm_hash64 builds the table with
// the same function): every 16-bit token field reaches the low bits
DEVINL unsigned lm_home_slot(unsigned long long x, unsigned mask) {
  x ^= x >> 33;
  x *= 0xFF51AFD7ED558CCDull;
  x ^= x >> 33;
  x *= 0xC4CEB9FE1A85EC53ull;
  x ^= x >> 33;
  return (unsigned)x & mask;
}
DEVINL bool lm_find(const BeamLm& lm, unsigned long long key, unsigned hi, float2* out) {
  unsigned slot = lm_home_slot(key ^ ((unsigned long long)hi * 0x9E3779B97F4A7C15ull), lm.mask);
  for (;;) {
    const unsigned long long k = __ldg(lm.keys + slot);
    if (k == key && (lm.keys_hi == nullptr || __ldg(lm.keys_hi + slot) == hi)) {
      *out = __ldg(lm.vals + slot);
      return true;
    }
    if (k == 0ull) return false;
    slot = (slot + 1) & lm.mask;
  }
}
// ln P(c | last order-1 tokens of the prefix, left-padded with <s>); any token without a unigram -> OOV_SCORE (-1000)
DEVINL float lm_log_cond_prob(const BeamLm& lm, const BeamEntry& e, int c) {
  if (!__ldg(lm.in_lm + c)) return -1000.0f;
  const int n = lm.order - 1;  // context length (<= 4)
  unsigned long long ctx[4];   // oldest .. most recent
  {
    const int raw[4] = {e.prev3, e.prev2, e.prev1, e.last};
    for (int i = 0; i < 4; ++i) {
      const int v = raw[i];
      if (i >= 4 - n && v >= 0 && !__ldg(lm.in_lm + v)) return -1000.0f;  // only tokens inside the n-gram window
      ctx[i] = v >= 0 ? (unsigned long long)(v + 2) : 1ull;  // <s> padding
    }
  }
  const unsigned long long w = (unsigned long long)(c + 2);
  float bo = 0.f;
  for (int start = 4 - n; start <= 4; ++start) {
    // n-gram ctx[start..3] + w: the four most recent tokens go to the 64-bit key, a fifth (oldest) one to `hi`
    const int len = 4 - start;  // context tokens in this n-gram
    unsigned long long kc = 0ull;
    for (int i = start; i < 4; ++i) kc = (kc << 16) | ctx[i];
    float2 v;
    {
      const unsigned long long lo = (kc << 16) | w;  // with four context tokens the 64-bit shift drops the oldest one
      const unsigned hi = (len == 4) ? (unsigned)ctx[start] : 0u;
      if (lm_find(lm, lo, hi, &v)) return bo + v.x;
    }
    if (start < 4 && lm_find(lm, kc, 0u, &v)) bo += v.y;  // back-off weight of the context (<= 4 tokens)
  }
  return -1000.0f;
}

// candidate -> 64-bit key whose ASCENDING order is prefix_compare's: score desc, then last char asc, then slot asc
DEVINL unsigned long long beam_key(float score, int last, int slot) {
  const unsigned u = __float_as_uint(score);
  const unsigned up = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // monotone increasing in score (-inf lowest)
  return ((unsigned long long)(~up) << 32) | ((unsigned long long)(unsigned)(last + 1) << 16) | (unsigned)slot;
}
DEVINL float beam_key_score(unsigned long long k) {
  const unsigned up = ~(unsigned)(k >> 32);
  return __uint_as_float((up & 0x80000000u) ? (up & 0x7fffffffu) : ~up);
}

struct BeamSmemLayout {  // byte offsets into dynamic shared memory
  int ebuf, stay_b, stay_nb, merge_nb, htab, fid, flp, keys, keys2, total;
  int bcap, hcap, kcap;
};
__host__ __device__ inline int beam_pow2_ge(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}
__host__ __device__ inline BeamSmemLayout beam_smem_layout(int beam, int topn) {
  BeamSmemLayout L;
  L.bcap = (beam + 3) & ~3;
  L.hcap = beam_pow2_ge(4 * beam < 64 ? 64 : 4 * beam);
  L.kcap = beam_pow2_ge(beam * topn + beam < 32 ? 32 : beam * topn + beam);
  int o = 0;
  L.keys = o, o += L.kcap * 8;
  L.keys2 = o, o += (beam <= 32 ? (L.kcap / 2 < 32 ? 32 : L.kcap / 2) : 0) * 8;  // tournament ping-pong buffer (small beams)
  L.ebuf = o, o += 2 * L.bcap * (int)sizeof(BeamEntry);
  L.stay_b = o, o += L.bcap * 4;
  L.stay_nb = o, o += L.bcap * 4;
  L.merge_nb = o, o += L.bcap * 4;
  L.htab = o, o += L.hcap * 4;
  L.fid = o, o += BEAM_MAXC * 4;
  L.flp = o, o += BEAM_MAXC * 4;
  L.total = o;
  return L;
}

template <int NT>
DEVINL void beam_sync() {
  if (NT == 32)
    __syncwarp();
  else
    __syncthreads();
}

template <int NT>
__global__ void __launch_bounds__(NT)
ctc_prefix_beam_kernel(const int* __restrict__ cnt, const int* __restrict__ cid, const float* __restrict__ clp, int T,
                       const int* __restrict__ frame_lens, int beam, int topn, int blank, int max_frames,
                       BeamStateHeader* __restrict__ states, size_t state_stride_bytes, int node_cap,
                       const float* __restrict__ probs, int V, const BeamLm lm) {
  const int b = blockIdx.x;
  uint8_t* sp = reinterpret_cast<uint8_t*>(states) + (size_t)b * state_stride_bytes;
  BeamStateHeader* hdr = reinterpret_cast<BeamStateHeader*>(sp);
  BeamEntry* gbeam = reinterpret_cast<BeamEntry*>(sp + sizeof(BeamStateHeader));
  int2* nodes = reinterpret_cast<int2*>(sp + sizeof(BeamStateHeader) + sizeof(BeamEntry) * BEAM_MAXB);

  extern __shared__ __align__(16) uint8_t dyn_smem[];
  const BeamSmemLayout L = beam_smem_layout(beam, topn);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(dyn_smem + L.keys);
  unsigned long long* keys2 = reinterpret_cast<unsigned long long*>(dyn_smem + L.keys2);
  BeamEntry* ebuf = reinterpret_cast<BeamEntry*>(dyn_smem + L.ebuf);
  float* stay_b = reinterpret_cast<float*>(dyn_smem + L.stay_b);
  float* stay_nb = reinterpret_cast<float*>(dyn_smem + L.stay_nb);
  float* merge_nb = reinterpret_cast<float*>(dyn_smem + L.merge_nb);
  int* htab = reinterpret_cast<int*>(dyn_smem + L.htab);
  int* fid = reinterpret_cast<int*>(dyn_smem + L.fid);
  float* flp = reinterpret_cast<float*>(dyn_smem + L.flp);
  __shared__ int s_nb, s_next_id, s_ncand, s_full_beam;
  __shared__ float s_min_cutoff;

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int nframes = frame_lens ? min(T, frame_lens[b]) : T;
  // a stream fed past the capacity the state was sized for: refuse (the id table would overflow) and flag it
  if (hdr->frames + nframes > max_frames) {
    if (tid == 0) hdr->pad = 1;
    return;
  }
  if (tid == 0) {
    s_nb = hdr->nb;
    s_next_id = hdr->next_id;
  }
  beam_sync<NT>();
  int cur = 0;
  for (int i = tid; i < s_nb; i += NT) ebuf[i] = gbeam[i];

  // candidate list of the next frame, prefetched into registers while the current frame is processed
  constexpr int PF = (BEAM_MAXC + NT - 1) / NT;
  int pf_id[PF];
  float pf_lp[PF];
  int pf_nc = 0;
  auto prefetch = [&](int t) {
    if (t < nframes) {
      const size_t row = (size_t)b * T + t;
      pf_nc = __ldg(cnt + row);
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int i = tid + u * NT;
        if (i < BEAM_MAXC) {
          pf_id[u] = __ldg(cid + row * BEAM_MAXC + i);
          pf_lp[u] = __ldg(clp + row * BEAM_MAXC + i);
        }
      }
    }
  };
  prefetch(0);
  beam_sync<NT>();

  for (int t = 0; t < nframes; ++t) {
    const size_t row = (size_t)b * T + t;
    const int nc = pf_nc;
    const int nb = s_nb;
    BeamEntry* E = ebuf + cur * L.bcap;
    BeamEntry* N = ebuf + (cur ^ 1) * L.bcap;
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      const int i = tid + u * NT;
      if (i < nc) {
        fid[i] = pf_id[u];
        flp[i] = pf_lp[u];
      }
    }
    prefetch(t + 1);
    for (int i = tid; i < nb; i += NT) stay_b[i] = stay_nb[i] = merge_nb[i] = -INFINITY;
    for (int i = tid; i < L.hcap; i += NT) htab[i] = 0;
    if (tid == 0) {
      s_ncand = 0;
      // with a scorer: min_cutoff = worst beam score + ln p(blank) - max(0, beta); applies when the beam is full
      // (ctc_beam_search_decoder.cpp; the beam is kept sorted best-first, so E[nb-1] is the worst entry)
      s_full_beam = 0;
      s_min_cutoff = -INFINITY;
      if (lm.keys != nullptr) {
        const float pb = probs[row * V + blank];
        s_min_cutoff = E[nb - 1].score + (pb > 0.f ? logf(pb) : -INFINITY) - fmaxf(0.f, lm.beta);
        s_full_beam = nb == beam;
      }
    }
    beam_sync<NT>();
    // identity table of the beam: hash -> entry index + 1 (open addressing)
    for (int k = tid; k < nb; k += NT) {
      unsigned slot = (unsigned)(E[k].hash >> 24) & (unsigned)(L.hcap - 1);
      while (atomicCAS(&htab[slot], 0, k + 1) != 0) slot = (slot + 1) & (unsigned)(L.hcap - 1);
    }
    beam_sync<NT>();
    // every (prefix i, char c) pair
    const int npair = nb * nc;
    for (int pidx = tid; pidx < npair; pidx += NT) {
      const int i = pidx / nc, ci = pidx - i * nc;
      const int c = fid[ci];
      const float lp = flp[ci];
      const BeamEntry e = E[i];
      if (s_full_beam && lp + e.score < s_min_cutoff) continue;
      if (c == blank) {
        stay_b[i] = lp + e.score;
        continue;
      }
      float log_p = -INFINITY;
      if (c == e.last) {
        stay_nb[i] = lp + e.nb_prev;
        if (e.b_prev > -INFINITY) log_p = lp + e.b_prev;
      } else {
        log_p = lp + e.score;
      }
      if (lm.keys != nullptr && log_p > -INFINITY) log_p += lm.alpha * lm_log_cond_prob(lm, e, c) + lm.beta;
      // does the extension land on a prefix that is already in the beam? (identity = hash of the id string)
      const unsigned long long hx = e.hash * 0x9E3779B97F4A7C15ull + (unsigned long long)(c + 1);
      int j = -1;
      {
        unsigned slot = (unsigned)(hx >> 24) & (unsigned)(L.hcap - 1);
        for (;;) {
          const int k1 = htab[slot];
          if (k1 == 0) break;
          const BeamEntry& o = E[k1 - 1];
          if (o.hash == hx && o.len == e.len + 1 && o.last == c) {
            j = k1 - 1;
            break;
          }
          slot = (slot + 1) & (unsigned)(L.hcap - 1);
        }
      }
      if (j >= 0) {
        merge_nb[j] = log_p;
      } else if (log_p > -INFINITY) {  // fresh extensions need a finite score
        keys[atomicAdd(&s_ncand, 1)] = beam_key(log_p, c, pidx);
      }
    }
    beam_sync<NT>();
    // prefixes already in the beam stay candidates even at -inf
    for (int j = tid; j < nb; j += NT) {
      const float nbn = lse2(stay_nb[j], merge_nb[j]);
      stay_nb[j] = nbn;
      keys[atomicAdd(&s_ncand, 1)] = beam_key(lse2(stay_b[j], nbn), E[j].last, npair + j);
    }
    beam_sync<NT>();
    // ---- top-`beam` = the first `beam` keys in ascending order ----
    const int n = s_ncand;
    if (n <= 32) {
      if (tid < 32) {
        unsigned long long k = lane < n ? keys[lane] : ~0ull;
        warp_sort32_keys(k, lane);
        keys[lane] = k;
      }
    } else if (beam <= 32) {
      // tournament: every warp sorts 64-key chunks in registers and keeps each chunk's best `keep` >= beam keys; the
      // survivors are sorted again until one chunk is left. The overall best `beam` keys survive every round.
      const int keep = beam <= 4 ? 4 : (beam <= 8 ? 8 : (beam <= 16 ? 16 : 32));
      const int warp = tid >> 5;
      constexpr int NW = NT / 32;
      unsigned long long* src = keys;
      unsigned long long* dst = keys2;
      int cur_n = n;
      for (;;) {
        const int nch = (cur_n + 63) >> 6;
        for (int ch = warp; ch < nch; ch += NW) {
          const int base = ch << 6;
          unsigned long long k0 = base + lane < cur_n ? src[base + lane] : ~0ull;
          unsigned long long k1 = base + 32 + lane < cur_n ? src[base + 32 + lane] : ~0ull;
          warp_sort64_keys(k0, k1, lane);
          if (lane < keep) dst[ch * keep + lane] = k0;
        }
        beam_sync<NT>();
        cur_n = nch * keep;
        unsigned long long* t2 = src;
        src = dst;
        dst = t2;
        if (nch == 1) break;
      }
      if (src != keys) {
        if (tid < keep) keys[tid] = src[tid];
      }
    } else {
      const int np2 = beam_pow2_ge(n);
      for (int i = n + tid; i < np2; i += NT) keys[i] = ~0ull;
      beam_sync<NT>();
      for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int x = tid; x < (np2 >> 1); x += NT) {
            const int lo = ((x & ~(j - 1)) << 1) | (x & (j - 1));
            const int hi = lo | j;
            const unsigned long long a = keys[lo], bq = keys[hi];
            const bool asc = (lo & k) == 0;
            if ((a > bq) == asc) {
              keys[lo] = bq;
              keys[hi] = a;
            }
          }
          beam_sync<NT>();
        }
      }
    }
    beam_sync<NT>();
    // build the next beam (already in best-first order)
    const int nsel = min(n, beam);
    for (int r = tid; r < nsel; r += NT) {
      const unsigned long long k = keys[r];
      const int slot = (int)(k & 0xffffu);
      BeamEntry ne;
      if (slot >= npair) {
        const int j = slot - npair;
        ne = E[j];
        ne.b_prev = stay_b[j];
        ne.nb_prev = stay_nb[j];
        ne.score = lse2(stay_b[j], stay_nb[j]);
      } else {
        const int i = slot / nc, ci = slot - i * nc;
        const BeamEntry e = E[i];
        const int c = fid[ci];
        const float log_p = beam_key_score(k);  // as computed above (includes the scorer terms)
        const int id = atomicAdd(&s_next_id, 1);
        ne.id = id;
        ne.hash = e.hash * 0x9E3779B97F4A7C15ull + (unsigned long long)(c + 1);
        ne.parent_id = e.id;
        ne.last = c;
        ne.prev1 = e.last;
        ne.prev2 = e.prev1;
        ne.prev3 = e.prev2;
        ne.len = e.len + 1;
        ne.b_prev = -INFINITY;
        ne.nb_prev = log_p;
        ne.score = log_p;
        if (id < node_cap) nodes[id] = make_int2(e.id, c);
      }
      N[r] = ne;
    }
    beam_sync<NT>();
    if (tid == 0) s_nb = nsel;
    cur ^= 1;
    beam_sync<NT>();
  }
  // write the state back
  for (int i = tid; i < s_nb; i += NT) gbeam[i] = ebuf[cur * L.bcap + i];
  if (tid == 0) {
    hdr->nb = s_nb;
    hdr->next_id = s_next_id;
    hdr->frames += nframes;
  }
}

// Results: the beam is kept best first (score desc; ties char asc), so entry r IS rank r; ids by walking the
// (parent, char) table. Only the first `nbest` entries are written: out_ids [B, nbest, lmax], out_lens / out_scores
// [B, nbest]; out_lens = -1 where the beam has fewer entries, -2 if the state overflowed its max_frames.
__global__ void ctc_beam_result_kernel(const BeamStateHeader* __restrict__ states, size_t state_stride_bytes,
                                       int node_cap, int nbest, int* __restrict__ out_ids, int lmax,
                                       int* __restrict__ out_lens, float* __restrict__ out_scores) {
  const int b = blockIdx.x;
  const uint8_t* sp = reinterpret_cast<const uint8_t*>(states) + (size_t)b * state_stride_bytes;
  const BeamStateHeader* hdr = reinterpret_cast<const BeamStateHeader*>(sp);
  const BeamEntry* gbeam = reinterpret_cast<const BeamEntry*>(sp + sizeof(BeamStateHeader));
  const int2* nodes = reinterpret_cast<const int2*>(sp + sizeof(BeamStateHeader) + sizeof(BeamEntry) * BEAM_MAXB);
  const int nb = hdr->nb;
  for (int r = threadIdx.x; r < nbest; r += blockDim.x) {
    if (hdr->pad != 0) {
      out_lens[b * nbest + r] = -2;
      out_scores[b * nbest + r] = -INFINITY;
      continue;
    }
    if (r >= nb) {
      out_lens[b * nbest + r] = -1;
      out_scores[b * nbest + r] = -INFINITY;
      continue;
    }
    const BeamEntry me = gbeam[r];
    int* ids = out_ids + ((size_t)b * nbest + r) * lmax;
    const int len = min(me.len, lmax);
    int id = me.id;
    for (int pos = me.len - 1; pos >= 0 && id > 0; --pos) {
      const int2 nd = (id < node_cap) ? nodes[id] : make_int2(0, 0);
      if (pos < lmax) ids[pos] = nd.y;
      id = nd.x;
    }
    out_lens[b * nbest + r] = len;
    out_scores[b * nbest + r] = me.score;
  }
}

__global__ void ctc_beam_reset_kernel(BeamStateHeader* states, size_t state_stride_bytes) {
  uint8_t* sp = reinterpret_cast<uint8_t*>(states) + (size_t)blockIdx.x * state_stride_bytes;
  BeamStateHeader* hdr = reinterpret_cast<BeamStateHeader*>(sp);
  BeamEntry* gbeam = reinterpret_cast<BeamEntry*>(sp + sizeof(BeamStateHeader));
  if (threadIdx.x == 0) {
    hdr->nb = 1;
    hdr->next_id = 1;  // id 0 = root (empty prefix)
    hdr->frames = 0;
    hdr->pad = 0;      // overflow flag
    BeamEntry root;
    root.id = 0, root.parent_id = -1, root.last = -1, root.len = 0;
    root.prev1 = root.prev2 = root.prev3 = -1;
    root.hash = 0x243F6A8885A308D3ull;
    root.b_prev = 0.f;  // root.log_prob_b_prev = 0.0, root.score = 0.0 (ctc_beam_search_decoder.cpp)
    root.nb_prev = -INFINITY;
    root.score = 0.f;
    gbeam[0] = root;
  }
}

size_t beam_state_stride(int node_cap) {
  size_t s = sizeof(BeamStateHeader) + sizeof(BeamEntry) * BEAM_MAXB + sizeof(int2) * (size_t)node_cap;
  return (s + 255) & ~size_t(255);
}

cudaError_t launch_beam_reset(void* states, int B, int node_cap, cudaStream_t st) {
  ctc_beam_reset_kernel<<<B, 32, 0, st>>>(reinterpret_cast<BeamStateHeader*>(states), beam_state_stride(node_cap));
  count_launch();
  return cudaGetLastError();
}

template <int NT>
static cudaError_t launch_beam_advance_nt(const int* cnt, const int* cid, const float* clp, int B, int T,
                                          const int* frame_lens, int beam, int topn, int blank, int max_frames, void* states,
                                          int node_cap, cudaStream_t st, const float* probs, int V, const BeamLm& lm) {
  const BeamSmemLayout L = beam_smem_layout(beam, topn);
  static std::once_flag once;
  static cudaError_t cfg_err = cudaSuccess;
  std::call_once(once, [] {
    cfg_err = cudaFuncSetAttribute(ctc_prefix_beam_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
  });
  if (cfg_err != cudaSuccess) return cfg_err;
  if (L.total > 226 * 1024) return cudaErrorInvalidValue;
  ctc_prefix_beam_kernel<NT><<<B, NT, L.total, st>>>(cnt, cid, clp, T, frame_lens, beam, topn, blank, max_frames,
                                                      reinterpret_cast<BeamStateHeader*>(states), beam_state_stride(node_cap),
                                                      node_cap, probs, V, lm);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_beam_advance(const int* cnt, const int* cid, const float* clp, int B, int T, const int* frame_lens,
                                int beam, int topn, int blank, int max_frames, void* states, int node_cap, cudaStream_t st,
                                const float* probs, int V, const BeamLm* lm) {
  if (beam < 1 || beam > BEAM_MAXB || topn < 1 || topn > BEAM_MAXC) return cudaErrorInvalidValue;
  BeamLm none{};
  const BeamLm& l = lm ? *lm : none;
  // beams <= 32: four warps per utterance, top-`beam` by a register-sort tournament; wider beams: shared-memory bitonic sort
  if (beam <= 32)
    return launch_beam_advance_nt<128>(cnt, cid, clp, B, T, frame_lens, beam, topn, blank, max_frames, states, node_cap, st, probs, V, l);
  if (beam <= 128)
    return launch_beam_advance_nt<256>(cnt, cid, clp, B, T, frame_lens, beam, topn, blank, max_frames, states, node_cap, st, probs, V, l);
  return launch_beam_advance_nt<512>(cnt, cid, clp, B, T, frame_lens, beam, topn, blank, max_frames, states, node_cap, st, probs, V, l);
}

cudaError_t launch_beam_result(const void* states, int B, int node_cap, int nbest, int* out_ids, int lmax, int* out_lens,
                               float* out_scores, cudaStream_t st) {
  ctc_beam_result_kernel<<<B, 128, 0, st>>>(reinterpret_cast<const BeamStateHeader*>(states), beam_state_stride(node_cap),
                                            node_cap, nbest, out_ids, lmax, out_lens, out_scores);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ppasr
