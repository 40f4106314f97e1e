// CTC prefix beam search on the GPU (no external scorer).
//
// Replaces the C++ library behind ppasr/decoders/swig_wrapper.py:35-121 / beam_search_decoder.py:45-96
// (`paddlespeech_ctcdecoders`: ctc_beam_search_decoder.cpp, decoder_utils.cpp, path_trie.cpp -- not vendored
// in the reference, restated in oracle/decoders_oracle.py).
//
// Two kernels:
//  (1) ctc_prune_kernel -- the HBM-bound scan of the posterior [B,T,V]: one warp per frame stages the
//      row in shared memory with 16-byte loads and extracts, in descending probability order (ties: lower
//      index first), the smallest head whose cumulative probability reaches cutoff_prob, capped at
//      cutoff_top_n (decoder_utils.cpp get_pruned_log_probs). Output per frame: n, ids[n], log(p + FLT_MIN)[n].
//      Every frame is independent, so the scan runs at full grid width; the sequential part never touches
//      the [B,T,V] tensor.
//  (2) ctc_prefix_beam_kernel -- one CTA per utterance walks the frames over the compact candidate lists.
//      The prefix trie of the reference is replaced by an equivalent flat form: a beam entry carries a 64-bit
//      hash of its id string (its identity) and an id into a (parent id, char) table used only to read the
//      string back; an extension (prefix i, char c) merges into beam entry j iff hash[j] == H(hash[i], c),
//      otherwise it is a fresh prefix (a revived trie node is reset to -inf in path_trie.cpp, i.e.
//      indistinguishable from a fresh one).
//      Top-`beam` selection per frame = `beam` rounds of a block-wide arg-max (score desc, then char asc
//      like prefix_compare). State (beam + id table) lives in global memory so the same kernel serves the
//      streaming decode_chunk API.
#include <float.h>

#include "kernels.h"
#include "ptx.cuh"

namespace ppasr {

void count_launch();

constexpr int BEAM_MAXB = 128;   // max beam size
constexpr int BEAM_MAXC = 64;    // max cutoff_top_n
constexpr int BEAM_THREADS = 256;

// ------------------------------------------------------------------------------------------------
// (1) pruning scan
// ------------------------------------------------------------------------------------------------
// Selection order everywhere: probability descending, index ascending on ties.
DEVINL bool prune_before(float pa, int ia, float pb, int ib) { return pa > pb || (pa == pb && ia < ib); }

constexpr int PRUNE_CAND_MAX = 256;  // compacted candidates per row; more (mass ties) falls back to full-row rounds
constexpr int PRUNE_WARPS = 4;

// One warp per frame, no row staging (occupancy, i.e. bytes in flight, is what saturates HBM):
//   pass 1 streams the row from HBM with 16-byte loads (8 in flight per lane) keeping each lane's two largest values;
//          the limit-th largest of those 64 values is a lower bound tau of the limit-th largest of the row;
//   pass 2 re-reads the row (L2 hits: 17 KB touched microseconds earlier) and compacts every element >= tau;
//   then at most `limit` arg-max rounds over the (typically ~64) candidates, with the cumulative cut-off.
__global__ void __launch_bounds__(PRUNE_WARPS * 32) ctc_prune_kernel(const float* __restrict__ probs, int V, int rows,
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
  // ---- threshold: walk the 64 per-lane maxima in descending order. They are row elements, so the k-th of them is a lower
  //      bound of the k-th largest of the row and their running sum a lower bound of the true cumulative mass: as soon as it
  //      reaches cutoff_prob (or k = limit) everything that can be selected is >= tau = the value just taken.
  float tau = -INFINITY;
  {
    float a1 = t1, a2 = t2;
    float cumx = 0.f;
    for (int k = 0; k < limit; ++k) {
      const float m = warp_max(fmaxf(a1, a2));
      tau = m;
      if (m == -INFINITY) break;
      cumx += m;
      if (cutoff_prob < 1.0f && cumx >= cutoff_prob * 1.0001f) break;  // margin: fp32 sum vs the Kahan sum used below
      const unsigned has = __ballot_sync(0xffffffffu, a1 == m || a2 == m);
      if (lane == __ffs(has) - 1) {
        if (a1 == m)
          a1 = -INFINITY;
        else
          a2 = -INFINITY;
      }
    }
    if (!prune) tau = -INFINITY;
  }
  // ---- pass 2: compact elements >= tau (warp-uniform control flow; lane 0 writes) ----
  int c = 0;
  auto emit = [&](float ev, int idx) {  // called with warp-uniform arguments
    if (ev >= tau && ev > -INFINITY) {
      if (c < PRUNE_CAND_MAX && lane == 0) {
        cp[c] = ev;
        ci[c] = idx;
      }
      ++c;
    }
  };
  {
    const float hv = (lane < head) ? src[lane] : -INFINITY;
    const float tv = (tail0 + lane < V) ? src[tail0 + lane] : -INFINITY;
    unsigned hit = __ballot_sync(0xffffffffu, hv >= tau && hv > -INFINITY);
    while (hit) {
      const int sl = __ffs(hit) - 1;
      hit &= hit - 1;
      emit(__shfl_sync(0xffffffffu, hv, sl), sl);
    }
    hit = __ballot_sync(0xffffffffu, tv >= tau && tv > -INFINITY);
    while (hit) {
      const int sl = __ffs(hit) - 1;
      hit &= hit - 1;
      emit(__shfl_sync(0xffffffffu, tv, sl), tail0 + sl);
    }
  }
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
      unsigned hit = __ballot_sync(0xffffffffu, mx >= tau && mx > -INFINITY);
      while (hit) {
        const int sl = __ffs(hit) - 1;
        hit &= hit - 1;
        const int base = head + 4 * (q0 + u * 32 + sl);
        emit(__shfl_sync(0xffffffffu, x[u].x, sl), base);
        emit(__shfl_sync(0xffffffffu, x[u].y, sl), base + 1);
        emit(__shfl_sync(0xffffffffu, x[u].z, sl), base + 2);
        emit(__shfl_sync(0xffffffffu, x[u].w, sl), base + 3);
      }
    }
  }
  __syncwarp();
  float last_p = INFINITY;
  int last_i = -1;
  float cum = 0.f, cum_c = 0.f;  // Kahan-compensated fp32 running sum (stands in for the reference's double)
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
    } else {  // mass ties (e.g. uniform rows): rounds over the whole row
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

cudaError_t launch_ctc_prune(const float* probs, int V, int rows, float cutoff_prob, int top_n, int* cnt, int* cid,
                             float* clp, cudaStream_t st) {
  if (rows <= 0) return cudaSuccess;
  ctc_prune_kernel<<<(rows + PRUNE_WARPS - 1) / PRUNE_WARPS, PRUNE_WARPS * 32, 0, st>>>(probs, V, rows, cutoff_prob, top_n, cnt,
                                                                                 cid, clp);
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
DEVINL bool lm_find(const BeamLm& lm, unsigned long long key, float2* out) {
  unsigned slot = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 17) & lm.mask;
  for (;;) {
    const unsigned long long k = __ldg(lm.keys + slot);
    if (k == key) {
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
  const int n = lm.order - 1;  // context length
  unsigned long long ctx[3];   // oldest .. most recent
  {
    const int raw[3] = {e.prev2, e.prev1, e.last};
    for (int i = 0; i < 3; ++i) {
      const int v = raw[i];
      if (i >= 3 - n && v >= 0 && !__ldg(lm.in_lm + v)) return -1000.0f;  // only tokens inside the n-gram window
      ctx[i] = v >= 0 ? (unsigned long long)(v + 2) : 1ull;  // <s> padding
    }
  }
  const unsigned long long w = (unsigned long long)(c + 2);
  float bo = 0.f;
  for (int start = 3 - n; start <= 3; ++start) {
    unsigned long long kc = 0ull;
    for (int i = start; i < 3; ++i) kc = (kc << 16) | ctx[i];
    float2 v;
    if (lm_find(lm, (kc << 16) | w, &v)) return bo + v.x;
    if (start < 3 && lm_find(lm, kc, &v)) bo += v.y;
  }
  return -1000.0f;
}

__global__ void __launch_bounds__(BEAM_THREADS)
ctc_prefix_beam_kernel(const int* __restrict__ cnt, const int* __restrict__ cid, const float* __restrict__ clp, int T,
                       const int* __restrict__ frame_lens, int beam, int blank, BeamStateHeader* __restrict__ states,
                       size_t state_stride_bytes, int node_cap, const float* __restrict__ probs, int V, const BeamLm lm) {
  const int b = blockIdx.x;
  uint8_t* sp = reinterpret_cast<uint8_t*>(states) + (size_t)b * state_stride_bytes;
  BeamStateHeader* hdr = reinterpret_cast<BeamStateHeader*>(sp);
  BeamEntry* gbeam = reinterpret_cast<BeamEntry*>(sp + sizeof(BeamStateHeader));
  int2* nodes = reinterpret_cast<int2*>(sp + sizeof(BeamStateHeader) + sizeof(BeamEntry) * BEAM_MAXB);

  __shared__ BeamEntry ebuf[2][BEAM_MAXB];
  __shared__ float stay_b[BEAM_MAXB], stay_nb[BEAM_MAXB], merge_nb[BEAM_MAXB];
  extern __shared__ uint8_t dyn_smem[];
  float* cscore = reinterpret_cast<float*>(dyn_smem);                                   // [beam*MAXC + beam]
  int* ckey = reinterpret_cast<int*>(dyn_smem) + (beam * BEAM_MAXC + beam);             // tie-break key (char); INT_MAX = consumed
  __shared__ int sel[BEAM_MAXB];
  __shared__ int fid[BEAM_MAXC];
  __shared__ float flp[BEAM_MAXC];
  __shared__ float red_s[BEAM_THREADS / 32];
  __shared__ int red_k[BEAM_THREADS / 32], red_i[BEAM_THREADS / 32];
  __shared__ int s_nb, s_next_id;
  __shared__ float s_min_cutoff;
  __shared__ int s_full_beam;

  const int tid = threadIdx.x;
  if (tid == 0) {
    s_nb = hdr->nb;
    s_next_id = hdr->next_id;
  }
  __syncthreads();
  int cur = 0;
  for (int i = tid; i < s_nb; i += BEAM_THREADS) ebuf[0][i] = gbeam[i];
  __syncthreads();

  const int nframes = frame_lens ? min(T, frame_lens[b]) : T;
  for (int t = 0; t < nframes; ++t) {
    const size_t row = (size_t)b * T + t;
    const int nc = cnt[row];
    const int nb = s_nb;
    BeamEntry* E = ebuf[cur];
    BeamEntry* N = ebuf[cur ^ 1];
    for (int i = tid; i < nc; i += BEAM_THREADS) {
      fid[i] = cid[row * BEAM_MAXC + i];
      flp[i] = clp[row * BEAM_MAXC + i];
    }
    for (int i = tid; i < nb; i += BEAM_THREADS) stay_b[i] = stay_nb[i] = merge_nb[i] = -INFINITY;
    const int npair = nb * nc;
    for (int i = tid; i < npair + nb; i += BEAM_THREADS) {
      cscore[i] = -INFINITY;
      ckey[i] = 0;
    }
    if (tid == 0) {
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
    __syncthreads();
    // every (prefix i, char c) pair
    for (int pidx = tid; pidx < npair; pidx += BEAM_THREADS) {
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
      for (int k = 0; k < nb; ++k)
        if (E[k].hash == hx && E[k].len == e.len + 1 && E[k].last == c) {
          j = k;
          break;
        }
      if (j >= 0) {
        merge_nb[j] = log_p;
      } else {
        cscore[pidx] = log_p;
        ckey[pidx] = c;
      }
    }
    __syncthreads();
    for (int j = tid; j < nb; j += BEAM_THREADS) {
      const float nbn = lse2(stay_nb[j], merge_nb[j]);
      stay_nb[j] = nbn;
      cscore[npair + j] = lse2(stay_b[j], nbn);
      ckey[npair + j] = E[j].last;
    }
    __syncthreads();
    // top-`beam` selection: rounds of block-wide arg-max (score desc, char asc, slot asc)
    const int nslots = npair + nb;
    int nsel = 0;
    for (int r = 0; r < beam; ++r) {
      float bs = -INFINITY;
      int bk = 0x7fffffff, bi = -1;
      for (int i = tid; i < nslots; i += BEAM_THREADS) {
        const float sc = cscore[i];
        const int k = ckey[i];
        // fresh extensions need a finite score; prefixes already in the beam stay candidates even at -inf
        if (k != 0x7fffffff && (i >= npair || sc > -INFINITY)) {
          if (bi < 0 || sc > bs || (sc == bs && (k < bk || (k == bk && i < bi)))) {
            bs = sc;
            bk = k;
            bi = i;
          }
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float os = __shfl_xor_sync(0xffffffffu, bs, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (oi >= 0 && (bi < 0 || os > bs || (os == bs && (ok < bk || (ok == bk && oi < bi))))) {
          bs = os;
          bk = ok;
          bi = oi;
        }
      }
      if ((tid & 31) == 0) {
        red_s[tid >> 5] = bs;
        red_k[tid >> 5] = bk;
        red_i[tid >> 5] = bi;
      }
      __syncthreads();
      if (tid == 0) {
        for (int w = 1; w < BEAM_THREADS / 32; ++w) {
          const float os = red_s[w];
          const int ok = red_k[w], oi = red_i[w];
          if (oi >= 0 && (bi < 0 || os > bs || (os == bs && (ok < bk || (ok == bk && oi < bi))))) {
            bs = os;
            bk = ok;
            bi = oi;
          }
        }
        sel[r] = bi;
        if (bi >= 0) ckey[bi] = 0x7fffffff;  // consumed
      }
      __syncthreads();
      if (sel[r] < 0) break;
      ++nsel;
    }
    // build the next beam
    for (int r = tid; r < nsel; r += BEAM_THREADS) {
      const int slot = sel[r];
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
        const float log_p = cscore[slot];  // as computed above (includes the scorer terms)
        const int id = atomicAdd(&s_next_id, 1);
        ne.id = id;
        ne.hash = e.hash * 0x9E3779B97F4A7C15ull + (unsigned long long)(c + 1);
        ne.parent_id = e.id;
        ne.last = c;
        ne.prev1 = e.last;
        ne.prev2 = e.prev1;
        ne.len = e.len + 1;
        ne.b_prev = -INFINITY;
        ne.nb_prev = log_p;
        ne.score = log_p;
        if (id < node_cap) nodes[id] = make_int2(e.id, c);
      }
      N[r] = ne;
    }
    __syncthreads();
    if (tid == 0) s_nb = nsel;
    cur ^= 1;
    __syncthreads();
  }
  // write the state back
  for (int i = tid; i < s_nb; i += BEAM_THREADS) gbeam[i] = ebuf[cur][i];
  if (tid == 0) {
    hdr->nb = s_nb;
    hdr->next_id = s_next_id;
    hdr->frames += nframes;
  }
}

// Results: beam entries sorted by score (desc; ties char asc), ids by walking the (parent, char) table.
__global__ void ctc_beam_result_kernel(const BeamStateHeader* __restrict__ states, size_t state_stride_bytes,
                                       int node_cap, int beam, int* __restrict__ out_ids, int lmax,
                                       int* __restrict__ out_lens, float* __restrict__ out_scores) {
  const int b = blockIdx.x;
  const uint8_t* sp = reinterpret_cast<const uint8_t*>(states) + (size_t)b * state_stride_bytes;
  const BeamStateHeader* hdr = reinterpret_cast<const BeamStateHeader*>(sp);
  const BeamEntry* gbeam = reinterpret_cast<const BeamEntry*>(sp + sizeof(BeamStateHeader));
  const int2* nodes = reinterpret_cast<const int2*>(sp + sizeof(BeamStateHeader) + sizeof(BeamEntry) * BEAM_MAXB);
  const int nb = hdr->nb;
  const int r = threadIdx.x;
  if (r >= beam) return;
  int* ids = out_ids + ((size_t)b * beam + r) * lmax;
  if (r >= nb) {
    out_lens[b * beam + r] = -1;
    out_scores[b * beam + r] = -INFINITY;
    return;
  }
  // rank of entry r
  const BeamEntry me = gbeam[r];
  int rank = 0;
  for (int k = 0; k < nb; ++k) {
    const BeamEntry o = gbeam[k];
    if (k != r && (o.score > me.score || (o.score == me.score && (o.last < me.last || (o.last == me.last && k < r))))) ++rank;
  }
  ids = out_ids + ((size_t)b * beam + rank) * lmax;
  const int len = min(me.len, lmax);
  int id = me.id;
  for (int pos = me.len - 1; pos >= 0 && id > 0; --pos) {
    const int2 nd = (id < node_cap) ? nodes[id] : make_int2(0, 0);
    if (pos < lmax) ids[pos] = nd.y;
    id = nd.x;
  }
  out_lens[b * beam + rank] = len;
  out_scores[b * beam + rank] = me.score;
}

__global__ void ctc_beam_reset_kernel(BeamStateHeader* states, size_t state_stride_bytes) {
  uint8_t* sp = reinterpret_cast<uint8_t*>(states) + (size_t)blockIdx.x * state_stride_bytes;
  BeamStateHeader* hdr = reinterpret_cast<BeamStateHeader*>(sp);
  BeamEntry* gbeam = reinterpret_cast<BeamEntry*>(sp + sizeof(BeamStateHeader));
  if (threadIdx.x == 0) {
    hdr->nb = 1;
    hdr->next_id = 1;  // id 0 = root (empty prefix)
    hdr->frames = 0;
    BeamEntry root;
    root.id = 0, root.parent_id = -1, root.last = -1, root.len = 0;
    root.prev1 = root.prev2 = -1, root.pad = 0;
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

cudaError_t launch_beam_advance(const int* cnt, const int* cid, const float* clp, int B, int T, const int* frame_lens,
                                int beam, int blank, void* states, int node_cap, cudaStream_t st, const float* probs, int V,
                                const BeamLm* lm) {
  if (beam < 1 || beam > BEAM_MAXB) return cudaErrorInvalidValue;
  const size_t dyn = (size_t)(beam * BEAM_MAXC + beam) * 8;
  static size_t configured = 0;
  if (dyn > configured) {
    cudaError_t e = cudaFuncSetAttribute(ctc_prefix_beam_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)((size_t)(BEAM_MAXB * BEAM_MAXC + BEAM_MAXB) * 8));
    if (e != cudaSuccess) return e;
    configured = (size_t)(BEAM_MAXB * BEAM_MAXC + BEAM_MAXB) * 8;
  }
  BeamLm none{};
  ctc_prefix_beam_kernel<<<B, BEAM_THREADS, dyn, st>>>(cnt, cid, clp, T, frame_lens, beam, blank,
                                                     reinterpret_cast<BeamStateHeader*>(states),
                                                     beam_state_stride(node_cap), node_cap, probs, V, lm ? *lm : none);
  count_launch();
  return cudaGetLastError();
}

cudaError_t launch_beam_result(const void* states, int B, int node_cap, int beam, int* out_ids, int lmax, int* out_lens,
                               float* out_scores, cudaStream_t st) {
  ctc_beam_result_kernel<<<B, BEAM_MAXB, 0, st>>>(reinterpret_cast<const BeamStateHeader*>(states),
                                                 beam_state_stride(node_cap), node_cap, beam, out_ids, lmax, out_lens,
                                                 out_scores);
  count_launch();
  return cudaGetLastError();
}

}  // namespace ppasr
