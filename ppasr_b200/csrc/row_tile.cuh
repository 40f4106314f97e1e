// Helpers shared by the row-tile chain kernels (fused_ffn.cu, fused_attn_out.cu): a CTA owns 128 complete
// 256-wide rows; 256 "compute" threads = two per row (128 columns each, TMEM lane = row).
#pragma once
#include "ptx.cuh"

namespace ppasr {

// running (count, mean, M2) LayerNorm statistics, merged with Chan's parallel update
struct FfnStat {
  float n, mean, m2;
};
DEVINL void ffn_chan(FfnStat& a, float nb, float mb, float m2b) {
  const float n = a.n + nb;
  const float d = mb - a.mean;
  a.mean += d * (nb / n);
  a.m2 += m2b + d * d * (a.n * nb / n);
  a.n = n;
}
DEVINL void ffn_add_chunk(FfnStat& a, const float (&v)[32]) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  const float m = s * (1.0f / 32.0f);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) q += (v[j] - m) * (v[j] - m);
  if (a.n == 0.f)
    a.n = 32.f, a.mean = m, a.m2 = q;
  else
    ffn_chan(a, 32.f, m, q);
}
// combine with the partner thread that owns the other 128 columns of the row (fixed order: bit-identical on both)
DEVINL void ffn_exchange(FfnStat& a, float4* sc, int r, int half, int barrier_id) {
  sc[r * 2 + half] = make_float4(a.n, a.mean, a.m2, 0.f);
  named_bar_sync(barrier_id, 256);
  const float4 o = sc[r * 2 + (half ^ 1)];
  const float4 lo = half ? o : make_float4(a.n, a.mean, a.m2, 0.f);
  const float4 hi = half ? make_float4(a.n, a.mean, a.m2, 0.f) : o;
  FfnStat t{lo.x, lo.y, lo.z};
  ffn_chan(t, hi.x, hi.y, hi.z);
  a = t;
}


}  // namespace ppasr
