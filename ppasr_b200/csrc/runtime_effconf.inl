// Efficient Conformer (model_type 3) on the Conformer kernels (included by runtime.cu inside its anonymous namespace).
//
// Reference: ppasr/model_utils/efficient_conformer/encoder.py:212-264 (forward), :455-548 (stride block),
// attention.py:128-193 (grouped attention), convolution.py:80-138 (strided conv module). Same pre-norm macaron block and
// the same fused kernels as run_encoder(); what changes per block:
//   * grouped blocks: QKV epilogue writes the (T/3 tokens x 4 heads x 192) views, grouped_attention_kernel;
//   * the stride block: strided depthwise conv, AvgPool1D(2, ceil) on the residual, then everything (rows, lengths,
//     positional rows) at half rate; blocks after it use depthwise kernel 7.

int run_encoder_effconf(ppasr_b200_ctx* c, cudaStream_t st) {
  Plan& p = c->plan;
  const auto& cfg = c->cfg;
  const int D = cfg.d_model, H = cfg.n_heads, FF = cfg.ffn_dim, L = cfg.n_layers;
  const float eps = 1e-5f;
  int rc = run_subsampling_convs(c, st);
  if (rc) return rc;
  float* xc = p.x;
  int Tc = p.Tp, Mc = p.M;
  const int* vl = p.vlen;
  const CUtensorMap* tmpos = &c->tm_pos;
  {
    EpiResidLN<BN_WIDE> e{p.x, c->emb_b, D, p.M, D, std::sqrt((float)D), 0, nullptr, p.Tp, 0, 0, c->layers[0].ln_ffm_g,
                          c->layers[0].ln_ffm_b, nullptr, nullptr, p.y, eps};
    PROF(PC_EMBED);
    PPASR_CUDA_CHECK((gemm<BN_WIDE, ST_WIDE>(c, p.tm_c2, c->tm_emb_w, p.M, D, c->Kemb, e, st)));
  }
  int gi = 0;  // index of the grouped block (its padded positional rows live in p.pg slot gi)
  for (int l = 0; l < L; ++l) {
    const LayerW& w = c->layers[l];
    const auto& m = c->lmaps[l];
    const int K = c->layer_k[l];
    const bool grouped = (c->eff_group_mask >> l) & 1;
    const bool strided = l == c->eff_stride_idx;
    // ---- macaron FFN (conformer/encoder.py:380-390)
    {
      PROF(PC_FUSED_FFN);
      PPASR_CUDA_CHECK(launch_fused_ffn(p.tm_y, nullptr, m.ffm_w1_128, m.ffm_w2s, Mc, FF, xc, p.y, w.ffm_b1, w.ffm_b2s, w.ln_mha_g,
                                        w.ln_mha_b, nullptr, nullptr, eps, nullptr, nullptr, nullptr, nullptr, Tc, st));
    }
    // ---- attention
    if (grouped) {
      // pad4group of p = linear_pos(pos_emb): rows >= T are zero (attention.py:73-77)
      PPASR_CUDA_CHECK(launch_grouped_pos(c->pos_tab, L * D, l * D, Tc, 3 * p.Tg, p.pg + (size_t)gi * p.Tg * 768, st));
      EpiQKVGrouped<BN_NARROW> e{p.q2g, p.kkg, p.vtg, w.bqkv, w.pos_u, w.pos_v, Mc, Tc, H, p.Tg, p.Tgp};
      { PROF(PC_QKV); PPASR_CUDA_CHECK((gemm<BN_NARROW, ST_NARROW>(c, p.tm_y, m.wqkv, Mc, 3 * D, D, e, st))); }
      GroupedAttnParams gp{p.B, H, Tc, p.Tg, vl, p.att};
      { PROF(PC_ATTENTION); PPASR_CUDA_CHECK(launch_grouped_attention(p.tm_qg, p.tm_kg, p.tm_pg[gi], p.tm_vtg, gp, st)); }
      ++gi;
    } else {
      AttnParams ap{};
      ap.B = p.B, ap.H = H, ap.T1 = Tc, ap.D = D, ap.pos_col0 = l * D, ap.out = p.att, ap.q_rows_per_bh = Tc;
      EpiQKV<BN_NARROW> e{p.q2, p.kk, p.vt, w.bqkv, w.pos_u, w.pos_v, Mc, Tc, H, Tc, p.Tkp, 0};
      { PROF(PC_QKV); PPASR_CUDA_CHECK((gemm<BN_NARROW, ST_NARROW>(c, p.tm_y, m.wqkv, Mc, 3 * D, D, e, st))); }
      ap.T2 = Tc, ap.k_rows_per_bh = Tc, ap.k_row0 = 0, ap.pos_row0 = 0, ap.klens = vl;
      { PROF(PC_ATTENTION); PPASR_CUDA_CHECK(launch_rel_attention(p.tm_q, p.tm_k, *tmpos, p.tm_vt, ap, st)); }
    }
    // ---- x += Wo att ; y = mask(norm_conv(x)) ; g = GLU(pw1 y)
    {
      PROF(PC_FUSED_ATTN_OUT);
      PPASR_CUDA_CHECK(launch_fused_attn_out(p.tm_att, m.wo, m.pw1, Mc, xc, p.g, w.bo, w.ln_conv_g, w.ln_conv_b, w.pw1_b, vl, Tc,
                                             eps, st));
    }
    // ---- depthwise conv + norm + swish -> z (half rate in the stride block)
    const int lpad = cfg.causal ? K - 1 : (K - 1) / 2;
    const bool fdw = !strided && cfg.causal && c->fused_dwconv;
    if (fdw) {
      // computed in the prologue of the chained FFN kernel below
    } else if (!strided) {
      PROF(PC_DWCONV);
      PPASR_CUDA_CHECK(launch_dwconv_norm_swish(p.g, w.dw_w, w.dw_b, cfg.causal ? w.glu_pad : nullptr, w.cn_g, w.cn_b,
                                                cfg.conv_norm == 0, p.z, p.B, Tc, Tc, D, K, lpad, eps, vl, st));
    } else {
      const int T2 = (Tc + 1) / 2;
      PPASR_CUDA_CHECK(launch_halve_lens(p.vlen, p.vlen2, p.B, st));
      { PROF(PC_DWCONV);
        PPASR_CUDA_CHECK(launch_dwconv_stride(p.g, w.dw_w, w.dw_b, cfg.causal ? w.glu_pad : nullptr, w.cn_g, w.cn_b,
                                              cfg.conv_norm == 0, p.z, p.B, Tc, T2, D, K, lpad, 2, eps, p.vlen2, st)); }
      // residual through AvgPool1D(2, 2, ceil_mode) (efficient_conformer/encoder.py:523-526)
      PPASR_CUDA_CHECK(launch_avgpool2(xc, p.x2, p.B, Tc, T2, D, st));
      xc = p.x2, Tc = T2, Mc = p.B * T2, vl = p.vlen2, tmpos = &c->tm_pos2;
    }
    // ---- x += mask * (pw2 z) ; FFN ; norm_final ; y = next block's first LayerNorm (or after_norm)
    {
      const float* g2 = (l + 1 < L) ? c->layers[l + 1].ln_ffm_g : c->after_g;
      const float* b2 = (l + 1 < L) ? c->layers[l + 1].ln_ffm_b : c->after_b;
      PROF(PC_FUSED_FFN);
      FfnDw dw{p.g, w.dw_w, w.dw_b, w.glu_pad, w.cn_g, w.cn_b, K, cfg.conv_norm == 0};
      PPASR_CUDA_CHECK(launch_fused_ffn(p.tm_z, &m.pw2, m.ff_w1_128, m.ff_w2s, Mc, FF, xc, p.y, w.ff_b1, w.ff_b2s, w.ln_fin_g,
                                        w.ln_fin_b, g2, b2, eps, w.pw2_b, w.ln_ff_g, w.ln_ff_b, vl, Tc, st, 0, nullptr, nullptr,
                                        nullptr, fdw ? &dw : nullptr));
    }
  }
  return PPASR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// forward_chunk (efficient_conformer/encoder.py:266-394) on device-resident, APPEND-ONLY caches. PPASRPredictor streams with
// required_cache_size < 0 (unlimited history, predict.py:304-306), so the first cached key is always absolute frame 0 and
// pad4group's re-grouping from the first cached frame (attention.py:153-160) is simply group = absolute_frame // 3:
//   * grouped blocks keep K as [B,H,group,192] and V^T as [B,H,192,group] (the EpiQKVGrouped operand layout) inside the
//     same per-layer cache slices the plain blocks use; a new frame f lands in group f / 3 at feature (f % 3) * 256 + c;
//     the caches start zeroed, so the missing frames of a partially filled last group read as zero (a real, unmasked key);
//   * queries are grouped from the CHUNK start (zero padded before pos_bias_u/v are added);
//   * p = linear_pos(pos_emb[0 : t_total]) zero-padded to a multiple of 3 is rebuilt per chunk (grouped_pos_kernel);
//   * blocks after the stride block run at half rate: keys at positions 2 j (tm_pos2), caches hold kend / 2 frames
//     (the reference stores them repeated x2 and reads ::2, encoder.py:351,368), conv caches of K/2 - 1 = 6 rows;
//   * the stride block's stride-2 "valid" conv over [cache 14 | chunk] is aligned with the offline run because chunks
//     start at even frames; its residual goes through AvgPool1D(2, 2, ceil).
// tests/test_effconf_chunk_layout_spec_cpu.py states this layout in NumPy and checks it against the reference code's own
// streaming outputs.
constexpr int EFF_GCAP = 256;  // key groups a grouped block can attend (grouped_attention_kernel keeps the score row in TMEM)

int run_encoder_effconf_chunk(ppasr_b200_ctx* c, cudaStream_t st) {
  Plan& p = c->plan;
  auto& ss = c->ss;
  const auto& cfg = c->cfg;
  const int D = cfg.d_model, H = cfg.n_heads, FF = cfg.ffn_dim, L = cfg.n_layers, B = p.B;
  const int lmax = cfg.conv_kernel - 1;
  const float eps = 1e-5f;
  int rc = run_subsampling_convs(c, st);
  if (rc) return rc;
  float* xc = p.x;
  int Tc = p.Tp, Mc = p.M;
  int t0 = ss.kend;  // cached frames at the current rate
  const CUtensorMap* tmpos = &c->tm_pos;
  int rate = 1;
  {
    EpiResidLN<BN_WIDE> e{p.x, c->emb_b, D, p.M, D, std::sqrt((float)D), 0, nullptr, p.Tp, 0, 0, c->layers[0].ln_ffm_g,
                          c->layers[0].ln_ffm_b, nullptr, nullptr, p.y, eps};
    PROF(PC_EMBED);
    PPASR_CUDA_CHECK((gemm<BN_WIDE, ST_WIDE>(c, p.tm_c2, c->tm_emb_w, p.M, D, c->Kemb, e, st)));
  }
  int gi = 0;
  for (int l = 0; l < L; ++l) {
    const LayerW& w = c->layers[l];
    const auto& m = c->lmaps[l];
    const int K = c->layer_k[l];
    const bool grouped = (c->eff_group_mask >> l) & 1;
    const bool strided = l == c->eff_stride_idx;
    const size_t lk = (size_t)l * B * H * ss.Tcap * 64;
    // ---- macaron FFN
    {
      PROF(PC_FUSED_FFN);
      PPASR_CUDA_CHECK(launch_fused_ffn(p.tm_y, nullptr, m.ffm_w1_128, m.ffm_w2s, Mc, FF, xc, p.y, w.ffm_b1, w.ffm_b2s, w.ln_mha_g,
                                        w.ln_mha_b, nullptr, nullptr, eps, nullptr, nullptr, nullptr, nullptr, Tc, st));
    }
    // ---- attention over [cache | chunk]
    const int T2 = t0 + Tc;
    if (grouped) {
      const int G2 = (T2 + 2) / 3;
      // positional operand of the T2 keys at positions j * rate, zero beyond T2 inside the last group
      PPASR_CUDA_CHECK(launch_grouped_pos(c->pos_tab, L * D * rate, l * D, T2, 3 * G2, ss.pgc + (size_t)gi * EFF_GCAP * 768, st));
      EpiQKVGrouped<BN_NARROW> e{p.q2g, ss.kk + lk, ss.vt + lk, w.bqkv, w.pos_u, w.pos_v, Mc, Tc, H, (Tc + 2) / 3, EFF_GCAP};
      e.kofs = t0, e.Tgk = EFF_GCAP;
      { PROF(PC_QKV); PPASR_CUDA_CHECK((gemm<BN_NARROW, ST_NARROW>(c, p.tm_y, m.wqkv, Mc, 3 * D, D, e, st))); }
      GroupedAttnParams gp{B, H, Tc, (Tc + 2) / 3, nullptr, p.att};
      gp.Tgk = G2, gp.k_pitch = EFF_GCAP;
      { PROF(PC_ATTENTION); PPASR_CUDA_CHECK(launch_grouped_attention(ss.tm_qgc, ss.tm_k[l], ss.tm_pgc[gi], ss.tm_vt[l], gp, st)); }
      ++gi;
    } else {
      AttnParams ap{};
      ap.B = B, ap.H = H, ap.T1 = Tc, ap.D = D, ap.pos_col0 = l * D, ap.out = p.att, ap.q_rows_per_bh = Tc;
      EpiQKV<BN_NARROW> e{p.q2, ss.kk + lk, ss.vt + lk, w.bqkv, w.pos_u, w.pos_v, Mc, Tc, H, ss.Tcap, ss.Tcap, t0};
      { PROF(PC_QKV); PPASR_CUDA_CHECK((gemm<BN_NARROW, ST_NARROW>(c, p.tm_y, m.wqkv, Mc, 3 * D, D, e, st))); }
      ap.T2 = T2, ap.k_rows_per_bh = ss.Tcap, ap.k_row0 = 0, ap.pos_row0 = 0, ap.klens = nullptr;
      { PROF(PC_ATTENTION); PPASR_CUDA_CHECK(launch_rel_attention(p.tm_q, ss.tm_k[l], *tmpos, ss.tm_vt[l], ap, st)); }
    }
    // ---- x += Wo att + bo ; y = norm_conv(x)   (no padding inside a chunk)
    {
      EpiResidLN<BN_WIDE> e{xc, w.bo, D, Mc, D, 1.0f, 1, nullptr, Tc, 0, 0, w.ln_conv_g, w.ln_conv_b, nullptr, nullptr, p.y, eps};
      PROF(PC_OUTPROJ);
      PPASR_CUDA_CHECK((gemm<BN_WIDE, ST_WIDE>(c, p.tm_att, m.wo, Mc, D, D, e, st)));
    }
    // ---- conv module over [cnn_cache ; chunk]: pw1 + GLU, "valid" depthwise conv (stride 2 in the stride block)
    const int lorder = K - 1;
    const int Tcat = lorder + Tc, Mcat = B * Tcat;
    PPASR_CUDA_CHECK(launch_conv_cache_concat(ss.cnn + (size_t)l * B * lmax * D, p.y, p.ycat, B, Tc, lorder, D, st, nullptr));
    {
      EpiGLU<BN_WIDE> eg{p.gcat, w.pw1_b, D, Mcat, 2 * D};
      PROF(PC_PW1_GLU);
      PPASR_CUDA_CHECK((gemm<BN_WIDE, ST_WIDE>(c, p.tm_ycat, m.pw1, Mcat, 2 * D, D, eg, st)));
    }
    if (!strided) {
      PROF(PC_DWCONV);
      PPASR_CUDA_CHECK(launch_dwconv_norm_swish(p.gcat, w.dw_w, w.dw_b, nullptr, w.cn_g, w.cn_b, cfg.conv_norm == 0, p.z, B, Tcat,
                                                Tc, D, K, 0, eps, nullptr, st));
    } else {
      const int Tout = (Tcat - K) / 2 + 1;  // == ceil(Tc / 2), the pooled residual's length (encoder.py:523-526)
      { PROF(PC_DWCONV);
        PPASR_CUDA_CHECK(launch_dwconv_stride(p.gcat, w.dw_w, w.dw_b, nullptr, w.cn_g, w.cn_b, cfg.conv_norm == 0, p.z, B, Tcat,
                                              Tout, D, K, 0, 2, eps, nullptr, st)); }
      PPASR_CUDA_CHECK(launch_avgpool2(xc, p.x2, B, Tc, Tout, D, st));
      PPASR_CUDA_CHECK(launch_halve_lens(p.vlen, p.vlen2, B, st));  // every row of a chunk is valid: ceil(Tp / 2) = Tout
      xc = p.x2, Tc = Tout, Mc = B * Tout, tmpos = &c->tm_pos2, rate = 2, t0 = t0 / 2;
    }
    // ---- x += pw2 z + b ; FFN ; norm_final ; y = next block's first LayerNorm (or after_norm)
    {
      const float* g2 = (l + 1 < L) ? c->layers[l + 1].ln_ffm_g : c->after_g;
      const float* b2 = (l + 1 < L) ? c->layers[l + 1].ln_ffm_b : c->after_b;
      PROF(PC_FUSED_FFN);
      PPASR_CUDA_CHECK(launch_fused_ffn(p.tm_z, &m.pw2, m.ff_w1_128, m.ff_w2s, Mc, FF, xc, p.y, w.ff_b1, w.ff_b2s, w.ln_fin_g,
                                        w.ln_fin_b, g2, b2, eps, w.pw2_b, w.ln_ff_g, w.ln_ff_b, nullptr, Tc, st));
    }
  }
  return PPASR_OK;
}
