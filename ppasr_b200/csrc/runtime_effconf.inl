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
