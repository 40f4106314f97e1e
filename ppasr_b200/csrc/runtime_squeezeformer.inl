// Squeezeformer encoder on the same kernels as the Conformer (included by runtime.cu inside its anonymous namespace).
//
// Reference: ppasr/model_utils/squeezeformer/encoder.py:172-236 (forward), :435-506 (post-norm block
// MHA -> LN -> FFN -> LN -> Conv -> LN -> FFN -> LN), attention.py:96-162, convolution.py:102-163, positionwise.py:55-65,
// subsampling.py:53-68, time_reduction.py:61-84,183-206.
//
// Mapping onto kernels (x = fp32 residual stream, y = bf16 operand of the next GEMM):
//  * every "x = LN(x + module(ada(x)))" ends in an epilogue that writes x = LN(...) and y = ada_next * x + ada_bias_next
//    (EpiResidLN / fused_ffn with y_affine), so the adaptive scale/bias never needs its own pass;
//  * subsampling: same conv1 / conv2-as-9-tap-GEMM kernels; pos_enc's x * sqrt(D) sits before input_proj, so the
//    projection runs with alpha = 16 and bias / 16, fused with preln;
//  * time reduction: masked depthwise k1/k5 stride-2 kernel -> pointwise GEMM (no norm) into the reduced stream x2;
//    blocks reduce_idx..recover_idx-1 run on B * ceil(T'/2) compact rows with every second row of the positional table;
//  * recover: one GEMM on the reduced rows whose epilogue adds its row to the two full-rate rows it is repeated to.

int finalize_squeezeformer(ppasr_b200_ctx* c) {
  const auto& cfg = c->cfg;
  const int D = cfg.d_model, L = cfg.n_layers, FF = cfg.ffn_dim, K = cfg.conv_kernel, V = cfg.vocab_size;
  auto& sq = c->sq;
  sq.reduce_idx = cfg.reduce_idx, sq.recover_idx = cfg.recover_idx, sq.tr_k = cfg.time_reduce_kernel;
  std::string missing;
  std::vector<std::string> names = {"encoder.global_cmvn.mean", "encoder.global_cmvn.istd", "encoder.embed.pw_conv.weight",
                                    "encoder.embed.pw_conv.bias", "encoder.embed.dw_conv.weight", "encoder.embed.dw_conv.bias",
                                    "encoder.embed.input_proj.0.weight", "encoder.embed.input_proj.0.bias",
                                    "encoder.preln.weight", "encoder.preln.bias", "ctc.ctc_lo.weight", "ctc.ctc_lo.bias"};
  if (sq.reduce_idx >= 0)
    for (const char* s : {"time_reduction_layer.dw_conv.weight", "time_reduction_layer.dw_conv.bias",
                          "time_reduction_layer.pw_conv.weight", "time_reduction_layer.pw_conv.bias",
                          "time_recover_layer.weight", "time_recover_layer.bias"})
      names.push_back(std::string("encoder.") + s);
  const char* mods[4] = {"self_attn", "ffn1", "conv_module", "ffn2"};
  for (int l = 0; l < L; ++l) {
    const std::string p = "encoder.encoders." + std::to_string(l) + ".";
    for (int i = 1; i <= 4; ++i) {
      names.push_back(p + "layer_norm" + std::to_string(i) + ".weight");
      names.push_back(p + "layer_norm" + std::to_string(i) + ".bias");
    }
    for (const char* s : {"ffn1", "ffn2"})
      for (const char* t : {".w_1.weight", ".w_1.bias", ".w_2.weight", ".w_2.bias"}) names.push_back(p + s + t);
    for (const char* s : {"linear_q", "linear_k", "linear_v", "linear_out", "linear_pos"}) {
      names.push_back(p + "self_attn." + s + ".weight");
      names.push_back(p + "self_attn." + s + ".bias");
    }
    names.push_back(p + "self_attn.pos_bias_u");
    names.push_back(p + "self_attn.pos_bias_v");
    for (const char* s : {"pointwise_conv1", "depthwise_conv", "pointwise_conv2", "norm"}) {
      names.push_back(p + "conv_module." + s + ".weight");
      names.push_back(p + "conv_module." + s + ".bias");
    }
    if (cfg.conv_norm == 1) {
      names.push_back(p + "conv_module.norm._mean");
      names.push_back(p + "conv_module.norm._variance");
    }
  }
  for (auto& n : names) find(c, n, &missing);
  if (!missing.empty()) {
    set_last_error("missing parameters: " + missing);
    return PPASR_ERR_STATE;
  }

  size_t bytes = 0;
  bytes += (size_t)L * ((size_t)4 * D * FF + 3 * D * D + D * D + 2 * D * D + D * D) * 2;
  bytes += (size_t)D * 9 * D * 2 + (size_t)D * c->Kemb * 2 + (size_t)c->Vpad * D * 2 + (size_t)2 * D * D * 2;
  bytes += (size_t)cfg.max_len * L * D * 2 + (size_t)cfg.max_len * D * 2 + (size_t)L * D * D * 2;
  bytes += (size_t)L * (40 * D + 2 * FF + D * K + 64) * 4 + (size_t)(c->Vpad + 16 * D + 2 * L * D + 4096) * 4;
  bytes += 4u << 20;
  PPASR_CUDA_CHECK(c->wslab.reserve(bytes));
  c->wslab.used = 0;

  std::string err;
  auto vecf = [&](const std::string& n) -> const float* { return upload(c, c->host[n].data); };
  {
    std::vector<float> one(D, 1.f), zero((size_t)std::max(L * D, 4096), 0.f);
    sq.ones = upload(c, one);
    sq.zeros = upload(c, zero);
    c->zero_bias = sq.zeros;
  }
  // adaptive scale / bias [1,1,D]; absent (adaptive_scale: false) => identity
  auto ada = [&](const std::string& mod, const float** s, const float** b) {
    auto is = c->host.find(mod + ".ada_scale");
    auto ib = c->host.find(mod + ".ada_bias");
    *s = (is != c->host.end() && is->second.numel() == D) ? upload(c, is->second.data) : sq.ones;
    *b = (ib != c->host.end() && ib->second.numel() == D) ? upload(c, ib->second.data) : sq.zeros;
  };

  // ---- front end (subsampling.py:33-44): pw_conv = Conv2d(1->D), dw_conv = full Conv2d(D->D) since dw_stride is False ----
  c->cmvn_mean = vecf("encoder.global_cmvn.mean");
  c->cmvn_istd = vecf("encoder.global_cmvn.istd");
  {
    const HostTensor& w = c->host["encoder.embed.pw_conv.weight"];
    PPASR_REQUIRE(w.numel() == (int64_t)D * 9, "embed.pw_conv.weight shape");
    c->conv1_w = upload(c, w.data);
    c->conv1_b = vecf("encoder.embed.pw_conv.bias");
  }
  {
    const HostTensor& w = c->host["encoder.embed.dw_conv.weight"];
    PPASR_REQUIRE(w.numel() == (int64_t)D * D * 9, "embed.dw_conv.weight shape (depthwise dw_stride=True is not supported)");
    std::vector<float> p((size_t)D * 9 * D);
    for (int o = 0; o < D; ++o)
      for (int i = 0; i < D; ++i)
        for (int t = 0; t < 9; ++t) p[(size_t)o * 9 * D + (size_t)t * D + i] = w.data[((size_t)o * D + i) * 9 + t];
    c->conv2_w = upload(c, to_bf16(p));
    c->conv2_b = vecf("encoder.embed.dw_conv.bias");
  }
  {
    const HostTensor& w = c->host["encoder.embed.input_proj.0.weight"];
    PPASR_REQUIRE(w.shape.size() == 2 && w.shape[0] == (int64_t)D * c->F2 && w.shape[1] == D, "embed.input_proj.0.weight shape");
    std::vector<float> p((size_t)D * c->Kemb);
    for (int ch = 0; ch < D; ++ch)
      for (int f = 0; f < c->F2; ++f)
        for (int o = 0; o < D; ++o) p[(size_t)o * c->Kemb + (size_t)f * D + ch] = w.data[((size_t)ch * c->F2 + f) * D + o];
    c->emb_w = upload(c, to_bf16(p));
    // x * sqrt(D) precedes the projection (subsampling.py:66-67): Linear(16 x) = 16 * (x W + b / 16); exact for D = 256
    std::vector<float> b = c->host["encoder.embed.input_proj.0.bias"].data;
    const float xs = std::sqrt((float)D);
    for (auto& v : b) v /= xs;
    sq.emb_b_scaled = upload(c, b);
  }
  sq.preln_g = vecf("encoder.preln.weight");
  sq.preln_b = vecf("encoder.preln.bias");
  {
    const HostTensor& w = c->host["ctc.ctc_lo.weight"];
    PPASR_REQUIRE(w.shape.size() == 2 && w.shape[0] == D && w.shape[1] == V, "ctc_lo.weight shape");
    c->ctc_w = upload(c, to_bf16(transpose_in_out(w, c->Vpad)));
    std::vector<float> b(c->Vpad, 0.f);
    std::memcpy(b.data(), c->host["ctc.ctc_lo.bias"].data.data(), sizeof(float) * V);
    c->ctc_b = upload(c, b);
  }
  if (sq.reduce_idx >= 0) {
    const HostTensor& dw = c->host["encoder.time_reduction_layer.dw_conv.weight"];  // [D,1,k]
    PPASR_REQUIRE(dw.numel() == (int64_t)D * sq.tr_k, "time_reduction_layer.dw_conv.weight shape");
    sq.tr_dw_w = upload(c, dw.data);
    sq.tr_dw_b = vecf("encoder.time_reduction_layer.dw_conv.bias");
    const HostTensor& pw = c->host["encoder.time_reduction_layer.pw_conv.weight"];  // [out, in, 1]
    PPASR_REQUIRE(pw.numel() == (int64_t)D * D, "time_reduction_layer.pw_conv.weight shape");
    sq.tr_pw = upload(c, to_bf16(pw.data));
    sq.tr_pw_b = vecf("encoder.time_reduction_layer.pw_conv.bias");
    sq.rec_w = upload(c, to_bf16(transpose_in_out(c->host["encoder.time_recover_layer.weight"])));
    sq.rec_b = vecf("encoder.time_recover_layer.bias");
    if (!make_tmap_2d(&sq.tm_tr_pw, sq.tr_pw, D, D, (uint64_t)D * 2, BN_WIDE, &err) ||
        !make_tmap_2d(&sq.tm_rec_w, sq.rec_w, D, D, (uint64_t)D * 2, BN_WIDE, &err)) {
      set_last_error(err);
      return PPASR_ERR_CUDA;
    }
  }

  // ---- blocks ----------------------------------------------------------------------------------
  sq.layers.resize(L);
  sq.maps.resize(L);
  std::vector<float> wpos_all((size_t)L * D * D), bpos_all((size_t)L * D);
  for (int l = 0; l < L; ++l) {
    const std::string p = "encoder.encoders." + std::to_string(l) + ".";
    auto& w = sq.layers[l];
    for (int i = 0; i < 4; ++i) {
      w.ln_g[i] = vecf(p + "layer_norm" + std::to_string(i + 1) + ".weight");
      w.ln_b[i] = vecf(p + "layer_norm" + std::to_string(i + 1) + ".bias");
      ada(p + mods[i], &w.ada_s[i], &w.ada_b[i]);
    }
    for (int f = 0; f < 2; ++f) {
      const std::string q = p + (f == 0 ? "ffn1" : "ffn2");
      w.w1[f] = upload(c, to_bf16(transpose_in_out(c->host[q + ".w_1.weight"])));
      w.w2[f] = upload(c, to_bf16(transpose_in_out(c->host[q + ".w_2.weight"])));
      w.b1[f] = vecf(q + ".w_1.bias");
      w.b2[f] = vecf(q + ".w_2.bias");
    }
    {
      std::vector<float> qkv((size_t)3 * D * D), bq((size_t)3 * D);
      const char* nm[3] = {"linear_q", "linear_k", "linear_v"};
      for (int s = 0; s < 3; ++s) {
        auto t = transpose_in_out(c->host[p + "self_attn." + nm[s] + ".weight"]);
        std::memcpy(qkv.data() + (size_t)s * D * D, t.data(), sizeof(float) * D * D);
        std::memcpy(bq.data() + (size_t)s * D, c->host[p + "self_attn." + nm[s] + ".bias"].data.data(), sizeof(float) * D);
      }
      w.wqkv = upload(c, to_bf16(qkv));
      w.bqkv = upload(c, bq);
    }
    w.wo = upload(c, to_bf16(transpose_in_out(c->host[p + "self_attn.linear_out.weight"])));
    w.bo = vecf(p + "self_attn.linear_out.bias");
    w.pos_u = vecf(p + "self_attn.pos_bias_u");
    w.pos_v = vecf(p + "self_attn.pos_bias_v");
    {
      auto t = transpose_in_out(c->host[p + "self_attn.linear_pos.weight"]);
      std::memcpy(wpos_all.data() + (size_t)l * D * D, t.data(), sizeof(float) * D * D);
      std::memcpy(bpos_all.data() + (size_t)l * D, c->host[p + "self_attn.linear_pos.bias"].data.data(), sizeof(float) * D);
    }
    {
      const HostTensor& pw = c->host[p + "conv_module.pointwise_conv1.weight"];
      const HostTensor& pb = c->host[p + "conv_module.pointwise_conv1.bias"];
      PPASR_REQUIRE(pw.numel() == (int64_t)2 * D * D, "pointwise_conv1.weight shape");
      std::vector<float> wi((size_t)2 * D * D), bi((size_t)2 * D);
      for (int ch = 0; ch < D; ++ch) {
        std::memcpy(&wi[(size_t)(2 * ch) * D], &pw.data[(size_t)ch * D], sizeof(float) * D);
        std::memcpy(&wi[(size_t)(2 * ch + 1) * D], &pw.data[(size_t)(ch + D) * D], sizeof(float) * D);
        bi[2 * ch] = pb.data[ch];
        bi[2 * ch + 1] = pb.data[ch + D];
      }
      w.pw1 = upload(c, to_bf16(wi));
      w.pw1_b = upload(c, bi);
      float* pad = c->wslab.take<float>(D);
      PPASR_CUDA_CHECK(launch_glu_pad(w.pw1_b, pad, D, 0));
      w.glu_pad = pad;
    }
    {
      const HostTensor& dw = c->host[p + "conv_module.depthwise_conv.weight"];
      PPASR_REQUIRE(dw.numel() == (int64_t)D * K, "depthwise_conv.weight shape");
      w.dw_w = upload(c, dw.data);
      w.dw_b = vecf(p + "conv_module.depthwise_conv.bias");
    }
    if (cfg.conv_norm == 0) {
      w.cn_g = vecf(p + "conv_module.norm.weight");
      w.cn_b = vecf(p + "conv_module.norm.bias");
    } else {
      const auto& g = c->host[p + "conv_module.norm.weight"].data;
      const auto& b = c->host[p + "conv_module.norm.bias"].data;
      const auto& mu = c->host[p + "conv_module.norm._mean"].data;
      const auto& var = c->host[p + "conv_module.norm._variance"].data;
      std::vector<float> sc(D), sh(D);
      for (int i = 0; i < D; ++i) {
        sc[i] = g[i] / std::sqrt(var[i] + 1e-5f);
        sh[i] = b[i] - mu[i] * sc[i];
      }
      w.cn_g = upload(c, sc);
      w.cn_b = upload(c, sh);
    }
    {
      const HostTensor& pw = c->host[p + "conv_module.pointwise_conv2.weight"];
      PPASR_REQUIRE(pw.numel() == (int64_t)D * D, "pointwise_conv2.weight shape");
      w.pw2 = upload(c, to_bf16(pw.data));
      w.pw2_b = vecf(p + "conv_module.pointwise_conv2.bias");
    }
    auto& m = sq.maps[l];
    bool ok = make_tmap_2d(&m.w1_128[0], w.w1[0], D, FF, (uint64_t)D * 2, 128, &err) &&
              make_tmap_2d(&m.w1_128[1], w.w1[1], D, FF, (uint64_t)D * 2, 128, &err) &&
              make_tmap_2d(&m.w2[0], w.w2[0], FF, D, (uint64_t)FF * 2, BN_WIDE, &err) &&
              make_tmap_2d(&m.w2[1], w.w2[1], FF, D, (uint64_t)FF * 2, BN_WIDE, &err) &&
              make_tmap_2d(&m.wqkv, w.wqkv, D, 3 * D, (uint64_t)D * 2, BN_NARROW, &err) &&
              make_tmap_2d(&m.wo, w.wo, D, D, (uint64_t)D * 2, BN_WIDE, &err) &&
              make_tmap_2d(&m.pw1, w.pw1, D, 2 * D, (uint64_t)D * 2, BN_WIDE, &err) &&
              make_tmap_2d(&m.pw2, w.pw2, D, D, (uint64_t)D * 2, BN_WIDE, &err);
    if (!ok) {
      set_last_error(err);
      return PPASR_ERR_CUDA;
    }
  }
  if (!make_tmap_2d(&c->tm_conv2_w, c->conv2_w, (uint64_t)9 * D, D, (uint64_t)9 * D * 2, BN_WIDE, &err) ||
      !make_tmap_2d(&c->tm_emb_w, c->emb_w, c->Kemb, D, (uint64_t)c->Kemb * 2, BN_WIDE, &err) ||
      !make_tmap_2d(&c->tm_ctc_w, c->ctc_w, D, c->Vpad, (uint64_t)D * 2, BN_NARROW, &err)) {
    set_last_error(err);
    return PPASR_ERR_CUDA;
  }
  // ---- pos_tab[pos, l*D + n] = linear_pos_l(pe[pos]) + bias_l (attention.py:28,142-144: WITH bias) ----
  {
    const int ML = cfg.max_len;
    std::vector<float> pe((size_t)ML * D);
    for (int pos = 0; pos < ML; ++pos)
      for (int i = 0; i < D / 2; ++i) {
        const float div = std::exp((float)(2 * i) * -(std::log(10000.0f) / (float)D));
        pe[(size_t)pos * D + 2 * i] = std::sin((float)pos * div);
        pe[(size_t)pos * D + 2 * i + 1] = std::cos((float)pos * div);
      }
    const __nv_bfloat16* pe_d = upload(c, to_bf16(pe));
    const __nv_bfloat16* wpos_d = upload(c, to_bf16(wpos_all));
    const float* bpos_d = upload(c, bpos_all);
    __nv_bfloat16* tab = c->wslab.take<__nv_bfloat16>((size_t)ML * L * D);
    CUtensorMap ta, tb;
    if (!make_tmap_2d(&ta, pe_d, D, ML, (uint64_t)D * 2, GEMM_BLOCK_M, &err) ||
        !make_tmap_2d(&tb, wpos_d, D, (uint64_t)L * D, (uint64_t)D * 2, BN_WIDE, &err) ||
        !make_tmap_2d(&c->tm_pos, tab, (uint64_t)L * D, ML, (uint64_t)L * D * 2, 128, &err) ||
        !make_tmap_2d(&sq.tm_pos2, tab, (uint64_t)L * D, ML / 2, (uint64_t)2 * L * D * 2, 128, &err)) {
      set_last_error(err);
      return PPASR_ERR_CUDA;
    }
    GemmShape s = make_shape(ML, L * D, D, BN_WIDE);
    EpiStoreBF16<BN_WIDE, ACT_NONE> epi{tab, bpos_d, L * D, ML, L * D};
    PPASR_CUDA_CHECK((launch_gemm<BN_WIDE, ST_WIDE, false>(ta, tb, s, epi, c->sms, 0)));
    c->pos_tab = tab;
  }
  PPASR_CUDA_CHECK(cudaDeviceSynchronize());
  if (c->wslab.used > c->wslab.cap) {
    set_last_error("internal error: weight slab overflow");
    return PPASR_ERR_STATE;
  }
  c->host.clear();
  c->finalized = true;
  return PPASR_OK;
}

int run_encoder_squeezeformer(ppasr_b200_ctx* c, cudaStream_t st, bool chunk = false) {
  Plan& p = c->plan;
  const auto& cfg = c->cfg;
  auto& sq = c->sq;
  const int D = cfg.d_model, H = cfg.n_heads, FF = cfg.ffn_dim, L = cfg.n_layers, K = cfg.conv_kernel;
  const float eps = 1e-5f;
  int rc = run_subsampling_convs(c, st);
  if (rc) return rc;
  // current resolution of the residual stream
  // chunk mode (forward_chunk, encoder.py:260-383): no padding masks; K/V and the conv-module inputs of earlier chunks come
  // from the device-resident caches in c->ss. The half-rate blocks keep their cache at half rate (the reference stores every
  // key twice and reads every second one, encoder.py:351,368): positions / lengths are halved for them, which needs even
  // chunk and cache sizes (checked by the caller).
  auto& ss = c->ss;
  float* xc = p.x;
  int Tc = p.Tp, Mc = p.M;
  int rate = 1;  // 2 inside the time-reduced section
  const int* vl = chunk ? nullptr : p.vlen;
  const CUtensorMap* tmpos = &c->tm_pos;
  const int Tr = (p.Tp + 1) / 2, Mr = p.B * Tr;

  // y-affine of the epilogue that feeds module `mod` of block l (mod 4 = first module of block l + 1)
  auto feed = [&](int l, int mod, const float** s, const float** b) {
    if (mod == 4) ++l, mod = 0;
    const bool plain = (l >= L) || (mod == 0 && (l == sq.recover_idx || l == sq.reduce_idx));
    *s = plain ? sq.ones : sq.layers[l].ada_s[mod];
    *b = plain ? sq.zeros : sq.layers[l].ada_b[mod];
  };
  auto post_ln = [&](int cls, const CUtensorMap& ta, const CUtensorMap& tb, int Kdim, const float* bias, float alpha, int residual,
                     const int* lens, int mask_resid, const float* g, const float* b, const float* ys, const float* yb,
                     int no_norm) -> int {
    EpiResidLN<BN_WIDE> e{xc, bias, D, Mc, D, alpha, residual, lens, Tc, mask_resid, 0, g, b, ys, yb, p.y, eps, 1, no_norm};
    PROF(cls);
    PPASR_CUDA_CHECK((gemm<BN_WIDE, ST_WIDE>(c, ta, tb, Mc, D, Kdim, e, st)));
    return PPASR_OK;
  };
  const float *ys, *yb;
  // input_proj(16 * conv) -> preln; y = ada_mha(0)                       (subsampling.py:66-67, encoder.py:205)
  feed(0, 0, &ys, &yb);
  if ((rc = post_ln(PC_EMBED, p.tm_c2, c->tm_emb_w, c->Kemb, sq.emb_b_scaled, std::sqrt((float)D), 0, nullptr, 0, sq.preln_g,
                    sq.preln_b, ys, yb, 0)))
    return rc;
  for (int l = 0; l < L; ++l) {
    const auto& w = sq.layers[l];
    const auto& m = sq.maps[l];
    if (l == sq.reduce_idx) {
      // ---- time reduction (encoder.py:211-214): masked depthwise stride-2 conv -> pointwise conv; masks / pos_emb [::2]
      if (!chunk) PPASR_CUDA_CHECK(launch_halve_lens(p.vlen, p.vlen2, p.B, st));
      { PROF(PC_DWCONV);
        PPASR_CUDA_CHECK(launch_time_reduce_dw(p.x, sq.tr_dw_w, sq.tr_dw_b, vl, p.z, p.B, p.Tp, Tr, D, sq.tr_k,
                                               sq.tr_k > 2 ? sq.tr_k - 2 : 0, st)); }
      xc = p.x2, Tc = Tr, Mc = Mr, vl = chunk ? nullptr : p.vlen2, tmpos = &sq.tm_pos2, rate = 2;
      ys = w.ada_s[0], yb = w.ada_b[0];
      if ((rc = post_ln(PC_PW2, p.tm_z, sq.tm_tr_pw, D, sq.tr_pw_b, 1.0f, 0, nullptr, 0, sq.ones, sq.zeros, ys, yb, 1))) return rc;
    }
    if (l == sq.recover_idx) {
      // ---- recover (encoder.py:216-230): x_full += Linear(repeat_interleave(x_reduced, 2))[:T']; y = ada_mha(l)(x_full)
      EpiRecover<BN_WIDE> e{p.x, p.y, sq.rec_b, w.ada_s[0], w.ada_b[0], Mr, D, Tr, p.Tp};
      { PROF(PC_OUTPROJ); PPASR_CUDA_CHECK((gemm<BN_WIDE, ST_WIDE>(c, p.tm_y, sq.tm_rec_w, Mr, D, D, e, st))); }
      xc = p.x, Tc = p.Tp, Mc = p.M, vl = chunk ? nullptr : p.vlen, tmpos = &c->tm_pos, rate = 1;
    }
    // ---- MHA: x = LN1(x + Wo attn(ada x)); y = ada_ffn1(x)                 (encoder.py:468-478)
    {
      AttnParams ap{};
      ap.B = p.B, ap.H = H, ap.T1 = Tc, ap.D = D, ap.pos_col0 = l * D, ap.out = p.att, ap.q_rows_per_bh = Tc;
      if (!chunk) {
        EpiQKV<BN_NARROW> e{p.q2, p.kk, p.vt, w.bqkv, w.pos_u, w.pos_v, Mc, Tc, H, Tc, p.Tkp, 0};
        { PROF(PC_QKV); PPASR_CUDA_CHECK((gemm<BN_NARROW, ST_NARROW>(c, p.tm_y, m.wqkv, Mc, 3 * D, D, e, st))); }
        ap.T2 = Tc, ap.k_rows_per_bh = Tc, ap.k_row0 = 0, ap.pos_row0 = 0, ap.klens = vl;
        { PROF(PC_ATTENTION); PPASR_CUDA_CHECK(launch_rel_attention(p.tm_q, p.tm_k, *tmpos, p.tm_vt, ap, st)); }
      } else {
        // this chunk's K/V are appended to the block's cache at its own rate; keys kstart .. kend + chunk are attended with
        // positions offset - cache_t .. (attention.py:127-134, encoder.py:302,351)
        const size_t lk = (size_t)l * ss.B * H * ss.Tcap * 64;
        const int kend = ss.kend / rate, kstart = ss.kstart / rate, cache_t = kend - kstart;
        EpiQKV<BN_NARROW> e{p.q2, ss.kk + lk, ss.vt + lk, w.bqkv, w.pos_u, w.pos_v, Mc, Tc, H, ss.Tcap, ss.Tcap, kend};
        { PROF(PC_QKV); PPASR_CUDA_CHECK((gemm<BN_NARROW, ST_NARROW>(c, p.tm_y, m.wqkv, Mc, 3 * D, D, e, st))); }
        ap.T2 = cache_t + Tc, ap.k_rows_per_bh = ss.Tcap, ap.k_row0 = kstart, ap.pos_row0 = (ss.offset - (ss.kend - ss.kstart)) / rate;
        ap.klens = nullptr;
        { PROF(PC_ATTENTION); PPASR_CUDA_CHECK(launch_rel_attention(p.tm_q, ss.tm_k[l], *tmpos, ss.tm_vt[l], ap, st)); }
      }
      if (!c->fused_ffn) {
        feed(l, 1, &ys, &yb);
        if ((rc = post_ln(PC_OUTPROJ, p.tm_att, m.wo, D, w.bo, 1.0f, 1, nullptr, 0, w.ln_g[0], w.ln_b[0], ys, yb, 0))) return rc;
      }
    }
    // ---- FFN1: x = LN2(x + W2 swish(W1 y)); y = ada_conv(x), pad frames zeroed  (encoder.py:480-487, convolution.py:119-127)
    //      fused: the attention out-projection + residual + layer_norm1 run in the kernel's chained pre-GEMM slot
    {
      feed(l, 2, &ys, &yb);
      PROF(PC_FUSED_FFN);
      if (c->fused_ffn) {
        PPASR_CUDA_CHECK(launch_fused_ffn(p.tm_att, &m.wo, m.w1_128[0], m.w2[0], Mc, FF, xc, p.y, w.b1[0], w.b2[0], w.ln_g[1],
                                          w.ln_b[1], ys, yb, eps, w.bo, w.ln_g[0], w.ln_b[0], nullptr, Tc, st, 1, vl,
                                          w.ada_s[1], w.ada_b[1]));
      } else {
        PPASR_CUDA_CHECK(launch_fused_ffn(p.tm_y, nullptr, m.w1_128[0], m.w2[0], Mc, FF, xc, p.y, w.b1[0], w.b2[0], w.ln_g[1],
                                          w.ln_b[1], ys, yb, eps, nullptr, nullptr, nullptr, nullptr, Tc, st, 1, vl));
      }
    }
    // ---- conv module: x = LN3(x + mask * pw2 swish(norm(dw(glu(pw1 y))))); y = ada_ffn2(x)   (encoder.py:489-496)
    {
      if (!chunk) {
        EpiGLU<BN_WIDE> eg{p.g, w.pw1_b, D, Mc, 2 * D};
        { PROF(PC_PW1_GLU); PPASR_CUDA_CHECK((gemm<BN_WIDE, ST_WIDE>(c, p.tm_y, m.pw1, Mc, 2 * D, D, eg, st))); }
        const int lpad = cfg.causal ? K - 1 : (K - 1) / 2;
        if (!(cfg.causal && c->fused_ffn && c->fused_dwconv)) {
          PROF(PC_DWCONV);
          PPASR_CUDA_CHECK(launch_dwconv_norm_swish(p.g, w.dw_w, w.dw_b, cfg.causal ? w.glu_pad : nullptr, w.cn_g, w.cn_b,
                                                    cfg.conv_norm == 0, p.z, p.B, Tc, Tc, D, K, lpad, eps, vl, st));
        }
      } else {
        // [cnn_cache ; chunk] -> pw1 + GLU -> "valid" depthwise conv; cache <- last K-1 input rows (convolution.py:129-138)
        const int lorder = K - 1, Tcat = Tc + lorder, Mcat = p.B * Tcat;
        PPASR_CUDA_CHECK(launch_conv_cache_concat(ss.cnn + (size_t)l * ss.B * lorder * D, p.y, p.ycat, p.B, Tc, lorder, D, st));
        EpiGLU<BN_WIDE> eg{p.gcat, w.pw1_b, D, Mcat, 2 * D};
        { PROF(PC_PW1_GLU); PPASR_CUDA_CHECK((gemm<BN_WIDE, ST_WIDE>(c, p.tm_ycat, m.pw1, Mcat, 2 * D, D, eg, st))); }
        { PROF(PC_DWCONV);
          PPASR_CUDA_CHECK(launch_dwconv_norm_swish(p.gcat, w.dw_w, w.dw_b, nullptr, w.cn_g, w.cn_b, cfg.conv_norm == 0, p.z, p.B,
                                                    Tcat, Tc, D, K, 0, eps, nullptr, st)); }
      }
      if (!c->fused_ffn) {
        feed(l, 3, &ys, &yb);
        if ((rc = post_ln(PC_PW2, p.tm_z, m.pw2, D, w.pw2_b, 1.0f, 1, vl, 1, w.ln_g[2], w.ln_b[2], ys, yb, 0))) return rc;
      }
    }
    // ---- FFN2: x = LN4(x + W2 swish(W1 y)); y = input of whatever comes next       (encoder.py:498-506)
    {
      feed(l, 4, &ys, &yb);
      PROF(PC_FUSED_FFN);
      if (c->fused_ffn) {
        // pointwise_conv2 (pad rows masked) + residual + layer_norm3 in the chained pre-GEMM slot
        FfnDw dw{p.g, w.dw_w, w.dw_b, w.glu_pad, w.cn_g, w.cn_b, K, cfg.conv_norm == 0};
        const bool fdw = !chunk && cfg.causal && c->fused_dwconv;
        PPASR_CUDA_CHECK(launch_fused_ffn(p.tm_z, &m.pw2, m.w1_128[1], m.w2[1], Mc, FF, xc, p.y, w.b1[1], w.b2[1], w.ln_g[3],
                                          w.ln_b[3], ys, yb, eps, w.pw2_b, w.ln_g[2], w.ln_b[2], vl, Tc, st, 1, nullptr,
                                          w.ada_s[3], w.ada_b[3], fdw ? &dw : nullptr));
      } else {
        PPASR_CUDA_CHECK(launch_fused_ffn(p.tm_y, nullptr, m.w1_128[1], m.w2[1], Mc, FF, xc, p.y, w.b1[1], w.b2[1], w.ln_g[3],
                                          w.ln_b[3], ys, yb, eps, nullptr, nullptr, nullptr, nullptr, Tc, st, 1, nullptr));
      }
    }
  }
  return PPASR_OK;
}
