// DeepSpeech2 (model_type 2) on top of the shared GEMM / CTC code (included by runtime.cu inside its anonymous namespace).
//
// Reference: ppasr/model_utils/deepspeech2/encoder.py:61-104 (CRNNEncoder.forward), conv.py:16-21, model.py:62-72.
// Per layer:  xg = y W_ih^T + b_ih  (tcgen05 GEMM, both directions in one N = nd*G*H launch)
//             -> rnn_layer_kernel (recurrence, sequence_length masking, final states)
//             -> row LayerNorm -> y (bf16 operand of the next layer / of the CTC head).

int finalize_ds2(ppasr_b200_ctx* c) {
  const auto& cfg = c->cfg;
  const int H = cfg.d_model, L = cfg.n_layers, V = cfg.vocab_size;
  auto& ds = c->ds;
  ds.nd = cfg.causal ? 1 : 2;
  ds.gru = cfg.use_gru ? 1 : 0;
  const int G = ds.gru ? 3 : 4, nd = ds.nd;
  ds.kin0 = 32 * c->F2;
  ds.kin0_pad = (ds.kin0 + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K * GEMM_BLOCK_K;
  std::string missing;
  std::vector<std::string> names = {"encoder.global_cmvn.mean", "encoder.global_cmvn.istd", "encoder.conv.conv.0.weight",
                                    "encoder.conv.conv.0.bias", "encoder.conv.conv.2.weight", "encoder.conv.conv.2.bias",
                                    "decoder.ctc_lo.weight", "decoder.ctc_lo.bias"};
  const char* sufs[2] = {"", "_reverse"};
  for (int l = 0; l < L; ++l) {
    const std::string p = "encoder.rnn." + std::to_string(l) + ".";
    for (int d = 0; d < nd; ++d)
      for (const char* n : {"weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"}) names.push_back(p + n + sufs[d]);
    names.push_back("encoder.layernorm_list." + std::to_string(l) + ".weight");
    names.push_back("encoder.layernorm_list." + std::to_string(l) + ".bias");
  }
  for (auto& n : names) find(c, n, &missing);
  if (!missing.empty()) {
    set_last_error("missing parameters: " + missing);
    return PPASR_ERR_STATE;
  }
  const int Kout = nd * H;
  size_t bytes = (size_t)L * ((size_t)nd * G * H * (size_t)std::max(ds.kin0_pad, Kout) + (size_t)nd * G * H * H) * 2;
  bytes += (size_t)c->Vpad * Kout * 2 + (size_t)L * (2 * nd * G * H + 2 * Kout) * 4 + (size_t)(c->Vpad + 16384) * 4;
  bytes += 4u << 20;
  PPASR_CUDA_CHECK(c->wslab.reserve(bytes));
  c->wslab.used = 0;
  std::string err;
  auto vecf = [&](const std::string& n) -> const float* { return upload(c, c->host[n].data); };
  c->cmvn_mean = vecf("encoder.global_cmvn.mean");
  c->cmvn_istd = vecf("encoder.global_cmvn.istd");
  {
    const HostTensor& w1 = c->host["encoder.conv.conv.0.weight"];
    const HostTensor& w2 = c->host["encoder.conv.conv.2.weight"];
    PPASR_REQUIRE(w1.numel() == 32 * 9 && w2.numel() == 32 * 32 * 9, "deepspeech2 conv weights must be [32,1,3,3] / [32,32,3,3]");
    c->conv1_w = upload(c, w1.data);
    c->conv1_b = vecf("encoder.conv.conv.0.bias");
    std::vector<float> t((size_t)288 * 32);  // [cin*9 + k][cout]
    for (int co = 0; co < 32; ++co)
      for (int ci = 0; ci < 32; ++ci)
        for (int k = 0; k < 9; ++k) t[(size_t)(ci * 9 + k) * 32 + co] = w2.data[((size_t)co * 32 + ci) * 9 + k];
    ds.conv2_wt = upload(c, t);
    c->conv2_b = vecf("encoder.conv.conv.2.bias");
  }
  {
    const HostTensor& w = c->host["decoder.ctc_lo.weight"];  // [nd*H, V]
    PPASR_REQUIRE(w.shape.size() == 2 && w.shape[0] == Kout && w.shape[1] == V, "decoder.ctc_lo.weight shape");
    c->ctc_w = upload(c, to_bf16(transpose_in_out(w, c->Vpad)));
    std::vector<float> b(c->Vpad, 0.f);
    std::memcpy(b.data(), c->host["decoder.ctc_lo.bias"].data.data(), sizeof(float) * V);
    c->ctc_b = upload(c, b);
    c->ctc_k = Kout;
  }
  ds.layers.resize(L);
  for (int l = 0; l < L; ++l) {
    auto& w = ds.layers[l];
    const std::string p = "encoder.rnn." + std::to_string(l) + ".";
    const int kin = l == 0 ? ds.kin0 : Kout, kpad = l == 0 ? ds.kin0_pad : Kout;
    std::vector<float> wih((size_t)nd * G * H * kpad, 0.f), bih((size_t)nd * G * H), whh((size_t)nd * G * H * H), bhh((size_t)nd * G * H);
    for (int d = 0; d < nd; ++d) {
      const HostTensor& a = c->host[p + "weight_ih_l0" + sufs[d]];  // [G*H, kin] (already K-major)
      const HostTensor& r = c->host[p + "weight_hh_l0" + sufs[d]];  // [G*H, H]
      PPASR_REQUIRE(a.numel() == (int64_t)G * H * kin && r.numel() == (int64_t)G * H * H, "rnn weight shape");
      for (int row = 0; row < G * H; ++row)
        std::memcpy(&wih[((size_t)d * G * H + row) * kpad], &a.data[(size_t)row * kin], sizeof(float) * kin);
      std::memcpy(&whh[(size_t)d * G * H * H], r.data.data(), sizeof(float) * G * H * H);
      std::memcpy(&bih[(size_t)d * G * H], c->host[p + "bias_ih_l0" + sufs[d]].data.data(), sizeof(float) * G * H);
      std::memcpy(&bhh[(size_t)d * G * H], c->host[p + "bias_hh_l0" + sufs[d]].data.data(), sizeof(float) * G * H);
    }
    w.wih = upload(c, to_bf16(wih));
    w.whh = upload(c, to_bf16(whh));
    w.bih = upload(c, bih);
    w.bhh = upload(c, bhh);
    w.ln_g = vecf("encoder.layernorm_list." + std::to_string(l) + ".weight");
    w.ln_b = vecf("encoder.layernorm_list." + std::to_string(l) + ".bias");
    if (!make_tmap_2d(&w.tm_wih, w.wih, kpad, (uint64_t)nd * G * H, (uint64_t)kpad * 2, BN_WIDE, &err)) {
      set_last_error(err);
      return PPASR_ERR_CUDA;
    }
  }
  if (!make_tmap_2d(&c->tm_ctc_w, c->ctc_w, Kout, c->Vpad, (uint64_t)Kout * 2, BN_NARROW, &err)) {
    set_last_error(err);
    return PPASR_ERR_CUDA;
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

// workspace for a (B, T) DeepSpeech2 problem; fills the Plan fields the ctc_* entry points read
int build_plan_ds2(ppasr_b200_ctx* c, int B, int T) {
  Plan& p = c->plan;
  if (p.B == B && p.T == T) return PPASR_OK;
  const auto& cfg = c->cfg;
  auto& ds = c->ds;
  const int H = cfg.d_model, F = cfg.feat_dim, nd = ds.nd, G = ds.gru ? 3 : 4, Kout = nd * H;
  PPASR_REQUIRE(T >= 7, "need at least 7 feature frames");
  PPASR_REQUIRE(B <= 64, "deepspeech2: at most 64 utterances per call in this build");
  Plan n;
  n.B = B, n.T = T;
  n.T1 = (T - 1) / 2;
  n.Tp = (n.T1 - 1) / 2;
  n.M = B * n.Tp;
  const size_t M = n.M;
  const int Ky = std::max(ds.kin0_pad, Kout);
  size_t bytes = 0;
  auto acc = [&](size_t b) { bytes += ((b + 255) & ~size_t(255)) + 256; };
  acc((size_t)B * T * F * 4);
  acc(B * 4);
  acc(M * Ky * 2);                      // y
  acc(M * (size_t)nd * G * H * 4);      // xg
  acc(M * Kout * 4);                    // rnn out
  acc((size_t)2 * nd * B * H * 4 + 64); // hbuf + barrier
  acc(M * c->Vld * 4);
  acc(M * (size_t)cfg.vocab_size * 4);
  acc(M * c->ctc_parts * 4 * 3);
  acc(M * 4 * 3 + (size_t)B * 4 * 3);
  bytes += 1u << 20;
  if (bytes > c->aslab.cap) {
    PPASR_CUDA_CHECK(cudaDeviceSynchronize());
    PPASR_CUDA_CHECK(c->aslab.reserve(bytes));
    PPASR_CUDA_CHECK(cudaMemset(c->aslab.base, 0, c->aslab.cap));
  }
  c->aslab.used = 0;
  auto& a = c->aslab;
  n.feats = a.take<float>((size_t)B * T * F);
  n.vlen = a.take<int>(B);
  n.y = a.take<__nv_bfloat16>(M * Ky);
  ds.xg = a.take<float>(M * (size_t)nd * G * H);
  ds.rnn_out = a.take<float>(M * Kout);
  ds.hbuf = a.take<float>((size_t)2 * nd * B * H);
  ds.bar = a.take<unsigned>(16);
  n.logits = a.take<float>(M * c->Vld);
  n.probs = a.take<float>(M * (size_t)cfg.vocab_size);
  n.pmax = a.take<float>(M * c->ctc_parts);
  n.psum = a.take<float>(M * c->ctc_parts);
  n.parg = a.take<int>(M * c->ctc_parts);
  n.idx = a.take<int>(M);
  n.maxp = a.take<float>(M);
  n.ids = a.take<int>(M);
  n.out_len = a.take<int>(B);
  n.score = a.take<float>(B);
  if (a.used > a.cap) {
    set_last_error("internal error: activation slab overflow");
    return PPASR_ERR_STATE;
  }
  std::string err;
  // layer 0 reads y as [M, kin0_pad]; later layers and the CTC head read it as [M, nd*H]
  if (!make_tmap_2d(&ds.tm_y0, n.y, ds.kin0_pad, M, (uint64_t)ds.kin0_pad * 2, GEMM_BLOCK_M, &err) ||
      !make_tmap_2d(&n.tm_y, n.y, Kout, M, (uint64_t)Kout * 2, GEMM_BLOCK_M, &err)) {
    set_last_error(err);
    return PPASR_ERR_CUDA;
  }
  n.Mc = n.M, n.Tc = n.Tp, n.vc = n.vlen;
  p = n;
  return PPASR_OK;
}

// keep_state: use / update the streaming states in ds.h_state / ds.c_state (predict_chunk_deepspeech,
// inference_predictor.py:147-182); otherwise zero initial states (get_encoder_out, model.py:62-65)
int run_encoder_ds2(ppasr_b200_ctx* c, cudaStream_t st, bool keep_state) {
  Plan& p = c->plan;
  const auto& cfg = c->cfg;
  auto& ds = c->ds;
  const int H = cfg.d_model, L = cfg.n_layers, nd = ds.nd, G = ds.gru ? 3 : 4, Kout = nd * H, M = p.M;
  { PROF(PC_CONV1);
    PPASR_CUDA_CHECK(launch_ds2_conv(p.feats, c->cmvn_mean, c->cmvn_istd, c->conv1_w, c->conv1_b, ds.conv2_wt, c->conv2_b, p.y,
                                     p.B, p.T, cfg.feat_dim, p.Tp, ds.kin0_pad, st)); }
  for (int l = 0; l < L; ++l) {
    const auto& w = ds.layers[l];
    const int kpad = l == 0 ? ds.kin0_pad : Kout;
    {
      EpiResidF32<BN_WIDE> e{ds.xg, w.bih, nd * G * H, M, nd * G * H, 1.0f, 0, nullptr, p.Tp};
      PROF(PC_FFN1);
      PPASR_CUDA_CHECK((gemm<BN_WIDE, ST_WIDE>(c, l == 0 ? ds.tm_y0 : p.tm_y, w.tm_wih, M, nd * G * H, kpad, e, st)));
    }
    {
      float* hs = keep_state ? ds.h_state + (size_t)l * nd * p.B * H : nullptr;
      float* cs = (keep_state && !ds.gru) ? ds.c_state + (size_t)l * nd * p.B * H : nullptr;
      PROF(PC_ATTENTION);
      PPASR_CUDA_CHECK(launch_rnn_layer(ds.xg, w.whh, w.bhh, p.vlen, hs, cs, hs, cs, ds.rnn_out, ds.hbuf, ds.bar, p.B, p.Tp, H, nd,
                                        ds.gru, c->sms, st));
    }
    { PROF(PC_LAYERNORM); PPASR_CUDA_CHECK(launch_row_layernorm(ds.rnn_out, w.ln_g, w.ln_b, p.y, M, Kout, 1e-5f, st)); }
  }
  return PPASR_OK;
}
