/* Minimal C client of the ppasr_b200 C-ABI (include/ppasr_b200.h): what a non-Python host (or the cgo / JNI / N-API stub
 * of another runtime) links against. Build:
 *   gcc -std=c99 -Iinclude examples/c_abi_demo.c -Lppasr_b200/lib -lppasr_b200 -Wl,-rpath,$PWD/ppasr_b200/lib -o c_abi_demo
 * Without weights it only exercises the life cycle and the error path (finalize reports the missing parameters). */
#include <stdio.h>
#include <string.h>

#include "ppasr_b200.h"

int main(void) {
  ppasr_b200_config cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.model_type = 0; /* conformer.yml */
  cfg.feat_dim = 80;
  cfg.d_model = 256;
  cfg.n_heads = 4;
  cfg.ffn_dim = 2048;
  cfg.n_layers = 12;
  cfg.conv_kernel = 15;
  cfg.causal = 1;
  cfg.vocab_size = 4233;
  cfg.max_len = 5000;
  ppasr_b200_ctx* ctx = NULL;
  if (ppasr_b200_create(&cfg, &ctx) != 0) {
    fprintf(stderr, "create failed: %s\n", ppasr_b200_last_error());
    return 1;
  }
  printf("abi %d, out_frames(998) = %d, fbank_frames(160000) = %d\n", ppasr_b200_abi_version(), ppasr_b200_out_frames(ctx, 998),
         ppasr_b200_fbank_frames(160000));
  /* a real client now calls ppasr_b200_load_tensor(ctx, "encoder.embed.conv.0.weight", data, 4, shape) for every
   * parameter of the checkpoint and then ppasr_b200_finalize / ppasr_b200_encode / ppasr_b200_ctc_greedy */
  if (ppasr_b200_finalize(ctx) != 0) printf("finalize (no weights loaded, expected to fail): %.80s...\n", ppasr_b200_last_error());
  ppasr_b200_destroy(ctx);
  return 0;
}
