/* A plain-C host of the mjv_b200 C ABI (include/mjv_b200.h) — what a non-Python caller of the scoring path links against.
 *
 *   gcc -std=c99 -Iinclude examples/abi_host.c -o abi_host -Lmj_video_b200 -lmjv_b200 -Wl,-rpath,$PWD/mj_video_b200
 *
 * It exercises the entry points that need no GPU (ABI version, struct sizes, the tile-grid choice of dynamic_preprocess,
 * data.py:66-99, and Pillow's bicubic coefficient tables) and shows that a compute entry point fails loudly, with a
 * message, on a machine without a CUDA device instead of falling back to anything. */
#include <stdio.h>
#include <string.h>

#include "mjv_b200.h"

int main(void) {
    int gi = 0, gj = 0;
    mjv_config cfg;
    mjv_handle* h = NULL;
    int rc;

    printf("abi %d; sizeof gemm_desc %d (lib) / %d (here), batch %d / %d\n", mjv_abi_version(), mjv_abi_struct_size(0),
           (int)sizeof(mjv_gemm_desc), mjv_abi_struct_size(3), (int)sizeof(mjv_batch));
    if (mjv_abi_struct_size(0) != (int)sizeof(mjv_gemm_desc) || mjv_abi_struct_size(1) != (int)sizeof(mjv_attn_desc) ||
        mjv_abi_struct_size(2) != (int)sizeof(mjv_config) || mjv_abi_struct_size(3) != (int)sizeof(mjv_batch) ||
        mjv_abi_struct_size(4) != (int)sizeof(mjv_outputs)) {
        printf("struct layout mismatch\n");
        return 2;
    }
    /* a 1344 x 896 frame with max_num = 6 is cut into a 3 x 2 grid (+ thumbnail = 7 tiles) */
    rc = mjv_choose_grid(1344, 896, 1, 6, 448, &gi, &gj);
    printf("grid rc %d: %d x %d\n", rc, gi, gj);
    if (rc != 0 || gi != 3 || gj != 2) return 3;

    memset(&cfg, 0, sizeof(cfg));
    cfg.image_size = 448; cfg.patch_size = 14;
    cfg.vit_hidden = 1024; cfg.vit_heads = 16; cfg.vit_mlp = 4096; cfg.vit_layers = 24; cfg.vit_eps = 1e-6f;
    cfg.llm_hidden = 2048; cfg.llm_inter = 8192; cfg.llm_layers = 24; cfg.llm_heads = 16; cfg.llm_kv_heads = 8;
    cfg.vocab = 92553; cfg.rms_eps = 1e-5f; cfg.max_pos = 32768; cfg.mlp1_eps = 1e-5f;
    cfg.num_objectives = 28; cfg.num_aspects = 5; cfg.gating_hidden = 1024; cfg.gating_n_hidden = 3;
    cfg.gating_temperature = 1.0f;
    rc = mjv_create(&cfg, &h);
    if (rc == 0) {
        printf("handle created (a CUDA device is present)\n");
        mjv_destroy(h);
    } else {
        printf("mjv_create rc %d: %s\n", rc, mjv_last_error());   /* no GPU: MJV_ERR_CUDA, no CPU fallback */
        if (rc != MJV_ERR_CUDA) return 4;
    }
    return 0;
}
