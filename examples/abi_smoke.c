/* A C host binding libskit_b200.so directly (no Python, no torch): what a maintainer of a compiled host would write.
 *   gcc -I include examples/abi_smoke.c -o /tmp/abi_smoke -L visual-tactile-synthesis_b200/csrc -lskit_b200 \
 *       -Wl,-rpath,$PWD/visual-tactile-synthesis_b200/csrc
 * Runs without a GPU: it exercises the argument validation and error reporting of the boundary (every entry point returns
 * 0 or a negative code and never throws; skit_last_error() holds the message), which happen before any CUDA call. */
#include <stdio.h>
#include <string.h>

#include "skit_b200.h"

int main(void) {
    int failures = 0;
    /* 1. a null operand is rejected with SKIT_ERR_INVALID and a message naming the entry point */
    int rc = skit_conv2d_fwd(NULL, NULL, 1, 0, 8, 8, NULL, NULL, NULL, SKIT_NORM_NONE, SKIT_IMPL_AUTO, NULL);
    printf("skit_conv2d_fwd(NULL...) -> %d, \"%s\"\n", rc, skit_last_error());
    failures += !(rc == SKIT_ERR_INVALID && strstr(skit_last_error(), "conv2d_fwd") != NULL);
    /* 2. shape contracts are checked on the host: an LPIPS stem operand must be fp32 [n][h+2][w+2][3] */
    float dummy[4] = {0};
    skit_operand op;
    memset(&op, 0, sizeof op);
    op.p0 = dummy; op.fmt = SKIT_FMT_F32; op.n = 1; op.hp = 10; op.wp = 10; op.c = 4;   /* wrong channel count */
    rc = skit_lpips_scale_fwd(dummy, 1, 3, 8, 8, &op, NULL);
    printf("skit_lpips_scale_fwd(bad operand) -> %d, \"%s\"\n", rc, skit_last_error());
    failures += !(rc == SKIT_ERR_INVALID);
    /* 3. so are value ranges: the StyleGAN2 sub-pixel fold exists for 3x3 filters only */
    rc = skit_sg2_weight_prep(dummy, 1, 1, 5, 2, dummy, NULL);
    printf("skit_sg2_weight_prep(k=5, mode 2) -> %d, \"%s\"\n", rc, skit_last_error());
    failures += !(rc == SKIT_ERR_INVALID);
    printf(failures ? "FAILED\n" : "ok\n");
    return failures;
}
