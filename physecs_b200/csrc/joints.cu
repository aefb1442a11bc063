// Joint rows on the device (placeholder: upload rejects joints until the row builders land).
#include "pb_ctx.h"
int pb_joints_upload(pb_ctx* ctx, int n, const int* type, const int* row0, const int* row1, const float* a0p, const float* a0q,
                     const float* a1p, const float* a1q, const float* params8, const int* color) {
    if (n == 0) { ctx->nJoints = 0; return PB_OK; }
    return pb_fail(ctx, PB_EUNSUPPORTED, "joints are not implemented on the device path yet");
}
void pb_joints_free(pb_ctx* ctx) {}
int pb_joint_prep(pb_ctx* ctx, float h) { return PB_OK; }
int pb_joint_solve(pb_ctx* ctx, float h, int warmStart) { return PB_OK; }
