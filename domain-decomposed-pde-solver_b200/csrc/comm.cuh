// comm.cuh — multi-GPU plumbing (definitions in comm.cu).
#pragma once
#include "common.cuh"

namespace heat {
int comm_destroy(heat_ctx *ctx);
int comm_allreduce_sum(heat_ctx *ctx, double *buf, int count);     // in place, on ctx->stream
int halo_begin(heat_ctx *ctx, heat_matrix *A, double *x);           // pack + send/recv (side stream)
int halo_end(heat_ctx *ctx, heat_matrix *A);                        // main stream waits for the ghosts
int comm_gather_reduced(heat_ctx *ctx, const std::vector<double> &x_owned, int64_t n_global,
                        std::vector<double> &x_global_on_root);
}  // namespace heat
