// comm.cuh — multi-GPU plumbing (definitions in comm.cu).
#pragma once
#include "common.cuh"
#include "peer.cuh"

namespace heat {
int comm_destroy(heat_ctx *ctx);
int comm_allreduce_sum(heat_ctx *ctx, double *buf, int count);     // in place, on ctx->stream
int halo_begin(heat_ctx *ctx, heat_matrix *A, double *x);           // pack + send/recv (side stream)
int halo_end(heat_ctx *ctx, heat_matrix *A);                        // main stream waits for the ghosts
int comm_gather_reduced(heat_ctx *ctx, const std::vector<double> &x_owned, int64_t n_global,
                        std::vector<double> &x_global_on_root);
// peer-memory path (peer.cuh)
PeerRed peer_red_of(const heat_ctx *ctx);
// collective; leaves A->peer null on failure.  need_z: also map the two polynomial-iterate buffers of the
// Chebyshev path.  With ONE rank it builds the degenerate (no neighbour) state the fused kernels run on.
int peer_matrix_setup(heat_ctx *ctx, heat_matrix *A, bool need_z = false);
void peer_matrix_teardown(heat_matrix *A);
}  // namespace heat

struct PeerMatrixState {
    heat::PeerPush push[4];            // push plans into the neighbours' p buffers (0, 1) and z buffers (2, 3)
    bool has_z = false;
    heat::PeerHalo halo;               // my flags
    std::vector<void *> mapped;        // cudaIpcOpenMemHandle mappings to close
};
