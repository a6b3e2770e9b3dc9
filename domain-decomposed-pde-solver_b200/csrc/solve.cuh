// solve.cuh — host drivers (definitions in solve.cu).
#pragma once
#include <functional>

#include "common.cuh"
#include "kernels.cuh"

namespace heat {
int sm_count(int device);
int spmv_halo(heat_ctx *ctx, heat_matrix *A, double *x, double *y, CgGate gate, double *dot_out, bool with_yy = false);
int spmv_peer_once(heat_ctx *ctx, heat_matrix *A, const double *x, double *y, int repeat, double *xy_global, double *kernel_ms);
int power_method_device(heat_ctx *ctx, heat_matrix *A, int niters, double tol, uint64_t seed, heat_power_info *info);
int ensure_workspace(heat_matrix *A, bool single_reduce, bool cheb);
// on_poll (optional) is called with the iteration count after every host poll (every check_every
// iterations and when the stopping test fired), while x holds exactly that iterate
int solve_device(heat_ctx *ctx, heat_matrix *A, double *x, const double *b, const heat_solve_opts &o,
                 heat_solve_info *info, const std::function<int(int)> *on_poll = nullptr);
int cheb_apply(heat_ctx *ctx, heat_matrix *A, const heat_solve_opts &o, double lmax, const double *r,
               double *z, CgGate gate, int grid);
int estimate_lambda_max(heat_ctx *ctx, heat_matrix *A, double *lmax_out);
// gmres.cu
int gmres_device(heat_ctx *ctx, heat_matrix *A, double *x, const double *b, const heat_solve_opts &o,
                 heat_solve_info *info, const std::function<int(int)> *on_poll);
// ilu.cu
int ilu0_setup(heat_ctx *ctx, heat_matrix *A);
int ilu_apply(heat_ctx *ctx, heat_matrix *A, const double *v, double *z);
int ilu_export(const heat_matrix *A, double *lu_host, int *n_levels_lower, int *n_levels_upper);
void ilu_free(heat_matrix *A);
}  // namespace heat
