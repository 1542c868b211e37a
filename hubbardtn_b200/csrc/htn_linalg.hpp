// Internal tensor-level linear algebra (see htn_linalg.cpp).  Callers hold ctx->mu.
#pragma once
#include <functional>

#include "htn_program.hpp"

namespace htn {

constexpr int MD_MAXVEC_HOST = 64;  // == MD_MAXVEC of the multidot kernel

struct KrylovInfo {
  double value = 0, residual = 0;
  int applies = 0, converged = 0;
};
typedef std::function<int32_t(const double* x, double* y)> ApplyFn;

void htn_drop_krylov_graphs(htn_ctx* ctx);
int32_t t_copy(const htn_tensor* src, htn_tensor* dst);
int32_t t_transpose(const htn_tensor* src, htn_tensor* dst, int mode);
int32_t t_fill_level(htn_tensor* env, int level, int mode);
int32_t t_qr_inplace(htn_tensor* A, htn_tensor* R);
int32_t t_dot_dev(const htn_tensor* like, const double* x, const double* y, double* out_dev);
int32_t t_dot_host(const htn_tensor* like, const double* x, const double* y, double* out);
int32_t t_normalize(const htn_tensor* like, double* x);
int32_t ensure_krylov(htn_ctx* ctx, int64_t nvec, int64_t stride, int64_t nchunks);
int32_t lanczos_lowest(const htn_tensor* like, const ApplyFn& apply, const double* x0, double* x_out, int krylovdim,
                       double tol, int maxiter, KrylovInfo* info);
int32_t arnoldi_dominant(const htn_tensor* like, const ApplyFn& apply, double* x, int krylovdim, double tol, int maxiter,
                         KrylovInfo* info);
int32_t gmres_solve(const htn_tensor* like, const ApplyFn& apply, const double* b, double* x, int krylovdim, double tol,
                    int maxiter, KrylovInfo* info);

}  // namespace htn
