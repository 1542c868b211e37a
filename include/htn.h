/* htn.h — C ABI of the B200-native hot path behind HubbardTN's MPSKit calls.
 *
 * The reference (DaanVrancken/HubbardTN) has no FFI layer: its hot path is reached by
 * Julia dispatch into MPSKit 0.13.1 at src/HubbardFunctions.jl:1010 (IDMRG2), :1012/:1017
 * (VUMPS), :1013/:1016/:1018/:1363-1365 (changebonds) and :1025-1027 (VUMPS &
 * GradientGrassmann).  This header declares the entry points a Julia `ccall` shim (shown
 * in INTEGRATION.md) binds instead.  Each entry point cites the reference interface it
 * replaces.  Conventions:
 *   - plain pointers and sizes only; every function returns int32_t:
 *       0 = ok, >0 = finished but not converged (result usable), <0 = hard error;
 *     no exception ever crosses the boundary; htn_last_error_string() explains <0;
 *   - the library owns all device memory; the caller owns every host buffer it passes;
 *   - handles are opaque and freed explicitly (htn_*_destroy); a context may be used from
 *     any host thread, one call at a time (internal mutex);
 *   - there is NO CPU fallback: without a CUDA device htn_ctx_create fails with
 *     HTN_ERR_NO_DEVICE.
 *
 * Sector labels are int32 triples (p, q1, n):
 *   HTN_SYM_SU2U1: (fermion parity, 2j, U(1) charge)   fZ2 x SU2 x U1  (HubbardFunctions.jl:250;
 *                  the fZ2 x SU2 mu-models of :342 use n = 0)
 *   HTN_SYM_U1U1 : (fermion parity, 2Sz, U(1) charge)   fZ2 x U1 x U1   (HubbardFunctions.jl:247)
 * Block data are real FP64 (the reference stores ComplexF64 but every operator entry of the
 * ground-state path is real: HubbardFunctions.jl:264-290,302-336).
 */
#ifndef HTN_H
#define HTN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTN_OK 0
#define HTN_NOT_CONVERGED 1
#define HTN_ERR_INVALID (-1)
#define HTN_ERR_NO_DEVICE (-2)
#define HTN_ERR_CUDA (-3)
#define HTN_ERR_OOM (-4)
#define HTN_ERR_SHAPE (-5)

#define HTN_SYM_SU2U1 0
#define HTN_SYM_U1U1 1

#define HTN_SIDE_LEFT 0
#define HTN_SIDE_RIGHT 1

/* tensor kinds (block label layout in htn_tensor_blocktable) */
#define HTN_T_MPS 0  /* A / AC : (V_l (x) P) <- V_r   labels (l, s, r)   block [n_l , n_r ] */
#define HTN_T_BOND 1 /* C      : V <- V               labels (c, c, c)   block [n_c , n_c ] */
#define HTN_T_ENVL 2 /* GL     : (bra, level, ket)    labels (a, l', l)  block [n_l', n_l ] */
#define HTN_T_ENVR 3 /* GR     : (ket, level, bra)    labels (b, r, r')  block [n_r , n_r'] */
#define HTN_T_MPST 4 /* A^T    : blockwise transpose  labels (l, s, r)   block [n_r , n_l ], grouped by l */
#define HTN_T_MPS2 5 /* AC2    : (V_l (x) P (x) P) <- V_r in the fusion-tree basis (l,s1->m),(m,s2->r):
                       labels (l, s1, m, s2, r), m = position in htn_tensor_mid_sectors; block [n_l , n_r ] */

typedef struct htn_ctx htn_ctx;
typedef struct htn_space htn_space;   /* graded bond space  (TensorKit GradedSpace)          */
typedef struct htn_legs htn_legs;     /* list of single multiplets: physical space or the     */
                                      /* MPO virtual levels (BlockTensorKit SumSpace)         */
typedef struct htn_tensor htn_tensor; /* block-sparse tensor in the HBM arena                 */
typedef struct htn_mpo htn_mpo;       /* one site of an InfiniteMPOHamiltonian (reduced form) */
typedef struct htn_plan htn_plan;     /* H_eff application plan (MPSKit derivative operator)  */

/* ---- context ------------------------------------------------------------------------ */
/* Replaces: module initialisation of the reference (BLAS/thread policy, HubbardFunctions.jl:1-55). */
int32_t htn_ctx_create(int32_t device, htn_ctx** out);
int32_t htn_ctx_destroy(htn_ctx* ctx);
const char* htn_last_error_string(htn_ctx* ctx); /* valid until the next call on ctx; ctx may be NULL */
int32_t htn_version(void);
int32_t htn_ctx_synchronize(htn_ctx* ctx);
/* the library's CUDA stream (cudaStream_t) of this context: collectives of the caller (NCCL allreduce of a sharded y)
   are enqueued on it so that they are ordered after the apply without a host synchronisation */
int32_t htn_ctx_stream(htn_ctx* ctx, void** stream);

/* ---- spaces ------------------------------------------------------------------------- */
/* Replaces: Vect[I](sector => multiplicity ...) (HubbardFunctions.jl:248,251,261,282,343).
 * Sectors are stored in the canonical order (sorted; see DESIGN.md "Block tables"); the
 * caller may pass them in any order. */
int32_t htn_space_create(htn_ctx* ctx, int32_t sym, int32_t nsec, const int32_t* labels /*[nsec][3]*/,
                         const int32_t* mult /*[nsec]*/, htn_space** out);
int32_t htn_space_destroy(htn_space* s);
int32_t htn_space_info(const htn_space* s, int32_t* nsec, int32_t* labels /*[nsec][3] or NULL*/,
                       int32_t* mult /*[nsec] or NULL*/);
/* ordered multiplets, one each (order is kept as given) */
int32_t htn_legs_create(htn_ctx* ctx, int32_t sym, int32_t n, const int32_t* labels /*[n][3]*/,
                        htn_legs** out);
int32_t htn_legs_destroy(htn_legs* l);

/* ---- tensors ------------------------------------------------------------------------ */
/* Replaces: TensorMap(zeros, T, codomain <- domain) and the InfiniteMPS site tensors
 * (HubbardFunctions.jl:264-290, 958, 990); storage = packed HBM arena + block table. */
int32_t htn_tensor_create_mps(htn_ctx* ctx, const htn_space* Vl, const htn_legs* P, const htn_space* Vr,
                              htn_tensor** out);
int32_t htn_tensor_create_bond(htn_ctx* ctx, const htn_space* V, htn_tensor** out);
/* identity_level >= 0 flags the level that holds the unit tensor (GL[1] / GR[chi]); -1: none */
int32_t htn_tensor_create_env(htn_ctx* ctx, int32_t side, const htn_space* V, const htn_legs* M,
                              int32_t identity_level, htn_tensor** out);
/* two-site tensor x2 = AC (x) AR (MPSKit `AC2`, reached from IDMRG2 at HubbardFunctions.jl:1010) */
int32_t htn_tensor_create_mps2(htn_ctx* ctx, const htn_space* Vl, const htn_legs* P1, const htn_legs* P2,
                               const htn_space* Vr, htn_tensor** out);
/* intermediate sectors of a two-site tensor (canonical order); labels may be NULL to query n */
int32_t htn_tensor_mid_sectors(const htn_tensor* t, int32_t* n, int32_t* labels /*[n][3]*/);
/* block table with 5 labels per block (kinds with 3 labels report 0 for the last two) */
int32_t htn_tensor_blocktable5(const htn_tensor* t, int32_t* nblocks, int64_t* nelem, int32_t* labels /*[n][5]*/,
                               int32_t* rows, int32_t* cols, int64_t* offsets);
int32_t htn_tensor_create_like(const htn_tensor* t, htn_tensor** out);
/* blockwise-transposed companion (kind HTN_T_MPST) of an MPS tensor; filled by htn_tensor_transpose */
int32_t htn_tensor_create_transposed(const htn_tensor* t, htn_tensor** out);
/* dst = blockwise transpose of src (MPS <-> MPST, or bond -> bond); weighted != 0 multiplies block
 * (l,s,r) by sqrt(dim r / dim l) (MPS -> MPST) or its inverse (MPST -> MPS): the isometric weights
 * of the right-orthonormal form */
int32_t htn_tensor_transpose(const htn_tensor* src, htn_tensor* dst, int32_t weighted);
int32_t htn_tensor_destroy(htn_tensor* t);
int32_t htn_tensor_kind(const htn_tensor* t); /* HTN_T_* */
/* raw device arena of a tensor (padded layout, zero padding; nelem_padded doubles) -- for collectives over
 * NVLink that reduce whole vectors in place (NCCL allreduce of the sharded H_eff sums) */
int32_t htn_tensor_device_ptr(const htn_tensor* t, void** ptr, int64_t* nelem_padded);
/* copy of the bond space a tensor was built on: which = 0 left / first, 1 right / second (caller destroys) */
int32_t htn_tensor_space(const htn_tensor* t, int32_t which, htn_space** out);
/* Block table: nblocks rows of (label0,label1,label2) POSITIONS into the spaces, rows, cols,
 * packed host offset (elements, row-major blocks back to back, no padding).  Pass NULL
 * arrays to query nblocks / nelem only. */
int32_t htn_tensor_blocktable(const htn_tensor* t, int32_t* nblocks, int64_t* nelem, int32_t* labels,
                              int32_t* rows, int32_t* cols, int64_t* offsets);
/* host <-> device, packed host layout as described by htn_tensor_blocktable */
int32_t htn_tensor_upload(htn_tensor* t, const double* host, int64_t nelem);
int32_t htn_tensor_download(const htn_tensor* t, double* host, int64_t nelem);

/* ---- MPO ---------------------------------------------------------------------------- */
/* Replaces: one site tensor of the InfiniteMPOHamiltonian built by @mpoham
 * (HubbardFunctions.jl:435-465).  entry i: (a, s', s, b) = idx[i][0..3] positions into
 * (Ml, P, P, Mr); coupled sector label c = clabel[i][0..2]; reduced value val[i]. */
int32_t htn_mpo_create(htn_ctx* ctx, const htn_legs* Ml, const htn_legs* P, const htn_legs* Mr,
                       int32_t nnz, const int32_t* idx /*[nnz][4]*/, const int32_t* clabel /*[nnz][3]*/,
                       const double* val, htn_mpo** out);
/* Same from a DENSE invariant tensor dense[a m_a][s' m'][s m][b m_b] (row-major, every multiplet
 * expanded with m = -j..+j): the library does the Wigner-Eckart projection (entries with |w| <= tol are
 * dropped) and fails with HTN_ERR_INVALID if the tensor is not invariant.  This is the form in which a
 * host can hand over `TensorMap` operator data without knowing the library's reduced convention. */
int32_t htn_mpo_create_dense(htn_ctx* ctx, const htn_legs* Ml, const htn_legs* P, const htn_legs* Mr,
                             const double* dense, double tol, htn_mpo** out);
/* reduced entries of an MPO tensor (arrays may be NULL to query nnz) */
int32_t htn_mpo_entries(const htn_mpo* w, int32_t* nnz, int32_t* idx /*[nnz][4]*/, int32_t* clabel /*[nnz][3]*/,
                        double* val);
int32_t htn_mpo_destroy(htn_mpo* w);

/* ---- H_eff -------------------------------------------------------------------------- */
/* Replaces: MPSKit `AC_hamiltonian(site, psi, H, psi, envs)` / `∂AC` reached from
 * find_groundstate at HubbardFunctions.jl:1012,1017,1027.  The plan references GL, GR (not
 * copied; they must outlive the plan) and copies W. `like` fixes the block structure of x,y. */
int32_t htn_plan_heff_ac(htn_ctx* ctx, const htn_tensor* GL, const htn_mpo* W, const htn_tensor* GR,
                         const htn_tensor* like, htn_plan** out);
/* Multi-GPU: shard `shard` of `nshards` of the same application ("partitioned by symmetry sector and
   by MPO virtual index, NCCL allreduce only for the sharded sums"; reference analogue: MPSKit's per-term task
   parallelism, HubbardFunctions.jl:37).  The plan computes a PARTIAL y; allreduce(sum) over the shards gives H_AC x.
   Units are the left symmetry sectors of the output (heavy sectors are split further by MPO level). */
int32_t htn_plan_heff_ac_sharded(htn_ctx* ctx, const htn_tensor* GL, const htn_mpo* W, const htn_tensor* GR,
                                 const htn_tensor* like, int32_t nshards, int32_t shard, htn_plan** out);
/* Replaces: MPSKit `AC2_hamiltonian` / `∂∂AC2` (two-site effective Hamiltonian of IDMRG2,
 * HubbardFunctions.jl:1010): y2 = GL . x2 . W1 . W2 . GR; GL = left environment of the first site,
 * GR = right environment of the second site; `like` is a HTN_T_MPS2 tensor. */
int32_t htn_plan_heff_ac2(htn_ctx* ctx, const htn_tensor* GL, const htn_mpo* W1, const htn_mpo* W2,
                          const htn_tensor* GR, const htn_tensor* like, htn_plan** out);
/* x2[l,s1,m,s2,r] = A1[l,s1,m] . A2[m,s2,r]  (x2 created with htn_tensor_create_mps2) */
int32_t htn_contract_two_site(const htn_tensor* A1, const htn_tensor* A2, htn_tensor* x2);
/* Replaces: MPSKit `C_hamiltonian` / `∂C` (zero-site effective Hamiltonian, same call sites).
 * GL = left environment on the bond of C (i.e. of the NEXT site), GR = right environment of
 * this site; `like` is a bond tensor. */
int32_t htn_plan_heff_c(htn_ctx* ctx, const htn_tensor* GL, const htn_tensor* GR, const htn_tensor* like,
                        htn_plan** out);
int32_t htn_plan_destroy(htn_plan* p);
/* y = H_AC x on the device (x, y created with the plan's `like` structure; x != y) */
int32_t htn_heff_apply(htn_plan* p, const htn_tensor* x, htn_tensor* y);
/* same through host buffers: upload x, apply, download y (the e2e path of bench.py) */
int32_t htn_heff_apply_host(htn_plan* p, const double* x_host, double* y_host, int64_t nelem);
/* plan statistics: stats[0]=algorithmic flops, [1]=stage-L flops, [2]=stage-R
 * flops, [3]=#stage-L GEMMs, [4]=#stage-R GEMM segments, [5]=#mix targets, [6]=#mix sources,
 * [7]=workspace bytes, [8]=#stage-L tiles, [9]=#stage-R tiles, [10]=padded (executed) flops,
 * [11]=kernel launches per apply */
int32_t htn_plan_stats(const htn_plan* p, double* stats, int32_t n);
/* time `reps` applies with CUDA events on the library stream: ms[0]=total per apply,
 * ms[1..3] = stage L / W / R per apply (each stage timed in its own pass) */
int32_t htn_plan_profile(htn_plan* p, const htn_tensor* x, htn_tensor* y, int32_t reps, float* ms /*[4]*/);

/* `reps` back-to-back applies bracketed by two CUDA events on the library stream (device time
 * of the whole timed region, ms) — what bench.py reports as ms_per_step * steps */
int32_t htn_heff_time(htn_plan* p, const htn_tensor* x, htn_tensor* y, int32_t reps, float* ms_total);

/* ---- environment transfers ---------------------------------------------------------- */
/* Replaces: MPSKit `TransferMatrix` application / `left_cyclethrough` used by `environments`
 * (reached from find_groundstate, HubbardFunctions.jl:1010-1027).
 *   side LEFT : GL'[b,r',r] = sum_a  A^T[l',s',r'] GL[a,l',l] W[a,b] A[l,s,r]   (env on the left bond -> right bond)
 *   side RIGHT: GR'[a,l,l'] = sum_b  A[l,s,r] W[a,b] GR[b,r,r'] A^T[l',s',r']   (right bond -> left bond)
 * At = htn_tensor_transpose(A) (kind HTN_T_MPST).  env_in/env_out fix the block structure; bond
 * tensors are accepted as one-level environments (MPO-free transfer matrix, W = one trivial level). */
int32_t htn_plan_transfer(htn_ctx* ctx, int32_t side, const htn_mpo* W, const htn_tensor* A, const htn_tensor* At,
                          const htn_tensor* env_in, const htn_tensor* env_out, htn_plan** out);
int32_t htn_transfer_apply(htn_plan* p, const htn_tensor* A, const htn_tensor* At, const htn_tensor* env_in,
                           htn_tensor* env_out);

/* ---- Krylov eigensolver --------------------------------------------------------------- */
/* Replaces: KrylovKit `eigsolve(H_eff, x0, 1, :SR, Lanczos(krylovdim, tol, maxiter))` as called by
 * MPSKit's VUMPS / IDMRG.  x: unit-norm eigenvector with <x0,x> >= 0.
 * Returns HTN_NOT_CONVERGED (>0) with the last Ritz pair when maxiter restarts did not reach tol. */
int32_t htn_eigsolve(htn_plan* p, const htn_tensor* x0, htn_tensor* x, int32_t krylovdim, double tol, int32_t maxiter,
                     double* eigenvalue, double* residual, int32_t* applies);

/* ---- gauge fixing --------------------------------------------------------------------- */
/* Replaces: TensorKit `leftorth!(A, alg = QRpos())`: A = Q R per coupled (right) sector, diag R > 0.
 * A: MPS tensor (Q same structure, R bond tensor on the right space) or bond tensor. */
int32_t htn_qrpos(const htn_tensor* A, htn_tensor* Q, htn_tensor* R);
/* Replaces: TensorKit `rightorth!(A, alg = LQpos())`: A = L Q per left sector, diag L > 0. */
int32_t htn_lqpos(const htn_tensor* A, htn_tensor* L, htn_tensor* Q);
/* Replaces: MPSKit `regauge!`: AL = Q(AC) Q(C)^T. */
int32_t htn_regauge(const htn_tensor* AC, const htn_tensor* C, htn_tensor* AL);
/* Replaces: the bond products MPSKit writes as `A * C` / `C * A` (`_mul_tail`, `_mul_front`; AC = AL C = C AR):
 * right != 0: out[l,s,r] = A[l,s,r] . C[r];  right == 0: out[l,s,r] = C[l] . A[l,s,r].  A new MPS tensor with the
 * structure of A is created (caller destroys it). */
int32_t htn_mul_bond(const htn_tensor* A, const htn_tensor* C, int32_t right, htn_tensor** out);
/* Replaces: MPSKit `uniform_rightorth!` (entered through InfiniteMPS(...), HubbardFunctions.jl:958,990
 * and VUMPS's gauge step): from left-orthonormal AL[0..n) and a guess for C[n-1] compute AR[i], C[i]
 * (C[i] on the bond right of site i, unit norm) with AL[i] C[i] = C[i-1] AR[i]. */
int32_t htn_gauge_right(htn_ctx* ctx, int32_t nsites, htn_tensor* const* AL, const htn_tensor* C_guess,
                        htn_tensor* const* AR, htn_tensor* const* C, double tol, int32_t maxiter, int32_t* iterations,
                        double* delta);

/* ---- truncated SVD ---------------------------------------------------------------------- */
/* Replaces: TensorKit `tsvd!(ac2; trunc = truncbelow(cut))` / `truncdim` as used by IDMRG2 and
 * `changebonds(.., SvdCut)` (HubbardFunctions.jl:1010,1013,1018,1363-1365): x2 = AL . C . AR with C the
 * diagonal of kept Schmidt values.  Keeps sigma >= cut * ||x2||, at most maxdim multiplets (maxdim <= 0:
 * no cap).  Creates the new middle space and the three factors (caller destroys them). */
int32_t htn_tsvd(const htn_tensor* x2, double cut, int32_t maxdim, htn_space** Vm, htn_tensor** AL, htn_tensor** C,
                 htn_tensor** AR, double* discarded_weight, int32_t* kept);

/* Mirror image (MPSKit `uniform_leftorth!`): from right-orthonormal AR[0..n) and a guess for C[n-1]
 * compute AL[i], C[i] with AL[i] C[i] = C[i-1] AR[i]. */
int32_t htn_gauge_left(htn_ctx* ctx, int32_t nsites, htn_tensor* const* AR, const htn_tensor* C_guess,
                       htn_tensor* const* AL, htn_tensor* const* C, double tol, int32_t maxiter, int32_t* iterations,
                       double* delta);

/* ---- environments and ground-state driver ------------------------------------------------ */
/* Replaces: MPSKit `environments(psi, H)` / `recalculate!` for an InfiniteMPOHamiltonian in Jordan
 * form (level 0 and level chi-1 carry the identity): GL[i] on the bond left of site i (created with
 * identity_level 0), GR[i] on the bond right of site i (identity_level chi-1); the identity-diagonal
 * level is solved by GMRES.  energy_left/right = energy per UNIT CELL seen by either side. */
int32_t htn_environments(htn_ctx* ctx, int32_t nsites, htn_tensor* const* AL, htn_tensor* const* AR,
                         htn_tensor* const* C, const htn_mpo* const* W, htn_tensor* const* GL, htn_tensor* const* GR,
                         double tol, int32_t krylovdim, int32_t maxiter, double* energy_left, double* energy_right);
/* Replaces: `find_groundstate(psi, H, VUMPS(; tol, maxiter))` (HubbardFunctions.jl:1012,1017,1025-1027)
 * on fixed bond spaces.  In/out: AL, AR, C, AC (mixed gauge), GL, GR.  delta = final Galerkin error
 * (the `δ` MPSKit returns, HF:1027).  log (may be NULL): rows of 8 doubles per iteration
 * (galerkin error, energy per site, gauge iterations, H_eff applies, seconds spent in the eigensolves,
 * in gauge fixing, in the environments, GMRES operator applications). Returns HTN_NOT_CONVERGED when
 * maxiter is reached (state usable, as MPSKit does). */
int32_t htn_vumps(htn_ctx* ctx, int32_t nsites, htn_tensor* const* AL, htn_tensor* const* AR, htn_tensor* const* C,
                  htn_tensor* const* AC, const htn_mpo* const* W, htn_tensor* const* GL, htn_tensor* const* GR,
                  double tol, int32_t maxiter, int32_t krylovdim, double* delta, double* energy_per_site,
                  int32_t* iterations, double* log, int32_t log_cap);
/* Replaces: the `GradientGrassmann(; maxiter, tol)` stage of `VUMPS(...) & GradientGrassmann(...)`
 * (HubbardFunctions.jl:1025-1027; MPSKit GrassmannMPS + OptimKit): Riemannian conjugate-gradient descent
 * of the energy per unit cell over the left isometries AL on fixed bond spaces (metric C C^T regularised,
 * positive-QR retraction, Polak-Ribiere+ directions, backtracking on the energy).  In/out as htn_vumps;
 * delta = final Galerkin error (gradient norm).  log rows of 8 doubles per iteration (galerkin error,
 * energy per site, accepted step, energy evaluations so far, beta, slope, seconds, 0).  The converged
 * state is the VUMPS fixed point; the iteration path is not OptimKit's.  HTN_NOT_CONVERGED at maxiter. */
int32_t htn_gradient_grassmann(htn_ctx* ctx, int32_t nsites, htn_tensor* const* AL, htn_tensor* const* AR,
                               htn_tensor* const* C, htn_tensor* const* AC, const htn_mpo* const* W,
                               htn_tensor* const* GL, htn_tensor* const* GR, double tol, int32_t maxiter,
                               int32_t krylovdim, double* delta, double* energy_per_site, int32_t* iterations,
                               double* log, int32_t log_cap);
/* Replaces: `find_groundstate(psi, H, IDMRG2(trscheme = truncbelow(cut), tol, maxiter))`
 * (HubbardFunctions.jl:1010): two-site infinite DMRG over a unit cell of nsites >= 2 with bond spaces
 * re-defined by the truncated SVD (keep Schmidt values >= cut, at most maxdim multiplets if maxdim > 0, at most a full
 * dimension of |maxdim| if maxdim < 0).
 * In/out handle arrays: the library destroys every tensor it replaces and stores the new handle; the
 * caller destroys the final ones.  delta = || C_new - C_old || on the common subspace of the edge bond.
 * log (may be NULL): rows of 8 doubles (delta, sum of D_red over the bonds, cumulative H_AC2 applies, cumulative
 * seconds in host planning / Lanczos / truncated SVD / environment growth, 0).
 * The result is NOT yet a consistent uniform MPS: call htn_mixed_gauge (MPSKit does `InfiniteMPS(psi.AR)`).
 * krylovdim <= 0 skips the eigensolves: one such iteration is the truncation-only sweep of
 * `changebonds(psi, SvdCut(trscheme = truncdim(D) | truncbelow(cut)))` (HubbardFunctions.jl:1013,1018,1365). */
int32_t htn_idmrg2(htn_ctx* ctx, int32_t nsites, htn_tensor** AL, htn_tensor** AR, htn_tensor** C, htn_tensor** AC,
                   const htn_mpo* const* W, double cut, double tol, int32_t maxiter, int32_t krylovdim, double eig_tol,
                   int32_t maxdim, double* delta, int32_t* iterations, double* log, int32_t log_cap);
/* Replaces: `InfiniteMPS(A...)` gauge fixing (MPSKit `uniform_leftorth!/uniform_rightorth!`).
 *   from_right = 0: AL[i] hold left isometries (modified in place: AL <- Q(AL));
 *   from_right = 1: AR[i] hold right isometries -- the list whose bond spaces chain after an IDMRG2
 *                   iteration (MPSKit: `InfiniteMPS(psi.AR)`); AL, C follow by the iterated QR.
 * Then AR, C by the iterated LQ and AC = AL C.  AL[i], AR[i], AC[i] share one block structure, C[i] is a
 * bond tensor on the right space of site i, C_guess lives on the last bond. */
int32_t htn_mixed_gauge(htn_ctx* ctx, int32_t nsites, htn_tensor* const* AL, const htn_tensor* C_guess,
                        htn_tensor* const* AR, htn_tensor* const* C, htn_tensor* const* AC, int32_t from_right,
                        double tol, int32_t maxiter, int32_t* iterations);
/* Replaces: `changebonds(psi, SvdCut(; trscheme))` (kind 0; HubbardFunctions.jl:1013,1018,1365) and
 * `changebonds(psi, H, VUMPSSvdCut(; trscheme))` (kind 1; HubbardFunctions.jl:1016,1363; unit cells of >= 2 sites, MPSKit
 * `changebonds_n`).  trscheme: cut > 0 = truncbelow(cut) relative to the norm of the two-site tensor; maxdim > 0 caps the
 * kept multiplets sum_c n_c per bond; maxdim < 0 = truncdim(|maxdim|): caps the FULL dimension sum_c dim(c) n_c (what
 * TensorKit counts and `dim_state` reports, HF:1363-1365, 1402).  The kept set is decided in one pass from the singular
 * values.  In/out handle arrays as in htn_idmrg2 (replaced tensors are destroyed by the library); on return the state is a
 * consistent uniform MPS in mixed gauge.  krylovdim / eig_tol (kind 1; <= 0: 30 / 1e-10), tol_gauge (<= 0: 1e-12). */
int32_t htn_changebonds(htn_ctx* ctx, int32_t kind, int32_t nsites, htn_tensor** AL, htn_tensor** AR, htn_tensor** C,
                        htn_tensor** AC, const htn_mpo* const* W, double cut, int32_t maxdim, int32_t krylovdim,
                        double eig_tol, double tol_gauge);
/* Replaces: `expectation_value(psi, i => op)` (HubbardFunctions.jl:1448-1449,1507,1533) for one-site
 * operators that are scalars on every physical multiplet (n, n_up, n_dn): values[s] per multiplet s. */
int32_t htn_expval_diag(const htn_tensor* AC, const double* values, int32_t nvalues, double* out);

/* Replaces: MPSKit `entanglement_spectrum(psi, site)` (the Schmidt spectrum the north star names as a parity
 * target; C from HubbardFunctions.jl's ground state): singular values of every block of the bond matrix C,
 * descending per sector, concatenated in block order (nout = sum_c n_c; sector c counts dim(c) times in
 * the full spectrum). */
int32_t htn_entanglement_spectrum(const htn_tensor* C, double* out, int64_t nout);

/* ---- vector algebra on tensors of identical structure (KrylovKit inner products) ------ */
/* <x,y> = sum_blocks dim(coupled sector) tr(x^T y)   (TensorKit inner product) */
int32_t htn_tensor_dot(const htn_tensor* x, const htn_tensor* y, double* out);
int32_t htn_tensor_axpby(double alpha, const htn_tensor* x, double beta, htn_tensor* y); /* y = a x + b y */

/* ---- test hooks --------------------------------------------------------------------- */
/* recoupling network N(l',s',r'; l,s,r; a,b,c) of DESIGN.md (labels are int32 triples) */
int32_t htn_network_coefficient(int32_t sym, const int32_t* nine_labels /*[9][3]*/, double* out);
/* device-timed Gram-Schmidt pass of the Krylov solvers on vectors shaped like `like` with nvec basis
 * vectors: ms[0] multidot, ms[1] multiaxpy (per pass), bytes[0..1] the algorithmic bytes of either */
int32_t htn_probe_krylov(const htn_tensor* like, int32_t nvec, int32_t reps, float* ms, double* bytes);
/* host-only: dominant eigenpair of a small upper-Hessenberg matrix (the Arnoldi inner problem) */
int32_t htn_test_hessenberg_dominant(int32_t m, const double* H, double* theta, double* y);
/* FP64 peak probes (dependent-free DMMA.8x8x4 / DFMA loops): TFLOP/s on the ctx device */
int32_t htn_probe_fp64_peak(htn_ctx* ctx, int32_t which /*0=DMMA,1=DFMA*/, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* HTN_H */
