/*
 * fcx.h -- C ABI of libfcx.so: B200-native (sm_100a, fp64) per-quadrature-point
 * constitutive updates, the drop-in for fenics-constitutive's
 *     IncrSmallStrainModel.evaluate(t, del_t, grad_del_u, stress, tangent, history)
 * (reference: src/fenics_constitutive/models/interfaces.py:82-101).
 *
 * This header is what a maintainer of the reference would bind (ctypes / cffi /
 * pybind11 / pyo3-free) in place of `fenics_constitutive._bindings`
 * (reference: bindings/src/lib.rs:76-129, whose batch driver is
 * comfe-rs/src/interfaces.rs:354-456).  Same granularity: ONE call per
 * `evaluate`, covering all quadrature points (QPs) of the rank.
 *
 * Array contract (identical to the reference): flat, C-contiguous
 * float64.  g = geometric dim, s = Mandel dim of the constraint.
 *     grad_del_u [n][g][g]   read-only, ufl.nabla_grad convention
 *     stress     [n][s]      in: sigma_n, out: sigma_{n+1}   (Mandel)
 *     tangent    [n][s][s]   write-only, row-major
 *     history    [n][dim]    in: committed values, out: trial values
 * Mandel order [xx, yy, zz, xy, xz, yz], shear scaled by Python's 1/2**0.5
 * (reference: models/utils.py:199-204).
 *
 * `tangent` may be NULL in every fcx_<model>_evaluate[_host] call and in
 * fcx_mises_form: a STRESS-ONLY evaluate, the reference's `tangent: Option<..>`
 * (bindings/src/lib.rs:83,109-113; comfe-rs/src/interfaces.rs:368,441-455).
 * Stress and history are updated exactly as with a tangent; no tangent byte is
 * formed, stored or sent over PCIe (VonMises3D: 280 instead of 568 B per QP).
 *
 * Two families of entry points:
 *   fcx_<model>_evaluate       DEVICE pointers; enqueue-only on `stream`
 *                              (a cudaStream_t passed as void*, NULL = default
 *                              stream); never synchronises.
 *   fcx_<model>_evaluate_host  HOST pointers (pageable or pinned); chunked,
 *                              multi-stream H2D -> kernel -> D2H pipeline;
 *                              returns after the results are in host memory.
 * All functions return FCX_OK (0), a negative FCX_ERR_* code, or -- for the
 * Mises host entry point -- the positive number of QPs whose return-mapping
 * Newton iteration did not converge (the reference raises RuntimeError there,
 * models/mises_plasticity_isotropic_hardening.py:141-143).
 * No function throws, aborts or keeps a caller pointer past its return
 * (device entry points: past completion of the enqueued work).
 *
 * Threading (reference: evaluate is synchronous and not re-entrant, one call per
 * MPI rank at a time): the device entry points keep no per-call state (parameters
 * travel by value) and may be called from several threads on different streams;
 * the *_host entry points share one pipeline context per process (streams, chunk
 * buffers, pinned ring slots, the host-thread pool) and serialise on a mutex.
 * The fcx_tune / fcx_host_* knobs are process-wide settings meant to be set
 * before the calls they affect, not concurrently with them.
 */
#ifndef FCX_H
#define FCX_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FCX_VERSION 200 /* 0.2.0 */

/* StressStrainConstraint values, reference models/interfaces.py:23-27 */
enum fcx_constraint {
    FCX_UNIAXIAL_STRAIN = 1, /* s=1 g=1 */
    FCX_UNIAXIAL_STRESS = 2, /* s=1 g=1 */
    FCX_PLANE_STRAIN = 3,    /* s=4 g=2 */
    FCX_PLANE_STRESS = 4,    /* s=4 g=2 */
    FCX_FULL = 5             /* s=6 g=3 */
};

enum fcx_status {
    FCX_OK = 0,
    FCX_ERR_CONSTRAINT = -1, /* unknown / unsupported constraint for the model   */
    FCX_ERR_TIMESTEP = -2,   /* del_t <= 0 (reference asserts, spring_kelvin_model.py:72) */
    FCX_ERR_NULL = -3,       /* a required pointer is NULL                        */
    FCX_ERR_CUDA = -4,       /* a CUDA runtime call failed; see fcx_last_cuda_error() */
    FCX_ERR_ARG = -5         /* other invalid argument                            */
};

/* History layout of the Mises plastic strain `eps_n`:
 * AOS = the reference contract [n][6]; SOA = six planes [6][n] (device-resident
 * state owned by a GPU solver; coalesced without staging). */
enum fcx_layout { FCX_LAYOUT_AOS = 0, FCX_LAYOUT_SOA = 1 };

int fcx_version(void);
const char *fcx_strerror(int code);
/* Text of the last CUDA error seen by this thread's calls ("" if none). */
const char *fcx_last_cuda_error(void);
/* s and g of a constraint (reference models/interfaces.py:30-73); -1 if unknown. */
int fcx_stress_strain_dim(int constraint);
int fcx_geometric_dim(int constraint);
/* Bind the calling thread to a CUDA device (one rank per GPU: LOCAL_RANK). */
int fcx_set_device(int device);

/* ------------------------------------------------------------------ device */

/* LinearElasticityModel.evaluate -- reference models/linear_elasticity_model.py:26-45:
 *   stress += strain_from_grad_u(grad) @ D ;  tangent[:] = tile(D.flatten(), n)
 * D: HOST pointer to s*s doubles (row-major), e.g. get_elastic_tangent(E,nu,c)
 * (models/utils.py:25-93) or comfe-rs' 2mu*P_dev+3kappa*P_vol
 * (comfe-rs/src/linear_elasticity.rs:60-73).  Copied at call time. */
int fcx_elastic_evaluate(int constraint, const double *D_host, size_t n,
                         const double *grad_del_u, double *stress, double *tangent,
                         void *stream);

/* VonMises3D.evaluate -- reference models/mises_plasticity_isotropic_hardening.py:57-175.
 * params = {p_ka, p_mu, p_y0, p_y00, p_w} (HOST, :51-55).  FULL constraint only.
 * eps_n [n][6] (or [6][n] if eps_layout == FCX_LAYOUT_SOA), alpha [n].
 * plastic_flag: optional DEVICE u8[n], 1 where the trial state is plastic
 *   (phitr > 0, :98); may be NULL.
 * status: optional DEVICE int[2]; status[0] is incremented once per QP whose
 *   Newton loop exceeded 100 iterations (:141-143), status[1] receives
 *   min(QP index) of such points via atomicMin (caller initialises to INT_MAX);
 *   may be NULL. */
int fcx_mises_evaluate(const double *params_host, size_t n, const double *grad_del_u,
                       double *stress, double *tangent, double *eps_n, double *alpha,
                       int eps_layout, unsigned char *plastic_flag, int *status,
                       void *stream);

/* MisesPlasticityLinearHardening3D.evaluate -- the reference's Rust model
 * comfe-rs/src/mises_plasticity.rs:58-126 behind models/rust_models.py:144-161
 * (batch driver comfe-rs/src/interfaces.rs:354-456).  params = {mu, kappa, y_0, h}
 * (HOST).  history: ONE array [n][7] = [alpha, plastic_strain[6]] per QP
 * (key "history", bindings/src/lib.rs:90-100).  plastic_flag: optional u8[n]. */
int fcx_mises_linear_hardening_evaluate(const double *params_host, size_t n,
                                        const double *grad_del_u, double *stress,
                                        double *tangent, double *history,
                                        unsigned char *plastic_flag, void *stream);

/* DruckerPrager3D / DruckerPragerHyperbolic3D.evaluate -- the reference's Rust models
 * comfe-rs/src/plasticity/{general.rs:105-266, drucker_prager_classic.rs:75-108,
 * drucker_prager_hyperbolic.rs:64-102} behind models/rust_models.py:96-141.
 * hyperbolic = 0: params = {mu, kappa, a, b, b_flow}; 1: {mu, kappa, a, b, d, b_flow} (HOST).
 * history: ONE array [n][7] = [alpha, plastic_strain[6]].  plastic_flag: optional u8[n].
 * status: as for fcx_mises_evaluate -- counts the points where the Rust code would panic
 * (Newton not converged after 25 iterations, or the classic model's apex assert); such
 * points are left unchanged. */
int fcx_drucker_prager_evaluate(int hyperbolic, const double *params_host, size_t n,
                                const double *grad_del_u, double *stress, double *tangent,
                                double *history, unsigned char *plastic_flag, int *status,
                                void *stream);

/* SpringKelvinModel.evaluate -- reference models/spring_kelvin_model.py:43-88.
 * D0 (s*s), I2 (s): HOST pointers to the constants of __init__ (:24-41);
 * mu0, lam0, mu1, tau as built there.  history: strain_visco [n][s], strain [n][s]. */
int fcx_kelvin_evaluate(int constraint, const double *D0_host, const double *I2_host,
                        double mu0, double lam0, double mu1, double tau, double del_t,
                        size_t n, const double *grad_del_u, double *stress, double *tangent,
                        double *strain_visco, double *strain, void *stream);

/* SpringMaxwellModel.evaluate -- reference models/spring_maxwell_model.py:40-88.
 * D0, D1 (s*s): HOST pointers to the constants of __init__ (:24-38). */
int fcx_maxwell_evaluate(int constraint, const double *D0_host, const double *D1_host,
                         double mu1, double tau, double del_t, size_t n,
                         const double *grad_del_u, double *stress, double *tangent,
                         double *strain_visco, double *strain, void *stream);

/* strain_from_grad_u -- reference models/utils.py:132-208.  strain [n][s]. */
int fcx_strain_from_grad_u(int constraint, size_t n, const double *grad_del_u,
                           double *strain, void *stream);

/* 3D -> 1D/2D adapters, reference models/utils.py:211-412 (UniaxialStrainFrom3D,
 * PlaneStrainFrom3D).  constraint = FCX_UNIAXIAL_STRAIN or FCX_PLANE_STRAIN.
 * fcx_embed_3d writes the mapped components of grad_del_u [n][g][g] / stress [n][s]
 * into the persistent 3D scratch arrays grad3d [n][9] / stress3d [n][6] (other
 * components keep their values, as in the reference :285-292, :365-386);
 * fcx_extract_from_3d copies stress3d[:, :s] and the leading s x s block of
 * tangent3d [n][36] back (:293-302, :388-412). */
int fcx_embed_3d(int constraint, size_t n, const double *grad_del_u, const double *stress,
                 double *grad3d, double *stress3d, void *stream);
int fcx_extract_from_3d(int constraint, size_t n, const double *stress3d, const double *tangent3d,
                        double *stress, double *tangent, void *stream);

/* Companion gather: IncrementalDisplacement.evaluate_local_incremental_gradient,
 * reference solver/_incrementalunknowns.py:19-27,40-49:
 *   grad[c][q][i][j] = d(u - u_prev)_j / dx_i   (ufl.nabla_grad)
 * for affine cells, from precomputed tables:
 *   dphi_ref [nq][nd][gdim]  reference-element basis gradients at the QPs (DEVICE)
 *   Jinv     [ncells][gdim][gdim]  dX/dx per cell (DEVICE)
 *   dofmap   [ncells][nd]    node index of each local basis function (DEVICE, int32)
 *   u, u_prev  blocked nodal vectors [nnodes][gdim] (DEVICE); u_prev may be NULL.
 * Output grad [ncells*nq][gdim][gdim] is exactly the grad_del_u of the
 * evaluate calls above (cell-major QP order, reference tests/solver/test_maps.py:119-121). */
int fcx_gather_grad(int gdim, size_t ncells, int nq, int nd, const int *dofmap,
                    const double *u, const double *u_prev, const double *dphi_ref,
                    const double *Jinv, double *grad_del_u, void *stream);

/* du = u - u_prev for blocked nodal vectors of n doubles (DEVICE): the increment the gather differentiates
 * (reference solver/_incrementalunknowns.py:25-27, nabla_grad(u - u_prev)), formed ONCE per call as a nodal
 * vector.  fcx_gather_grad(u = du, u_prev = NULL) then gathers each nodal value once instead of twice:
 * bit-identical (the same subtraction, done before instead of after the fetch) and 0.12 instead of 0.16 ms
 * for 998 250 P2 tets. */
int fcx_nodal_increment(size_t n, const double *u, const double *u_prev, double *du, void *stream);

/* Fused form() pipeline for VonMises3D on affine P1/P2 tetrahedra -- one launch
 * for what LawOnSubMesh.evaluate does per Newton iteration (reference
 * solver/_lawonsubmesh.py:72-95 with an IdentityMap): gather grad_del_u
 * (_incrementalunknowns.py:40-49), history trial reset history_1 <- history_0
 * (_history.py:64-79), sigma_local <- stress.previous (_lawonsubmesh.py:58-61),
 * law.evaluate (:86-94), results -> stress.current / tangent (:63-70).
 * The committed arrays (stress_prev, eps_n0, alpha0) are only read, the trial
 * arrays (stress_cur, eps_n1, alpha1, tangent) only written; a pair may alias
 * (in-place).  grad_out: optional [ncells*nq][3][3] copy of grad_del_u.
 * (nd, nq) in {(10,4), (4,1), (4,4)}; FCX_ERR_ARG otherwise (use
 * fcx_gather_grad + fcx_mises_evaluate).
 * cells: NULL for a law that owns every cell (IdentityMap).  Otherwise the law's cell list
 *   (DEVICE int32 [ncells], reference solver/maps.py:62-123 SubSpaceMap / _lawonsubmesh.py:58-70):
 *   dofmap, Jinv, eps_n0/1, alpha0/1, grad_out and plastic_flag are indexed by the law's LOCAL cell
 *   number (sub-mesh arrays), stress_prev / stress_cur / tangent / tangent_rec are rows cells[c] of the
 *   PARENT arrays -- map_to_sub and map_to_parent happen inside the kernel's bulk copies.
 * tangent_rec: optional [ncells*nq][10], 16-byte aligned: the tangent of each point as
 *   {ka + cpp*2/3, ka - cpp/3, cpp, cnn, xn[6]} (see fcx_tangent_apply_rec). */
int fcx_mises_form(const double *params_host, size_t ncells, const int *cells, int nq, int nd,
                   const int *dofmap,
                   const double *u, const double *u_prev, const double *dphi_ref,
                   const double *Jinv, const double *stress_prev, double *stress_cur,
                   double *tangent, const double *eps_n0, double *eps_n1, const double *alpha0,
                   double *alpha1, double *grad_out, double *tangent_rec,
                   unsigned char *plastic_flag, int *status,
                   void *stream);

/* Residual and Jacobian action of IncrSmallStrainProblem on the device
 * (reference solver/_solver.py:87-101: R_form = inner(eps(v), sigma) dx,
 * dR_form = inner(eps(du), C eps(v)) dx, eps = ufl_mandel_strain,
 * solver/utils.py:10-62).  Affine simplex cells; tables as for fcx_gather_grad
 * plus weights [nq] (reference-cell quadrature weights) and detJ [ncells]
 * (|det dx/dX|).  Each call writes ELEMENT vectors fe [ncells][nd][fs] with
 * fs = fcx_fe_stride(gdim) doubles per (cell, local node) slot (3-D slots are
 * padded to 4 doubles = one 32-byte sector; the pad entry is written as 0):
 *   fcx_internal_force   fe = sum_q w|J| B_q^T stress_q          (stress [ncells*nq][sdim])
 *   fcx_tangent_apply    fe = sum_q w|J| B_q^T C_q^T B_q p_e     (p [nnodes][gdim], tangent [ncells*nq][sdim][sdim])
 *   fcx_tangent_diag     fe = diagonal of the element matrix
 * fcx_gather_sum then forms out[node][j] = beta*out + alpha * sum over the
 * node's (cell, local index) adjacency (adj_ptr [nnodes+1] int64, adj_idx
 * int32 = cell*nd + a) in a fixed order -- deterministic, no atomics.
 * fe_pos (optional, 3-D elements with 4 quadrature points only): [ncells][nd] int32, the slot of
 * (cell, local node) in the node-sorted adjacency (the inverse of adj_idx).  The element kernel
 * then writes fe in NODE-major slot order and fcx_gather_sum is called with adj_idx = NULL: it
 * reads each node's contributions as one contiguous run (HBM fetches 64 bytes per missed sector,
 * so the cell-major layout costs 2.7x the bytes on the read side, profiles/r1n).
 * fcx_tangent_apply_rec is fcx_tangent_apply for VonMises3D tangents given as the 10-double
 * records fcx_mises_form can emit (tangent_rec [ncells*nq][10] = ka+cpp*2/3, ka-cpp/3, cpp, cnn,
 * xn[6]; tangent = ka*xioi + cpp*xpp + cnn*outer(xn,xn), reference
 * models/mises_plasticity_isotropic_hardening.py:170-175): 80 instead of 288 bytes per QP per
 * Krylov iteration; gdim 3 only. */
int fcx_fe_stride(int gdim);
int fcx_internal_force(int gdim, int sdim, size_t ncells, int nq, int nd, const double *dphi_ref,
                       const double *weights, const double *Jinv, const double *detJ,
                       const double *stress, double *fe, const int *fe_pos, void *stream);
int fcx_tangent_apply(int gdim, int sdim, size_t ncells, int nq, int nd, const int *dofmap,
                      const double *p, const double *dphi_ref, const double *weights,
                      const double *Jinv, const double *detJ, const double *tangent, double *fe,
                      const int *fe_pos, void *stream);
int fcx_tangent_apply_rec(int gdim, int sdim, size_t ncells, int nq, int nd, const int *dofmap,
                          const double *p, const double *dphi_ref, const double *weights,
                          const double *Jinv, const double *detJ, const double *tangent_rec,
                          double *fe, const int *fe_pos, void *stream);
int fcx_tangent_diag(int gdim, int sdim, size_t ncells, int nq, int nd, const double *dphi_ref,
                     const double *weights, const double *Jinv, const double *detJ,
                     const double *tangent, double *fe, const int *fe_pos, void *stream);
int fcx_gather_sum(int gdim, size_t nnodes, const long long *adj_ptr, const int *adj_idx,
                   const double *fe, double *out, double alpha, double beta, void *stream);

/* SubSpaceMap.map_to_sub / map_to_parent on the device -- reference solver/maps.py:62-123 (the
 * index gather / scatter between the parent mesh's quadrature arrays and those of a law's sub-mesh,
 * called from solver/_lawonsubmesh.py:58-70).  A quadrature array is [cell][row_doubles] flat
 * (row_doubles = quadrature points per cell x entries per point, reference
 * tests/solver/test_maps.py:119-121); cells: DEVICE int32 [nrows_sub], parent cell of sub cell i.
 *   fcx_map_rows_to_sub     sub[i][:]           = parent[cells[i]][:]
 *   fcx_map_rows_to_parent  parent[cells[i]][:] = sub[i][:]      (other parent rows untouched) */
int fcx_map_rows_to_sub(size_t nrows_sub, size_t row_doubles, const int *cells, const double *parent,
                        double *sub, void *stream);
int fcx_map_rows_to_parent(size_t nrows_sub, size_t row_doubles, const int *cells, const double *sub,
                           double *parent, void *stream);

/* Fused vector kernels of one Jacobi-preconditioned CG iteration (the linear solve inside the
 * stand-in NewtonSolver; the reference leaves it to PETSc via dolfinx.nls.petsc.NewtonSolver).
 * All scalars are DEVICE doubles; reductions are deterministic (per-CTA partials summed in index
 * order by the last CTA).  minv = inverse Jacobi diagonal, 0 on constrained dofs (also the mask).
 * scratch: fcx_pcg_scratch_doubles() doubles; ticket: one zero-initialised unsigned.
 *   fcx_pcg_pap        out[0]  = sum_free p.Ap
 *   fcx_pcg_update_xr  alpha = rz/pAp (0 if pAp <= 0); x += alpha p; r -= alpha Ap on free dofs;
 *                      out2[0] = sum r.(minv r), out2[1] = sum r.r
 *   fcx_pcg_update_p   p = minv r + (rz_new/rz) p */
size_t fcx_pcg_scratch_doubles(void);
int fcx_pcg_pap(size_t n, const double *p, const double *Ap, const double *minv, double *scratch,
                unsigned *ticket, double *out, void *stream);
int fcx_pcg_update_xr(size_t n, double *x, double *r, const double *p, const double *Ap,
                      const double *minv, const double *rz, const double *pAp, double *scratch,
                      unsigned *ticket, double *out2, void *stream);
int fcx_pcg_update_p(size_t n, double *p, const double *r, const double *minv, const double *rz_new,
                     const double *rz, void *stream);

/* Device-resident Krylov loop (csrc/fcx_krylov.cu): single-reduction (Chronopoulos-Gear) Jacobi-PCG for the
 * linear solve of a Newton step -- in the reference PETSc's job behind dolfinx.nls.petsc.NewtonSolver, with
 * the ghost exchange and the dot-product reductions of the MPI-partitioned mesh
 * (reference solver/_solver.py:64-68, tests/solver/test_solver_mpi.py:93-121).  One process per GPU; the
 * reduction (an all-gather of three partial sums, added in rank order on every rank) and the ghost update
 * of the matvec input are peer-memory stores from inside the kernels over NVLink -- no NCCL, no host
 * round trip per iteration.  Call order:
 *   fcx_krylov_create    allocates the rank's vectors and its communication block; returns the block's
 *                        CUDA IPC handle (64 bytes) for the other ranks.  Local numbering: the nnodes_owned
 *                        owned nodes first, then the ghosts (ghost entries of the matvec input are written by
 *                        their owners only)
 *   fcx_krylov_connect   handles of all ranks (world x 64 bytes, rank order)
 *   fcx_krylov_set_halo  per neighbour: my local nodes to send and the neighbour's local index of each
 *   fcx_krylov_set_operator  the Jacobian action: arguments of fcx_tangent_apply_rec (mode 3) or
 *                        fcx_tangent_apply (mode 1) and the adjacency of fcx_gather_sum; the local cells
 *                        [0, ncells_interior) touch no ghost node -- the element kernel runs on them while the
 *                        neighbours' ghost stores arrive, then waits, then runs on the rest (pass ncells for
 *                        no split)
 *   fcx_krylov_begin     x = 0, r = rhs where minv != 0 (minv = inverse Jacobi diagonal, 0 on constrained
 *                        AND ghost dofs), first ghost push
 *   fcx_krylov_iterate   enqueue `iters` iterations (6 launches each, 3 on one rank); never synchronises
 *   fcx_krylov_status    (after a stream synchronisation) out[0] iterations done, out[1] r.r at the start
 *                        of the last one, out[2] r.r of the right-hand side, out[3] 1 = breakdown (p.Ap <= 0),
 *                        2 = a peer rank never arrived (bounded spin timed out)
 *   fcx_krylov_solution  x_out <- x
 *   fcx_krylov_set_tolerance  residual test ON THE DEVICE: with rtol > 0 the solve freezes itself once
 *                        r.r <= rtol^2 (r.r of iteration 0) -- the iteration that sees it leaves x untouched, every
 *                        later kernel of the loop returns at once (also after a breakdown) -- so the host may
 *                        enqueue blocks of iterations AHEAD of knowing the outcome instead of draining the
 *                        stream at every check; takes effect with the next fcx_krylov_begin; rtol <= 0 = never
 *   fcx_krylov_snapshot / fcx_krylov_wait_snapshot  enqueue a copy of the control block into pinned host slot
 *                        0 / 1 with an event behind it / wait for that event: out6 = {frozen, iteration at which
 *                        it froze, r.r at the start of the latest live iteration, r.r of iteration 0,
 *                        1 = breakdown / 2 = a peer never arrived, live iterations so far} */
int fcx_krylov_create(int rank, int world, int gdim, size_t nnodes, size_t nnodes_owned, void **handle_out,
                      void **comm_out, unsigned char *ipc_handle_out);
int fcx_krylov_connect(void *handle, const unsigned char *ipc_handles);
int fcx_krylov_set_halo(void *handle, int n_nbr, const int *nbr_rank, const int *send_ptr, const int *send_src,
                        const int *send_dst);
int fcx_krylov_set_operator(void *handle, int mode, int sdim, size_t ncells, int nq, int nd, const int *dofmap,
                            const double *dphi_ref, const double *weights, const double *Jinv, const double *detJ,
                            const double *tangent, double *fe, const int *fe_pos, const long long *adj_ptr,
                            const int *adj_idx, size_t ncells_interior);
int fcx_krylov_begin(void *handle, const double *rhs, const double *minv, void *stream);
int fcx_krylov_iterate(void *handle, int iters, void *stream);
int fcx_krylov_status(void *handle, double *out4);
int fcx_krylov_solution(void *handle, double *x_out, void *stream);
int fcx_krylov_set_tolerance(void *handle, double rtol);
int fcx_krylov_snapshot(void *handle, int slot, void *stream);
int fcx_krylov_wait_snapshot(void *handle, int slot, double *out6);
/* Ghost entries of the nodal vector x (owned nodes first) <- their owners' values, through the same peer-memory
 * push (reference: PETSc ghostUpdate / scatter_forward of the displacement, solver/_incrementalunknowns.py:36-38).
 * Collective over the ranks of the solver; enqueue-only; not during a solve. */
int fcx_krylov_halo_update(void *handle, double *x, void *stream);
void fcx_krylov_destroy(void *handle);

/* -------------------------------------------------------------------- host */

int fcx_elastic_evaluate_host(int constraint, const double *D, size_t n,
                              const double *grad_del_u, double *stress, double *tangent);

/* Returns >0 = number of non-converged QPs (see above). plastic_flag: optional HOST u8[n]. */
int fcx_mises_evaluate_host(const double *params, size_t n, const double *grad_del_u,
                            double *stress, double *tangent, double *eps_n, double *alpha,
                            unsigned char *plastic_flag);

int fcx_mises_linear_hardening_evaluate_host(const double *params, size_t n,
                                             const double *grad_del_u, double *stress,
                                             double *tangent, double *history,
                                             unsigned char *plastic_flag);

int fcx_drucker_prager_evaluate_host(int hyperbolic, const double *params, size_t n,
                                     const double *grad_del_u, double *stress, double *tangent,
                                     double *history, unsigned char *plastic_flag);

int fcx_kelvin_evaluate_host(int constraint, const double *D0, const double *I2, double mu0,
                             double lam0, double mu1, double tau, double del_t, size_t n,
                             const double *grad_del_u, double *stress, double *tangent,
                             double *strain_visco, double *strain);

int fcx_maxwell_evaluate_host(int constraint, const double *D0, const double *D1, double mu1,
                              double tau, double del_t, size_t n, const double *grad_del_u,
                              double *stress, double *tangent, double *strain_visco,
                              double *strain);

/* Page-lock / unlock a caller-owned host array so the *_host entry points DMA
 * it directly (cudaHostRegister).  A solver would call this once per
 * quadrature array at set-up (reference solver/_solver.py:75-85 is where those
 * arrays are created). */
int fcx_host_register(void *ptr, size_t bytes);
int fcx_host_unregister(void *ptr);
/* Pageable caller arrays (ordinary numpy memory) are staged through pinned ring
 * slots by a pool of host threads, both ways, overlapped with the DMA (memcpy
 * only -- no arithmetic on the CPU).  fcx_host_staging(0/1) switches that off/on
 * (off = let the driver stage; -1 = query), fcx_host_threads(n) sets the pool
 * size (0 = query; default min(8, cores - 2)).  Both return the old value. */
int fcx_host_staging(int on);
int fcx_host_threads(int n);
/* The plastic models' *_host entry points send their results over a download wire (stress for
 * every point, a flag byte, and for PLASTIC points only a compacted record; the host threads
 * scatter the records and copy the constant elastic tangent -- bit-identical arrays).
 * 1 = records carry tangent (VonMises3D: its 21 upper-triangle entries) + history;
 * 1 was the default of version 0.1 (3 = auto is now);
 * 2 = additionally, a page-locked caller tangent array gets the plastic tangents stored in place by
 * a kernel through its device alias, records carry the history only (fewer host-thread bytes, but
 * slower on the hosts measured so far, profiles/r1zf_host_wire_stats.jsonl);
 * 0 = plain D2H of every array, 3 = auto (below), -1 = query; returns the old value. */
int fcx_host_wire(int on);
/* 3 = AUTO, the default since 0.2: wire 1, except that stress-only calls take wire 0 and, with several
 * ranks per host (LOCAL_WORLD_SIZE >= FCX_WIRE_AUTO_RANKS, default 4) and every result array page-locked,
 * the call takes wire 2 (measured best at 4 and 8 ranks per host, profiles/r2f_e2e_sweep_pinned_n*.jsonl;
 * plain DMA, wire 0, is the slowest there: every variant is bound by the host's DRAM).
 * fcx_host_wire_used(): what the last plastic host call resolved to (0/1/2; -1 before the first call). */
int fcx_host_wire_used(void);
/* Mixed download of a record-wire call (wire 1) whose result arrays are ALL page-locked: `percent` of the chunks
 * leave by plain DMA straight into the caller's arrays (392 B per point over the link, no host-thread byte), the
 * others by records (161 B per point over the link, 392 B per point written by the pool threads) -- the link and
 * the threads work side by side on different chunks.  0 = off, 1..100, -1 = AUTO (default: FCX_WIRE_MIX_AUTO
 * percent with one rank per host, 0 with several); values below -1 only query.  Returns the old setting.
 * Bit-identical arrays either way (tests/test_gpu_parity.py::test_host_path_memory_kinds[mixed_wire-*]).
 * Measured on this pool's 16-core hosts it is SLOWER at every share (158 -> 122 M QP/s from 0 to 100 %,
 * profiles/r2s_e2e_mix_pinned.jsonl), so AUTO is 0 unless FCX_WIRE_MIX_AUTO says otherwise.
 * fcx_host_wire_mix_used(): the share the last plastic host call ran with (0 if its arrays were not page-locked). */
int fcx_host_wire_mix(int percent);
int fcx_host_wire_mix_used(void);
/* NUMA placement of the host pipeline (pool threads, drain thread, pinned ring slots) on the node
 * the bound GPU hangs off; 1 = on (default), 0 = off, -1 = query; returns the old value.  A no-op on
 * single-node hosts and where sysfs hides the topology.  fcx_host_numa_info: out[0..4) = GPU's NUMA
 * node (-1 unknown), usable CPUs of that node, CPUs this process may use, 1 if threads are pinned. */
int fcx_host_numa(int on);
int fcx_host_numa_info(int *out, int n);
/* Host-side roofline probes for the e2e path, measured on this box (GB/s, best of a few passes):
 * fcx_diag_host_bandwidth: `threads` pool threads (0 = the pool's default) over two `bytes`-sized
 *   buffers: out[0] memcpy (bytes read + bytes written per second), out[1] streaming-store fill,
 *   out[2] read, out[3] threads used.
 * fcx_diag_pcie: page-locked <-> device copies of `bytes`: out[0] H2D alone, out[1] D2H alone,
 *   out[2] H2D and out[3] D2H while the other direction runs (two copy engines). */
int fcx_diag_host_bandwidth(int threads, size_t bytes, double *out, int nout);
int fcx_diag_pcie(size_t bytes, double *out, int nout);
/* Chunks in flight in the *_host pipelines (streams / device buffers / pinned ring slots):
 * 2..8, default 6; 0 = query.  Returns the old value. */
int fcx_host_slots(int n);
/* Where the wall time of the last staged / wire host call went: out[0..12) = total_s,
 * main_wait_slot_s, main_stage_in_s, main_enqueue_s, drain_event_wait_s, drain_expand_s,
 * gpu_h2d_s, gpu_kernel_s, gpu_pack_s, gpu_d2h_s (per-chunk event intervals summed over chunks;
 * recorded only after fcx_host_trace(1)), chunks, chunk_qps.  Returns 12. */
int fcx_host_stats(double *out, int n);
int fcx_host_trace(int on);
/* DIAGNOSTIC ONLY: leave phases of the staged host pipeline out (1 uploads, 2 wire kernels,
 * 4 download DMAs, 8 host expansion, 16 model kernel) to time the others; results are garbage
 * while any bit is set.  -1 = query. */
int fcx_host_debug_skip(int mask);
/* Per-chunk timeline of the last traced call, 11 doubles per chunk (seconds since the call began):
 * chunk, slot, host slot-acquired, host enqueued, gpu chain start, h2d done, kernel done, pack done,
 * d2h done, host drain woke, host expansion done.  Returns the number of chunks. */
int fcx_host_timeline(double *out, int max_rows);
/* QPs per pipeline chunk of the *_host entry points (default 1<<18); 0 = query. */
size_t fcx_host_chunk_qps(size_t new_value);
/* Release the streams / staging buffers cached by the *_host entry points. */
void fcx_host_release(void);

/* ------------------------------------------------------------- diagnostics */

/* Number of kernel launches issued by this library in this process. */
unsigned long long fcx_launch_count(void);
/* Tunables: "ctas_per_sm" (persistent-grid CTAs per SM, 0 = occupancy query),
 * "mises_nmax" (Newton iteration cap, reference value 100).  Returns the
 * previous value, or FCX_ERR_ARG for an unknown key. */
int fcx_tune(const char *key, int value);

/* Streaming kernel with the Mises kernel's read:write byte mix (176 B read,
 * 392 B written per QP-equivalent), perfectly coalesced: measures the practical
 * DRAM ceiling for that mix.  src holds >= 22*n doubles, dst >= 49*n doubles
 * (DEVICE).  Returns the QP-equivalents actually processed (<= n_qps). */
long long fcx_diag_stream_mix(const double *src, double *dst, size_t n_qps, void *stream);
/* fp64 peak of the current device from chains of independent DFMAs (no memory traffic), in TFLOP/s
 * (2 flops per DFMA): the denominator for any fp64-pipe utilisation that is quoted.  < 0 on error. */
double fcx_diag_dfma_peak(void);

#ifdef __cplusplus
}
#endif
#endif /* FCX_H */
