// fcx_models.cuh -- per-quadrature-point arithmetic of the four constitutive
// laws, written as Model policies for the tile pipeline (fcx_tile.cuh) and the
// uniaxial elementwise kernel.  Operation order follows the reference
// (file:line cited per block; fc = src/fenics_constitutive) so results track
// its numpy models to rounding; the library is compiled with -fmad=false so
// nvcc does not fuse a*b+c differently from numpy.
#pragma once
#include "fcx_tile.cuh"

namespace fcx {

// Python's `1 / 2**0.5` (fc/models/utils.py:202-204) = 0x1.6a09e667f3bccp-1,
// one ulp below the correctly rounded 1/sqrt(2).
__device__ __forceinline__ double shear_factor() { return 0x1.6a09e667f3bccp-1; }
// np.sqrt(2 / 3)
__host__ __device__ __forceinline__ double sqrt23() { return 0x1.a20bd700c2c3ep-1; }

// fc/models/utils.py:132-208 (strain_from_grad_u), one QP.
template <int S, int G>
__device__ __forceinline__ void mandel_strain(const double *g, double *e)
{
    if (G == 1) {
        e[0] = g[0];  // :153-156
    } else if (G == 2) {
        e[0] = g[0];  // :165-168 / :182-185
        e[1] = g[3];
        e[2] = 0.0;
        e[3] = shear_factor() * (g[1] + g[2]);
    } else {
        e[0] = g[0];  // :199-204
        e[1] = g[4];
        e[2] = g[8];
        e[3] = shear_factor() * (g[1] + g[3]);
        e[4] = shear_factor() * (g[2] + g[6]);
        e[5] = shear_factor() * (g[5] + g[7]);
    }
}

// y_j = sum_i x_i M[i][j]  (numpy `x @ M`), M in kernel-parameter space.
template <int S>
__device__ __forceinline__ void vec_mat(const double *x, const double *M, double *y)
{
#pragma unroll
    for (int j = 0; j < S; ++j) {
        double acc = 0.0;
#pragma unroll
        for (int i = 0; i < S; ++i)
            acc += x[i] * M[i * S + j];
        y[j] = acc;
    }
}

// Dense, coalesced store of a tile's tangent when every QP has the same s*s
// matrix (elastic, Kelvin, Maxwell): Dsm holds the matrix in shared memory;
// consecutive threads write consecutive 16-byte pairs of the [cnt][s*s] block.
template <int S>
__device__ __forceinline__ void store_const_tangent(const double *Dsm, double *tang, int cnt,
                                                    int tid, int nthreads, bool vec_ok)
{
    constexpr int SS = S * S;
    if (vec_ok) {
        constexpr int NP = SS / 2;  // 16-byte pairs per QP
        const int npairs = cnt * NP;
        for (int p = tid; p < npairs; p += nthreads) {
            const int ij = p % NP;
            const double2 v = reinterpret_cast<const double2 *>(Dsm)[ij];
            st_stream_v2(tang + 2 * (size_t)p, v.x, v.y);
        }
    } else {
        for (int p = tid; p < cnt * SS; p += nthreads)
            tang[p] = Dsm[p % SS];
    }
}

// ===========================================================================
// LinearElasticityModel -- fc/models/linear_elasticity_model.py:26-45
//   segments: 0 grad_del_u [g*g] (read)   1 stress [s] (in place)
// ===========================================================================
template <int S, int G>
struct ElasticModel {
    struct Params {
        double D[S * S];  // get_elastic_tangent (fc/models/utils.py:25-93), row-major
    };
    static constexpr __host__ __device__ int nseg() { return 2; }
    static constexpr __host__ __device__ int w(int k) { return k == 0 ? G * G : S; }
    static constexpr __host__ __device__ int off(int k) { return k == 0 ? 0 : G * G; }
    static constexpr __host__ __device__ int wsum() { return G * G + S; }
    static constexpr __host__ __device__ bool wr(int k) { return k == 1; }
    static constexpr __host__ __device__ bool soa(int) { return false; }
    static constexpr __host__ __device__ int sdim() { return S; }
    // aux = the s*s tangent repeated for CQ QPs: source block of the bulk tangent stores
    static constexpr __host__ __device__ int const_tangent_qps() { return 32; }
    static constexpr __host__ __device__ int aux_doubles(int) { return 32 * S * S; }
    static constexpr __host__ __device__ int min_ctas(int tile) { return 512 / tile; }
    static constexpr __host__ __device__ bool has_flag() { return false; }

    __device__ static void init_aux(const Params &p, double *aux, int tid, int nthreads)
    {
        for (int i = tid; i < 32 * S * S; i += nthreads)
            aux[i] = p.D[i % (S * S)];
    }

    // stress += strain_increment @ D   (:44)
    __device__ static __forceinline__ void update(const Params &p, const double *g, double *sig)
    {
        double e[S], de[S];
        mandel_strain<S, G>(g, e);
        vec_mat<S>(e, p.D, de);
#pragma unroll
        for (int k = 0; k < S; ++k)
            sig[k] += de[k];
    }

    template <class V>
    __device__ static __forceinline__ void qp(const Params &p, const V &v, double *, int, bool &,
                                              bool &)
    {
        double g[G * G], sig[S];
#pragma unroll
        for (int i = 0; i < G * G; ++i)
            g[i] = v.template ld<0>(i);
#pragma unroll
        for (int i = 0; i < S; ++i)
            sig[i] = v.template ld<1>(i);
        update(p, g, sig);
#pragma unroll
        for (int i = 0; i < S; ++i)
            v.template st<1>(i, sig[i]);
    }

    // tangent[:] = tile(D.flatten(), n)   (:45)
    __device__ static __forceinline__ void store_tangent(const Params &, const double *aux,
                                                         double *tang, int cnt, int tid,
                                                         int nthreads, bool vec_ok)
    {
        store_const_tangent<S>(aux, tang, cnt, tid, nthreads, vec_ok);
    }

    // uniaxial form: a[0] = grad, a[1] = stress
    __device__ static __forceinline__ void qp1(const Params &p, double *a, double &tang)
    {
        update(p, &a[0], &a[1]);
        tang = p.D[0];
    }
};

// ===========================================================================
// SpringKelvinModel -- fc/models/spring_kelvin_model.py:43-88
//   segments: 0 grad [g*g] (read)  1 stress [s]  2 strain_visco [s]  3 strain [s]
// ===========================================================================
template <int S, int G>
struct KelvinModel {
    struct Params {
        double D0[S * S];  // self.D_0 (:38)
        double Dt[S * S];  // (1 - mu0/(tau*mu1*factor)) * D_0   (:85)
        double I2[S];      // get_identity (:39)
        double inv_factor; // 1 / factor, factor = 1/dt + 1/tau + mu0/(tau*mu1)   (:73,:75-76)
        double c_sig;      // 1 / (tau*2*mu1)      (:78)
        double c_ev;       // 1 / tau              (:79)
        double c_e;        // mu0 / (tau*mu1)      (:80)
        double c_tr;       // lam0 / (tau*2*mu1)   (:81)
        double two_mu0;    // 2 * mu0              (:84)
    };
    static constexpr __host__ __device__ int nseg() { return 4; }
    static constexpr __host__ __device__ int w(int k) { return k == 0 ? G * G : S; }
    static constexpr __host__ __device__ int off(int k) { return k == 0 ? 0 : G * G + (k - 1) * S; }
    static constexpr __host__ __device__ int wsum() { return G * G + 3 * S; }
    static constexpr __host__ __device__ bool wr(int k) { return k >= 1; }
    static constexpr __host__ __device__ bool soa(int) { return false; }
    static constexpr __host__ __device__ int sdim() { return S; }
    // aux = the s*s tangent repeated for CQ QPs: source block of the bulk tangent stores
    static constexpr __host__ __device__ int const_tangent_qps() { return 32; }
    static constexpr __host__ __device__ int aux_doubles(int) { return 32 * S * S; }
    static constexpr __host__ __device__ int min_ctas(int tile) { return 512 / tile; }
    static constexpr __host__ __device__ bool has_flag() { return false; }

    __device__ static void init_aux(const Params &p, double *aux, int tid, int nthreads)
    {
        for (int i = tid; i < 32 * S * S; i += nthreads)
            aux[i] = p.Dt[i % (S * S)];
    }

    __device__ static __forceinline__ void update(const Params &p, const double *g, double *sig,
                                                  double *ev, double *et)
    {
        double e[S], eD[S];
        mandel_strain<S, G>(g, e);
        double tr = 0.0;  // np.sum(strain_increment[:, :geometric_dim], axis=1)   (:69-71)
#pragma unroll
        for (int k = 0; k < G; ++k)
            tr += e[k];
        vec_mat<S>(e, p.D0, eD);
        const double ctr = p.c_tr * tr;
#pragma unroll
        for (int k = 0; k < S; ++k) {
            // (:74-83)
            const double dv =
                p.inv_factor * (p.c_sig * sig[k] - p.c_ev * ev[k] + p.c_e * e[k] + ctr * p.I2[k]);
            sig[k] += eD[k] - p.two_mu0 * dv;  // (:84)
            ev[k] += dv;                       // (:87)
            et[k] += e[k];                     // (:88)
        }
    }

    template <class V>
    __device__ static __forceinline__ void qp(const Params &p, const V &v, double *, int, bool &,
                                              bool &)
    {
        double g[G * G], sig[S], ev[S], et[S];
#pragma unroll
        for (int i = 0; i < G * G; ++i)
            g[i] = v.template ld<0>(i);
#pragma unroll
        for (int i = 0; i < S; ++i) {
            sig[i] = v.template ld<1>(i);
            ev[i] = v.template ld<2>(i);
            et[i] = v.template ld<3>(i);
        }
        update(p, g, sig, ev, et);
#pragma unroll
        for (int i = 0; i < S; ++i) {
            v.template st<1>(i, sig[i]);
            v.template st<2>(i, ev[i]);
            v.template st<3>(i, et[i]);
        }
    }

    __device__ static __forceinline__ void store_tangent(const Params &, const double *aux,
                                                         double *tang, int cnt, int tid,
                                                         int nthreads, bool vec_ok)
    {
        store_const_tangent<S>(aux, tang, cnt, tid, nthreads, vec_ok);
    }

    __device__ static __forceinline__ void qp1(const Params &p, double *a, double &tang)
    {
        update(p, &a[0], &a[1], &a[2], &a[3]);
        tang = p.Dt[0];
    }
};

// ===========================================================================
// SpringMaxwellModel -- fc/models/spring_maxwell_model.py:40-88
//   segments as for Kelvin.
// ===========================================================================
template <int S, int G>
struct MaxwellModel {
    struct Params {
        double D1[S * S];   // self.D_1 (:37)
        double D01[S * S];  // D_0 + D_1 (:80)
        double Dt[S * S];   // D_0 + (1 - 1/(tau*factor)) * D_1   (:82)
        double inv_factor;  // 1 / factor, factor = 1/dt + 1/tau    (:70,:72-73)
        double c_tot;       // 1 / (tau*2*mu1)   (:75)
        double c_ev;        // 1 / tau           (:76)
        double two_mu1;     // 2 * mu1           (:80)
    };
    static constexpr __host__ __device__ int nseg() { return 4; }
    static constexpr __host__ __device__ int w(int k) { return k == 0 ? G * G : S; }
    static constexpr __host__ __device__ int off(int k) { return k == 0 ? 0 : G * G + (k - 1) * S; }
    static constexpr __host__ __device__ int wsum() { return G * G + 3 * S; }
    static constexpr __host__ __device__ bool wr(int k) { return k >= 1; }
    static constexpr __host__ __device__ bool soa(int) { return false; }
    static constexpr __host__ __device__ int sdim() { return S; }
    // aux = the s*s tangent repeated for CQ QPs: source block of the bulk tangent stores
    static constexpr __host__ __device__ int const_tangent_qps() { return 32; }
    static constexpr __host__ __device__ int aux_doubles(int) { return 32 * S * S; }
    static constexpr __host__ __device__ int min_ctas(int tile) { return 512 / tile; }
    static constexpr __host__ __device__ bool has_flag() { return false; }

    __device__ static void init_aux(const Params &p, double *aux, int tid, int nthreads)
    {
        for (int i = tid; i < 32 * S * S; i += nthreads)
            aux[i] = p.Dt[i % (S * S)];
    }

    __device__ static __forceinline__ void update(const Params &p, const double *g, double *sig,
                                                  double *ev, double *et)
    {
        double e[S], tot[S], totD[S], eD[S];
        mandel_strain<S, G>(g, e);
#pragma unroll
        for (int k = 0; k < S; ++k)
            tot[k] = p.c_tot * (et[k] + e[k]);  // (:69, :75)
        vec_mat<S>(tot, p.D1, totD);
        vec_mat<S>(e, p.D01, eD);
#pragma unroll
        for (int k = 0; k < S; ++k) {
            const double dv = p.inv_factor * (totD[k] - p.c_ev * ev[k]);  // (:71-78)
            sig[k] += eD[k] - p.two_mu1 * dv;                             // (:80-81)
            ev[k] += dv;                                                  // (:85)
            et[k] += e[k];                                                // (:86)
        }
    }

    template <class V>
    __device__ static __forceinline__ void qp(const Params &p, const V &v, double *, int, bool &,
                                              bool &)
    {
        double g[G * G], sig[S], ev[S], et[S];
#pragma unroll
        for (int i = 0; i < G * G; ++i)
            g[i] = v.template ld<0>(i);
#pragma unroll
        for (int i = 0; i < S; ++i) {
            sig[i] = v.template ld<1>(i);
            ev[i] = v.template ld<2>(i);
            et[i] = v.template ld<3>(i);
        }
        update(p, g, sig, ev, et);
#pragma unroll
        for (int i = 0; i < S; ++i) {
            v.template st<1>(i, sig[i]);
            v.template st<2>(i, ev[i]);
            v.template st<3>(i, et[i]);
        }
    }

    __device__ static __forceinline__ void store_tangent(const Params &, const double *aux,
                                                         double *tang, int cnt, int tid,
                                                         int nthreads, bool vec_ok)
    {
        store_const_tangent<S>(aux, tang, cnt, tid, nthreads, vec_ok);
    }

    __device__ static __forceinline__ void qp1(const Params &p, double *a, double &tang)
    {
        update(p, &a[0], &a[1], &a[2], &a[3]);
        tang = p.Dt[0];
    }
};

// ===========================================================================
// VonMises3D -- fc/models/mises_plasticity_isotropic_hardening.py:57-175
//   segments: 0 grad [9] (read)  1 stress [6]  2 eps_n [6] (AoS or SoA)  3 alpha [1]
//   per-QP tangent record in shared memory (REC doubles, odd stride):
//     [0] A = ka + cpp*xpp_diag  [1] B = ka + cpp*xpp_off  [2] cpp = 2mu(1 - 2mu*xc2)
//     [3] cnn = 4mu^2 (xc2 - xc1)   [4..9] xn
// ===========================================================================
struct MisesParams {
    double ka, mu, y0, y00, w;  // p_ka, p_mu, p_y0, p_y00, p_w   (:51-55)
    int nmax;                   // Newton iteration cap, 100 in the reference (:106)
};

// Scalar Newton for the plastic multiplier, :100-151.  Stays in registers.
// exp(-w*alpha) of the trial check is the first iterate's exp (gamma_0 = 0).
__device__ __forceinline__ void mises_return_map(const MisesParams &P, double sigtrn,
                                                 double alpha_n, double phitr, double exp0,
                                                 double &gamma, double &xg_final, bool &failed)
{
    const double c23 = sqrt23();
    const double two_mu = 2 * P.mu;
    const double dy = P.y00 - P.y0;
    const double dfc = (2.0 / 3.0) * dy * P.w;  // (2/3)*(y00-y0)*w   (:124-125)
    double gamma_0 = 1, gamma_1 = 0, xr = 1;
    int it = 0;
    // first pass of the loop (:129-139) with gamma_0 = 0: f(0) == phitr
    gamma_0 = gamma_1;
    it = 1;
    xr = phitr;
    double xg = -two_mu - dfc * exp0;
    gamma_1 = gamma_0 - xr / xg;
    while (fabs(xr) > 1e-12 && fabs(gamma_1 - gamma_0) > 1e-8 * fabs(gamma_1)) {
        gamma_0 = gamma_1;
        it = it + 1;
        const double ex = exp(-P.w * (alpha_n + c23 * gamma_0));
        xr = sigtrn - two_mu * gamma_0 - c23 * (P.y0 + dy * (1 - ex));  // f   (:111-121)
        xg = -two_mu - dfc * ex;                                        // df  (:123-126)
        gamma_1 = gamma_0 - xr / xg;
        if (it > P.nmax) {  // (:141-143)
            failed = true;
            break;
        }
    }
    xg_final = -two_mu - dfc * exp(-P.w * (alpha_n + c23 * gamma_1));  // (:147)
    gamma = gamma_1;
}

// One quadrature point of VonMises3D.evaluate (:75-175), registers only.
// In/out: sig (sigma_n -> sigma_{n+1}), ep (eps_n), alpha.  Out: the four
// tangent coefficients
//   coef[0] = ka + cpp*xpp_diag   coef[1] = ka + cpp*xpp_off   (volumetric 3x3 block)
//   coef[2] = cpp = 2mu(1 - 2mu*xc2)                            (shear diagonal)
//   coef[3] = cnn = 4mu^2 (xc2 - xc1)                           (weight of xn (x) xn)
// and the flow direction xn (zero for an elastic point), from which
// aah = ka*xioi + cpp*xpp + cnn*outer(xn, xn) (:170-175) is regenerated.
__device__ __forceinline__ void mises_point(const MisesParams &P, const double *g, double *sig,
                                            double *ep, double &alpha, double *coef, double *xn,
                                            bool &plastic, bool &failed)
{
    const double alpha_n = alpha;
    const double c23 = sqrt23();
    const double two_mu = 2 * P.mu;
    double eps[6];
    mandel_strain<6, 3>(g, eps);
    const double tr_eps = (eps[0] + eps[1]) + eps[2];  // :75
    const double tr_sig = (sig[0] + sig[1]) + sig[2];  // :81
    const double te3 = tr_eps / 3;                     // tr_eps * I2 / 3   (:76)
    const double ts3 = tr_sig / 3;                     // (:81)
    double del_sigtr[6], sigtr[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const double eps_dev = (k < 3) ? eps[k] - te3 : eps[k];       // :76
        del_sigtr[k] = two_mu * eps_dev;                              // :79
        const double stress_n_dev = (k < 3) ? sig[k] - ts3 : sig[k];  // :80-82
        sigtr[k] = stress_n_dev + del_sigtr[k];                       // :83-85
    }
    double dot = 0.0;
#pragma unroll
    for (int k = 0; k < 6; ++k)
        dot += sigtr[k] * sigtr[k];
    const double sigtrn = sqrt(dot);  // :88
    const double exp0 = exp(-P.w * alpha_n);
    const double phitr = sigtrn - c23 * (P.y0 + (P.y00 - P.y0) * (1 - exp0));  // :91-94

    double gamma_1 = 0, xc1 = 0, xc2 = 0;
    plastic = phitr > 0;  // :98
    if (plastic) {
        double xg;
        mises_return_map(P, sigtrn, alpha_n, phitr, exp0, gamma_1, xg, failed);
        const double inv_n = 1.0 / sigtrn;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            xn[k] = sigtr[k] * inv_n;  // flow direction (:108)
        xc1 = -1 / xg;                 // :150
        xc2 = gamma_1 * inv_n;         // :151
    } else {
#pragma unroll
        for (int k = 0; k < 6; ++k)
            xn[k] = 0.0;  // :154-158
    }
    const double ktr = P.ka * tr_eps;
    const double tmg = two_mu * gamma_1;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        ep[k] += gamma_1 * xn[k];                                                       // :161
        const double sh = ((k < 3) ? ktr + del_sigtr[k] : del_sigtr[k]) - tmg * xn[k];  // :165
        sig[k] += sh;                                                                   // :167
    }
    alpha = alpha_n + c23 * gamma_1;  // :162

    // ka*xioi + cpp*xpp takes four values per QP: coef[0] on the volumetric
    // diagonal, coef[1] on its off-diagonal, cpp on the shear diagonal, 0
    // elsewhere (xioi :33-42, xpp = I4 - (1/3) xioi :48).
    const double cpp = two_mu * (1 - two_mu * xc2);  // :172
    const double third = (1.0 / 3.0) * 1.0;
    coef[0] = P.ka + cpp * (1.0 - third);
    coef[1] = P.ka + cpp * (0.0 - third);
    coef[2] = cpp;
    coef[3] = 4 * P.mu * P.mu * (xc2 - xc1);  // :173
}

template <bool EPS_SOA>
struct MisesModel {
    using Params = MisesParams;
    static constexpr int REC = 11;  // 10 used; odd stride -> conflict-free 64-bit smem access
    static constexpr __host__ __device__ int nseg() { return 4; }
    static constexpr __host__ __device__ int w(int k) { return k == 0 ? 9 : (k == 3 ? 1 : 6); }
    static constexpr __host__ __device__ int off(int k) { return k == 0 ? 0 : (k == 1 ? 9 : (k == 2 ? 15 : 21)); }
    static constexpr __host__ __device__ int wsum() { return 22; }
    static constexpr __host__ __device__ bool wr(int k) { return k >= 1; }
    static constexpr __host__ __device__ bool soa(int k) { return EPS_SOA && k == 2; }
    static constexpr __host__ __device__ int sdim() { return 6; }
    static constexpr __host__ __device__ int const_tangent_qps() { return 0; }
    static constexpr __host__ __device__ int aux_doubles(int tile) { return REC * tile; }
    static constexpr __host__ __device__ int min_ctas(int tile) { return 512 / tile; }
    static constexpr __host__ __device__ bool has_flag() { return true; }

    __device__ static void init_aux(const Params &, double *, int, int) {}

    template <class V>
    __device__ static __forceinline__ void qp(const Params &P, const V &v, double *aux, int t,
                                              bool &plastic, bool &failed)
    {
        double g[9], sig[6], ep[6];
#pragma unroll
        for (int i = 0; i < 9; ++i)
            g[i] = v.template ld<0>(i);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            sig[i] = v.template ld<1>(i);
            ep[i] = v.template ld<2>(i);
        }
        double alpha = v.template ld<3>(0);
        double coef[4], xn[6];
        mises_point(P, g, sig, ep, alpha, coef, xn, plastic, failed);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            v.template st<1>(i, sig[i]);
            v.template st<2>(i, ep[i]);
        }
        v.template st<3>(0, alpha);
        if (aux == nullptr)  // stress-only instantiation: no tangent record
            return;
        double *rec = aux + t * REC;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            rec[k] = coef[k];
#pragma unroll
        for (int k = 0; k < 6; ++k)
            rec[4 + k] = xn[k];
    }

    // aah = ka*xioi + cpp*xpp + cnn*outer(xn, xn), row-major 36 per QP (:170-175),
    // generated pair-by-pair so that consecutive threads write consecutive
    // 16-byte chunks of the tile's [cnt][36] block.
    __device__ static __forceinline__ double entry(const double *rec, int i, int j)
    {
        const bool vol = (i < 3) && (j < 3);
        const bool diag = (i == j);
        const double base = vol ? (diag ? rec[0] : rec[1]) : (diag ? rec[2] : 0.0);
        return base + rec[3] * (rec[4 + i] * rec[4 + j]);
    }

    __device__ static __forceinline__ void store_tangent(const Params &, const double *aux,
                                                         double *tang, int cnt, int tid,
                                                         int nthreads, bool vec_ok)
    {
        if (vec_ok) {
            const int npairs = cnt * 18;
            for (int p = tid; p < npairs; p += nthreads) {
                const int q = p / 18;
                const int pr = p - q * 18;
                const int i = pr / 3;
                const int j = 2 * (pr - 3 * i);
                const double *rec = aux + q * REC;
                st_stream_v2(tang + 2 * (size_t)p, entry(rec, i, j), entry(rec, i, j + 1));
            }
        } else {
            for (int p = tid; p < cnt * 36; p += nthreads) {
                const int q = p / 36;
                const int ij = p - q * 36;
                const int i = ij / 6, j = ij - 6 * i;
                tang[p] = entry(aux + q * REC, i, j);
            }
        }
    }
};

// ===========================================================================
// comfe-rs MisesPlasticity3D (linear isotropic hardening, closed-form radial
// return) -- comfe-rs/src/mises_plasticity.rs:58-126, exported by the reference
// as MisesPlasticityLinearHardening3D (fc/models/rust_models.py:144-161).
//   segments: 0 grad [9] (read)  1 stress [6]  2 history [7] = [alpha, plastic_strain[6]]
//   (ONE history array per QP, comfe-rs/src/mises_plasticity.rs:43-51)
// Mandel strain uses Rust's correctly rounded FRAC_1_SQRT_2 (mandel.rs:147).
// The tangent kappa*1(x)1 + 2mu*theta*P_dev + 2mu*theta_bar*n n^T has the same
// four-coefficient + direction structure as VonMises3D's, so the tangent
// record / cooperative store of MisesModel is reused.
// ===========================================================================
struct MisesLinParams {
    double mu, kappa, y_0, h;
};

struct MisesLinModel {
    using Params = MisesLinParams;
    static constexpr int REC = MisesModel<false>::REC;
    static constexpr __host__ __device__ int nseg() { return 3; }
    static constexpr __host__ __device__ int w(int k) { return k == 0 ? 9 : (k == 1 ? 6 : 7); }
    static constexpr __host__ __device__ int off(int k) { return k == 0 ? 0 : (k == 1 ? 9 : 15); }
    static constexpr __host__ __device__ int wsum() { return 22; }
    static constexpr __host__ __device__ bool wr(int k) { return k >= 1; }
    static constexpr __host__ __device__ bool soa(int) { return false; }
    static constexpr __host__ __device__ int sdim() { return 6; }
    static constexpr __host__ __device__ int const_tangent_qps() { return 0; }
    static constexpr __host__ __device__ int aux_doubles(int tile) { return REC * tile; }
    static constexpr __host__ __device__ int min_ctas(int tile) { return 512 / tile; }
    static constexpr __host__ __device__ bool has_flag() { return true; }

    __device__ static void init_aux(const Params &, double *, int, int) {}

    template <class V>
    __device__ static __forceinline__ void qp(const Params &P, const V &v, double *aux, int t,
                                              bool &plastic, bool &)
    {
        double g[9], sig[6], hist[7];
#pragma unroll
        for (int i = 0; i < 9; ++i)
            g[i] = v.template ld<0>(i);
#pragma unroll
        for (int i = 0; i < 6; ++i)
            sig[i] = v.template ld<1>(i);
#pragma unroll
        for (int i = 0; i < 7; ++i)
            hist[i] = v.template ld<2>(i);
        const double f = 0.70710678118654752440;  // f64::consts::FRAC_1_SQRT_2
        const double e[6] = {g[0], g[4], g[8], f * (g[1] + g[3]), f * (g[2] + g[6]), f * (g[5] + g[7])};
        const double alpha = hist[0];
        const double p_0 = ((sig[0] + sig[1]) + sig[2]) / 3.0;   // vol_dev (mandel.rs:53-60)
        const double eps_trace = (e[0] + e[1]) + e[2];           // trace_dev (mandel.rs:62-68)
        const double ev = eps_trace / 3.0;
        const double p_1 = p_0 + P.kappa * eps_trace;            // :86
        const double two_mu = 2. * P.mu;
        double s_tr[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double s0 = (k < 3) ? sig[k] + (-p_0) : sig[k];
            const double ed = (k < 3) ? e[k] + (-ev) : e[k];
            s_tr[k] = s0 + two_mu * ed;  // :88
        }
        const double tv = ((s_tr[0] + s_tr[1]) + s_tr[2]) / 3.0;  // mises_norm (mandel.rs:19-34)
        double nsq = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double d = (k < 3) ? s_tr[k] + (-tv) : s_tr[k];
            nsq += d * d;
        }
        const double s_tr_eq = sqrt(3.0 * (0.5 * nsq));  // :89
        const double sigma_y = P.y_0 + P.h * alpha;      // :91
        double theta = 1.0, theta_bar = 0.0, nn[6];
        plastic = !(s_tr_eq < sigma_y);  // :94
        if (plastic) {
            const double del_alpha = (s_tr_eq - sigma_y) / (3. * P.mu + P.h);  // :104
            const double del_gamma = sqrt(3. / 2.) * del_alpha;                // :105
            theta = 1. - (3. * P.mu * del_alpha) / s_tr_eq;                    // :106
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                nn[k] = s_tr[k] / s_tr_eq;       // :110
                hist[1 + k] += del_gamma * nn[k];  // :111
            }
            hist[0] += del_alpha;                                                // :112
            theta_bar = 1.0 / (1.0 + (P.h / (3.0 * P.mu))) - (1.0 - theta);      // :117
        } else {
#pragma unroll
            for (int k = 0; k < 6; ++k)
                nn[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < 6; ++k)
            sig[k] = ((k < 3) ? p_1 : 0.0) + (plastic ? theta * s_tr[k] : s_tr[k]);  // :96 / :114
#pragma unroll
        for (int i = 0; i < 6; ++i)
            v.template st<1>(i, sig[i]);
#pragma unroll
        for (int i = 0; i < 7; ++i)
            v.template st<2>(i, hist[i]);
        if (aux == nullptr)  // stress-only instantiation: no tangent record
            return;
        // kappa*1(x)1 + (2mu theta)*P_dev + (2mu theta_bar)*n n^T   (:98-100, :118-120)
        const double c = two_mu * theta;
        const double third = (1.0 * (1.0 / 3.0)) * -1.0;  // consts.rs:106-115
        double *rec = aux + t * REC;
        rec[0] = P.kappa + c * (1.0 + third);
        rec[1] = P.kappa + c * (0.0 + third);
        rec[2] = c;
        rec[3] = two_mu * theta_bar;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            rec[4 + k] = nn[k];
    }

    __device__ static __forceinline__ void store_tangent(const Params &, const double *aux,
                                                         double *tang, int cnt, int tid,
                                                         int nthreads, bool vec_ok)
    {
        MisesModel<false>::store_tangent(MisesParams{}, aux, tang, cnt, tid, nthreads, vec_ok);
    }
};

// ===========================================================================
// comfe-rs Drucker-Prager models behind the generic implicit return mapping
//   comfe-rs/src/plasticity/general.rs:105-266           (8x8 Newton on [sigma, del_lambda, kappa])
//   comfe-rs/src/plasticity/drucker_prager_classic.rs:75-108      f = sqrt(J2) + b I1 - a
//   comfe-rs/src/plasticity/drucker_prager_hyperbolic.rs:64-102   f = sqrt(J2 + d^2) + b I1 - a
// exported by the reference as DruckerPrager3D / DruckerPragerHyperbolic3D
// (fc/models/rust_models.py:96-141).  Same segments as MisesLinModel:
//   0 grad [9] (read)  1 stress [6]  2 history [7] = [alpha, plastic_strain[6]]
//
// The reference solves the dense 8x8 system with LU every Newton iteration.
// For both models that system has closed structure: with s = dev(sigma),
// c1 = df/dJ2, c2 = d2f/dJ2^2,
//   dg/dsigma   = c1 P_dev + c2 s s^T                      (:98 / :87)
//   A := I + dl C dg/dsigma = I + alpha P_dev - beta s s^T,  alpha = 2 mu dl c1, beta = -2 mu dl c2
//   A^-1 x      = x_vol + x_dev/(1+alpha) + beta (s.x) / ((1+alpha) den) s,   den = 1 + alpha - beta |s|^2
// (Sherman-Morrison), the del_lambda column is C g = 3 kappa b_flow 1 + 2 mu c1 s and
// the kappa row/column decouples (df/dkappa = dg/dkappa = dk/dkappa = 0, the
// struct defaults).  So each Newton step -- THE SAME step, iterate for iterate,
// stop rule for stop rule (:219-227) -- costs O(6) instead of an 8x8 LU, the
// consistent tangent inverse(dres)[0:6,0:6] C (:255-262) is
//   3 kappa P_vol + 2mu/(1+alpha) P_dev + 2 mu gamma s s^T - y2 w^T / (c.y2)
// and fits a 12-double record (6 coefficients + s) that the CTA expands into
// the dense [TILE][36] block exactly like the Mises kernels do.
// Quirk reproduced, not fixed: res_kappa = alpha_1 - alpha_0 - k has no
// del_lambda factor (:208) while its Jacobian row has (:60-67), so alpha grows
// by sqrt(2/3)|g| per plastic step.
// ===========================================================================
struct DruckerPragerParams {
    double mu, kappa, a, b, d2, b_flow;
    double apex;  // a / b
};

//
// VAR selects how the slow fp64 operations (division, square root: 20-30 instructions and ~60-80 cycles of
// dependent latency each on sm_100a) are spelled.  The kernel is latency-bound on exactly those chains (ncu,
// profiles/r1t_rs_ncu_full.json: 3 warps per scheduler, 2.8 cycles of fixed-latency stall per issue), so
//   VAR 0  the reference's spelling, operation for operation: x / 3.0, sqrt then 1 / r, one division per
//          quotient -- 51 slow operations for a plastic point with four Newton steps;
//   VAR 1  the same algorithm -- same iterates, same stop rule -- with x * (1/3), r = x * rsqrt(x), one
//          reciprocal of opa * den for both quotients of the step, reciprocals instead of repeated divisions
//          in the tangent and the commit: 24 slow operations, results within a few ulp of VAR 0 (the model's
//          tolerance against the restated Rust algorithm is 1e-10; tests/test_dp_host_harness.py checks both
//          variants on the host, tests/test_drucker_prager.py on the GPU).
template <bool HYP, int VAR = 1>
struct DruckerPragerModel {
    using Params = DruckerPragerParams;
    static constexpr bool FAST = VAR != 0;
    __host__ __device__ static __forceinline__ double third(double x) { return FAST ? x * (1.0 / 3.0) : x / 3.0; }
    // a * b + c: fused in VAR 1 (the library is compiled with -fmad=false so that the Python models' kernels
    // round like numpy; this model's tolerance is 1e-10 against a restated algorithm, and ncu counted 4 unfused
    // fp64 instructions for every fused one)
    __host__ __device__ static __forceinline__ double mad(double a, double b, double c)
    {
        return FAST ? fma(a, b, c) : a * b + c;
    }
    // r = sqrt(x), inv_r = 1 / r  (x >= 0; x == 0 gives r = 0, inv_r = inf like the reference's 1 / sqrt(0))
    __host__ __device__ static __forceinline__ void sqrt_and_inverse(double x, double &r, double &inv_r)
    {
        if (FAST) {
            inv_r = rsqrt(x);
            r = x > 0.0 ? x * inv_r : 0.0;
        } else {
            r = sqrt(x);
            inv_r = 1.0 / r;
        }
    }
    static constexpr int REC = 13;  // 12 used; odd stride -> conflict-free 64-bit smem access
    static constexpr __host__ __device__ int nseg() { return 3; }
    static constexpr __host__ __device__ int w(int k) { return k == 0 ? 9 : (k == 1 ? 6 : 7); }
    static constexpr __host__ __device__ int off(int k) { return k == 0 ? 0 : (k == 1 ? 9 : 15); }
    static constexpr __host__ __device__ int wsum() { return 22; }
    static constexpr __host__ __device__ bool wr(int k) { return k >= 1; }
    static constexpr __host__ __device__ bool soa(int) { return false; }
    static constexpr __host__ __device__ int sdim() { return 6; }
    static constexpr __host__ __device__ int const_tangent_qps() { return 0; }
    static constexpr __host__ __device__ int aux_doubles(int tile) { return REC * tile; }
    static constexpr __host__ __device__ int min_ctas(int tile) { return 384 / tile; }
    static constexpr __host__ __device__ bool has_flag() { return true; }

    // slot 12 of every record stays 0.0: the "no base term" operand of entry_fixed
    __device__ static void init_aux(const Params &, double *aux, int tid, int tile)
    {
        for (int t = tid; t < tile; t += blockDim.x)
            aux[t * REC + 12] = 0.0;
    }

    struct State {
        double s[6], nsq, f, c1, c2, gn, inv_gn;
        bool apex;
    };

    // set_model_state (classic :75-108, hyperbolic :64-102)
    __host__ __device__ static __forceinline__ void state(const Params &P, const double *sg, double bfe, State &S)
    {
        const double i_1 = (sg[0] + sg[1]) + sg[2];
        const double m = third(i_1);
#pragma unroll
        for (int k = 0; k < 6; ++k)
            S.s[k] = (k < 3) ? sg[k] + (-m) : sg[k];
        double nsq = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            nsq = mad(S.s[k], S.s[k], nsq);
        S.nsq = nsq;
        const double j_2 = 0.5 * nsq;
        // one division per state: c1 = 1/(2r), c2 = -1/(4 r^3) with r = sqrt(J2 [+ d^2])
        double r, inv_r;
        sqrt_and_inverse(HYP ? j_2 + P.d2 : j_2, r, inv_r);
        S.apex = HYP ? false : !(i_1 < P.apex);  // assert!(i_1 < a/b), classic :86
        S.f = mad(P.b, i_1, r) - P.a;
        S.c1 = 0.5 * inv_r;
        S.c2 = -0.25 * (inv_r * inv_r) * inv_r;
        double gsq = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double gk = mad(S.c1, S.s[k], (k < 3) ? bfe : 0.0);
            gsq = mad(gk, gk, gsq);
        }
        if (FAST)
            sqrt_and_inverse(gsq, S.gn, S.inv_gn);
        else
            S.gn = sqrt(gsq);
    }

    __host__ __device__ static __forceinline__ double normsq6(const double *x)
    {
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            acc = fma(x[k], x[k], acc);
        return acc;
    }

    // Two-phase update (fcx_tile.cuh TwoPhase): `trial` classifies the point and finishes it if it is elastic
    // (or fails the apex assert); returns true if the return mapping has to run -- then `qp` below is called
    // for the point by whichever thread picks it from the CTA's list (it repeats the trial arithmetic, so
    // the classification and every bit of the result are those of the single-phase update).
    // MEASURED and switched OFF (profiles/r2g_models.json, 16 M points, 52 % plastic, 425 GPU tests green with it
    // on): classic 2.85 ms against 2.6 ms single-phase, hyperbolic 3.29 against 3.3.  With 128 points per tile
    // the list fills two warps and a few lanes of a third, so only one warp of four sits the return mapping
    // out, and that is paid for with two more CTA barriers per tile and the repeated trial arithmetic.
    static constexpr __host__ __device__ bool two_phase() { return false; }

    template <class V>
    __host__ __device__ static __forceinline__ bool trial(const Params &P, const V &v, double *aux, int t,
                                                 bool &plastic, bool &failed)
    {
        double g[9], sig0[6];
#pragma unroll
        for (int i = 0; i < 9; ++i)
            g[i] = v.template ld<0>(i);
#pragma unroll
        for (int i = 0; i < 6; ++i)
            sig0[i] = v.template ld<1>(i);
        const double fr = 0.70710678118654752440;
        const double e[6] = {g[0], g[4], g[8], fr * (g[1] + g[3]), fr * (g[2] + g[6]), fr * (g[5] + g[7])};
        const double two_mu = 2.0 * P.mu, k3 = 3.0 * P.kappa;
        const double bfe = (P.b == P.b_flow) ? P.b : P.b_flow;
        const double etr = (e[0] + e[1]) + e[2];
        double sigtr[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double ev = (k < 3) ? third(etr) : 0.0;
            sigtr[k] = mad(two_mu, e[k] - ev, k3 * ev) + sig0[k];
        }
        State S;
        state(P, sigtr, bfe, S);
        failed = S.apex;
        plastic = !(S.f <= 0.0);
        if (plastic && !failed)
            return true;
        if (aux != nullptr) {  // elastic tangent 3 kappa P_vol + 2 mu P_dev (also what a failed point reports)
            double *rec = aux + t * REC;
            rec[0] = P.kappa + two_mu * (2.0 / 3.0);
            rec[1] = P.kappa - third(two_mu);
            rec[2] = two_mu;
#pragma unroll
            for (int k = 3; k < 12; ++k)
                rec[k] = 0.0;
        }
        if (!failed) {
#pragma unroll
            for (int i = 0; i < 6; ++i)
                v.template st<1>(i, sigtr[i]);
        }
        return false;
    }

    template <class V>
    __host__ __device__ static __forceinline__ void qp(const Params &P, const V &v, double *aux, int t,
                                              bool &plastic, bool &failed)
    {
        double g[9], sig0[6], hist[7];
#pragma unroll
        for (int i = 0; i < 9; ++i)
            g[i] = v.template ld<0>(i);
#pragma unroll
        for (int i = 0; i < 6; ++i)
            sig0[i] = v.template ld<1>(i);
#pragma unroll
        for (int i = 0; i < 7; ++i)
            hist[i] = v.template ld<2>(i);
        const double fr = 0.70710678118654752440;  // f64::consts::FRAC_1_SQRT_2 (mandel.rs:147)
        const double e[6] = {g[0], g[4], g[8], fr * (g[1] + g[3]), fr * (g[2] + g[6]), fr * (g[5] + g[7])};
        const double two_mu = 2.0 * P.mu, k3 = 3.0 * P.kappa;
        const double bfe = (P.b == P.b_flow) ? P.b : P.b_flow;  // associated / non-associated flow
        const double c23 = sqrt23();
        // sigma_tr = C de + sigma_0   (general.rs:126)
        const double etr = (e[0] + e[1]) + e[2];
        double sol[6], sigtr[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double ev = (k < 3) ? third(etr) : 0.0;
            sigtr[k] = mad(two_mu, e[k] - ev, k3 * ev) + sig0[k];
            sol[k] = sigtr[k];
        }
        State S;
        state(P, sigtr, bfe, S);
        failed = S.apex;
        plastic = !(S.f <= 0.0);  // :133
        // the tangent record is built in registers and leaves for shared memory at every exit
        // (stress-only instantiation, aux == nullptr: dead code)
        double rec[REC];
        auto flush = [&]() {
            if (aux != nullptr) {
#pragma unroll
                for (int k = 0; k < 12; ++k)
                    aux[t * REC + k] = rec[k];
            }
        };
        rec[0] = P.kappa + two_mu * (2.0 / 3.0);  // elastic tangent 3 kappa P_vol + 2 mu P_dev
        rec[1] = P.kappa - third(two_mu);          // (also what a failed point reports)
        rec[2] = two_mu;
#pragma unroll
        for (int k = 3; k < 12; ++k)
            rec[k] = 0.0;
        if (!plastic || failed) {
            if (!failed) {
#pragma unroll
                for (int i = 0; i < 6; ++i)
                    v.template st<1>(i, sigtr[i]);
            }
            flush();
            return;
        }
        const double alpha_0 = hist[0];
        double dl = 0.0, al = alpha_0;
        double rs[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0}, rf = S.f, rk = 0.0;
        const double atol = 1e-8, rtol = 1e-8;
        double opa = 1.0, den = 1.0, beta = 0.0;
        for (int it = 0;; ++it) {
            // ---- one Newton step with the matrix of the current iterate (:176-190) ----
            const double alpha = two_mu * dl * S.c1;
            beta = -(two_mu * dl * S.c2);
            opa = 1.0 + alpha;
            den = mad(-beta, S.nsq, opa);
            const double tr = ((rs[0] + rs[1]) + rs[2]) * (1.0 / 3.0);
            double sdot = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k)
                sdot = fma(S.s[k], rs[k], sdot);
            double inv_opa, inv_den;
            if (FAST) {  // one reciprocal for both
                const double t = 1.0 / (opa * den);
                inv_opa = den * t;
                inv_den = opa * t;
            } else {
                inv_opa = 1.0 / opa;
                inv_den = 1.0 / den;
            }
            const double cf = beta * sdot * (inv_opa * inv_den);
            const double h = two_mu * S.c1 * inv_den, q1 = k3 * bfe;
            double y1[6], y2[6];
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const double vol = (k < 3) ? tr : 0.0;
                y1[k] = fma(cf, S.s[k], fma(rs[k] - vol, inv_opa, vol));
                y2[k] = fma(h, S.s[k], (k < 3) ? q1 : 0.0);
            }
            double sy1 = 0.0, sy2 = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                sy1 = fma(S.s[k], y1[k], sy1);
                sy2 = fma(S.s[k], y2[k], sy2);
            }
            const double cy1 = mad(S.c1, sy1, P.b * ((y1[0] + y1[1]) + y1[2]));
            const double cy2 = mad(S.c1, sy2, P.b * ((y2[0] + y2[1]) + y2[2]));
            const double dlam = (cy1 - rf) / cy2;
            double dsig[6], sds = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                dsig[k] = fma(-y2[k], dlam, y1[k]);
                sds = fma(S.s[k], dsig[k], sds);
            }
            const double kk = c23 * S.gn;
            const double dkc = (FAST ? c23 * S.inv_gn : c23 / S.gn) * S.c1 * mad(S.c2, S.nsq, S.c1);  // dk/dsigma = dkc s
            const double dkap = mad(kk, dlam, mad(dl, dkc * sds, rk));
#pragma unroll
            for (int k = 0; k < 6; ++k)
                sol[k] -= dsig[k];
            dl -= dlam;
            al -= dkap;
            // ---- new state and residual (:200-217) ----
            state(P, sol, bfe, S);
            if (S.apex || !(cy2 != 0.0)) {
                failed = true;
                break;
            }
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const double cg = mad(two_mu * S.c1, S.s[k], (k < 3) ? k3 * bfe : 0.0);
                rs[k] = mad(dl, cg, sol[k] - sigtr[k]);
            }
            rk = mad(-c23, S.gn, al - alpha_0);
            rf = S.f;
            // |x| < t  <=>  x.x < t^2 for t > 0: two of the three square roots go
            const bool conv_res = normsq6(rs) < atol * atol && fabs(rk) < atol && fabs(rf) < atol;
            bool conv_inc = fabs(dkap) < atol + rtol * fabs(al) && fabs(dlam) < atol + rtol * fabs(dl);
            if (!FAST || (conv_inc && !conv_res)) {  // FAST: the square root only when it decides (same outcome)
                const double tinc = atol + rtol * sqrt(normsq6(sol));
                conv_inc = conv_inc && normsq6(dsig) < tinc * tinc;
            }
            if (conv_res || conv_inc)
                break;
            if (it > 25) {  // maxit, :167,:228
                failed = true;
                break;
            }
        }
        if (failed) {
            flush();
            return;  // the Rust code panics; nothing is written for this point
        }
        // ---- commit: stress, alpha, plastic strain += de - C^-1 (sigma_1 - sigma_0)  (:249-252) ----
        double x[6];
#pragma unroll
        for (int k = 0; k < 6; ++k)
            x[k] = sol[k] - sig0[k];
        const double xv = third((x[0] + x[1]) + x[2]);
        hist[0] = al;
        const double i2mu = FAST ? 1.0 / two_mu : 0.0, ik3 = FAST ? 1.0 / k3 : 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const double vol = (k < 3) ? xv : 0.0;
            hist[1 + k] += e[k] - (FAST ? (x[k] - vol) * i2mu + vol * ik3 : (x[k] - vol) / two_mu + vol / k3);
        }
#pragma unroll
        for (int i = 0; i < 6; ++i)
            v.template st<1>(i, sol[i]);
#pragma unroll
        for (int i = 0; i < 7; ++i)
            v.template st<2>(i, hist[i]);
        // ---- consistent tangent at the converged state (:255-262) ----
        {
            const double alpha = two_mu * dl * S.c1;
            beta = -(two_mu * dl * S.c2);
            opa = 1.0 + alpha;
            den = mad(-beta, S.nsq, opa);
            const double q1 = k3 * bfe, q2 = k3 * P.b;
            if (FAST) {  // two reciprocals instead of eight divisions
                const double t = 1.0 / (opa * den);
                const double inv_opa = den * t, inv_den = opa * t;
                const double m1 = two_mu * inv_opa, m2 = two_mu * (beta * t);
                const double h = two_mu * S.c1 * inv_den;
                const double D = fma(S.c1, h * S.nsq, P.b * (3.0 * q1));
                const double iD = 1.0 / D;
                const double A0 = P.kappa - q1 * q2 * iD;
                rec[0] = A0 + m1 * (2.0 / 3.0);
                rec[1] = A0 - m1 * (1.0 / 3.0);
                rec[2] = m1;
                rec[3] = m2 - h * h * iD;
                rec[4] = -(q1 * h) * iD;
                rec[5] = -(h * q2) * iD;
            } else {
                const double m1 = two_mu / opa, m2 = two_mu * (beta / (opa * den));
                const double h = two_mu * S.c1 / den;
                const double D = P.b * (3.0 * q1) + S.c1 * (h * S.nsq);
                const double A0 = P.kappa - q1 * q2 / D;
                rec[0] = A0 + m1 * (2.0 / 3.0);
                rec[1] = A0 - m1 / 3.0;
                rec[2] = m1;
                rec[3] = m2 - h * h / D;
                rec[4] = -(q1 * h) / D;
                rec[5] = -(h * q2) / D;
            }
#pragma unroll
            for (int k = 0; k < 6; ++k)
                rec[6 + k] = S.s[k];
        }
        flush();
    }

    // M_ij = [vol block] + m1 delta_ij + A3 s_i s_j + A4 1_i s_j + A5 s_i 1_j
    // record: rec[0..2] = vol-block diagonal / off-diagonal / shear diagonal, rec[3..5] = A3, A4, A5,
    // rec[6..11] = s, rec[12] = 0.0
    __host__ __device__ static __forceinline__ int base_slot(int i, int j)
    {
        return (i < 3 && j < 3) ? (i == j ? 0 : 1) : (i == j ? 2 : 12);
    }
    // VAR 1: everything that depends on (i, j) only -- the record slot of the base term and the two 0/1 factors --
    // comes in as an argument, so a thread that owns a fixed (i, j) pays six loads and five fused operations
    __host__ __device__ static __forceinline__ double entry_fixed(const double *rec, int i, int j, int slot,
                                                                  double ci, double cj)
    {
        const double si = rec[6 + i];
        return fma(cj * rec[5], si, fma(rec[6 + j], fma(rec[3], si, ci * rec[4]), rec[slot]));
    }
    __host__ __device__ static __forceinline__ double entry(const double *rec, int i, int j)
    {
        if (FAST)
            return entry_fixed(rec, i, j, base_slot(i, j), i < 3 ? 1.0 : 0.0, j < 3 ? 1.0 : 0.0);
        const bool vi = i < 3, vj = j < 3, diag = (i == j);
        const double base = (vi && vj) ? (diag ? rec[0] : rec[1]) : (diag ? rec[2] : 0.0);
        const double si = rec[6 + i], sj = rec[6 + j];
        return base + rec[3] * (si * sj) + (vi ? rec[4] * sj : 0.0) + (vj ? rec[5] * si : 0.0);
    }

    __device__ static __forceinline__ void store_tangent(const Params &, const double *aux,
                                                         double *tang, int cnt, int tid,
                                                         int nthreads, bool vec_ok)
    {
        if (vec_ok && FAST) {
            // 18 consecutive threads write the 18 16-byte pairs of one point's 6x6 block, nthreads / 18 points
            // per round: a thread keeps ONE (row, column pair) for the whole kernel, so the index arithmetic and
            // the selects of entry() leave the loop (ncu on the p-strided loop below: 57 % of the kernel's
            // instructions, profiles/r2m_dp_ncu_full.json); stores stay consecutive 16-byte chunks
            const int qpr = nthreads / 18;
            const int ql = tid / 18, pr = tid - 18 * ql;
            if (ql < qpr) {
                const int i = pr / 3, j = 2 * (pr - 3 * i);
                const int b0 = base_slot(i, j), b1 = base_slot(i, j + 1);
                const double ci = i < 3 ? 1.0 : 0.0, cj0 = j < 3 ? 1.0 : 0.0, cj1 = j + 1 < 3 ? 1.0 : 0.0;
                for (int q = ql; q < cnt; q += qpr) {
                    const double *rec = aux + q * REC;
                    st_stream_v2(tang + (size_t)q * 36 + 2 * pr, entry_fixed(rec, i, j, b0, ci, cj0),
                                 entry_fixed(rec, i, j + 1, b1, ci, cj1));
                }
            }
        } else if (vec_ok) {
            const int npairs = cnt * 18;
            for (int p = tid; p < npairs; p += nthreads) {
                const int q = p / 18;
                const int pr = p - q * 18;
                const int i = pr / 3;
                const int j = 2 * (pr - 3 * i);
                const double *rec = aux + q * REC;
                st_stream_v2(tang + 2 * (size_t)p, entry(rec, i, j), entry(rec, i, j + 1));
            }
        } else {
            for (int p = tid; p < cnt * 36; p += nthreads) {
                const int q = p / 36;
                const int ij = p - q * 36;
                const int i = ij / 6, j = ij - 6 * i;
                tang[p] = entry(aux + q * REC, i, j);
            }
        }
    }
};

}  // namespace fcx
