// dp_host.cu -- TEST INFRASTRUCTURE ONLY (never linked into libfcx.so, never imported by the package).
// Host instantiation of the PRODUCT's per-point Drucker-Prager arithmetic (csrc/fcx_models.cuh,
// DruckerPragerModel<HYP>::qp / entry, the functions the CUDA tile kernel calls for every quadrature point),
// so that `-m "not gpu"` tests can compare the kernel's own source -- Newton step, stop rule, tangent
// record and its expansion -- with the C restatement of comfe-rs/src/plasticity/general.rs:105-266 in
// oracle/fcx_oracle.c without a GPU.  Host sqrt / division are IEEE like the device's; only rsqrt differs
// in its last bit, well inside the 1e-10 tolerance of the model.
#include <cstddef>

#include "../../fenics_constitutive_b200/csrc/fcx_models.cuh"

namespace {
struct HostView {
    const double *g;
    double *sig, *hist;
    template <int K>
    __host__ __device__ double ld(int i) const
    {
        return K == 0 ? g[i] : (K == 1 ? sig[i] : hist[i]);
    }
    template <int K>
    __host__ __device__ void st(int i, double v) const
    {
        if (K == 1)
            sig[i] = v;
        else if (K == 2)
            hist[i] = v;
    }
};

template <bool HYP, int VAR>
int run(const fcx::DruckerPragerParams &P, size_t n, const double *grad, double *stress, double *tangent,
        double *hist, unsigned char *flag)
{
    using M = fcx::DruckerPragerModel<HYP, VAR>;
    int nfail = 0;
    for (size_t q = 0; q < n; ++q) {
        double rec[M::REC];
        rec[12] = 0.0;  // what init_aux leaves in every record of the tile kernel
        bool plastic = false, failed = false;
        HostView v{grad + 9 * q, stress + 6 * q, hist + 7 * q};
        M::qp(P, v, tangent != nullptr ? rec : nullptr, 0, plastic, failed);
        if (flag != nullptr)
            flag[q] = plastic ? 1 : 0;
        nfail += failed ? 1 : 0;
        if (tangent != nullptr)
            for (int i = 0; i < 6; ++i)
                for (int j = 0; j < 6; ++j)
                    tangent[36 * q + 6 * i + j] = M::entry(rec, i, j);
    }
    return nfail;
}
}  // namespace

// params as in fcx_drucker_prager_evaluate (include/fcx.h): classic [mu, kappa, a, b, b_flow], hyperbolic
// [mu, kappa, a, b, d, b_flow]; returns the number of failed points
// variant: the model's VAR (0 = the reference's spelling of every division / square root, 1 = the shipped one)
extern "C" int dp_host_evaluate(int hyperbolic, int variant, const double *params, size_t n, const double *grad, double *stress,
                                double *tangent, double *hist, unsigned char *flag)
{
    fcx::DruckerPragerParams P;
    P.mu = params[0];
    P.kappa = params[1];
    P.a = params[2];
    P.b = params[3];
    P.d2 = hyperbolic ? params[4] * params[4] : 0.0;
    P.b_flow = hyperbolic ? params[5] : params[4];
    P.apex = P.a / P.b;
    if (variant == 0)
        return hyperbolic ? run<true, 0>(P, n, grad, stress, tangent, hist, flag)
                          : run<false, 0>(P, n, grad, stress, tangent, hist, flag);
    return hyperbolic ? run<true, 1>(P, n, grad, stress, tangent, hist, flag)
                      : run<false, 1>(P, n, grad, stress, tangent, hist, flag);
}
