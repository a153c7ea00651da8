/*
 * vb_nlls.cuh - non-linear least squares per voxel (--method=nlls, NLLSInferenceTechnique): the whole optimisation
 * of a voxel in one thread.
 *
 * Reference: inference_nlls.cc - DoCalculations :90-229 (start at the model's initial posterior means, optimise,
 * precision = J'J / mse with a 1e-6 floor on the diagonal, 1e-12 I when that cannot be inverted), the cost
 * function NLLSCF :232-293 (cf = |y - g(p)|^2 over the unmasked samples, grad = -2 J'(y - g), Gauss-Newton
 * hess = 2 J'J, J from LinearizedFwdModel::ReCentre).
 * The optimiser itself is MISCMATHS::nonlin (NL_LM) - FSL's miscmaths/nonlin.cpp, a dependency that is not part of
 * the reference tree. Its Levenberg(-Marquardt) driver is restated in oracle/vb_oracle.cc (nlls_levmar, with the
 * algorithm written out and what pins it: the reference's golden test/outdata_linear_nlls) and followed here:
 *      lambda = 0.1; a step is accepted when it lowers cf (lambda /= 10, stop when the relative drop is <= 1e-8),
 *      refused otherwise (lambda *= 10, the Hessian keeps its nudged diagonal and has the old nudge taken out,
 *      stop when lambda > 1e20); at most 200 accepted steps; Levenberg adds lambda to the diagonal (fabber's
 *      default), --lm scales it by 1 + lambda.
 *
 * Re-design: J is never stored. An accepted step costs one pass with the 2P+1 finite-difference evaluations
 * (J'J and J'r accumulated on the fly, as in the VB kernels) and every trial one plain evaluation pass for cf.
 */
#pragma once
#include "vb_voxelwise.cuh"

namespace fab
{
/* cf(p): sum of squared residuals over the unmasked samples, in the reference's plain multiply-add order */
template <class Model>
FAB_DEV double nlls_cf(const VbArgs &a, const typename Model::Ctx &mc, int v, const double (&c)[Model::P])
{
    constexpr int P = Model::P;
    double p0[P];
#pragma unroll
    for (int i = 0; i < P; i++)
        p0[i] = to_model(a.params[i].transform, c[i]);
    const float *yp = a.data + v;
    const size_t stride = (size_t)a.N;
    double s = 0.0;
#pragma unroll 1
    for (int t = 0; t < a.T; t++)
    {
        if (a.pattern && a.pattern[t] == FAB_PAT_MASKED)
            continue;
        const double d = (double)__ldg(yp + (size_t)t * stride) - Model::eval(mc, t, p0);
        s = __dadd_rn(s, __dmul_rn(d, d));
    }
    return s;
}

/* ReCentre about c fused with the Gauss-Newton sums over the unmasked samples: S.A = J'J, S.b = J'(y - g).
 * (ALL_SAMPLES: masked ones too - the final precision, see the kernel.)
 * Returns 0 or FABBER_VOX_NONFINITE_* (every sample is looked at, masked or not: ReCentre comes before MaskRows) */
template <class Model, bool ALL_SAMPLES = false>
FAB_DEV int nlls_jacobian(const VbArgs &a, const typename Model::Ctx &mc, int v, const double (&c)[Model::P],
    Stats<Model::P> &S)
{
    constexpr int P = Model::P;
    double p0[P], pp[P], pn[P], rden[P];
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        const char code = a.params[i].transform;
        double delta = c[i] * 1e-5;
        if (delta < 0)
            delta = -delta;
        if (delta < 1e-10)
            delta = 1e-10;
        const double c2 = c[i] + delta, c3 = c[i] - delta;
        p0[i] = to_model(code, c[i]);
        pp[i] = to_model(code, c2);
        pn[i] = to_model(code, c3);
        rden[i] = 1.0 / (c2 - c3);
    }
    S.zero();
    const float *yp = a.data + v;
    const size_t stride = (size_t)a.N;
    bool bad_g = false, bad_j = false;
#pragma unroll 1
    for (int t = 0; t < a.T; t++)
    {
        typename Model::Sample smp;
        Model::sample(mc, t, smp);
        double g, gp[P], gn[P], J[P];
        /* the library exponential, not the table one: cf() above evaluates the model with eval(), and the two must
         * agree to the last bit for the accept / refuse test of a step to mean anything near convergence */
        Model::template eval_fd<false>(mc, smp, p0, pp, pn, g, gp, gn);
        bad_g = bad_g || !finite_d(g);
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            J[i] = (gp[i] - gn[i]) * rden[i];
            bad_j = bad_j || !finite_d(J[i]);
        }
        if (!ALL_SAMPLES && a.pattern && a.pattern[t] == FAB_PAT_MASKED)
            continue;
        S.add((double)__ldg(yp + (size_t)t * stride) - g, J);
    }
    return bad_g ? FABBER_VOX_NONFINITE_OFFSET : (bad_j ? FABBER_VOX_NONFINITE_JACOBIAN : 0);
}

template <class Model>
__global__ void __launch_bounds__(VB_BLOCK) nlls_kernel(const __grid_constant__ VbArgs a)
{
    constexpr int P = Model::P;
    constexpr int NT = NTri<P>::value;
    extern __shared__ double smem[];
    Model::stage(a, smem);
    __syncthreads();
    const int v = a.v_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.v_end)
        return;
    const typename Model::Ctx mc = Model::make_ctx(a, smem);
    const size_t N = (size_t)a.N;

    /* initialFwdPosterior->means: the model's hard-coded posterior in Fabber space, or the fwd-inital-posterior
     * file; the same for every voxel - no InitVoxelPosterior here (inference_nlls.cc:66-83,141-143) */
    double p[P];
#pragma unroll
    for (int i = 0; i < P; i++)
        p[i] = a.nlls_have_start ? a.nlls_start[i] : to_fabber(a.params[i].transform, a.params[i].post_mean);

    int status = 0, niter = 0;
    const int maxiter = 200;
    double lambda = 0.1, olambda = 0.0;
    double cf = nlls_cf<Model>(a, mc, v, p);
    bool success = true;
    double H[NT], g[P];
    Stats<P> S;
    for (;;)
    {
        if (success && niter++ >= maxiter)
            break;
        if (success)
        {
            const int err = nlls_jacobian<Model>(a, mc, v, p, S);
            if (err)
            {
                status = err;
                break;
            }
#pragma unroll
            for (int i = 0; i < NT; i++)
                H[i] = 2 * S.A[i];
#pragma unroll
            for (int i = 0; i < P; i++)
                g[i] = -2 * S.b[i];
        }
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            if (a.nlls_lm)
                H[tri(i, i)] = ((1.0 + lambda) / (1.0 + olambda)) * H[tri(i, i)];
            else
                H[tri(i, i)] = H[tri(i, i)] + lambda - olambda;
        }
        double Hinv[NT], ld, trial[P], ncf = 0.0;
        const bool solved = ldl_inverse<P>(H, Hinv, ld, false);
        if (solved)
        {
            double step[P];
            symv<P>(Hinv, g, step);
#pragma unroll
            for (int i = 0; i < P; i++)
                trial[i] = p[i] + -step[i];
            ncf = nlls_cf<Model>(a, mc, v, trial);
        }
        if (solved && (success = (ncf < cf)))
        {
            olambda = 0.0;
#pragma unroll
            for (int i = 0; i < P; i++)
                p[i] = trial[i];
            lambda = lambda / 10.0;
            const bool conv = 2.0 * fabs(cf - ncf) <= 1e-8 * (fabs(cf) + fabs(ncf) + 2.0e-16);
            cf = ncf;
            if (conv)
                break;
        }
        else
        {
            success = false;
            olambda = lambda;
            lambda = 10.0 * lambda;
            if (lambda > 1e20)
                break;
        }
    }
    if (niter > maxiter)
        niter = maxiter;

    /* ---- the NLLS precision (inference_nlls.cc:168-196): (J'J) / mse, zero diagonal elements lifted to 1e-6 ---- */
    double cov[NT];
#pragma unroll
    for (int i = 0; i < NT; i++)
        cov[i] = 0.0;
    if (status == 0)
    {
        /* QUIRK KEPT: inference_nlls.cc:172 calls MaskRows(J, ..) and drops its result, so this J'J is over ALL
         * samples, masked ones included, while sqerr and the degrees of freedom leave them out */
        const int err = nlls_jacobian<Model, true>(a, mc, v, p, S);
        if (err)
            status = err;
        else
        {
            const double sqerr = nlls_cf<Model>(a, mc, v, p);
            const double mse = sqerr / (double)(a.n_unmasked - P);
            double prec[NT];
            bool fin = true;
#pragma unroll
            for (int i = 0; i < NT; i++)
                prec[i] = S.A[i] / mse;
#pragma unroll
            for (int i = 0; i < P; i++)
                if (prec[tri(i, i)] < 1e-6)
                    prec[tri(i, i)] = 1e-6;
#pragma unroll
            for (int i = 0; i < NT; i++)
                fin = fin && finite_d(prec[i]);
            double ld;
            if (!fin)
            {
                /* mse = 0/0 (no more samples than parameters) or a perfect fit: NEWMAT's inverse does not throw on NaN,
                 * the covariance is NaN and the voxel is not an error (test/test_inference.cc:79-105 runs this) */
#pragma unroll
                for (int i = 0; i < NT; i++)
                    cov[i] = __longlong_as_double(0x7ff8000000000000ll);
            }
            else if (!mvn_inverse<P>(prec, cov, ld, false))
            {
                /* "precision matrix is probably singular so set manually": precisions 1e-12 I (:212-221) */
                status = FABBER_VOX_SINGULAR;
#pragma unroll
                for (int i = 0; i < NT; i++)
                    cov[i] = 0.0;
#pragma unroll
                for (int i = 0; i < P; i++)
                    cov[tri(i, i)] = 1 / 1e-12;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < P; i++)
        a.mean[i * N + v] = p[i];
#pragma unroll
    for (int i = 0; i < NT; i++)
        a.cov[i * N + v] = cov[i];
    if (a.iterations)
        a.iterations[v] = niter;
    a.status[v] = status;
}

} // namespace fab
