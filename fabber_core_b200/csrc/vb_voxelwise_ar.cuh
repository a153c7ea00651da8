/*
 * vb_voxelwise_ar.cuh - non-spatial VB with AR(1) noise (Ar1cNoiseModel, num-echoes=1,
 * ar1-cross-terms=none): all iterations of a voxel in one thread.
 *
 * Reference: noisemodel_ar.cc - UpdateAlpha :447-528, UpdatePhi :530-556, UpdateTheta :558-634,
 * CalcFreeEnergy :643-747, Precalculate :749-769, HardcodedInitialDists :379-403, matrix cache :83-223.
 *
 * Re-design. The reference stores >= 5 dense T x T "alpha matrices" PER VOXEL (noisemodel_ar.h:38-39,
 * 1.6 MB at T = 200) and pushes T x T products through OperatorKLJ ~6 times an iteration. With one
 * echo those matrices are
 *      M00 = diag(0,1,..,1)    M20 = diag(1,..,1,0)    M10 = -1 on the two first off-diagonals
 * and the marginal is Q = M00 + E[a] M10 + E[a^2] M20, so every product the model needs is a lag-0 or
 * lag-1 sum. One pass over the time-series accumulates
 *      S0 = (sum J_t J_t', sum J_t r_t, sum r_t^2)                                  lag 0
 *      S1 = (sum J_t J_t+1' + J_t+1 J_t', sum J_t r_t+1 + J_t+1 r_t, sum r_t r_t+1)  lag 1 (symmetrised)
 *      first and last sample (J_0, r_0), (J_T-1, r_T-1)
 * from which J'MJ, J'Mr, r'Mr follow for M00 (S0 minus first), M20 (S0 minus last), M10 (-S1) and any
 * combination Q; k'Mk uses k = r + J d as in the white-noise kernel. O(T) work, 40 doubles of state.
 */
#pragma once
#include "vb_voxelwise.cuh"

namespace fab
{
template <int P> struct ArStats
{
    Stats<P> S0; /* lag 0 */
    Stats<P> S1; /* lag 1, symmetrised; rr holds sum r_t r_t+1 (NOT doubled) */
    double Jf[P], rf, Jl[P], rl;
};

/* (J'MJ, J'Mr, r'Mr) for M = w00 M00 + w10 M10 + w20 M20 */
template <int P> FAB_DEV void ar_combine(const ArStats<P> &S, double w00, double w10, double w20, Stats<P> &out)
{
#pragma unroll
    for (int i = 0; i < P; i++)
    {
#pragma unroll
        for (int j = 0; j <= i; j++)
        {
            const double m00 = S.S0.A[tri(i, j)] - S.Jf[i] * S.Jf[j];
            const double m20 = S.S0.A[tri(i, j)] - S.Jl[i] * S.Jl[j];
            out.A[tri(i, j)] = w00 * m00 - w10 * S.S1.A[tri(i, j)] + w20 * m20;
        }
        const double b00 = S.S0.b[i] - S.Jf[i] * S.rf;
        const double b20 = S.S0.b[i] - S.Jl[i] * S.rl;
        out.b[i] = w00 * b00 - w10 * S.S1.b[i] + w20 * b20;
    }
    const double r00 = S.S0.rr - S.rf * S.rf;
    const double r20 = S.S0.rr - S.rl * S.rl;
    out.rr = w00 * r00 - w10 * (2.0 * S.S1.rr) + w20 * r20;
}

/* k'Mk + tr(Sigma J'MJ) with k = r + J d  (OperatorKLJ, noisemodel_ar.cc:433-445) */
template <int P>
FAB_DEV double ar_klj(const Stats<P> &M, const double (&d)[P], const double (&Sig)[NTri<P>::value])
{
    double bd = 0.0;
#pragma unroll
    for (int j = 0; j < P; j++)
        bd += M.b[j] * d[j];
    return (M.rr + 2.0 * bd + quadform<P>(M.A, d)) + trace_prod<P>(Sig, M.A);
}

template <class Model, int FAST, bool BASIS>
FAB_DEV void recentre_loop_ar(const VbArgs &a, const typename Model::Ctx &mc, int v, const double (&p0)[Model::P],
    const double (&pp)[Model::P], const double (&pn)[Model::P], const double (&rden)[Model::P],
    ArStats<Model::P> &S, volatile double *first)
{
    constexpr int P = Model::P;
    const float *yp = a.data + v;
    const size_t stride = (size_t)a.N;
    /* three-deep software prefetch and unchecked arithmetic: see recentre_loop (vb_voxelwise.cuh); the lag-0
     * sums S0.rr / S0.A_ii carry every sample, recentre_stats_ar() inspects them after the pass */
    float q0 = __ldg(yp), q1 = 0.f, q2 = 0.f;
    if (1 < a.T)
        q1 = __ldg(yp + stride);
    if (2 < a.T)
        q2 = __ldg(yp + 2 * stride);
    double Jprev[P], rprev = 0.0, jscale[P];
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        Jprev[i] = 0.0;
        jscale[i] = (pp[i] - pn[i]) * rden[i]; /* BASIS: see recentre_loop (vb_voxelwise.cuh) */
    }
    typename Model::Sample smp;
    Model::sample(mc, 0, smp);
#pragma unroll 1
    for (int t = 0; t < a.T; t++)
    {
        const double y = (double)q0;
        q0 = q1;
        q1 = q2;
        if (t + 3 < a.T)
            q2 = __ldg(yp + (size_t)(t + 3) * stride);
        typename Model::Sample nxt;
        Model::sample(mc, t + 1, nxt);
        double g, J[P];
        if constexpr (BASIS)
        {
            Model::basis_row(mc, smp, p0, g, J);
#pragma unroll
            for (int i = 0; i < P; i++)
                J[i] = J[i] * jscale[i];
        }
        else
        {
            double gp[P], gn[P];
            Model::template eval_fd<(FAST != 0)>(mc, smp, p0, pp, pn, g, gp, gn);
#pragma unroll
            for (int i = 0; i < P; i++)
                J[i] = (gp[i] - gn[i]) * rden[i];
        }
        smp = nxt;
        const double r = y - g;
        S.S0.add(r, J);
        /* lag-1 terms: Jprev/rprev are zero at t == 0, so the first pass adds exact zeros */
        S.S1.rr = fma(rprev, r, S.S1.rr);
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            S.S1.b[i] = fma(Jprev[i], r, fma(J[i], rprev, S.S1.b[i]));
#pragma unroll
            for (int j = 0; j <= i; j++)
                S.S1.A[tri(i, j)] = fma(Jprev[i], J[j], fma(J[i], Jprev[j], S.S1.A[tri(i, j)]));
        }
        if (t == 0)
        {
            /* the first sample is only needed after the loop: parked in shared memory, not in registers */
#pragma unroll
            for (int i = 0; i < P; i++)
                first[i * VB_BLOCK] = J[i];
            first[P * VB_BLOCK] = r;
        }
#pragma unroll
        for (int i = 0; i < P; i++)
            Jprev[i] = J[i];
        rprev = r;
    }
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        S.Jl[i] = Jprev[i];
        S.Jf[i] = first[i * VB_BLOCK];
    }
    S.rl = rprev;
    S.rf = first[P * VB_BLOCK];
}

template <class Model>
FAB_DEV int recentre_stats_ar(const VbArgs &a, const typename Model::Ctx &mc, int v, const double (&c)[Model::P],
    ArStats<Model::P> &S, volatile double *first)
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
    S.S0.zero();
    S.S1.zero();
    bool bad_g = false, bad_j = false;
    const bool fast = Model::HAS_FAST && Model::fast_ok(mc, a.T, p0, pp, pn);
    bool basis = false; /* opt-in, and only for models that hand out their basis row */
    if constexpr (Model::LINEAR)
        basis = a.basis_jacobian != 0;
    if (basis)
    {
        if constexpr (Model::LINEAR)
            recentre_loop_ar<Model, 0, true>(a, mc, v, p0, pp, pn, rden, S, first);
    }
    else if (fast)
        recentre_loop_ar<Model, 1, false>(a, mc, v, p0, pp, pn, rden, S, first);
    else
        recentre_loop_ar<Model, 0, false>(a, mc, v, p0, pp, pn, rden, S, first);
    bool sums_finite = finite_d(S.S0.rr);
#pragma unroll
    for (int i = 0; i < P; i++)
        sums_finite = sums_finite && finite_d(S.S0.A[tri(i, i)]);
    if (!sums_finite)
    {
        if (fast)
            recentre_diagnose<Model, true>(a, mc, p0, pp, pn, rden, bad_g, bad_j);
        else
            recentre_diagnose<Model, false>(a, mc, p0, pp, pn, rden, bad_g, bad_j);
    }
    return bad_g ? FABBER_VOX_NONFINITE_OFFSET : (bad_j ? FABBER_VOX_NONFINITE_JACOBIAN : 0);
}

/* 2x2 symmetric inverse via the same LDL^T path as everything else */
FAB_DEV bool inv2(const double (&A)[3], double (&Inv)[3], double &logdet) { return ldl_inverse<2>(A, Inv, logdet); }

template <class Model> struct ArVoxel
{
    static constexpr int P = Model::P;
    static constexpr int NT = NTri<P>::value;
    /* theta posterior / prior */
    double m[P], Lam[NT], Sig[NT], m0[P], L0[P];
    double logdetLam;
    /* noise posterior: phi ~ Gamma(b, c); alpha ~ N(am, aprec^-1) (2 entries; the second is inert
     * without cross terms but is carried because it is part of the MVN output and of F) */
    double nb, nc, am[2], aprec[3];
    /* alpha marginal (Ar1cMatrixCache::Update :197-222): Q = M00 + qa M10 + qcp M20 */
    double qa, qcp;

    static constexpr int STASH_DOUBLES = 3 * P + 2 * NT + 1 + 2 + 2 + 3 + 2 + (P + 1); /* + first sample */
    static constexpr int SNAP_DOUBLES = 3 * P + NT + 2 + 2 + 3 + 2;

    template <bool WITH_SIG> FAB_DEV void put(volatile double *s) const
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            s[(k++) * VB_BLOCK] = m[i];
            s[(k++) * VB_BLOCK] = m0[i];
            s[(k++) * VB_BLOCK] = L0[i];
        }
#pragma unroll
        for (int i = 0; i < NT; i++)
        {
            s[(k++) * VB_BLOCK] = Lam[i];
            if (WITH_SIG)
                s[(k++) * VB_BLOCK] = Sig[i];
        }
        if (WITH_SIG)
            s[(k++) * VB_BLOCK] = logdetLam;
        s[(k++) * VB_BLOCK] = nb;
        s[(k++) * VB_BLOCK] = nc;
        s[(k++) * VB_BLOCK] = am[0];
        s[(k++) * VB_BLOCK] = am[1];
        s[(k++) * VB_BLOCK] = aprec[0];
        s[(k++) * VB_BLOCK] = aprec[1];
        s[(k++) * VB_BLOCK] = aprec[2];
        s[(k++) * VB_BLOCK] = qa;
        s[(k++) * VB_BLOCK] = qcp;
    }
    template <bool WITH_SIG> FAB_DEV void get(const volatile double *s)
    {
        int k = 0;
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            m[i] = s[(k++) * VB_BLOCK];
            m0[i] = s[(k++) * VB_BLOCK];
            L0[i] = s[(k++) * VB_BLOCK];
        }
#pragma unroll
        for (int i = 0; i < NT; i++)
        {
            Lam[i] = s[(k++) * VB_BLOCK];
            if (WITH_SIG)
                Sig[i] = s[(k++) * VB_BLOCK];
        }
        if (WITH_SIG)
            logdetLam = s[(k++) * VB_BLOCK];
        nb = s[(k++) * VB_BLOCK];
        nc = s[(k++) * VB_BLOCK];
        am[0] = s[(k++) * VB_BLOCK];
        am[1] = s[(k++) * VB_BLOCK];
        aprec[0] = s[(k++) * VB_BLOCK];
        aprec[1] = s[(k++) * VB_BLOCK];
        aprec[2] = s[(k++) * VB_BLOCK];
        qa = s[(k++) * VB_BLOCK];
        qcp = s[(k++) * VB_BLOCK];
    }

    /* Ar1cMatrixCache::Update: covarPlus = Cov(alpha) + alpha alpha'. Returns false if singular. */
    FAB_DEV bool update_marginal()
    {
        double acov[3], ld;
        if (!mvn_inverse<2>(aprec, acov, ld, false))
            return false;
        qa = am[0];
        qcp = acov[0] + am[0] * am[0];
        return true;
    }

    FAB_DEV double apply_prior(const VbArgs &a, int k, int v, int it)
    {
        const fabber_cuda_param &p = a.params[k];
        if (p.prior_type == 'A')
        {
            const double new_cov = m[k] * m[k] + Sig[tri(k, k)];
            if (it == 0)
            {
                L0[k] = 1.0 / p.prior_var;
                m0[k] = p.prior_mean;
            }
            else
                L0[k] = 1.0 / new_cov;
            const double b = 2 / new_cov;
            return -1.5 * (log(b) + digamma_fsl(0.5)) - 0.5 - gammaln(0.5) - 0.5 * log(b);
        }
        m0[k] = (p.prior_type == 'I') ? a.image_prior[k][v] : p.prior_mean;
        L0[k] = p.prior_prec;
        return 0.0;
    }

    /* noisemodel_ar.cc:558-610 (LMalpha is ignored by the AR model) */
    FAB_DEV bool update_theta(const ArStats<P> &S, const double (&c)[P], bool want_logdet)
    {
        const double w = nb * nc;
        Stats<P> Q;
        ar_combine<P>(S, w, w * qa, w * qcp, Q); /* X = phi_bar Q */
#pragma unroll
        for (int i = 0; i < NT; i++)
            Lam[i] = Q.A[i];
#pragma unroll
        for (int i = 0; i < P; i++)
            Lam[tri(i, i)] = L0[i] + Q.A[tri(i, i)];
        if (!mvn_inverse<P>(Lam, Sig, logdetLam, want_logdet))
            return false;
        double Ac[P], rhs[P];
        symv<P>(Q.A, c, Ac);
#pragma unroll
        for (int i = 0; i < P; i++)
            rhs[i] = (Q.b[i] + Ac[i]) + L0[i] * m0[i];
        symv<P>(Sig, rhs, m);
        return true;
    }

    /* UpdateAlpha then UpdatePhi (noisemodel_ar.cc:405-410). Returns 0 or a FABBER_VOX_* code. */
    FAB_DEV int update_noise(const VbArgs &a, const ArStats<P> &S, const double (&c)[P])
    {
        double d[P];
#pragma unroll
        for (int i = 0; i < P; i++)
            d[i] = c[i] - m[i];
        const double w = nb * nc;
        const double prior_prec = a.ar_alpha_prior_prec;
        Stats<P> M;
        /* alpha precisions(1,1) += phi_bar * OpKLJ(M20) */
        ar_combine<P>(S, 0.0, 0.0, 1.0, M);
        aprec[0] = prior_prec + w * ar_klj<P>(M, d, Sig);
        aprec[1] = 0.0;
        aprec[2] = prior_prec;
        if (!finite_d(aprec[0]))
            return FABBER_VOX_NONFINITE_F; /* "Non-finite values in alpha precisions" :484 */
        double acov[3], ld;
        if (!ldl_inverse<2>(aprec, acov, ld, false))
            return FABBER_VOX_SINGULAR;
        if (fmin(acov[0], acov[2]) < 0)
            return FABBER_VOX_AR_NEG_VARIANCE;
        /* means = Cov * (prior_prec * prior_means + [-0.5 phi_bar OpKLJ(M10), 0]); prior means are 0 */
        ar_combine<P>(S, 0.0, 1.0, 0.0, M);
        const double t0 = 0.0 + -0.5 * w * ar_klj<P>(M, d, Sig);
        double acov2[3];
        if (!mvn_inverse<2>(aprec, acov2, ld, false))
            return FABBER_VOX_SINGULAR;
        am[0] = acov2[0] * t0 + acov2[1] * 0.0;
        am[1] = acov2[1] * t0 + acov2[2] * 0.0;
        if (!update_marginal())
            return FABBER_VOX_SINGULAR;
        /* UpdatePhi with the new marginal */
        ar_combine<P>(S, 1.0, qa, qcp, M);
        const double tmp = ar_klj<P>(M, d, Sig);
        nb = 1 / (tmp * 0.5 + 1 / a.noise_prior_b[0]);
        nc = ((double)a.T - 1) * 0.5 + a.noise_prior_c[0];
        return 0;
    }

    /* noisemodel_ar.cc:643-747 with c == m */
    FAB_DEV double free_energy(const VbArgs &a, const ArStats<P> &S) const
    {
        const double log2pi = FAB_LOG_2PI;
        const double w = nb * nc;
        Stats<P> Q;
        ar_combine<P>(S, w, w * qa, w * qcp, Q);
        double acov[3], ldA;
        ldl_inverse<2>(aprec, acov, ldA);
        const double elAlpha = 0.5 * ldA - 0.5 * 2 * (log2pi + 1);
        const double elTheta = 0.5 * logdetLam - 0.5 * P * (log2pi + 1);
        const double si = nb, ci = nc, siP = a.noise_prior_b[0], ciP = a.noise_prior_c[0];
        const double dg = digamma_fsl(ci), lsi = log(si);
        const double elPhi = -gammaln(ci) - ci * lsi - ci + (ci - 1) * (dg + lsi);
        const double p0 = (dg + lsi) * (((double)a.T - 1) * 0.5 + ciP - 1);
        const double p9 = -2 * gammaln(ciP) - 2 * ciP * log(siP) - si * ci / siP;
        const double p1 = -log2pi * ((double)a.T - 1 + 0.5 * 2 + 0.5 * P);
        const double p2 = -0.5 * Q.rr - 0.5 * trace_prod<P>(Q.A, Sig);
        double q = 0.0, tr0 = 0.0;
        LogProd lp0;
        lp0.init();
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            lp0.mul(L0[i]);
            const double dm = m[i] - m0[i];
            q += dm * L0[i] * dm;
            tr0 += Sig[tri(i, i)] * L0[i];
        }
        const double pp = a.ar_alpha_prior_prec;
        const double p3 = 0.5 * lp0.value();
        const double p4 = -0.5 * q;
        const double p5 = -0.5 * tr0;
        const double p6 = 0.5 * (log(fabs(pp)) + log(fabs(pp)));
        const double p7 = -0.5 * (am[0] * pp * am[0] + am[1] * pp * am[1]);
        const double p8 = -0.5 * (acov[0] * pp + acov[2] * pp);
        double F = -elAlpha - elTheta - elPhi;
        F += p0;
        F += p1;
        F += p2;
        F += p3;
        F += p4;
        F += p5;
        F += p6;
        F += p7;
        F += p8;
        F += p9;
        return F;
    }
};

#ifndef FAB_AR_MIN_BLOCKS
#define FAB_AR_MIN_BLOCKS FAB_MIN_BLOCKS
#endif
template <class Model>
__global__ void __launch_bounds__(VB_BLOCK, FAB_AR_MIN_BLOCKS) vb_voxelwise_ar_kernel(const __grid_constant__ VbArgs a)
{
    constexpr int P = Model::P;
    constexpr int NT = NTri<P>::value;
    typedef ArVoxel<Model> Vox;
    extern __shared__ double smem[];
    Model::stage(a, smem);
    volatile double *park = smem + Model::smem_bytes(a.T) / sizeof(double) + threadIdx.x;
    volatile double *snap = park + Vox::STASH_DOUBLES * VB_BLOCK;
    __syncthreads();
    const int v = a.v_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.v_end)
        return;
    const typename Model::Ctx mc = Model::make_ctx(a, smem);
    const size_t N = (size_t)a.N;

    Vox X;
    int status = 0;
    double F = 1234.5678;
    int it = 0;

    /* ---- SetupPerVoxelDists ------------------------------------------------------------------ */
    if (a.init_mean)
    {
#pragma unroll
        for (int i = 0; i < P; i++)
            X.m[i] = a.init_mean[i * N + v];
#pragma unroll
        for (int i = 0; i < NT; i++)
            X.Sig[i] = a.init_cov[i * N + v];
        double ld;
        if (!mvn_inverse<P>(X.Sig, X.Lam, ld))
            status = FABBER_VOX_SINGULAR | FABBER_VOX_SETUP_FLAG;
        X.logdetLam = -ld;
    }
    else
    {
        double var[P];
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            X.m[i] = (a.params[i].prior_type == 'I') ? a.image_prior[i][v] : a.params[i].post_mean;
            var[i] = a.params[i].post_var;
        }
        Model::init_voxel(a, v, X.m);
#pragma unroll
        for (int i = 0; i < NT; i++)
        {
            X.Sig[i] = 0.0;
            X.Lam[i] = 0.0;
        }
        X.logdetLam = 0.0;
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            const char code = a.params[i].transform;
            X.m[i] = to_fabber(code, X.m[i]);
            const double fv = to_fabber_var(code, var[i]);
            X.Sig[tri(i, i)] = fv;
            X.Lam[tri(i, i)] = 1.0 / fv;
            X.logdetLam += log(fabs(X.Lam[tri(i, i)]));
        }
    }
    /* noise: hard-coded initial dists (noisemodel_ar.cc:379-403) or the restart values, then
     * Precalculate (:749-769): marginal from the initial alpha, c = c_prior + (T-1)/2 */
    if (a.init_noise)
    {
        X.nb = a.init_noise[0 * N + v];
        X.nc = a.init_noise[1 * N + v];
        X.am[0] = a.init_noise[2 * N + v];
        X.am[1] = a.init_noise[3 * N + v];
        X.aprec[0] = a.init_noise[4 * N + v];
        X.aprec[1] = a.init_noise[5 * N + v];
        X.aprec[2] = a.init_noise[6 * N + v];
    }
    else
    {
        X.nb = a.noise_post_b[0];
        X.nc = a.noise_post_c[0];
        X.am[0] = X.am[1] = 0.0;
        X.aprec[0] = X.aprec[2] = a.ar_alpha_prior_prec;
        X.aprec[1] = 0.0;
    }
    if (!X.update_marginal())
        status = FABBER_VOX_SINGULAR | FABBER_VOX_SETUP_FLAG;
    X.nc = a.noise_prior_c[0] + ((double)a.T - 1) * 0.5;
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        X.m0[i] = 0.0;
        X.L0[i] = 1.0;
    }

    enum
    {
        PH_SETUP,
        PH_ITER,
        PH_REVERT
    };
    ArStats<P> S;
    double c[P];
    Conv conv;
    conv.init(a.conv_type, a.max_iterations, a.fchange, a.max_trials);
    /* only trialmode / freduce keep a real snapshot (see vb_voxelwise.cuh); the launcher sizes the
     * shared-memory snapshot region accordingly */
    const bool use_snap = a.conv_type == FABBER_CONV_TRIALMODE || a.conv_type == FABBER_CONV_FREDUCE;
    double Fprior = 0.0;
    int phase = PH_SETUP;
    while (status == 0)
    {
#pragma unroll
        for (int i = 0; i < P; i++)
            c[i] = X.m[i];
        X.template put<true>(park);
        const int err = recentre_stats_ar<Model>(a, mc, v, c, S, park + (Vox::STASH_DOUBLES - (P + 1)) * VB_BLOCK);
        X.template get<true>(park);
        if (phase == PH_SETUP)
        {
            if (err)
            {
                status = err | FABBER_VOX_SETUP_FLAG;
                break;
            }
            if (use_snap)
                X.template put<false>(snap);
            phase = PH_ITER;
        }
        else
        {
            if (err)
            {
                status = err;
                break;
            }
            if (a.need_f)
            {
                F = X.free_energy(a, S) + Fprior;
                if (!finite_d(F))
                {
                    status = FABBER_VOX_NONFINITE_F;
                    break;
                }
            }
            if (phase == PH_REVERT)
                break;
            if (a.f_history && it < a.f_history_len)
                a.f_history[it * N + v] = F;
            ++it;
            if (conv.test(F))
            {
                if (use_snap)
                {
                    if (conv.need_save())
                        X.template put<false>(snap);
                    if (conv.need_revert())
                    {
                        X.template get<false>(snap);
                        if (!mvn_inverse<P>(X.Lam, X.Sig, X.logdetLam, a.need_f != 0))
                        {
                            status = FABBER_VOX_SINGULAR;
                            break;
                        }
                        phase = PH_REVERT;
                        continue;
                    }
                }
                break;
            }
        }
        if (use_snap && conv.need_save())
            X.template put<false>(snap);
#pragma unroll
        for (int k = 0; k < P; k++)
            Fprior = X.apply_prior(a, k, v, it);
        if (!X.update_theta(S, c, a.need_f != 0))
        {
            status = FABBER_VOX_SINGULAR;
            break;
        }
        const int nerr = X.update_noise(a, S, c);
        if (nerr)
        {
            status = nerr;
            break;
        }
    }

    if (a.f_history)
        for (int h = it; h < a.f_history_len; h++)
            a.f_history[h * N + v] = F;
#pragma unroll
    for (int i = 0; i < P; i++)
        a.mean[i * N + v] = X.m[i];
    const bool zero_cov = (status & 0xff) == FABBER_VOX_SINGULAR;
#pragma unroll
    for (int i = 0; i < NT; i++)
        a.cov[i * N + v] = zero_cov ? 0.0 : X.Sig[i];
    a.noise[0 * N + v] = X.nb;
    a.noise[1 * N + v] = X.nc;
    a.noise[2 * N + v] = X.am[0];
    a.noise[3 * N + v] = X.am[1];
    a.noise[4 * N + v] = X.aprec[0];
    a.noise[5 * N + v] = X.aprec[1];
    a.noise[6 * N + v] = X.aprec[2];
    if (a.free_energy)
        a.free_energy[v] = F;
    if (a.iterations)
        a.iterations[v] = it;
    a.status[v] = status;
}

} // namespace fab
