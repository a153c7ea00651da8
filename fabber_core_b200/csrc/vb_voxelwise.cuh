/*
 * vb_voxelwise.cuh - non-spatial VB with white noise: all iterations of a voxel in one thread.
 *
 * Replaces, per voxel, Vb::SetupPerVoxelDists + the body of Vb::DoCalculationsVoxelwise
 * (inference_vb.cc:144-248, 415-576) with the operators it calls:
 *   LinearizedFwdModel::ReCentre        fwdmodel_linear.cc:126-181
 *   Default/Image/ARD Prior::ApplyToMVN priors.cc:108-181
 *   WhiteNoiseModel::UpdateTheta        noisemodel_white.cc:275-363 (incl. the LM branch)
 *   WhiteNoiseModel::UpdateNoise        noisemodel_white.cc:228-273
 *   WhiteNoiseModel::CalcFreeEnergy     noisemodel_white.cc:365-454
 *   ConvergenceDetector::Test etc.      convergence.cc
 *
 * Re-design (not a translation): the reference keeps a T x P Jacobian, T-vectors and T x T diagonal
 * matrices per voxel and walks them ~10 times per iteration. Here one pass over the time-series per
 * iteration evaluates the model at the 2P+1 finite-difference points and folds each sample straight
 * into the sufficient statistics
 *      A = J^T Q J (packed),  b = J^T Q r,  rr = r^T Q r,   r = y - g(c)
 * per noise precision phi. Everything UpdateTheta / UpdateNoise / CalcFreeEnergy need is a function
 * of (A, b, rr) and P-sized state, because k = y - g + J (c - m) = r + J d:
 *      k^T Q k = rr + 2 b.d + d^T A d,   J^T Q (y - g + J c) = b + A c,   tr(Sigma J^T Q J) = tr(Sigma A).
 * J and g are never stored. The y read is coalesced ([T][N], voxel fastest); the design matrix and
 * noise pattern are staged in shared memory; P x P algebra is fully unrolled in registers.
 */
#pragma once
#include "vb_models.cuh"

namespace fab
{
#define FAB_MAX_PHIS FABBER_CUDA_MAX_PHIS
#define FAB_PAT_MASKED 255
constexpr int VB_BLOCK = 128;

/* Kernel argument block (passed by value as a __grid_constant__ parameter; < 4 KB). */
struct VbArgs
{
    int N, T;
    int basis_jacobian; /* 1: opt-in basis-row Jacobian for linear-in-parameter models (FABBER_B200_BASIS_JACOBIAN=1), see recentre_loop; 0 = the reference's 2P+1 evaluations */
    int v_begin, v_end; /* voxelwise kernels work on voxels [v_begin, v_end) of the N (N stays the array stride) */
    const float *data;          /* [T][N] */
    const double *design;       /* device [T][P], linear model */
    const unsigned char *pattern; /* device [T]: phi index per sample, 255 = masked; NULL = all phi 0 */
    fabber_cuda_param params[FABBER_CUDA_MAX_PARAMS];
    double exp_dt;
    double model_consts[FABBER_CUDA_MODEL_CONSTS]; /* plug-in models */
    int design_len;                                /* plug-in models: doubles behind `design` */
    int n_phis;
    int n_per_phi[FAB_MAX_PHIS]; /* unmasked samples using each phi (Qi.Trace()) */
    int n_unmasked;              /* T - #masked */
    double noise_prior_b[FAB_MAX_PHIS], noise_prior_c[FAB_MAX_PHIS];
    double noise_post_b[FAB_MAX_PHIS], noise_post_c[FAB_MAX_PHIS];
    double locked_noise_stdev;
    double ar_alpha_prior_prec;
    int nlls_lm, nlls_have_start; /* --method=nlls (vb_nlls.cuh) */
    double nlls_start[FABBER_CUDA_MAX_PARAMS];
    int ar_n_alphas; /* AR(1) on two echoes: 2 / 3 / 4 alphas for ar1-cross-terms none / same / dual */
    int conv_type, max_iterations, max_trials, need_f, f_history_len;
    double fchange;
    const double *image_prior[FABBER_CUDA_MAX_PARAMS];
    const double *init_mean, *init_cov, *init_noise;
    const double *lock_centre; /* spatial only */
    double *mean, *cov, *noise, *free_energy, *f_history;
    int *iterations, *status;
    const double *fit_mean; /* model_fit_kernel: [P][N] Fabber-space means in, */
    double *fit_out;        /*                   [T][N] model prediction out (double), or */
    float *fit_out_f32;     /*                   [T][N] model prediction, float32, and/or */
    float *resid_out_f32;   /*                   [T][N] data - prediction, float32 */
    unsigned long long *check; /* [2] failure count, site code: the debug build's index checks (FAB_CHECK); else NULL */
};

/* fit[t][v] = g(ToModel(mean[:, v])): the modelfit / residuals output (inference.cc:190-191) */
template <class Model> __global__ void __launch_bounds__(256) model_fit_kernel(const __grid_constant__ VbArgs a)
{
    constexpr int P = Model::P;
    extern __shared__ double smem[];
    Model::stage(a, smem);
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.N)
        return;
    const typename Model::Ctx mc = Model::make_ctx(a, smem);
    const size_t N = (size_t)a.N;
    double p[P];
#pragma unroll
    for (int i = 0; i < P; i++)
        p[i] = to_model(a.params[i].transform, a.fit_mean[i * N + v]);
    for (int t = 0; t < a.T; t++)
    {
        const double g = Model::eval(mc, t, p);
        if (a.fit_out)
            a.fit_out[t * N + v] = g;
        if (a.fit_out_f32)
            a.fit_out_f32[t * N + v] = (float)g;
        if (a.resid_out_f32)
            a.resid_out_f32[t * N + v] = (float)((double)a.data[t * N + v] - g);
    }
}

template <int P> struct Stats
{
    double A[NTri<P>::value];
    double b[P];
    double rr;
    FAB_DEV void zero()
    {
#pragma unroll
        for (int i = 0; i < NTri<P>::value; i++)
            A[i] = 0.0;
#pragma unroll
        for (int i = 0; i < P; i++)
            b[i] = 0.0;
        rr = 0.0;
    }
    FAB_DEV void add(double r, const double (&J)[P])
    {
        rr = fma(r, r, rr);
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            b[i] = fma(J[i], r, b[i]);
#pragma unroll
            for (int j = 0; j <= i; j++)
                A[tri(i, j)] = fma(J[i], J[j], A[tri(i, j)]);
        }
    }
};

/* the pass over the time-series: model at the 2P+1 finite-difference points, Jacobian row, statistics */
/* CHECK: test g and every Jacobian entry of every sample for non-finite values (ReCentre throws on them,
 * fwdmodel_linear.cc:134,174). With a single phi and no masked samples (NPHI == 1) every sample enters
 * rr = sum r^2 and A_ii = sum J_i^2, and a non-finite r or J_i leaves those sums non-finite for good - so the
 * hot loop runs unchecked (11 of its 80 instructions were the tests) and recentre_stats() looks at the sums
 * afterwards, re-walking the series with recentre_diagnose() only in the rare case they are not finite. */
/* BASIS (models with LINEAR = true): the reference differentiates every model numerically,
 *     J_i(t) = [g(c + d e_i) - g(c - d e_i)] / [(c_i + d) - (c_i - d)]        (fwdmodel_linear.cc:142-172).
 * For g = sum_j phi_j(t) p_j with p = transform(c) the numerator is phi_i(t) [p_i(c_i+d) - p_i(c_i-d)] plus
 * the rounding error of subtracting two nearly equal sums - error, not signal: ~eps |g| / d, up to 1e-8
 * relative in the reference's own J. So J_i(t) = phi_i(t) * s_i with s_i = (pp_i - pn_i) * rden_i formed once
 * per pass is the same Jacobian without that noise, and the pass evaluates the model ONCE per sample:
 * 23 + P FP64 instructions per sample instead of 54 at P = 4. The results stay inside the reference's own
 * noise floor (same parity rule, same tests). OPT-IN (FABBER_B200_BASIS_JACOBIAN=1): the default is the reference's
 * literal 2P+1 evaluations per sample. */
template <class Model, int NPHI, int FAST, bool CHECK, bool BASIS, bool COLD>
FAB_DEV void recentre_loop(const VbArgs &a, const typename Model::Ctx &mc, const unsigned char *pat, int v,
    const double (&p0)[Model::P], const double (&pp)[Model::P], const double (&pn)[Model::P],
    const double (&rden)[Model::P], Stats<Model::P> (&S)[NPHI], bool &bad_g, bool &bad_j)
{
    constexpr int P = Model::P;
    double jscale[P];
#pragma unroll
    for (int i = 0; i < P; i++)
        jscale[i] = (pp[i] - pn[i]) * rden[i];
    const float *yp = a.data + v;
    const size_t stride = (size_t)a.N;
    /* software prefetch, three samples deep: one sample is ~60 FP64 instructions (~120 issue cycles per
     * warp), a miss to HBM several hundred cycles */
    float q0 = __ldg(yp), q1 = 0.f, q2 = 0.f;
    if (1 < a.T)
        q1 = __ldg(yp + stride);
    if (2 < a.T)
        q2 = __ldg(yp + 2 * stride);
    typename Model::Sample smp;
    Model::sample(mc, 0, smp);
    /* COLD: a pass whose series comes from HBM (one pass per launch: the spatial kernels) also pulls it into L2
     * well ahead. The voxelwise kernels run ~10 passes per launch over a series that stays in L2 after the first:
     * there the extra instruction per sample cost 1-4 % (measured), so they do not. */
    if (COLD)
    {
#pragma unroll
        for (int k = 3; k < FAB_L2_PREFETCH_AHEAD; k++)
            if (k < a.T)
                prefetch_l2(yp + (size_t)k * stride);
    }
#pragma unroll 1
    for (int t = 0; t < a.T; t++)
    {
        const double y = (double)q0;
        q0 = q1;
        q1 = q2;
        if (t + 3 < a.T)
            q2 = __ldg(yp + (size_t)(t + 3) * stride);
        if (COLD && t + FAB_L2_PREFETCH_AHEAD < a.T)
            prefetch_l2(yp + (size_t)(t + FAB_L2_PREFETCH_AHEAD) * stride);
        /* the next sample's model-side constants (poly: integer powers -> double, a long-latency conversion)
         * are formed one sample ahead */
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
        if (CHECK)
        {
            bad_g = bad_g || !finite_d(g);
#pragma unroll
            for (int i = 0; i < P; i++)
                bad_j = bad_j || !finite_d(J[i]);
        }
        const double r = y - g;
        if (NPHI == 1)
            S[0].add(r, J);
        else
        {
            const int ph = pat[t];
#pragma unroll
            for (int i = 0; i < NPHI; i++)
                if (ph == i)
                    S[i].add(r, J);
        }
        smp = nxt;
    }
}

/* the rare path of the unchecked loop: which of g / J was non-finite (if any: sums of finite squares can
 * overflow too, and then the reference does not throw here either) */
template <class Model, bool FAST>
FAB_DEV void recentre_diagnose(const VbArgs &a, const typename Model::Ctx &mc, const double (&p0)[Model::P],
    const double (&pp)[Model::P], const double (&pn)[Model::P], const double (&rden)[Model::P], bool &bad_g, bool &bad_j)
{
    constexpr int P = Model::P;
#pragma unroll 1
    for (int t = 0; t < a.T; t++)
    {
        typename Model::Sample smp;
        Model::sample(mc, t, smp);
        double g, gp[P], gn[P];
        Model::template eval_fd<FAST>(mc, smp, p0, pp, pn, g, gp, gn);
        bad_g = bad_g || !finite_d(g);
#pragma unroll
        for (int i = 0; i < P; i++)
            bad_j = bad_j || !finite_d((gp[i] - gn[i]) * rden[i]);
    }
}

/*
 * LinearizedFwdModel::ReCentre fused with the statistics pass. Returns 0, or FABBER_VOX_NONFINITE_*
 * exactly where the reference throws (offset checked before the Jacobian, fwdmodel_linear.cc:134,174).
 * NPHI == 1: single phi, no masked samples (fast path). NPHI > 1: `pat` gives the phi per sample.
 */
template <class Model, int NPHI, bool COLD = false>
FAB_DEV int recentre_stats(const VbArgs &a, const typename Model::Ctx &mc, const unsigned char *pat, int v,
    const double (&c)[Model::P], Stats<Model::P> (&S)[NPHI])
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
#pragma unroll
    for (int i = 0; i < NPHI; i++)
        S[i].zero();
    bool bad_g = false, bad_j = false;
    /* models with a range-limited cheaper evaluation (exp: table-based exponential) take it when every
     * argument of this pass is inside its range - checked once here, not per sample */
    constexpr bool CHECK = NPHI > 1; /* masked samples never reach the sums: test them one by one */
    const bool fast = Model::HAS_FAST && Model::fast_ok(mc, a.T, p0, pp, pn);
    bool basis = false; /* opt-in, and only for models that hand out their basis row */
    if constexpr (Model::LINEAR)
        basis = a.basis_jacobian != 0;
    if (basis)
    {
        if constexpr (Model::LINEAR)
            recentre_loop<Model, NPHI, 0, CHECK, true, COLD>(a, mc, pat, v, p0, pp, pn, rden, S, bad_g, bad_j);
    }
    else if (fast)
        recentre_loop<Model, NPHI, 1, CHECK, false, COLD>(a, mc, pat, v, p0, pp, pn, rden, S, bad_g, bad_j);
    else
        recentre_loop<Model, NPHI, 0, CHECK, false, COLD>(a, mc, pat, v, p0, pp, pn, rden, S, bad_g, bad_j);
    if (!CHECK)
    {
        bool sums_finite = finite_d(S[0].rr);
#pragma unroll
        for (int i = 0; i < P; i++)
            sums_finite = sums_finite && finite_d(S[0].A[tri(i, i)]);
        if (!sums_finite)
        {
            if (fast)
                recentre_diagnose<Model, true>(a, mc, p0, pp, pn, rden, bad_g, bad_j);
            else
                recentre_diagnose<Model, false>(a, mc, p0, pp, pn, rden, bad_g, bad_j);
        }
    }
    return bad_g ? FABBER_VOX_NONFINITE_OFFSET : (bad_j ? FABBER_VOX_NONFINITE_JACOBIAN : 0);
}

/* WhiteNoiseModel::CalcFreeEnergy with c == m (always true where Vb reads F: after ReCentre). */
/* What the free energy re-derives every iteration although it never changes: gammaln / digamma of the noise
 * shape c (constant from the first UpdateNoise on: c = (n-1)/2 + c0) and the prior's own constant. Measured
 * on C3 (LM, F every iteration): the two gammaln + digamma were 9 % of the kernel. Keyed on the value of c,
 * so a changed c (first iteration, restarts) just recomputes. */
template <int NPHI> struct FCache
{
    double key[NPHI], lgam[NPHI], dgam[NPHI], prior[NPHI];
    FAB_DEV void init(const VbArgs &a)
    {
#pragma unroll
        for (int i = 0; i < NPHI; i++)
        {
            key[i] = -1.0; /* a Gamma shape is positive: never matches */
            lgam[i] = dgam[i] = prior[i] = 0.0;
            if (a.need_f && i < a.n_phis)
                prior[i] = -gammaln(a.noise_prior_c[i]) - a.noise_prior_c[i] * log(a.noise_prior_b[i]);
        }
    }
    FAB_DEV void lookup(int i, double ci, double &lg, double &dg)
    {
        if (ci != key[i])
        {
            key[i] = ci;
            lgam[i] = gammaln(ci);
            dgam[i] = digamma_fsl(ci);
        }
        lg = lgam[i];
        dg = dgam[i];
    }
};

template <int P, int NPHI>
FAB_DEV double white_free_energy(const VbArgs &a, const Stats<P> (&S)[NPHI], const double (&m)[P],
    const double (&Sig)[NTri<P>::value], double logdetLam, const double (&m0)[P], const double (&L0)[P],
    const double (&nb)[NPHI], const double (&nc)[NPHI], FCache<NPHI> &fc)
{
    const double log2pi = FAB_LOG_2PI;
    const double elTheta = 0.5 * logdetLam - 0.5 * P * (log2pi + 1);
    double elPhi = 0.0, p0 = 0.0, p2 = 0.0, p9 = 0.0;
#pragma unroll
    for (int i = 0; i < NPHI; i++)
    {
        if (i < a.n_phis)
        {
            const double si = nb[i], ci = nc[i];
            const double siP = a.noise_prior_b[i], ciP = a.noise_prior_c[i];
            double lg, dg;
            fc.lookup(i, ci, lg, dg);
            const double lsi = log(si);
            elPhi += -lg - ci * lsi - ci + (ci - 1) * (dg + lsi);
            p0 += (dg + lsi) * ((double)a.n_per_phi[i] * 0.5 + ciP - 1);
            p9 += fc.prior[i] - si * ci / siP;
            p2 += -0.5 * si * ci * S[i].rr - 0.5 * trace_prod<P>(S[i].A, Sig);
        }
    }
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
    const double ld0 = lp0.value();
    const double p3 = 0.5 * ld0 - 0.5 * a.n_unmasked * log2pi - 0.5 * P * log2pi;
    const double p4 = -0.5 * q;
    const double p5 = -0.5 * tr0;
    double F = -elTheta - elPhi;
    F += p0;
    F += p2;
    F += p3;
    F += p4;
    F += p5;
    F += p9;
    return F;
}

template <class Model, int NPHI, bool SNAP> struct WhiteVoxel
{
    static constexpr int P = Model::P;
    static constexpr int NT = NTri<P>::value;

    double m[P], Lam[NT], Sig[NT], m0[P], L0[P], nb[NPHI], nc[NPHI];
    double logdetLam;

    /* Shared-memory parking. The state above (~40 doubles at P = 4) is dead during the pass over the
     * time-series; left in registers it would cost ~80 registers of occupancy in the hot loop. It is
     * parked in shared memory around the pass instead ([slot][thread] layout: conflict-free 64-bit
     * accesses; volatile so the compiler cannot forward the values through registers).
     * The trialmode / freduce snapshot (inference_vb.cc:432-434,451-458) lives there permanently. */
    FCache<NPHI> fc;
    static constexpr int STASH_DOUBLES = 3 * P + 2 * NT + 2 * NPHI + 1 + 4 * NPHI;
    static constexpr int SNAP_DOUBLES = SNAP ? 3 * P + NT + 2 * NPHI : 0;
    FAB_DEV void stash(volatile double *s) const
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
            s[(k++) * VB_BLOCK] = Sig[i];
        }
#pragma unroll
        for (int i = 0; i < NPHI; i++)
        {
            s[(k++) * VB_BLOCK] = nb[i];
            s[(k++) * VB_BLOCK] = nc[i];
        }
        s[(k++) * VB_BLOCK] = logdetLam;
#pragma unroll
        for (int i = 0; i < NPHI; i++)
        {
            s[(k++) * VB_BLOCK] = fc.key[i];
            s[(k++) * VB_BLOCK] = fc.lgam[i];
            s[(k++) * VB_BLOCK] = fc.dgam[i];
            s[(k++) * VB_BLOCK] = fc.prior[i];
        }
    }
    FAB_DEV void unstash(const volatile double *s)
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
            Sig[i] = s[(k++) * VB_BLOCK];
        }
#pragma unroll
        for (int i = 0; i < NPHI; i++)
        {
            nb[i] = s[(k++) * VB_BLOCK];
            nc[i] = s[(k++) * VB_BLOCK];
        }
        logdetLam = s[(k++) * VB_BLOCK];
#pragma unroll
        for (int i = 0; i < NPHI; i++)
        {
            fc.key[i] = s[(k++) * VB_BLOCK];
            fc.lgam[i] = s[(k++) * VB_BLOCK];
            fc.dgam[i] = s[(k++) * VB_BLOCK];
            fc.prior[i] = s[(k++) * VB_BLOCK];
        }
    }
    FAB_DEV void save(volatile double *s) const
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
            s[(k++) * VB_BLOCK] = Lam[i];
#pragma unroll
        for (int i = 0; i < NPHI; i++)
        {
            s[(k++) * VB_BLOCK] = nb[i];
            s[(k++) * VB_BLOCK] = nc[i];
        }
    }
    FAB_DEV void restore(const volatile double *s)
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
            Lam[i] = s[(k++) * VB_BLOCK];
#pragma unroll
        for (int i = 0; i < NPHI; i++)
        {
            nb[i] = s[(k++) * VB_BLOCK];
            nc[i] = s[(k++) * VB_BLOCK];
        }
    }

    /* priors.cc:108-181. Returns the free-energy contribution of parameter k. */
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

    /* noisemodel_white.cc:275-363. Returns false if the posterior precision is singular. */
    FAB_DEV bool update_theta(const VbArgs &a, const Stats<P> (&S)[NPHI], const double (&c)[P], double alpha)
    {
        double Aw[NT], bw[P];
#pragma unroll
        for (int i = 0; i < NT; i++)
            Aw[i] = 0.0;
#pragma unroll
        for (int i = 0; i < P; i++)
            bw[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NPHI; i++)
            if (i < a.n_phis)
            {
                const double x = nb[i] * nc[i]; /* GammaDist::CalcMean */
#pragma unroll
                for (int j = 0; j < NT; j++)
                    Aw[j] = fma(x, S[i].A[j], Aw[j]);
#pragma unroll
                for (int j = 0; j < P; j++)
                    bw[j] = fma(x, S[i].b[j], bw[j]);
            }
#pragma unroll
        for (int i = 0; i < NT; i++)
            Lam[i] = Aw[i];
#pragma unroll
        for (int i = 0; i < P; i++)
            Lam[tri(i, i)] = L0[i] + Aw[tri(i, i)];
        double P0m0[P];
#pragma unroll
        for (int i = 0; i < P; i++)
            P0m0[i] = L0[i] * m0[i];
        if (alpha <= 0.0)
        {
            if (!mvn_inverse<P>(Lam, Sig, logdetLam, a.need_f != 0))
                return false;
            double Ac[P], rhs[P];
            symv<P>(Aw, c, Ac);
#pragma unroll
            for (int i = 0; i < P; i++)
                rhs[i] = (bw[i] + Ac[i]) + P0m0[i];
            symv<P>(Sig, rhs, m);
        }
        else
        {
            double D[NT], Dinv[NT], Delta[P], step[P], ld;
#pragma unroll
            for (int i = 0; i < NT; i++)
                D[i] = Lam[i];
#pragma unroll
            for (int i = 0; i < P; i++)
            {
                D[tri(i, i)] = Lam[tri(i, i)] + alpha * Lam[tri(i, i)];
                Delta[i] = bw[i] + P0m0[i] - L0[i] * c[i];
            }
            if (ldl_inverse<P>(D, Dinv, ld, false))
            {
                symv<P>(Dinv, Delta, step);
#pragma unroll
                for (int i = 0; i < P; i++)
                    m[i] = c[i] + step[i];
            }
            /* Sigma is first needed by UpdateNoise (theta.GetCovariance(), noisemodel_white.cc:252) */
            if (!mvn_inverse<P>(Lam, Sig, logdetLam, a.need_f != 0))
                return false;
        }
        return true;
    }

    /* noisemodel_white.cc:228-273 */
    FAB_DEV void update_noise(const VbArgs &a, const Stats<P> (&S)[NPHI], const double (&c)[P])
    {
        double d[P];
#pragma unroll
        for (int i = 0; i < P; i++)
            d[i] = c[i] - m[i];
#pragma unroll
        for (int i = 0; i < NPHI; i++)
            if (i < a.n_phis)
            {
                double bd = 0.0;
#pragma unroll
                for (int j = 0; j < P; j++)
                    bd += S[i].b[j] * d[j];
                const double kk = S[i].rr + 2.0 * bd + quadform<P>(S[i].A, d);
                const double tmp = kk + trace_prod<P>(Sig, S[i].A);
                nb[i] = 1 / (tmp * 0.5 + 1 / a.noise_prior_b[i]);
                nc[i] = ((double)a.n_per_phi[i] - 1) * 0.5 + a.noise_prior_c[i];
                if (a.locked_noise_stdev > 0)
                    nb[i] = 1 / nc[i] / a.locked_noise_stdev / a.locked_noise_stdev;
            }
    }
};

#ifndef FAB_MIN_BLOCKS
#define FAB_MIN_BLOCKS 3 /* measured on B200: 3 x 128 threads/SM (<=168 regs, no spills) beats 2 and 4, profiles/ */
#endif
template <class Model, int NPHI, bool SNAP>
__global__ void __launch_bounds__(VB_BLOCK, FAB_MIN_BLOCKS) vb_voxelwise_white_kernel(const __grid_constant__ VbArgs a)
{
    constexpr int P = Model::P;
    constexpr int NT = NTri<P>::value;
    extern __shared__ double smem[];
    Model::stage(a, smem);
    typedef WhiteVoxel<Model, NPHI, SNAP> Vox;
    /* dynamic shared memory: [model constants][state parking][snapshot][noise pattern] */
    volatile double *park = smem + Model::smem_bytes(a.T) / sizeof(double) + threadIdx.x;
    volatile double *snap = park + Vox::STASH_DOUBLES * VB_BLOCK;
    unsigned char *pat = reinterpret_cast<unsigned char *>(
        smem + Model::smem_bytes(a.T) / sizeof(double) + (Vox::STASH_DOUBLES + Vox::SNAP_DOUBLES) * VB_BLOCK);
    if (NPHI > 1)
        for (int i = threadIdx.x; i < a.T; i += blockDim.x)
            pat[i] = a.pattern ? a.pattern[i] : 0;
    __syncthreads();
    const int v = a.v_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.v_end)
        return;
    const typename Model::Ctx mc = Model::make_ctx(a, smem);
    const size_t N = (size_t)a.N;

    Vox X;
    int status = 0;
    double F = 1234.5678; /* inference_vb.cc:438 */
    int it = 0;

    /* ---- SetupPerVoxelDists (inference_vb.cc:207-247, fwdmodel.cc:284-324) ------------------- */
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
#pragma unroll
    for (int i = 0; i < NPHI; i++)
    {
        X.nb[i] = a.init_noise ? a.init_noise[(2 * i) * N + v] : a.noise_post_b[i];
        X.nc[i] = a.init_noise ? a.init_noise[(2 * i + 1) * N + v] : a.noise_post_c[i];
        if (i >= a.n_phis)
            X.nb[i] = X.nc[i] = 1.0;
    }
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        X.m0[i] = 0.0; /* fwd_prior starts as N(0, I), inference_vb.cc:159 */
        X.L0[i] = 1.0;
    }

    /* One call site for the pass over the time-series (keeps the hot loop a single copy in the
     * instruction cache): a small state machine walks set-up -> iterations -> optional revert.
     *   SETUP  : ReCentre of SetupPerVoxelDists (:235). A failure here is never caught by the reference.
     *            The second ReCentre at :443 recomputes the same values from the same centre.
     *   ITER   : ReCentre at the end of an iteration (:490), then F, ++it and Test() (:495-500).
     *   REVERT : ReCentre + F after restoring the snapshot (:516-525).                              */
    enum
    {
        PH_SETUP,
        PH_ITER,
        PH_REVERT
    };
    X.fc.init(a);
    Stats<P> S[NPHI];
    double c[P];
    Conv conv;
    conv.init(a.conv_type, a.max_iterations, a.fchange, a.max_trials);
    double Fprior = 0.0;
    int phase = PH_SETUP;
    while (status == 0)
    {
#pragma unroll
        for (int i = 0; i < P; i++)
            c[i] = X.m[i];
        X.stash(park);
        const int err = recentre_stats<Model, NPHI>(a, mc, pat, v, c, S);
        X.unstash(park);
        if (phase == PH_SETUP)
        {
            if (err)
            {
                status = err | FABBER_VOX_SETUP_FLAG;
                break;
            }
            if (SNAP)
                X.save(snap); /* pre-loop copies, inference_vb.cc:432-434 */
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
                F = white_free_energy<P, NPHI>(a, S, X.m, X.Sig, X.logdetLam, X.m0, X.L0, X.nb, X.nc, X.fc) + Fprior;
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
                /* LM: m_save is always true (convergence.cc:270), so the snapshot taken after the
                 * loop (:506-513) is the current state and the revert at :516-525 restores exactly
                 * that, re-centres on the same means and recomputes the same F: a no-op on every
                 * output, so only detectors with a real snapshot (SNAP) take the revert path. */
                if (SNAP)
                {
                    if (conv.need_save())
                        X.save(snap);
                    if (conv.need_revert())
                    {
                        X.restore(snap);
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
        /* ---- one VB iteration: priors, theta, noise (inference_vb.cc:451-480) ---- */
        if (SNAP && conv.need_save())
            X.save(snap);
#pragma unroll
        for (int k = 0; k < P; k++)
            Fprior = X.apply_prior(a, k, v, it); /* '=' not '+=': inference_vb.cc:462 */
        if (!X.update_theta(a, S, c, conv.lm_alpha()))
        {
            status = FABBER_VOX_SINGULAR;
            break;
        }
        X.update_noise(a, S, c);
    }

    /* ---- results (inference_vb.cc:546-570; padded F history :1041-1044) ----------------------- */
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
#pragma unroll
    for (int i = 0; i < NPHI; i++)
        if (i < a.n_phis)
        {
            a.noise[(2 * i) * N + v] = X.nb[i];
            a.noise[(2 * i + 1) * N + v] = X.nc[i];
        }
    if (a.free_energy)
        a.free_energy[v] = F;
    if (a.iterations)
        a.iterations[v] = it;
    a.status[v] = status;
}

} // namespace fab
