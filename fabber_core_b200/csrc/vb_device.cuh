/*
 * vb_device.cuh - register-resident building blocks of the per-voxel VB update.
 *
 * Everything here is per-thread (voxel-per-thread layout): packed symmetric P x P algebra with
 * compile-time P so that every array index is a constant after unrolling and the matrices live in
 * registers, the special functions the free energy needs, the parameter transforms and the
 * convergence-detector state machines.
 *
 * Reference behaviour restated (file:line relative to the fabber_core tree):
 *   dist_mvn.cc:197-265     lazy precision <-> covariance inverse with the "+1e-10 I and retry" fallback
 *   tools.cc:87-98          gammaln (6-term Lanczos, not lgamma)
 *   MISCMATHS::digamma      single-precision AS 103 (FSL, not vendored - see DESIGN.md)
 *   transforms.h:114-242    I / L / S / F / A parameter transforms
 *   convergence.cc:43-378   maxits / pointzeroone / freduce / trialmode / lm detectors
 */
#pragma once
#include <cuda_runtime.h>

#include "../../include/fabber_cuda.h"

namespace fab
{
#define FAB_DEV __device__ __forceinline__

/* Index checks for the DEBUG build (make checked: -DFAB_BOUNDS_CHECK): every look-up through an index that came
 * out of a table (neighbour lists, plane order, slab ghost maps, mailbox slots) is tested against the size of
 * what it indexes; a failure is COUNTED in args.check[0] and its site code kept in args.check[1] (largest seen) -
 * the kernel carries on, the host reads the two words with fabber_cuda_check_report(). compute-sanitizer is not
 * available on the GPU pool this was developed on; the checked build is run over the spatial / slab / two-echo
 * tests instead (tests/test_gpu_checked_build.py). The production build compiles the checks away. */
#ifdef FAB_BOUNDS_CHECK
#define FAB_CHECK(args, cond, code)                                                                                  \
    do                                                                                                               \
    {                                                                                                                \
        if ((args).check && !(cond))                                                                                 \
        {                                                                                                            \
            atomicAdd((args).check, 1ull);                                                                           \
            atomicMax((args).check + 1, (unsigned long long)(code));                                                 \
        }                                                                                                            \
    } while (0)
#else
#define FAB_CHECK(args, cond, code)                                                                                  \
    do                                                                                                               \
    {                                                                                                                \
    } while (0)
#endif
#define FAB_CHECK_INDEX(args, idx, n, code) FAB_CHECK(args, (long long)(idx) >= 0 && (long long)(idx) < (long long)(n), code)

/* Pull one line of the voxel series into L2 well ahead of its use. The time loops prefetch three samples into
 * registers, which covers an L2 hit but not a DRAM miss: with one pass per launch (sp_setup, sp_noise) the series
 * streams from HBM, and ncu showed 54 % of sp_noise's stall samples on the
 * consumer of that load (long scoreboard, FP64 pipe 59 % active; profiles/r2n_ncu_full_c5_noise.txt). No register
 * cost, one instruction per sample. */
#ifndef FAB_L2_AHEAD
#define FAB_L2_AHEAD 8
#endif
constexpr int FAB_L2_PREFETCH_AHEAD = FAB_L2_AHEAD;
FAB_DEV void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

/* packed lower triangle by rows: (0,0),(1,0),(1,1),(2,0).. - the order of MVNDist::Save */
__host__ __device__ constexpr int tri(int i, int j)
{
    return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i;
}
template <int P> struct NTri
{
    static constexpr int value = P * (P + 1) / 2;
};

FAB_DEV bool finite_d(double x)
{
    /* exponent field all ones <=> inf or nan; integer test keeps the FP64 pipe free */
    return ((unsigned)__double2hiint(x) & 0x7ff00000u) != 0x7ff00000u;
}

/*
 * log|x1 x2 ... xn| with ONE logarithm: the mantissas (in [1,2)) are multiplied, the exponents added -
 * what NEWMAT's LogAndSign does for LogDeterminant(). n <= 1000 mantissas cannot overflow the product.
 * Zero, denormal and non-finite factors take the plain log() so that the result matches sum(log|x_i|).
 */
struct LogProd
{
    double m, extra;
    int e;
    FAB_DEV void init()
    {
        m = 1.0;
        extra = 0.0;
        e = 0;
    }
    FAB_DEV void mul(double x)
    {
        const int hi = __double2hiint(x) & 0x7fffffff;
        const int ex = hi >> 20;
        if (ex == 0 || ex == 0x7ff)
            extra += log(fabs(x));
        else
        {
            m *= __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(x));
            e += ex - 1023;
        }
    }
    FAB_DEV double value() const { return (log(m) + (double)e * 0.69314718055994530942) + extra; }
};

/* log(2 pi), the value log() returns for 2 * 3.14159265358979323846 */
#define FAB_LOG_2PI 1.8378770664093453

/*
 * LDL^T factorisation of a packed symmetric matrix, inverse and log|det|.
 * Returns false when a pivot is exactly zero or not finite (the condition under which the
 * reference's LU-based inverse raises). Works for indefinite matrices too (negative prior
 * "precisions" are possible with the log transform, transforms.h:153-156).
 */
template <int P>
FAB_DEV bool ldl_inverse(const double (&A)[NTri<P>::value], double (&Inv)[NTri<P>::value], double &logdet,
    bool want_logdet = true)
{
    double L[NTri<P>::value]; /* strictly-lower part: L, diagonal: d */
    double dinv[P];
    bool ok = true;
    LogProd lp; /* only the free energy reads log|det|: four logarithms per inverse were 7 % of a maxits run */
    lp.init();
#pragma unroll
    for (int j = 0; j < P; j++)
    {
        double w[P]; /* w_k = L_jk d_k */
        double d = A[tri(j, j)];
#pragma unroll
        for (int k = 0; k < j; k++)
        {
            w[k] = L[tri(j, k)] * L[tri(k, k)];
            d -= L[tri(j, k)] * w[k];
        }
        ok = ok && finite_d(d) && d != 0.0;
        L[tri(j, j)] = d;
        dinv[j] = 1.0 / d;
        if (want_logdet)
            lp.mul(d);
#pragma unroll
        for (int i = j + 1; i < P; i++)
        {
            double s = A[tri(i, j)];
#pragma unroll
            for (int k = 0; k < j; k++)
                s -= L[tri(i, k)] * w[k];
            L[tri(i, j)] = s * dinv[j];
        }
    }
    /* M = L^-1 (unit lower triangular), stored in the strictly-lower slots of Mx */
    double M[NTri<P>::value];
#pragma unroll
    for (int j = 0; j < P; j++)
    {
        M[tri(j, j)] = 1.0;
#pragma unroll
        for (int i = j + 1; i < P; i++)
        {
            double s = -L[tri(i, j)];
#pragma unroll
            for (int k = j + 1; k < i; k++)
                s -= L[tri(i, k)] * M[tri(k, j)];
            M[tri(i, j)] = s;
        }
    }
    /* Inv = M^T D^-1 M */
#pragma unroll
    for (int i = 0; i < P; i++)
#pragma unroll
        for (int j = 0; j <= i; j++)
        {
            double s = 0.0;
#pragma unroll
            for (int k = i; k < P; k++)
                s += M[tri(k, i)] * M[tri(k, j)] * dinv[k];
            Inv[tri(i, j)] = s;
        }
    logdet = want_logdet ? lp.value() : 0.0;
    return ok;
}

/* MVNDist::GetCovariance / GetPrecisions semantics (dist_mvn.cc:197-265): invert, on failure
 * retry once with 1e-10 added to the diagonal, on a second failure report singular. */
template <int P>
FAB_DEV bool mvn_inverse(const double (&A)[NTri<P>::value], double (&Inv)[NTri<P>::value], double &logdet,
    bool want_logdet = true)
{
    if (ldl_inverse<P>(A, Inv, logdet, want_logdet))
        return true;
    double B[NTri<P>::value];
#pragma unroll
    for (int i = 0; i < NTri<P>::value; i++)
        B[i] = A[i];
#pragma unroll
    for (int i = 0; i < P; i++)
        B[tri(i, i)] += 1e-10;
    double ld2;
    return ldl_inverse<P>(B, Inv, ld2, false);
}

/* y = S x for packed symmetric S */
template <int P> FAB_DEV void symv(const double (&S)[NTri<P>::value], const double (&x)[P], double (&y)[P])
{
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < P; j++)
            s += S[tri(i, j)] * x[j];
        y[i] = s;
    }
}

/* x^T S x */
template <int P> FAB_DEV double quadform(const double (&S)[NTri<P>::value], const double (&x)[P])
{
    double q = 0.0;
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < P; j++)
            s += S[tri(i, j)] * x[j];
        q += x[i] * s;
    }
    return q;
}

/* trace(S1 S2) for two packed symmetric matrices */
template <int P> FAB_DEV double trace_prod(const double (&S1)[NTri<P>::value], const double (&S2)[NTri<P>::value])
{
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < P; i++)
#pragma unroll
        for (int j = 0; j < P; j++)
            t += S1[tri(i, j)] * S2[tri(i, j)];
    return t;
}

/* ------------------------------------------------------------------------------------------------
 * Special functions used by the free energy
 * ---------------------------------------------------------------------------------------------- */
/* tools.cc:87-98 */
FAB_DEV double gammaln(double x)
{
    const double s0 = 2.5066282746310005, s1 = 76.18009172947146, s2 = -86.50532032941677,
                 s3 = 24.01409824083091, s4 = -1.231739572450155, s5 = 0.1208650973866179e-2,
                 s6 = -0.5395239384953e-5;
    double total = 1.000000000190015;
    total += s1 / (x + 1.0);
    total += s2 / (x + 2.0);
    total += s3 / (x + 3.0);
    total += s4 / (x + 4.0);
    total += s5 / (x + 5.0);
    total += s6 / (x + 6.0);
    return log(s0 * total / x) + (x + 0.5) * log(x + 5.5) - x - 5.5;
}

/* MISCMATHS::digamma is `float digamma(const float)` (AS 103). Explicit round-to-nearest float
 * intrinsics stop nvcc contracting the polynomial into FMAs, so the value is the same float the
 * CPU computes. */
FAB_DEV double digamma_fsl(double xin)
{
    const float s = 1e-5f, c = 8.5f, s3 = 8.333333333e-2f, s4 = 8.333333333e-3f, s5 = 3.968253968e-3f,
                d1 = -0.5772156649f;
    float y = (float)xin;
    float dg = 0.0f;
    if (y <= s)
        return (double)__fsub_rn(d1, __fdiv_rn(1.0f, y));
    while (y < c)
    {
        dg = __fsub_rn(dg, __fdiv_rn(1.0f, y));
        y = __fadd_rn(y, 1.0f);
    }
    float r = __fdiv_rn(1.0f, y);
    dg = (float)__dsub_rn(__dadd_rn((double)dg, (double)(float)log((double)y)), __dmul_rn(0.5, (double)r));
    r = __fmul_rn(r, r);
    float poly = __fsub_rn(s3, __fmul_rn(r, __fsub_rn(s4, __fmul_rn(r, s5))));
    dg = __fsub_rn(dg, __fmul_rn(r, poly));
    return (double)dg;
}

/* ------------------------------------------------------------------------------------------------
 * Parameter transforms (transforms.h:114-242, transforms.cc:17-25)
 * ---------------------------------------------------------------------------------------------- */
FAB_DEV double to_model(char code, double v)
{
    switch (code)
    {
    case 'L':
        return exp(v);
    case 'S':
        return v < 10 ? log(1 + exp(v)) : v;
    case 'F':
        return 1 / (1 + exp(v));
    case 'A':
        return fabs(v);
    default:
        return v;
    }
}
FAB_DEV double to_fabber(char code, double v)
{
    switch (code)
    {
    case 'L':
        return log(v);
    case 'S':
        return v < 10 ? log(exp(v) - 1) : v;
    case 'F':
        return log(1 / v - 1);
    default:
        return v;
    }
}
FAB_DEV double to_fabber_var(char code, double v)
{
    switch (code)
    {
    case 'L':
        return log(v);
    case 'I':
    case 'F':
        return v;
    default:
    {
        double t = to_fabber(code, to_model(code, 0.0) + sqrt(v));
        return t * t;
    }
    }
}

/* ------------------------------------------------------------------------------------------------
 * Convergence detectors (convergence.cc). One struct, switch on type: all threads of a launch
 * run the same detector so the branches are warp-uniform.
 * ---------------------------------------------------------------------------------------------- */
struct Conv
{
    int type;
    int its, max_its;
    double prev_f, min_fchange;
    bool revert, save;
    int trials, max_trials;
    bool trialmode;
    bool lm;
    double alpha;

    FAB_DEV void init(int type_, int max_its_, double fchange, int max_trials_)
    {
        type = type_;
        max_its = max_its_ + (type_ == FABBER_CONV_TRIALMODE ? 1 : 0); /* convergence.cc:145 */
        min_fchange = fchange;
        max_trials = max_trials_;
        its = 0;
        prev_f = -99e99; /* convergence.h:46 */
        revert = false;
        trials = 0;
        trialmode = false;
        lm = false;
        alpha = 0.0;
        save = (type_ == FABBER_CONV_TRIALMODE || type_ == FABBER_CONV_LM); /* :158, :270 */
    }
    FAB_DEV bool need_save() const { return type != FABBER_CONV_MAXITS && save; }
    FAB_DEV bool need_revert() const { return type != FABBER_CONV_MAXITS && revert; }
    /* NoiseModel::UpdateTheta takes `float LMalpha` (noisemodel.h:134): alpha is truncated */
    FAB_DEV double lm_alpha() const { return type == FABBER_CONV_LM ? (double)(float)alpha : 0.0; }

    FAB_DEV bool counting()
    {
        ++its;
        return its >= max_its;
    }
    FAB_DEV bool fchange(double F)
    {
        double diff = F - prev_f;
        prev_f = F;
        diff = diff > 0 ? diff : -diff;
        if (diff < min_fchange)
            return true;
        return counting();
    }
    FAB_DEV bool test(double F)
    {
        switch (type)
        {
        case FABBER_CONV_MAXITS:
            return counting();
        case FABBER_CONV_FCHANGE:
            return fchange(F);
        case FABBER_CONV_FREDUCE:
            if (F - prev_f < 0)
            {
                revert = true;
                return true;
            }
            return fchange(F);
        case FABBER_CONV_TRIALMODE:
            return trial(F);
        case FABBER_CONV_LM:
            return lm_test(F);
        }
        return true;
    }
    FAB_DEV bool trial(double F)
    {
        double diff = F - prev_f;
        double absdiff = diff > 0 ? diff : -diff;
        if (!trialmode)
        {
            if (diff < 0)
            {
                its = 1;
                trials = 1;
                trialmode = true;
                revert = true;
                save = false;
                return false;
            }
            if (absdiff < min_fchange)
            {
                revert = false;
                save = false;
                return true;
            }
            save = true;
            revert = false;
            prev_f = F;
            ++its;
            return its >= max_its;
        }
        ++trials;
        if (diff > 0)
        {
            if (absdiff < min_fchange)
            {
                revert = false;
                save = false;
                return true;
            }
            trialmode = false;
            trials = 0;
            save = true;
            revert = false;
            prev_f = F;
            return false;
        }
        if (trials >= max_trials)
        {
            save = false;
            revert = true;
            return true;
        }
        save = false;
        revert = false;
        return false;
    }
    FAB_DEV bool lm_test(double F)
    {
        const double alphastart = 1e-6, alphamax = 1e6;
        double diff = F - prev_f;
        double absdiff = diff < 0 ? -diff : diff;
        if (!lm)
        {
            if (diff < 0)
            {
                lm = true;
                revert = true;
                alpha = alphastart;
                return false;
            }
            if (absdiff < min_fchange)
            {
                revert = false;
                return true;
            }
            if (its >= max_its)
            {
                revert = false;
                return true;
            }
            prev_f = F;
            ++its;
            return false;
        }
        if (diff > 0)
        {
            if (alpha == alphastart)
                lm = false;
            else
                alpha /= 10;
            revert = false;
            prev_f = F;
            ++its;
            return false;
        }
        if (alpha >= alphamax)
        {
            revert = true;
            return true;
        }
        if (its >= max_its)
        {
            revert = false;
            return true;
        }
        alpha *= 10;
        revert = true;
        return false;
    }
};

} // namespace fab
