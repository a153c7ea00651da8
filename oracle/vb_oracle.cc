/*
 * vb_oracle.cc - TEST INFRASTRUCTURE ONLY. CPU (FP64, single thread) restatement of fabber_core's
 * Variational Bayes update loop, used as the parity checker for the CUDA path and as the
 * "port" CPU baseline in bench.py. Nothing in the product (fabber_core_b200/) may include, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs use it.
 *
 * It deliberately follows the reference's *order of sub-steps and quirks* in "direct form"
 * (T-length vectors, stored T x P Jacobian, full P x P matrices with lazily synchronised
 * precision/covariance), i.e. it does NOT share the sufficient-statistics restructuring of the
 * CUDA kernels, so agreement between the two is evidence, not tautology.
 *
 * Reference files followed (all under /root/reference):
 *   inference_vb.cc:144-248 (SetupPerVoxelDists), :415-576 (DoCalculationsVoxelwise),
 *                  :578-767 (DoCalculationsSpatial), :266-297 (IgnoreVoxel), :830-964 (CalcNeighbours)
 *   fwdmodel_linear.cc:126-181 (LinearizedFwdModel::ReCentre), :92-96 (LinearFwdModel::EvaluateModel)
 *   fwdmodel.cc:284-324 (GetInitialPosterior/ToFabber), :365-382 (EvaluateFabber)
 *   fwdmodel_poly.cc:62-80, examples/fwdmodel_exp.cc:65-91
 *   transforms.h:114-242, transforms.cc:17-25
 *   noisemodel_white.cc:127-454, noisemodel_ar.cc:83-223,379-769 (num-echoes=1 and 2, all ar1-cross-terms)
 *   priors.cc:108-181 (Default/Image/ARD), :221-488 (SpatialPrior)
 *   convergence.cc:34-378, convergence.h
 *   dist_mvn.cc:57-100,197-265, dist_gamma.cc:21-33, tools.cc:87-98 (gammaln)
 *   inference_nlls.cc:57-293 (NLLSInferenceTechnique, NLLSCF) - see the NLLS section for its optimiser
 *
 * Third-party arithmetic that is NOT in the reference tree (FSL armawrap/NEWMAT and
 * MISCMATHS::digamma, un-pinned - see DESIGN.md "oracle"): matrix inverse and log-determinant are
 * restated as LU with partial pivoting; digamma is restated as Bernardo's AS 103 evaluated in
 * single precision (the FSL signature is `float digamma(const float)`). Free energy therefore has
 * "parity unpinned" status with respect to the original binary: no golden in the reference pins it.
 * The NLLS optimiser is a third FSL dependency outside the tree (MISCMATHS::nonlin): restated from its published
 * algorithm and pinned by the reference's golden test/outdata_linear_nlls (means, z-statistics, finalMVN; the golden
 * tells Levenberg from Levenberg-Marquardt and reproduces the reference's own 1e-4 convergence error to 6e-6);
 * "parity unpinned" beyond that linear case - oracle/_ref cannot run NLLS (built with NO_NLLS).
 *
 * Pinned against: test/outdata_linear_vb, test/outdata_linear_nlls and test/outdata_poly goldens on the 18 voxels of
 * test/test_data_small.nii.gz (tests/test_oracle_golden.py), test/test_convergence.cc sequences
 * (tests/test_convergence.py), test/test_priors.cc and test/test_spatialvb.cc expectations.
 */
#include "../include/fabber_cuda.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace
{
/* Every quantity of the algorithm is a `real`: double in the oracle proper (NEWMAT::Real is double in the
 * reference). -DORACLE_LONG_DOUBLE builds the SAME source in 80-bit extended precision (64-bit mantissa: rounding
 * noise 2048 x smaller) - the "truth" build of tests/parity.py: where the reference's own FP64 arithmetic is the
 * limit (ill-conditioned normal equations), the CUDA path is required to be as close to that truth as the
 * reference's arithmetic is. Inputs, outputs and the C interface stay double / float. */
#ifdef ORACLE_LONG_DOUBLE
typedef long double real;
#else
typedef double real;
#endif
typedef std::vector<real> Vec;

struct InternalError : std::runtime_error
{
    int code;
    InternalError(int c, const char *m)
        : std::runtime_error(m)
        , code(c)
    {
    }
};
struct SingularError : std::runtime_error
{
    SingularError()
        : std::runtime_error("matrix is singular")
    {
    }
};

// ---------------------------------------------------------------------------------------------
// Tiny dense matrix (row-major, 0-based)
// ---------------------------------------------------------------------------------------------
struct Mat
{
    int r, c;
    Vec a;
    Mat()
        : r(0)
        , c(0)
    {
    }
    Mat(int r_, int c_, real v = 0.0)
        : r(r_)
        , c(c_)
        , a((size_t)r_ * c_, v)
    {
    }
    real &operator()(int i, int j) { return a[(size_t)i * c + j]; }
    real operator()(int i, int j) const { return a[(size_t)i * c + j]; }
    static Mat identity(int n)
    {
        Mat m(n, n);
        for (int i = 0; i < n; i++)
            m(i, i) = 1.0;
        return m;
    }
};

Mat mul(const Mat &A, const Mat &B)
{
    Mat C(A.r, B.c);
    for (int i = 0; i < A.r; i++)
        for (int j = 0; j < B.c; j++)
        {
            real s = 0;
            for (int k = 0; k < A.c; k++)
                s += A(i, k) * B(k, j);
            C(i, j) = s;
        }
    return C;
}
Mat transpose(const Mat &A)
{
    Mat B(A.c, A.r);
    for (int i = 0; i < A.r; i++)
        for (int j = 0; j < A.c; j++)
            B(j, i) = A(i, j);
    return B;
}
Mat add(const Mat &A, const Mat &B)
{
    Mat C(A.r, A.c);
    for (size_t i = 0; i < A.a.size(); i++)
        C.a[i] = A.a[i] + B.a[i];
    return C;
}
Vec mulv(const Mat &A, const Vec &x)
{
    Vec y(A.r);
    for (int i = 0; i < A.r; i++)
    {
        real s = 0;
        for (int k = 0; k < A.c; k++)
            s += A(i, k) * x[k];
        y[i] = s;
    }
    return y;
}
real trace(const Mat &A)
{
    real s = 0;
    for (int i = 0; i < A.r; i++)
        s += A(i, i);
    return s;
}
bool all_finite(const Mat &A)
{
    for (size_t i = 0; i < A.a.size(); i++)
        if (!std::isfinite(A.a[i]))
            return false;
    return true;
}
bool all_finite(const Vec &v)
{
    for (size_t i = 0; i < v.size(); i++)
        if (!std::isfinite(v[i]))
            return false;
    return true;
}

// LU with partial pivoting (restates LAPACK getrf as used by armawrap's .i() / LogDeterminant()).
struct LU
{
    Mat lu;
    std::vector<int> piv;
    int sign;
    bool singular;
    explicit LU(const Mat &A)
        : lu(A)
        , piv(A.r)
        , sign(1)
        , singular(false)
    {
        int n = A.r;
        for (int k = 0; k < n; k++)
        {
            int p = k;
            real best = std::fabs(lu(k, k));
            for (int i = k + 1; i < n; i++)
                if (std::fabs(lu(i, k)) > best)
                {
                    best = std::fabs(lu(i, k));
                    p = i;
                }
            piv[k] = p;
            if (!(best > 0.0) || !std::isfinite(best))
            {
                singular = true;
                continue;
            }
            if (p != k)
            {
                for (int j = 0; j < n; j++)
                    std::swap(lu(k, j), lu(p, j));
                sign = -sign;
            }
            for (int i = k + 1; i < n; i++)
            {
                lu(i, k) /= lu(k, k);
                real f = lu(i, k);
                for (int j = k + 1; j < n; j++)
                    lu(i, j) -= f * lu(k, j);
            }
        }
    }
};

Mat inverse(const Mat &A)
{
    int n = A.r;
    if (!all_finite(A))
        throw SingularError();
    LU f(A);
    if (f.singular)
        throw SingularError();
    Mat inv(n, n);
    for (int col = 0; col < n; col++)
    {
        Vec b(n, 0.0);
        b[col] = 1.0;
        for (int k = 0; k < n; k++)
            if (f.piv[k] != k)
                std::swap(b[k], b[f.piv[k]]);
        for (int i = 1; i < n; i++)
        {
            real s = b[i];
            for (int j = 0; j < i; j++)
                s -= f.lu(i, j) * b[j];
            b[i] = s;
        }
        for (int i = n - 1; i >= 0; i--)
        {
            real s = b[i];
            for (int j = i + 1; j < n; j++)
                s -= f.lu(i, j) * b[j];
            b[i] = s / f.lu(i, i);
        }
        for (int i = 0; i < n; i++)
            inv(i, col) = b[i];
    }
    // NEWMAT SymmetricMatrix assignment keeps the lower triangle (lossy symmetric assignment)
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++)
            inv(i, j) = inv(j, i);
    return inv;
}

struct LogAndSign
{
    real logval;
    int sign;
};
LogAndSign log_determinant(const Mat &A)
{
    LU f(A);
    LogAndSign r;
    r.logval = 0;
    r.sign = f.sign;
    for (int i = 0; i < A.r; i++)
    {
        real d = f.lu(i, i);
        if (d == 0.0 || f.singular)
        {
            r.sign = 0;
            r.logval = -INFINITY;
            return r;
        }
        if (d < 0)
            r.sign = -r.sign;
        r.logval += std::log(std::fabs(d));
    }
    return r;
}

// ---------------------------------------------------------------------------------------------
// Special functions
// ---------------------------------------------------------------------------------------------
// tools.cc:87-98 - 6-term Lanczos, NOT lgamma
real gammaln(real x)
{
    static const real series[7] = { 2.5066282746310005, 76.18009172947146, -86.50532032941677,
        24.01409824083091, -1.231739572450155, 0.1208650973866179e-2, -0.5395239384953e-5 };
    real total = 1.000000000190015;
    for (int i = 2; i <= 7; i++)
        total += series[i - 1] / (x + i - 1);
    return std::log(series[0] * total / x) + (x + 0.5) * std::log(x + 5.5) - x - 5.5;
}

// MISCMATHS::digamma (FSL, not in the reference tree): `float digamma(const float x)`,
// Bernardo's algorithm AS 103 in single precision. Restated; see header comment (unpinned).
double digamma_fsl(double xin)
{
    const float s = 1e-5f, c = 8.5f, s3 = 8.333333333e-2f, s4 = 8.333333333e-3f,
                s5 = 3.968253968e-3f, d1 = -0.5772156649f;
    float y = (float)xin;
    float dg = 0.0f;
    if (y <= s)
        return (double)(d1 - 1.0f / y);
    while (y < c)
    {
        dg = dg - 1.0f / y;
        y = y + 1.0f;
    }
    float r = 1.0f / y;
    dg = (float)((double)dg + (double)(float)std::log((double)y) - 0.5 * (double)r);
    r = r * r;
    dg = dg - r * (s3 - r * (s4 - r * s5));
    return (double)dg;
}

// ---------------------------------------------------------------------------------------------
// Transforms (transforms.h:114-242, transforms.cc:17-25)
// ---------------------------------------------------------------------------------------------
real t_to_model(char code, real v)
{
    switch (code)
    {
    case 'L':
        return std::exp(v);
    case 'S':
        return v < 10 ? std::log(1 + std::exp(v)) : v;
    case 'F':
        return 1 / (1 + std::exp(v));
    case 'A':
        return std::fabs(v);
    default:
        return v;
    }
}
real t_to_fabber(char code, real v)
{
    switch (code)
    {
    case 'L':
        return std::log(v);
    case 'S':
        return v < 10 ? std::log(std::exp(v) - 1) : v;
    case 'F':
        return std::log(1 / v - 1);
    default:
        return v;
    }
}
real t_to_fabber_var(char code, real v)
{
    switch (code)
    {
    case 'L':
        return std::log(v);
    case 'I':
    case 'F':
        return v;
    default: // generic rule transforms.cc:22-25
        return std::pow(t_to_fabber(code, t_to_model(code, 0) + std::sqrt(v)), 2);
    }
}

// ---------------------------------------------------------------------------------------------
// MVN with lazily synchronised precision / covariance (dist_mvn.cc:197-265)
// ---------------------------------------------------------------------------------------------
struct MVN
{
    int n;
    Vec means;
    mutable Mat prec, cov;
    mutable bool pv, cv;
    explicit MVN(int n_ = 0)
        : n(n_)
        , means(n_, 0.0)
        , prec(Mat::identity(n_))
        , cov(Mat::identity(n_))
        , pv(true)
        , cv(true)
    {
    }
    const Mat &GetPrecisions() const
    {
        if (!pv)
        {
            try
            {
                prec = inverse(cov);
            }
            catch (SingularError &)
            {
                Mat tmp = cov;
                for (int i = 0; i < n; i++)
                    tmp(i, i) += 1e-10;
                prec = inverse(tmp);
            }
            pv = true;
        }
        return prec;
    }
    const Mat &GetCovariance() const
    {
        if (!cv)
        {
            try
            {
                cov = inverse(prec);
            }
            catch (SingularError &)
            {
                Mat tmp = prec;
                for (int i = 0; i < n; i++)
                    tmp(i, i) += 1e-10;
                cov = inverse(tmp);
            }
            cv = true;
        }
        return cov;
    }
    void SetPrecisions(const Mat &p)
    {
        prec = p;
        pv = true;
        cv = false;
    }
    void SetCovariance(const Mat &c)
    {
        cov = c;
        cv = true;
        pv = false;
    }
};

// ---------------------------------------------------------------------------------------------
// Forward models (A3) and the linearised model (A2)
// ---------------------------------------------------------------------------------------------
struct ModelCtx
{
    const fabber_cuda_vb_problem *prob;
    int T, P;
};

// Noise-floor probe only (oracle/Makefile target libvb_oracle_ulp.so, -DORACLE_EXP_ULP_PROBE): model
// the forward model being linked against a different, equally valid libm whose exp() is accurate to
// <= 1 ULP instead of glibc's: half of the results are moved to an adjacent real, chosen by a hash of
// the bits. Never defined for the oracle proper.
inline real model_exp(real x)
{
    real e = std::exp(x);
#ifdef ORACLE_EXP_ULP_PROBE
    unsigned long long u;
    std::memcpy(&u, &e, sizeof(u));
    unsigned long long h = u * 0x9E3779B97F4A7C15ull;
    h ^= h >> 29;
    if (std::isfinite(e) && e != 0.0)
    {
        if ((h & 3ull) == 1ull)
            e = std::nextafter(e, INFINITY);
        else if ((h & 3ull) == 2ull)
            e = std::nextafter(e, -INFINITY);
    }
#endif
    return e;
}

static const int ORACLE_MODEL_SINE = 101; // cuda_abi.MODEL_ORACLE_SINE

void evaluate_model(const ModelCtx &mc, const Vec &p, Vec &result)
{
    const fabber_cuda_model &m = mc.prob->model;
    const int T = mc.T;
    result.assign(T, 0.0);
    if (m.id == FABBER_MODEL_LINEAR)
    {
        // fwdmodel_linear.cc:95  result = J*(params - centre) + offset, centre = offset = 0
        for (int t = 0; t < T; t++)
        {
            real s = 0;
            for (int j = 0; j < mc.P; j++)
                s += m.design[(size_t)t * mc.P + j] * (p[j] - 0.0);
            result[t] = s + 0.0;
        }
    }
    else if (m.id == FABBER_MODEL_POLY)
    {
        // fwdmodel_poly.cc:68-79 - note the *int* power accumulator (wraps for large i^n)
        for (int i = 1; i <= T; i++)
        {
            real res = 0;
            unsigned int pw = 1; // unsigned arithmetic == two's complement wrap of the reference's int
            for (int n = 0; n <= m.poly_degree; n++)
            {
                res += p[n] * (real)(int)pw;
                pw *= (unsigned int)i;
            }
            result[i - 1] = res;
        }
    }
    else if (m.id == FABBER_MODEL_EXP)
    {
        // examples/fwdmodel_exp.cc:71-81
        for (int k = 0; k < m.exp_num; k++)
        {
            real amp = p[2 * k];
            real r = p[2 * k + 1];
            for (int i = 0; i < T; i++)
            {
                real t = real(i) * m.exp_dt;
                real val = amp * model_exp(-r * t);
                result[i] += val;
            }
        }
    }
    else if (m.id == ORACLE_MODEL_SINE)
    {
        // not a reference model: the example PLUG-IN model of fabber_core_b200/examples/sine_model.cu, restated
        // here so that the plug-in path has an independent checker.  g = a sin(b (t - c)) + d, t = i dt
        for (int i = 0; i < T; i++)
        {
            real t = real(i) * m.consts[0];
            result[i] = p[0] * std::sin(p[1] * (t - p[2])) + p[3];
        }
    }
    else
        throw std::runtime_error("unknown model id");
}

// fwdmodel.cc:365-382
void evaluate_fabber(const ModelCtx &mc, const Vec &theta, Vec &result)
{
    Vec tp(mc.P);
    for (int i = 0; i < mc.P; i++)
        tp[i] = t_to_model(mc.prob->params[i].transform, theta[i]);
    evaluate_model(mc, tp, result);
}

struct LinModel
{
    Vec centre, offset;
    Mat J; // T x P
    // fwdmodel_linear.cc:126-181
    void ReCentre(const ModelCtx &mc, const Vec &about)
    {
        centre = about;
        evaluate_fabber(mc, centre, offset);
        if (!all_finite(offset))
            throw InternalError(FABBER_VOX_NONFINITE_OFFSET,
                "LinearizedFwdModel::ReCentre: Non-finite values found in offset");
        J = Mat(mc.T, mc.P);
        Vec c2, c3, o2, o3;
        for (int i = 0; i < mc.P; i++)
        {
            real delta = centre[i] * 1e-5;
            if (delta < 0)
                delta = -delta;
            if (delta < 1e-10)
                delta = 1e-10;
            c3 = centre;
            c2 = centre;
            c2[i] += delta;
            c3[i] -= delta;
            evaluate_fabber(mc, c2, o2);
            evaluate_fabber(mc, c3, o3);
            real den = c2[i] - c3[i];
            for (int t = 0; t < mc.T; t++)
                J(t, i) = (o2[t] - o3[t]) / den;
        }
        if (!all_finite(J))
            throw InternalError(FABBER_VOX_NONFINITE_JACOBIAN,
                "LinearizedFwdModel::ReCentre: Non-finite values found in jacobian");
    }
};

// ---------------------------------------------------------------------------------------------
// Noise models
// ---------------------------------------------------------------------------------------------
struct Gamma
{
    real b, c;
};

struct NoiseParams
{
    std::vector<Gamma> phis;
    // AR(1): alpha MVN (size 2) and the marginal tridiagonal Q = M00 + M10 E[a] + M20 E[a^2]
    MVN alpha;
    Vec q_diag, q_off;
    // AR(1), two echoes: alpha MVN of size 2 / 3 / 4 and one marginal per echo, sparse symmetric
    // ((row, col) with row >= col stands for both triangles)
    typedef std::map<std::pair<int, int>, real> SymSparse;
    std::vector<SymSparse> Qm;
    NoiseParams()
        : alpha(2)
    {
    }
};

struct NoiseModel
{
    const fabber_cuda_vb_problem *prob;
    int T, P;
    bool ar, ar2;
    int nPhis, nAlphas;
    std::vector<Vec> Qis; // white: 0/1 diagonal masks per phi (noisemodel_white.cc:166-226)

    void init(const fabber_cuda_vb_problem *p)
    {
        prob = p;
        T = p->n_times;
        P = p->model.n_params;
        ar = p->noise_type == FABBER_NOISE_AR1;
        nPhis = ar ? (p->n_phis == 2 ? 2 : 1) : p->n_phis;
        ar2 = ar && nPhis == 2;
        nAlphas = 2 + (ar2 ? p->ar_cross_terms : 0); // Ar1cNoiseModel::NumAlphas, noisemodel_ar.cc:367-377
        if (!ar)
        {
            Qis.assign(nPhis, Vec(T, 0.0));
            for (int d = 0; d < T; d++)
            {
                int phi = p->phi_pattern ? p->phi_pattern[d] : 0;
                bool masked = p->time_masked && p->time_masked[d];
                if (!masked)
                    Qis[phi][d] = 1;
            }
        }
    }

    // k = data - offset + J*(centre - means)
    Vec calc_k(const MVN &theta, const LinModel &lin, const Vec &data) const
    {
        Vec d(P);
        for (int i = 0; i < P; i++)
            d[i] = lin.centre[i] - theta.means[i];
        Vec Jd = mulv(lin.J, d);
        Vec k(T);
        for (int t = 0; t < T; t++)
            k[t] = data[t] - lin.offset[t] + Jd[t];
        return k;
    }

    // ------------------------------ white ------------------------------------------------
    // (Sigma * J' * Q * J).Trace()  evaluated left to right as NEWMAT does
    static real trace_SJtQJ(const Mat &Sigma, const Mat &J, const Vec &q)
    {
        Mat SJt = mul(Sigma, transpose(J)); // P x T
        for (int i = 0; i < SJt.r; i++)
            for (int t = 0; t < SJt.c; t++)
                SJt(i, t) *= q[t];
        return trace(mul(SJt, J));
    }

    void white_update_noise(NoiseParams &post, const NoiseParams &prior, const MVN &theta,
        const LinModel &lin, const Vec &data) const
    {
        Vec k = calc_k(theta, lin, data);
        for (int i = 0; i < nPhis; i++)
        {
            const Vec &Qi = Qis[i];
            real kqk = 0;
            for (int t = 0; t < T; t++)
                kqk += k[t] * Qi[t] * k[t];
            real tmp = kqk + trace_SJtQJ(theta.GetCovariance(), lin.J, Qi);
            post.phis[i].b = 1 / (tmp * 0.5 + 1 / prior.phis[i].b);
            real nTimes = 0;
            for (int t = 0; t < T; t++)
                nTimes += Qi[t];
            post.phis[i].c = (nTimes - 1) * 0.5 + prior.phis[i].c;
            if (prob->locked_noise_stdev > 0)
                post.phis[i].b = 1 / post.phis[i].c / prob->locked_noise_stdev / prob->locked_noise_stdev;
        }
    }

    void white_update_theta(const NoiseParams &noise, MVN &theta, const MVN &thetaPrior,
        const LinModel &lin, const Vec &data, float LMalpha) const
    {
        const Vec &ml = lin.centre;
        const Vec &gml = lin.offset;
        const Mat &J = lin.J;
        Vec X(T, 0.0);
        for (int i = 0; i < nPhis; i++)
        {
            real mean = noise.phis[i].b * noise.phis[i].c;
            for (int t = 0; t < T; t++)
                X[t] += Qis[i][t] * mean;
        }
        // Ltmp << J.t() * X * J  (lower triangle kept)
        Mat JtX = transpose(J);
        for (int i = 0; i < P; i++)
            for (int t = 0; t < T; t++)
                JtX(i, t) *= X[t];
        Mat Ltmp = mul(JtX, J);
        for (int i = 0; i < P; i++)
            for (int j = i + 1; j < P; j++)
                Ltmp(i, j) = Ltmp(j, i);
        theta.SetPrecisions(add(thetaPrior.GetPrecisions(), Ltmp));

        Vec Jml = mulv(J, ml);
        Vec resid(T);
        for (int t = 0; t < T; t++)
            resid[t] = data[t] - gml[t] + Jml[t];
        Vec mTmp = mulv(JtX, resid);
        Vec P0m0 = mulv(thetaPrior.GetPrecisions(), thetaPrior.means);
        if (LMalpha <= 0.0)
        {
            Vec rhs(P);
            for (int i = 0; i < P; i++)
                rhs[i] = mTmp[i] + P0m0[i];
            theta.means = mulv(theta.GetCovariance(), rhs);
        }
        else
        {
            Mat prec = theta.GetPrecisions();
            Vec r2(T);
            for (int t = 0; t < T; t++)
                r2[t] = data[t] - gml[t];
            Vec JtXr = mulv(JtX, r2);
            Vec P0ml = mulv(thetaPrior.GetPrecisions(), ml);
            Vec Delta(P);
            for (int i = 0; i < P; i++)
                Delta[i] = JtXr[i] + P0m0[i] - P0ml[i];
            Mat damped = prec;
            for (int i = 0; i < P; i++)
                damped(i, i) = prec(i, i) + (real)LMalpha * prec(i, i);
            try
            {
                Vec step = mulv(inverse(damped), Delta);
                for (int i = 0; i < P; i++)
                    theta.means[i] = ml[i] + step[i];
            }
            catch (SingularError &)
            {
                // WARN_ONCE("matrix was singular in LM update") - means left unchanged
            }
        }
    }

    real white_free_energy(const NoiseParams &noise, const NoiseParams &noisePrior,
        const MVN &theta, const MVN &thetaPrior, const LinModel &lin, const Vec &data) const
    {
        const Mat &J = lin.J;
        Vec k = calc_k(theta, lin, data);
        const Mat &Linv = theta.GetCovariance();
        int n_masked = 0;
        if (prob->time_masked)
            for (int t = 0; t < T; t++)
                n_masked += prob->time_masked[t] ? 1 : 0;
        int nTimes = T - n_masked;
        int nTheta = P;

        real expectedLogThetaDist
            = +0.5 * log_determinant(theta.GetPrecisions()).logval - 0.5 * nTheta * (std::log(2 * M_PI) + 1);
        real expectedLogPhiDist = 0;
        real parts[10] = { 0 };
        for (int i = 0; i < nPhis; i++)
        {
            real si = noise.phis[i].b, ci = noise.phis[i].c;
            real siPrior = noisePrior.phis[i].b, ciPrior = noisePrior.phis[i].c;
            expectedLogPhiDist
                += -gammaln(ci) - ci * std::log(si) - ci + (ci - 1) * (digamma_fsl(ci) + std::log(si));
            real qtrace = 0;
            for (int t = 0; t < T; t++)
                qtrace += Qis[i][t];
            parts[0] += (digamma_fsl(ci) + std::log(si)) * (qtrace * 0.5 + ciPrior - 1);
            parts[9] += -gammaln(ciPrior) - ciPrior * std::log(siPrior) - si * ci / siPrior;
            real kk = 0;
            for (int t = 0; t < T; t++)
            {
                real ki = Qis[i][t] * k[t];
                kk += ki * ki;
            }
            Mat Ji = J;
            for (int t = 0; t < T; t++)
                for (int p = 0; p < P; p++)
                    Ji(t, p) *= Qis[i][t];
            parts[2] += -0.5 * si * ci * kk - 0.5 * trace(mul(mul(transpose(Ji), Ji), Linv));
        }
        parts[3] = +0.5 * log_determinant(thetaPrior.GetPrecisions()).logval
            - 0.5 * nTimes * std::log(2 * M_PI) - 0.5 * nTheta * std::log(2 * M_PI);
        Vec dm(P);
        for (int i = 0; i < P; i++)
            dm[i] = theta.means[i] - thetaPrior.means[i];
        Vec Pdm = mulv(thetaPrior.GetPrecisions(), dm);
        real q = 0;
        for (int i = 0; i < P; i++)
            q += dm[i] * Pdm[i];
        parts[4] = -0.5 * q;
        parts[5] = -0.5 * trace(mul(Linv, thetaPrior.GetPrecisions()));
        real F = -expectedLogThetaDist - expectedLogPhiDist;
        for (int i = 0; i < 10; i++)
            F += parts[i];
        if (!(F - F == 0))
            throw InternalError(FABBER_VOX_NONFINITE_F, "WhiteNoiseModel::Non-finite free energy!");
        return F;
    }

    // ------------------------------ AR(1), num-echoes=1 -----------------------------------
    // Symmetric tridiagonal "alpha matrices" (noisemodel_ar.cc:130-179, nPhis == 1):
    //   M00 = diag(0,1,..,1)   M20 = diag(1,..,1,0)   M10 = -1 on the first off-diagonals.
    // Stored as (diag, off) so products skip exact zeros (bit-identical to the dense sums).
    static real quad_tri(const Vec &k, const Vec &d, const Vec &e)
    {
        int T = (int)k.size();
        real s = 0;
        for (int i = 0; i < T; i++)
        {
            real mk = 0;
            if (i > 0)
                mk += e[i - 1] * k[i - 1];
            mk += d[i] * k[i];
            if (i + 1 < T)
                mk += e[i] * k[i + 1];
            s += k[i] * mk;
        }
        return s;
    }
    // J' * M * J for tridiagonal M (P x P)
    static Mat JtMJ_tri(const Mat &J, const Vec &d, const Vec &e)
    {
        int T = J.r, P = J.c;
        Mat MJ(T, P);
        for (int i = 0; i < T; i++)
            for (int p = 0; p < P; p++)
            {
                real s = 0;
                if (i > 0)
                    s += e[i - 1] * J(i - 1, p);
                s += d[i] * J(i, p);
                if (i + 1 < T)
                    s += e[i] * J(i + 1, p);
                MJ(i, p) = s;
            }
        return mul(transpose(J), MJ);
    }
    // OperatorKLJ (noisemodel_ar.cc:433-445): k'Mk + Trace(L.i() * J' M J)
    static real op_klj(const Vec &k, const Mat &Sigma, const Mat &J, const Vec &d, const Vec &e)
    {
        return quad_tri(k, d, e) + trace(mul(Sigma, JtMJ_tri(J, d, e)));
    }

    void ar_alpha_matrices(Vec &d00, Vec &d20, Vec &e10) const
    {
        d00.assign(T, 1.0);
        d00[0] = 0.0;
        d20.assign(T, 1.0);
        d20[T - 1] = 0.0;
        e10.assign(T > 0 ? T - 1 : 0, -1.0);
    }

    // Ar1cMatrixCache::Update (noisemodel_ar.cc:197-222), nAlphas == 2
    void ar_update_marginal(NoiseParams &np) const
    {
        Vec d00, d20, e10;
        ar_alpha_matrices(d00, d20, e10);
        real a = np.alpha.means[0];
        real covarPlus = np.alpha.GetCovariance()(0, 0) + a * a;
        np.q_diag.assign(T, 0.0);
        np.q_off.assign(T > 0 ? T - 1 : 0, 0.0);
        for (int t = 0; t < T; t++)
            np.q_diag[t] = d00[t] + 0.0 * a + d20[t] * covarPlus;
        for (int t = 0; t + 1 < T; t++)
            np.q_off[t] = 0.0 + e10[t] * a + 0.0 * covarPlus;
    }

    void ar_precalculate(NoiseParams &post, const NoiseParams &prior) const
    {
        ar_update_marginal(post);
        post.phis[0].c = prior.phis[0].c + (T - 1) * 0.5; // noisemodel_ar.cc:765-768
    }

    void ar_update_noise(NoiseParams &post, const NoiseParams &prior, const MVN &theta,
        const LinModel &lin, const Vec &data) const
    {
        // UpdateAlpha (noisemodel_ar.cc:447-528)
        Vec d00, d20, e10;
        ar_alpha_matrices(d00, d20, e10);
        Vec zeros_off(T > 0 ? T - 1 : 0, 0.0), zeros_diag(T, 0.0);
        {
            Vec k = calc_k(theta, lin, data);
            real si_ci = post.phis[0].b * post.phis[0].c;
            const Mat &Sigma = theta.GetCovariance(); // OpKLJ uses L.i() with L = precisions
            Mat alphaPrec = prior.alpha.GetPrecisions();
            alphaPrec(0, 0) += si_ci * op_klj(k, Sigma, lin.J, d20, zeros_off);
            post.alpha.SetPrecisions(alphaPrec);
            if (!all_finite(alphaPrec))
                throw InternalError(FABBER_VOX_NONFINITE_F,
                    "Ar1cNoiseModel::UpdateAlpha Non-finite values in alpha precisions!");
            Mat chk = inverse(alphaPrec);
            real mn = chk(0, 0);
            for (int i = 1; i < chk.r; i++)
                mn = std::min(mn, chk(i, i));
            if (mn < 0)
                throw InternalError(
                    FABBER_VOX_AR_NEG_VARIANCE, "Ar1cNoiseModel::UpdateAlpha Negative variance!");
            Vec tmp = mulv(prior.alpha.GetPrecisions(), prior.alpha.means);
            tmp[0] += -0.5 * si_ci * op_klj(k, Sigma, lin.J, zeros_diag, e10);
            post.alpha.means = mulv(post.alpha.GetCovariance(), tmp);
            ar_update_marginal(post);
        }
        // UpdatePhi (noisemodel_ar.cc:530-556)
        {
            Vec k = calc_k(theta, lin, data);
            real tmp = quad_tri(k, post.q_diag, post.q_off)
                + trace(mul(theta.GetCovariance(), JtMJ_tri(lin.J, post.q_diag, post.q_off)));
            post.phis[0].b = 1 / (tmp * 0.5 + 1 / prior.phis[0].b);
            post.phis[0].c = (T - 1) * 0.5 + prior.phis[0].c;
        }
    }

    void ar_update_theta(const NoiseParams &noise, MVN &theta, const MVN &thetaPrior,
        const LinModel &lin, const Vec &data) const
    {
        real si_ci = noise.phis[0].b * noise.phis[0].c;
        Vec xd(T), xe(T > 0 ? T - 1 : 0);
        for (int t = 0; t < T; t++)
            xd[t] = si_ci * noise.q_diag[t];
        for (int t = 0; t + 1 < T; t++)
            xe[t] = si_ci * noise.q_off[t];
        Mat Ltmp = JtMJ_tri(lin.J, xd, xe);
        for (int i = 0; i < P; i++)
            for (int j = i + 1; j < P; j++)
                Ltmp(i, j) = Ltmp(j, i);
        theta.SetPrecisions(add(thetaPrior.GetPrecisions(), Ltmp));
        Vec Jml = mulv(lin.J, lin.centre);
        Vec resid(T);
        for (int t = 0; t < T; t++)
            resid[t] = data[t] - lin.offset[t] + Jml[t];
        // J' * X * resid : (J'X) first, as NEWMAT evaluates left to right
        Mat Xm(T, 1);
        for (int i = 0; i < T; i++)
        {
            real s = 0;
            if (i > 0)
                s += xe[i - 1] * resid[i - 1];
            s += xd[i] * resid[i];
            if (i + 1 < T)
                s += xe[i] * resid[i + 1];
            Xm(i, 0) = s;
        }
        Vec mTmp(P);
        for (int p = 0; p < P; p++)
        {
            real s = 0;
            for (int t = 0; t < T; t++)
                s += lin.J(t, p) * Xm(t, 0);
            mTmp[p] = s;
        }
        Vec P0m0 = mulv(thetaPrior.GetPrecisions(), thetaPrior.means);
        Vec rhs(P);
        for (int i = 0; i < P; i++)
            rhs[i] = mTmp[i] + P0m0[i];
        theta.means = mulv(theta.GetCovariance(), rhs);
    }

    real ar_free_energy(const NoiseParams &post, const NoiseParams &prior, const MVN &theta,
        const MVN &thetaPrior, const LinModel &lin, const Vec &data) const
    {
        Vec k = calc_k(theta, lin, data);
        const Mat &Linv = theta.GetCovariance();
        const Gamma &phi1 = post.phis[0];
        real w = phi1.b * phi1.c;
        Vec qd(T), qe(T > 0 ? T - 1 : 0);
        for (int t = 0; t < T; t++)
            qd[t] = post.q_diag[t] * w;
        for (int t = 0; t + 1 < T; t++)
            qe[t] = post.q_off[t] * w;
        int nTimes = T;
        int nTheta = P;
        int nAlphas = 2;
        real expectedLogAlphaDist = +0.5 * log_determinant(post.alpha.GetPrecisions()).logval
            - 0.5 * nAlphas * (std::log(2 * M_PI) + 1);
        real expectedLogThetaDist = +0.5 * log_determinant(theta.GetPrecisions()).logval
            - 0.5 * nTheta * (std::log(2 * M_PI) + 1);
        real expectedLogPhiDist = 0;
        real parts[10] = { 0 };
        {
            real si = phi1.b, ci = phi1.c;
            real siPrior = prior.phis[0].b, ciPrior = prior.phis[0].c;
            expectedLogPhiDist
                += -gammaln(ci) - ci * std::log(si) - ci + (ci - 1) * (digamma_fsl(ci) + std::log(si));
            parts[0] += (digamma_fsl(ci) + std::log(si)) * ((nTimes - 1) * 0.5 + ciPrior - 1);
            parts[9] += -2 * gammaln(ciPrior) - 2 * ciPrior * std::log(siPrior) - si * ci / siPrior;
        }
        parts[1] = -std::log(2 * M_PI) * (nTimes - 1 + 0.5 * nAlphas + 0.5 * nTheta);
        parts[2] = -0.5 * quad_tri(k, qd, qe) - 0.5 * trace(mul(JtMJ_tri(lin.J, qd, qe), Linv));
        parts[3] = +0.5 * log_determinant(thetaPrior.GetPrecisions()).logval;
        Vec dm(P);
        for (int i = 0; i < P; i++)
            dm[i] = theta.means[i] - thetaPrior.means[i];
        Vec Pdm = mulv(thetaPrior.GetPrecisions(), dm);
        real q = 0;
        for (int i = 0; i < P; i++)
            q += dm[i] * Pdm[i];
        parts[4] = -0.5 * q;
        parts[5] = -0.5 * trace(mul(Linv, thetaPrior.GetPrecisions()));
        parts[6] = +0.5 * log_determinant(prior.alpha.GetPrecisions()).logval;
        Vec da(2);
        for (int i = 0; i < 2; i++)
            da[i] = post.alpha.means[i] - prior.alpha.means[i];
        Vec Pda = mulv(prior.alpha.GetPrecisions(), da);
        parts[7] = -0.5 * (da[0] * Pda[0] + da[1] * Pda[1]);
        parts[8] = -0.5 * trace(mul(post.alpha.GetCovariance(), prior.alpha.GetPrecisions()));
        real F = -expectedLogAlphaDist - expectedLogThetaDist - expectedLogPhiDist;
        for (int i = 0; i < 10; i++)
            F += parts[i];
        if (!(F - F == 0))
            throw InternalError(
                FABBER_VOX_NONFINITE_F, "Ar1cNoiseModel::CalcFreeEnergy Non-finite free energy!");
        return F;
    }

    // ------------------------------ AR(1), num-echoes=2 -----------------------------------
    // The series interleaves the echoes, TE1 TE2 TE1 TE2 .. (noisemodel_ar.cc:126-129), nTimes = T / 2
    // samples each. Every alpha matrix is ONE diagonal line of nTimes-1 entries of +-1, reflected to keep the
    // matrix symmetric (:108-179). n = 1, 2 is the echo; (a12pow, a34pow) the powers of the echo's own alpha
    // and of its cross-term alpha the matrix multiplies in the marginal.
    typedef NoiseParams::SymSparse SymSparse;
    SymSparse ar2_matrix(int n, int a12pow, int a34pow) const
    {
        const int nTimes = T / nPhis;
        int row, col; // 1-based, as in the reference
        switch (a12pow * 10 + a34pow)
        {
        case 0:
            row = col = 1 + nPhis;
            break;
        case 10:
            row = 1;
            col = 1 + nPhis;
            break;
        case 20:
            row = col = 1;
            break;
        case 1:
            row = 4;
            col = 3;
            break;
        case 11:
            row = 4;
            col = 1;
            break;
        case 2:
            row = col = 4;
            break;
        default:
            throw InternalError(FABBER_VOX_SINGULAR, "Ar1cMatrixCache::Update Invalid row/col");
        }
        const real value = (a12pow + a34pow == 1) ? -1 : 1;
        if (n == 2)
        {
            row = row - 1 + 2 * (row % 2); // 2n->2n-1, 2n-1->2n: the other echo
            col = col - 1 + 2 * (col % 2);
        }
        SymSparse m;
        for (int count = 0; count < nTimes - 1; count++, row += nPhis, col += nPhis)
            m[std::make_pair(std::max(row, col) - 1, std::min(row, col) - 1)] = value;
        return m;
    }
    static void sym_axpy(SymSparse &dst, real w, const SymSparse &src)
    {
        for (SymSparse::const_iterator e = src.begin(); e != src.end(); ++e)
            dst[e->first] += w * e->second;
    }
    static Vec sym_apply(const SymSparse &M, const Vec &x)
    {
        Vec y(x.size(), 0.0);
        for (SymSparse::const_iterator e = M.begin(); e != M.end(); ++e)
        {
            const int r = e->first.first, c = e->first.second;
            y[r] += e->second * x[c];
            if (r != c)
                y[c] += e->second * x[r];
        }
        return y;
    }
    static real sym_quad(const SymSparse &M, const Vec &k)
    {
        Vec y = sym_apply(M, k);
        real s = 0;
        for (size_t i = 0; i < k.size(); i++)
            s += k[i] * y[i];
        return s;
    }
    static Mat sym_JtMJ(const Mat &J, const SymSparse &M)
    {
        Mat MJ(J.r, J.c);
        Vec col(J.r);
        for (int p = 0; p < J.c; p++)
        {
            for (int t = 0; t < J.r; t++)
                col[t] = J(t, p);
            Vec y = sym_apply(M, col);
            for (int t = 0; t < J.r; t++)
                MJ(t, p) = y[t];
        }
        return mul(transpose(J), MJ);
    }
    static real op_klj2(const Vec &k, const Mat &Sigma, const Mat &J, const SymSparse &M)
    {
        return sym_quad(M, k) + trace(mul(Sigma, sym_JtMJ(J, M)));
    }

    // Ar1cMatrixCache::Update (noisemodel_ar.cc:197-222)
    void ar2_update_marginal(NoiseParams &np) const
    {
        const Mat &cov = np.alpha.GetCovariance();
        Mat covarPlus(nAlphas, nAlphas);
        for (int i = 0; i < nAlphas; i++)
            for (int j = 0; j < nAlphas; j++)
                covarPlus(i, j) = cov(i, j) + np.alpha.means[i] * np.alpha.means[j];
        np.Qm.assign(nPhis, SymSparse());
        for (int n = 1; n <= nPhis; n++)
        {
            SymSparse &Q = np.Qm[n - 1];
            sym_axpy(Q, 1.0, ar2_matrix(n, 0, 0));
            sym_axpy(Q, np.alpha.means[n - 1], ar2_matrix(n, 1, 0));
            sym_axpy(Q, covarPlus(n - 1, n - 1), ar2_matrix(n, 2, 0));
            if (nAlphas >= 3)
            {
                const int Tn = (nAlphas == 4) ? 2 + n : 3; // 1-based index of this echo's cross-term alpha
                sym_axpy(Q, np.alpha.means[Tn - 1], ar2_matrix(n, 0, 1));
                sym_axpy(Q, covarPlus(n - 1, Tn - 1), ar2_matrix(n, 1, 1));
                sym_axpy(Q, covarPlus(Tn - 1, Tn - 1), ar2_matrix(n, 0, 2));
            }
        }
    }

    void ar2_precalculate(NoiseParams &post, const NoiseParams &prior) const
    {
        const int nTimes = T / nPhis;
        ar2_update_marginal(post);
        for (int i = 0; i < nPhis; i++)
            post.phis[i].c = prior.phis[i].c + (nTimes - 1) * 0.5; // noisemodel_ar.cc:765-768
    }

    void ar2_update_noise(NoiseParams &post, const NoiseParams &prior, const MVN &theta,
        const LinModel &lin, const Vec &data) const
    {
        const int nTimes = T / nPhis;
        // UpdateAlpha (noisemodel_ar.cc:447-528)
        {
            Vec k = calc_k(theta, lin, data);
            Vec si_ci(nPhis);
            for (int i = 0; i < nPhis; i++)
                si_ci[i] = post.phis[i].b * post.phis[i].c;
            const Mat &Sigma = theta.GetCovariance(); // OpKLJ uses L.i() with L = precisions
            const Mat &J = lin.J;
            Mat alphaPrec = prior.alpha.GetPrecisions();
            const int Tx = nAlphas; // "use same code for nAlphas == 3 or 4" (:470)
            for (int i = 1; i <= nPhis; i++)
                alphaPrec(i - 1, i - 1) += si_ci[i - 1] * op_klj2(k, Sigma, J, ar2_matrix(i, 2, 0));
            if (Tx > 2)
            {
                real x = 0.5 * si_ci[0] * op_klj2(k, Sigma, J, ar2_matrix(1, 1, 1));
                alphaPrec(2, 0) += x;
                alphaPrec(0, 2) = alphaPrec(2, 0);
                x = 0.5 * si_ci[1] * op_klj2(k, Sigma, J, ar2_matrix(2, 1, 1));
                alphaPrec(Tx - 1, 1) += x;
                alphaPrec(1, Tx - 1) = alphaPrec(Tx - 1, 1);
                alphaPrec(2, 2) += si_ci[0] * op_klj2(k, Sigma, J, ar2_matrix(1, 0, 2));
                alphaPrec(Tx - 1, Tx - 1) += si_ci[1] * op_klj2(k, Sigma, J, ar2_matrix(2, 0, 2));
            }
            post.alpha.SetPrecisions(alphaPrec);
            if (!all_finite(alphaPrec))
                throw InternalError(FABBER_VOX_NONFINITE_F,
                    "Ar1cNoiseModel::UpdateAlpha Non-finite values in alpha precisions!");
            Mat chk = inverse(alphaPrec);
            real mn = chk(0, 0);
            for (int i = 1; i < chk.r; i++)
                mn = std::min(mn, chk(i, i));
            if (mn < 0)
                throw InternalError(
                    FABBER_VOX_AR_NEG_VARIANCE, "Ar1cNoiseModel::UpdateAlpha Negative variance!");
            Vec tmp = mulv(prior.alpha.GetPrecisions(), prior.alpha.means);
            for (int i = 1; i <= nPhis; i++)
                tmp[i - 1] += -0.5 * si_ci[i - 1] * op_klj2(k, Sigma, J, ar2_matrix(i, 1, 0));
            if (Tx > 2)
            {
                tmp[2] += -0.5 * si_ci[0] * op_klj2(k, Sigma, J, ar2_matrix(1, 0, 1));
                tmp[Tx - 1] += -0.5 * si_ci[1] * op_klj2(k, Sigma, J, ar2_matrix(2, 0, 1));
            }
            post.alpha.means = mulv(post.alpha.GetCovariance(), tmp);
            ar2_update_marginal(post);
        }
        // UpdatePhi (noisemodel_ar.cc:530-556)
        {
            Vec k = calc_k(theta, lin, data);
            for (int i = 0; i < nPhis; i++)
            {
                real tmp = sym_quad(post.Qm[i], k) + trace(mul(theta.GetCovariance(), sym_JtMJ(lin.J, post.Qm[i])));
                post.phis[i].b = 1 / (tmp * 0.5 + 1 / prior.phis[i].b);
                post.phis[i].c = (nTimes - 1) * 0.5 + prior.phis[i].c;
            }
        }
    }

    SymSparse ar2_weighted_marginals(const NoiseParams &noise) const
    {
        SymSparse X;
        for (int i = 0; i < nPhis; i++)
            sym_axpy(X, noise.phis[i].b * noise.phis[i].c, noise.Qm[i]);
        return X;
    }

    void ar2_update_theta(const NoiseParams &noise, MVN &theta, const MVN &thetaPrior,
        const LinModel &lin, const Vec &data) const
    {
        // noisemodel_ar.cc:558-610 (LMalpha is ignored by the AR model)
        SymSparse X = ar2_weighted_marginals(noise);
        Mat Ltmp = sym_JtMJ(lin.J, X);
        for (int i = 0; i < P; i++)
            for (int j = i + 1; j < P; j++)
                Ltmp(i, j) = Ltmp(j, i);
        theta.SetPrecisions(add(thetaPrior.GetPrecisions(), Ltmp));
        Vec Jml = mulv(lin.J, lin.centre);
        Vec resid(T);
        for (int t = 0; t < T; t++)
            resid[t] = data[t] - lin.offset[t] + Jml[t];
        Vec Xr = sym_apply(X, resid);
        Vec mTmp(P);
        for (int p = 0; p < P; p++)
        {
            real s = 0;
            for (int t = 0; t < T; t++)
                s += lin.J(t, p) * Xr[t];
            mTmp[p] = s;
        }
        Vec P0m0 = mulv(thetaPrior.GetPrecisions(), thetaPrior.means);
        Vec rhs(P);
        for (int i = 0; i < P; i++)
            rhs[i] = mTmp[i] + P0m0[i];
        theta.means = mulv(theta.GetCovariance(), rhs);
    }

    real ar2_free_energy(const NoiseParams &post, const NoiseParams &prior, const MVN &theta,
        const MVN &thetaPrior, const LinModel &lin, const Vec &data) const
    {
        // noisemodel_ar.cc:643-747
        Vec k = calc_k(theta, lin, data);
        const Mat &Linv = theta.GetCovariance();
        SymSparse Qsum = ar2_weighted_marginals(post);
        const int nTimes = T / nPhis;
        const int nTheta = P;
        real expectedLogAlphaDist = +0.5 * log_determinant(post.alpha.GetPrecisions()).logval
            - 0.5 * nAlphas * (std::log(2 * M_PI) + 1);
        real expectedLogThetaDist = +0.5 * log_determinant(theta.GetPrecisions()).logval
            - 0.5 * nTheta * (std::log(2 * M_PI) + 1);
        real expectedLogPhiDist = 0;
        real parts[10] = { 0 };
        for (int i = 0; i < nPhis; i++)
        {
            real si = post.phis[i].b, ci = post.phis[i].c;
            real siPrior = prior.phis[i].b, ciPrior = prior.phis[i].c;
            expectedLogPhiDist
                += -gammaln(ci) - ci * std::log(si) - ci + (ci - 1) * (digamma_fsl(ci) + std::log(si));
            parts[0] += (digamma_fsl(ci) + std::log(si)) * ((nTimes - 1) * 0.5 + ciPrior - 1);
            parts[9] += -2 * gammaln(ciPrior) - 2 * ciPrior * std::log(siPrior) - si * ci / siPrior;
        }
        parts[1] = -std::log(2 * M_PI) * (nTimes - 1 + 0.5 * nAlphas + 0.5 * nTheta);
        parts[2] = -0.5 * sym_quad(Qsum, k) - 0.5 * trace(mul(sym_JtMJ(lin.J, Qsum), Linv));
        parts[3] = +0.5 * log_determinant(thetaPrior.GetPrecisions()).logval;
        Vec dm(P);
        for (int i = 0; i < P; i++)
            dm[i] = theta.means[i] - thetaPrior.means[i];
        Vec Pdm = mulv(thetaPrior.GetPrecisions(), dm);
        real q = 0;
        for (int i = 0; i < P; i++)
            q += dm[i] * Pdm[i];
        parts[4] = -0.5 * q;
        parts[5] = -0.5 * trace(mul(Linv, thetaPrior.GetPrecisions()));
        parts[6] = +0.5 * log_determinant(prior.alpha.GetPrecisions()).logval;
        Vec da(nAlphas);
        for (int i = 0; i < nAlphas; i++)
            da[i] = post.alpha.means[i] - prior.alpha.means[i];
        Vec Pda = mulv(prior.alpha.GetPrecisions(), da);
        real qa = 0;
        for (int i = 0; i < nAlphas; i++)
            qa += da[i] * Pda[i];
        parts[7] = -0.5 * qa;
        parts[8] = -0.5 * trace(mul(post.alpha.GetCovariance(), prior.alpha.GetPrecisions()));
        real F = -expectedLogAlphaDist - expectedLogThetaDist - expectedLogPhiDist;
        for (int i = 0; i < 10; i++)
            F += parts[i];
        if (!(F - F == 0))
            throw InternalError(
                FABBER_VOX_NONFINITE_F, "Ar1cNoiseModel::CalcFreeEnergy Non-finite free energy!");
        return F;
    }

    // ------------------------------ dispatch ---------------------------------------------
    void hardcoded_initial(NoiseParams &prior, NoiseParams &post) const
    {
        prior.phis.resize(nPhis);
        post.phis.resize(nPhis);
        for (int i = 0; i < nPhis; i++)
        {
            prior.phis[i].b = prob->noise_prior_b[i];
            prior.phis[i].c = prob->noise_prior_c[i];
            post.phis[i].b = prob->noise_post_b[i];
            post.phis[i].c = prob->noise_post_c[i];
        }
        if (ar)
        {
            prior.alpha = MVN(nAlphas);
            post.alpha = MVN(nAlphas);
            Mat p = Mat::identity(nAlphas);
            for (int i = 0; i < nAlphas; i++)
                p(i, i) = prob->ar_alpha_prior_prec;
            prior.alpha.SetPrecisions(p);
            post.alpha.SetPrecisions(p);
        }
    }
    void precalculate(NoiseParams &post, const NoiseParams &prior) const
    {
        if (ar2)
            ar2_precalculate(post, prior);
        else if (ar)
            ar_precalculate(post, prior);
    }
    void update_theta(const NoiseParams &noise, MVN &theta, const MVN &thetaPrior,
        const LinModel &lin, const Vec &data, float LMalpha) const
    {
        if (ar2)
            ar2_update_theta(noise, theta, thetaPrior, lin, data);
        else if (ar)
            ar_update_theta(noise, theta, thetaPrior, lin, data);
        else
            white_update_theta(noise, theta, thetaPrior, lin, data, LMalpha);
    }
    void update_noise(NoiseParams &post, const NoiseParams &prior, const MVN &theta,
        const LinModel &lin, const Vec &data) const
    {
        if (ar2)
            ar2_update_noise(post, prior, theta, lin, data);
        else if (ar)
            ar_update_noise(post, prior, theta, lin, data);
        else
            white_update_noise(post, prior, theta, lin, data);
    }
    real free_energy(const NoiseParams &post, const NoiseParams &prior, const MVN &theta,
        const MVN &thetaPrior, const LinModel &lin, const Vec &data) const
    {
        if (ar2)
            return ar2_free_energy(post, prior, theta, thetaPrior, lin, data);
        return ar ? ar_free_energy(post, prior, theta, thetaPrior, lin, data)
                  : white_free_energy(post, prior, theta, thetaPrior, lin, data);
    }
};

// ---------------------------------------------------------------------------------------------
// Convergence detectors (convergence.cc:34-378)
// ---------------------------------------------------------------------------------------------
struct Conv
{
    int type;
    int m_its, m_max_its;
    real m_prev_f, m_min_fchange;
    bool m_revert, m_save;
    int m_trials, m_max_trials;
    bool m_trialmode;
    // LM
    bool m_LM;
    real m_alpha, m_alphastart, m_alphamax;

    void Initialize(int type_, int max_its, real fchange, int max_trials)
    {
        type = type_;
        m_max_its = max_its;
        m_min_fchange = fchange;
        m_max_trials = max_trials;
        if (type == FABBER_CONV_TRIALMODE)
            m_max_its += 1; // convergence.cc:145
        Reset();
    }
    void Reset(real F = -99e99)
    {
        m_its = 0;
        m_prev_f = F;
        m_save = false;
        m_revert = false;
        m_trials = 0;
        m_trialmode = false;
        m_LM = false;
        m_alpha = 0.0;
        m_alphastart = 1e-6;
        m_alphamax = 1e6;
        if (type == FABBER_CONV_TRIALMODE || type == FABBER_CONV_LM)
            m_save = true;
    }
    bool NeedSave() const { return type == FABBER_CONV_MAXITS ? false : m_save; }
    bool NeedRevert() const { return type == FABBER_CONV_MAXITS ? false : m_revert; }
    float LMalpha() const { return type == FABBER_CONV_LM ? (float)m_alpha : 0.0f; }

    bool counting_test()
    {
        ++m_its;
        return m_its >= m_max_its;
    }
    bool fchange_test(real F)
    {
        real diff = F - m_prev_f;
        m_prev_f = F;
        diff = diff > 0 ? diff : -diff;
        if (diff < m_min_fchange)
            return true;
        return counting_test();
    }
    bool Test(real F)
    {
        switch (type)
        {
        case FABBER_CONV_MAXITS:
            return counting_test();
        case FABBER_CONV_FCHANGE:
            return fchange_test(F);
        case FABBER_CONV_FREDUCE:
        {
            real diff = F - m_prev_f;
            if (diff < 0)
            {
                m_revert = true;
                return true;
            }
            return fchange_test(F);
        }
        case FABBER_CONV_TRIALMODE:
            return trial_test(F);
        case FABBER_CONV_LM:
            return lm_test(F);
        }
        return true;
    }
    bool trial_test(real F)
    {
        real diff = F - m_prev_f;
        if (!m_trialmode)
        {
            if (diff < 0)
            {
                m_its = 1;
                m_trials = 1;
                m_trialmode = true;
                m_revert = true;
                m_save = false;
                return false;
            }
            real absdiff = diff > 0 ? diff : -diff;
            if (absdiff < m_min_fchange)
            {
                m_revert = false;
                m_save = false;
                return true;
            }
            m_save = true;
            m_revert = false;
            m_prev_f = F;
            ++m_its;
            return (m_its >= m_max_its);
        }
        ++m_trials;
        if (diff > 0)
        {
            real absdiff = diff > 0 ? diff : -diff;
            if (absdiff < m_min_fchange)
            {
                m_revert = false;
                m_save = false;
                return true;
            }
            m_trialmode = false;
            m_trials = 0;
            m_save = true;
            m_revert = false;
            m_prev_f = F;
            return false;
        }
        else if (m_trials >= m_max_trials)
        {
            m_save = false;
            m_revert = true;
            return true;
        }
        m_save = false;
        m_revert = false;
        return false;
    }
    bool lm_test(real F)
    {
        real diff = F - m_prev_f;
        real absdiff = diff;
        if (diff < 0)
            absdiff = -diff;
        if (!m_LM)
        {
            if (diff < 0)
            {
                m_LM = true;
                m_revert = true;
                m_alpha = m_alphastart;
                return false;
            }
            else if (absdiff < m_min_fchange)
            {
                m_revert = false;
                return true;
            }
            else if (m_its >= m_max_its)
            {
                m_revert = false;
                return true;
            }
            m_prev_f = F;
            ++m_its;
            return false;
        }
        if (diff > 0)
        {
            if (m_alpha == m_alphastart)
                m_LM = false;
            else
            {
                m_alpha /= 10;
                m_LM = true;
            }
            m_revert = false;
            m_prev_f = F;
            ++m_its;
            return false;
        }
        else if (m_alpha >= m_alphamax)
        {
            m_revert = true;
            return true;
        }
        else if (m_its >= m_max_its)
        {
            m_revert = false;
            return true;
        }
        m_alpha *= 10;
        m_revert = true;
        return false;
    }
};

// ---------------------------------------------------------------------------------------------
// Run context + priors (run_context.h, priors.cc)
// ---------------------------------------------------------------------------------------------
struct RunContext
{
    int it, v, nvoxels;
    std::vector<int> ignore_voxels;
    std::vector<MVN> fwd_prior, fwd_post;
    std::vector<NoiseParams> noise_prior, noise_post;
    std::vector<std::vector<int> > neighbours, neighbours2;
};

struct Prior
{
    const fabber_cuda_vb_problem *prob;
    const fabber_cuda_vb_buffers *buf;
    int idx;
    char type;
    real mean, prec, var;
    real aK; // spatial

    bool is_spatial() const { return type == 'M' || type == 'm' || type == 'P' || type == 'p'; }

    // SpatialPrior::CalculateaK priors.cc:221-344
    real CalculateaK(const RunContext &ctx) const
    {
        const int dims = prob->spatial_dims;
        real trace_term = 0.0, term2 = 0.0;
        for (int v = 1; v <= ctx.nvoxels; v++)
        {
            if (std::find(ctx.ignore_voxels.begin(), ctx.ignore_voxels.end(), v) != ctx.ignore_voxels.end())
                continue;
            real sigmaK = ctx.fwd_post[v - 1].GetCovariance()(idx, idx);
            int nn = (int)ctx.neighbours[v - 1].size();
            if (type == 'm')
                trace_term += sigmaK * dims * 2;
            else if (type == 'M')
                trace_term += sigmaK * (nn + 1e-8);
            else if (type == 'p')
                trace_term += sigmaK * (4 * dims * dims + 2 * dims);
            else
                trace_term += sigmaK * (nn * nn + nn);
            real wK = ctx.fwd_post[v - 1].means[idx];
            real SwK = 0.0;
            for (size_t j = 0; j < ctx.neighbours[v - 1].size(); j++)
                SwK += wK - ctx.fwd_post[ctx.neighbours[v - 1][j] - 1].means[idx];
            if (type == 'p' || type == 'm')
                SwK += wK * (dims * 2 - (real)ctx.neighbours[v - 1].size());
            if (type == 'm' || type == 'M')
                term2 += SwK * wK;
            else
                term2 += SwK * SwK;
        }
        real gk = 1 / (0.5 * trace_term + 0.5 * term2 + 1 / prob->spatial_q1);
        real hK = (ctx.nvoxels * 0.5 + prob->spatial_q2);
        real a = gk * hK;
        if (a < 1e-50)
            a = 1e-50;
        real aKMax = a * prob->spatial_speed;
        if (aKMax < 0.5)
            aKMax = 0.5;
        if ((prob->spatial_speed > 0) && (a > aKMax))
            a = aKMax;
        return a;
    }

    real ApplyToMVN(MVN *prior, RunContext &ctx)
    {
        if (type == 'N' || type == '-' || type == 'I')
        {
            // priors.cc:108-117, :133-142
            prior->means[idx] = (type == 'I') ? buf->image_prior[idx][ctx.v - 1] : mean;
            Mat p = prior->GetPrecisions();
            p(idx, idx) = prec;
            prior->SetPrecisions(p);
            return 0;
        }
        if (type == 'A')
        {
            // priors.cc:150-181
            Mat cov = prior->GetCovariance();
            real post_mean = ctx.fwd_post[ctx.v - 1].means[idx];
            real post_cov = ctx.fwd_post[ctx.v - 1].GetCovariance()(idx, idx);
            real new_cov = post_mean * post_mean + post_cov;
            if (ctx.it == 0)
            {
                cov(idx, idx) = var;
                prior->means[idx] = mean;
            }
            else
                cov(idx, idx) = new_cov;
            prior->SetCovariance(cov);
            real b = 2 / new_cov;
            return -1.5 * (std::log(b) + digamma_fsl(0.5)) - 0.5 - gammaln(0.5) - 0.5 * std::log(b);
        }
        // SpatialPrior::ApplyToMVN priors.cc:346-488
        if (ctx.v == 1 && (ctx.it > 0 || prob->update_first_iter))
            aK = CalculateaK(ctx);
        int nn = (int)ctx.neighbours[ctx.v - 1].size();
        real contrib_nn = 0.0;
        for (size_t j = 0; j < ctx.neighbours[ctx.v - 1].size(); j++)
            contrib_nn += ctx.fwd_post[ctx.neighbours[ctx.v - 1][j] - 1].means[idx];
        int nn2 = (int)ctx.neighbours2[ctx.v - 1].size();
        real contrib_nn2 = 0.0;
        for (size_t j = 0; j < ctx.neighbours2[ctx.v - 1].size(); j++)
            contrib_nn2 += -ctx.fwd_post[ctx.neighbours2[ctx.v - 1][j] - 1].means[idx];
        const int dims = prob->spatial_dims;
        if (type == 'p' || type == 'm')
        {
            nn = 2 * dims;
            nn2 = 4 * dims * dims - nn;
        }
        real spatial_prec = 0;
        if (type == 'M')
            spatial_prec = aK * (nn + 1e-8);
        else if (type == 'm')
            spatial_prec = aK * nn;
        else
            spatial_prec = aK * (nn * nn + nn);
        Mat precs = prior->GetPrecisions();
        if (type == 'p' || type == 'm')
            precs(idx, idx) = spatial_prec;
        else
            precs(idx, idx) = prec + spatial_prec;
        prior->SetPrecisions(precs);
        real spatial_mean;
        if (type == 'm' || type == 'M')
        {
            real rec = 1 / real(nn);
            spatial_mean = contrib_nn * rec;
        }
        else if (nn != 0)
        {
            // priors.cc:455 - INTEGER division quirk: `real rec = 1 / (8*nn - nn2);`
            int den = 8 * nn - nn2;
            real rec;
            if (den == 0)
                rec = INFINITY; // the reference would trap (SIGFPE); not reachable for 3D grids
            else
                rec = (real)(1 / den);
            spatial_mean = (8 * contrib_nn + contrib_nn2) * rec;
        }
        else
            spatial_mean = 0;
        if (type == 'm' || type == 'M')
            prior->means[idx] = prior->GetCovariance()(idx, idx) * spatial_prec * spatial_mean;
        else
            prior->means[idx]
                = prior->GetCovariance()(idx, idx) * (spatial_prec * spatial_mean + prec * mean);
        return 0;
    }
};

std::vector<Prior> make_priors(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf)
{
    std::vector<Prior> priors(prob->model.n_params);
    for (int k = 0; k < prob->model.n_params; k++)
    {
        Prior &p = priors[k];
        p.prob = prob;
        p.buf = buf;
        p.idx = k;
        p.type = prob->params[k].prior_type;
        p.mean = prob->params[k].prior_mean;
        p.prec = prob->params[k].prior_prec;
        p.var = prob->params[k].prior_var;
        p.aK = 1e-8; // priors.cc:185
    }
    return priors;
}

// ---------------------------------------------------------------------------------------------
// Per-voxel set-up (inference_vb.cc:144-248, fwdmodel.cc:284-324)
// ---------------------------------------------------------------------------------------------
struct Engine
{
    const fabber_cuda_vb_problem *prob;
    const fabber_cuda_vb_buffers *buf;
    int N, T, P, NN;
    ModelCtx mc;
    NoiseModel noise;
    bool needF;

    Engine(const fabber_cuda_vb_problem *p, const fabber_cuda_vb_buffers *b)
        : prob(p)
        , buf(b)
    {
        N = p->n_voxels;
        T = p->n_times;
        P = p->model.n_params;
        mc.prob = p;
        mc.T = T;
        mc.P = P;
        noise.init(p);
        NN = noise.ar2 ? FABBER_CUDA_AR2_NOISE_FIELDS(noise.nAlphas)
                       : noise.ar ? FABBER_CUDA_AR_NOISE_FIELDS : 2 * noise.nPhis;
        if (noise.ar2 && (T % 2 != 0 || T < 4))
            throw std::runtime_error("num-echoes=2 needs an even number of time points");
        needF = p->need_f != 0;
    }

    Vec voxel_data(int v) const
    {
        Vec y(T);
        for (int t = 0; t < T; t++)
            y[t] = (real)buf->data[(size_t)t * N + v];
        return y;
    }

    void initial_posterior(int v, const Vec &y, MVN &post) const
    {
        post = MVN(P);
        if (buf->init_mean)
        {
            // continue-from-mvn (inference_vb.cc:209-216): means + covariance given in Fabber space
            Mat cov(P, P);
            int idx = 0;
            for (int r = 0; r < P; r++)
                for (int c = 0; c <= r; c++, idx++)
                    cov(r, c) = cov(c, r) = buf->init_cov[(size_t)idx * N + v];
            for (int i = 0; i < P; i++)
                post.means[i] = buf->init_mean[(size_t)i * N + v];
            post.SetCovariance(cov);
            return;
        }
        Mat cov = post.GetCovariance();
        for (int i = 0; i < P; i++)
        {
            if (prob->params[i].prior_type == 'I')
                post.means[i] = buf->image_prior[i][v];
            else
                post.means[i] = prob->params[i].post_mean;
            cov(i, i) = prob->params[i].post_var;
        }
        post.SetCovariance(cov);
        if (prob->model.id == FABBER_MODEL_EXP)
        {
            // ExpFwdModel::InitVoxelPosterior examples/fwdmodel_exp.cc:84-91
            real mx = y[0];
            for (int t = 1; t < T; t++)
                mx = std::max(mx, y[t]);
            for (int k = 0; k < prob->model.exp_num; k++)
                post.means[2 * k] = mx / (prob->model.exp_num + k);
        }
        // FwdModel::ToFabber fwdmodel.cc:315-324
        cov = post.GetCovariance();
        for (int i = 0; i < P; i++)
        {
            post.means[i] = t_to_fabber(prob->params[i].transform, post.means[i]);
            cov(i, i) = t_to_fabber_var(prob->params[i].transform, cov(i, i));
        }
        post.SetCovariance(cov);
    }

    void initial_noise(int v, NoiseParams &prior, NoiseParams &post) const
    {
        noise.hardcoded_initial(prior, post);
        if (buf->init_noise)
        {
            // InputFromMVN is done by the caller of the ABI; here the raw fields are given
            if (noise.ar2)
            {
                const int nA = noise.nAlphas;
                for (int i = 0; i < 2; i++)
                {
                    post.phis[i].b = buf->init_noise[(size_t)(2 * i) * N + v];
                    post.phis[i].c = buf->init_noise[(size_t)(2 * i + 1) * N + v];
                }
                Mat pr(nA, nA);
                int f = 4;
                for (int i = 0; i < nA; i++)
                    post.alpha.means[i] = buf->init_noise[(size_t)(f++) * N + v];
                for (int r = 0; r < nA; r++)
                    for (int c = 0; c <= r; c++)
                        pr(r, c) = pr(c, r) = buf->init_noise[(size_t)(f++) * N + v];
                post.alpha.SetPrecisions(pr);
            }
            else if (noise.ar)
            {
                post.phis[0].b = buf->init_noise[(size_t)0 * N + v];
                post.phis[0].c = buf->init_noise[(size_t)1 * N + v];
                post.alpha.means[0] = buf->init_noise[(size_t)2 * N + v];
                post.alpha.means[1] = buf->init_noise[(size_t)3 * N + v];
                Mat pr(2, 2);
                pr(0, 0) = buf->init_noise[(size_t)4 * N + v];
                pr(1, 0) = pr(0, 1) = buf->init_noise[(size_t)5 * N + v];
                pr(1, 1) = buf->init_noise[(size_t)6 * N + v];
                post.alpha.SetPrecisions(pr);
            }
            else
                for (int i = 0; i < noise.nPhis; i++)
                {
                    post.phis[i].b = buf->init_noise[(size_t)(2 * i) * N + v];
                    post.phis[i].c = buf->init_noise[(size_t)(2 * i + 1) * N + v];
                }
        }
        noise.precalculate(post, prior);
    }

    void write_result(int v, const MVN &post, const NoiseParams &np, real F, int its, int status) const
    {
        for (int i = 0; i < P; i++)
            buf->mean[(size_t)i * N + v] = post.means[i];
        Mat cov(P, P);
        try
        {
            cov = post.GetCovariance();
        }
        catch (SingularError &)
        {
            cov = Mat(P, P); // dist_mvn.cc:70-77: elements set to 0 on exception
        }
        int idx = 0;
        for (int r = 0; r < P; r++)
            for (int c = 0; c <= r; c++, idx++)
                buf->cov[(size_t)idx * N + v] = cov(r, c);
        if (noise.ar2)
        {
            const int nA = noise.nAlphas;
            for (int i = 0; i < 2; i++)
            {
                buf->noise[(size_t)(2 * i) * N + v] = np.phis[i].b;
                buf->noise[(size_t)(2 * i + 1) * N + v] = np.phis[i].c;
            }
            int f = 4;
            for (int i = 0; i < nA; i++)
                buf->noise[(size_t)(f++) * N + v] = np.alpha.means[i];
            const Mat &pr = np.alpha.GetPrecisions();
            for (int r = 0; r < nA; r++)
                for (int c = 0; c <= r; c++)
                    buf->noise[(size_t)(f++) * N + v] = pr(r, c);
        }
        else if (noise.ar)
        {
            buf->noise[(size_t)0 * N + v] = np.phis[0].b;
            buf->noise[(size_t)1 * N + v] = np.phis[0].c;
            buf->noise[(size_t)2 * N + v] = np.alpha.means[0];
            buf->noise[(size_t)3 * N + v] = np.alpha.means[1];
            const Mat &pr = np.alpha.GetPrecisions();
            buf->noise[(size_t)4 * N + v] = pr(0, 0);
            buf->noise[(size_t)5 * N + v] = pr(1, 0);
            buf->noise[(size_t)6 * N + v] = pr(1, 1);
        }
        else
            for (int i = 0; i < noise.nPhis; i++)
            {
                buf->noise[(size_t)(2 * i) * N + v] = np.phis[i].b;
                buf->noise[(size_t)(2 * i + 1) * N + v] = np.phis[i].c;
            }
        if (buf->free_energy)
            buf->free_energy[v] = F;
        if (buf->iterations)
            buf->iterations[v] = its;
        buf->status[v] = status;
    }
};

// ---------------------------------------------------------------------------------------------
// Neighbours (inference_vb.cc:795-964)
// ---------------------------------------------------------------------------------------------
int binary_search(const std::vector<int> &data, int num)
{
    int first = 1, last = (int)data.size();
    while (first <= last)
    {
        int test = (first + last) / 2;
        if (data[test - 1] < num)
            first = test + 1;
        else if (data[test - 1] > num)
            last = test - 1;
        else
            return test;
    }
    return -1;
}

int calc_neighbours(const int *coords, int nVoxels, int spatial_dims,
    std::vector<std::vector<int> > &neighbours, std::vector<std::vector<int> > &neighbours2)
{
    neighbours.assign(nVoxels, std::vector<int>());
    neighbours2.assign(nVoxels, std::vector<int>());
    if (nVoxels == 0)
        return 0;
    const int *cx = coords, *cy = coords + nVoxels, *cz = coords + 2 * (size_t)nVoxels;
    // CheckCoordMatrixCorrectlyOrdered inference_vb.cc:769-793
    for (int v = 0; v + 1 < nVoxels; v++)
    {
        int sx = (cx[v + 1] > cx[v]) - (cx[v + 1] < cx[v]);
        int sy = (cy[v + 1] > cy[v]) - (cy[v + 1] < cy[v]);
        int sz = (cz[v + 1] > cz[v]) - (cz[v + 1] < cz[v]);
        int d = sx + 10 * sy + 100 * sz;
        if (d <= 0)
            return -1;
    }
    int xsize = *std::max_element(cx, cx + nVoxels) + 1;
    int ysize = *std::max_element(cy, cy + nVoxels) + 1;
    std::vector<int> offsets(nVoxels);
    for (int v = 0; v < nVoxels; v++)
        offsets[v] = cz[v] * xsize * ysize + cy[v] * xsize + cx[v];
    int delta[6] = { 1, -1, xsize, -xsize, xsize * ysize, -xsize * ysize };
    int max_delta = spatial_dims * 2 - 1;
    for (int vid = 1; vid <= nVoxels; vid++)
    {
        int pos = offsets[vid - 1];
        for (int n = 0; n <= max_delta; n++)
        {
            int id = binary_search(offsets, pos + delta[n]);
            if (id < 0)
                continue;
            if (n < 4)
            {
                bool ignore = false;
                if (delta[n] > 0)
                {
                    int test = delta[n + 2];
                    if (test > 0)
                        ignore = (pos % test) >= test - delta[n];
                }
                else
                {
                    int test = -delta[n + 2];
                    if (test > 0)
                        ignore = (pos % test) < -delta[n];
                }
                if (ignore)
                    continue;
            }
            neighbours[vid - 1].push_back(id);
        }
    }
    for (int vid = 1; vid <= nVoxels; vid++)
        for (size_t n1 = 0; n1 < neighbours[vid - 1].size(); n1++)
        {
            int n1id = neighbours[vid - 1][n1];
            int check = 0;
            for (size_t n2 = 0; n2 < neighbours[n1id - 1].size(); n2++)
            {
                int n2id = neighbours[n1id - 1][n2];
                if (n2id != vid)
                    neighbours2[vid - 1].push_back(n2id);
                else
                    check++;
            }
            if (check != 1)
                return -2;
        }
    return 0;
}

// ---------------------------------------------------------------------------------------------
// NLLS (inference_nlls.cc:90-293). The optimiser is MISCMATHS::nonlin - FSL's miscmaths/nonlin.{h,cpp}, an FSL
// dependency that is NOT in the reference tree (CMakeLists.txt links libmiscmaths; FSL 5.0 / 6.0). Its
// Levenberg(-Marquardt) driver `levmar` is restated here from FSL's published source as the builder knows it:
//   cf0 = cf(p); success = true; olambda = 0; lambda = 0.1 (NonlinParam default)
//   while NextIter(success):                 // counts SUCCESSFUL steps only, at most 200
//     if success: g = grad(p), H = hess(p)
//     H_ii <- H_ii (1+lambda)/(1+olambda)    (LM)      or   H_ii + lambda - olambda    (Levenberg, fabber's default)
//     step = -H^-1 g; ncf = cf(p + step)     // a failing solve counts as an unsuccessful step
//     if ncf < cf: p += step; lambda /= 10; olambda = 0; stop if 2|cf-ncf| <= 1e-8 (|cf|+|ncf|+2e-16); cf = ncf
//     else:        olambda = lambda; lambda *= 10; stop if lambda > 1e20
// PARITY OF THIS PART IS PINNED ONLY THROUGH THE REFERENCE'S GOLDEN test/outdata_linear_nlls (means, z-stats and
// finalMVN of the shipped regression case, tests/test_oracle_golden.py): oracle/_ref is built with NO_NLLS
// (inference_nlls.cc:1) because nonlin is not there. Everything around the optimiser (cost function, Jacobian,
// precision = J'J / mse with the 1e-6 floor, exception path) is the reference's own code restated.
// ---------------------------------------------------------------------------------------------
struct NllsCost // NLLSCF, inference_nlls.cc:232-293
{
    const ModelCtx &mc;
    std::vector<int> keep; // unmasked samples (MaskRows)
    Vec data;
    NllsCost(const ModelCtx &m, const Vec &y, const fabber_cuda_vb_problem *prob)
        : mc(m)
        , data(y)
    {
        for (int t = 0; t < mc.T; t++)
            if (!(prob->time_masked && prob->time_masked[t]))
                keep.push_back(t);
    }
    real cf(const Vec &p) const
    {
        Vec pred;
        evaluate_fabber(mc, p, pred);
        real s = 0;
        for (size_t k = 0; k < keep.size(); k++)
        {
            real d = data[keep[k]] - pred[keep[k]];
            s += d * d;
        }
        return s;
    }
    // gradient -2 J'(data - pred) and Gauss-Newton Hessian 2 J'J about p
    void grad_hess(const Vec &p, Vec &g, Mat &H) const
    {
        LinModel lin;
        lin.ReCentre(mc, p);
        Vec pred;
        evaluate_fabber(mc, p, pred);
        g.assign(mc.P, 0.0);
        H = Mat(mc.P, mc.P);
        for (int i = 0; i < mc.P; i++)
        {
            real s = 0;
            for (size_t k = 0; k < keep.size(); k++)
                s += lin.J(keep[k], i) * (data[keep[k]] - pred[keep[k]]);
            g[i] = -2 * s;
            for (int j = 0; j < mc.P; j++)
            {
                real h = 0;
                for (size_t k = 0; k < keep.size(); k++)
                    h += lin.J(keep[k], i) * lin.J(keep[k], j);
                H(i, j) = 2 * h;
            }
        }
    }
};

// MISCMATHS::nonlin with NL_LM (see the header of this section). Returns the number of successful steps.
int nlls_levmar(const NllsCost &cost, Vec &p, bool lm)
{
    const int P = (int)p.size();
    const int maxiter = 200;
    const real cftol = 1e-8, ltol = 1e20;
    real lambda = 0.1, olambda = 0.0;
    real cf = cost.cf(p);
    bool success = true;
    int niter = 0;
    Vec g;
    Mat H;
    for (;;)
    {
        if (success && niter++ >= maxiter)
            break;
        if (success)
            cost.grad_hess(p, g, H);
        for (int i = 0; i < P; i++)
        {
            if (lm)
                H(i, i) = ((1.0 + lambda) / (1.0 + olambda)) * H(i, i);
            else
                H(i, i) = H(i, i) + lambda - olambda;
        }
        Vec trial(P);
        real ncf = 0;
        bool inv_fail = false;
        try
        {
            Vec step = mulv(inverse(H), g);
            for (int i = 0; i < P; i++)
                trial[i] = p[i] + -step[i];
            ncf = cost.cf(trial);
        }
        catch (SingularError &)
        {
            inv_fail = true;
        }
        if (!inv_fail && (success = (ncf < cf)))
        {
            olambda = 0.0;
            p = trial;
            lambda = lambda / 10.0;
            const bool conv = 2.0 * std::fabs(cf - ncf) <= cftol * (std::fabs(cf) + std::fabs(ncf) + 2.0e-16);
            cf = ncf;
            if (conv)
                break;
        }
        else
        {
            success = false;
            olambda = lambda;
            lambda = 10.0 * lambda;
            if (lambda > ltol)
                break;
        }
    }
    return niter > maxiter ? maxiter : niter;
}

int nlls_run(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf)
{
    Engine eng(prob, buf);
    const int N = eng.N, P = eng.P, T = eng.T;
    int n_masked = 0;
    if (prob->time_masked)
        for (int t = 0; t < T; t++)
            n_masked += prob->time_masked[t] ? 1 : 0;
    const int Nsamples = T - n_masked;
    // initialFwdPosterior: the model's hard-coded posterior (Fabber space) or the fwd-inital-posterior file;
    // NOT the per-voxel InitVoxelPosterior (inference_nlls.cc:66-83,141-143)
    Vec start(P);
    for (int i = 0; i < P; i++)
        start[i] = prob->nlls_have_start ? (real)prob->nlls_start[i]
                                         : t_to_fabber(prob->params[i].transform, prob->params[i].post_mean);
    for (int v = 0; v < N; v++)
    {
        Vec y = eng.voxel_data(v);
        NllsCost cost(eng.mc, y, prob);
        Vec p = start;
        int status = 0, its = 0;
        Mat cov(P, P);
        try
        {
            its = nlls_levmar(cost, p, prob->nlls_lm != 0);
            LinModel lin;
            lin.ReCentre(eng.mc, p);
            const real sqerr = cost.cf(p);
            const real mse = sqerr / (Nsamples - P);
            Mat prec(P, P);
            // QUIRK KEPT: `MaskRows(J, m_masked_tpoints);` at inference_nlls.cc:172 takes J by value and its result is
            // discarded, so the precision is built from the Jacobian of ALL samples - masked ones included - while
            // sqerr and the degrees of freedom leave them out (found by running the reference's own code,
            // tests/test_reference_build.py::test_nlls_poly_masked_timepoints)
            for (int i = 0; i < P; i++)
                for (int j = 0; j < P; j++)
                {
                    real h = 0;
                    for (int t = 0; t < T; t++)
                        h += lin.J(t, i) * lin.J(t, j);
                    prec(i, j) = h / mse;
                }
            for (int i = 0; i < P; i++)
                if (prec(i, i) < 1e-6)
                    prec(i, i) = 1e-6;
            if (!all_finite(prec))
            {
                // mse = 0/0 (as many samples as parameters, test/test_inference.cc:79-105) or a perfect fit: NEWMAT's
                // inverse does not throw on NaN, the covariance is NaN and the voxel is NOT an error
                for (int i = 0; i < P; i++)
                    for (int j = 0; j < P; j++)
                        cov(i, j) = std::numeric_limits<real>::quiet_NaN();
            }
            else
            {
                MVN post(P);
                post.SetPrecisions(prec);
                cov = post.GetCovariance();
            }
        }
        catch (SingularError &)
        {
            // "precision matrix is probably singular so set manually" (:212-221): means kept, precisions 1e-12 I
            status = FABBER_VOX_SINGULAR;
            cov = Mat(P, P);
            for (int i = 0; i < P; i++)
                cov(i, i) = 1 / (real)1e-12;
        }
        catch (InternalError &e)
        {
            status = e.code; // a non-finite offset / Jacobian ends the reference's run whatever allow-bad-voxels says
            cov = Mat(P, P);
        }
        for (int i = 0; i < P; i++)
            buf->mean[(size_t)i * N + v] = p[i];
        int idx = 0;
        for (int r = 0; r < P; r++)
            for (int c = 0; c <= r; c++, idx++)
                buf->cov[(size_t)idx * N + v] = cov(r, c);
        if (buf->iterations)
            buf->iterations[v] = its;
        buf->status[v] = status;
        if (status != 0 && !prob->allow_bad_voxels)
            return FABBER_CUDA_ERR_BAD_VOXEL;
    }
    return FABBER_CUDA_OK;
}

} // namespace

// ---------------------------------------------------------------------------------------------
// Entry points (C linkage, host pointers in the buffers struct)
// ---------------------------------------------------------------------------------------------
extern "C" {

/* Vb::DoCalculationsVoxelwise, inference_vb.cc:415-576 */
int vb_oracle_voxelwise(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf)
{
    if (prob->method == FABBER_METHOD_NLLS)
        return nlls_run(prob, buf);
    Engine eng(prob, buf);
    const int N = eng.N, P = eng.P;
    std::vector<Prior> priors = make_priors(prob, buf);
    RunContext ctx;
    ctx.nvoxels = N;
    ctx.fwd_post.resize(1);
    ctx.fwd_prior.resize(1);
    int n_bad = 0;

    for (int v = 0; v < N; v++)
    {
        Vec y = eng.voxel_data(v);
        MVN &post = ctx.fwd_post[0];
        MVN &prior = ctx.fwd_prior[0];
        prior = MVN(P); // inference_vb.cc:159
        NoiseParams nprior, npost;
        LinModel lin;
        Conv conv;
        int status = 0;
        real F = 1234.5678;
        real Fprior = 0;
        ctx.v = 1; // priors index ctx.fwd_post[ctx.v-1]; single-slot context here
        ctx.it = 0;
        // SetupPerVoxelDists (inference_vb.cc:207-247) is outside the reference's try block:
        // a failure there aborts the run regardless of allow-bad-voxels.
        try
        {
            eng.initial_posterior(v, y, post);
            eng.initial_noise(v, nprior, npost);
            lin.ReCentre(eng.mc, post.means); // inference_vb.cc:235
        }
        catch (InternalError &e)
        {
            eng.write_result(v, post, npost, F, 0, e.code | FABBER_VOX_SETUP_FLAG);
            return FABBER_CUDA_ERR_BAD_VOXEL;
        }
        try
        {
            conv.Initialize(prob->conv_type, prob->max_iterations, prob->fchange, prob->max_trials);

            NoiseParams noiseSave = npost;
            MVN postSave = post, priorSave = prior;

            lin.ReCentre(eng.mc, post.means); // :443
            conv.Reset();
            do
            {
                if (conv.NeedSave())
                {
                    noiseSave = npost;
                    postSave = post;
                    priorSave = prior;
                }
                for (int k = 0; k < P; k++)
                {
                    // image priors index by the true voxel: temporarily map ctx.v
                    if (priors[k].type == 'I')
                    {
                        prior.means[k] = buf->image_prior[k][v];
                        Mat pm = prior.GetPrecisions();
                        pm(k, k) = priors[k].prec;
                        prior.SetPrecisions(pm);
                        Fprior = 0;
                    }
                    else
                        Fprior = priors[k].ApplyToMVN(&prior, ctx); // '=' not '+=' (:462)
                }
                if (eng.needF)
                    F = eng.noise.free_energy(npost, nprior, post, prior, lin, y) + Fprior;
                eng.noise.update_theta(npost, post, prior, lin, y, conv.LMalpha());
                if (eng.needF)
                    F = eng.noise.free_energy(npost, nprior, post, prior, lin, y) + Fprior;
                eng.noise.update_noise(npost, nprior, post, lin, y);
                if (eng.needF)
                    F = eng.noise.free_energy(npost, nprior, post, prior, lin, y) + Fprior;
                lin.ReCentre(eng.mc, post.means);
                if (eng.needF)
                    F = eng.noise.free_energy(npost, nprior, post, prior, lin, y) + Fprior;
                if (buf->f_history && ctx.it < prob->f_history_len)
                    buf->f_history[(size_t)ctx.it * N + v] = F;
                ++ctx.it;
            } while (!conv.Test(F));

            if (conv.NeedSave())
            {
                noiseSave = npost;
                postSave = post;
                priorSave = prior;
            }
            if (conv.NeedRevert())
            {
                npost = noiseSave;
                post = postSave;
                prior = priorSave;
                lin.ReCentre(eng.mc, post.means);
                if (eng.needF)
                    F = eng.noise.free_energy(npost, nprior, post, prior, lin, y) + Fprior;
            }
        }
        catch (InternalError &e)
        {
            status = e.code;
        }
        catch (SingularError &)
        {
            status = FABBER_VOX_SINGULAR;
        }
        if (buf->f_history)
            for (int h = ctx.it; h < prob->f_history_len; h++)
                buf->f_history[(size_t)h * N + v] = F; // padded with the last value (:1041-1044)
        eng.write_result(v, post, npost, F, ctx.it, status);
        if (status != 0)
        {
            n_bad++;
            if (!prob->allow_bad_voxels)
                return FABBER_CUDA_ERR_BAD_VOXEL;
        }
    }
    return FABBER_CUDA_OK;
}

/* Vb::DoCalculationsSpatial, inference_vb.cc:578-767 */
int vb_oracle_spatial(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf)
{
    Engine eng(prob, buf);
    const int N = eng.N, P = eng.P;
    RunContext ctx;
    ctx.nvoxels = N;
    ctx.it = 0;
    ctx.fwd_post.resize(N);
    ctx.fwd_prior.assign(N, MVN(P));
    ctx.noise_post.resize(N);
    ctx.noise_prior.resize(N);
    std::vector<LinModel> lin(N);
    std::vector<Vec> ys(N);
    std::vector<int> status(N, 0);
    std::vector<double> resultF(N, 9999.0);

    // SetupPerVoxelDists
    for (int v = 0; v < N; v++)
    {
        ys[v] = eng.voxel_data(v);
        try
        {
            eng.initial_posterior(v, ys[v], ctx.fwd_post[v]);
            eng.initial_noise(v, ctx.noise_prior[v], ctx.noise_post[v]);
            if (buf->lock_centre) // inference_vb.cc:227-231
            {
                Vec centre(P);
                for (int i = 0; i < P; i++)
                    centre[i] = buf->lock_centre[(size_t)i * N + v];
                lin[v].ReCentre(eng.mc, centre);
            }
            else
                lin[v].ReCentre(eng.mc, ctx.fwd_post[v].means);
        }
        catch (InternalError &e)
        {
            return FABBER_CUDA_ERR_BAD_VOXEL; // set-up errors are not caught by the reference
        }
    }
    int rc = calc_neighbours(buf->coords, N, prob->spatial_dims, ctx.neighbours, ctx.neighbours2);
    if (rc != 0)
        return FABBER_CUDA_ERR_INVALID;
    std::vector<Prior> priors = make_priors(prob, buf);

    Conv conv;
    conv.Initialize(FABBER_CONV_MAXITS, prob->max_iterations, 0.01, 10);

    struct Local
    {
        static void ignore_voxel(RunContext &ctx, int v)
        {
            ctx.ignore_voxels.push_back(v);
            std::vector<int> nn = ctx.neighbours[v - 1];
            for (size_t i = 0; i < nn.size(); i++)
            {
                std::vector<int> &n2 = ctx.neighbours[nn[i] - 1];
                n2.erase(std::remove(n2.begin(), n2.end(), v), n2.end());
            }
            nn = ctx.neighbours2[v - 1];
            for (size_t i = 0; i < nn.size(); i++)
            {
                std::vector<int> &n2 = ctx.neighbours2[nn[i] - 1];
                n2.erase(std::remove(n2.begin(), n2.end(), v), n2.end());
            }
        }
        static bool ignored(const RunContext &ctx, int v)
        {
            return std::find(ctx.ignore_voxels.begin(), ctx.ignore_voxels.end(), v) != ctx.ignore_voxels.end();
        }
    };

    real Fglobal = 1234.5678;
    do
    {
        real Fprior = 0;
        for (int v = 1; v <= N; v++)
        {
            ctx.v = v;
            try
            {
                Fprior = 0;
                for (int k = 0; k < P; k++)
                {
                    if (buf->spatial_ak && v == 1 && priors[k].is_spatial())
                    {
                        // record after the (possible) update below
                    }
                    Fprior += priors[k].ApplyToMVN(&ctx.fwd_prior[v - 1], ctx);
                }
                if (v == 1 && buf->spatial_ak)
                    for (int k = 0; k < P; k++)
                        buf->spatial_ak[(size_t)ctx.it * P + k] = priors[k].is_spatial() ? priors[k].aK : 0.0;
                if (Local::ignored(ctx, v))
                    continue;
                if (eng.needF)
                    resultF[v - 1] = eng.noise.free_energy(ctx.noise_post[v - 1], ctx.noise_prior[v - 1],
                                         ctx.fwd_post[v - 1], ctx.fwd_prior[v - 1], lin[v - 1], ys[v - 1])
                        + Fprior;
                eng.noise.update_theta(ctx.noise_post[v - 1], ctx.fwd_post[v - 1], ctx.fwd_prior[v - 1],
                    lin[v - 1], ys[v - 1], 0);
                if (eng.needF)
                    resultF[v - 1] = eng.noise.free_energy(ctx.noise_post[v - 1], ctx.noise_prior[v - 1],
                                         ctx.fwd_post[v - 1], ctx.fwd_prior[v - 1], lin[v - 1], ys[v - 1])
                        + Fprior;
            }
            catch (InternalError &e)
            {
                if (!prob->allow_bad_voxels)
                    return FABBER_CUDA_ERR_BAD_VOXEL;
                status[v - 1] = e.code;
                Local::ignore_voxel(ctx, v);
            }
            catch (SingularError &)
            {
                if (!prob->allow_bad_voxels)
                    return FABBER_CUDA_ERR_BAD_VOXEL;
                status[v - 1] = FABBER_VOX_SINGULAR;
                Local::ignore_voxel(ctx, v);
            }
        }
        Fglobal = 0;
        for (int v = 1; v <= N; v++)
        {
            try
            {
                if (Local::ignored(ctx, v))
                    continue;
                eng.noise.update_noise(ctx.noise_post[v - 1], ctx.noise_prior[v - 1], ctx.fwd_post[v - 1],
                    lin[v - 1], ys[v - 1]);
                if (eng.needF)
                    resultF[v - 1] = eng.noise.free_energy(ctx.noise_post[v - 1], ctx.noise_prior[v - 1],
                                         ctx.fwd_post[v - 1], ctx.fwd_prior[v - 1], lin[v - 1], ys[v - 1])
                        + Fprior;
                if (!buf->lock_centre) // :695
                    lin[v - 1].ReCentre(eng.mc, ctx.fwd_post[v - 1].means);
                real F = 1234.5678;
                if (eng.needF)
                {
                    F = eng.noise.free_energy(ctx.noise_post[v - 1], ctx.noise_prior[v - 1],
                            ctx.fwd_post[v - 1], ctx.fwd_prior[v - 1], lin[v - 1], ys[v - 1])
                        + Fprior; // stale Fprior of the last voxel of the first loop (:700)
                    resultF[v - 1] = F;
                }
                Fglobal += F;
            }
            catch (InternalError &e)
            {
                if (!prob->allow_bad_voxels)
                    return FABBER_CUDA_ERR_BAD_VOXEL;
                status[v - 1] = e.code;
                Local::ignore_voxel(ctx, v);
            }
            catch (SingularError &)
            {
                if (!prob->allow_bad_voxels)
                    return FABBER_CUDA_ERR_BAD_VOXEL;
                status[v - 1] = FABBER_VOX_SINGULAR;
                Local::ignore_voxel(ctx, v);
            }
        }
        ++ctx.it;
    } while (!conv.Test(Fglobal));

    if (buf->spatial_ak)
        for (int k = 0; k < P; k++)
            buf->spatial_ak[(size_t)ctx.it * P + k] = priors[k].is_spatial() ? priors[k].aK : 0.0;
    for (int v = 0; v < N; v++)
        eng.write_result(v, ctx.fwd_post[v], ctx.noise_post[v], resultF[v], ctx.it, status[v]);
    return FABBER_CUDA_OK;
}

/* neighbours as flat CSR (1-based ids as in the reference); returns 0 ok */
int vb_oracle_neighbours(const int *coords, int n_voxels, int spatial_dims, int *nn_offsets /*[N+1]*/,
    int *nn_ids, int nn_cap, int *nn2_offsets, int *nn2_ids, int nn2_cap)
{
    std::vector<std::vector<int> > n1, n2;
    int rc = calc_neighbours(coords, n_voxels, spatial_dims, n1, n2);
    if (rc != 0)
        return rc;
    int o = 0, o2 = 0;
    for (int v = 0; v < n_voxels; v++)
    {
        nn_offsets[v] = o;
        for (size_t j = 0; j < n1[v].size(); j++)
        {
            if (o >= nn_cap)
                return -3;
            nn_ids[o++] = n1[v][j];
        }
        nn2_offsets[v] = o2;
        for (size_t j = 0; j < n2[v].size(); j++)
        {
            if (o2 >= nn2_cap)
                return -3;
            nn2_ids[o2++] = n2[v][j];
        }
    }
    nn_offsets[n_voxels] = o;
    nn2_offsets[n_voxels] = o2;
    return 0;
}

/* fit[t][v] = model(ToModel(mean[:,v]))  (inference.cc:190-191) */
int vb_oracle_model_fit(const fabber_cuda_vb_problem *prob, const double *mean, double *fit)
{
    ModelCtx mc;
    mc.prob = prob;
    mc.T = prob->n_times;
    mc.P = prob->model.n_params;
    const int N = prob->n_voxels;
    Vec th(mc.P), out;
    for (int v = 0; v < N; v++)
    {
        for (int i = 0; i < mc.P; i++)
            th[i] = mean[(size_t)i * N + v];
        evaluate_fabber(mc, th, out);
        for (int t = 0; t < mc.T; t++)
            fit[(size_t)t * N + v] = out[t];
    }
    return 0;
}

/* Convergence detector driver for unit tests mirroring test/test_convergence.cc:
 * feeds F[0..n) to Test() and records the return value plus NeedSave/NeedRevert/LMalpha after each. */
int vb_oracle_convergence_trace(int conv_type, int max_its, double fchange, int max_trials, const double *F,
    int n, int *out_test, int *out_save, int *out_revert, float *out_alpha)
{
    Conv c;
    c.Initialize(conv_type, max_its, fchange, max_trials);
    for (int i = 0; i < n; i++)
    {
        out_test[i] = c.Test(F[i]) ? 1 : 0;
        out_save[i] = c.NeedSave() ? 1 : 0;
        out_revert[i] = c.NeedRevert() ? 1 : 0;
        out_alpha[i] = c.LMalpha();
    }
    return 0;
}

double vb_oracle_gammaln(double x)
{
    return gammaln(x);
}
double vb_oracle_digamma(double x)
{
    return digamma_fsl(x);
}

} // extern "C"
