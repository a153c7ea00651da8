/*
 * vb_voxelwise_ar2.cuh - non-spatial VB with AR(1) noise on TWO interleaved echoes (Ar1cNoiseModel with
 * num-echoes=2, ar1-cross-terms = none / same / dual): all iterations of a voxel in one thread.
 *
 * Reference: noisemodel_ar.cc - alpha matrices and marginals :83-223, UpdateAlpha :447-528, UpdatePhi :530-556,
 * UpdateTheta :558-634, CalcFreeEnergy :643-747, Precalculate :749-769, HardcodedInitialDists :379-403.
 *
 * What the reference holds. The series interleaves the echoes, TE1 TE2 TE1 TE2 .. (:126-129), nT = T/2 samples
 * each. Per voxel it builds twelve dense T x T "alpha matrices" M(n, a, c), n = echo, each ONE diagonal line of
 * nT-1 entries of +-1 (reflected to stay symmetric), and two marginals
 *      Q_n = M(n,0,0) + E[a_n] M(n,1,0) + E[a_n^2] M(n,2,0)
 *                     + E[x_n] M(n,0,1) + E[a_n x_n] M(n,1,1) + E[x_n^2] M(n,0,2)
 * where a_n is the echo's own AR coefficient and x_n its cross-term coefficient (alpha 3 for both echoes with
 * "same", alpha 3 / alpha 4 with "dual", absent with "none"). This is the expectation of
 *      sum_{t>=2} ( k_n,t - a_n k_n,t-1 - x_n k_m,t )^2        m = the other echo
 * so with echo-1 sample E1_t and echo-2 sample E2_t the nine lines are
 *      a  E1 E1, t >= 2     b  E1 E1, t <= nT-1     c  E1_t E1_t-1
 *      d  E2 E2, t >= 2     e  E2 E2, t <= nT-1     f  E2_t E2_t-1
 *      g  E2_t E1_t (t>=2)  h  E2_t E1_t-1          i  E1_t E2_t-1
 *      M(1,00)=a  M(1,10)=-c  M(1,20)=b  M(1,01)=-g  M(1,11)=+h  M(1,02)=d
 *      M(2,00)=d  M(2,10)=-f  M(2,20)=e  M(2,01)=-g  M(2,11)=+i  M(2,02)=a        (off-diagonal lines count twice)
 *
 * Re-design. Nothing T x T exists here. An iteration is two passes over the voxel's time series:
 *   pass 1 (before UpdateTheta): the noise posterior is fixed, so X = phi1 Q_1 + phi2 Q_2 is a banded matrix with
 *          SEVEN known line weights; one pass accumulates (J'XJ, J'Xr, r'Xr) directly - the statistics the free
 *          energy and the theta update need (k'Xk = rr + 2 b.d + d'Ad as in the white kernel).
 *   pass 2 (after UpdateTheta): OperatorKLJ(M) = k'Mk + tr(Sigma J'MJ) is LINEAR in M, and with the new theta
 *          (d = centre - mean, Sigma) known it is a scalar per line: nine running sums of
 *          q(u, v) = k_u k_v + J_u' Sigma J_v. UpdateAlpha and UpdatePhi are then scalar algebra on those nine.
 * J and the residual are re-evaluated in pass 2 (deterministic: the same centre gives the same values): two model
 * passes per iteration against the reference's ~14 T x T products. O(T) work, no per-voxel matrices.
 */
#pragma once
#include "vb_voxelwise_ar.cuh"

namespace fab
{
constexpr int AR2_NTA = 10; /* packed 4 x 4; the leading n(n+1)/2 entries are the packed n x n matrix */

/* n x n inverse (n = 2, 3, 4 at run time) on the packed 4 x 4 storage; MVN = with the 1e-10 retry */
template <bool MVN> FAB_DEV bool ar2_inverse(int nA, const double (&A)[AR2_NTA], double (&Inv)[AR2_NTA], double &logdet,
    bool want_logdet)
{
    bool ok;
#pragma unroll
    for (int i = 0; i < AR2_NTA; i++)
        Inv[i] = 0.0;
    if (nA == 2)
    {
        double a[3] = { A[0], A[1], A[2] }, inv[3];
        ok = MVN ? mvn_inverse<2>(a, inv, logdet, want_logdet) : ldl_inverse<2>(a, inv, logdet, want_logdet);
#pragma unroll
        for (int i = 0; i < 3; i++)
            Inv[i] = inv[i];
    }
    else if (nA == 3)
    {
        double a[6], inv[6];
#pragma unroll
        for (int i = 0; i < 6; i++)
            a[i] = A[i];
        ok = MVN ? mvn_inverse<3>(a, inv, logdet, want_logdet) : ldl_inverse<3>(a, inv, logdet, want_logdet);
#pragma unroll
        for (int i = 0; i < 6; i++)
            Inv[i] = inv[i];
    }
    else
        ok = MVN ? mvn_inverse<4>(A, Inv, logdet, want_logdet) : ldl_inverse<4>(A, Inv, logdet, want_logdet);
    return ok;
}

/* the seven line weights of X = phi1 Q_1 + phi2 Q_2 (element values; off-diagonal lines appear in both triangles) */
struct Ar2Lines
{
    double d1a, d1b; /* E1 diagonal: t >= 2, t <= nT-1 */
    double d2a, d2b; /* E2 diagonal */
    double o11, o22; /* E1_t E1_t-1, E2_t E2_t-1 */
    double o12, oh, oi; /* E2_t E1_t, E2_t E1_t-1, E1_t E2_t-1 (zero without cross terms) */
};

/* the nine OperatorKLJ line sums (raw sums of q(u, v): no line weight, no factor 2) */
struct Ar2Klj
{
    double a, b, c, d, e, f, g, h, i;
};

/* expectations of the alphas the marginals need (Ar1cMatrixCache::Update :197-222) */
struct Ar2Moments
{
    double a1, a2, x1, x2;         /* E[a_1], E[a_2], E[x_1], E[x_2] */
    double a1a1, a2a2, a1x1, a2x2; /* covarPlus entries */
    double x1x1, x2x2;
};

/* one (J, r) sample of the linearised model at the pass's centre */
template <class Model, int FAST, bool BASIS>
FAB_DEV void ar2_sample(const typename Model::Ctx &mc, int s, double y, const double (&p0)[Model::P],
    const double (&pp)[Model::P], const double (&pn)[Model::P], const double (&rden)[Model::P],
    const double (&jscale)[Model::P], double (&J)[Model::P], double &r)
{
    constexpr int P = Model::P;
    typename Model::Sample smp;
    Model::sample(mc, s, smp);
    double g;
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
    r = y - g;
}

/* walk the series pair by pair: body(t, J1, r1, J2, r2) */
template <class Model, int FAST, bool BASIS, class Body>
FAB_DEV void ar2_pairs(const VbArgs &a, const typename Model::Ctx &mc, int v, const double (&p0)[Model::P],
    const double (&pp)[Model::P], const double (&pn)[Model::P], const double (&rden)[Model::P], Body &body)
{
    constexpr int P = Model::P;
    const int nT = a.T >> 1;
    const float *yp = a.data + v;
    const size_t stride = (size_t)a.N;
    double jscale[P];
#pragma unroll
    for (int i = 0; i < P; i++)
        jscale[i] = (pp[i] - pn[i]) * rden[i];
    /* two pairs of software prefetch */
    float q0 = __ldg(yp), q1 = __ldg(yp + stride), q2 = 0.f, q3 = 0.f;
    if (1 < nT)
    {
        q2 = __ldg(yp + 2 * stride);
        q3 = __ldg(yp + 3 * stride);
    }
#pragma unroll 1
    for (int t = 0; t < nT; t++)
    {
        const double y1 = (double)q0, y2 = (double)q1;
        q0 = q2;
        q1 = q3;
        if (t + 2 < nT)
        {
            q2 = __ldg(yp + (size_t)(2 * t + 4) * stride);
            q3 = __ldg(yp + (size_t)(2 * t + 5) * stride);
        }
        double J1[P], r1, J2[P], r2;
        ar2_sample<Model, FAST, BASIS>(mc, 2 * t, y1, p0, pp, pn, rden, jscale, J1, r1);
        ar2_sample<Model, FAST, BASIS>(mc, 2 * t + 1, y2, p0, pp, pn, rden, jscale, J2, r2);
        body(t, J1, r1, J2, r2);
    }
}

template <class Model, class Body>
FAB_DEV void ar2_walk(const VbArgs &a, const typename Model::Ctx &mc, int v, const double (&p0)[Model::P],
    const double (&pp)[Model::P], const double (&pn)[Model::P], const double (&rden)[Model::P], bool fast, Body &body)
{
    bool basis = false; /* opt-in, see recentre_loop (vb_voxelwise.cuh) */
    if constexpr (Model::LINEAR)
        basis = a.basis_jacobian != 0;
    if (basis)
    {
        if constexpr (Model::LINEAR)
            ar2_pairs<Model, 0, true>(a, mc, v, p0, pp, pn, rden, body);
    }
    else if (fast)
        ar2_pairs<Model, 1, false>(a, mc, v, p0, pp, pn, rden, body);
    else
        ar2_pairs<Model, 0, false>(a, mc, v, p0, pp, pn, rden, body);
}

/* finite-difference points of LinearizedFwdModel::ReCentre about c (fwdmodel_linear.cc:142-172) */
template <class Model>
FAB_DEV void ar2_centre(const VbArgs &a, const double (&c)[Model::P], double (&p0)[Model::P], double (&pp)[Model::P],
    double (&pn)[Model::P], double (&rden)[Model::P])
{
#pragma unroll
    for (int i = 0; i < Model::P; i++)
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
}

/* pass 1: (J'XJ, J'Xr, r'Xr) for the banded X with line weights L. Returns 0 or FABBER_VOX_NONFINITE_*. */
template <class Model>
FAB_DEV int ar2_pass_theta(const VbArgs &a, const typename Model::Ctx &mc, int v, const double (&c)[Model::P],
    const Ar2Lines &L, bool cross, Stats<Model::P> &Q)
{
    constexpr int P = Model::P;
    double p0[P], pp[P], pn[P], rden[P];
    ar2_centre<Model>(a, c, p0, pp, pn, rden);
    const bool fast = Model::HAS_FAST && Model::fast_ok(mc, a.T, p0, pp, pn);
    const int nT = a.T >> 1;
    Q.zero();
    double J1p[P], J2p[P], r1p = 0.0, r2p = 0.0, chk = 0.0;
#pragma unroll
    for (int i = 0; i < P; i++)
        J1p[i] = J2p[i] = 0.0;
    /* x'Xx = sum_s x_s (D_s x_s + 2 sum_{earlier u} O_su x_u): each sample meets u = D/2 x + (its earlier
     * neighbours on the lines), and the symmetric products J u' + u J' carry both triangles. The previous pair is
     * all zeros at t = 0, so the off-diagonal lines start at t = 1 without a test. */
    auto body = [&](int t, const double (&J1)[P], double r1, const double (&J2)[P], double r2) {
        const double D1 = 0.5 * ((t >= 1 ? L.d1a : 0.0) + (t < nT - 1 ? L.d1b : 0.0));
        const double D2 = 0.5 * ((t >= 1 ? L.d2a : 0.0) + (t < nT - 1 ? L.d2b : 0.0));
        const double o12 = t >= 1 ? L.o12 : 0.0;
        double u1[P], u2[P];
        double ur1 = fma(D1, r1, L.o11 * r1p), ur2 = fma(D2, r2, L.o22 * r2p);
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            u1[i] = fma(D1, J1[i], L.o11 * J1p[i]);
            u2[i] = fma(D2, J2[i], L.o22 * J2p[i]);
        }
        if (cross)
        {
            ur1 = fma(L.oi, r2p, ur1);
            ur2 = fma(o12, r1, fma(L.oh, r1p, ur2));
#pragma unroll
            for (int i = 0; i < P; i++)
            {
                u1[i] = fma(L.oi, J2p[i], u1[i]);
                u2[i] = fma(o12, J1[i], fma(L.oh, J1p[i], u2[i]));
            }
        }
        Q.rr = fma(2.0 * r1, ur1, fma(2.0 * r2, ur2, Q.rr));
        chk = fma(r1, r1, fma(r2, r2, chk));
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            Q.b[i] = fma(J1[i], ur1, fma(u1[i], r1, fma(J2[i], ur2, fma(u2[i], r2, Q.b[i]))));
            chk = fma(J1[i], J1[i], fma(J2[i], J2[i], chk));
#pragma unroll
            for (int j = 0; j <= i; j++)
                Q.A[tri(i, j)] = fma(J1[i], u1[j], fma(u1[i], J1[j], fma(J2[i], u2[j], fma(u2[i], J2[j], Q.A[tri(i, j)]))));
        }
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            J1p[i] = J1[i];
            J2p[i] = J2[i];
        }
        r1p = r1;
        r2p = r2;
    };
    ar2_walk<Model>(a, mc, v, p0, pp, pn, rden, fast, body);
    bool bad_g = false, bad_j = false;
    if (!finite_d(chk)) /* the rare path: which of offset / Jacobian it was (fwdmodel_linear.cc:134,174) */
    {
        if (fast)
            recentre_diagnose<Model, true>(a, mc, p0, pp, pn, rden, bad_g, bad_j);
        else
            recentre_diagnose<Model, false>(a, mc, p0, pp, pn, rden, bad_g, bad_j);
    }
    return bad_g ? FABBER_VOX_NONFINITE_OFFSET : (bad_j ? FABBER_VOX_NONFINITE_JACOBIAN : 0);
}

/* pass 2: the nine OperatorKLJ line sums for k = r + J d and Sigma (noisemodel_ar.cc:433-445) */
template <class Model>
FAB_DEV void ar2_pass_noise(const VbArgs &a, const typename Model::Ctx &mc, int v, const double (&c)[Model::P],
    const double (&d)[Model::P], const double (&Sig)[NTri<Model::P>::value], bool cross, Ar2Klj &K)
{
    constexpr int P = Model::P;
    double p0[P], pp[P], pn[P], rden[P];
    ar2_centre<Model>(a, c, p0, pp, pn, rden);
    const bool fast = Model::HAS_FAST && Model::fast_ok(mc, a.T, p0, pp, pn);
    const int nT = a.T >> 1;
    K.a = K.b = K.c = K.d = K.e = K.f = K.g = K.h = K.i = 0.0;
    double z1p[P], z2p[P], k1p = 0.0, k2p = 0.0;
#pragma unroll
    for (int i = 0; i < P; i++)
        z1p[i] = z2p[i] = 0.0;
    auto body = [&](int t, const double (&J1)[P], double r1, const double (&J2)[P], double r2) {
        double z1[P], z2[P], k1 = r1, k2 = r2;
        symv<P>(Sig, J1, z1);
        symv<P>(Sig, J2, z2);
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            k1 = fma(J1[i], d[i], k1);
            k2 = fma(J2[i], d[i], k2);
        }
        double q11 = k1 * k1, q22 = k2 * k2, q1p = k1 * k1p, q2p = k2 * k2p;
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            q11 = fma(J1[i], z1[i], q11);
            q22 = fma(J2[i], z2[i], q22);
            q1p = fma(J1[i], z1p[i], q1p);
            q2p = fma(J2[i], z2p[i], q2p);
        }
        if (t >= 1)
        {
            K.a += q11;
            K.d += q22;
        }
        if (t < nT - 1)
        {
            K.b += q11;
            K.e += q22;
        }
        K.c += q1p; /* the previous pair is all zeros at t = 0 */
        K.f += q2p;
        if (cross)
        {
            double q21 = k2 * k1, q21p = k2 * k1p, q12p = k1 * k2p;
#pragma unroll
            for (int i = 0; i < P; i++)
            {
                q21 = fma(J2[i], z1[i], q21);
                q21p = fma(J2[i], z1p[i], q21p);
                q12p = fma(J1[i], z2p[i], q12p);
            }
            if (t >= 1)
                K.g += q21;
            K.h += q21p;
            K.i += q12p;
        }
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            z1p[i] = z1[i];
            z2p[i] = z2[i];
        }
        k1p = k1;
        k2p = k2;
    };
    ar2_walk<Model>(a, mc, v, p0, pp, pn, rden, fast, body);
}

template <class Model> struct Ar2Voxel
{
    static constexpr int P = Model::P;
    static constexpr int NT = NTri<P>::value;
    double m[P], Lam[NT], Sig[NT], m0[P], L0[P];
    double logdetLam;
    /* noise posterior: phi_n ~ Gamma(nb, nc), alpha ~ N(am, aprec^-1) of size nA */
    double nb[2], nc[2], am[4], aprec[AR2_NTA];

    static constexpr int STASH_DOUBLES = 3 * P + 2 * NT + 1 + 4 + 4 + AR2_NTA;
    static constexpr int SNAP_DOUBLES = 3 * P + NT + 4 + 4 + AR2_NTA;

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
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            s[(k++) * VB_BLOCK] = nb[i];
            s[(k++) * VB_BLOCK] = nc[i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            s[(k++) * VB_BLOCK] = am[i];
#pragma unroll
        for (int i = 0; i < AR2_NTA; i++)
            s[(k++) * VB_BLOCK] = aprec[i];
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
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            nb[i] = s[(k++) * VB_BLOCK];
            nc[i] = s[(k++) * VB_BLOCK];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            am[i] = s[(k++) * VB_BLOCK];
#pragma unroll
        for (int i = 0; i < AR2_NTA; i++)
            aprec[i] = s[(k++) * VB_BLOCK];
    }

    /* Ar1cMatrixCache::Update :197-222: covarPlus = Cov(alpha) + alpha alpha'. The cross-term alpha is number 3
     * for both echoes with "same" (nA = 3), 3 and 4 with "dual" (nA = 4). False if the precisions are singular. */
    FAB_DEV bool moments(int nA, Ar2Moments &M) const
    {
        double acov[AR2_NTA], ld;
        if (!ar2_inverse<true>(nA, aprec, acov, ld, false))
            return false;
        M.a1 = am[0];
        M.a2 = am[1];
        M.a1a1 = fma(am[0], am[0], acov[tri(0, 0)]);
        M.a2a2 = fma(am[1], am[1], acov[tri(1, 1)]);
        M.x1 = M.x2 = M.a1x1 = M.a2x2 = M.x1x1 = M.x2x2 = 0.0;
        if (nA >= 3)
        {
            const bool dual = nA == 4;
            M.x1 = am[2];
            M.x2 = dual ? am[3] : am[2];
            M.a1x1 = fma(am[0], am[2], acov[tri(2, 0)]);
            M.a2x2 = dual ? fma(am[1], am[3], acov[tri(3, 1)]) : fma(am[1], am[2], acov[tri(2, 1)]);
            M.x1x1 = fma(am[2], am[2], acov[tri(2, 2)]);
            M.x2x2 = dual ? fma(am[3], am[3], acov[tri(3, 3)]) : M.x1x1;
        }
        return true;
    }

    /* line weights of X = phi1 Q_1 + phi2 Q_2 (see the table at the top) */
    FAB_DEV bool lines(int nA, Ar2Lines &L) const
    {
        Ar2Moments M;
        if (!moments(nA, M))
            return false;
        const double w1 = nb[0] * nc[0], w2 = nb[1] * nc[1];
        L.d1a = fma(w2, M.x2x2, w1);
        L.d1b = w1 * M.a1a1;
        L.d2a = fma(w1, M.x1x1, w2);
        L.d2b = w2 * M.a2a2;
        L.o11 = -(w1 * M.a1);
        L.o22 = -(w2 * M.a2);
        L.o12 = -fma(w1, M.x1, w2 * M.x2);
        L.oh = w1 * M.a1x1;
        L.oi = w2 * M.a2x2;
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

    /* noisemodel_ar.cc:558-610 with Q = (J'XJ, J'Xr, r'Xr) (LMalpha is ignored by the AR model) */
    FAB_DEV bool update_theta(const Stats<P> &Q, const double (&c)[P], bool want_logdet)
    {
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

    /* UpdateAlpha :447-528 then UpdatePhi :530-556 from the nine line sums. Returns 0 or a FABBER_VOX_* code. */
    FAB_DEV int update_noise(const VbArgs &a, int nA, const Ar2Klj &K)
    {
        const double w1 = nb[0] * nc[0], w2 = nb[1] * nc[1];
        const double pp = a.ar_alpha_prior_prec;
        const bool dual = nA == 4;
#pragma unroll
        for (int i = 0; i < AR2_NTA; i++)
            aprec[i] = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            aprec[tri(i, i)] = (i < nA) ? pp : 0.0;
        /* OpKLJ(M(n,2,0)) = b / e;  0.5 OpKLJ(M(n,1,1)) = h / i;  OpKLJ(M(n,0,2)) = d / a */
        aprec[tri(0, 0)] += w1 * K.b;
        aprec[tri(1, 1)] += w2 * K.e;
        if (nA > 2)
        {
            aprec[tri(2, 0)] += 0.5 * w1 * (2.0 * K.h);
            aprec[tri(2, 2)] += w1 * K.d;
            if (dual)
            {
                aprec[tri(3, 1)] += 0.5 * w2 * (2.0 * K.i);
                aprec[tri(3, 3)] += w2 * K.a;
            }
            else
            {
                aprec[tri(2, 1)] += 0.5 * w2 * (2.0 * K.i);
                aprec[tri(2, 2)] += w2 * K.a;
            }
        }
        bool fin = true;
#pragma unroll
        for (int i = 0; i < AR2_NTA; i++)
            fin = fin && finite_d(aprec[i]);
        if (!fin)
            return FABBER_VOX_NONFINITE_F; /* "Non-finite values in alpha precisions" :489 */
        double acov[AR2_NTA], ld;
        if (!ar2_inverse<false>(nA, aprec, acov, ld, false))
            return FABBER_VOX_SINGULAR;
        double mn = fmin(acov[tri(0, 0)], acov[tri(1, 1)]);
        if (nA > 2)
            mn = fmin(mn, acov[tri(2, 2)]);
        if (dual)
            mn = fmin(mn, acov[tri(3, 3)]);
        if (mn < 0)
            return FABBER_VOX_AR_NEG_VARIANCE;
        /* means = Cov (prior_prec prior_means + tmp), prior means 0; -0.5 OpKLJ(M(n,1,0)) = c / f, (n,0,1) = g */
        double tmp[4] = { 0.0, 0.0, 0.0, 0.0 };
        tmp[0] += -0.5 * w1 * (-2.0 * K.c);
        tmp[1] += -0.5 * w2 * (-2.0 * K.f);
        if (nA > 2)
        {
            tmp[2] += -0.5 * w1 * (-2.0 * K.g);
            if (dual)
                tmp[3] += -0.5 * w2 * (-2.0 * K.g);
            else
                tmp[2] += -0.5 * w2 * (-2.0 * K.g);
        }
        if (!ar2_inverse<true>(nA, aprec, acov, ld, false))
            return FABBER_VOX_SINGULAR;
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                s += acov[tri(i, j)] * tmp[j];
            am[i] = (i < nA) ? s : 0.0;
        }
        Ar2Moments M;
        if (!moments(nA, M))
            return FABBER_VOX_SINGULAR;
        /* UpdatePhi: k'Q_n k + tr(Sigma J'Q_n J) = OpKLJ(Q_n), linear in the lines */
        const double t1 = K.a + M.a1 * (-2.0 * K.c) + M.a1a1 * K.b + M.x1 * (-2.0 * K.g) + M.a1x1 * (2.0 * K.h) + M.x1x1 * K.d;
        const double t2 = K.d + M.a2 * (-2.0 * K.f) + M.a2a2 * K.e + M.x2 * (-2.0 * K.g) + M.a2x2 * (2.0 * K.i) + M.x2x2 * K.a;
        const double half = ((double)(a.T >> 1) - 1) * 0.5;
        nb[0] = 1 / (t1 * 0.5 + 1 / a.noise_prior_b[0]);
        nc[0] = half + a.noise_prior_c[0];
        nb[1] = 1 / (t2 * 0.5 + 1 / a.noise_prior_b[1]);
        nc[1] = half + a.noise_prior_c[1];
        return 0;
    }

    /* noisemodel_ar.cc:643-747 with c == m (d = 0): k'Xk = Q.rr */
    FAB_DEV double free_energy(const VbArgs &a, int nA, const Stats<P> &Q) const
    {
        const double log2pi = FAB_LOG_2PI;
        const double nTm1 = (double)(a.T >> 1) - 1;
        double acov[AR2_NTA], ldA;
        ar2_inverse<false>(nA, aprec, acov, ldA, true);
        const double elAlpha = 0.5 * ldA - 0.5 * nA * (log2pi + 1);
        const double elTheta = 0.5 * logdetLam - 0.5 * P * (log2pi + 1);
        double elPhi = 0.0, p0 = 0.0, p9 = 0.0;
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            const double si = nb[i], ci = nc[i], siP = a.noise_prior_b[i], ciP = a.noise_prior_c[i];
            const double dg = digamma_fsl(ci), lsi = log(si);
            elPhi += -gammaln(ci) - ci * lsi - ci + (ci - 1) * (dg + lsi);
            p0 += (dg + lsi) * (nTm1 * 0.5 + ciP - 1);
            p9 += -2 * gammaln(ciP) - 2 * ciP * log(siP) - si * ci / siP;
        }
        const double p1 = -log2pi * (nTm1 + 0.5 * nA + 0.5 * P);
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
        double p6 = 0.0, p7 = 0.0, p8 = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < nA)
            {
                p6 += log(fabs(pp));
                p7 += am[i] * pp * am[i];
                p8 += acov[tri(i, i)] * pp;
            }
        p6 *= 0.5;
        p7 *= -0.5;
        p8 *= -0.5;
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

template <class Model>
__global__ void __launch_bounds__(VB_BLOCK, FAB_AR_MIN_BLOCKS) vb_voxelwise_ar2_kernel(const __grid_constant__ VbArgs a)
{
    constexpr int P = Model::P;
    constexpr int NT = NTri<P>::value;
    typedef Ar2Voxel<Model> Vox;
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
    const int nA = a.ar_n_alphas;
    const bool cross = nA > 2;
    const int n_tri_a = nA * (nA + 1) / 2;

    Vox X;
    int status = 0;
    double F = 1234.5678;
    int it = 0;

    /* ---- SetupPerVoxelDists (as vb_voxelwise_ar_kernel) ------------------------------------------ */
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
    /* noise: hard-coded initial dists (noisemodel_ar.cc:379-403) or the restart values, then Precalculate
     * (:749-769): c = c_prior + (nT-1)/2 for both echoes */
#pragma unroll
    for (int i = 0; i < 4; i++)
        X.am[i] = 0.0;
#pragma unroll
    for (int i = 0; i < AR2_NTA; i++)
        X.aprec[i] = 0.0;
    if (a.init_noise)
    {
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            X.nb[i] = a.init_noise[(size_t)(2 * i) * N + v];
            X.nc[i] = a.init_noise[(size_t)(2 * i + 1) * N + v];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < nA)
                X.am[i] = a.init_noise[(size_t)(4 + i) * N + v];
#pragma unroll
        for (int i = 0; i < AR2_NTA; i++)
            if (i < n_tri_a)
                X.aprec[i] = a.init_noise[(size_t)(4 + nA + i) * N + v];
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 2; i++)
        {
            X.nb[i] = a.noise_post_b[i];
            X.nc[i] = a.noise_post_c[i];
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
            if (i < nA)
                X.aprec[tri(i, i)] = a.ar_alpha_prior_prec;
    }
#pragma unroll
    for (int i = 0; i < 2; i++)
        X.nc[i] = a.noise_prior_c[i] + ((double)(a.T >> 1) - 1) * 0.5;
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
    Stats<P> Q;
    double c[P];
    Conv conv;
    conv.init(a.conv_type, a.max_iterations, a.fchange, a.max_trials);
    const bool use_snap = a.conv_type == FABBER_CONV_TRIALMODE || a.conv_type == FABBER_CONV_FREDUCE;
    double Fprior = 0.0;
    int phase = PH_SETUP;
    while (status == 0)
    {
#pragma unroll
        for (int i = 0; i < P; i++)
            c[i] = X.m[i];
        Ar2Lines L;
        if (!X.lines(nA, L))
        {
            status = FABBER_VOX_SINGULAR | (phase == PH_SETUP ? FABBER_VOX_SETUP_FLAG : 0);
            break;
        }
        X.template put<true>(park);
        const int err = ar2_pass_theta<Model>(a, mc, v, c, L, cross, Q);
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
                F = X.free_energy(a, nA, Q) + Fprior;
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
        if (!X.update_theta(Q, c, a.need_f != 0))
        {
            status = FABBER_VOX_SINGULAR;
            break;
        }
        /* second pass: the OperatorKLJ line sums about the NEW theta */
        double d[P];
#pragma unroll
        for (int i = 0; i < P; i++)
            d[i] = c[i] - X.m[i];
        Ar2Klj K;
        X.template put<true>(park);
        ar2_pass_noise<Model>(a, mc, v, c, d, X.Sig, cross, K);
        X.template get<true>(park);
        const int nerr = X.update_noise(a, nA, K);
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
#pragma unroll
    for (int i = 0; i < 2; i++)
    {
        a.noise[(size_t)(2 * i) * N + v] = X.nb[i];
        a.noise[(size_t)(2 * i + 1) * N + v] = X.nc[i];
    }
#pragma unroll
    for (int i = 0; i < 4; i++)
        if (i < nA)
            a.noise[(size_t)(4 + i) * N + v] = X.am[i];
#pragma unroll
    for (int i = 0; i < AR2_NTA; i++)
        if (i < n_tri_a)
            a.noise[(size_t)(4 + nA + i) * N + v] = X.aprec[i];
    if (a.free_energy)
        a.free_energy[v] = F;
    if (a.iterations)
        a.iterations[v] = it;
    a.status[v] = status;
}

} // namespace fab
