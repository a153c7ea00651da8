/*
 * vb_models.cuh - forward models as __device__ Evaluate hooks.
 *
 * Each model is a struct with
 *   P                      number of parameters (compile time)
 *   Ctx                    per-thread constants (pointers into shared memory, scalars)
 *   smem_bytes(args)       dynamic shared memory the model wants (design / timing vectors)
 *   stage(args, smem)      block-cooperative staging into shared memory
 *   make_ctx(args, smem)   builds Ctx
 *   eval(ctx, t, p)        one sample of EvaluateModel at model-space parameters p
 *   sample(ctx, t, smp); eval_fd(ctx, smp, p0, pp, pn, g, gp, gn)
 *                          the 2P+1 evaluations LinearizedFwdModel::ReCentre needs at sample t:
 *                          g = f(p0), gp[i] = f(p0 with p0[i] -> pp[i]), gn[i] likewise with pn[i].
 *                          Every value is bit-identical to calling eval() on the perturbed vector;
 *                          models override it only to share sub-expressions that provably do not
 *                          depend on the perturbed element (e.g. exp(-r t) when an amplitude moves).
 *   LINEAR / basis_row()   models that are LINEAR IN THEIR MODEL-SPACE PARAMETERS (g = sum_j phi_j(t) p_j) say so
 *                          and hand out phi: the central difference of such a model is phi_j times the
 *                          difference quotient of the parameter transform, up to the rounding noise of the
 *                          subtraction, so the pass needs one evaluation per sample instead of 2P+1
 *                          (recentre_loop in vb_voxelwise.cuh; opt-in with FABBER_B200_BASIS_JACOBIAN=1)
 *   HAS_FAST / fast_ok()   optional: a cheaper eval_fd<true> that is valid only for a range of parameters
 *                          (exp: the table-based exponential of vb_exp.cuh needs |r t| < 708); the pass
 *                          checks fast_ok() once, outside the time loop, and otherwise runs eval_fd<false>
 *   init_voxel(...)        FwdModel::InitVoxelPosterior hook (model-space means)
 *
 * The arithmetic inside eval() follows the reference's EvaluateModel operation by operation, with
 * explicit round-to-nearest intrinsics so that nvcc does not contract mul+add into FMA: the
 * finite-difference Jacobian amplifies last-bit differences of g by ~1e5 (SURVEY.md hard part 1).
 *
 * Reference: fwdmodel_linear.cc:92-96, fwdmodel_poly.cc:62-80, examples/fwdmodel_exp.cc:65-91.
 */
#pragma once
#include "vb_device.cuh"
#include "vb_exp.cuh"

namespace fab
{

struct VbArgs;

/* ---------------------------------------------------------------------------------------------
 * linear:  result = design * (params - 0) + 0          (fwdmodel_linear.cc:95)
 * design is staged in shared memory, row-major [T][P]; all threads of a warp read the same row
 * (broadcast, conflict free).
 * ------------------------------------------------------------------------------------------- */
template <int P_> struct LinearModel
{
    static constexpr int P = P_;
    static constexpr int ID = FABBER_MODEL_LINEAR;
    struct Ctx
    {
        const double *design;
    };
    /* a design too long for shared memory next to the kernels' parked state (> 64 KB: T P > 8192) is read from
     * global memory instead - every thread of a warp reads the same row, so it is one broadcast load from L1 / L2 */
    static constexpr size_t STAGE_MAX_BYTES = 64 * 1024;
    static __host__ __device__ size_t smem_bytes(int T)
    {
        const size_t b = (size_t)T * P * sizeof(double);
        return b <= STAGE_MAX_BYTES ? b : 0;
    }
    template <class Args> static FAB_DEV void stage(const Args &a, double *smem)
    {
        if (smem_bytes(a.T) == 0)
            return;
        for (int i = threadIdx.x; i < a.T * P; i += blockDim.x)
            smem[i] = a.design[i];
    }
    template <class Args> static FAB_DEV Ctx make_ctx(const Args &a, double *smem)
    {
        Ctx c;
        c.design = smem_bytes(a.T) ? smem : a.design;
        return c;
    }
    static FAB_DEV double eval(const Ctx &c, int t, const double (&p)[P])
    {
        /* (0.0 + x) is x: the first add of the reference's zero-initialised accumulator is skipped */
        const double *row = c.design + t * P;
        double s = __dmul_rn(row[0], p[0]);
#pragma unroll
        for (int j = 1; j < P; j++)
            s = __dadd_rn(s, __dmul_rn(row[j], p[j]));
        return s;
    }
    static constexpr bool HAS_FAST = false;
    static FAB_DEV bool fast_ok(const Ctx &, int, const double (&)[P], const double (&)[P], const double (&)[P])
    {
        return false;
    }
    struct Sample
    {
        int t;
    };
    static FAB_DEV void sample(const Ctx &, int t, Sample &s) { s.t = t; }
    /* linear in its model-space parameters: g = sum_j phi_j(t) p_j. basis_row returns g (eval()'s operation
     * order) and phi - see recentre_loop for what that buys */
    static constexpr bool LINEAR = true;
    static FAB_DEV void basis_row(const Ctx &c, const Sample &smp, const double (&p0)[P], double &g, double (&phi)[P])
    {
        const double *row = c.design + smp.t * P;
#pragma unroll
        for (int j = 0; j < P; j++)
            phi[j] = row[j];
        double s = __dmul_rn(phi[0], p0[0]);
#pragma unroll
        for (int j = 1; j < P; j++)
            s = __dadd_rn(s, __dmul_rn(phi[j], p0[j]));
        g = s;
    }
    template <bool FAST>
    static FAB_DEV void eval_fd(const Ctx &c, const Sample &smp, const double (&p0)[P], const double (&pp)[P],
        const double (&pn)[P], double &g, double (&gp)[P], double (&gn)[P])
    {
        const double *row = c.design + smp.t * P;
        double d[P], prod[P];
#pragma unroll
        for (int j = 0; j < P; j++)
        {
            d[j] = row[j];
            prod[j] = __dmul_rn(d[j], p0[j]);
        }
        /* prefix[j] = sum of the first j products, in the reference's left-to-right order
         * (prefix[1] = 0.0 + prod[0] = prod[0]: the add to the zero-initialised accumulator is skipped) */
        double prefix[P + 1];
        prefix[0] = 0.0;
        prefix[1] = prod[0];
#pragma unroll
        for (int j = 1; j < P; j++)
            prefix[j + 1] = __dadd_rn(prefix[j], prod[j]);
        g = prefix[P];
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            double sp = __dmul_rn(d[i], pp[i]);
            double sn = __dmul_rn(d[i], pn[i]);
            if (i > 0)
            {
                sp = __dadd_rn(prefix[i], sp);
                sn = __dadd_rn(prefix[i], sn);
            }
#pragma unroll
            for (int j = i + 1; j < P; j++)
            {
                sp = __dadd_rn(sp, prod[j]);
                sn = __dadd_rn(sn, prod[j]);
            }
            gp[i] = sp;
            gn[i] = sn;
        }
    }
    template <class Args> static FAB_DEV void init_voxel(const Args &, int, double (&)[P]) {}
};

/* ---------------------------------------------------------------------------------------------
 * poly:  result(i) = sum_n c_n * i^n, i = 1..T, with the power kept in an `int`
 * (fwdmodel_poly.cc:68-79; wraps like the reference for i^(n) >= 2^31).
 * ------------------------------------------------------------------------------------------- */
template <int P_> struct PolyModel
{
    static constexpr int P = P_; /* degree + 1 */
    static constexpr int ID = FABBER_MODEL_POLY;
    struct Ctx
    {
    };
    static __host__ __device__ size_t smem_bytes(int) { return 0; }
    template <class Args> static FAB_DEV void stage(const Args &, double *) {}
    template <class Args> static FAB_DEV Ctx make_ctx(const Args &, double *) { return Ctx(); }
    /* measured and NOT kept (round 2): staging the powers as a [T][P] table in shared memory (no IMAD / I2F in the
     * time loop) made C2 slower, 6.15 -> 6.56 ms - the loop is FP64-pipe bound and the integer / conversion
     * instructions ride in its shadow, while four more LDS per sample sit on the dependency chain */
    static FAB_DEV void powers(int t, double (&pw)[P])
    {
        unsigned int x = 1u, i = (unsigned int)(t + 1);
#pragma unroll
        for (int n = 0; n < P; n++)
        {
            pw[n] = (double)(int)x;
            x *= i;
        }
    }
    static FAB_DEV double eval(const Ctx &, int t, const double (&p)[P])
    {
        double pw[P];
        powers(t, pw);
        double s = __dmul_rn(p[0], pw[0]);
#pragma unroll
        for (int n = 1; n < P; n++)
            s = __dadd_rn(s, __dmul_rn(p[n], pw[n]));
        return s;
    }
    static constexpr bool HAS_FAST = false;
    static FAB_DEV bool fast_ok(const Ctx &, int, const double (&)[P], const double (&)[P], const double (&)[P])
    {
        return false;
    }
    struct Sample
    {
        double pw[P]; /* (t+1)^n, n = 0..P-1 */
    };
    static FAB_DEV void sample(const Ctx &, int t, Sample &s) { powers(t, s.pw); }
    static constexpr bool LINEAR = true; /* g = sum_n (t+1)^n c_n */
    static FAB_DEV void basis_row(const Ctx &, const Sample &smp, const double (&p0)[P], double &g, double (&phi)[P])
    {
#pragma unroll
        for (int n = 0; n < P; n++)
            phi[n] = smp.pw[n];
        double s = __dmul_rn(p0[0], phi[0]);
#pragma unroll
        for (int n = 1; n < P; n++)
            s = __dadd_rn(s, __dmul_rn(p0[n], phi[n]));
        g = s;
    }
    template <bool FAST>
    static FAB_DEV void eval_fd(const Ctx &, const Sample &smp, const double (&p0)[P], const double (&pp)[P],
        const double (&pn)[P], double &g, double (&gp)[P], double (&gn)[P])
    {
        double prod[P], prefix[P + 1];
        const double(&pw)[P] = smp.pw;
        prefix[0] = 0.0;
#pragma unroll
        for (int j = 0; j < P; j++)
        {
            prod[j] = __dmul_rn(p0[j], pw[j]);
            prefix[j + 1] = j == 0 ? prod[0] : __dadd_rn(prefix[j], prod[j]);
        }
        g = prefix[P];
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            double sp = __dmul_rn(pp[i], pw[i]);
            double sn = __dmul_rn(pn[i], pw[i]);
            if (i > 0)
            {
                sp = __dadd_rn(prefix[i], sp);
                sn = __dadd_rn(prefix[i], sn);
            }
#pragma unroll
            for (int j = i + 1; j < P; j++)
            {
                sp = __dadd_rn(sp, prod[j]);
                sn = __dadd_rn(sn, prod[j]);
            }
            gp[i] = sp;
            gn[i] = sn;
        }
    }
    template <class Args> static FAB_DEV void init_voxel(const Args &, int, double (&)[P]) {}
};

/* ---------------------------------------------------------------------------------------------
 * exp:  result(i) = sum_k amp_k * exp(-r_k * (double(i) * dt)), i = 0..T-1
 * (examples/fwdmodel_exp.cc:71-81), parameters (amp_1, r_1, amp_2, r_2, ...).
 * InitVoxelPosterior: amp_k = max(data) / (n + k)  (:84-91).
 * ------------------------------------------------------------------------------------------- */
template <int NE> struct ExpModel
{
    static constexpr int P = 2 * NE;
    static constexpr int ID = FABBER_MODEL_EXP;
    struct Ctx
    {
        double dt;
        const double *tab; /* exp table in shared memory (vb_exp.cuh) */
    };
    static __host__ __device__ size_t smem_bytes(int) { return EXP_TAB_DOUBLES * sizeof(double); }
    template <class Args> static FAB_DEV void stage(const Args &, double *smem) { exp_table_stage(smem); }
    template <class Args> static FAB_DEV Ctx make_ctx(const Args &a, double *smem)
    {
        Ctx c;
        c.dt = a.exp_dt;
        c.tab = smem;
        return c;
    }
    static constexpr bool HAS_FAST = true;
    /* every exponent the pass will form stays inside the table method's range */
    static FAB_DEV bool fast_ok(const Ctx &c, int T, const double (&p0)[P], const double (&pp)[P], const double (&pn)[P])
    {
        const double t_max = __dmul_rn((double)(T > 0 ? T - 1 : 0), c.dt);
        bool ok = true;
#pragma unroll
        for (int k = 0; k < NE; k++)
            ok = ok && exp_fast_range_ok(p0[2 * k + 1], t_max) && exp_fast_range_ok(pp[2 * k + 1], t_max)
                && exp_fast_range_ok(pn[2 * k + 1], t_max);
        return ok;
    }
    /* Inside the table pass the two PERTURBED values of each rate come by series from the centre exponential:
     *     exp(-(r + dr) t) = exp(-r t) * exp(z),  exp(z) = 1 + z (1 + z/2 + z^2/6 + z^3/24 + z^4/120),  z = -dr t.
     * dr = r(c +- delta) - r(c) is five orders of magnitude below r (the finite-difference step is 1e-5 |c|,
     * fwdmodel_linear.cc:157-161): z ~ 1e-5 |c| r t. For |z| < 2^-8 the series is exact to < 1e-17 (z^6/720) - the
     * same <= 1 ULP result a library exp gives, for 7 FP64 instructions instead of 11, and both perturbed values
     * share the rounding error of exp(-r t), which cancels in the difference the Jacobian is made of. A larger z
     * needs r t > 390 / |c|, where exp(-r t) itself has all but vanished: the truncation error of the VALUE is
     * exp(-x) (1e-5 |c| x)^6 / 720 <= 1.6e-31 |c|^6 (x = r t; x^6 e^-x peaks at 116) relative to the amplitude, and
     * of the Jacobian entry 5e-20 of it - both far below the rounding of the sum they enter. So the series is used
     * for every sample of the table pass, without a test (a per-pass choice between passes made whole warps run
     * both once a few voxels' rates had grown: the C5 noise kernel went from 14.5 to 36 ms over ten iterations; a
     * per-sample branch cost C3 15 % - both measured). The bound holds for any |c| a double can reach before
     * r = exp(c) overflows the table pass's own range test. */
    /* e0 * exp(z), |z| small (see above) */
    static FAB_DEV double scaled_exp_small(double e0, double z)
    {
        double q = fma(z, 1.0 / 120.0, 1.0 / 24.0);
        q = fma(z, q, 1.0 / 6.0);
        q = fma(z, q, 0.5);
        q = fma(z, q, 1.0);
        return fma(e0, z * q, e0);
    }
    template <bool FAST> static FAB_DEV double ex(const Ctx &c, double x) { return FAST ? exp_fast(x, c.tab) : exp(x); }
    static FAB_DEV double eval(const Ctx &c, int t, const double (&p)[P])
    {
        double tt = __dmul_rn((double)t, c.dt);
        double s = __dmul_rn(p[0], exp(__dmul_rn(-p[1], tt)));
#pragma unroll
        for (int k = 1; k < NE; k++)
            s = __dadd_rn(s, __dmul_rn(p[2 * k], exp(__dmul_rn(-p[2 * k + 1], tt))));
        return s;
    }
    static constexpr bool LINEAR = false;
    struct Sample
    {
        double tt; /* double(t) * dt */
    };
    static FAB_DEV void sample(const Ctx &c, int t, Sample &s) { s.tt = __dmul_rn((double)t, c.dt); }
    template <bool FAST>
    static FAB_DEV void eval_fd(const Ctx &c, const Sample &smp, const double (&p0)[P], const double (&pp)[P],
        const double (&pn)[P], double &g, double (&gp)[P], double (&gn)[P])
    {
        const double tt = smp.tt;
        double e0[NE], term[NE];
#pragma unroll
        for (int k = 0; k < NE; k++)
        {
            e0[k] = ex<FAST>(c, __dmul_rn(-p0[2 * k + 1], tt));
            term[k] = __dmul_rn(p0[2 * k], e0[k]);
        }
        double s = term[0];
#pragma unroll
        for (int k = 1; k < NE; k++)
            s = __dadd_rn(s, term[k]);
        g = s;
#pragma unroll
        for (int k = 0; k < NE; k++)
        {
            /* amplitude k moves: exp(-r_k t) is unchanged; rate k moves: two new exponentials */
            double ta[4];
            ta[0] = __dmul_rn(pp[2 * k], e0[k]);
            ta[1] = __dmul_rn(pn[2 * k], e0[k]);
            if (FAST)
            {
                const double zp = (p0[2 * k + 1] - pp[2 * k + 1]) * tt, zn = (p0[2 * k + 1] - pn[2 * k + 1]) * tt;
                ta[2] = __dmul_rn(p0[2 * k], scaled_exp_small(e0[k], zp));
                ta[3] = __dmul_rn(p0[2 * k], scaled_exp_small(e0[k], zn));
            }
            else
            {
                ta[2] = __dmul_rn(p0[2 * k], exp(__dmul_rn(-pp[2 * k + 1], tt)));
                ta[3] = __dmul_rn(p0[2 * k], exp(__dmul_rn(-pn[2 * k + 1], tt)));
            }
            double out[4];
#pragma unroll
            for (int q = 0; q < 4; q++)
            {
                double acc = (k == 0) ? ta[q] : term[0];
#pragma unroll
                for (int j = 1; j < NE; j++)
                    acc = __dadd_rn(acc, j == k ? ta[q] : term[j]);
                out[q] = acc;
            }
            gp[2 * k] = out[0];
            gn[2 * k] = out[1];
            gp[2 * k + 1] = out[2];
            gn[2 * k + 1] = out[3];
        }
    }
    template <class Args> static FAB_DEV void init_voxel(const Args &a, int v, double (&m)[P])
    {
        double mx = (double)a.data[v];
        for (int t = 1; t < a.T; t++)
            mx = fmax(mx, (double)a.data[(size_t)t * a.N + v]);
#pragma unroll
        for (int k = 0; k < NE; k++)
            m[2 * k] = mx / (double)(NE + k);
    }
};

} // namespace fab
