/*
 * sine_model.cu - device half of the example plug-in model "sine":
 *     g(t) = a * sin(b * (t_i - c)) + d,   t_i = i * dt,  i = 0..T-1,   parameters (a, b, c, d)
 * One translation unit per model: the struct with the __device__ hooks, then vb_inst.cu, which instantiates
 * the library's VB kernels (white / AR(1) / spatial / model fit) for it and exports the launcher table.
 * The arithmetic is written with explicit round-to-nearest intrinsics for the same reason as the built-in
 * models (no FMA contraction under the finite-difference Jacobian, vb_models.cuh).
 */
#include "vb_models.cuh"

namespace fab
{
struct SineModel
{
    static constexpr int P = 4;
    static constexpr int ID = FABBER_MODEL_PLUGIN;
    struct Ctx
    {
        double dt;
    };
    struct Sample
    {
        double tt;
    };
    static __host__ __device__ size_t smem_bytes(int) { return 0; }
    template <class Args> static FAB_DEV void stage(const Args &, double *) {}
    template <class Args> static FAB_DEV Ctx make_ctx(const Args &a, double *)
    {
        Ctx c;
        c.dt = a.model_consts[0];
        return c;
    }
    static FAB_DEV double one(double tt, double a, double b, double c, double d)
    {
        return __dadd_rn(__dmul_rn(a, sin(__dmul_rn(b, __dadd_rn(tt, -c)))), d);
    }
    static FAB_DEV double eval(const Ctx &c, int t, const double (&p)[P])
    {
        return one(__dmul_rn((double)t, c.dt), p[0], p[1], p[2], p[3]);
    }
    static constexpr bool HAS_FAST = false;
    static constexpr bool LINEAR = false;
    static FAB_DEV bool fast_ok(const Ctx &, int, const double (&)[P], const double (&)[P], const double (&)[P])
    {
        return false;
    }
    static FAB_DEV void sample(const Ctx &c, int t, Sample &s) { s.tt = __dmul_rn((double)t, c.dt); }
    /* the 2P+1 evaluations of one sample; sin(b (t - c)) is shared where only a or d moves */
    template <bool FAST>
    static FAB_DEV void eval_fd(const Ctx &, const Sample &smp, const double (&p0)[P], const double (&pp)[P],
        const double (&pn)[P], double &g, double (&gp)[P], double (&gn)[P])
    {
        const double tt = smp.tt;
        const double s0 = sin(__dmul_rn(p0[1], __dadd_rn(tt, -p0[2])));
        const double as0 = __dmul_rn(p0[0], s0);
        g = __dadd_rn(as0, p0[3]);
        gp[0] = __dadd_rn(__dmul_rn(pp[0], s0), p0[3]);
        gn[0] = __dadd_rn(__dmul_rn(pn[0], s0), p0[3]);
        gp[1] = one(tt, p0[0], pp[1], p0[2], p0[3]);
        gn[1] = one(tt, p0[0], pn[1], p0[2], p0[3]);
        gp[2] = one(tt, p0[0], p0[1], pp[2], p0[3]);
        gn[2] = one(tt, p0[0], p0[1], pn[2], p0[3]);
        gp[3] = __dadd_rn(as0, pp[3]);
        gn[3] = __dadd_rn(as0, pn[3]);
    }
    template <class Args> static FAB_DEV void init_voxel(const Args &, int, double (&)[P]) {}
};
} // namespace fab

#define FAB_MODEL_TYPE SineModel
#define FAB_GETTER fabber_example_sine_launchers
#include "vb_inst.cu"
