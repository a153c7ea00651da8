/*
 * exp_model.cu - device half of the plug-in's "exp" model. In the reference the sum-of-exponentials model
 * IS the example plug-in (examples/exp_models.cc, fwdmodel_exp.cc); here the same device struct the core
 * library compiles in (ExpModel<2>, vb_models.cuh) is built a second time, as a plug-in, so that the
 * plug-in path can be checked bit for bit against the built-in one.
 */
#include "vb_models.cuh"

namespace fab
{
/* the bi-exponential, reading its sample spacing from the plug-in constants instead of VbArgs::exp_dt */
struct PluginBiExpModel : ExpModel<2>
{
    template <class Args> static FAB_DEV Ctx make_ctx(const Args &a, double *smem)
    {
        Ctx c;
        c.dt = a.model_consts[0];
        c.tab = smem;
        return c;
    }
};
} // namespace fab

#define FAB_MODEL_TYPE PluginBiExpModel
#define FAB_GETTER fabber_example_exp_launchers
#include "vb_inst.cu"
