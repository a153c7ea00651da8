/*
 * fabber_model_plugin.h - forward models as plug-in libraries (the --loadmodels / fabber_load_models
 * mechanism of the reference: fwdmodel.cc:63-129, rundata.cc:440-443, fabber_capi.cc:60-76; example
 * library: examples/exp_models.cc).
 *
 * A plug-in is a shared library that exports the reference's three symbols plus an ABI check:
 *
 *     int          get_num_models(void);
 *     const char  *get_model_name(int index);
 *     NewInstanceFptr get_new_instance_func(const char *name);   // fabber_b200::FwdModel *(*)(void)
 *     int          fabber_b200_plugin_abi(void);                  // returns FABBER_B200_PLUGIN_ABI
 *
 * Where the reference's plug-in carries a C++ EvaluateModel, this one carries KERNELS: each model is
 *   (1) a device struct with the hooks documented in fabber_core_b200/csrc/vb_models.cuh (P, Ctx, Sample,
 *       stage, make_ctx, eval, sample, eval_fd, init_voxel ...) compiled with the library's kernel templates
 *       by including vb_inst.cu - one translation unit per model:
 *
 *           #include "vb_models.cuh"
 *           struct MyModel { static constexpr int P = 3; ... };
 *           #define FAB_MODEL_TYPE MyModel
 *           #define FAB_GETTER my_model_launchers      // extern "C" const void *my_model_launchers(void)
 *           #include "vb_inst.cu"
 *
 *   (2) a host class derived from fabber_b200::FwdModel (fabber_core_b200/host/fabber_host.h: options,
 *       parameter defaults, EvaluateModel for --evaluate) whose GetDeviceModel() fills
 *
 *           m.id = FABBER_MODEL_PLUGIN;  m.n_params = P;  m.plugin_launchers = my_model_launchers();
 *           m.consts[0..15] = scalars the device hooks read from args.model_consts;
 *           m.design / m.design_len = an optional vector (HOST pointer) the hooks find behind args.design.
 *
 * Build: `make -C fabber_core_b200/csrc plugin PLUGIN_SRC="a.cu b.cu host.cc" PLUGIN_OUT=libmine.so`
 * (nvcc for sm_100a, linked against libfabber_cuda.so and libfabbercore_b200.so).
 * fabber_core_b200/examples/ holds a complete one (models "sine" and "exp").
 */
#ifndef FABBER_MODEL_PLUGIN_H
#define FABBER_MODEL_PLUGIN_H

#include "fabber_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* the exports of a plug-in library, as the loader resolves them */
int get_num_models(void);
const char *get_model_name(int index);
void *(*get_new_instance_func(const char *name))(void); /* really fabber_b200::FwdModel *(*)(void) */
int fabber_b200_plugin_abi(void);

#ifdef __cplusplus
}
#endif
#endif
