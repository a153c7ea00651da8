/*
 * vb_launch.h - host-side launcher table for the templated VB kernels.
 *
 * Every (model family, parameter count) pair is compiled in its own translation unit
 * (vb_inst.cu with -DFAB_FAMILY/-DFAB_K, see Makefile) so the build parallelises; each unit exports
 * one getter returning the launchers for that model.
 */
#pragma once
#include <cuda_runtime.h>

namespace fab
{
struct VbArgs;
struct SpArgs;

typedef cudaError_t (*VbLaunchFn)(const VbArgs &, cudaStream_t);
typedef cudaError_t (*SpLaunchFn)(const SpArgs &, cudaStream_t);

struct ModelLaunchers
{
    VbLaunchFn white_fast;      /* one phi, no masked samples, detectors without a real snapshot */
    VbLaunchFn white_fast_snap; /* same, trialmode / freduce (snapshot + revert) */
    VbLaunchFn white_general;   /* noise patterns (<= FABBER_CUDA_MAX_PHIS) and masked samples */
    VbLaunchFn ar1;             /* AR(1) noise */
    VbLaunchFn model_fit;       /* batched model evaluation */
    /* spatial mode (vb_spatial.cuh) */
    SpLaunchFn sp_setup, sp_ak_partial, sp_ak_final, sp_theta, sp_sweep, sp_noise;
    /* load every spatial kernel variant on the current device NOW. CUDA loads kernels lazily, and loading one may
     * have to wait for kernels that are running - with slabs that spin on each other's flags a first launch in
     * the middle of an iteration would deadlock until the spin gives up. Appended: older plug-ins leave it NULL. */
    cudaError_t (*sp_preload)(void);
    VbLaunchFn ar2; /* AR(1) noise on two interleaved echoes, with or without cross terms (appended, may be NULL) */
    VbLaunchFn nlls; /* --method=nlls: non-linear least squares per voxel (appended, may be NULL) */
};

/* number of blocks the aK partial reduction is launched with (size of SpArgs::ak_partial) */
constexpr int SP_AK_BLOCKS = 1184; /* 148 SMs x 8 */

/* returns NULL when no device Evaluate hook is compiled for that size */
const ModelLaunchers *find_model(int model_id, int n_params);

} // namespace fab
