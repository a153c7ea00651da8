/*
 * vb_inst.cu - one translation unit per forward model: compiled repeatedly by the Makefile with
 *   -DFAB_FAMILY=<LinearModel|PolyModel|ExpModel> -DFAB_K=<template argument> -DFAB_GETTER=<symbol>
 * A model plug-in (include/fabber_model_plugin.h) includes this file after defining its model struct, with
 *   #define FAB_MODEL_TYPE <the struct>  and  #define FAB_GETTER <symbol>
 * Instantiates every kernel that model needs and exports its launcher table.
 */
#include <cstdlib>

#include "vb_launch.h"
#include "vb_voxelwise.cuh"
#include "vb_voxelwise_ar.cuh"
#include "vb_voxelwise_ar2.cuh"
#include "vb_nlls.cuh"
#include "vb_spatial.cuh"

#if !defined(FAB_GETTER) || !(defined(FAB_MODEL_TYPE) || (defined(FAB_FAMILY) && defined(FAB_K)))
#error "compile with -DFAB_FAMILY=.. -DFAB_K=.. -DFAB_GETTER=.. (or define FAB_MODEL_TYPE and FAB_GETTER)"
#endif

namespace fab
{
void count_launch();

#ifdef FAB_MODEL_TYPE
typedef FAB_MODEL_TYPE M;
#else
typedef FAB_FAMILY<FAB_K> M;
#endif

template <int NPHI, bool SNAP> static cudaError_t launch_white(const VbArgs &a, cudaStream_t s)
{
    if (a.v_end <= a.v_begin)
        return cudaSuccess;
    typedef WhiteVoxel<M, NPHI, SNAP> Vox;
    const size_t smem = M::smem_bytes(a.T) + (size_t)(Vox::STASH_DOUBLES + Vox::SNAP_DOUBLES) * VB_BLOCK * sizeof(double)
        + (NPHI > 1 ? (size_t)a.T : 0);
    auto kern = vb_voxelwise_white_kernel<M, NPHI, SNAP>;
    if (smem > 48 * 1024)
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
    }
    const unsigned grid = (unsigned)((a.v_end - a.v_begin + VB_BLOCK - 1) / VB_BLOCK);
    kern<<<grid, VB_BLOCK, smem, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t launch_ar(const VbArgs &a, cudaStream_t s)
{
    if (a.v_end <= a.v_begin)
        return cudaSuccess;
    typedef ArVoxel<M> Vox;
    const bool use_snap = a.conv_type == FABBER_CONV_TRIALMODE || a.conv_type == FABBER_CONV_FREDUCE;
    const size_t smem = M::smem_bytes(a.T)
        + (size_t)(Vox::STASH_DOUBLES + (use_snap ? Vox::SNAP_DOUBLES : 0)) * VB_BLOCK * sizeof(double);
    auto kern = vb_voxelwise_ar_kernel<M>;
    if (smem > 48 * 1024)
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
    }
    const unsigned grid = (unsigned)((a.v_end - a.v_begin + VB_BLOCK - 1) / VB_BLOCK);
    kern<<<grid, VB_BLOCK, smem, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t launch_ar2(const VbArgs &a, cudaStream_t s)
{
    if (a.v_end <= a.v_begin)
        return cudaSuccess;
    typedef Ar2Voxel<M> Vox;
    const bool use_snap = a.conv_type == FABBER_CONV_TRIALMODE || a.conv_type == FABBER_CONV_FREDUCE;
    const size_t smem = M::smem_bytes(a.T)
        + (size_t)(Vox::STASH_DOUBLES + (use_snap ? Vox::SNAP_DOUBLES : 0)) * VB_BLOCK * sizeof(double);
    auto kern = vb_voxelwise_ar2_kernel<M>;
    if (smem > 48 * 1024)
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
    }
    const unsigned grid = (unsigned)((a.v_end - a.v_begin + VB_BLOCK - 1) / VB_BLOCK);
    kern<<<grid, VB_BLOCK, smem, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t launch_nlls(const VbArgs &a, cudaStream_t s)
{
    if (a.v_end <= a.v_begin)
        return cudaSuccess;
    const size_t smem = M::smem_bytes(a.T);
    auto kern = nlls_kernel<M>;
    if (smem > 48 * 1024)
    {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess)
            return e;
    }
    const unsigned grid = (unsigned)((a.v_end - a.v_begin + VB_BLOCK - 1) / VB_BLOCK);
    kern<<<grid, VB_BLOCK, smem, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t launch_model_fit(const VbArgs &a, cudaStream_t s)
{
    if (a.N <= 0)
        return cudaSuccess;
    const size_t smem = M::smem_bytes(a.T);
    auto kern = model_fit_kernel<M>;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)((a.N + 255) / 256), 256, smem, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

static cudaError_t launch_sp_setup(const SpArgs &s, cudaStream_t st)
{
    const size_t smem = M::smem_bytes(s.v.T);
    auto kern = sp_setup_kernel<M>;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)((s.v.N + VB_BLOCK - 1) / VB_BLOCK), VB_BLOCK, smem, st>>>(s);
    count_launch();
    return cudaGetLastError();
}
static cudaError_t launch_sp_noise(const SpArgs &s, cudaStream_t st)
{
    const size_t smem = M::smem_bytes(s.v.T) + (size_t)NTri<M::P>::value * VB_BLOCK * sizeof(double);
    auto kern = sp_noise_kernel<M>;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<(unsigned)((s.v.N + VB_BLOCK - 1) / VB_BLOCK), VB_BLOCK, smem, st>>>(s);
    count_launch();
    return cudaGetLastError();
}
static cudaError_t launch_sp_ak_partial(const SpArgs &s, cudaStream_t st)
{
    const bool slab = s.link.world > 1;
    if (s.ignore_bad)
    {
        if (slab)
            sp_ak_partial_kernel<M::P, true, true><<<SP_AK_BLOCKS, 256, 0, st>>>(s);
        else
            sp_ak_partial_kernel<M::P, true, false><<<SP_AK_BLOCKS, 256, 0, st>>>(s);
    }
    else if (slab)
        sp_ak_partial_kernel<M::P, false, true><<<SP_AK_BLOCKS, 256, 0, st>>>(s);
    else
        sp_ak_partial_kernel<M::P, false, false><<<SP_AK_BLOCKS, 256, 0, st>>>(s);
    count_launch();
    return cudaGetLastError();
}
static cudaError_t launch_sp_ak_final(const SpArgs &s, cudaStream_t st)
{
    sp_ak_final_kernel<M::P><<<1, 256, 0, st>>>(s);
    count_launch();
    return cudaGetLastError();
}
static cudaError_t launch_sp_theta(const SpArgs &s, cudaStream_t st)
{
    sp_theta_kernel<M::P><<<(unsigned)((s.v.N + VB_BLOCK - 1) / VB_BLOCK), VB_BLOCK, 0, st>>>(s);
    count_launch();
    return cudaGetLastError();
}
static cudaError_t launch_sp_sweep(const SpArgs &s, cudaStream_t st)
{
    /* persistent cooperative kernel: as many CTAs as can be co-resident (grid-wide barrier inside) */
    static int grid = 0, max_grid = 0;
    /* one instantiation per (failed voxels struck from neighbour lists?, z-slab coupling?): both are uniform for
     * a launch, and the plain one-GPU sweep must not carry the others' look-ups */
    const bool slab = s.link.world > 1;
    const void *kern_any = s.ignore_bad ? (slab ? (const void *)sp_sweep_kernel<M::P, true, true>
                                                : (const void *)sp_sweep_kernel<M::P, true, false>)
                                        : (slab ? (const void *)sp_sweep_kernel<M::P, false, true>
                                                : (const void *)sp_sweep_kernel<M::P, false, false>);
    auto kern = sp_sweep_kernel<M::P, true, true>; /* the largest: sizes the grid for all four */
    if (grid == 0)
    {
        int dev = 0, sms = 0, per_sm = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, SP_SWEEP_BLOCK, 0);
        if (e != cudaSuccess)
            return e;
        if (per_sm < 1)
            return cudaErrorLaunchOutOfResources;
        const int want = 1; /* one CTA of 480 voxel threads + the barrier warp per SM */
        max_grid = sms * per_sm;
        grid = sms * (per_sm > want ? want : per_sm);
    }
    int launch_grid = grid;
    if (s.sweep_share > 1 && max_grid / s.sweep_share < launch_grid)
        launch_grid = max_grid / s.sweep_share > 0 ? max_grid / s.sweep_share : 1;
    if (s.sweep_max_ctas > 0 && s.sweep_max_ctas < launch_grid)
        launch_grid = s.sweep_max_ctas;
    cudaMemsetAsync(s.sweep_barrier, 0, sizeof(unsigned), st); /* the barrier's arrival counter only grows */
    void *args[] = { (void *)&s };
    cudaError_t e = cudaLaunchCooperativeKernel(kern_any, dim3(launch_grid), dim3(SP_SWEEP_BLOCK), args, 0, st);
    count_launch();
    return e;
}

static cudaError_t preload_spatial()
{
    const void *kernels[] = { (const void *)sp_setup_kernel<M>, (const void *)sp_noise_kernel<M>,
        (const void *)sp_theta_kernel<M::P>, (const void *)sp_ak_final_kernel<M::P>,
        (const void *)sp_ak_partial_kernel<M::P, false, false>, (const void *)sp_ak_partial_kernel<M::P, false, true>,
        (const void *)sp_ak_partial_kernel<M::P, true, false>, (const void *)sp_ak_partial_kernel<M::P, true, true>,
        (const void *)sp_sweep_kernel<M::P, false, false>, (const void *)sp_sweep_kernel<M::P, false, true>,
        (const void *)sp_sweep_kernel<M::P, true, false>, (const void *)sp_sweep_kernel<M::P, true, true> };
    for (size_t i = 0; i < sizeof(kernels) / sizeof(kernels[0]); i++)
    {
        cudaFuncAttributes at;
        cudaError_t e = cudaFuncGetAttributes(&at, kernels[i]);
        if (e != cudaSuccess)
            return e;
    }
    return cudaSuccess;
}

static const ModelLaunchers g_launchers = {
    launch_white<1, false>,
    launch_white<1, true>,
    launch_white<FABBER_CUDA_MAX_PHIS, true>,
    launch_ar,
    launch_model_fit,
    launch_sp_setup,
    launch_sp_ak_partial,
    launch_sp_ak_final,
    launch_sp_theta,
    launch_sp_sweep,
    launch_sp_noise,
    preload_spatial,
    launch_ar2,
    launch_nlls,
};

#ifdef FAB_MODEL_TYPE
extern "C" const void *FAB_GETTER() { return &g_launchers; } /* plug-ins hand the table out through the C ABI */
#else
const ModelLaunchers *FAB_GETTER() { return &g_launchers; }
#endif

} // namespace fab
