/*
 * vb_inst.cu - one translation unit per forward model: compiled repeatedly by the Makefile with
 *   -DFAB_FAMILY=<LinearModel|PolyModel|ExpModel> -DFAB_K=<template argument> -DFAB_GETTER=<symbol>
 * Instantiates every kernel that model needs and exports its launcher table.
 */
#include "vb_launch.h"
#include "vb_voxelwise.cuh"

#if !defined(FAB_FAMILY) || !defined(FAB_K) || !defined(FAB_GETTER)
#error "compile with -DFAB_FAMILY=.. -DFAB_K=.. -DFAB_GETTER=.."
#endif

namespace fab
{
void count_launch();

typedef FAB_FAMILY<FAB_K> M;

template <int NPHI, bool SNAP> static cudaError_t launch_white(const VbArgs &a, cudaStream_t s)
{
    if (a.N <= 0)
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
    const unsigned grid = (unsigned)((a.N + VB_BLOCK - 1) / VB_BLOCK);
    kern<<<grid, VB_BLOCK, smem, s>>>(a);
    count_launch();
    return cudaGetLastError();
}

static const ModelLaunchers g_launchers = {
    launch_white<1, false>,
    launch_white<1, true>,
    launch_white<FABBER_CUDA_MAX_PHIS, true>,
    nullptr,
    nullptr,
    nullptr,
    nullptr,
    nullptr,
};

const ModelLaunchers *FAB_GETTER() { return &g_launchers; }

} // namespace fab
