/*
 * fabber_cuda.cu - implementation of the thin host <-> device C ABI declared in
 * include/fabber_cuda.h: argument validation, staging of the small per-run constants (design
 * matrix, noise pattern) and dispatch to the templated sm_100a kernels.
 *
 * There is deliberately no CPU path in here: if CUDA is unavailable every entry point fails with
 * FABBER_CUDA_ERR_CUDA.
 */
#include <cub/cub.cuh>

#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "vb_launch.h"
#define FAB_SPATIAL_HOST_KERNELS
#include "vb_spatial.cuh"
#include "vb_voxelwise.cuh"

namespace fab
{
static std::atomic<unsigned long long> g_launches(0);
static std::atomic<double> g_last_multi_ms(0.0);
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static thread_local std::string g_last_error;
static int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}
static int cuda_fail(cudaError_t e, const char *what)
{
    return fail(FABBER_CUDA_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

/* Scratch comes from the stream-ordered allocator. By default the pool hands freed memory back to the
 * driver at the next synchronisation, so a spatial run would re-acquire ~14 GB from the OS on every call
 * (seconds, and erratic - measured). Keep it cached in the pool instead. */
static void keep_pool_memory()
{
    static std::atomic<unsigned long long> done_mask(0); /* one bit per device: a run may span several */
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
        return;
    if (done_mask.load(std::memory_order_relaxed) & (1ull << dev))
        return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
    {
        unsigned long long threshold = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    done_mask.fetch_or(1ull << dev, std::memory_order_relaxed);
}

/* getters exported by the per-model translation units (vb_inst.cu) */
#define FAB_MODEL(id, p, name) const ModelLaunchers *name();
#include "vb_models.inc"
#undef FAB_MODEL

const ModelLaunchers *find_model(const fabber_cuda_model &m, int n_params)
{
    if (m.id == FABBER_MODEL_PLUGIN)
        return static_cast<const ModelLaunchers *>(m.plugin_launchers); /* NULL: reported by the caller */
    return find_model(m.id, n_params);
}

const ModelLaunchers *find_model(int model_id, int n_params)
{
    struct Entry
    {
        int id, p;
        const ModelLaunchers *(*get)();
    };
    static const Entry table[] = {
#define FAB_MODEL(id, p, name) { id, p, name },
#include "vb_models.inc"
#undef FAB_MODEL
    };
    for (size_t i = 0; i < sizeof(table) / sizeof(table[0]); i++)
        if (table[i].id == model_id && table[i].p == n_params)
            return table[i].get();
    return nullptr;
}

/* the debug build's index-check counters (vb_device.cuh FAB_CHECK): [0] failures, [1] largest site code. The helper
 * kernels of this file count into them directly, the model kernels through VbArgs::check. */
__device__ unsigned long long g_fab_check[2];
#ifdef FAB_BOUNDS_CHECK
#define FAB_CHECK_G(cond, code)                                                                                      \
    do                                                                                                               \
    {                                                                                                                \
        if (!(cond))                                                                                                 \
        {                                                                                                            \
            atomicAdd(&g_fab_check[0], 1ull);                                                                        \
            atomicMax(&g_fab_check[1], (unsigned long long)(code));                                                  \
        }                                                                                                            \
    } while (0)
#else
#define FAB_CHECK_G(cond, code)                                                                                      \
    do                                                                                                               \
    {                                                                                                                \
    } while (0)
#endif


/* ------------------------------------------------------------------------------------------------
 * small utility kernels
 * ---------------------------------------------------------------------------------------------- */
__global__ void gather_voxels_kernel(const float *__restrict__ full, size_t n_grid, int T,
    const int *__restrict__ index, int N, float *__restrict__ out)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)T * N;
    if (i >= total)
        return;
    const size_t t = i / N, v = i - t * N;
    out[i] = full[t * n_grid + (size_t)index[v]];
}

__global__ void scatter_voxels_kernel(const double *__restrict__ in, int n_rows, int N,
    const int *__restrict__ index, size_t n_grid, float *__restrict__ out_full)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n_rows * N;
    if (i >= total)
        return;
    const size_t r = i / N, v = i - r * N;
    out_full[r * n_grid + (size_t)index[v]] = (float)in[i];
}

__global__ void status_scan_kernel(const int *__restrict__ status, int N, int *out /* [count, first, code] */)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const int s = status[i];
    if (s != 0)
    {
        atomicAdd(&out[0], 1);
        atomicMin(&out[1], i);
    }
}

/* row-wise voxel permutation of a [rows][N] array: gather out[r][i] = in[r][perm[i]], or its inverse
 * scatter out[r][perm[i]] = in[r][i] (spatial VB keeps its state in hyper-plane-major voxel order) */
template <class T, bool GATHER>
__global__ void permute_rows_kernel(const T *__restrict__ in, T *__restrict__ out, const int *__restrict__ perm,
    int rows, int N)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * N)
        return;
    const size_t r = idx / N, i = idx - r * N;
    FAB_CHECK_G(perm[i] >= 0 && perm[i] < N, 201);
    if (GATHER)
        out[idx] = in[r * N + (size_t)perm[i]];
    else
        out[r * N + (size_t)perm[i]] = in[idx];
}
template <class T, bool GATHER>
static void permute_rows(const T *in, T *out, const int *perm, int rows, int N, cudaStream_t st)
{
    const size_t total = (size_t)rows * N;
    if (total == 0)
        return;
    permute_rows_kernel<T, GATHER><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(in, out, perm, rows, N);
    count_launch();
}
__global__ void iota_kernel(int *out, int N)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N)
        out[i] = i;
}
__global__ void rank_kernel(const int *order, int N, int *rank)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N)
    {
        FAB_CHECK_G(order[i] >= 0 && order[i] < N, 202);
        rank[order[i]] = i;
    }
}
/* z-slab mode: ghost voxels carry FABBER_VOX_GHOST so every update kernel skips them */
__global__ void mark_ghosts_kernel(const unsigned char *ghost, const int *order, int N, int *status_p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && ghost[order[i]])
        status_p[i] = FABBER_VOX_GHOST;
}
/* halo buffers: buf[k][i] <-> mean_p[k][rank[idx[i]]] */
template <bool PACK>
__global__ void halo_kernel(double *mean_p, const int *idx, const int *rank, int n, int P, int N, double *buf)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    FAB_CHECK_G(idx[i] >= 0 && idx[i] < N && rank[idx[i]] >= 0 && rank[idx[i]] < N, 203);
    const size_t pos = (size_t)rank[idx[i]];
    for (int k = 0; k < P; k++)
    {
        if (PACK)
            buf[(size_t)k * n + i] = mean_p[(size_t)k * N + pos];
        else
            mean_p[(size_t)k * N + pos] = buf[(size_t)k * n + i];
    }
}

/* neighbour table in permuted numbering: nnp[j][pos] = rank[nn[j][order[pos]]] */
__global__ void renumber_neighbours_kernel(const int *nn, const int *order, const int *rank, int N, int *nnp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N)
        return;
    const int v = order[i];
    for (int j = 0; j < 6; j++)
    {
        const int n = nn[(size_t)j * N + v];
        FAB_CHECK_G(v >= 0 && v < N && n >= -1 && n < N && n != v, 204);
        nnp[(size_t)j * N + i] = n >= 0 ? rank[n] : -1;
    }
}

/* SaveResults on the device: model-space mean / std / zstat / var per parameter, noise means and
 * standard deviations, the packed finalMVN rows, F and its history - all float32 (memory-bound map). */
struct SaveArgs
{
    int N, P, n_noise, ar, n_phis, n_alphas, f_len;
    char transform[FABBER_CUDA_MAX_PARAMS];
    const double *mean, *cov, *noise, *free_energy, *f_history;
    const int *iterations;
    fabber_cuda_vb_outputs out;
};
__device__ inline int save_tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
__device__ inline double save_to_model_var(char code, double v)
{
    switch (code)
    {
    case 'L':
        return exp(v);
    case 'I':
    case 'F':
        return v;
    default: /* transforms.cc:17-20 */
    {
        const double d = fab::to_model(code, sqrt(v)) - fab::to_model(code, 0.0);
        return d * d;
    }
    }
}
__global__ void __launch_bounds__(256) save_results_kernel(const __grid_constant__ SaveArgs a)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.N)
        return;
    const size_t N = (size_t)a.N;
    const int P = a.P, Nn = a.n_noise, NA = P + Nn;
    for (int i = 0; i < P; i++)
    {
        const double mean = fab::to_model(a.transform[i], a.mean[i * N + v]);
        const double var = save_to_model_var(a.transform[i], a.cov[save_tri(i, i) * N + v]);
        const double sd = sqrt(var);
        if (a.out.mean)
            a.out.mean[i * N + v] = (float)mean;
        if (a.out.std)
            a.out.std[i * N + v] = (float)sd;
        if (a.out.zstat)
            a.out.zstat[i * N + v] = (float)(mean / sd);
        if (a.out.var)
            a.out.var[i * N + v] = (float)var;
    }
    /* noise block of the result MVN (WhiteParams / Ar1cParams::OutputAsMVN) */
    double nmean[6], ncov[6][6];
    for (int i = 0; i < 6; i++)
    {
        nmean[i] = 0.0;
        for (int j = 0; j < 6; j++)
            ncov[i][j] = 0.0;
    }
    if (a.ar && a.n_phis == 2)
    {
        /* two echoes: the alphas (means, covariance = inverse of the packed precisions), then the two phis */
        const int nA = a.n_alphas;
        double w[4][8];
        for (int r = 0; r < nA; r++)
        {
            nmean[r] = a.noise[(size_t)(4 + r) * N + v];
            for (int c = 0; c < nA; c++)
            {
                w[r][c] = a.noise[(size_t)(4 + nA + save_tri(r, c)) * N + v];
                w[r][nA + c] = r == c ? 1.0 : 0.0;
            }
        }
        for (int k = 0; k < nA; k++) /* Gauss-Jordan with partial pivoting on [prec | I] */
        {
            int piv = k;
            for (int r = k + 1; r < nA; r++)
                if (fabs(w[r][k]) > fabs(w[piv][k]))
                    piv = r;
            for (int c = 0; c < 2 * nA; c++)
            {
                const double t = w[k][c];
                w[k][c] = w[piv][c];
                w[piv][c] = t;
            }
            const double inv = 1.0 / w[k][k];
            for (int c = 0; c < 2 * nA; c++)
                w[k][c] *= inv;
            for (int r = 0; r < nA; r++)
                if (r != k)
                {
                    const double f = w[r][k];
                    for (int c = 0; c < 2 * nA; c++)
                        w[r][c] -= f * w[k][c];
                }
        }
        for (int r = 0; r < nA; r++)
            for (int c = 0; c < nA; c++)
                ncov[r][c] = w[r][nA + c];
        for (int i = 0; i < 2; i++)
        {
            const double b = a.noise[(size_t)(2 * i) * N + v], c = a.noise[(size_t)(2 * i + 1) * N + v];
            nmean[nA + i] = b * c;
            ncov[nA + i][nA + i] = b * b * c;
        }
    }
    else if (a.ar)
    {
        const double b = a.noise[0 * N + v], c = a.noise[1 * N + v];
        const double p11 = a.noise[4 * N + v], p21 = a.noise[5 * N + v], p22 = a.noise[6 * N + v];
        const double det = p11 * p22 - p21 * p21;
        nmean[0] = a.noise[2 * N + v];
        nmean[1] = a.noise[3 * N + v];
        nmean[2] = b * c;
        ncov[0][0] = p22 / det;
        ncov[1][1] = p11 / det;
        ncov[0][1] = ncov[1][0] = -p21 / det;
        ncov[2][2] = b * b * c;
    }
    else
        for (int i = 0; i < a.n_phis; i++)
        {
            const double b = a.noise[(2 * i) * N + v], c = a.noise[(2 * i + 1) * N + v];
            nmean[i] = b * c;         /* GammaDist::CalcMean, dist_gamma.cc:21 */
            ncov[i][i] = b * b * c;   /* CalcVariance :25 */
        }
    for (int i = 0; i < Nn; i++)
    {
        if (a.out.noise_mean)
            a.out.noise_mean[i * N + v] = (float)nmean[i];
        if (a.out.noise_std)
            a.out.noise_std[i * N + v] = (float)sqrt(ncov[i][i]);
    }
    if (a.out.final_mvn)
    {
        int idx = 0;
        for (int r = 0; r < NA; r++)
            for (int c = 0; c <= r; c++, idx++)
            {
                double val = 0.0;
                if (r < P)
                    val = a.cov[save_tri(r, c) * N + v];
                else if (c >= P)
                    val = ncov[r - P][c - P];
                a.out.final_mvn[idx * N + v] = (float)val;
            }
        for (int i = 0; i < P; i++)
            a.out.final_mvn[(idx + i) * N + v] = (float)a.mean[i * N + v];
        for (int i = 0; i < Nn; i++)
            a.out.final_mvn[(idx + P + i) * N + v] = (float)nmean[i];
        a.out.final_mvn[(idx + NA) * N + v] = 1.0f;
    }
    if (a.out.free_energy && a.free_energy)
        a.out.free_energy[v] = (float)a.free_energy[v];
    if (a.out.f_history && a.f_history)
    {
        const int its = a.iterations ? a.iterations[v] : a.f_len;
        for (int r = 0; r < a.out.f_history_rows; r++)
            a.out.f_history[r * N + v]
                = (float)((r < its && r < a.f_len) ? a.f_history[r * N + v] : (a.free_energy ? a.free_energy[v] : 0.0));
    }
}
__global__ void max_int_kernel(const int *values, int n, int *result)
{
    int m = INT_MIN;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        m = max(m, values[i]);
    for (int o = 16; o > 0; o >>= 1)
        m = max(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0)
        atomicMax(result, m);
}

/* accuracy probe of the table-based exponential (vb_exp.cuh) against the CUDA library exp() */
__global__ void exp_probe_kernel(const double *x, double *fast, double *ref, int n)
{
    __shared__ double tab[EXP_TAB_DOUBLES];
    exp_table_stage(tab);
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    fast[i] = exp_fast(x[i], tab);
    ref[i] = exp(x[i]);
}

/* dependent-free DFMA loop: 8 independent accumulators per thread */
__global__ void fp64_peak_kernel(double *out, int iters, double x)
{
    double a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
           a7 = a0 + 7;
    const double y = x * 0.5;
#pragma unroll 1
    for (int i = 0; i < iters; i++)
    {
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            a0 = fma(a0, x, y);
            a1 = fma(a1, x, y);
            a2 = fma(a2, x, y);
            a3 = fma(a3, x, y);
            a4 = fma(a4, x, y);
            a5 = fma(a5, x, y);
            a6 = fma(a6, x, y);
            a7 = fma(a7, x, y);
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

/* ------------------------------------------------------------------------------------------------
 * problem -> kernel argument block
 * ---------------------------------------------------------------------------------------------- */
struct Staged
{
    double *design = nullptr;
    unsigned char *pattern = nullptr;
    void release(cudaStream_t s)
    {
        if (design)
            cudaFreeAsync(design, s);
        if (pattern)
            cudaFreeAsync(pattern, s);
        design = nullptr;
        pattern = nullptr;
    }
};

static int build_args(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf, cudaStream_t s,
    VbArgs &a, Staged &st, bool &general)
{
    if (!prob || !buf)
        return fail(FABBER_CUDA_ERR_INVALID, "null problem or buffers");
    keep_pool_memory();
    const int P = prob->model.n_params, T = prob->n_times, N = prob->n_voxels;
    if (P < 1 || P > FABBER_CUDA_MAX_PARAMS)
        return fail(FABBER_CUDA_ERR_INVALID, "n_params out of range");
    if (T < 1 || N < 0)
        return fail(FABBER_CUDA_ERR_INVALID, "bad n_times / n_voxels");
    if (prob->max_iterations <= 0)
        return fail(FABBER_CUDA_ERR_INVALID, "max_iterations must be positive"); /* convergence.cc:38 */
    if (prob->conv_type != FABBER_CONV_MAXITS && !(prob->fchange > 0))
        return fail(FABBER_CUDA_ERR_INVALID, "fchange must be positive"); /* convergence.cc:73,261 */
    if (prob->conv_type == FABBER_CONV_TRIALMODE && prob->max_trials <= 0)
        return fail(FABBER_CUDA_ERR_INVALID, "max_trials must be positive"); /* convergence.cc:148 */
    if (N > 0 && (!buf->data || !buf->mean || !buf->cov || !buf->noise || !buf->status))
        return fail(FABBER_CUDA_ERR_INVALID, "missing data / mean / cov / noise / status buffer");
    if ((buf->init_mean == nullptr) != (buf->init_cov == nullptr))
        return fail(FABBER_CUDA_ERR_INVALID, "init_mean and init_cov must be given together");

    memset(&a, 0, sizeof(a));
    a.N = N;
    a.T = T;
    a.data = buf->data;
    for (int i = 0; i < P; i++)
    {
        a.params[i] = prob->params[i];
        if (a.params[i].prior_type == '-')
            a.params[i].prior_type = 'N';
        const char tr = a.params[i].transform;
        if (tr != 'I' && tr != 'L' && tr != 'S' && tr != 'F' && tr != 'A')
            return fail(FABBER_CUDA_ERR_INVALID, "unknown transform code");
        a.image_prior[i] = buf->image_prior[i];
        if (a.params[i].prior_type == 'I' && N > 0 && !buf->image_prior[i])
            return fail(FABBER_CUDA_ERR_INVALID, "image prior without image");
    }
    a.exp_dt = prob->model.exp_dt;
    memcpy(a.model_consts, prob->model.consts, sizeof(a.model_consts));
    a.design_len = prob->model.design_len;
    {
        /* opt-in, default off: see recentre_loop (vb_voxelwise.cuh) */
        const char *bj = getenv("FABBER_B200_BASIS_JACOBIAN");
        a.basis_jacobian = (bj && bj[0] == '1') ? 1 : 0;
    }
    if (prob->model.id == FABBER_MODEL_POLY && prob->model.poly_degree + 1 != P)
        return fail(FABBER_CUDA_ERR_INVALID, "poly: n_params != degree + 1");
    if (prob->model.id == FABBER_MODEL_EXP && 2 * prob->model.exp_num != P)
        return fail(FABBER_CUDA_ERR_INVALID, "exp: n_params != 2 * num_exps");

    /* noise */
    const bool ar = prob->noise_type == FABBER_NOISE_AR1;
    a.n_phis = ar ? (prob->n_phis == 0 ? 1 : prob->n_phis) : prob->n_phis;
    if (a.n_phis < 1 || a.n_phis > FABBER_CUDA_MAX_PHIS)
        return fail(FABBER_CUDA_ERR_INVALID, "n_phis out of range");
    a.ar_n_alphas = 2;
    if (ar)
    {
        /* noisemodel_ar.cc:318-349 */
        if (a.n_phis > 2)
            return fail(FABBER_CUDA_ERR_INVALID, "AR noise model: num-echoes must be 1 or 2");
        if (prob->ar_cross_terms < FABBER_AR_CROSS_NONE || prob->ar_cross_terms > FABBER_AR_CROSS_DUAL)
            return fail(FABBER_CUDA_ERR_INVALID, "AR noise model: unknown ar1-cross-terms");
        if (a.n_phis == 1 && prob->ar_cross_terms != FABBER_AR_CROSS_NONE)
            return fail(FABBER_CUDA_ERR_INVALID, "AR noise model: ar1-cross-terms needs num-echoes=2");
        /* two echoes interleave TE1 TE2 ..; an odd series gives the reference alpha matrices of the wrong size
         * (a NEWMAT dimension exception), and a single pair has no lag at all */
        if (a.n_phis == 2 && (T % 2 != 0 || T < 4))
            return fail(FABBER_CUDA_ERR_INVALID, "AR noise model: num-echoes=2 needs an even number (>= 4) of time points");
        a.ar_n_alphas = 2 + prob->ar_cross_terms;
    }
    if (ar && prob->time_masked)
        for (int t = 0; t < T; t++)
            if (prob->time_masked[t]) /* noisemodel_ar.cc: masked time points not supported */
                return fail(FABBER_CUDA_ERR_INVALID, "AR noise model does not support masked time points");
    std::vector<unsigned char> pat(T, 0);
    general = false;
    int n_masked = 0;
    for (int i = 0; i < FABBER_CUDA_MAX_PHIS; i++)
        a.n_per_phi[i] = 0;
    for (int t = 0; t < T; t++)
    {
        int ph = (!ar && prob->phi_pattern) ? prob->phi_pattern[t] : 0;
        if (!ar && ph >= a.n_phis)
            return fail(FABBER_CUDA_ERR_INVALID, "phi_pattern entry >= n_phis");
        if (prob->time_masked && prob->time_masked[t])
        {
            pat[t] = FAB_PAT_MASKED;
            n_masked++;
            general = true;
        }
        else
        {
            pat[t] = (unsigned char)ph;
            a.n_per_phi[ph]++;
        }
    }
    if (!ar && a.n_phis > 1)
        general = true;
    a.n_unmasked = T - n_masked;
    for (int i = 0; i < FABBER_CUDA_MAX_PHIS; i++)
    {
        a.noise_prior_b[i] = prob->noise_prior_b[i];
        a.noise_prior_c[i] = prob->noise_prior_c[i];
        a.noise_post_b[i] = prob->noise_post_b[i];
        a.noise_post_c[i] = prob->noise_post_c[i];
    }
    a.locked_noise_stdev = prob->locked_noise_stdev;
    a.ar_alpha_prior_prec = prob->ar_alpha_prior_prec;
    a.nlls_lm = prob->nlls_lm;
    a.nlls_have_start = prob->nlls_have_start;
    for (int i = 0; i < FABBER_CUDA_MAX_PARAMS; i++)
        a.nlls_start[i] = prob->nlls_start[i];
    a.check = nullptr;
#ifdef FAB_BOUNDS_CHECK
    {
        void *sym = nullptr;
        if (cudaGetSymbolAddress(&sym, g_fab_check) == cudaSuccess)
            a.check = (unsigned long long *)sym;
    }
#endif
    a.conv_type = prob->conv_type;
    a.max_iterations = prob->max_iterations;
    a.max_trials = prob->max_trials;
    a.fchange = prob->fchange;
    a.need_f = (prob->need_f || prob->conv_type != FABBER_CONV_MAXITS) ? 1 : 0;
    a.f_history_len = buf->f_history ? prob->f_history_len : 0;
    a.init_mean = buf->init_mean;
    a.init_cov = buf->init_cov;
    a.init_noise = buf->init_noise;
    a.lock_centre = buf->lock_centre;
    a.mean = buf->mean;
    a.cov = buf->cov;
    a.noise = buf->noise;
    a.free_energy = buf->free_energy;
    a.f_history = a.f_history_len > 0 ? buf->f_history : nullptr;
    a.iterations = buf->iterations;
    a.status = buf->status;

    /* stage the per-run constants */
    cudaError_t e;
    const bool plugin_vec = prob->model.id == FABBER_MODEL_PLUGIN && prob->model.design && prob->model.design_len > 0;
    if (prob->model.id == FABBER_MODEL_LINEAR || plugin_vec)
    {
        if (!prob->model.design)
            return fail(FABBER_CUDA_ERR_INVALID, "linear model without design matrix");
        const size_t bytes = (plugin_vec ? (size_t)prob->model.design_len : (size_t)T * P) * sizeof(double);
        if ((e = cudaMallocAsync((void **)&st.design, bytes, s)) != cudaSuccess)
            return cuda_fail(e, "cudaMallocAsync(design)");
        if ((e = cudaMemcpyAsync(st.design, prob->model.design, bytes, cudaMemcpyHostToDevice, s)) != cudaSuccess)
            return cuda_fail(e, "cudaMemcpyAsync(design)");
        a.design = st.design;
    }
    if (general)
    {
        if ((e = cudaMallocAsync((void **)&st.pattern, (size_t)T, s)) != cudaSuccess)
            return cuda_fail(e, "cudaMallocAsync(pattern)");
        if ((e = cudaMemcpyAsync(st.pattern, pat.data(), (size_t)T, cudaMemcpyHostToDevice, s)) != cudaSuccess)
            return cuda_fail(e, "cudaMemcpyAsync(pattern)");
        /* pat is pageable host memory: the copy is staged before the call returns */
        a.pattern = st.pattern;
    }
    return FABBER_CUDA_OK;
}

/* device scratch of one spatial run; freed in reverse on any exit path */
struct Scratch
{
    cudaStream_t s;
    std::vector<void *> ptrs;
    explicit Scratch(cudaStream_t st)
        : s(st)
    {
    }
    template <class T> T *get(size_t n)
    {
        void *p = nullptr;
        if (cudaMallocAsync(&p, (n ? n : 1) * sizeof(T), s) != cudaSuccess)
            return nullptr;
        ptrs.push_back(p);
        return (T *)p;
    }
    ~Scratch()
    {
        for (size_t i = ptrs.size(); i-- > 0;)
            cudaFreeAsync(ptrs[i], s);
    }
};


} // namespace fab

using namespace fab;

extern "C" {

/* debug build (make checked): failures counted by the kernels' index checks since the last call, summed over
 * the devices, and the largest failing site code; the counters are reset. Returns 1 when the checks are compiled
 * in, 0 when they are not (out is zeroed), < 0 on a CUDA error. */
int fabber_cuda_check_report(unsigned long long *out)
{
    if (!out)
        return fail(FABBER_CUDA_ERR_INVALID, "null argument");
    out[0] = out[1] = 0;
#ifdef FAB_BOUNDS_CHECK
    int n = 0, cur = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || cudaGetDevice(&cur) != cudaSuccess)
        return fail(FABBER_CUDA_ERR_CUDA, "no device");
    for (int d = 0; d < n; d++)
    {
        unsigned long long w[2] = { 0, 0 }, zero[2] = { 0, 0 };
        cudaError_t e = cudaSetDevice(d);
        if (e == cudaSuccess)
            e = cudaDeviceSynchronize();
        if (e == cudaSuccess)
            e = cudaMemcpyFromSymbol(w, g_fab_check, sizeof(w));
        if (e == cudaSuccess)
            e = cudaMemcpyToSymbol(g_fab_check, zero, sizeof(zero));
        if (e != cudaSuccess)
        {
            cudaSetDevice(cur);
            return cuda_fail(e, "check_report");
        }
        out[0] += w[0];
        out[1] = w[1] > out[1] ? w[1] : out[1];
    }
    cudaSetDevice(cur);
    return 1;
#else
    return 0;
#endif
}

/* proves the check machinery is live: one deliberate failure with site code 999 (debug build; else a no-op) */
__global__ void check_selftest_kernel() { FAB_CHECK_G(threadIdx.x != 0, 999); }
int fabber_cuda_check_selftest(void)
{
    check_selftest_kernel<<<1, 32>>>();
    cudaError_t e = cudaDeviceSynchronize();
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "check_selftest");
}

int fabber_cuda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
        return 0;
    return n;
}

int fabber_cuda_set_device(int dev)
{
    cudaError_t e = cudaSetDevice(dev);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaSetDevice");
}

int fabber_cuda_get_device(void)
{
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess)
        return -1;
    return dev;
}

const char *fabber_cuda_last_error(void) { return g_last_error.c_str(); }

void *fabber_cuda_malloc(unsigned long long bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
    if (e != cudaSuccess)
    {
        cuda_fail(e, "cudaMalloc");
        return nullptr;
    }
    return p;
}
void fabber_cuda_free(void *dptr)
{
    if (dptr)
        cudaFree(dptr);
}
void *fabber_cuda_host_alloc(unsigned long long bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess)
    {
        cuda_fail(e, "cudaMallocHost");
        return nullptr;
    }
    return p;
}
void fabber_cuda_host_free(void *hptr)
{
    if (hptr)
        cudaFreeHost(hptr);
}
int fabber_cuda_memcpy_h2d(void *dst, const void *src, unsigned long long bytes, void *stream)
{
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaMemcpyAsync(h2d)");
}
int fabber_cuda_memcpy_d2h(void *dst, const void *src, unsigned long long bytes, void *stream)
{
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaMemcpyAsync(d2h)");
}
int fabber_cuda_memset(void *dst, int value, unsigned long long bytes, void *stream)
{
    cudaError_t e = cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaMemsetAsync");
}
int fabber_cuda_stream_sync(void *stream)
{
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaStreamSynchronize");
}

int fabber_cuda_gather_voxels(const float *full, unsigned long long n_grid, int n_times, const int *voxel_index,
    int n_voxels, float *out, void *stream)
{
    if (n_voxels <= 0 || n_times <= 0)
        return FABBER_CUDA_OK;
    const size_t total = (size_t)n_times * n_voxels;
    gather_voxels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        full, (size_t)n_grid, n_times, voxel_index, n_voxels, out);
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "gather_voxels");
}

int fabber_cuda_scatter_voxels(const double *in, int n_rows, int n_voxels, const int *voxel_index,
    unsigned long long n_grid, float *out_full, void *stream)
{
    cudaError_t e = cudaMemsetAsync(out_full, 0, (size_t)n_rows * n_grid * sizeof(float), (cudaStream_t)stream);
    if (e != cudaSuccess)
        return cuda_fail(e, "cudaMemsetAsync(scatter)");
    if (n_voxels <= 0 || n_rows <= 0)
        return FABBER_CUDA_OK;
    const size_t total = (size_t)n_rows * n_voxels;
    scatter_voxels_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        in, n_rows, n_voxels, voxel_index, (size_t)n_grid, out_full);
    count_launch();
    e = cudaGetLastError();
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "scatter_voxels");
}

int fabber_cuda_vb_voxelwise(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf, void *stream)
{
    return fabber_cuda_vb_voxelwise_range(prob, buf, 0, prob ? prob->n_voxels : 0, stream);
}

int fabber_cuda_vb_voxelwise_range(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf, int v_begin,
    int v_end, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (prob && (v_begin < 0 || v_end > prob->n_voxels || v_begin > v_end))
        return fail(FABBER_CUDA_ERR_INVALID, "voxel range outside [0, n_voxels]");
    VbArgs a;
    Staged st;
    bool general = false;
    int rc = build_args(prob, buf, s, a, st, general);
    if (rc != FABBER_CUDA_OK)
    {
        st.release(s);
        return rc;
    }
    const int P = prob->model.n_params;
    for (int i = 0; i < P; i++)
    {
        const char ty = a.params[i].prior_type;
        if (ty != 'N' && ty != 'I' && ty != 'A')
        {
            st.release(s);
            return fail(FABBER_CUDA_ERR_INVALID,
                "prior type is spatial or unknown: use fabber_cuda_vb_spatial (inference_vb.cc:334-358)");
        }
    }
    const ModelLaunchers *ml = find_model(prob->model, P);
    if (!ml)
    {
        st.release(s);
        return fail(FABBER_CUDA_ERR_INVALID, "no device Evaluate hook compiled for this model / parameter count");
    }
    VbLaunchFn fn = nullptr;
    if (prob->method == FABBER_METHOD_NLLS)
        fn = ml->nlls;
    else if (prob->noise_type == FABBER_NOISE_AR1)
        fn = a.n_phis == 2 ? ml->ar2 : ml->ar1;
    else if (prob->noise_type == FABBER_NOISE_WHITE)
    {
        const bool snap = prob->conv_type == FABBER_CONV_TRIALMODE || prob->conv_type == FABBER_CONV_FREDUCE;
        fn = general ? ml->white_general : (snap ? ml->white_fast_snap : ml->white_fast);
    }
    if (!fn)
    {
        st.release(s);
        return fail(FABBER_CUDA_ERR_INVALID, "noise model / inference technique not available for this model");
    }
    a.v_begin = v_begin;
    a.v_end = v_end;
    if (prob->method == FABBER_METHOD_NLLS && buf->noise && v_end > v_begin)
    {
        /* NLLS has no noise parameters: the two rows a white-noise caller allocates read as zeros, not as whatever
         * the allocation held */
        cudaError_t ez = cudaMemset2DAsync(buf->noise + v_begin, (size_t)prob->n_voxels * sizeof(double), 0,
            (size_t)(v_end - v_begin) * sizeof(double), 2, s);
        if (ez != cudaSuccess)
        {
            st.release(s);
            return cuda_fail(ez, "cudaMemset2DAsync(noise)");
        }
    }
    cudaError_t e = fn(a, s);
    st.release(s);
    if (e != cudaSuccess)
        return cuda_fail(e, "vb_voxelwise launch");
    return FABBER_CUDA_OK;
}

/* events and a second stream: what a host needs to overlap the upload of one block of voxels with the
 * calculation on the previous one (fabber_cuda_vb_voxelwise_range) */
void *fabber_cuda_stream_create(void)
{
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess)
        return nullptr;
    return s;
}
void fabber_cuda_stream_destroy(void *stream)
{
    if (stream)
        cudaStreamDestroy((cudaStream_t)stream);
}
void *fabber_cuda_event_create(void)
{
    cudaEvent_t e = nullptr;
    if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess)
        return nullptr;
    return e;
}
void fabber_cuda_event_destroy(void *event)
{
    if (event)
        cudaEventDestroy((cudaEvent_t)event);
}
int fabber_cuda_event_record(void *event, void *stream)
{
    cudaError_t e = cudaEventRecord((cudaEvent_t)event, (cudaStream_t)stream);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaEventRecord");
}
int fabber_cuda_stream_wait_event(void *stream, void *event)
{
    cudaError_t e = cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaStreamWaitEvent");
}
int fabber_cuda_memcpy2d_h2d(void *dst, unsigned long long dst_pitch, const void *src, unsigned long long src_pitch,
    unsigned long long width_bytes, unsigned long long rows, void *stream)
{
    cudaError_t e = cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyHostToDevice,
        (cudaStream_t)stream);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaMemcpy2DAsync");
}

int fabber_cuda_memcpy2d_d2h(void *dst, unsigned long long dst_pitch, const void *src, unsigned long long src_pitch,
    unsigned long long width_bytes, unsigned long long rows, void *stream)
{
    cudaError_t e = cudaMemcpy2DAsync(dst, dst_pitch, src, src_pitch, width_bytes, rows, cudaMemcpyDeviceToHost,
        (cudaStream_t)stream);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaMemcpy2DAsync(d2h)");
}
int fabber_cuda_event_sync(void *event)
{
    cudaError_t e = cudaEventSynchronize((cudaEvent_t)event);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "cudaEventSynchronize");
}
int fabber_cuda_host_is_pinned(const void *host_ptr)
{
    cudaPointerAttributes at;
    if (!host_ptr || cudaPointerGetAttributes(&at, host_ptr) != cudaSuccess)
    {
        cudaGetLastError(); /* older runtimes report unregistered memory as an error: clear it */
        return 0;
    }
    return at.type == cudaMemoryTypeHost ? 1 : 0;
}

int fabber_cuda_vb_spatial(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf, void *stream)
{
    return fabber_cuda_vb_spatial_slab(prob, buf, nullptr, stream);
}

/* FABBER_B200_SLAB_EXCHANGE=side moves the halo exchange to the side stream (under sp_noise); the default keeps
 * it on the main stream, in front of sp_noise, until the side-stream variant has been measured over NCCL */
static bool slab_exchange_on_main()
{
    const char *e = getenv("FABBER_B200_SLAB_EXCHANGE");
    return !(e && strcmp(e, "side") == 0);
}

} // extern "C"

namespace fab
{
/* own voxels are the positions [own0, own1) of the caller's list: everything else is a ghost */
__global__ void mark_ghost_range_kernel(const int *order, int N, int own0, int own1, int *status_p)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N && (order[i] < own0 || order[i] >= own1))
        status_p[i] = FABBER_VOX_GHOST;
}
/* one thread: wait until both ghost planes of this slab hold the neighbours' values of iteration it-1 (the aK
 * sums of iteration it read them). A kernel of its own, one thread: a full-GPU kernel spinning in every CTA would
 * keep slabs that share the GPU from ever running the sweep that raises the flag. */
__global__ void slab_wait_kernel(const unsigned long long *flags, int wait_fwd, int wait_hi, int it, int *error)
{
    if (wait_fwd)
        slab_wait(flags + SLAB_FLAG_FWD, (unsigned long long)it * SLAB_IT_STRIDE, error);
    if (wait_hi)
        slab_wait(flags + SLAB_FLAG_HI, (unsigned long long)it, error);
}
/* links between two neighbouring slabs, in plane-major positions: for every voxel of the caller's list of the
 * LOWER slab that is also in the UPPER slab's list (global ids [g0, g1)), lower position <-> upper position */
__global__ void slab_up_pos_kernel(const int *rank_lo, int v0_lo, const int *rank_hi, int v0_hi, int g0, int g1,
    int own1_lo /* global */, int *up_pos, int n_lo, int n_hi)
{
    const int g = g0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= g1 || g >= own1_lo)
        return; /* only the lower slab's OWN voxels forward upwards */
    FAB_CHECK_G(g - v0_lo >= 0 && g - v0_lo < n_lo && g - v0_hi >= 0 && g - v0_hi < n_hi, 205);
    FAB_CHECK_G(rank_lo[g - v0_lo] >= 0 && rank_lo[g - v0_lo] < n_lo && rank_hi[g - v0_hi] >= 0 && rank_hi[g - v0_hi] < n_hi, 206);
    up_pos[rank_lo[g - v0_lo]] = rank_hi[g - v0_hi];
}
__global__ void slab_dn_list_kernel(const int *rank_hi, int v0_hi, const int *rank_lo, int v0_lo, int g0, int n,
    int *dn_src, int *dn_dst, int n_hi, int n_lo)
{
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n)
        return;
    const int g = g0 + j; /* own voxels of the upper slab that the lower slab holds as its upper ghosts */
    FAB_CHECK_G(g - v0_hi >= 0 && g - v0_hi < n_hi && g - v0_lo >= 0 && g - v0_lo < n_lo, 207);
    dn_src[j] = rank_hi[g - v0_hi];
    dn_dst[j] = rank_lo[g - v0_lo];
}

/* One spatial VB run on one device: the state fabber_cuda_vb_spatial_slab used to keep in locals, so that
 * several of them (one per z-slab / GPU) can be stepped in lock-step by one host thread. */
struct SpRun
{
    int device = 0;
    cudaStream_t st = nullptr;
    const fabber_cuda_vb_problem *prob = nullptr;
    fabber_cuda_vb_buffers buf;
    SpArgs sp;
    Staged staged;
    Scratch *sc = nullptr;
    const ModelLaunchers *ml = nullptr;
    int P = 0, N = 0, NT = 0, H = 0, max_it = 0;
    bool any_spatial = false, any_coupled = false;
    int *order = nullptr, *rank = nullptr, *status_p = nullptr, *status_prev = nullptr, *its_p = nullptr;
    float *y_p = nullptr;
    double *mean_p = nullptr, *cov_p = nullptr, *noise_p = nullptr, *F_p = nullptr, *hist_p = nullptr;
    unsigned gridN = 0;

    ~SpRun()
    {
        if (sc)
        {
            cudaSetDevice(device);
            staged.release(st);
            delete sc;
        }
    }
    template <class T> T *get(size_t n) { return sc->get<T>(n); }

    /* validation, scratch, neighbours, hyper-plane renumbering, inputs into plane-major order, sp_setup.
     * Returns FABBER_CUDA_OK with N == 0 handled by the caller. */
    int prepare(const fabber_cuda_vb_problem *prob_, const fabber_cuda_vb_buffers *buf_, cudaStream_t stream,
        int n_global)
    {
        prob = prob_;
        buf = *buf_;
        st = stream;
        cudaGetDevice(&device);
        memset(&sp, 0, sizeof(sp));
        bool general = false;
        int rc = build_args(prob, &buf, st, sp.v, staged, general);
        sc = new Scratch(st);
        if (rc != FABBER_CUDA_OK)
            return rc;
        if (prob->noise_type != FABBER_NOISE_WHITE || general)
            return fail(FABBER_CUDA_ERR_INVALID,
                "spatial VB kernels support white noise with one phi and no masked time points");
        P = prob->model.n_params;
        N = prob->n_voxels;
        NT = P * (P + 1) / 2;
        for (int i = 0; i < P; i++)
        {
            const char ty = sp.v.params[i].prior_type;
            if (ty == 'M' || ty == 'm')
                any_coupled = true;
            if (ty == 'M' || ty == 'm' || ty == 'P' || ty == 'p')
                any_spatial = true;
            else if (ty != 'N' && ty != 'I' && ty != 'A')
                return fail(FABBER_CUDA_ERR_INVALID, "unknown prior type");
        }
        if (prob->spatial_dims < 0 || prob->spatial_dims > 3)
            return fail(FABBER_CUDA_ERR_INVALID, "spatial-dims must be 0, 1, 2 or 3"); /* priors.cc:191-194 */
        ml = find_model(prob->model, P);
        if (!ml)
            return fail(FABBER_CUDA_ERR_INVALID, "no device Evaluate hook compiled for this model / parameter count");
        if (N == 0)
            return FABBER_CUDA_OK;
        if (!buf.coords)
            return fail(FABBER_CUDA_ERR_INVALID, "spatial VB needs voxel coordinates");
        const int nx = prob->nx, ny = prob->ny, nz = prob->nz;
        if (nx <= 0 || ny <= 0 || nz <= 0)
            return fail(FABBER_CUDA_ERR_INVALID, "spatial VB needs the bounding grid nx, ny, nz");
        const size_t n_grid = (size_t)nx * ny * nz;
        const int n_planes = nx + ny + nz;
        max_it = prob->max_iterations;

        int *grid2vox = get<int>(n_grid), *nn_idx = get<int>((size_t)6 * N), *plane_of = get<int>(N);
        int *hist = get<int>(n_planes + 1);
        order = get<int>(N);
        rank = get<int>(N);
        int *bad = get<int>(1), *iota = get<int>(N), *plane_sorted = get<int>(N), *nnp = get<int>((size_t)6 * N);
        int *plane_starts = get<int>(n_planes + 2);
        sp.centre = get<double>((size_t)P * N);
        sp.stats = get<double>((size_t)(NT + P + 1) * N);
        sp.m0 = get<double>((size_t)P * N);
        sp.L0 = get<double>((size_t)P * N);
        sp.rhs = get<double>((size_t)P * N);
        sp.logdet = get<double>(N);
        sp.aK = get<double>(P);
        sp.ak_hist = get<double>((size_t)(max_it + 1) * P);
        sp.ak_partial = get<double>((size_t)SP_AK_BLOCKS * 2 * P);
        sp.fprior_last = get<double>(1);
        sp.ak_sums = get<double>(2 * P);
        sp.sweep_barrier = get<unsigned>(1);
        sp.n_global = n_global > 0 ? n_global : N;
        sp.ak_phase = 0;
        /* the run's inputs and outputs in hyper-plane-major voxel order (see below) */
        H = sp.v.f_history_len;
        y_p = get<float>((size_t)prob->n_times * N);
        mean_p = get<double>((size_t)P * N);
        cov_p = get<double>((size_t)NT * N);
        noise_p = get<double>((size_t)2 * N);
        F_p = get<double>(N);
        hist_p = get<double>((size_t)H * N);
        its_p = get<int>(N);
        status_p = get<int>(N);
        /* allow-bad-voxels: the status words as they stood when the iteration began (Vb::IgnoreVoxel, see
         * nbr_alive in vb_spatial.cuh); without it any failure ends the run and the live array serves */
        status_prev = prob->allow_bad_voxels ? get<int>(N) : status_p;
        if (!grid2vox || !nn_idx || !plane_of || !order || !hist || !rank || !bad || !iota || !plane_sorted || !nnp
            || !plane_starts || !sp.centre || !sp.stats || !sp.m0 || !sp.L0 || !sp.rhs || !sp.logdet || !sp.aK
            || !sp.ak_hist || !sp.ak_partial || !sp.fprior_last || !sp.ak_sums || !sp.sweep_barrier || !y_p || !mean_p || !cov_p || !noise_p
            || !F_p || !hist_p || !its_p || !status_p || !status_prev)
            return fail(FABBER_CUDA_ERR_CUDA, "out of device memory for the spatial VB state");
        sp.status_prev = status_prev;
        sp.ignore_bad = prob->allow_bad_voxels ? 1 : 0;
        sp.ak_blocks = SP_AK_BLOCKS;
        sp.spatial_dims = prob->spatial_dims;
        sp.update_first_iter = prob->update_first_iter;
        sp.any_coupled = any_coupled ? 1 : 0;
        sp.q1 = prob->spatial_q1;
        sp.q2 = prob->spatial_q2;
        sp.speed = prob->spatial_speed;

        /* ---- neighbours (Vb::CalcNeighbours) and the hyper-plane-major renumbering ----------------------
         * The ordered sweep walks planes x+y+z = h. In the caller's x-fastest order the voxels of a plane are
         * strided through memory and every access of the sweep is an uncoalesced 8-byte gather (measured:
         * 64 % of the step). So the whole spatial run works on a renumbered volume: voxels sorted by plane,
         * original order kept inside a plane (stable radix sort), which makes every per-voxel array access of
         * every kernel coalesced and keeps the -x/-y/-z neighbours of consecutive voxels consecutive too. The
         * time-series is permuted once on the way in, the results once on the way out. */
        cudaMemsetAsync(grid2vox, 0xff, n_grid * sizeof(int), st);
        cudaMemsetAsync(hist, 0, (n_planes + 1) * sizeof(int), st);
        cudaMemsetAsync(bad, 0, sizeof(int), st);
        cudaMemsetAsync(sp.fprior_last, 0, sizeof(double), st);
        {
            std::vector<double> ak0(P, 1e-8); /* SpatialPrior::m_aK initial value, priors.cc:185 */
            cudaMemcpyAsync(sp.aK, ak0.data(), P * sizeof(double), cudaMemcpyHostToDevice, st);
            cudaStreamSynchronize(st);
        }
        gridN = (unsigned)((N + 255) / 256);
        sp_grid_kernel<<<gridN, 256, 0, st>>>(buf.coords, N, nx, ny, nz, grid2vox, bad);
        count_launch();
        {
            /* nothing below may index by plane or grid cell before the coordinates are known to be inside the
             * grid: a caller's bad coords must come back as ERR_INVALID, not as an illegal address */
            int h_bad0 = 0;
            cudaMemcpyAsync(&h_bad0, bad, sizeof(int), cudaMemcpyDeviceToHost, st);
            cudaError_t be = cudaStreamSynchronize(st);
            if (be != cudaSuccess)
                return cuda_fail(be, "spatial coordinate check");
            if (h_bad0 == 1)
                return fail(FABBER_CUDA_ERR_INVALID, "voxel coordinates outside the nx, ny, nz grid");
            if (h_bad0 == 2)
                return fail(FABBER_CUDA_ERR_INVALID,
                    "coordinates must be in increasing order (x fastest, then y, then z), inference_vb.cc:769-793");
        }
        sp_neighbour_kernel<<<gridN, 256, 0, st>>>(buf.coords, N, nx, ny, nz, grid2vox, prob->spatial_dims, nn_idx,
            plane_of, hist);
        count_launch();
        iota_kernel<<<gridN, 256, 0, st>>>(iota, N);
        count_launch();
        {
            int bits = 1;
            while ((1 << bits) <= n_planes)
                bits++;
            size_t tmp_bytes = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, plane_of, plane_sorted, iota, order, N, 0, bits, st);
            void *tmp = get<char>(tmp_bytes);
            if (!tmp)
                return fail(FABBER_CUDA_ERR_CUDA, "out of device memory for the spatial VB state");
            cudaError_t se = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, plane_of, plane_sorted, iota, order, N, 0,
                bits, st);
            if (se != cudaSuccess)
                return cuda_fail(se, "hyper-plane sort");
            count_launch();
        }
        rank_kernel<<<gridN, 256, 0, st>>>(order, N, rank);
        count_launch();
        renumber_neighbours_kernel<<<gridN, 256, 0, st>>>(nn_idx, order, rank, N, nnp);
        count_launch();
        std::vector<int> h_hist(n_planes + 1), h_begin(n_planes + 2);
        cudaMemcpyAsync(h_hist.data(), hist, (n_planes + 1) * sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess)
            return cuda_fail(e, "spatial neighbour set-up");
        int acc = 0;
        for (int h = 0; h <= n_planes; h++)
        {
            h_begin[h] = acc;
            acc += h_hist[h];
        }
        h_begin[n_planes + 1] = acc;
        cudaMemcpyAsync(plane_starts, h_begin.data(), (n_planes + 2) * sizeof(int), cudaMemcpyHostToDevice, st);
        sp.plane_starts = plane_starts;
        sp.n_planes = n_planes + 1;
        sp.nn_idx = nnp;
        sp.order = order;
        sp.last_pos = rank + (N - 1);

        /* inputs into plane-major order */
        permute_rows<float, true>(buf.data, y_p, order, prob->n_times, N, st);
        sp.v.data = y_p;
        for (int i = 0; i < P; i++)
            if (sp.v.image_prior[i])
            {
                double *img = get<double>(N);
                if (!img)
                    return fail(FABBER_CUDA_ERR_CUDA, "out of device memory for the spatial VB state");
                permute_rows<double, true>(sp.v.image_prior[i], img, order, 1, N, st);
                sp.v.image_prior[i] = img;
            }
        if (sp.v.init_mean)
        {
            double *im = get<double>((size_t)P * N), *ic = get<double>((size_t)NT * N);
            if (!im || !ic)
                return fail(FABBER_CUDA_ERR_CUDA, "out of device memory for the spatial VB state");
            permute_rows<double, true>(sp.v.init_mean, im, order, P, N, st);
            permute_rows<double, true>(sp.v.init_cov, ic, order, NT, N, st);
            sp.v.init_mean = im;
            sp.v.init_cov = ic;
        }
        if (sp.v.lock_centre)
        {
            double *lc = get<double>((size_t)P * N);
            if (!lc)
                return fail(FABBER_CUDA_ERR_CUDA, "out of device memory for the spatial VB state");
            permute_rows<double, true>(sp.v.lock_centre, lc, order, P, N, st);
            sp.v.lock_centre = lc;
        }
        if (sp.v.init_noise)
        {
            double *in = get<double>((size_t)2 * N);
            if (!in)
                return fail(FABBER_CUDA_ERR_CUDA, "out of device memory for the spatial VB state");
            permute_rows<double, true>(sp.v.init_noise, in, order, 2, N, st);
            sp.v.init_noise = in;
        }
        sp.v.mean = mean_p;
        sp.v.cov = cov_p;
        sp.v.noise = noise_p;
        sp.v.free_energy = F_p;
        sp.v.f_history = H > 0 ? hist_p : nullptr;
        sp.v.iterations = its_p;
        sp.v.status = status_p;
        cudaStreamSynchronize(st); /* h_begin is pageable host memory */
        sp.it = 0;
        cudaError_t le = ml->sp_setup(sp, st);
        if (le != cudaSuccess)
            return cuda_fail(le, "spatial VB sp_setup");
        return FABBER_CUDA_OK;
    }

    void snapshot_status()
    {
        if (status_prev != status_p)
            cudaMemcpyAsync(status_prev, status_p, (size_t)N * sizeof(int), cudaMemcpyDeviceToDevice, st);
    }

    /* results back into the caller's voxel order; synchronises */
    bool finish_queued = false;
    int finish()
    {
        if (!finish_queued)
            finish_async();
        cudaError_t e = cudaStreamSynchronize(st);
        if (e != cudaSuccess)
            return cuda_fail(e, "spatial VB");
        return FABBER_CUDA_OK;
    }
    void finish_async()
    {
        finish_queued = true;
        permute_rows<double, false>(mean_p, buf.mean, order, P, N, st);
        permute_rows<double, false>(cov_p, buf.cov, order, NT, N, st);
        permute_rows<double, false>(noise_p, buf.noise, order, 2, N, st);
        permute_rows<int, false>(status_p, buf.status, order, 1, N, st);
        if (buf.free_energy)
            permute_rows<double, false>(F_p, buf.free_energy, order, 1, N, st);
        if (buf.iterations)
            permute_rows<int, false>(its_p, buf.iterations, order, 1, N, st);
        if (H > 0)
            permute_rows<double, false>(hist_p, buf.f_history, order, H, N, st);
        if (buf.spatial_ak)
            cudaMemcpyAsync(buf.spatial_ak, sp.ak_hist, (size_t)(max_it + 1) * P * sizeof(double),
                cudaMemcpyDeviceToHost, st);
    }
};
} // namespace fab

extern "C" {

int fabber_cuda_vb_spatial_slab(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf,
    const fabber_cuda_slab *slab, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!prob || !buf)
        return fail(FABBER_CUDA_ERR_INVALID, "null problem or buffers");
    if (slab
        && (!slab->ghost || !slab->allreduce_sum || !slab->exchange || !slab->forward
            || slab->n_global_voxels < prob->n_voxels || slab->n_blocks < 1 || slab->block_planes < 1 || slab->world < 1
            || slab->rank < 0 || slab->rank >= slab->world || !slab->fwd_send_start || !slab->fwd_recv_start))
        return fail(FABBER_CUDA_ERR_INVALID, "incomplete slab description");
    SpRun R;
    int rc = R.prepare(prob, buf, st, slab ? slab->n_global_voxels : 0);
    if (rc != FABBER_CUDA_OK || R.N == 0)
        return rc;
    SpArgs &sp = R.sp;
    const ModelLaunchers *ml = R.ml;
    const int P = R.P, N = R.N, max_it = R.max_it;
    const bool any_spatial = R.any_spatial, any_coupled = R.any_coupled;
    double *mean_p = R.mean_p;
    int *rank = R.rank;
    double *halo_send_lo = nullptr, *halo_send_hi = nullptr, *halo_recv_lo = nullptr, *halo_recv_hi = nullptr;
    double *fwd_send_buf = nullptr, *fwd_recv_buf = nullptr;
    if (slab)
    {
        halo_send_lo = R.get<double>((size_t)P * slab->n_send_lo);
        halo_send_hi = R.get<double>((size_t)P * slab->n_send_hi);
        halo_recv_lo = R.get<double>((size_t)P * slab->n_recv_lo);
        halo_recv_hi = R.get<double>((size_t)P * slab->n_recv_hi);
        size_t max_fwd = 1;
        for (int b = 0; b < slab->n_blocks; b++)
        {
            max_fwd = std::max<size_t>(max_fwd, slab->fwd_send_start[b + 1] - slab->fwd_send_start[b]);
            max_fwd = std::max<size_t>(max_fwd, slab->fwd_recv_start[b + 1] - slab->fwd_recv_start[b]);
        }
        fwd_send_buf = R.get<double>((size_t)P * max_fwd);
        fwd_recv_buf = R.get<double>((size_t)P * max_fwd);
        if (!halo_send_lo || !halo_send_hi || !halo_recv_lo || !halo_recv_hi || !fwd_send_buf || !fwd_recv_buf)
            return fail(FABBER_CUDA_ERR_CUDA, "out of device memory for the spatial VB state");
        mark_ghosts_kernel<<<R.gridN, 256, 0, st>>>(slab->ghost, R.order, N, R.status_p);
        count_launch();
    }

    /* ---- the iteration-major loop --------------------------------------------------------------- */
#define FAB_SP_LAUNCH(fn)                                \
    do                                                   \
    {                                                    \
        cudaError_t le = ml->fn(sp, st);                 \
        if (le != cudaSuccess)                           \
            return cuda_fail(le, "spatial VB " #fn);     \
    } while (0)
    /* z-slab mode: the aK sums of iteration it+1 depend on the means and covariances as the sweep and the halo
     * exchange of iteration it leave them - sp_noise does not touch either. The slabs finish their sweeps
     * staggered (slab r starts r * nz_local hyper-planes after slab 0), so an all-reduce in front of the next
     * iteration would make every slab wait for the LAST slab's sp_noise. Instead the local sums and the
     * all-reduce are issued on a side stream right after the halo exchange and run under this slab's own
     * sp_noise; the next iteration only waits for the event. (Only difference: with allow_bad_voxels a voxel
     * that fails in sp_noise of iteration it still contributes to the aK of iteration it+1.) */
    struct SideStream
    {
        cudaStream_t s = nullptr;
        cudaEvent_t swept = nullptr, ak = nullptr;
        ~SideStream()
        {
            if (swept)
                cudaEventDestroy(swept);
            if (ak)
                cudaEventDestroy(ak);
            if (s)
                cudaStreamDestroy(s);
        }
    } side;
    if (slab && any_spatial)
    {
        /* FABBER_B200_SLAB_PRIORITY=1: highest stream priority, so that the side stream's small kernels are
         * dispatched between sp_noise's thread blocks instead of behind the last of them (see DESIGN.md section
         * 7: without it the 8-GPU run measured no gain from the side stream). Opt-in until measured. */
        const char *pe = getenv("FABBER_B200_SLAB_PRIORITY");
        int least = 0, greatest = 0;
        cudaError_t se = cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (se == cudaSuccess)
            se = cudaStreamCreateWithPriority(&side.s, cudaStreamNonBlocking, (pe && pe[0] == '1') ? greatest : least);
        if (se == cudaSuccess)
            se = cudaEventCreateWithFlags(&side.swept, cudaEventDisableTiming);
        if (se == cudaSuccess)
            se = cudaEventCreateWithFlags(&side.ak, cudaEventDisableTiming);
        if (se != cudaSuccess)
            return cuda_fail(se, "spatial VB side stream");
    }
    bool ak_ahead = false; /* this iteration's all-reduced sums are in flight on (or done by) the side stream */
    for (int it = 0; it < max_it; it++)
    {
        sp.it = it;
        /* SpatialPrior::ApplyToMVN at v == 1: aK is refreshed from the current posteriors unless this is
         * the first iteration (priors.cc:350-358) */
        sp.ak_update = (any_spatial && (it > 0 || prob->update_first_iter)) ? 1 : 0;
        if (slab && sp.ak_update)
        {
            /* local sums -> all-reduce over the slabs -> aK (identical on every rank) */
            if (ak_ahead)
                cudaStreamWaitEvent(st, side.ak, 0);
            else
            {
                FAB_SP_LAUNCH(sp_ak_partial);
                sp.ak_phase = 1;
                FAB_SP_LAUNCH(sp_ak_final);
                if (slab->allreduce_sum(slab->user, sp.ak_sums, 2 * P, st) != 0)
                    return fail(FABBER_CUDA_ERR_CUDA, "slab all-reduce callback failed");
            }
            sp.ak_phase = 2;
        }
        else if (sp.ak_update)
            FAB_SP_LAUNCH(sp_ak_partial);
        FAB_SP_LAUNCH(sp_ak_final);
        sp.ak_phase = 0;
        R.snapshot_status();
        FAB_SP_LAUNCH(sp_theta);
        sp.plane_first = 0;
        sp.plane_last = sp.n_planes;
        if (any_coupled && !slab)
            FAB_SP_LAUNCH(sp_sweep);
        if (any_coupled && slab)
        {
            /* pipelined exact sweep over the slabs: block b = global planes [b B, (b+1) B) */
            const int B = slab->block_planes;
            for (int step = 0; step < slab->n_blocks + slab->world - 1; step++)
            {
                const int b = step - slab->rank;
                int n_send = 0, n_recv = 0;
                if (b >= 0 && b < slab->n_blocks)
                {
                    sp.plane_first = std::min(std::max(b * B - slab->plane_offset, 0), sp.n_planes);
                    sp.plane_last = std::min(std::max((b + 1) * B - slab->plane_offset, 0), sp.n_planes);
                    if (sp.plane_last > sp.plane_first)
                        FAB_SP_LAUNCH(sp_sweep);
                    n_send = slab->fwd_send_start[b + 1] - slab->fwd_send_start[b];
                    if (n_send > 0)
                    {
                        halo_kernel<true><<<(n_send + 255) / 256, 256, 0, st>>>(
                            mean_p, slab->fwd_send + slab->fwd_send_start[b], rank, n_send, P, N, fwd_send_buf);
                        count_launch();
                    }
                }
                const int b_in = b + 1; /* the block this rank sweeps next: its lower ghosts arrive now */
                if (b_in >= 0 && b_in < slab->n_blocks)
                    n_recv = slab->fwd_recv_start[b_in + 1] - slab->fwd_recv_start[b_in];
                if (slab->forward(slab->user, step, fwd_send_buf, n_send, fwd_recv_buf, n_recv, st) != 0)
                    return fail(FABBER_CUDA_ERR_CUDA, "slab forward callback failed");
                if (n_recv > 0)
                {
                    halo_kernel<false><<<(n_recv + 255) / 256, 256, 0, st>>>(
                        mean_p, slab->fwd_recv + slab->fwd_recv_start[b_in], rank, n_recv, P, N, fwd_recv_buf);
                    count_launch();
                }
            }
        }
        /* z-slab mode, after the sweep: halo exchange of the posterior means (own boundary planes out, ghost
         * planes in), then - if another iteration follows - the local aK sums of that iteration and their
         * all-reduce. Nothing in sp_noise reads a ghost or a neighbour, so when another iteration follows all of
         * this goes on the side stream and runs under this slab's sp_noise: the main stream does not even wait
         * for the upper neighbour's sweep to end (its bottom plane is this slab's upper ghost). */
        ak_ahead = false;
        if (slab)
        {
            const bool ahead = any_spatial && it + 1 < max_it && !slab_exchange_on_main() && !prob->allow_bad_voxels;
            cudaStream_t xs = ahead ? side.s : st;
            if (ahead)
            {
                cudaEventRecord(side.swept, st);
                cudaStreamWaitEvent(side.s, side.swept, 0);
            }
            if (slab->n_send_lo > 0)
                halo_kernel<true><<<(slab->n_send_lo + 255) / 256, 256, 0, xs>>>(
                    mean_p, slab->send_lo, rank, slab->n_send_lo, P, N, halo_send_lo);
            if (slab->n_send_hi > 0)
                halo_kernel<true><<<(slab->n_send_hi + 255) / 256, 256, 0, xs>>>(
                    mean_p, slab->send_hi, rank, slab->n_send_hi, P, N, halo_send_hi);
            count_launch();
            if (slab->exchange(slab->user, halo_send_lo, slab->n_send_lo, halo_send_hi, slab->n_send_hi, halo_recv_lo,
                    slab->n_recv_lo, halo_recv_hi, slab->n_recv_hi, xs)
                != 0)
                return fail(FABBER_CUDA_ERR_CUDA, "slab halo-exchange callback failed");
            if (slab->n_recv_lo > 0)
                halo_kernel<false><<<(slab->n_recv_lo + 255) / 256, 256, 0, xs>>>(
                    mean_p, slab->recv_lo, rank, slab->n_recv_lo, P, N, halo_recv_lo);
            if (slab->n_recv_hi > 0)
                halo_kernel<false><<<(slab->n_recv_hi + 255) / 256, 256, 0, xs>>>(
                    mean_p, slab->recv_hi, rank, slab->n_recv_hi, P, N, halo_recv_hi);
            count_launch();
            /* with allow_bad_voxels the sums must see the failures sp_noise of THIS iteration records
             * (CalculateaK skips ignored voxels, priors.cc:238-241): then they are formed in order, at the top of
             * the next iteration, not under sp_noise */
            if (any_spatial && it + 1 < max_it && !prob->allow_bad_voxels)
            {
                if (!ahead)
                {
                    cudaEventRecord(side.swept, st);
                    cudaStreamWaitEvent(side.s, side.swept, 0);
                }
                SpArgs nxt = sp;
                nxt.it = it + 1;
                nxt.ak_update = 1;
                nxt.ak_phase = 1;
                cudaError_t le = ml->sp_ak_partial(nxt, side.s);
                if (le == cudaSuccess)
                    le = ml->sp_ak_final(nxt, side.s);
                if (le != cudaSuccess)
                    return cuda_fail(le, "spatial VB aK sums (side stream)");
                if (slab->allreduce_sum(slab->user, sp.ak_sums, 2 * P, side.s) != 0)
                    return fail(FABBER_CUDA_ERR_CUDA, "slab all-reduce callback failed");
                cudaEventRecord(side.ak, side.s);
                ak_ahead = true;
            }
        }
        FAB_SP_LAUNCH(sp_noise);
    }
    sp.it = max_it;
    sp.ak_update = 0;
    FAB_SP_LAUNCH(sp_ak_final);
#undef FAB_SP_LAUNCH
    return R.finish();
}

/* ---- device-driven z-slabs: ONE process, one slab per part, kernels coupled through peer memory ---------- */
int fabber_cuda_vb_spatial_multi(const fabber_cuda_vb_problem *prob, int n_parts, const fabber_cuda_slab_part *parts)
{
    if (!prob || !parts || n_parts < 1 || n_parts > SLAB_MAX_WORLD)
        return fail(FABBER_CUDA_ERR_INVALID, "spatial multi: 1..16 parts");
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    struct Restore
    {
        int dev;
        ~Restore() { cudaSetDevice(dev); }
    } restore = { prev_dev };
    if (n_parts == 1)
    {
        cudaSetDevice(parts[0].device);
        fabber_cuda_vb_problem p1 = *prob;
        p1.n_voxels = parts[0].v1 - parts[0].v0;
        return fabber_cuda_vb_spatial(&p1, &parts[0].buf, nullptr);
    }
    const int W = n_parts;
    const int n_global = prob->n_voxels;
    for (int r = 0; r < W; r++)
    {
        const fabber_cuda_slab_part &pt = parts[r];
        if (pt.v0 < 0 || pt.v1 > n_global || pt.own0 < pt.v0 || pt.own1 > pt.v1 || pt.own0 >= pt.own1
            || pt.own_z0 >= pt.own_z1)
            return fail(FABBER_CUDA_ERR_INVALID, "spatial multi: bad part ranges");
        if (r > 0 && (pt.own0 != parts[r - 1].own1 || pt.v0 > parts[r - 1].own1 || pt.own_z0 != parts[r - 1].own_z1))
            return fail(FABBER_CUDA_ERR_INVALID, "spatial multi: parts must tile the voxel list in order");
        if (r + 1 < W && pt.v1 < pt.own1)
            return fail(FABBER_CUDA_ERR_INVALID, "spatial multi: bad part ranges");
    }
    if (parts[0].own0 != 0 || parts[W - 1].own1 != n_global)
        return fail(FABBER_CUDA_ERR_INVALID, "spatial multi: parts must cover the whole voxel list");
    /* peer access between neighbouring devices (and all-to-all for the aK mailboxes) */
    int ranks_on_device[64] = { 0 };
    for (int r = 0; r < W; r++)
    {
        if (parts[r].device < 0 || parts[r].device >= 64)
            return fail(FABBER_CUDA_ERR_INVALID, "spatial multi: bad device ordinal");
        ranks_on_device[parts[r].device]++;
    }
    for (int r = 0; r < W; r++)
        for (int q = 0; q < W; q++)
            if (parts[r].device != parts[q].device)
            {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, parts[r].device, parts[q].device);
                if (!can)
                    return fail(FABBER_CUDA_ERR_CUDA, "spatial multi: the devices cannot access each other's memory");
                cudaSetDevice(parts[r].device);
                cudaError_t pe = cudaDeviceEnablePeerAccess(parts[q].device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled)
                    return cuda_fail(pe, "cudaDeviceEnablePeerAccess");
                cudaGetLastError();
                /* the slabs' state lives in the stream-ordered allocator's pool, which peer access does not cover:
                 * the pool of device q must grant device r access explicitly */
                cudaMemPool_t pool;
                pe = cudaDeviceGetDefaultMemPool(&pool, parts[q].device);
                if (pe == cudaSuccess)
                {
                    cudaMemAccessDesc desc;
                    memset(&desc, 0, sizeof(desc));
                    desc.location.type = cudaMemLocationTypeDevice;
                    desc.location.id = parts[r].device;
                    desc.flags = cudaMemAccessFlagsProtReadWrite;
                    pe = cudaMemPoolSetAccess(pool, &desc, 1);
                }
                if (pe != cudaSuccess)
                    return cuda_fail(pe, "cudaMemPoolSetAccess (peer access to the slabs' scratch)");
            }
    std::vector<fabber_cuda_vb_problem> probs(W, *prob);
    std::vector<cudaStream_t> streams(W, nullptr);
    struct Streams
    {
        std::vector<cudaStream_t> &s;
        const fabber_cuda_slab_part *parts;
        ~Streams()
        {
            for (size_t i = 0; i < s.size(); i++)
                if (s[i])
                {
                    cudaSetDevice(parts[i].device);
                    cudaStreamDestroy(s[i]);
                }
        }
    } stream_guard = { streams, parts };
    /* per-device start / end events on the engine's streams: the device-side duration of the call (max over
     * the devices) for fabber_cuda_last_multi_ms() */
    std::vector<cudaEvent_t> ev0(W, nullptr), ev1(W, nullptr);
    struct Events
    {
        std::vector<cudaEvent_t> &a, &b;
        ~Events()
        {
            for (size_t i = 0; i < a.size(); i++)
            {
                if (a[i])
                    cudaEventDestroy(a[i]);
                if (b[i])
                    cudaEventDestroy(b[i]);
            }
        }
    } event_guard = { ev0, ev1 };
    /* declared after the streams: the runs hand their scratch back to the stream-ordered allocator ON their
     * stream when they go out of scope, so the streams must outlive them */
    std::vector<SpRun> runs(W);
    for (int r = 0; r < W; r++)
    {
        cudaSetDevice(parts[r].device);
        cudaError_t se = cudaStreamCreateWithFlags(&streams[r], cudaStreamNonBlocking);
        if (se == cudaSuccess)
            se = cudaEventCreate(&ev0[r]);
        if (se == cudaSuccess)
            se = cudaEventCreate(&ev1[r]);
        if (se == cudaSuccess)
            se = cudaEventRecord(ev0[r], streams[r]);
        if (se != cudaSuccess)
            return cuda_fail(se, "spatial multi: stream");
        probs[r].n_voxels = parts[r].v1 - parts[r].v0;
    }
    {
        /* set-up (neighbours, hyper-plane sort, permutation of the series) has host synchronisations inside:
         * one host thread per slab so that the devices prepare side by side */
        std::vector<int> rcs(W, FABBER_CUDA_OK);
        std::vector<std::string> msgs(W);
        std::vector<std::thread> workers;
        for (int r = 0; r < W; r++)
            workers.emplace_back([&, r]() {
                cudaSetDevice(parts[r].device);
                rcs[r] = runs[r].prepare(&probs[r], &parts[r].buf, streams[r], n_global);
                if (rcs[r] != FABBER_CUDA_OK)
                    msgs[r] = g_last_error;
            });
        for (size_t i = 0; i < workers.size(); i++)
            workers[i].join();
        for (int r = 0; r < W; r++)
            if (rcs[r] != FABBER_CUDA_OK)
                return fail(rcs[r], msgs[r]);
    }
    const int P = runs[0].P, max_it = runs[0].max_it;
    const bool any_spatial = runs[0].any_spatial, any_coupled = runs[0].any_coupled;
    /* flags, mailboxes, error word, neighbour links */
    std::vector<unsigned long long *> flags(W);
    std::vector<double *> mail(W);
    std::vector<int *> err(W);
    for (int r = 0; r < W; r++)
    {
        SpRun &R = runs[r];
        cudaSetDevice(parts[r].device);
        flags[r] = R.get<unsigned long long>(SLAB_FLAG_MAIL + SLAB_MAX_WORLD);
        mail[r] = R.get<double>((size_t)2 * W * (2 * P + 1));
        err[r] = R.get<int>(1);
        int *up_pos = R.get<int>(R.N);
        if (!flags[r] || !mail[r] || !err[r] || !up_pos)
            return fail(FABBER_CUDA_ERR_CUDA, "out of device memory for the spatial VB state");
        cudaMemsetAsync(flags[r], 0, (SLAB_FLAG_MAIL + SLAB_MAX_WORLD) * sizeof(unsigned long long), R.st);
        cudaMemsetAsync(mail[r], 0, (size_t)2 * W * (2 * P + 1) * sizeof(double), R.st);
        cudaMemsetAsync(err[r], 0, sizeof(int), R.st);
        cudaMemsetAsync(up_pos, 0xff, (size_t)R.N * sizeof(int), R.st);
        mark_ghost_range_kernel<<<R.gridN, 256, 0, R.st>>>(R.order, R.N, parts[r].own0 - parts[r].v0,
            parts[r].own1 - parts[r].v0, R.status_p);
        count_launch();
        SlabLinks &lk = R.sp.link;
        lk.world = W;
        lk.rank = r;
        lk.own_z0 = parts[r].own_z0;
        lk.own_z1 = parts[r].own_z1;
        lk.inplane_span = (prob->spatial_dims >= 3 || prob->spatial_dims == 0) ? prob->nx + prob->ny - 2
                                                                              : prob->nx + prob->ny - 2;
        lk.flags = flags[r];
        lk.error = err[r];
        {
            /* developer knob: FABBER_B200_SLAB_FORWARD=direct restores forwarding by the voxel threads */
            const char *fw = getenv("FABBER_B200_SLAB_FORWARD");
            lk.forward_late = (fw && std::string(fw) == "direct") ? 0 : 1;
        }
        lk.up_pos = (r + 1 < W) ? up_pos : nullptr;
        /* small slabs: planes hold few voxels, and the grid-wide barrier gets cheaper with every CTA less; ranks
         * that share a device (tests on one GPU) must all be co-resident, they spin on each other's flags */
        R.sp.sweep_share = ranks_on_device[parts[r].device];
        /* a slab's hyper-planes hold at most nx * ny voxels... and far fewer CTAs make the barrier cheaper */
        {
            const long long widest = (long long)prob->nx * prob->ny;
            R.sp.sweep_max_ctas = (int)std::max<long long>(8, std::min<long long>(1 << 20, (widest + SP_SWEEP_WORKERS - 1) / SP_SWEEP_WORKERS));
        }
    }
    for (int r = 0; r < W; r++)
        cudaStreamSynchronize(streams[r]); /* rank[] tables are read across devices below */
    for (int r = 0; r < W; r++)
    {
        SpRun &R = runs[r];
        SlabLinks &lk = R.sp.link;
        cudaSetDevice(parts[r].device);
        for (int q = 0; q < W; q++)
        {
            lk.mail[q] = mail[q];
            lk.mail_flags[q] = flags[q];
        }
        if (r + 1 < W)
        {
            /* our own voxels that the slab above also holds (its lower ghosts): global ids [v0_up, own0_up) */
            const fabber_cuda_slab_part &up = parts[r + 1];
            lk.up_flags = flags[r + 1];
            lk.up_mean = runs[r + 1].mean_p;
            lk.up_N = runs[r + 1].N;
            const int g0 = std::max(up.v0, parts[r].own0), g1 = up.own0;
            if (g1 > g0)
            {
                slab_up_pos_kernel<<<(g1 - g0 + 255) / 256, 256, 0, R.st>>>(R.rank, parts[r].v0, runs[r + 1].rank, up.v0,
                    g0, g1, parts[r].own1, const_cast<int *>(lk.up_pos), R.N, runs[r + 1].N);
                count_launch();
            }
        }
        if (r > 0)
        {
            /* our own voxels that the slab below also holds (its upper ghosts): global ids [own0, v1_dn) */
            const fabber_cuda_slab_part &dn = parts[r - 1];
            lk.dn_flags = flags[r - 1];
            lk.dn_mean = runs[r - 1].mean_p;
            lk.dn_N = runs[r - 1].N;
            const int g0 = parts[r].own0, n_dn = std::max(0, std::min(dn.v1, parts[r].own1) - g0);
            int *dn_src = R.get<int>(n_dn), *dn_dst = R.get<int>(n_dn);
            if (!dn_src || !dn_dst)
                return fail(FABBER_CUDA_ERR_CUDA, "out of device memory for the spatial VB state");
            if (n_dn > 0)
            {
                slab_dn_list_kernel<<<(n_dn + 255) / 256, 256, 0, R.st>>>(R.rank, parts[r].v0, runs[r - 1].rank, dn.v0, g0,
                    n_dn, dn_src, dn_dst, R.N, runs[r - 1].N);
                count_launch();
            }
            lk.dn_src = dn_src;
            lk.dn_dst = dn_dst;
            lk.n_dn = n_dn;
        }
    }
    for (int r = 0; r < W; r++)
    {
        /* every kernel of the loop must be loaded on its device before any slab starts to spin on another */
        cudaSetDevice(parts[r].device);
        cudaFuncAttributes at;
        cudaError_t pe = cudaFuncGetAttributes(&at, (const void *)slab_wait_kernel);
        if (pe == cudaSuccess && runs[r].ml->sp_preload)
            pe = runs[r].ml->sp_preload();
        if (pe != cudaSuccess)
            return cuda_fail(pe, "spatial multi: loading the kernels");
        cudaStreamSynchronize(streams[r]);
    }

    /* ---- the iteration-major loop: every launch is asynchronous; the slabs order themselves through the flags
     * in each other's memory, the host only queues work ---------------------------------------------------- */
    for (int it = 0; it < max_it; it++)
        for (int r = 0; r < W; r++)
        {
            SpRun &R = runs[r];
            SpArgs &sp = R.sp;
            cudaSetDevice(parts[r].device);
            cudaError_t le = cudaSuccess;
            sp.it = it;
            sp.ak_update = (any_spatial && (it > 0 || prob->update_first_iter)) ? 1 : 0;
            sp.ak_phase = 3;
            if (sp.ak_update)
            {
                if (it > 0)
                {
                    slab_wait_kernel<<<1, 1, 0, R.st>>>(flags[r], r > 0, r + 1 < W, it, err[r]);
                    count_launch();
                }
                le = R.ml->sp_ak_partial(sp, R.st);
            }
            if (le == cudaSuccess)
                le = R.ml->sp_ak_final(sp, R.st);
            R.snapshot_status();
            if (le == cudaSuccess)
                le = R.ml->sp_theta(sp, R.st);
            /* the sweep kernel also carries the slab coupling (forwarding, halo, flags): it runs even when no
             * parameter needs the ordered sweep, then with an empty plane range */
            sp.plane_first = 0;
            sp.plane_last = any_coupled ? sp.n_planes : 0;
            if (le == cudaSuccess)
                le = R.ml->sp_sweep(sp, R.st);
            if (le == cudaSuccess)
                le = R.ml->sp_noise(sp, R.st);
            if (le != cudaSuccess)
                return cuda_fail(le, "spatial VB (multi-device) launch");
        }
    int rc = FABBER_CUDA_OK;
    for (int r = 0; r < W; r++)
    {
        SpRun &R = runs[r];
        cudaSetDevice(parts[r].device);
        R.sp.it = max_it;
        R.sp.ak_update = 0;
        R.sp.ak_phase = 0;
        cudaError_t le = R.ml->sp_ak_final(R.sp, R.st);
        if (le != cudaSuccess)
            return cuda_fail(le, "spatial VB (multi-device) launch");
    }
    for (int r = 0; r < W; r++)
    {
        /* queue every slab's result permutation before waiting for any of them */
        cudaSetDevice(parts[r].device);
        runs[r].finish_async();
        cudaEventRecord(ev1[r], streams[r]);
    }
    double worst_ms = 0.0;
    for (int r = 0; r < W; r++)
    {
        cudaSetDevice(parts[r].device);
        const int frc = runs[r].finish();
        if (frc != FABBER_CUDA_OK && rc == FABBER_CUDA_OK)
            rc = frc;
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev0[r], ev1[r]) == cudaSuccess)
            worst_ms = std::max(worst_ms, (double)ms);
        int h_err = 0;
        cudaMemcpy(&h_err, err[r], sizeof(int), cudaMemcpyDeviceToHost);
        if (h_err && rc == FABBER_CUDA_OK)
            rc = fail(FABBER_CUDA_ERR_CUDA, "spatial multi: a slab timed out waiting for its neighbour");
    }
    g_last_multi_ms.store(worst_ms);
    return rc;
}

double fabber_cuda_last_multi_ms(void) { return g_last_multi_ms.load(); }

int fabber_cuda_check_status(const int *status, int n_voxels, int *first_bad_voxel, int *first_bad_code, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (first_bad_voxel)
        *first_bad_voxel = -1;
    if (first_bad_code)
        *first_bad_code = 0;
    if (n_voxels <= 0)
        return 0;
    int *d = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d, 3 * sizeof(int), s);
    if (e != cudaSuccess)
        return cuda_fail(e, "cudaMallocAsync(status)");
    int h[3] = { 0, 0x7fffffff, 0 };
    cudaMemcpyAsync(d, h, sizeof(h), cudaMemcpyHostToDevice, s);
    status_scan_kernel<<<(unsigned)((n_voxels + 255) / 256), 256, 0, s>>>(status, n_voxels, d);
    count_launch();
    cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, s);
    e = cudaStreamSynchronize(s);
    if (e == cudaSuccess && h[0] > 0)
    {
        int code = 0;
        e = cudaMemcpy(&code, status + h[1], sizeof(int), cudaMemcpyDeviceToHost);
        if (first_bad_voxel)
            *first_bad_voxel = h[1];
        if (first_bad_code)
            *first_bad_code = code;
    }
    cudaFreeAsync(d, s);
    if (e != cudaSuccess)
        return cuda_fail(e, "check_status");
    return h[0];
}

int fabber_cuda_max_int(const int *values, int n, int *result, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!result)
        return fail(FABBER_CUDA_ERR_INVALID, "null result");
    *result = 0;
    if (n <= 0)
        return FABBER_CUDA_OK;
    int *d = nullptr;
    cudaError_t e = cudaMallocAsync((void **)&d, sizeof(int), st);
    if (e != cudaSuccess)
        return cuda_fail(e, "cudaMallocAsync(max)");
    int init = INT_MIN;
    cudaMemcpyAsync(d, &init, sizeof(int), cudaMemcpyHostToDevice, st);
    max_int_kernel<<<592, 256, 0, st>>>(values, n, d);
    count_launch();
    cudaMemcpyAsync(result, d, sizeof(int), cudaMemcpyDeviceToHost, st);
    e = cudaStreamSynchronize(st);
    cudaFreeAsync(d, st);
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "max_int");
}

static int model_fit_impl(const fabber_cuda_vb_problem *prob, const double *mean, double *fit, const float *data,
    float *fit_f32, float *resid_f32, cudaStream_t st);

int fabber_cuda_vb_save_results(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf,
    const fabber_cuda_vb_outputs *out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (!prob || !buf || !out)
        return fail(FABBER_CUDA_ERR_INVALID, "null argument");
    const int P = prob->model.n_params, N = prob->n_voxels;
    if (P < 1 || P > FABBER_CUDA_MAX_PARAMS || N < 0)
        return fail(FABBER_CUDA_ERR_INVALID, "bad sizes");
    if (N == 0)
        return FABBER_CUDA_OK;
    const bool nlls = prob->method == FABBER_METHOD_NLLS; /* no noise block: the result MVN is the model's alone */
    if (!buf->mean || !buf->cov || (!buf->noise && !nlls))
        return fail(FABBER_CUDA_ERR_INVALID, "missing mean / cov / noise result arrays");
    SaveArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N;
    a.P = P;
    a.ar = prob->noise_type == FABBER_NOISE_AR1;
    a.n_phis = a.ar ? (prob->n_phis == 2 ? 2 : 1) : prob->n_phis;
    a.n_alphas = 2 + (a.ar && a.n_phis == 2 ? prob->ar_cross_terms : 0);
    if (a.n_alphas < 2 || a.n_alphas > 4)
        return fail(FABBER_CUDA_ERR_INVALID, "ar_cross_terms out of range");
    /* size of the noise block of the result MVN: the phis, or (alphas, phis) - Ar1cParams::OutputAsMVN */
    a.n_noise = a.ar ? a.n_alphas + a.n_phis : prob->n_phis;
    if (a.n_noise < 1 || a.n_noise > 6 || (!a.ar && a.n_noise > FABBER_CUDA_MAX_PHIS))
        return fail(FABBER_CUDA_ERR_INVALID, "n_phis out of range");
    if (nlls)
        a.ar = a.n_phis = a.n_noise = 0;
    a.f_len = buf->f_history ? prob->f_history_len : 0;
    for (int i = 0; i < P; i++)
        a.transform[i] = prob->params[i].transform;
    a.mean = buf->mean;
    a.cov = buf->cov;
    a.noise = buf->noise;
    a.free_energy = buf->free_energy;
    a.f_history = buf->f_history;
    a.iterations = buf->iterations;
    a.out = *out;
    if (out->mean || out->std || out->zstat || out->var || out->noise_mean || out->noise_std || out->final_mvn
        || out->free_energy || out->f_history)
    {
        save_results_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(a);
        count_launch();
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            return cuda_fail(e, "save_results launch");
    }
    if (out->model_fit || out->residuals)
    {
        if (out->residuals && !out->data)
            return fail(FABBER_CUDA_ERR_INVALID, "residuals need the input series");
        return model_fit_impl(prob, buf->mean, nullptr, out->data, out->model_fit, out->residuals, st);
    }
    return FABBER_CUDA_OK;
}

int fabber_cuda_model_fit(const fabber_cuda_vb_problem *prob, const double *mean, double *fit, void *stream)
{
    if (!fit)
        return fail(FABBER_CUDA_ERR_INVALID, "null argument");
    return model_fit_impl(prob, mean, fit, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

static int model_fit_impl(const fabber_cuda_vb_problem *prob, const double *mean, double *fit, const float *data,
    float *fit_f32, float *resid_f32, cudaStream_t st)
{
    if (!prob || !mean)
        return fail(FABBER_CUDA_ERR_INVALID, "null argument");
    const int P = prob->model.n_params, T = prob->n_times, N = prob->n_voxels;
    if (P < 1 || P > FABBER_CUDA_MAX_PARAMS || T < 1 || N < 0)
        return fail(FABBER_CUDA_ERR_INVALID, "bad sizes");
    const ModelLaunchers *ml = find_model(prob->model, P);
    if (!ml || !ml->model_fit)
        return fail(FABBER_CUDA_ERR_INVALID, "no device Evaluate hook compiled for this model / parameter count");
    VbArgs a;
    memset(&a, 0, sizeof(a));
    a.N = N;
    a.T = T;
    for (int i = 0; i < P; i++)
        a.params[i] = prob->params[i];
    a.exp_dt = prob->model.exp_dt;
    memcpy(a.model_consts, prob->model.consts, sizeof(a.model_consts));
    a.design_len = prob->model.design_len;
    a.fit_mean = mean;
    a.fit_out = fit;
    a.fit_out_f32 = fit_f32;
    a.resid_out_f32 = resid_f32;
    a.data = data;
    Staged staged;
    const bool plugin_vec = prob->model.id == FABBER_MODEL_PLUGIN && prob->model.design && prob->model.design_len > 0;
    if (prob->model.id == FABBER_MODEL_LINEAR || plugin_vec)
    {
        if (!prob->model.design)
            return fail(FABBER_CUDA_ERR_INVALID, "linear model without design matrix");
        const size_t bytes = (plugin_vec ? (size_t)prob->model.design_len : (size_t)T * P) * sizeof(double);
        cudaError_t e = cudaMallocAsync((void **)&staged.design, bytes, st);
        if (e != cudaSuccess)
            return cuda_fail(e, "cudaMallocAsync(design)");
        cudaMemcpyAsync(staged.design, prob->model.design, bytes, cudaMemcpyHostToDevice, st);
        a.design = staged.design;
    }
    cudaError_t e = ml->model_fit(a, st);
    staged.release(st);
    if (e != cudaSuccess)
        return cuda_fail(e, "model_fit launch");
    return FABBER_CUDA_OK;
}

double fabber_cuda_measure_fp64_peak(int repeats)
{
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess)
        return -1.0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 256, blocks = sms * 8, iters = 4096;
    double *out = nullptr;
    if (cudaMalloc((void **)&out, (size_t)threads * blocks * sizeof(double)) != cudaSuccess)
        return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    if (repeats < 1)
        repeats = 1;
    for (int r = 0; r < repeats + 1; r++)
    {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, threads>>>(out, iters, 1.0000001);
        count_launch();
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess)
        {
            best = -1.0;
            break;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * (double)iters * threads * blocks;
        const double gf = flops / (ms * 1e-3) * 1e-9;
        if (r > 0 && gf > best) /* first repeat is warm-up */
            best = gf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}

unsigned long long fabber_cuda_launch_count(void) { return g_launches.load(); }

int fabber_cuda_exp_probe(const double *x, double *fast, double *ref, int n, void *stream)
{
    if (n <= 0)
        return FABBER_CUDA_OK;
    exp_probe_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, fast, ref, n);
    count_launch();
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? FABBER_CUDA_OK : cuda_fail(e, "exp_probe");
}

/* struct sizes, so that language bindings can verify their mirror of the header */
int fabber_cuda_sizeof_problem(void) { return (int)sizeof(fabber_cuda_vb_problem); }
int fabber_cuda_sizeof_buffers(void) { return (int)sizeof(fabber_cuda_vb_buffers); }

} // extern "C"
