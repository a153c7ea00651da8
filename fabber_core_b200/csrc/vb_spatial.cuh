/*
 * vb_spatial.cuh - spatial VB (iteration-major), white noise.
 *
 * Replaces Vb::DoCalculationsSpatial (inference_vb.cc:578-767) with SpatialPrior::CalculateaK /
 * ApplyToMVN (priors.cc:221-488) and Vb::CalcNeighbours (inference_vb.cc:830-964).
 *
 * The reference sweeps the voxels sequentially: voxel v's MRF prior mean uses the posterior means of
 * its neighbours, those with a smaller index already updated in the same sweep (Gauss-Seidel). To
 * keep that exactly while running in parallel the sweep is split:
 *
 *   sp_theta_kernel   (parallel)  priors' precisions, Lambda = Lambda0 + phi A, Sigma = Lambda^-1 and the
 *                                 neighbour-independent part of the right-hand side. The spatial prior
 *                                 *precision* depends only on aK and the neighbour count (priors.cc:406-429).
 *   sp_sweep_kernel   (wavefront) m_v = Sigma_v (rhs_v + sum_k spatial_prec_k * mean_nn,k e_k), launched once
 *                                 per hyper-plane x+y+z = h in increasing h: in lexicographic voxel order
 *                                 (x fastest) the already-updated neighbours -x,-y,-z lie on plane h-1 and
 *                                 the not-yet-updated ones +x,+y,+z on plane h+1, so all voxels of a plane
 *                                 are independent and see exactly the values the sequential sweep sees.
 *   sp_noise_kernel   (parallel)  UpdateNoise, ReCentre (the pass over the time-series -> new sufficient
 *                                 statistics) and the free energy, second loop of the reference (:675-722).
 *   sp_ak_*           (reduction) the two global sums of CalculateaK per spatial parameter, deterministic
 *                                 two-stage reduction; aK itself is computed on the device.
 *
 * P/p (Penny) priors: `double rec = 1 / (8*nn - nn2)` is an INTEGER division in the reference
 * (priors.cc:455) and |8 nn - nn2| >= 3 for every reachable neighbour count, so rec == 0 and the
 * neighbour mean drops out of the prior mean; only the precision and aK see the neighbours. Those types
 * therefore need no ordered sweep (and no second-neighbour lists) - reproduced, not "fixed".
 *
 * Between kernels the per-voxel state lives in HBM as structure-of-arrays [field][N] doubles. All kernels
 * of a spatial run index voxels in HYPER-PLANE-MAJOR order (fabber_cuda.cu renumbers inputs on the way in
 * and results on the way out), so that a plane of the sweep is a contiguous, coalesced range.
 */
#pragma once
#include <cooperative_groups.h>

#include "vb_voxelwise.cuh"

namespace fab
{
/* ---- device-driven z-slab coupling (fabber_cuda_vb_spatial_multi) ----------------------------------------
 * One volume, cut into z-slabs over several GPUs. Each slab's kernels talk to the neighbouring slabs' memory
 * DIRECTLY (peer access over NVLink): no host callback, no separate pack / send / receive / unpack step.
 *   forward   the ordered sweep of slab r stores the fresh mean of every voxel of its TOP own plane straight
 *             into slab r+1's lower ghost voxel, and after the grid-wide barrier of hyper-plane H publishes
 *             "planes <= H done" in r+1's FWD flag (st.release.sys); slab r+1 spins on that flag
 *             (ld.acquire.sys) before it finishes hyper-plane H+1 - the exact Gauss-Seidel order across slabs
 *             with ONE hyper-plane of skew.
 *   halo      at the end of its sweep slab r stores its BOTTOM own plane into slab r-1's upper ghost voxels
 *             and bumps r-1's HI flag (the upper ghost must stay one iteration old during the sweep).
 *   aK sums   every slab writes its partial sums of CalculateaK into EVERY slab's mailbox, raises that slab's
 *             mail flag, waits for all its own flags and adds the world's partials in rank order: an
 *             all-gather fused into sp_ak_final, bit-identical on every slab.
 * Flags only ever grow (iteration number folded in), mailboxes are double-buffered by iteration parity; every
 * spin has a cycle budget and reports through `error` instead of hanging the GPU. */
constexpr int SLAB_MAX_WORLD = 16;
constexpr unsigned long long SLAB_IT_STRIDE = 1ull << 20; /* FWD flag = it * stride + hyper-planes done */
constexpr int SLAB_PUBLISH_EVERY = 4; /* the sweep publishes "planes done" to the slab above every so many planes */
enum
{
    SLAB_FLAG_FWD = 0, /* written by the slab below */
    SLAB_FLAG_HI = 1,  /* written by the slab above: iterations whose bottom plane has been published */
    SLAB_FLAG_MAIL = 2 /* [SLAB_FLAG_MAIL + q]: written by slab q: aK iterations mailed */
};
struct SlabLinks
{
    int world, rank; /* world == 0: not a slab run */
    int own_z0, own_z1, inplane_span; /* own z-planes [own_z0, own_z1); nx + ny - 2 restricted by spatial_dims */
    unsigned long long *flags;        /* this slab's flag block */
    unsigned long long *up_flags, *dn_flags; /* the neighbours' (peer memory), NULL at the ends */
    double *up_mean; /* slab r+1's mean array [P][up_N] (peer memory) */
    int up_N;
    const int *up_pos; /* [N]: position in slab r+1 of the same voxel (its lower ghost), -1 = none */
    double *dn_mean;   /* slab r-1's mean array [P][dn_N] (peer memory) */
    int dn_N;
    const int *dn_src, *dn_dst; /* [n_dn]: own bottom-plane position -> position in slab r-1 (its upper ghost) */
    int n_dn;
    double *mail[SLAB_MAX_WORLD]; /* every slab's mailbox [2][world][2P] (peer memory; own included) */
    unsigned long long *mail_flags[SLAB_MAX_WORLD]; /* every slab's flag block */
    int *error; /* set to 1 when a spin ran out of budget */
    /* 1: the sweep's top plane is forwarded by the idle lanes of the barrier warp one hyper-plane LATER, from the
     * means the voxel threads stored locally (default); 0: by the voxel threads themselves as they finish a voxel.
     * See sp_sweep_kernel. */
    int forward_late;
};

FAB_DEV unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
FAB_DEV void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
/* spin until *flag >= want; ~2 s budget (a dead peer must not hang the GPU: gpurun counts that as a strike) */
FAB_DEV unsigned long long slab_wait(const unsigned long long *flag, unsigned long long want, int *error)
{
    const long long t0 = clock64();
    unsigned long long seen;
    while ((seen = ld_acquire_sys(flag)) < want)
    {
        /* once any wait of this slab has given up, none waits again: the run is lost, end it quickly */
        if (*(volatile int *)error != 0)
            break;
        if (clock64() - t0 > 4000000000ll)
        {
            atomicExch(error, 1);
            break;
        }
        __nanosleep(64);
    }
    return seen; /* >= want unless the wait gave up */
}

struct SpArgs
{
    VbArgs v;
    const int *nn_idx;   /* [6][N]: +x,-x,+y,-y,+z,-z neighbour voxel or -1 (restricted by spatial_dims) */
    double *centre;      /* [P][N] */
    double *stats;       /* [NT + P + 1][N] : A, b, rr */
    double *m0, *L0;     /* [P][N] prior mean / diagonal prior precision */
    double *rhs;         /* [P][N] neighbour-independent right-hand side */
    double *logdet;      /* [N] log|det Lambda| */
    double *aK;          /* [P] current spatial precisions (device) */
    double *ak_hist;     /* [max_it + 1][P] (device) */
    double *ak_partial;  /* [ak_blocks][2][P] */
    double *fprior_last; /* [1] stale Fprior of the last voxel (inference_vb.cc:700) */
    const int *last_pos; /* [1] plane-major position of the caller's last voxel */
    const int *order;    /* original voxel index of each (plane-major) position */
    const int *plane_starts; /* [n_planes + 1] offsets into order (device) */
    int n_planes;
    int plane_first, plane_last; /* sp_sweep_kernel walks planes [plane_first, plane_last) */
    int ak_blocks;
    int it;
    int spatial_dims;
    int update_first_iter;
    int any_coupled; /* any 'M' / 'm' parameter */
    int ak_update;   /* sp_ak_final_kernel: recompute aK (else only record the history row) */
    int ak_phase;    /* 0: partials -> sums and aK in one go (one GPU); 1: partials -> ak_sums only;
                        2: aK from ak_sums (after the host all-reduced them across z-slabs) */
    double *ak_sums; /* [2][P] */
    int n_global;    /* voxels of the WHOLE volume (hK = N/2 + q2, priors.cc:313); == v.N on one GPU */
    /* [N] status words as they were when this iteration began (== v.status unless allow_bad_voxels): which
     * neighbours Vb::IgnoreVoxel had already struck from the lists, see nbr_alive() */
    const int *status_prev;
    int ignore_bad;     /* allow_bad_voxels: failed voxels are struck from their neighbours' lists (nbr_alive) */
    SlabLinks link;     /* link.world == 0 on one GPU and in the host-callback slab mode */
    int sweep_max_ctas; /* > 0: cap on the cooperative sweep grid (small slabs: the barrier gets cheaper) */
    unsigned *sweep_barrier; /* [1] arrival counter of the sweep's grid-wide barrier, zeroed before every launch */
    int sweep_share;    /* > 1: that many slabs share this GPU and spin on each other's flags - each sweep grid is
                           held to 1/share of what the GPU can keep resident, so that all of them fit */
    double q1, q2, speed;
};

FAB_DEV bool is_spatial_type(char t) { return t == 'M' || t == 'm' || t == 'P' || t == 'p'; }

/* Vb::IgnoreVoxel (inference_vb.cc:266-297): a voxel that failed under allow-bad-voxels is erased from its
 * neighbours' lists, so it no longer counts towards nn and its stale mean no longer feeds the MRF prior mean
 * or the aK sums. The lists here are static; a neighbour is skipped when its status word reports a failure.
 * Ghost voxels of a z-slab are alive (their owner updates them). */
template <bool IGNORE> FAB_DEV bool nbr_alive(const int *status, int n)
{
    if (n < 0)
        return false;
    if (!IGNORE) /* without allow-bad-voxels the first failure ends the run: no status look-ups (compile time: a
                    run-time test here cost the latency-bound kernels 15-25 %, measured) */
        return true;
    const int st = status[n];
    return st == 0 || st == FABBER_VOX_GHOST;
}

/* ---- set-up: initial posterior, noise, first ReCentre (Vb::SetupPerVoxelDists) ------------------- */
template <class Model> __global__ void __launch_bounds__(VB_BLOCK, FAB_MIN_BLOCKS) sp_setup_kernel(const __grid_constant__ SpArgs s)
{
    constexpr int P = Model::P;
    constexpr int NT = NTri<P>::value;
    const VbArgs &a = s.v;
    extern __shared__ double smem[];
    Model::stage(a, smem);
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.N)
        return;
    const typename Model::Ctx mc = Model::make_ctx(a, smem);
    const size_t N = (size_t)a.N;
    double m[P], Sig[NT];
    int status = 0;
    if (a.init_mean)
    {
#pragma unroll
        for (int i = 0; i < P; i++)
            m[i] = a.init_mean[i * N + v];
#pragma unroll
        for (int i = 0; i < NT; i++)
            Sig[i] = a.init_cov[i * N + v];
    }
    else
    {
        double var[P];
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            m[i] = (a.params[i].prior_type == 'I') ? a.image_prior[i][v] : a.params[i].post_mean;
            var[i] = a.params[i].post_var;
        }
        Model::init_voxel(a, v, m);
#pragma unroll
        for (int i = 0; i < NT; i++)
            Sig[i] = 0.0;
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            const char code = a.params[i].transform;
            m[i] = to_fabber(code, m[i]);
            Sig[tri(i, i)] = to_fabber_var(code, var[i]);
        }
    }
    Stats<P> S[1];
    double c[P];
#pragma unroll
    for (int i = 0; i < P; i++)
        c[i] = a.lock_centre ? a.lock_centre[i * N + v] : m[i]; /* inference_vb.cc:227-236 */
    const int err = recentre_stats<Model, 1, true>(a, mc, nullptr, v, c, S);
    if (err)
        status = err | FABBER_VOX_SETUP_FLAG;
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        a.mean[i * N + v] = m[i];
        s.centre[i * N + v] = c[i];
        s.m0[i * N + v] = 0.0;
        s.L0[i * N + v] = 1.0;
        s.stats[(NT + i) * N + v] = S[0].b[i];
    }
#pragma unroll
    for (int i = 0; i < NT; i++)
    {
        a.cov[i * N + v] = Sig[i];
        s.stats[i * N + v] = S[0].A[i];
    }
    s.stats[(NT + P) * N + v] = S[0].rr;
    a.noise[0 * N + v] = a.init_noise ? a.init_noise[0 * N + v] : a.noise_post_b[0];
    a.noise[1 * N + v] = a.init_noise ? a.init_noise[1 * N + v] : a.noise_post_c[0];
    s.logdet[v] = 0.0;
    if (a.free_energy)
        a.free_energy[v] = 9999.0; /* resultFs default, inference_vb.cc:165 */
    if (a.iterations)
        a.iterations[v] = 0;
    a.status[v] = status;
}

/* ---- aK: per-block partial sums of trace_term and term2 (priors.cc:233-294) ------------------------ */
template <int P, bool IGNORE, bool SLAB>
__global__ void __launch_bounds__(256) sp_ak_partial_kernel(const __grid_constant__ SpArgs s)
{
    const VbArgs &a = s.v;
    const size_t N = (size_t)a.N;
    /* (slab mode: the host queues slab_wait_kernel in front of this kernel - the sums read the ghost planes) */
    double tr[P], t2[P];
#pragma unroll
    for (int k = 0; k < P; k++)
        tr[k] = t2[k] = 0.0;
    const int dims = s.spatial_dims;
    for (int v = blockIdx.x * blockDim.x + threadIdx.x; v < a.N; v += gridDim.x * blockDim.x)
    {
        if (a.status[v] != 0)
            continue; /* ignore_voxels */
        int nbr[6], nn = 0;
#pragma unroll
        for (int j = 0; j < 6; j++)
        {
            nbr[j] = s.nn_idx[j * N + v];
            FAB_CHECK(a, nbr[j] >= -1 && nbr[j] < a.N && nbr[j] != v, 111);
            if (!nbr_alive<IGNORE>(a.status, nbr[j])) /* every failure so far: CalculateaK runs at v == 1 */
                nbr[j] = -1;
            nn += nbr[j] >= 0;
        }
#pragma unroll
        for (int k = 0; k < P; k++)
        {
            const char ty = a.params[k].prior_type;
            if (!is_spatial_type(ty))
                continue;
            const double sigmaK = a.cov[tri(k, k) * N + v];
            const double wK = a.mean[k * N + v];
            if (ty == 'm')
                tr[k] += sigmaK * dims * 2;
            else if (ty == 'M')
                tr[k] += sigmaK * (nn + 1e-8);
            else if (ty == 'p')
                tr[k] += sigmaK * (4 * dims * dims + 2 * dims);
            else
                tr[k] += sigmaK * (nn * nn + nn);
            double SwK = 0.0;
#pragma unroll
            for (int j = 0; j < 6; j++)
                if (nbr[j] >= 0) /* slab mode: ghosts are written by a peer GPU, never serve them from L1 */
                    SwK += wK - (SLAB ? __ldcg(a.mean + k * N + nbr[j]) : a.mean[k * N + nbr[j]]);
            if (ty == 'p' || ty == 'm')
                SwK += wK * (dims * 2 - (double)nn);
            if (ty == 'm' || ty == 'M')
                t2[k] += SwK * wK;
            else
                t2[k] += SwK * SwK;
        }
    }
    __shared__ double red[2 * P][256 / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < P; k++)
    {
        double x = tr[k], y = t2[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            x += __shfl_down_sync(0xffffffffu, x, o);
            y += __shfl_down_sync(0xffffffffu, y, o);
        }
        if (lane == 0)
        {
            red[k][warp] = x;
            red[P + k][warp] = y;
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * P)
    {
        double x = 0.0;
        for (int w = 0; w < 256 / 32; w++)
            x += red[threadIdx.x][w];
        s.ak_partial[(size_t)blockIdx.x * 2 * P + threadIdx.x] = x;
    }
}

/* one block: fixed-order (hence deterministic) final sum, then the Penny update for aK (priors.cc:296-343) */
template <int P> __global__ void __launch_bounds__(256) sp_ak_final_kernel(const __grid_constant__ SpArgs s)
{
    const VbArgs &a = s.v;
    __shared__ double red[256];
    __shared__ double sums[2 * P];
    if (s.ak_update && s.ak_phase == 2)
    {
        if (threadIdx.x < 2 * P)
            sums[threadIdx.x] = s.ak_sums[threadIdx.x];
        __syncthreads();
    }
    if (s.ak_update && s.ak_phase != 2)
        for (int q = 0; q < 2 * P; q++)
        {
            double x = 0.0;
            for (int b = threadIdx.x; b < s.ak_blocks; b += 256)
                x += s.ak_partial[(size_t)b * 2 * P + q];
            red[threadIdx.x] = x;
            __syncthreads();
            for (int o = 128; o > 0; o >>= 1)
            {
                if (threadIdx.x < o)
                    red[threadIdx.x] += red[threadIdx.x + o];
                __syncthreads();
            }
            if (threadIdx.x == 0)
                sums[q] = red[0];
            __syncthreads();
        }
    if (s.ak_phase == 1)
    {
        /* z-slab mode: hand the local sums to the host, which all-reduces them over the slabs */
        if (s.ak_update && threadIdx.x < 2 * P)
            s.ak_sums[threadIdx.x] = sums[threadIdx.x];
        return;
    }
    if (s.ak_phase == 3)
    {
        /* device-driven slab mode: all-gather through the slabs' mailboxes, then the same fixed-order sum on
         * every slab (priors.cc:233-343: two global sums per spatial parameter). The mail also carries the ARD
         * free-energy term of the volume's LAST voxel (inference_vb.cc:700 adds that voxel's stale Fprior to every
         * voxel's F): only the top slab owns it, every slab needs it. Runs every iteration. */
        constexpr int MP = 2 * P + 1;
        const int W = s.link.world, me = s.link.rank, slot = s.it & 1;
        __shared__ double fprior_mail;
        if (threadIdx.x == 0)
        {
            double fp = 0.0;
            if (me == W - 1)
            {
                const size_t N = (size_t)a.N;
                const int pos = *s.last_pos;
                for (int k = 0; k < P; k++)
                    if (a.params[k].prior_type == 'A')
                    {
                        const double mk = a.mean[k * N + pos];
                        const double bb = 2 / (mk * mk + a.cov[tri(k, k) * N + pos]);
                        fp += -1.5 * (log(bb) + digamma_fsl(0.5)) - 0.5 - gammaln(0.5) - 0.5 * log(bb);
                    }
            }
            fprior_mail = fp;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < MP * W; i += blockDim.x)
        {
            const int q = i / MP, j = i - q * MP;
            FAB_CHECK(a, q >= 0 && q < W && W <= SLAB_MAX_WORLD && me >= 0 && me < W && (slot == 0 || slot == 1) && s.link.mail[q] != nullptr, 112);
            s.link.mail[q][((size_t)slot * W + me) * MP + j] = j < 2 * P ? (s.ak_update ? sums[j] : 0.0) : fprior_mail;
        }
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < W)
            st_release_sys(s.link.mail_flags[threadIdx.x] + SLAB_FLAG_MAIL + me, (unsigned long long)s.it + 1);
        if (threadIdx.x < W)
            slab_wait(s.link.flags + SLAB_FLAG_MAIL + threadIdx.x, (unsigned long long)s.it + 1, s.link.error);
        __syncthreads();
        if (threadIdx.x < 2 * P)
        {
            double x = 0.0;
            for (int q = 0; q < W; q++)
                x += __ldcg(s.link.mail[me] + ((size_t)slot * W + q) * MP + threadIdx.x);
            sums[threadIdx.x] = x;
        }
        if (threadIdx.x == 0)
            *s.fprior_last = __ldcg(s.link.mail[me] + ((size_t)slot * W + (W - 1)) * MP + 2 * P);
        __syncthreads();
    }
    const int k = threadIdx.x;
    if (k >= P)
        return;
    const char ty = a.params[k].prior_type;
    if (s.ak_update && is_spatial_type(ty))
    {
        const double trace_term = sums[k], term2 = sums[P + k];
        const double gk = 1 / (0.5 * trace_term + 0.5 * term2 + 1 / s.q1);
        const double hK = (s.n_global * 0.5 + s.q2);
        double aK = gk * hK;
        if (aK < 1e-50)
            aK = 1e-50;
        double aKMax = aK * s.speed;
        if (aKMax < 0.5)
            aKMax = 0.5;
        if ((s.speed > 0) && (aK > aKMax))
            aK = aKMax;
        s.aK[k] = aK;
    }
    s.ak_hist[(size_t)s.it * P + k] = is_spatial_type(ty) ? s.aK[k] : 0.0;
}

/* ---- theta, neighbour-independent part (first loop of the reference, :614-650) --------------------- */
template <int P> __global__ void __launch_bounds__(VB_BLOCK) sp_theta_kernel(const __grid_constant__ SpArgs s)
{
    constexpr int NT = NTri<P>::value;
    const VbArgs &a = s.v;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.N)
        return;
    const size_t N = (size_t)a.N;
    if (a.status[v] != 0)
        return;
    double m[P], c[P], A[NT], b[P], L0[P], m0[P], sigd[P];
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        m[i] = a.mean[i * N + v];
        c[i] = s.centre[i * N + v];
        b[i] = s.stats[(NT + i) * N + v];
        sigd[i] = a.cov[tri(i, i) * N + v];
        m0[i] = s.m0[i * N + v];
    }
#pragma unroll
    for (int i = 0; i < NT; i++)
        A[i] = s.stats[i * N + v];
    int nn = 0;
#pragma unroll
    for (int j = 0; j < 6; j++)
        nn += s.ignore_bad ? nbr_alive<true>(s.status_prev, s.nn_idx[j * N + v]) : (s.nn_idx[j * N + v] >= 0);
    const int dims = s.spatial_dims;
    double Fprior = 0.0;
    bool coupled[P];
#pragma unroll
    for (int k = 0; k < P; k++)
    {
        const fabber_cuda_param &p = a.params[k];
        const char ty = p.prior_type;
        coupled[k] = false;
        if (ty == 'A')
        {
            const double new_cov = m[k] * m[k] + sigd[k];
            if (s.it == 0)
            {
                L0[k] = 1.0 / p.prior_var;
                m0[k] = p.prior_mean;
            }
            else
                L0[k] = 1.0 / new_cov;
            const double bb = 2 / new_cov;
            Fprior += -1.5 * (log(bb) + digamma_fsl(0.5)) - 0.5 - gammaln(0.5) - 0.5 * log(bb);
        }
        else if (is_spatial_type(ty))
        {
            const double aK = s.aK[k];
            int n1 = nn;
            if (ty == 'p' || ty == 'm')
                n1 = 2 * dims;
            double sp;
            if (ty == 'M')
                sp = aK * (n1 + 1e-8);
            else if (ty == 'm')
                sp = aK * n1;
            else
                sp = aK * (n1 * n1 + n1);
            L0[k] = (ty == 'p' || ty == 'm') ? sp : p.prior_prec + sp;
            if (ty == 'M' || ty == 'm')
                coupled[k] = true; /* prior mean needs the neighbours: sp_sweep_kernel */
            else
            {
                /* rec == 0 (integer division, priors.cc:455): spatial_mean == 0 */
                const double spatial_mean = 0.0;
                m0[k] = (1.0 / L0[k]) * (sp * spatial_mean + p.prior_prec * p.prior_mean);
            }
        }
        else
        {
            m0[k] = (ty == 'I') ? a.image_prior[k][v] : p.prior_mean;
            L0[k] = p.prior_prec;
        }
    }
    /* the reference's last voxel: its Fprior goes stale into the second loop (z-slabs: the top slab owns it, the
     * others got its value by mail in sp_ak_final) */
    if (s.order[v] == a.N - 1 && (s.link.world <= 1 || s.link.rank == s.link.world - 1))
        *s.fprior_last = Fprior;
    const double phi = a.noise[0 * N + v] * a.noise[1 * N + v];
    double Lam[NT], Sig[NT], ld;
#pragma unroll
    for (int i = 0; i < NT; i++)
        Lam[i] = phi * A[i];
#pragma unroll
    for (int i = 0; i < P; i++)
        Lam[tri(i, i)] = L0[i] + phi * A[tri(i, i)];
    if (!mvn_inverse<P>(Lam, Sig, ld, a.need_f != 0))
    {
        a.status[v] = FABBER_VOX_SINGULAR;
        return;
    }
    double Aw[NT], Ac[P], rhs[P];
#pragma unroll
    for (int i = 0; i < NT; i++)
        Aw[i] = phi * A[i];
    symv<P>(Aw, c, Ac);
#pragma unroll
    for (int i = 0; i < P; i++)
        rhs[i] = (phi * b[i] + Ac[i]) + (coupled[i] ? 0.0 : L0[i] * m0[i]);
#pragma unroll
    for (int i = 0; i < NT; i++)
        a.cov[i * N + v] = Sig[i];
    s.logdet[v] = ld;
#pragma unroll
    for (int i = 0; i < P; i++)
    {
        s.L0[i * N + v] = L0[i];
        if (!coupled[i])
            s.m0[i * N + v] = m0[i];
    }
    if (s.any_coupled)
    {
#pragma unroll
        for (int i = 0; i < P; i++)
            s.rhs[i * N + v] = rhs[i];
    }
    else
    {
        double mn[P];
        symv<P>(Sig, rhs, mn);
#pragma unroll
        for (int i = 0; i < P; i++)
            a.mean[i * N + v] = mn[i];
    }
}

/* ---- ordered sweep: MRF prior means + posterior means, in place, plane by plane ------------------------
 * One persistent kernel walks the hyper-planes with a grid-wide barrier between them (a launch per plane
 * costs ~13 us each, 64 % of the spatial step in the round-1 launch list). Neighbour means are read with
 * ld.global.cg: they were written by other SMs one barrier ago and must not be served from a stale L1 line.
 *
 * The step from one plane to the next is pure latency (766 planes at 256^3, a few thousand voxels each), so the
 * barrier is hand-made and SPLIT: a CTA is 480 voxel threads plus ONE EXTRA WARP that does nothing but the
 * barrier. After the voxel threads have stored a plane (bar.sync), that warp's leader fences and ARRIVES
 * (one atomic on a counter that only grows), the voxel threads meanwhile issue the next plane's
 * neighbour-independent loads (DRAM latency), and only then everyone WAITS for the leader to see all CTAs
 * arrived. With cooperative-groups' grid.sync() the fence sat in a warp that had those loads in flight and so
 * put a DRAM round trip on the critical path of every plane. The launch stays cooperative: every CTA must be
 * resident (one per SM). */
constexpr int SP_SWEEP_WORKERS = 480; /* 15 warps + the barrier warp = 16: four per SM sub-partition, 128 registers each */
constexpr int SP_SWEEP_BLOCK = SP_SWEEP_WORKERS + 32;

struct SweepBarrier
{
    unsigned *counter; /* zeroed by the host before the launch */
    unsigned target;
    FAB_DEV void arrive(unsigned n_ctas)
    {
        target += n_ctas;
        __threadfence(); /* release: this CTA's stores (ordered before us by bar.sync) before the arrival */
        atomicAdd(counter, 1u);
    }
    FAB_DEV void wait()
    {
        while (*(volatile unsigned *)counter < target)
        {
        }
        __threadfence(); /* acquire */
    }
};

template <int P, bool IGNORE, bool SLAB> struct SweepVoxel
{
    static constexpr int NT = NTri<P>::value;
    int v, nbr[6];
    double rhs[P], Sig[NT], L0[P];
    bool live;
    /* everything that does not depend on this sweep's updates: issued one barrier early */
    FAB_DEV void load_static(const SpArgs &s, int pos)
    {
        const VbArgs &a = s.v;
        const size_t N = (size_t)a.N;
        v = pos;
        FAB_CHECK_INDEX(a, pos, a.N, 101);
        live = a.status[pos] == 0;
        /* neighbours struck by IgnoreVoxel before this iteration are gone from the list (same count as
         * sp_theta used for the prior precision) */
#pragma unroll
        for (int j = 0; j < 6; j++)
        {
            nbr[j] = s.nn_idx[j * N + pos];
            FAB_CHECK(a, nbr[j] >= -1 && nbr[j] < a.N && nbr[j] != pos, 102);
            if (!nbr_alive<IGNORE>(s.status_prev, nbr[j]))
                nbr[j] = -1;
        }
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            rhs[i] = s.rhs[i * N + pos];
            L0[i] = s.L0[i * N + pos];
        }
#pragma unroll
        for (int i = 0; i < NT; i++)
            Sig[i] = a.cov[i * N + pos];
    }
    /* the part on the critical path: neighbour means (ld.global.cg - written by other SMs one barrier
     * ago, must not come from a stale L1 line), prior mean, posterior mean */
    FAB_DEV void finish(const SpArgs &s)
    {
        if (!live)
            return;
        const VbArgs &a = s.v;
        const size_t N = (size_t)a.N;
        int nn = 0;
#pragma unroll
        for (int j = 0; j < 6; j++)
            nn += nbr[j] >= 0;
        const int dims = s.spatial_dims;
#pragma unroll
        for (int k = 0; k < P; k++)
        {
            const char ty = a.params[k].prior_type;
            if (ty != 'M' && ty != 'm')
                continue;
            double contrib = 0.0;
#pragma unroll
            for (int j = 0; j < 6; j++)
                if (nbr[j] >= 0)
                    contrib += __ldcg(a.mean + k * N + nbr[j]);
            const int n1 = (ty == 'm') ? 2 * dims : nn;
            const double aK = s.aK[k];
            const double sp = (ty == 'M') ? aK * (n1 + 1e-8) : aK * n1;
            const double rec = 1 / double(n1);
            const double spatial_mean = contrib * rec;
            const double m0k = (1.0 / L0[k]) * sp * spatial_mean; /* priors.cc:469-470 */
            s.m0[k * N + v] = m0k;
            rhs[k] += L0[k] * m0k;
        }
        double mn[P];
        symv<P>(Sig, rhs, mn);
#pragma unroll
        for (int i = 0; i < P; i++)
            __stcg(a.mean + i * N + v, mn[i]);
        if (SLAB && s.link.up_pos && !s.link.forward_late)
        {
            /* top own plane of a z-slab: the slab above sweeps this voxel's +z neighbour one hyper-plane later
             * and must see THIS sweep's value - store it straight into that slab's lower ghost voxel */
            const int up = s.link.up_pos[v];
            FAB_CHECK(a, up >= -1 && up < s.link.up_N, 103);
            if (up >= 0)
            {
#pragma unroll
                for (int i = 0; i < P; i++)
                    s.link.up_mean[(size_t)i * s.link.up_N + up] = mn[i];
                /* no fence here: these stores are ordered before the flag the way NCCL orders its data before a
                 * step flag - the CTA barrier, the barrier warp's gpu-scope fence + arrival, and then ONE system
                 * fence and a st.release.sys by the thread that publishes the plane (release is cumulative). A
                 * __threadfence_system() in every forwarding thread put a system round trip on the critical path of
                 * every hyper-plane that touches the top plane: 13.2 -> see profiles/ ms per iteration on two GPUs */
            }
        }
    }
};

template <int P, bool IGNORE, bool SLAB>
__global__ void __launch_bounds__(SP_SWEEP_BLOCK, 1) sp_sweep_kernel(const __grid_constant__ SpArgs s)
{
    const bool worker = threadIdx.x < SP_SWEEP_WORKERS;
    const bool leader = threadIdx.x == SP_SWEEP_WORKERS; /* lane 0 of the barrier warp */
    const int stride = gridDim.x * SP_SWEEP_WORKERS, tid = blockIdx.x * SP_SWEEP_WORKERS + threadIdx.x;
    SweepBarrier bar = { s.sweep_barrier, 0u };
    const SlabLinks &lk = s.link;
    const bool slab = SLAB && lk.world > 1;
    const bool has_dn = slab && lk.rank > 0, has_up = slab && lk.rank + 1 < lk.world;
    const unsigned long long fwd_base = (unsigned long long)s.it * SLAB_IT_STRIDE;
    if (has_up)
    {
        /* the upper ghost plane must hold the slab above's values of the PREVIOUS iteration */
        if (leader)
            slab_wait(lk.flags + SLAB_FLAG_HI, (unsigned long long)s.it, lk.error);
        __syncthreads();
    }
    SweepVoxel<P, IGNORE, SLAB> cur;
    bool have = false;
    if (worker && s.plane_first < s.plane_last)
    {
        const int b = s.plane_starts[s.plane_first], e = s.plane_starts[s.plane_first + 1];
        if (b + tid < e)
        {
            cur.load_static(s, b + tid);
            have = true;
        }
    }
    /* how far the slab below is known to have got (its FWD flag as last read; uniform across the CTA). The flag is
     * read with a system-scope acquire - about a microsecond - so it is read only when the next hyper-plane is not
     * yet covered by what is known: the slab below starts own_z0 planes earlier and publishes in steps of
     * SLAB_PUBLISH_EVERY planes, so one read usually covers many planes. */
    __shared__ unsigned long long s_fwd_seen;
    unsigned long long fwd_known = 0;
    /* Late forwarding (lk.forward_late). A voxel thread that stores its fresh mean straight into the slab above puts
     * an NVLink round trip into this plane's barrier: the barrier warp's fence cannot complete before those remote
     * stores are acknowledged (about 2 us on each of the ~510 hyper-planes that touch the top plane - measured as
     * 6.1 -> 7.4 us per plane on two / eight GPUs against 4.9 us on one). Instead the 31 idle lanes of the barrier warp
     * copy the top-plane voxels of the PREVIOUS hyper-plane (already stored locally, visible after its barrier) into
     * the slab above WHILE the voxel threads work on the current one; the current plane's barrier then orders those
     * stores before the flag. The slab above runs one more hyper-plane behind - a constant lag. */
    const bool late = has_up && lk.forward_late != 0 && lk.up_pos != nullptr;
    const bool helper = threadIdx.x > SP_SWEEP_WORKERS;
    int pending = -1; /* hyper-plane finished on this GPU whose top-plane voxels are not forwarded yet (uniform) */
    auto forward_plane = [&](int q) {
        const size_t N = (size_t)s.v.N;
        const int pb = s.plane_starts[q], pe = s.plane_starts[q + 1];
        const int lanes = SP_SWEEP_BLOCK - SP_SWEEP_WORKERS - 1;
        for (int i = pb + blockIdx.x * lanes + (threadIdx.x - SP_SWEEP_WORKERS - 1); i < pe; i += gridDim.x * lanes)
        {
            const int up = lk.up_pos[i];
            FAB_CHECK(s.v, up >= -1 && up < lk.up_N, 107);
            if (up >= 0)
#pragma unroll
                for (int k = 0; k < P; k++)
                    lk.up_mean[(size_t)k * lk.up_N + up] = __ldcg(s.v.mean + k * N + i);
        }
    };
    for (int h = s.plane_first; h < s.plane_last; h++)
    {
        const int b = s.plane_starts[h], e = s.plane_starts[h + 1];
        FAB_CHECK(s.v, h >= 0 && h < s.n_planes && b >= 0 && b <= e && e <= s.v.N, 104);
        /* slab mode works on GLOBAL coordinates: local plane h IS hyper-plane x+y+z = h of the whole volume.
         * Its voxels on the bottom own plane (z = own_z0) have their -z neighbour in the slab below, on
         * hyper-plane h-1: wait until that slab has published it. */
        if (has_dn && e > b && h >= lk.own_z0 && h <= lk.own_z0 + lk.inplane_span
            && fwd_known < fwd_base + (unsigned long long)h)
        {
            if (leader)
                s_fwd_seen = slab_wait(lk.flags + SLAB_FLAG_FWD, fwd_base + (unsigned long long)h, lk.error);
            __syncthreads();
            fwd_known = s_fwd_seen; /* rewritten at the earliest after this plane's own CTA barrier */
        }
        if (late && pending >= 0 && helper && e > b)
            forward_plane(pending);
        if (worker)
        {
            if (have)
                cur.finish(s);
            for (int i = b + tid + stride; i < e; i += stride) /* planes wider than the grid */
            {
                SweepVoxel<P, IGNORE, SLAB> extra;
                extra.load_static(s, i);
                extra.finish(s);
            }
        }
        have = false;
        const bool sync = e > b; /* uniform across the grid: every thread skips the same empty planes */
        if (sync)
        {
            __syncthreads(); /* this CTA's stores of plane h are issued */
            if (leader)
                bar.arrive(gridDim.x);
        }
        if (worker && h + 1 < s.plane_last)
        {
            const int b2 = e, e2 = s.plane_starts[h + 2];
            if (b2 + tid < e2)
            {
                cur.load_static(s, b2 + tid); /* prefetch: in flight while the barrier completes */
                have = true;
            }
        }
        if (sync)
        {
            if (leader)
                bar.wait();
            __syncthreads();
        }
        /* hyper-plane h is done everywhere on this GPU: tell the slab above (only the planes that hold voxels of
         * the top own plane z = own_z1 - 1 matter to it) */
        if (has_up && !late && leader && blockIdx.x == 0 && h >= lk.own_z1 - 1 && h <= lk.own_z1 - 1 + lk.inplane_span
            && ((h - (lk.own_z1 - 1)) % SLAB_PUBLISH_EVERY == SLAB_PUBLISH_EVERY - 1 || h == lk.own_z1 - 1 + lk.inplane_span))
        {
            /* one release per SLAB_PUBLISH_EVERY hyper-planes: the slab above runs that many planes later (a constant
             * lag, tens of microseconds per sweep) and this warp pays the system-scope release a quarter as often.
             * st.release.sys is the fence (cumulative over everything ordered before it by the barriers above). */
            st_release_sys(lk.up_flags + SLAB_FLAG_FWD, fwd_base + (unsigned long long)h + 1);
        }
        if (late && sync)
        {
            /* this barrier also completed the helpers' forwarding of `pending` (their stores precede this plane's
             * CTA barrier, the barrier warp's fence and arrival): publish it, then this plane becomes pending */
            if (pending >= 0 && leader && blockIdx.x == 0
                && ((pending - (lk.own_z1 - 1)) % SLAB_PUBLISH_EVERY == SLAB_PUBLISH_EVERY - 1
                       || pending == lk.own_z1 - 1 + lk.inplane_span))
                st_release_sys(lk.up_flags + SLAB_FLAG_FWD, fwd_base + (unsigned long long)pending + 1);
            pending = (h >= lk.own_z1 - 1 && h <= lk.own_z1 - 1 + lk.inplane_span) ? h : -1;
        }
    }
    if (late && pending >= 0 && helper)
        forward_plane(pending); /* the last one: ordered before the final flag by the end-of-sweep barrier below */
    if (!slab)
        return;
    /* ---- end of the sweep ------------------------------------------------------------------------------
     * (a) no ordered sweep in this run (no 'M' / 'm' parameter): the means were written by sp_theta; the slab
     *     above still needs this iteration's values of our top plane for its aK sums */
    if (worker && has_up && s.plane_first >= s.plane_last)
    {
        const size_t N = (size_t)s.v.N;
        for (int v = tid; v < s.v.N; v += stride)
        {
            const int up = lk.up_pos[v];
            FAB_CHECK(s.v, up >= -1 && up < lk.up_N, 105);
            if (up >= 0)
                for (int i = 0; i < P; i++)
                    lk.up_mean[(size_t)i * lk.up_N + up] = __ldcg(s.v.mean + i * N + v);
        }
    }
    /* (b) our bottom own plane becomes the slab below's upper ghost - only now: during its own sweep that slab
     *     had to see last iteration's values, and it finished the planes that read them before we could finish
     *     ours (we waited for its flag plane by plane) */
    if (worker && has_dn)
    {
        const size_t N = (size_t)s.v.N;
        for (int j = tid; j < lk.n_dn; j += stride)
        {
            const int src = lk.dn_src[j], dst = lk.dn_dst[j];
            FAB_CHECK(s.v, src >= 0 && src < s.v.N && dst >= 0 && dst < lk.dn_N, 106);
            for (int i = 0; i < P; i++)
                lk.dn_mean[(size_t)i * lk.dn_N + dst] = __ldcg(s.v.mean + i * N + src);
        }
    }
    __threadfence_system();
    __syncthreads();
    if (leader)
    {
        bar.arrive(gridDim.x);
        bar.wait();
        if (blockIdx.x == 0)
        {
            __threadfence_system();
            if (has_up) /* everything up to the last hyper-plane is forwarded: releases the next iteration's waits */
                st_release_sys(lk.up_flags + SLAB_FLAG_FWD, fwd_base + SLAB_IT_STRIDE);
            if (has_dn)
                st_release_sys(lk.dn_flags + SLAB_FLAG_HI, (unsigned long long)s.it + 1);
        }
    }
}

/* ---- noise update, ReCentre and free energy (second loop of the reference, :675-722) --------------- */
template <class Model>
__global__ void __launch_bounds__(VB_BLOCK, FAB_MIN_BLOCKS) sp_noise_kernel(const __grid_constant__ SpArgs s)
{
    constexpr int P = Model::P;
    constexpr int NT = NTri<P>::value;
    const VbArgs &a = s.v;
    extern __shared__ double smem[];
    Model::stage(a, smem);
    volatile double *park = smem + Model::smem_bytes(a.T) / sizeof(double) + threadIdx.x;
    __syncthreads();
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= a.N)
        return;
    const size_t N = (size_t)a.N;
    /* the reference counts iterations once for the whole volume (m_ctx->it, inference_vb.cc:724): every
     * voxel, ignored ones included, reports the global count */
    if (a.iterations)
        a.iterations[v] = s.it + 1;
    if (a.status[v] != 0)
        return;
    const typename Model::Ctx mc = Model::make_ctx(a, smem);
    double m[P], Sig[NT];
    Stats<P> S[1];
    double nb, nc;
    {
        double d[P];
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            m[i] = a.mean[i * N + v];
            d[i] = s.centre[i * N + v] - m[i];
            S[0].b[i] = s.stats[(NT + i) * N + v];
        }
#pragma unroll
        for (int i = 0; i < NT; i++)
        {
            Sig[i] = a.cov[i * N + v];
            S[0].A[i] = s.stats[i * N + v];
        }
        S[0].rr = s.stats[(NT + P) * N + v];
        /* WhiteNoiseModel::UpdateNoise, noisemodel_white.cc:228-273 */
        double bd = 0.0;
#pragma unroll
        for (int j = 0; j < P; j++)
            bd += S[0].b[j] * d[j];
        const double kk = S[0].rr + 2.0 * bd + quadform<P>(S[0].A, d);
        const double tmp = kk + trace_prod<P>(Sig, S[0].A);
        nb = 1 / (tmp * 0.5 + 1 / a.noise_prior_b[0]);
        nc = ((double)a.n_per_phi[0] - 1) * 0.5 + a.noise_prior_c[0];
        if (a.locked_noise_stdev > 0)
            nb = 1 / nc / a.locked_noise_stdev / a.locked_noise_stdev;
        if (a.lock_centre)
            S[0].rr = kk; /* no re-centring below: F's k'Qk is taken about the locked centre */
    }
    a.noise[0 * N + v] = nb;
    a.noise[1 * N + v] = nc;
    if (!a.lock_centre) /* inference_vb.cc:695 */
    {
        /* park what F needs; the pass over the time-series only needs the new centre */
        {
            int k = 0;
#pragma unroll
            for (int i = 0; i < NT; i++)
                park[(k++) * VB_BLOCK] = Sig[i];
        }
        const int err = recentre_stats<Model, 1, true>(a, mc, nullptr, v, m, S);
        {
            int k = 0;
#pragma unroll
            for (int i = 0; i < NT; i++)
                Sig[i] = park[(k++) * VB_BLOCK];
        }
        if (err)
        {
            a.status[v] = err;
            return;
        }
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            s.centre[i * N + v] = m[i];
            s.stats[(NT + i) * N + v] = S[0].b[i];
        }
#pragma unroll
        for (int i = 0; i < NT; i++)
            s.stats[i * N + v] = S[0].A[i];
        s.stats[(NT + P) * N + v] = S[0].rr;
    }
    if (a.need_f)
    {
        double m0[P], L0[P], nbv[1] = { nb }, ncv[1] = { nc };
#pragma unroll
        for (int i = 0; i < P; i++)
        {
            m0[i] = s.m0[i * N + v];
            L0[i] = s.L0[i * N + v];
        }
        FCache<1> fc;
        fc.init(a);
        const double F = white_free_energy<P, 1>(a, S, m, Sig, s.logdet[v], m0, L0, nbv, ncv, fc) + *s.fprior_last;
        if (!finite_d(F))
        {
            a.status[v] = FABBER_VOX_NONFINITE_F;
            return;
        }
        if (a.free_energy)
            a.free_energy[v] = F;
        if (a.f_history && s.it < a.f_history_len)
            a.f_history[s.it * N + v] = F;
    }
}

/* ---- neighbour table and hyper-plane ordering (Vb::CalcNeighbours, inference_vb.cc:830-934) ---------
 * model independent: compiled once, in fabber_cuda.cu */
#ifdef FAB_SPATIAL_HOST_KERNELS
__global__ void sp_grid_kernel(const int *coords, int N, int nx, int ny, int nz, int *grid2vox, int *bad)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N)
        return;
    const int x = coords[v], y = coords[(size_t)N + v], z = coords[2 * (size_t)N + v];
    if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz)
    {
        atomicExch(bad, 1);
        return;
    }
    grid2vox[((size_t)z * ny + y) * nx + x] = v;
    /* CheckCoordMatrixCorrectlyOrdered (:769-793): x fastest, then y, then z, strictly increasing */
    if (v + 1 < N)
    {
        const int x2 = coords[v + 1], y2 = coords[(size_t)N + v + 1], z2 = coords[2 * (size_t)N + v + 1];
        const int sx = (x2 > x) - (x2 < x), sy = (y2 > y) - (y2 < y), sz = (z2 > z) - (z2 < z);
        if (sx + 10 * sy + 100 * sz <= 0)
            atomicExch(bad, 2);
    }
}

__global__ void sp_neighbour_kernel(const int *coords, int N, int nx, int ny, int nz, const int *grid2vox, int dims,
    int *nn_idx, int *plane_of, int *plane_hist)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= N)
        return;
    const int x = coords[v], y = coords[(size_t)N + v], z = coords[2 * (size_t)N + v];
    if (x < 0 || y < 0 || z < 0 || x >= nx || y >= ny || z >= nz)
        return; /* reported by sp_grid_kernel; never index plane_hist with it */
    const int dx[6] = { 1, -1, 0, 0, 0, 0 }, dy[6] = { 0, 0, 1, -1, 0, 0 }, dz[6] = { 0, 0, 0, 0, 1, -1 };
#pragma unroll
    for (int j = 0; j < 6; j++)
    {
        int id = -1;
        if (j < 2 * dims)
        {
            const int xx = x + dx[j], yy = y + dy[j], zz = z + dz[j];
            if (xx >= 0 && yy >= 0 && zz >= 0 && xx < nx && yy < ny && zz < nz)
                id = grid2vox[((size_t)zz * ny + yy) * nx + xx];
        }
        nn_idx[(size_t)j * N + v] = id;
    }
    const int h = x + y + z;
    plane_of[v] = h;
    atomicAdd(&plane_hist[h], 1);
}

#endif /* FAB_SPATIAL_HOST_KERNELS */

} // namespace fab
