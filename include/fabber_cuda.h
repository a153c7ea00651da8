/*
 * fabber_cuda.h - thin host <-> device C ABI underneath Vb::DoCalculations
 *
 * This is the *inner* drop-in boundary of the B200-native VB path: everything the reference does
 * between "voxel matrices are in RAM" and "resultMVNs / resultFs are filled"
 * (reference: inference_vb.cc:360-767, Vb::DoCalculations / DoCalculationsVoxelwise /
 * DoCalculationsSpatial) is one call into this ABI. Plain C structs, plain pointers and sizes;
 * no STL, no torch types, no exceptions. Errors are int return codes plus a per-voxel status word.
 *
 * All numerics are FP64 (NEWMAT::Real is double in the reference). The time-series itself is FP32
 * because every reference entry point stores FP32 (fabber_capi.h:128 `const float *data`,
 * rundata_newimage.cc:100 volume4D<float>) - the FP32 device copy is lossless.
 *
 * Layouts (all structure-of-arrays, voxel index fastest => coalesced voxel-per-thread access):
 *   data        float  [T][N]          row t = volume t (same as the C-API buffer restricted to mask;
 *                                      reference rundata_array.cc:100-133, Matrix T x Nvox)
 *   mean        double [P][N]          posterior means, Fabber space
 *   cov         double [P(P+1)/2][N]   posterior covariance, packed lower triangle by rows
 *                                      (1,1),(2,1),(2,2),(3,1).. = order of MVNDist::Save, dist_mvn.cc:419-432
 *   noise       double [NN][N]         white: (b_i, c_i) Gamma scale/shape per phi  -> NN = 2*n_phis
 *                                      ar1:   (b, c, a1_mean, a2_mean, a_prec11, a_prec21, a_prec22) -> NN = 7
 *   free_energy double [N]
 *   iterations  int    [N]             passes through the do{}while body (m_ctx->it, inference_vb.cc:499)
 *   status      int    [N]             FABBER_VOX_* code of the first numerical failure, 0 if none
 */
#ifndef FABBER_CUDA_H
#define FABBER_CUDA_H

#ifdef __cplusplus
extern "C" {
#endif

#define FABBER_CUDA_MAX_PARAMS 8
#define FABBER_CUDA_MAX_PHIS 4
#define FABBER_CUDA_AR_NOISE_FIELDS 7 /* AR(1), one echo: b c alpha1 alpha2 prec11 prec21 prec22 */
/* AR(1), two echoes (num-echoes=2): b1 c1 b2 c2, the nA alpha means, the alpha precisions as a packed lower
 * triangle (row-major: 11 21 22 31 ..). nA = 2 / 3 / 4 for ar1-cross-terms = none / same / dual
 * (Ar1cNoiseModel::NumAlphas, noisemodel_ar.cc:367-377) */
#define FABBER_CUDA_AR2_NOISE_FIELDS(nA) (4 + (nA) + (nA) * ((nA) + 1) / 2)
#define FABBER_CUDA_AR_MAX_NOISE_FIELDS 18

/* return codes */
#define FABBER_CUDA_OK 0
#define FABBER_CUDA_ERR_INVALID -1      /* bad argument / unsupported configuration */
#define FABBER_CUDA_ERR_CUDA -2         /* CUDA runtime error (message via fabber_cuda_last_error) */
#define FABBER_CUDA_ERR_BAD_VOXEL -3    /* numerical failure in a voxel and allow_bad_voxels == 0
                                           (reference: rethrow at inference_vb.cc:534,542) */

/* forward models with a __device__ Evaluate hook compiled in */
#define FABBER_MODEL_LINEAR 1 /* fwdmodel_linear.cc:92-96  */
#define FABBER_MODEL_POLY 2   /* fwdmodel_poly.cc:62-80    */
#define FABBER_MODEL_EXP 3    /* examples/fwdmodel_exp.cc:65-82 */
/* a model from a plug-in library (the --loadmodels mechanism, fwdmodel.cc:63-129): the kernels live in the
 * plug-in, fabber_cuda_model.plugin_launchers points at its launcher table (include/fabber_model_plugin.h) */
#define FABBER_MODEL_PLUGIN 100
#define FABBER_CUDA_MODEL_CONSTS 16

/* inference techniques */
#define FABBER_METHOD_VB 0   /* "vb" / "spatialvb" (inference_vb.cc) */
#define FABBER_METHOD_NLLS 1 /* "nlls" (inference_nlls.cc) */

/* noise models */
#define FABBER_NOISE_WHITE 0 /* noisemodel_white.cc */
#define FABBER_NOISE_AR1 1   /* noisemodel_ar.cc: n_phis = num-echoes (1 or 2), ar_cross_terms */
/* ar1-cross-terms (noisemodel_ar.cc:332,367-377); anything but NONE needs two echoes (:336-340) */
#define FABBER_AR_CROSS_NONE 0
#define FABBER_AR_CROSS_SAME 1
#define FABBER_AR_CROSS_DUAL 2

/* convergence detectors (registry names in setup.cc:49-57) */
#define FABBER_CONV_MAXITS 0    /* "maxits"        convergence.cc:43  */
#define FABBER_CONV_FCHANGE 1   /* "pointzeroone"  convergence.cc:86  */
#define FABBER_CONV_FREDUCE 2   /* "freduce"       convergence.cc:117 */
#define FABBER_CONV_TRIALMODE 3 /* "trialmode"     convergence.cc:162 */
#define FABBER_CONV_LM 4        /* "lm"            convergence.cc:278 */

/* per-voxel status codes (first failure wins) */
#define FABBER_VOX_OK 0
#define FABBER_VOX_NONFINITE_OFFSET 1   /* fwdmodel_linear.cc:134-140 */
#define FABBER_VOX_NONFINITE_JACOBIAN 2 /* fwdmodel_linear.cc:174-181 */
#define FABBER_VOX_NONFINITE_F 3        /* noisemodel_white.cc:445-451 */
#define FABBER_VOX_SINGULAR 4           /* NEWMAT exception from a matrix inverse */
#define FABBER_VOX_AR_NEG_VARIANCE 5    /* noisemodel_ar.cc:489-499 */
#define FABBER_VOX_IGNORED 6            /* spatial mode: voxel dropped by IgnoreVoxel, inference_vb.cc:266 */
/* OR-ed into the code when the failure happened in Vb::SetupPerVoxelDists (inference_vb.cc:235),
 * which the reference never catches: always fatal, even with allow_bad_voxels. */
#define FABBER_VOX_SETUP_FLAG 0x100

typedef struct fabber_cuda_model
{
    int id;       /* FABBER_MODEL_* */
    int n_params; /* P, 1..FABBER_CUDA_MAX_PARAMS */
    /* LINEAR: design matrix, HOST pointer, row-major [T][P] doubles (copied by the library) */
    const double *design;
    /* POLY */
    int poly_degree; /* P = degree + 1 */
    /* EXP */
    int exp_num;   /* number of exponentials, P = 2 * exp_num, params (amp_k, r_k) */
    double exp_dt; /* sample spacing */
    /* PLUGIN: launcher table of the plug-in's kernels (const fab::ModelLaunchers *), scalar constants the
     * plug-in's device hooks read, and optionally a vector of `design_len` doubles passed through `design`
     * (HOST pointer, copied by the library; e.g. a list of inversion times) */
    const void *plugin_launchers;
    double consts[FABBER_CUDA_MODEL_CONSTS];
    int design_len;
} fabber_cuda_model;

typedef struct fabber_cuda_param
{
    char transform;  /* 'I','L','S','F','A'  (transforms.h:20-24) */
    char prior_type; /* 'N','I','A','M','m','P','p' (priors.h) */
    char pad_[6];
    double prior_mean; /* Fabber space, after Transform::ToFabber (fwdmodel.cc:277) */
    double prior_prec; /* Fabber space precision = 1/ToFabberVar(var) */
    double prior_var;  /* Fabber space variance (used by ARD on the first iteration, priors.cc:164) */
    double post_mean;  /* MODEL space default initial posterior mean (fwdmodel.cc:302) */
    double post_var;   /* MODEL space default initial posterior variance (fwdmodel.cc:304) */
} fabber_cuda_param;

typedef struct fabber_cuda_vb_problem
{
    int n_voxels; /* N */
    int n_times;  /* T */
    fabber_cuda_model model;
    fabber_cuda_param params[FABBER_CUDA_MAX_PARAMS];

    /* noise */
    int noise_type; /* FABBER_NOISE_* */
    int n_phis;     /* white: number of distinct phis in noise-pattern (<= FABBER_CUDA_MAX_PHIS);
                       AR1: num-echoes, 1 or 2 (0 is read as 1). Two echoes: the series interleaves them,
                       TE1 TE2 TE1 TE2 .., n_times must be even (noisemodel_ar.cc:126-129) */
    const unsigned char *phi_pattern; /* HOST [T]: 0-based phi index per time point, NULL = all 0 */
    const unsigned char *time_masked; /* HOST [T]: 1 = masked time point (mt<n>), NULL = none */
    double noise_prior_b[FABBER_CUDA_MAX_PHIS]; /* Gamma scale of prior      (noisemodel_white.cc:144) */
    double noise_prior_c[FABBER_CUDA_MAX_PHIS]; /* Gamma shape of prior */
    double noise_post_b[FABBER_CUDA_MAX_PHIS];  /* initial posterior         (noisemodel_white.cc:148) */
    double noise_post_c[FABBER_CUDA_MAX_PHIS];
    double locked_noise_stdev; /* <= 0: off (noisemodel_white.cc:265) */
    double ar_alpha_prior_prec; /* AR1: prior/initial precision of alpha (1e-4, noisemodel_ar.cc:393) */
    int ar_cross_terms;         /* AR1: FABBER_AR_CROSS_* */

    /* convergence */
    int conv_type; /* FABBER_CONV_* */
    int max_iterations;
    double fchange; /* min-fchange (fchange/freduce/trialmode) or max-fchange (lm) */
    int max_trials;
    int need_f; /* compute free energy (m_needF, inference_vb.cc:242) */
    int f_history_len; /* rows available in buffers.f_history (0 = not recorded) */
    int allow_bad_voxels;

    /* spatial mode only (inference_vb.cc:578, priors.cc:183-488) */
    int spatial_dims;
    double spatial_speed; /* -1 = unlimited */
    double spatial_q1, spatial_q2;
    int update_first_iter;
    int nx, ny, nz; /* bounding grid of the coords, used for neighbour search */

    /* inference technique (setup.cc:28-33). FABBER_METHOD_NLLS (inference_nlls.cc): non-linear least squares per
     * voxel through the same entry points; noise / priors / convergence fields are ignored, the noise result array
     * (2 rows) is zeroed, free_energy and f_history are not written, `iterations` holds the optimiser's accepted
     * steps. */
    int method;          /* FABBER_METHOD_* */
    int nlls_lm;         /* --lm: Levenberg-Marquardt damping, else Levenberg (inference_nlls.cc:86,135-139) */
    int nlls_have_start; /* fwd-inital-posterior given: start from nlls_start (Fabber space) */
    double nlls_start[FABBER_CUDA_MAX_PARAMS];
} fabber_cuda_vb_problem;

typedef struct fabber_cuda_vb_buffers
{
    /* inputs (device pointers for fabber_cuda_*, host pointers for the test oracle) */
    const float *data;                                  /* [T][N] */
    const double *image_prior[FABBER_CUDA_MAX_PARAMS];  /* [N] for prior_type 'I', else NULL */
    const double *init_mean;  /* optional [P][N] Fabber-space restart means (continue-from-mvn) */
    const double *init_cov;   /* optional [P(P+1)/2][N] restart covariance, packed */
    const double *init_noise; /* optional [NN][N] restart noise posterior */
    /* optional [P][N] fixed linearisation centres (locked-linear-from-mvn, inference_vb.cc:171-178,227-231):
     * spatial runs linearise there once and never re-centre (:695); voxelwise runs ignore it, as the
     * reference does (:443 re-centres on the posterior unconditionally) */
    const double *lock_centre;
    const int *coords;        /* spatial: [3][N] integer voxel coordinates (x,y,z) */
    /* outputs */
    double *mean;        /* [P][N] */
    double *cov;         /* [P(P+1)/2][N] */
    double *noise;       /* [NN][N] */
    double *free_energy; /* [N] or NULL */
    double *f_history;   /* [f_history_len][N] or NULL */
    int *iterations;     /* [N] or NULL */
    int *status;         /* [N] */
    double *spatial_ak;  /* spatial: HOST [max_iterations+1][P] aK history or NULL */
} fabber_cuda_vb_buffers;

/* --- device management helpers (so a C/C++ host needs no CUDA headers) -------------------- */
int fabber_cuda_device_count(void);
/* debug build only (csrc: make checked, -DFAB_BOUNDS_CHECK): out[0] = index-check failures counted by the kernels
 * since the last call (all devices), out[1] = largest failing site code; resets the counters. Returns 1 when the
 * checks are compiled in, 0 when not (production build; out zeroed), < 0 on error. */
int fabber_cuda_check_report(unsigned long long *out);
int fabber_cuda_check_selftest(void); /* debug build: one deliberate failure, site code 999 */
int fabber_cuda_set_device(int dev);
int fabber_cuda_get_device(void); /* current device of the calling thread, < 0 on error */
const char *fabber_cuda_last_error(void);
void *fabber_cuda_malloc(unsigned long long bytes);
void fabber_cuda_free(void *dptr);
void *fabber_cuda_host_alloc(unsigned long long bytes); /* pinned */
void fabber_cuda_host_free(void *hptr);
int fabber_cuda_memcpy_h2d(void *dst, const void *src, unsigned long long bytes, void *stream);
int fabber_cuda_memcpy_d2h(void *dst, const void *src, unsigned long long bytes, void *stream);
int fabber_cuda_memset(void *dst, int value, unsigned long long bytes, void *stream);
int fabber_cuda_stream_sync(void *stream);

/* Gather masked voxels: full[t][i] (i over nx*ny*nz, x fastest) -> out[t][v] for v in index list
 * (reference rundata_array.cc:100-133 SetVoxelDataArray, done on the device). */
int fabber_cuda_gather_voxels(const float *full, unsigned long long n_grid, int n_times,
    const int *voxel_index, int n_voxels, float *out, void *stream);
/* Scatter back: out_full[r][i] = (float) in[r][v], 0 outside mask (rundata_array.cc:68-98). */
int fabber_cuda_scatter_voxels(const double *in, int n_rows, int n_voxels, const int *voxel_index,
    unsigned long long n_grid, float *out_full, void *stream);

/* --- the hot path --------------------------------------------------------------------------- */

/* Non-spatial VB: all iterations of every voxel fused in one launch.
 * Replaces Vb::SetupPerVoxelDists + Vb::DoCalculationsVoxelwise (inference_vb.cc:144-248,415-576).
 * `stream` is a cudaStream_t (NULL = default stream). Asynchronous: returns after enqueueing.
 * Returns FABBER_CUDA_OK / FABBER_CUDA_ERR_*. Per-voxel failures are reported in buffers.status;
 * use fabber_cuda_check_status() after synchronising to apply the halt-on-bad-voxel policy. */
int fabber_cuda_vb_voxelwise(
    const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf, void *stream);

/* The same on the voxels [v_begin, v_end) only; n_voxels stays the stride of every [field][N] array. Voxels
 * are independent in this mode (inference_vb.cc:423-571), so a host can upload the series in blocks of
 * voxels and start each block's calculation as soon as its data has arrived - upload and arithmetic
 * overlap. The helpers below are all that needs: a second (non-blocking) stream for the copies, events,
 * and a strided host -> device copy of one block of columns. */
int fabber_cuda_vb_voxelwise_range(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf,
    int v_begin, int v_end, void *stream);
void *fabber_cuda_stream_create(void);
void fabber_cuda_stream_destroy(void *stream);
void *fabber_cuda_event_create(void);
void fabber_cuda_event_destroy(void *event);
int fabber_cuda_event_record(void *event, void *stream);
int fabber_cuda_stream_wait_event(void *stream, void *event);
int fabber_cuda_memcpy2d_h2d(void *dst, unsigned long long dst_pitch, const void *src, unsigned long long src_pitch,
    unsigned long long width_bytes, unsigned long long rows, void *stream);
int fabber_cuda_memcpy2d_d2h(void *dst, unsigned long long dst_pitch, const void *src, unsigned long long src_pitch,
    unsigned long long width_bytes, unsigned long long rows, void *stream);
int fabber_cuda_event_sync(void *event);
/* 1 if `host_ptr` is page-locked memory known to the CUDA runtime (cudaMallocHost / cudaHostRegister: a
 * DMA engine can read it in place), 0 if it is ordinary pageable memory */
int fabber_cuda_host_is_pinned(const void *host_ptr);

/* Spatial VB (iteration-major; replaces Vb::DoCalculationsSpatial, inference_vb.cc:578-767,
 * SpatialPrior::CalculateaK / ApplyToMVN priors.cc:221-488, Vb::CalcNeighbours :830-964).
 * Synchronous. Single-GPU entry; multi-GPU z-slab sharding is driven from the host layer. */
int fabber_cuda_vb_spatial(
    const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf, void *stream);

/* Spatial VB on one z-slab of a volume that is partitioned over several GPUs (one process per GPU).
 *
 * The slab's voxel list holds its own voxels plus the GHOST planes just below and above it (their series
 * included, so their initial posterior is computed locally exactly as their owner computes it). Ghost
 * voxels are never updated locally; after every sweep the library packs the means of the slab's own
 * boundary planes, calls `exchange` and unpacks what the neighbours sent into the ghosts. The two global
 * sums of CalculateaK (priors.cc:233-343) go through `allreduce_sum`. Both callbacks receive DEVICE
 * pointers and the stream the work is queued on; the host implements them with its communication layer
 * (NCCL via torch.distributed in this repo) and returns 0 on success.
 *
 * The ordered sweep stays EXACT across slabs. The reference's sequential sweep means a voxel on the bottom
 * plane of slab r must see THIS iteration's value of the voxel below it (top plane of slab r-1), and last
 * iteration's value of the voxel above the slab's top plane. The hyper-planes x+y+z = H (global z) are cut
 * into `n_blocks` blocks of `block_planes` planes; at step s rank r sweeps block s - r, then `forward`s the
 * freshly updated means of its top plane that lie in that block to rank r+1, which sweeps the same block
 * one step later - a software pipeline with one block of skew, n_blocks + world - 1 steps per iteration.
 * After the sweep `exchange` refreshes all ghosts (the downward direction must wait for the sweep to end). */
#define FABBER_VOX_GHOST 0x200 /* status of a ghost voxel in a slab run */
typedef struct fabber_cuda_slab
{
    int n_global_voxels;        /* voxels of the whole volume (hK = N/2 + q2) */
    const unsigned char *ghost; /* DEVICE [N]: 1 = ghost voxel */
    /* DEVICE index lists (positions in this slab's voxel list): own boundary planes to send down / up,
     * ghost planes that receive from below / above. Counts may be 0 at the ends of the volume. */
    const int *send_lo, *send_hi, *recv_lo, *recv_hi;
    int n_send_lo, n_send_hi, n_recv_lo, n_recv_hi;
    /* pipelined sweep: global plane H = local plane + plane_offset (plane_offset = z of the first local
     * plane). fwd_send / fwd_recv: positions (this slab's voxel list) of the own top plane / the lower ghost
     * plane, grouped by block: block b is [fwd_*_start[b], fwd_*_start[b+1]). DEVICE lists, HOST starts. */
    int rank, world, n_blocks, block_planes, plane_offset;
    const int *fwd_send, *fwd_recv;
    const int *fwd_send_start, *fwd_recv_start; /* HOST [n_blocks + 1] */
    void *user;
    int (*allreduce_sum)(void *user, double *dev_values, int n, void *stream);
    /* buffers are [P][count] doubles */
    int (*exchange)(void *user, const double *send_lo, int n_send_lo, const double *send_hi, int n_send_hi,
        double *recv_lo, int n_recv_lo, double *recv_hi, int n_recv_hi, void *stream);
    /* step s of the pipelined sweep: send n_send values up to rank+1, receive n_recv from rank-1 */
    int (*forward)(void *user, int step, const double *send, int n_send, double *recv, int n_recv, void *stream);
} fabber_cuda_slab;
int fabber_cuda_vb_spatial_slab(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf,
    const fabber_cuda_slab *slab, void *stream);

/* Spatial VB of ONE volume over several GPUs driven by ONE process: the voxel list is cut into z-slabs, part r
 * holds its own voxels [own0, own1) of the list plus the ghost planes just below and above (the whole range
 * [v0, v1), series included), on device `device`. The slabs' kernels are coupled through each other's memory
 * (peer access over NVLink): the ordered sweep forwards its top plane into the next slab's ghosts hyper-plane by
 * hyper-plane behind release/acquire flags, the bottom plane goes down after the sweep, and the two aK sums
 * (priors.cc:233-343) are all-gathered through mailboxes inside the aK kernel - no host callback, one
 * cooperative sweep launch per iteration per GPU, and the same result as the one-GPU run (the aK sums differ
 * in summation order only). Replaces Vb::DoCalculationsSpatial (inference_vb.cc:578-767) for the whole box.
 *   prob      n_voxels = voxels of the WHOLE volume; nx, ny, nz the whole grid
 *   parts[r]  buf: DEVICE pointers on parts[r].device, every per-voxel array of v1 - v0 voxels (data
 *             [T][v1-v0], coords [3][v1-v0] with GLOBAL coordinates, outputs, optional image priors / restart
 *             arrays); own_z0 / own_z1: the z-planes this part owns. Parts tile the list in order; part r's
 *             ghosts are part r-1's last and part r+1's first plane. Outputs of ghost voxels are not
 *             meaningful (status FABBER_VOX_GHOST).
 * Synchronous. Several parts may name the same device (tests on one GPU). */
typedef struct fabber_cuda_slab_part
{
    int device;
    int v0, v1, own0, own1;
    int own_z0, own_z1;
    fabber_cuda_vb_buffers buf;
} fabber_cuda_slab_part;
int fabber_cuda_vb_spatial_multi(const fabber_cuda_vb_problem *prob, int n_parts, const fabber_cuda_slab_part *parts);
/* device-side duration of the last fabber_cuda_vb_spatial_multi call with more than one part: CUDA events on every
 * slab's stream around set-up, iterations and result permutation, maximum over the devices, in ms */
double fabber_cuda_last_multi_ms(void);

/* Scan status[] on the device; returns 0 if all OK, else the number of failed voxels and the
 * index / code of the first one (synchronises the stream). */
int fabber_cuda_check_status(
    const int *status, int n_voxels, int *first_bad_voxel, int *first_bad_code, void *stream);

/* Batched model evaluation at Fabber-space means: fit[t][v] = g(ToModel(mean[:,v]))
 * (reference InferenceTechnique::SaveResults inference.cc:190-191 EvaluateFabber per voxel). */
int fabber_cuda_model_fit(const fabber_cuda_vb_problem *prob, const double *mean /*[P][N]*/,
    double *fit /*[T][N]*/, void *stream);

/* Output maps of InferenceTechnique::SaveResults / Vb::SaveResults (inference.cc:112-157,
 * inference_vb.cc:966-1047, MVNDist::Save dist_mvn.cc:377-433) computed on the device from the result
 * arrays of fabber_cuda_vb_*, in float32 - the precision every reference output is stored in
 * (rundata_array.cc:68-98, NIfTI float). All pointers are DEVICE pointers; NULL = not wanted. */
typedef struct fabber_cuda_vb_outputs
{
    float *mean, *std, *zstat, *var; /* [P][N], model space (FwdModel::ToModel, fwdmodel.cc:326-337) */
    float *noise_mean, *noise_std;   /* [Nn][N]; Nn = n_phis (white) or n_alphas + num-echoes (AR1: the alphas, then the phis) */
    float *final_mvn;                /* [(P+Nn)(P+Nn+1)/2 + (P+Nn) + 1][N] packed covariance, means, 1 */
    float *free_energy;              /* [N] */
    float *f_history;                /* [f_history_rows][N]: rows >= iterations repeat the final F */
    int f_history_rows;
    const float *data;               /* [T][N] input series, needed for residuals */
    float *model_fit, *residuals;    /* [T][N] (inference.cc:160-239) */
} fabber_cuda_vb_outputs;
int fabber_cuda_vb_save_results(const fabber_cuda_vb_problem *prob, const fabber_cuda_vb_buffers *buf,
    const fabber_cuda_vb_outputs *out, void *stream);

/* max over an int array on the device (synchronises the stream); used for the F-history row count */
int fabber_cuda_max_int(const int *values, int n, int *result, void *stream);

/* Measured FP64 FMA throughput of the current device in GFLOP/s (dependent-free DFMA loop on all
 * SMs; used as the roofline denominator because MEASURED_PEAKS.json carries no FP64 figure). */
double fabber_cuda_measure_fp64_peak(int repeats);

/* Number of kernels this library has launched so far in this process. */
unsigned long long fabber_cuda_launch_count(void);

/* Accuracy probe: fast[i] = the kernels' table-based exp(x[i]) (csrc/vb_exp.cuh, valid for |x| < 708),
 * ref[i] = the CUDA library's exp(x[i]). Device pointers. */
int fabber_cuda_exp_probe(const double *x, double *fast, double *ref, int n, void *stream);

/* sizeof() of the two structs above as compiled, so language bindings can verify their mirror. */
int fabber_cuda_sizeof_problem(void);
int fabber_cuda_sizeof_buffers(void);

#ifdef __cplusplus
}
#endif
#endif
