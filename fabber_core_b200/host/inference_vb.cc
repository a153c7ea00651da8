/*
 * inference_vb.cc - Vb: the VB inference technique, host side.
 *
 *   Initialize      inference_vb.cc:100-130 + inference.cc:62-110 (noise model by name, F flags, masked
 *                   time points, halt-on-bad-voxel policy)
 *   DoCalculations  inference_vb.cc:360-413. Everything the reference does between "voxel matrices are in
 *                   RAM" and "resultMVNs / resultFs are filled" is ONE call into the device C ABI
 *                   (fabber_cuda_vb_voxelwise / fabber_cuda_vb_spatial, include/fabber_cuda.h): this
 *                   function only translates options into the plain-C problem description, moves the
 *                   float32 series to the GPU and brings the structure-of-arrays results back.
 *   SaveResults     inference_vb.cc:966-1051 + inference.cc:112-252 + dist_mvn.cc:377-433 (finalMVN).
 */
#include <algorithm>
#include <cmath>
#include <cstring>

#include <chrono>

#include "fabber_host.h"
#include "operators.h"

namespace fabber_b200
{
namespace
{
struct StopWatch
{
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    double lap_ms()
    {
        std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
        return ms;
    }
};
/* RAII for device allocations */
struct DeviceBuf
{
    void *p = nullptr;
    explicit DeviceBuf(size_t bytes)
    {
        p = fabber_cuda_malloc(bytes);
        if (!p)
            throw FabberInternalError(std::string("GPU allocation failed: ") + fabber_cuda_last_error());
    }
    ~DeviceBuf() { fabber_cuda_free(p); }
    DeviceBuf(const DeviceBuf &) = delete;
    DeviceBuf &operator=(const DeviceBuf &) = delete;
};
void check(int rc, const char *what)
{
    if (rc != FABBER_CUDA_OK)
        throw FabberInternalError(std::string(what) + ": " + fabber_cuda_last_error());
}
int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

/* noise-initial-prior / noise-initial-posterior (Vb::InitializeNoiseFromParam, inference_vb.cc:132-142): an MVN
 * matrix file [covariance means(:); means(:)' 1.0] (MVNDist::LoadFromMatrix, dist_mvn.cc:287-309); each phi's
 * Gamma is moment-matched to its mean and variance (WhiteParams::InputFromMVN, noisemodel_white.cc:70-79;
 * GammaDist::SetMeanVariance, dist_gamma.cc:29-33: b = v / m, c = m / b). */
void NoiseGammasFromMvnFile(FabberRunData &rundata, const std::string &key, int nphis, double *b, double *c)
{
    const std::string filename = rundata.GetStringDefault(key, "modeldefault");
    if (filename == "modeldefault")
        return;
    std::vector<double> m;
    int rows = 0, cols = 0;
    read_matrix_file(filename, m, rows, cols);
    const int n = rows - 1;
    bool ok = n >= 1 && rows == cols && m[(size_t)n * cols + n] == 1.0;
    for (int i = 0; ok && i < rows; i++)
        for (int j = 0; j < i; j++)
            ok = ok && m[(size_t)i * cols + j] == m[(size_t)j * cols + i];
    if (!ok)
        throw InvalidOptionValue(filename, "",
            "MVNs must be symmetric matrices (format = [covariance means(:); means(:) 1.0])");
    if (n < nphis)
        throw InvalidOptionValue(key, filename, "MVN has fewer rows than the noise model has precisions");
    for (int i = 0; i < nphis; i++)
    {
        for (int j = i + 1; j < n; j++)
            if (m[(size_t)i * cols + j] != 0.0)
                throw FabberRunDataError("Phis should have zero covariance!");
        const double mean = m[(size_t)i * cols + n], var = m[(size_t)i * cols + i];
        b[i] = var / mean;
        c[i] = mean / b[i];
    }
}
} // namespace

std::vector<std::string> Vb::GetKnownMethods()
{
    std::vector<std::string> k;
    k.push_back("nlls");
    k.push_back("spatialvb");
    k.push_back("vb");
    return k;
}
std::string Vb::GetDescription() { return "Variational Bayes inference technique (B200 GPU implementation)"; }
std::string Vb::GetDescription(const std::string &method)
{
    if (method == "nlls") /* inference_nlls.cc:48-51 */
        return "Non-linear least squares inference technique (B200 GPU implementation).";
    return GetDescription();
}
void Vb::GetOptions(std::vector<OptionSpec> &opts, const std::string &method)
{
    if (method != "nlls")
        return GetOptions(opts);
    /* inference_nlls.cc:31-36 (the reference lists only the first: NUM_OPTIONS = 1) */
    static const OptionSpec O[] = {
        { "vb-init", OPT_BOOL, "Whether NLLS is being run in isolation or as a pre-step for VB", true, "" },
        { "lm", OPT_BOOL, "Whether to use LM convergence (default is L)", true, "" },
    };
    const size_t listed = 1; /* the reference's table has both, its NUM_OPTIONS lists the first (--lm works all the same) */
    for (size_t i = 0; i < listed; i++)
        opts.push_back(O[i]);
}
void Vb::GetOptions(std::vector<OptionSpec> &opts)
{
    /* inference_vb.cc:31-76 */
    static const OptionSpec O[] = {
        { "noise", OPT_STR, "Noise model to use (white or ar1)", false, "" },
        { "convergence", OPT_STR, "Name of method for detecting convergence", true, "maxits" },
        { "max-iterations", OPT_STR, "number of iterations of VB to use with the maxits convergence detector", true,
            "10" },
        { "min-fchange", OPT_STR, "When using the fchange convergence detector, the change in F to stop at", true,
            "10" },
        { "max-trials", OPT_STR,
            "When using the trial mode convergence detector, the maximum number of trials after an initial reduction "
            "in F",
            true, "10" },
        { "print-free-energy", OPT_BOOL, "Output the free energy", true, "" },
        { "continue-from-mvn", OPT_MVN, "Continue previous run from output MVN files", true, "" },
        { "locked-linear-from-mvn", OPT_MVN, "MVN file containing fixed centres for linearization", true, "" },
        { "output-only", OPT_BOOL,
            "Skip model fitting, just output requested data based on supplied MVN. Can only be used with "
            "continue-from-mvn",
            true, "" },
        { "noise-initial-prior", OPT_MATRIX, "MVN of initial noise prior", true, "" },
        { "noise-initial-posterior", OPT_MATRIX, "MVN of initial noise posterior", true, "" },
        { "noise-pattern", OPT_STR,
            "repeating pattern of noise variances for each point (e.g. 12 gives odd and even data points different "
            "variances)",
            true, "1" },
        { "PSP_byname<n>", OPT_STR, "Name of model parameter to use image prior", true, "" },
        { "PSP_byname<n>_type", OPT_STR, "Type of image prior to use for parameter <n> - I=image prior", true, "" },
        { "PSP_byname<n>_image", OPT_IMAGE, "Image prior for parameter <n>", true, "" },
        { "PSP_byname<n>_prec", OPT_FLOAT, "Precision to apply to image prior for parameter <n>", true, "" },
        { "PSP_byname<n>_transform", OPT_STR, "Transform to apply to parameter <n>", true, "" },
        { "allow-bad-voxels", OPT_BOOL, "Continue if numerical error found in a voxel, rather than stopping", true,
            "" },
        { "mcsteps", OPT_INT, "Number of motion correction steps (motion correction is compiled out of the reference too)",
            true, "0" },
        { "distance-measure", OPT_STR, "", true, "dist1" },
        { "ar1-cross-terms", OPT_STR, "For AR1 noise, type of cross-linking (dual, same or none)", true, "dual" },
        { "spatial-dims", OPT_INT, "Number of spatial dimensions", true, "3" },
        { "spatial-speed", OPT_STR, "Restrict speed of spatial smoothing", true, "-1" },
        { "param-spatial-priors", OPT_STR,
            "Type of spatial priors for each parameter, as a sequence of characters. N=nonspatial, M=Markov random "
            "field, P=Penny, A=ARD",
            true, "N+" },
        { "update-spatial-prior-on-first-iteration", OPT_BOOL, "", true, "" },
    };
    for (size_t i = 0; i < sizeof(O) / sizeof(O[0]); i++)
        opts.push_back(O[i]);
}

void Vb::Initialize(FwdModel *model, FabberRunData &rundata)
{
    m_model = model;
    std::vector<Parameter> params;
    m_model->GetParameters(rundata, params);
    m_num_params = (int)params.size();
    if (m_num_params < 1 || m_num_params > FABBER_CUDA_MAX_PARAMS)
        throw InvalidOptionValue("model", stringify(m_num_params),
            "number of parameters outside the compiled device hooks (1.." + stringify(FABBER_CUDA_MAX_PARAMS) + ")");
    m_halt_bad_voxel = !rundata.GetBool("allow-bad-voxels"); /* inference.cc:93 */

    /* --method=nlls (NLLSInferenceTechnique::Initialize, inference_nlls.cc:57-88): no noise model, no priors, no
     * free energy; the same device plumbing (block-wise upload, one voxel range per GPU, SaveResults) */
    m_nlls = rundata.GetString("method") == "nlls";
    if (m_nlls)
    {
        m_ar = false;
        m_saveF = m_saveFsHistory = m_printF = false;
        rundata.GetBool("vb-init"); /* only changes what the reference logs for an ill-conditioned voxel */
        m_nlls_lm = rundata.GetBool("lm");
        m_nlls_start.clear();
        const std::string file = rundata.GetStringDefault("fwd-inital-posterior", "modeldefault"); /* sic */
        if (file != "modeldefault")
        {
            /* MVNDist::LoadFromMatrix: [covariance means(:); means(:) 1.0] (dist_mvn.cc:287-309) */
            std::vector<double> m;
            int rows = 0, cols = 0;
            read_matrix_file(file, m, rows, cols);
            if (rows != cols || rows != m_num_params + 1 || m[(size_t)(rows - 1) * cols + rows - 1] != 1.0)
                throw InvalidOptionValue("fwd-inital-posterior", file,
                    "MVNs must be symmetric matrices (format = [covariance means(:); means(:) 1.0]) of the model's size");
            for (int i = 0; i < m_num_params; i++)
                m_nlls_start.push_back(m[(size_t)i * cols + rows - 1]);
        }
        return;
    }
    const std::string noise = rundata.GetString("noise");
    if (noise == "white")
        m_ar = false;
    else if (noise == "ar")
        m_ar = true;
    else
        throw InvalidOptionValue("noise", noise, "Unrecognized noise type");
    m_saveF = rundata.GetBool("save-free-energy");
    m_saveFsHistory = rundata.GetBool("save-free-energy-history");
    m_printF = rundata.GetBool("print-free-energy");
}

bool Vb::IsSpatial(FabberRunData &rundata, const std::vector<Parameter> &params) const
{
    if (rundata.GetString("method") == "spatialvb")
        return true;
    for (size_t i = 0; i < params.size(); i++)
        if (std::string("MmPp").find(params[i].prior_type) != std::string::npos)
            return true;
    return false;
}

void Vb::ReleaseDevice()
{
    for (size_t g = 0; g < m_ctx.size(); g++)
    {
        DevCtx &c = m_ctx[g];
        DeviceScope scope(c.device);
        DevArray *all[] = { &c.mean, &c.cov, &c.noise, &c.F, &c.hist, &c.its, &c.status };
        bool any = false;
        for (DevArray *d : all)
            any = any || d->p;
        if (any)
            fabber_cuda_stream_sync(nullptr);
        for (DevArray *d : all)
        {
            cached_device_free(d->p, d->bytes);
            d->p = nullptr;
            d->bytes = 0;
        }
    }
    m_ctx.clear();
}

void Vb::DoCalculations(FabberRunData &rundata)
{
    VoxelData &data = rundata.MutableMainVoxelData();
    Prepare(rundata, data);
    if (m_nvoxels == 0 || m_output_only)
        return;
    LaunchAll(data);
    Finish(rundata);
}

/* Everything of DoCalculations up to the first kernel launch: options -> plain-C problem description, the
 * devices of the run, result arrays, per-voxel inputs on the devices. Split off so that the host library can
 * run it while the series is still being staged (FabberRunData::SetVoxelDataArray starts each block's kernel
 * the moment that block is on its way - see "speculative start" there). */
void Vb::Prepare(FabberRunData &rundata, VoxelData &data)
{
    StopWatch sw;
    m_output_only = false;
    m_launch_rc = FABBER_CUDA_OK;
    const size_t N = data.cols;
    const int T = data.rows;
    m_nvoxels = N;
    m_ntimes = T;
    std::vector<Parameter> params;
    m_model->GetParameters(rundata, params);
    const int P = m_num_params, NT = P * (P + 1) / 2;

    fabber_cuda_vb_problem &prob = m_prob;
    memset(&prob, 0, sizeof(prob));
    prob.n_voxels = (int)N;
    prob.n_times = T;
    m_model->GetDeviceModel(prob.model);
    if (prob.model.n_params != P)
        throw FabberInternalError("device model hook disagrees with GetParameterDefaults about the number of parameters");
    for (int i = 0; i < P; i++)
    {
        fabber_cuda_param &q = prob.params[i];
        q.transform = params[i].transform;
        q.prior_type = params[i].prior_type == '-' ? 'N' : params[i].prior_type;
        q.prior_mean = params[i].prior.mean();
        q.prior_var = params[i].prior.var();
        q.prior_prec = params[i].prior.prec();
        q.post_mean = params[i].post.mean();
        q.post_var = params[i].post.var();
    }

    /* ---- noise model, convergence detector, priors: the reference's operator classes (host/operators.h) know
     * their options, their hard-coded initial distributions and their device codes ----------------------- */
    m_masked.assign(T, 0);
    const std::vector<int> mt = rundata.GetIntList("mt", 1); /* inference.cc:96-103 */
    for (size_t i = 0; i < mt.size(); i++)
    {
        if (mt[i] > T)
            throw InvalidOptionValue("mt", stringify(mt[i]), "Masked time point beyond the end of the data");
        m_masked[mt[i] - 1] = 1;
    }
    if (m_nlls)
    {
        prob.method = FABBER_METHOD_NLLS;
        prob.nlls_lm = m_nlls_lm ? 1 : 0;
        prob.nlls_have_start = m_nlls_start.empty() ? 0 : 1;
        for (size_t i = 0; i < m_nlls_start.size(); i++)
            prob.nlls_start[i] = m_nlls_start[i];
        prob.noise_type = FABBER_NOISE_WHITE; /* unused by the NLLS kernel; keeps the shared checks quiet */
        prob.n_phis = 1;
        m_pattern.assign(T, 0);
        prob.noise_prior_b[0] = prob.noise_post_b[0] = 1;
        prob.noise_prior_c[0] = prob.noise_post_c[0] = 1;
        prob.locked_noise_stdev = -1;
        prob.conv_type = FABBER_CONV_MAXITS;
        prob.max_iterations = 1;
        prob.max_trials = 1;
        m_nphis = 1;
        m_noise_params = 0; /* InferenceTechnique::m_noise_params stays 0: the result MVN is the model's alone */
        prob.phi_pattern = m_pattern.data();
        prob.time_masked = mt.empty() ? nullptr : m_masked.data();
    }
    else
    {
        if (m_ar && !mt.empty())
            throw InvalidOptionValue("noise", "ar", "AR noise model does not support masked time points");
        std::unique_ptr<NoiseModel> noise(NoiseModel::NewFromName(rundata.GetString("noise")));
        noise->Initialize(rundata);
        noise->Describe(prob, T, m_pattern);
        if (m_ar)
        {
            /* noisemodel_ar.cc:305-349 */
            for (const char *key : { "noise-initial-prior", "noise-initial-posterior" })
                if (rundata.GetStringDefault(key, "modeldefault") != "modeldefault")
                    throw InvalidOptionValue(key, rundata.GetString(key),
                        "the AR(1) device kernel starts alpha at N(0, 1e4 I) only: not supported with noise=ar");
            m_nphis = prob.n_phis; /* num-echoes */
            m_nalphas = 2 + prob.ar_cross_terms;
            m_noise_params = m_nalphas + m_nphis; /* the alphas, then the phis: Ar1cParams::OutputAsMVN, noisemodel_ar.cc:287-300 */
        }
        else
        {
            m_nphis = prob.n_phis;
            m_noise_params = m_nphis;
            /* after the hard-coded values, as in the reference (inference_vb.cc:204-205) */
            NoiseGammasFromMvnFile(rundata, "noise-initial-prior", m_nphis, prob.noise_prior_b, prob.noise_prior_c);
            NoiseGammasFromMvnFile(rundata, "noise-initial-posterior", m_nphis, prob.noise_post_b, prob.noise_post_c);
        }
        prob.phi_pattern = m_pattern.data();
        prob.time_masked = mt.empty() ? nullptr : m_masked.data();

        /* convergence (setup.cc:49-57, convergence.cc Initialize functions) */
        std::unique_ptr<ConvergenceDetector> detector(
            ConvergenceDetector::NewFromName(rundata.GetStringDefault("convergence", "maxits")));
        detector->Initialize(rundata);
        detector->Describe(prob);
        /* (max-trials is validated by the trial-mode detector only, convergence.cc:147-149) */
        rundata.GetIntDefault("spatial-dims", 3, 0, 3); /* range-checked here first, as Vb::Initialize does (inference_vb.cc:119) */
        /* priors (priors.cc:490-528): the factory validates every parameter's prior type and options (image data
         * present, spatial-dims / speed in range); the types themselves travel in prob.params[] */
        {
            std::vector<Prior *> priors = PriorFactory(rundata).CreatePriors(params);
            for (size_t i = 0; i < priors.size(); i++)
                delete priors[i];
        }
    }
    /* locked-linear-from-mvn is loaded by every Vb run, spatial or not (inference_vb.cc:171-178) - a missing data set
     * is an error even where the centres are then not used (only the spatial method honours them, :695 vs :443) */
    if (!m_nlls && rundata.GetStringDefault("locked-linear-from-mvn", "") != "")
        rundata.GetVoxelData(rundata.GetString("locked-linear-from-mvn"));
    const bool spatial = !m_nlls && IsSpatial(rundata, params);
    const bool useF = !spatial && prob.conv_type != FABBER_CONV_MAXITS;
    m_needF = !m_nlls && (useF || m_printF || m_saveF || m_saveFsHistory); /* inference_vb.cc:242 */
    prob.need_f = m_needF ? 1 : 0;
    prob.allow_bad_voxels = m_halt_bad_voxel ? 0 : 1;
    /* F history: worst case per detector - lm never runs more than (max_its + 1) * 14 passes */
    m_fhist_len = 0;
    if (m_saveFsHistory)
        m_fhist_len = spatial ? prob.max_iterations
                              : (prob.conv_type == FABBER_CONV_LM ? (prob.max_iterations + 1) * 14
                                      : prob.conv_type == FABBER_CONV_TRIALMODE
                                      ? (prob.max_iterations + 2) * (prob.max_trials + 1)
                                      : prob.max_iterations + 1);
    prob.f_history_len = m_fhist_len;
    prob.spatial_dims = rundata.GetIntDefault("spatial-dims", 3, 0, 3);
    prob.spatial_speed = rundata.GetDoubleDefault("spatial-speed", -1);
    prob.spatial_q1 = rundata.GetDoubleDefault("spatial-q1", 10.0);
    prob.spatial_q2 = rundata.GetDoubleDefault("spatial-q2", 1.0);
    prob.update_first_iter = rundata.GetBool("update-spatial-prior-on-first-iteration") ? 1 : 0;
    prob.nx = rundata.Extent()[0];
    prob.ny = rundata.Extent()[1];
    prob.nz = rundata.Extent()[2];
    m_nn = !m_ar ? 2 * m_nphis : m_nphis == 2 ? FABBER_CUDA_AR2_NOISE_FIELDS(m_nalphas) : FABBER_CUDA_AR_NOISE_FIELDS;
    const int NN = m_nn;
    m_spatial = spatial;
    if (N == 0)
        return; /* zero voxels is not an error (test/test_inference.cc:57-73) */

    /* ---- the devices of this run: one contiguous voxel range per GPU (the series is normally already there,
     * dealt out while it was being set); spatial VB couples the voxels and runs on one device ------------- */
    ReleaseDevice();
    bool slabs = spatial && data.parts.size() > 1; /* z-slabs dealt out at set_data: the multi-device engine */
    FreeScratch();
    for (size_t g = 0; g < data.parts.size(); g++)
        slabs = slabs && data.parts[g].z1 > data.parts[g].z0;
    if (data.parts.empty() || (spatial && data.parts.size() != 1 && !slabs))
    {
        const std::vector<int> &devs = run_devices();
        if (devs.empty())
            throw FabberInternalError(std::string("no CUDA device: ") + fabber_cuda_last_error());
        data.upload_whole(devs[0]);
    }
    for (size_t g = 0; g < data.parts.size(); g++)
    {
        DevCtx c;
        c.device = data.parts[g].device;
        c.v0 = data.parts[g].v0;
        c.v1 = data.parts[g].v1;
        c.own0 = data.parts[g].own0;
        c.own1 = data.parts[g].own1;
        c.data = data.parts[g].dev;
        m_ctx.push_back(c);
    }
    /* device result arrays (cached blocks: no cudaMalloc on the steady-state path) */
    for (size_t g = 0; g < m_ctx.size(); g++)
    {
        DevCtx &c = m_ctx[g];
        DeviceScope scope(c.device);
        const size_t n = c.v1 - c.v0;
        auto dev = [](DevArray &d, size_t bytes) {
            d.bytes = bytes;
            d.p = cached_device_alloc(bytes);
        };
        dev(c.mean, (size_t)P * n * sizeof(double));
        dev(c.cov, (size_t)NT * n * sizeof(double));
        dev(c.noise, (size_t)NN * n * sizeof(double));
        dev(c.F, n * sizeof(double));
        dev(c.its, n * sizeof(int));
        dev(c.status, n * sizeof(int));
        if (m_fhist_len > 0)
            dev(c.hist, (size_t)m_fhist_len * n * sizeof(double));
    }

    /* ---- restart / output-only (inference_vb.cc:181-216, 385-389) ------------------------------------ */
    std::vector<double> init_mean, init_cov, init_noise;
    bool continue_from_mvn = false;
    try
    {
        if (m_nlls) /* continue-from-mvn / output-only are Vb options (inference_vb.cc:181): NLLS never reads them */
            throw DataNotFound("continue-from-mvn");
        const VoxelData &mvn = rundata.GetVoxelData("continue-from-mvn");
        continue_from_mvn = true;
        const int n_all = P + m_noise_params, n_cov_all = n_all * (n_all + 1) / 2;
        if (mvn.rows != n_cov_all + n_all + 1)
            throw FabberRunDataError("continue-from-mvn: MVN has the wrong number of parameters for this model / noise");
        init_mean.resize((size_t)P * N);
        init_cov.resize((size_t)NT * N);
        init_noise.assign((size_t)NN * N, 0.0);
        parallel_for(N, [&](size_t vb, size_t ve) {
            for (size_t v = vb; v < ve; v++)
            {
                for (int i = 0; i < P; i++)
                    init_mean[(size_t)i * N + v] = mvn.at(n_cov_all + i, v);
                for (int r = 0; r < P; r++)
                    for (int c = 0; c <= r; c++)
                        init_cov[(size_t)tri(r, c) * N + v] = mvn.at(tri(r, c), v);
                if (m_ar && m_nphis == 2)
                {
                    /* MVN order: the alphas, the two phis. Fields: b1 c1 b2 c2, alpha means, packed alpha precisions
                     * (= inverse of the MVN's alpha covariance block) */
                    const int nA = m_nalphas, ia = P, ip = P + nA;
                    double w[4][8];
                    for (int r = 0; r < nA; r++)
                        for (int c = 0; c < nA; c++)
                        {
                            w[r][c] = mvn.at(tri(ia + r, ia + c), v);
                            w[r][nA + c] = r == c ? 1.0 : 0.0;
                        }
                    for (int k = 0; k < nA; k++) /* Gauss-Jordan, partial pivoting */
                    {
                        int piv = k;
                        for (int r = k + 1; r < nA; r++)
                            if (std::fabs(w[r][k]) > std::fabs(w[piv][k]))
                                piv = r;
                        for (int c = 0; c < 2 * nA; c++)
                            std::swap(w[k][c], w[piv][c]);
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
                    for (int i = 0; i < 2; i++)
                    {
                        const double mean = mvn.at(n_cov_all + ip + i, v), var = mvn.at(tri(ip + i, ip + i), v);
                        init_noise[(size_t)(2 * i) * N + v] = var / mean;
                        init_noise[(size_t)(2 * i + 1) * N + v] = mean * mean / var;
                    }
                    for (int i = 0; i < nA; i++)
                        init_noise[(size_t)(4 + i) * N + v] = mvn.at(n_cov_all + ia + i, v);
                    for (int r = 0; r < nA; r++)
                        for (int c = 0; c <= r; c++)
                            init_noise[(size_t)(4 + nA + tri(r, c)) * N + v] = w[r][nA + c];
                }
                else if (m_ar)
                {
                    /* MVN order: alpha (2), phi (1). InputFromMVN: Gamma from mean/variance (dist_gamma.cc:29) */
                    const int ia = P, ip = P + 2;
                    const double a1 = mvn.at(n_cov_all + ia, v), a2 = mvn.at(n_cov_all + ia + 1, v);
                    const double c11 = mvn.at(tri(ia, ia), v), c21 = mvn.at(tri(ia + 1, ia), v),
                                 c22 = mvn.at(tri(ia + 1, ia + 1), v);
                    const double det = c11 * c22 - c21 * c21;
                    const double mean = mvn.at(n_cov_all + ip, v), var = mvn.at(tri(ip, ip), v);
                    init_noise[0 * N + v] = var / mean;
                    init_noise[1 * N + v] = mean * mean / var;
                    init_noise[2 * N + v] = a1;
                    init_noise[3 * N + v] = a2;
                    init_noise[4 * N + v] = c22 / det;
                    init_noise[5 * N + v] = -c21 / det;
                    init_noise[6 * N + v] = c11 / det;
                }
                else
                    for (int i = 0; i < m_nphis; i++)
                    {
                        const double mean = mvn.at(n_cov_all + P + i, v), var = mvn.at(tri(P + i, P + i), v);
                        init_noise[(size_t)(2 * i) * N + v] = var / mean;            /* b = variance / mean */
                        init_noise[(size_t)(2 * i + 1) * N + v] = mean * mean / var; /* c = mean^2 / variance */
                    }
            }
        });
    }
    catch (DataNotFound &)
    {
    }
    /* locked-linear-from-mvn (inference_vb.cc:128-129,171-178,227-231): the first P means of each voxel's MVN
     * are the fixed linearisation centre. Like the reference, no check that the MVN belongs to this model
     * beyond holding at least P parameters; only the spatial method honours it (:695 vs :443). */
    std::vector<double> lock_centre;
    const std::string lock_name = rundata.GetStringDefault("locked-linear-from-mvn", "");
    if (lock_name != "" && spatial)
    {
        const VoxelData &mvn = rundata.GetVoxelData(lock_name);
        /* rows = n(n+1)/2 + n + 1  ->  n */
        int n_all = 0;
        while ((n_all + 1) * (n_all + 2) / 2 < mvn.rows)
            n_all++;
        if ((n_all + 1) * (n_all + 2) / 2 != mvn.rows || n_all < P || (size_t)mvn.cols != N)
            throw FabberRunDataError("locked-linear-from-mvn: not an MVN for this mask / model");
        const int n_cov_all = n_all * (n_all + 1) / 2;
        for (size_t v = 0; v < N; v++)
            if (mvn.at(mvn.rows - 1, v) != 1) /* dist_mvn.cc:365-369 */
                throw FabberRunDataError("MVNDist::Load - Voxel data does not contain a valid MVN - last value != 1");
        lock_centre.resize((size_t)P * N);
        parallel_for(N, [&](size_t vb, size_t ve) {
            for (size_t v = vb; v < ve; v++)
                for (int i = 0; i < P; i++)
                    lock_centre[(size_t)i * N + v] = mvn.at(n_cov_all + i, v);
        });
        rundata.Log() << "Vb::Loading fixed linearization centres from the MVN '" << lock_name << "'" << std::endl;
    }
    /* rows x N host array -> the columns [v0, v1) of every row on the current device */
    auto copy_columns = [&](void *dst, const std::vector<double> &h, int rows, const DevCtx &c) {
        const size_t n = c.v1 - c.v0;
        check(fabber_cuda_memcpy2d_h2d(dst, n * sizeof(double), h.data() + c.v0, N * sizeof(double), n * sizeof(double),
                  (size_t)rows, nullptr),
            "copying to the GPU");
    };
    if (!m_nlls && rundata.GetBool("output-only"))
    {
        if (!continue_from_mvn)
            throw FabberRunDataError("output-only requires continue-from-mvn");
        for (size_t g = 0; g < m_ctx.size(); g++)
        {
            DevCtx &c = m_ctx[g];
            DeviceScope scope(c.device);
            copy_columns(c.mean.p, init_mean, P, c);
            copy_columns(c.cov.p, init_cov, NT, c);
            copy_columns(c.noise.p, init_noise, NN, c);
            check(fabber_cuda_memset(c.F.p, 0, c.F.bytes, nullptr), "output-only");
            check(fabber_cuda_memset(c.its.p, 0, c.its.bytes, nullptr), "output-only");
            check(fabber_cuda_stream_sync(nullptr), "output-only");
        }
        m_needF = false;
        m_output_only = true;
        rundata.Log() << "Vb::DoCalculations output-only set - not performing any calculations" << std::endl;
        return;
    }

    /* ---- inputs --------------------------------------------------------------------------------------- */
    m_description = std::string(m_nlls ? "NLLSInferenceTechnique::" : "Vb::") + (spatial ? "Spatial" : "Voxelwise")
        + " calculations on the GPU: " + stringify(N)
        + " voxels x " + stringify(T) + " time points, " + stringify(P) + " parameters, " + stringify(m_ctx.size())
        + " device" + (m_ctx.size() == 1 ? "" : "s");
    rundata.Log() << m_description << std::endl;
    rundata.Progress(0, (int)N);
    std::vector<Scratch> &scratch = m_scratch;
    auto upload_columns = [&](const std::vector<double> &h, int rows, const DevCtx &c) -> void * {
        Scratch sc;
        sc.device = c.device;
        sc.d.bytes = (size_t)rows * (c.v1 - c.v0) * sizeof(double);
        sc.d.p = cached_device_alloc(sc.d.bytes);
        scratch.push_back(sc);
        copy_columns(sc.d.p, h, rows, c);
        return sc.d.p;
    };
    std::vector<std::vector<double>> images(P);
    for (int i = 0; i < P; i++)
        if (params[i].prior_type == 'I')
        {
            /* ImagePrior (priors.cc:118-124): a one-volume voxel data item named by the parameter's image option */
            const VoxelData &img = rundata.GetVoxelData(params[i].options.find("image")->second);
            images[i].resize(N);
            for (size_t v = 0; v < N; v++)
                images[i][v] = img.at(0, v);
        }
    m_bufs.assign(m_ctx.size(), fabber_cuda_vb_buffers());
    m_probs.assign(m_ctx.size(), prob);
    std::vector<fabber_cuda_vb_buffers> &bufs = m_bufs;
    std::vector<fabber_cuda_vb_problem> &probs = m_probs;
    for (size_t g = 0; g < m_ctx.size(); g++)
    {
        DevCtx &c = m_ctx[g];
        DeviceScope scope(c.device);
        fabber_cuda_vb_buffers &buf = bufs[g];
        memset(&buf, 0, sizeof(buf));
        probs[g].n_voxels = (int)(c.v1 - c.v0);
        buf.data = c.data;
        for (int i = 0; i < P; i++)
            if (!images[i].empty())
                buf.image_prior[i] = (const double *)upload_columns(images[i], 1, c);
        if (continue_from_mvn)
        {
            buf.init_mean = (const double *)upload_columns(init_mean, P, c);
            buf.init_cov = (const double *)upload_columns(init_cov, NT, c);
            buf.init_noise = (const double *)upload_columns(init_noise, NN, c);
        }
        if (!lock_centre.empty())
            buf.lock_centre = (const double *)upload_columns(lock_centre, P, c);
        if (spatial)
        {
            /* [3][n] coordinates of this device's voxels (global coordinates) */
            const size_t n = c.v1 - c.v0;
            Scratch sc;
            sc.device = c.device;
            sc.d.bytes = 3 * n * sizeof(int);
            sc.d.p = cached_device_alloc(sc.d.bytes);
            scratch.push_back(sc);
            check(fabber_cuda_memcpy2d_h2d(sc.d.p, n * sizeof(int), rundata.Coords().data() + c.v0, N * sizeof(int),
                      n * sizeof(int), 3, nullptr),
                "copying to the GPU");
            buf.coords = (const int *)sc.d.p;
        }
        buf.mean = (double *)c.mean.p;
        buf.cov = (double *)c.cov.p;
        buf.noise = (double *)c.noise.p;
        buf.free_energy = (double *)c.F.p;
        buf.status = (int *)c.status.p;
        buf.iterations = (int *)c.its.p;
        buf.f_history = m_fhist_len > 0 ? (double *)c.hist.p : nullptr;
    }
    m_slabs = slabs;
    rundata.Log() << "Vb::timing: option translation + device buffers " << sw.lap_ms() << " ms" << std::endl;
}

void Vb::FreeScratch()
{
    for (size_t i = 0; i < m_scratch.size(); i++)
    {
        DeviceScope scope(m_scratch[i].device);
        fabber_cuda_stream_sync(nullptr);
        cached_device_free(m_scratch[i].d.p, m_scratch[i].d.bytes);
    }
    m_scratch.clear();
}

/* voxelwise VB on one uploaded block of voxels, on the block's device, behind the block's upload event */
void Vb::LaunchBlock(VoxelData &data, size_t b)
{
    if (m_launch_rc != FABBER_CUDA_OK)
        return;
    const VoxelData::Block &blk = data.blocks[b];
    const DevCtx &c = m_ctx[blk.part];
    DeviceScope scope(c.device);
    m_launch_rc = fabber_cuda_stream_wait_event(nullptr, blk.ready);
    if (m_launch_rc == FABBER_CUDA_OK)
        m_launch_rc = fabber_cuda_vb_voxelwise_range(&m_probs[blk.part], &m_bufs[blk.part], (int)(blk.v0 - c.v0),
            (int)(blk.v1 - c.v0), nullptr);
    if (m_launch_rc != FABBER_CUDA_OK)
        m_launch_error = fabber_cuda_last_error();
}

void Vb::LaunchAll(VoxelData &data)
{
    const bool spatial = m_spatial, slabs = m_slabs;
    fabber_cuda_vb_problem &prob = m_prob;
    std::vector<fabber_cuda_vb_buffers> &bufs = m_bufs;
    std::vector<fabber_cuda_vb_problem> &probs = m_probs;
    int rc = FABBER_CUDA_OK;
    if (spatial && slabs)
    {
        /* ONE volume over all the devices: z-slabs whose kernels talk through each other's memory */
        std::vector<fabber_cuda_slab_part> sp(m_ctx.size());
        for (size_t g = 0; g < m_ctx.size(); g++)
        {
            DeviceScope scope(m_ctx[g].device);
            data.wait_uploaded((int)g);
            check(fabber_cuda_stream_sync(nullptr), "copying data to the GPU"); /* the engine runs on streams of its own */
            sp[g].device = m_ctx[g].device;
            sp[g].v0 = (int)m_ctx[g].v0;
            sp[g].v1 = (int)m_ctx[g].v1;
            sp[g].own0 = (int)m_ctx[g].own0;
            sp[g].own1 = (int)m_ctx[g].own1;
            sp[g].own_z0 = data.parts[g].z0;
            sp[g].own_z1 = data.parts[g].z1;
            sp[g].buf = bufs[g];
        }
        rc = fabber_cuda_vb_spatial_multi(&prob, (int)sp.size(), sp.data());
    }
    else if (spatial)
    {
        DeviceScope scope(m_ctx[0].device);
        data.wait_uploaded(0);
        rc = fabber_cuda_vb_spatial(&probs[0], &bufs[0], nullptr);
    }
    else
    {
        /* voxels are independent (inference_vb.cc:423-571): one launch per uploaded block on the block's
         * device, each behind its block's event - block k computes while the blocks after it are still on
         * the PCIe bus, and every device works on its own range. All calls are asynchronous. */
        for (size_t b = 0; b < data.blocks.size(); b++)
            LaunchBlock(data, b);
        rc = m_launch_rc;
        if (data.blocks.empty()) /* uploaded in one piece */
            for (size_t g = 0; g < m_ctx.size() && rc == FABBER_CUDA_OK; g++)
            {
                DeviceScope scope(m_ctx[g].device);
                rc = fabber_cuda_vb_voxelwise(&probs[g], &bufs[g], nullptr);
            }
    }
    if (rc != FABBER_CUDA_OK && m_launch_rc == FABBER_CUDA_OK)
    {
        m_launch_rc = rc;
        m_launch_error = fabber_cuda_last_error();
    }
}

/* after the last launch: errors, the bad-voxel policy, scratch */
void Vb::Finish(FabberRunData &rundata)
{
    StopWatch sw;
    const size_t N = m_nvoxels;
    struct ScratchGuard
    {
        Vb &vb;
        ~ScratchGuard() { vb.FreeScratch(); }
    } guard = { *this };
    if (m_launch_rc == FABBER_CUDA_ERR_INVALID)
        throw FabberRunDataError(std::string("Vb: ") + m_launch_error);
    if (m_launch_rc != FABBER_CUDA_OK)
        throw FabberInternalError(std::string("VB kernels: ") + m_launch_error);

    /* ---- bad-voxel policy (inference_vb.cc:529-544; set-up failures are never caught, :235) ------------
     * the status words are scanned on the devices; only a count and the first offender come back */
    static const char *reason[] = { "", "LinearizedFwdModel::ReCentre: Non-finite values found in offset",
        "LinearizedFwdModel::ReCentre: Non-finite values found in jacobian", "Non-finite free energy!",
        "matrix is singular", "Ar1cNoiseModel::UpdateAlpha Negative variance!", "voxel ignored" };
    long n_bad = 0;
    long first = -1;
    int code = 0;
    for (size_t g = 0; g < m_ctx.size(); g++)
    {
        DevCtx &c = m_ctx[g];
        DeviceScope scope(c.device);
        int first_g = -1, code_g = 0;
        /* only the columns this device publishes (a z-slab's ghost voxels carry FABBER_VOX_GHOST) */
        const int bad_g = fabber_cuda_check_status((const int *)c.status.p + (c.own0 - c.v0), (int)(c.own1 - c.own0),
            &first_g, &code_g, nullptr);
        if (bad_g < 0)
            check(bad_g, "VB kernels");
        if (bad_g > 0 && first < 0) /* ranges are in voxel order: the first device with a failure holds the first */
        {
            first = (long)c.own0 + first_g;
            code = code_g;
        }
        n_bad += bad_g;
    }
    rundata.Log() << "Vb::timing: waiting for the kernels " << sw.lap_ms() << " ms" << std::endl;
    rundata.Progress((int)N, (int)N);
    if (n_bad > 0)
    {
        const int c = code & 0xff;
        const std::string why = c >= 1 && c <= 6 ? reason[c] : "numerical error";
        rundata.Log() << (m_nlls ? "NLLSInferenceTechnique::" : "Vb::") << "Internal error for voxel " << first + 1 << " : "
                      << why << std::endl;
        if (m_halt_bad_voxel || (code & FABBER_VOX_SETUP_FLAG))
            throw FabberInternalError(why);
        rundata.Log() << "Vb::" << n_bad << " voxels had numerical errors and kept their last state" << std::endl;
    }
}

void Vb::SaveResults(FabberRunData &rundata)
{
    StopWatch sw;
    const size_t N = m_nvoxels;
    const int P = m_num_params, T = m_ntimes;
    const std::vector<Parameter> &params = m_model->Params();
    const int NP_all = P + m_noise_params;
    /* Every requested output map is produced on the device in float32 (fabber_cuda_vb_save_results) and
     * downloaded straight into the pinned buffer fabber_get_data will read from - each device writes the
     * columns of its own voxel range. */
    struct Wanted
    {
        float *fabber_cuda_vb_outputs::*slot;
        int rows;
        std::vector<std::string> keys; /* one key per row group of `rows_per_key` rows */
        int rows_per_key;
    };
    std::vector<Wanted> wanted;
    auto want = [&](float *fabber_cuda_vb_outputs::*slot, int rows, const std::vector<std::string> &keys, int rows_per_key) {
        Wanted w;
        w.slot = slot;
        w.rows = rows;
        w.keys = keys;
        w.rows_per_key = rows_per_key;
        wanted.push_back(w);
    };
    auto per_param = [&](const std::string &prefix) {
        std::vector<std::string> k;
        for (int i = 0; i < P; i++)
            k.push_back(prefix + params[i].name);
        return k;
    };
    if (rundata.GetBool("save-mean"))
        want(&fabber_cuda_vb_outputs::mean, P, per_param("mean_"), 1);
    if (rundata.GetBool("save-std"))
        want(&fabber_cuda_vb_outputs::std, P, per_param("std_"), 1);
    if (rundata.GetBool("save-zstat"))
        want(&fabber_cuda_vb_outputs::zstat, P, per_param("zstat_"), 1);
    if (rundata.GetBool("save-var"))
        want(&fabber_cuda_vb_outputs::var, P, per_param("var_"), 1);
    /* Quirk kept: Ar1cNoiseModel::NumParams() returns nPhis (noisemodel_ar.cc:362-365) although its MVN
     * block is (the alphas, the phis), so the reference's noise_means / noise_stdevs hold num-echoes volumes for
     * AR noise - the first elements of that block, i.e. alpha1 (and alpha2 with two echoes)
     * (inference_vb.cc:982-988). */
    const int noise_rows = m_ar ? m_nphis : m_noise_params;
    if (rundata.GetBool("save-noise-mean") && m_noise_params > 0)
        want(&fabber_cuda_vb_outputs::noise_mean, m_noise_params, std::vector<std::string>(1, "noise_means"), noise_rows);
    if (rundata.GetBool("save-noise-std") && m_noise_params > 0)
        want(&fabber_cuda_vb_outputs::noise_std, m_noise_params, std::vector<std::string>(1, "noise_stdevs"), noise_rows);
    if (rundata.GetBool("save-mvn"))
    {
        const int rows = NP_all * (NP_all + 1) / 2 + NP_all + 1; /* MVNDist::Save, dist_mvn.cc:377-433 */
        want(&fabber_cuda_vb_outputs::final_mvn, rows, std::vector<std::string>(1, "finalMVN"), rows);
    }
    if (m_saveF && m_needF)
        want(&fabber_cuda_vb_outputs::free_energy, 1, std::vector<std::string>(1, "freeEnergy"), 1);
    int f_history_rows = 0;
    if (m_saveFsHistory && m_fhist_len > 0 && m_needF)
    {
        /* one row per pass plus the final value pushed after the loop (inference_vb.cc:553-554); voxels
         * that stopped early repeat their last value (:1038-1045) */
        int max_its = 0;
        for (size_t g = 0; g < m_ctx.size(); g++)
        {
            DeviceScope scope(m_ctx[g].device);
            int m = 0;
            check(fabber_cuda_max_int((const int *)m_ctx[g].its.p, (int)(m_ctx[g].v1 - m_ctx[g].v0), &m, nullptr),
                "free energy history");
            max_its = std::max(max_its, m);
        }
        f_history_rows = std::min(max_its + 1, m_fhist_len);
        want(&fabber_cuda_vb_outputs::f_history, f_history_rows, std::vector<std::string>(1, "freeEnergyHistory"),
            f_history_rows);
    }
    if (rundata.GetBool("save-model-fit"))
        want(&fabber_cuda_vb_outputs::model_fit, T, std::vector<std::string>(1, "modelfit"), T);
    const bool residuals = rundata.GetBool("save-residuals");
    if (residuals)
        want(&fabber_cuda_vb_outputs::residuals, T, std::vector<std::string>(1, "residuals"), T);
    /* host side of every output, created before any device starts filling its columns */
    for (size_t i = 0; i < wanted.size(); i++)
        for (size_t k = 0; k < wanted[i].keys.size(); k++)
            rundata.NewVoxelData(wanted[i].keys[k], wanted[i].rows_per_key);
    if (N == 0 || m_ctx.empty())
    {
        /* zero voxels is not an error and the outputs exist, with no columns (test/test_inference.cc:57-73) */
        rundata.Log() << "Vb::Done writing results." << std::endl;
        return;
    }

    struct Pending
    {
        int device;
        void *dev;
        size_t bytes;
    };
    std::vector<Pending> pending;
    int rc = FABBER_CUDA_OK;
    for (size_t g = 0; g < m_ctx.size() && rc == FABBER_CUDA_OK && !wanted.empty(); g++)
    {
        DevCtx &c = m_ctx[g];
        DeviceScope scope(c.device);
        const size_t n = c.v1 - c.v0;
        fabber_cuda_vb_problem prob = m_prob;
        prob.n_voxels = (int)n;
        fabber_cuda_vb_buffers buf;
        memset(&buf, 0, sizeof(buf));
        buf.mean = (double *)c.mean.p;
        buf.cov = (double *)c.cov.p;
        buf.noise = (double *)c.noise.p;
        buf.free_energy = m_needF ? (double *)c.F.p : nullptr;
        buf.iterations = (int *)c.its.p;
        buf.f_history = m_fhist_len > 0 ? (double *)c.hist.p : nullptr;
        fabber_cuda_vb_outputs out;
        memset(&out, 0, sizeof(out));
        out.f_history_rows = f_history_rows;
        if (residuals)
            out.data = c.data;
        std::vector<void *> dev_of(wanted.size());
        for (size_t i = 0; i < wanted.size(); i++)
        {
            Pending p;
            p.device = c.device;
            p.bytes = (size_t)wanted[i].rows * n * sizeof(float);
            p.dev = cached_device_alloc(p.bytes);
            pending.push_back(p);
            out.*(wanted[i].slot) = (float *)p.dev;
            dev_of[i] = p.dev;
        }
        rc = fabber_cuda_vb_save_results(&prob, &buf, &out, nullptr);
        for (size_t i = 0; i < wanted.size() && rc == FABBER_CUDA_OK; i++)
            for (size_t k = 0; k < wanted[i].keys.size() && rc == FABBER_CUDA_OK; k++)
            {
                VoxelData &vd = rundata.MutableVoxelData(wanted[i].keys[k]);
                rc = fabber_cuda_memcpy2d_d2h(vd.f + c.own0, N * sizeof(float),
                    (const float *)dev_of[i] + k * (size_t)wanted[i].rows_per_key * n + (c.own0 - c.v0), n * sizeof(float),
                    (c.own1 - c.own0) * sizeof(float), (size_t)wanted[i].rows_per_key, nullptr);
            }
    }
    int rc_sync = FABBER_CUDA_OK;
    for (size_t g = 0; g < m_ctx.size(); g++)
    {
        DeviceScope scope(m_ctx[g].device);
        const int r = fabber_cuda_stream_sync(nullptr);
        if (r != FABBER_CUDA_OK)
            rc_sync = r;
    }
    for (size_t i = 0; i < pending.size(); i++)
        cached_device_free(pending[i].dev, pending[i].bytes);
    ReleaseDevice();
    check(rc, "saving results");
    check(rc_sync, "saving results");
    /* file-based front ends write each output now (rundata_newimage.cc:140-183); a no-op for the array one */
    for (size_t i = 0; i < wanted.size(); i++)
        for (size_t k = 0; k < wanted[i].keys.size(); k++)
            rundata.SaveVoxelData(wanted[i].keys[k], wanted[i].keys[k] == "finalMVN" ? VDT_MVN : VDT_SCALAR);
    rundata.Log() << "Vb::timing: SaveResults " << sw.lap_ms() << " ms" << std::endl;
    rundata.Log() << "Vb::Done writing results." << std::endl;
}

} // namespace fabber_b200
