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

#include "fabber_host.h"

namespace fabber_b200
{
namespace
{
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
} // namespace

std::vector<std::string> Vb::GetKnownMethods()
{
    std::vector<std::string> k;
    k.push_back("spatialvb");
    k.push_back("vb");
    return k;
}
std::string Vb::GetDescription() { return "Variational Bayes inference technique (B200 GPU implementation)"; }
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
        { "output-only", OPT_BOOL,
            "Skip model fitting, just output requested data based on supplied MVN. Can only be used with "
            "continue-from-mvn",
            true, "" },
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

    const std::string noise = rundata.GetString("noise");
    if (noise == "white")
        m_ar = false;
    else if (noise == "ar")
        m_ar = true;
    else
        throw InvalidOptionValue("noise", noise, "Unrecognized noise model (white, ar)");
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

void Vb::DoCalculations(FabberRunData &rundata)
{
    const VoxelData &data = rundata.GetMainVoxelData();
    if (!data.is_float)
        throw FabberInternalError("main voxel data must be the float32 series set with fabber_set_data");
    const size_t N = data.cols;
    const int T = data.rows;
    m_nvoxels = N;
    m_ntimes = T;
    std::vector<Parameter> params;
    m_model->GetParameters(rundata, params);
    const int P = m_num_params, NT = P * (P + 1) / 2;

    fabber_cuda_vb_problem prob;
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

    /* ---- noise model options ---------------------------------------------------------------------- */
    std::vector<unsigned char> pattern(T, 0), masked(T, 0);
    const std::vector<int> mt = rundata.GetIntList("mt", 1); /* inference.cc:96-103 */
    for (size_t i = 0; i < mt.size(); i++)
    {
        if (mt[i] > T)
            throw InvalidOptionValue("mt", stringify(mt[i]), "Masked time point beyond the end of the data");
        masked[mt[i] - 1] = 1;
    }
    if (m_ar)
    {
        /* noisemodel_ar.cc:305-349: one echo, no cross terms is what the device kernels implement */
        if (!mt.empty())
            throw InvalidOptionValue("noise", "ar", "AR noise model does not support masked time points");
        if (rundata.GetIntDefault("num-echoes", 1) != 1)
            throw InvalidOptionValue("num-echoes", rundata.GetString("num-echoes"), "only 1 echo has a device kernel");
        if (rundata.GetStringDefault("ar1-cross-terms", "none") != "none")
            throw InvalidOptionValue("ar1-cross-terms", rundata.GetString("ar1-cross-terms"), "only 'none' has a device kernel");
        prob.noise_type = FABBER_NOISE_AR1;
        prob.n_phis = 1;
        prob.noise_prior_b[0] = 1e6; /* noisemodel_ar.cc:379-403 */
        prob.noise_prior_c[0] = 1e-6;
        prob.noise_post_b[0] = 1e-8;
        prob.noise_post_c[0] = 1e-6;
        prob.ar_alpha_prior_prec = 1e-4;
        m_noise_params = 3; /* alpha (2) + phi (1): Ar1cParams::OutputAsMVN, noisemodel_ar.cc:287-300 */
        m_nphis = 1;
    }
    else
    {
        prob.noise_type = FABBER_NOISE_WHITE;
        const std::string pat = rundata.GetStringDefault("noise-pattern", "1");
        std::vector<int> digits; /* noisemodel_white.cc:166-215 */
        for (size_t i = 0; i < pat.size(); i++)
        {
            const char ch = pat[i];
            if (ch >= '1' && ch <= '9')
                digits.push_back(ch - '0');
            else if (ch >= 'A' && ch <= 'Z')
                digits.push_back(ch - 'A' + 10);
            else if (ch >= 'a' && ch <= 'z')
                digits.push_back(ch - 'a' + 10);
            else
                throw InvalidOptionValue("noise-pattern", pat, "Invalid character in pattern");
        }
        int nphis = *std::max_element(digits.begin(), digits.end());
        if (nphis > FABBER_CUDA_MAX_PHIS)
            throw InvalidOptionValue("noise-pattern", pat, "more noise precisions than the device kernels carry");
        for (int t = 0; t < T; t++)
            pattern[t] = (unsigned char)(digits[t % digits.size()] - 1);
        prob.n_phis = nphis;
        m_nphis = nphis;
        m_noise_params = nphis;
        const double phiprior = rundata.GetDoubleDefault("prior-noise-stddev", -1);
        if (phiprior < 0 && phiprior != -1)
            throw InvalidOptionValue("prior-noise-stddev", stringify(phiprior), "Must be > 0");
        for (int i = 0; i < nphis; i++)
        {
            if (phiprior == -1)
            {
                prob.noise_prior_b[i] = 1e6; /* noisemodel_white.cc:142-149 */
                prob.noise_prior_c[i] = 1e-6;
                prob.noise_post_b[i] = 1e-8;
                prob.noise_post_c[i] = 50;
            }
            else
            {
                prob.noise_prior_c[i] = prob.noise_post_c[i] = 0.5; /* :156-161 */
                prob.noise_prior_b[i] = prob.noise_post_b[i] = 1 / (phiprior * phiprior * 0.5);
            }
        }
        prob.locked_noise_stdev = rundata.GetDoubleDefault("locked-noise-stdev", -1);
    }
    prob.phi_pattern = pattern.data();
    prob.time_masked = mt.empty() ? nullptr : masked.data();

    /* ---- convergence (setup.cc:49-57, convergence.cc Initialize functions) --------------------------- */
    const std::string conv = rundata.GetStringDefault("convergence", "maxits");
    if (conv == "maxits")
        prob.conv_type = FABBER_CONV_MAXITS;
    else if (conv == "pointzeroone")
        prob.conv_type = FABBER_CONV_FCHANGE;
    else if (conv == "freduce")
        prob.conv_type = FABBER_CONV_FREDUCE;
    else if (conv == "trialmode")
        prob.conv_type = FABBER_CONV_TRIALMODE;
    else if (conv == "lm")
        prob.conv_type = FABBER_CONV_LM;
    else
        throw InvalidOptionValue("convergence", conv, "Unrecognized convergence detector");
    prob.max_iterations = rundata.GetIntDefault("max-iterations", 10);
    if (prob.max_iterations <= 0)
        throw InvalidOptionValue("max-iterations", stringify(prob.max_iterations), "Must be positive");
    prob.fchange = (conv == "lm") ? rundata.GetDoubleDefault("max-fchange", 0.01)
                                  : rundata.GetDoubleDefault("min-fchange", 0.01);
    if (!(prob.fchange > 0))
        throw InvalidOptionValue(conv == "lm" ? "max-fchange" : "min-fchange", stringify(prob.fchange), "Must be positive");
    prob.max_trials = rundata.GetIntDefault("max-trials", 10);
    if (prob.max_trials <= 0)
        throw InvalidOptionValue("max-trials", stringify(prob.max_trials), "Must be positive");
    const bool spatial = IsSpatial(rundata, params);
    const bool useF = !spatial && prob.conv_type != FABBER_CONV_MAXITS;
    m_needF = useF || m_printF || m_saveF || m_saveFsHistory; /* inference_vb.cc:242 */
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

    const int NN = m_ar ? FABBER_CUDA_AR_NOISE_FIELDS : 2 * m_nphis;
    m_mean.assign((size_t)P * N, 0.0);
    m_cov.assign((size_t)NT * N, 0.0);
    m_noise.assign((size_t)NN * N, 0.0);
    m_F.assign(N, 9999.0);
    m_status.assign(N, 0);
    m_iterations.assign(N, 0);
    m_Fhist.assign((size_t)m_fhist_len * N, 0.0);
    if (N == 0)
        return; /* zero voxels is not an error (test/test_inference.cc:57-73) */

    /* ---- restart / output-only (inference_vb.cc:181-216, 385-389) ------------------------------------ */
    std::vector<double> init_mean, init_cov, init_noise;
    bool continue_from_mvn = false;
    try
    {
        const VoxelData &mvn = rundata.GetVoxelData("continue-from-mvn");
        continue_from_mvn = true;
        const int n_all = P + m_noise_params, n_cov_all = n_all * (n_all + 1) / 2;
        if (mvn.rows != n_cov_all + n_all + 1)
            throw FabberRunDataError("continue-from-mvn: MVN has the wrong number of parameters for this model / noise");
        init_mean.resize((size_t)P * N);
        init_cov.resize((size_t)NT * N);
        init_noise.assign((size_t)NN * N, 0.0);
        for (size_t v = 0; v < N; v++)
        {
            for (int i = 0; i < P; i++)
                init_mean[(size_t)i * N + v] = mvn.at(n_cov_all + i, v);
            for (int r = 0; r < P; r++)
                for (int c = 0; c <= r; c++)
                    init_cov[(size_t)tri(r, c) * N + v] = mvn.at(tri(r, c), v);
            if (m_ar)
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
                    init_noise[(size_t)(2 * i) * N + v] = var / mean;         /* b = variance / mean */
                    init_noise[(size_t)(2 * i + 1) * N + v] = mean * mean / var; /* c = mean^2 / variance */
                }
        }
    }
    catch (DataNotFound &)
    {
    }
    if (rundata.GetBool("output-only"))
    {
        if (!continue_from_mvn)
            throw FabberRunDataError("output-only requires continue-from-mvn");
        m_mean = init_mean;
        m_cov = init_cov;
        m_noise = init_noise;
        rundata.Log() << "Vb::DoCalculations output-only set - not performing any calculations" << std::endl;
        return;
    }

    /* ---- device buffers: inputs up, one launch, results down ------------------------------------------ */
    rundata.Log() << "Vb::" << (spatial ? "Spatial" : "Voxelwise") << " calculations on the GPU: " << N << " voxels x "
                  << T << " time points, " << P << " parameters" << std::endl;
    rundata.Progress(0, (int)N);
    fabber_cuda_vb_buffers buf;
    memset(&buf, 0, sizeof(buf));
    DeviceBuf d_data((size_t)T * N * sizeof(float));
    check(fabber_cuda_memcpy_h2d(d_data.p, data.f, (size_t)T * N * sizeof(float), nullptr), "copying data to the GPU");
    buf.data = (const float *)d_data.p;
    std::vector<std::unique_ptr<DeviceBuf>> keep;
    auto upload = [&](const std::vector<double> &h) -> const double * {
        keep.emplace_back(new DeviceBuf(h.size() * sizeof(double)));
        check(fabber_cuda_memcpy_h2d(keep.back()->p, h.data(), h.size() * sizeof(double), nullptr), "copying to the GPU");
        return (const double *)keep.back()->p;
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
            buf.image_prior[i] = upload(images[i]);
        }
    if (continue_from_mvn)
    {
        buf.init_mean = upload(init_mean);
        buf.init_cov = upload(init_cov);
        buf.init_noise = upload(init_noise);
    }
    std::unique_ptr<DeviceBuf> d_coords;
    if (spatial)
    {
        d_coords.reset(new DeviceBuf(3 * N * sizeof(int)));
        check(fabber_cuda_memcpy_h2d(d_coords->p, rundata.Coords().data(), 3 * N * sizeof(int), nullptr), "copying coordinates");
        buf.coords = (const int *)d_coords->p;
    }
    DeviceBuf d_mean(m_mean.size() * sizeof(double)), d_cov(m_cov.size() * sizeof(double)),
        d_noise(m_noise.size() * sizeof(double)), d_F(N * sizeof(double)), d_status(N * sizeof(int)),
        d_its(N * sizeof(int)), d_hist(std::max<size_t>(1, m_Fhist.size()) * sizeof(double));
    buf.mean = (double *)d_mean.p;
    buf.cov = (double *)d_cov.p;
    buf.noise = (double *)d_noise.p;
    buf.free_energy = (double *)d_F.p;
    buf.status = (int *)d_status.p;
    buf.iterations = (int *)d_its.p;
    buf.f_history = m_fhist_len > 0 ? (double *)d_hist.p : nullptr;
    check(fabber_cuda_stream_sync(nullptr), "uploading inputs");

    int rc = spatial ? fabber_cuda_vb_spatial(&prob, &buf, nullptr) : fabber_cuda_vb_voxelwise(&prob, &buf, nullptr);
    if (rc == FABBER_CUDA_ERR_INVALID)
        throw FabberRunDataError(std::string("Vb: ") + fabber_cuda_last_error());
    check(rc, "VB kernels");
    check(fabber_cuda_memcpy_d2h(m_mean.data(), d_mean.p, m_mean.size() * sizeof(double), nullptr), "results");
    check(fabber_cuda_memcpy_d2h(m_cov.data(), d_cov.p, m_cov.size() * sizeof(double), nullptr), "results");
    check(fabber_cuda_memcpy_d2h(m_noise.data(), d_noise.p, m_noise.size() * sizeof(double), nullptr), "results");
    check(fabber_cuda_memcpy_d2h(m_F.data(), d_F.p, N * sizeof(double), nullptr), "results");
    check(fabber_cuda_memcpy_d2h(m_status.data(), d_status.p, N * sizeof(int), nullptr), "results");
    check(fabber_cuda_memcpy_d2h(m_iterations.data(), d_its.p, N * sizeof(int), nullptr), "results");
    if (m_fhist_len > 0)
        check(fabber_cuda_memcpy_d2h(m_Fhist.data(), d_hist.p, m_Fhist.size() * sizeof(double), nullptr), "results");
    check(fabber_cuda_stream_sync(nullptr), "VB kernels");
    rundata.Progress((int)N, (int)N);

    /* ---- bad-voxel policy (inference_vb.cc:529-544; set-up failures are never caught, :235) ------------ */
    static const char *reason[] = { "", "LinearizedFwdModel::ReCentre: Non-finite values found in offset",
        "LinearizedFwdModel::ReCentre: Non-finite values found in jacobian", "Non-finite free energy!",
        "matrix is singular", "Ar1cNoiseModel::UpdateAlpha Negative variance!", "voxel ignored" };
    size_t n_bad = 0;
    for (size_t v = 0; v < N; v++)
    {
        const int st = m_status[v];
        if (st == 0)
            continue;
        n_bad++;
        const int code = st & 0xff;
        const std::string why = code >= 1 && code <= 6 ? reason[code] : "numerical error";
        if (n_bad <= 20)
            rundata.Log() << "Vb::Internal error for voxel " << v + 1 << " : " << why << std::endl;
        if (m_halt_bad_voxel || (st & FABBER_VOX_SETUP_FLAG))
            throw FabberInternalError(why);
    }
    if (n_bad)
        rundata.Log() << "Vb::" << n_bad << " voxels had numerical errors and kept their last state" << std::endl;
}

void Vb::SaveResults(FabberRunData &rundata)
{
    const size_t N = m_nvoxels;
    const int P = m_num_params, T = m_ntimes;
    const std::vector<Parameter> &params = m_model->Params();
    const int NP_all = P + m_noise_params;

    /* noise block of the result MVN: means and (co)variances (OutputAsMVN) */
    auto noise_mean = [&](int i, size_t v) -> double {
        if (m_ar)
            return i < 2 ? m_noise[(size_t)(2 + i) * N + v] : m_noise[0 * N + v] * m_noise[1 * N + v];
        return m_noise[(size_t)(2 * i) * N + v] * m_noise[(size_t)(2 * i + 1) * N + v];
    };
    auto noise_cov = [&](int i, int j, size_t v) -> double {
        if (m_ar)
        {
            if (i < 2 && j < 2)
            {
                const double p11 = m_noise[4 * N + v], p21 = m_noise[5 * N + v], p22 = m_noise[6 * N + v];
                const double det = p11 * p22 - p21 * p21;
                return i == j ? (i == 0 ? p22 / det : p11 / det) : -p21 / det;
            }
            if (i == 2 && j == 2)
                return m_noise[0 * N + v] * m_noise[0 * N + v] * m_noise[1 * N + v];
            return 0.0;
        }
        if (i != j)
            return 0.0;
        const double b = m_noise[(size_t)(2 * i) * N + v], c = m_noise[(size_t)(2 * i + 1) * N + v];
        return b * b * c; /* GammaDist::CalcVariance, dist_gamma.cc:25 */
    };

    if (rundata.GetBool("save-mvn"))
    {
        /* MVNDist::Save, dist_mvn.cc:377-433: packed lower triangle by rows, means, 1 */
        const int n_cov = NP_all * (NP_all + 1) / 2;
        VoxelData &out = rundata.NewVoxelData("finalMVN", n_cov + NP_all + 1);
        for (size_t v = 0; v < N; v++)
        {
            int idx = 0;
            for (int r = 0; r < NP_all; r++)
                for (int c = 0; c <= r; c++, idx++)
                {
                    double val = 0.0;
                    if (r < P)
                        val = m_cov[(size_t)tri(r, c) * N + v];
                    else if (c >= P)
                        val = noise_cov(r - P, c - P, v);
                    out.d[(size_t)idx * N + v] = val;
                }
            for (int i = 0; i < P; i++)
                out.d[(size_t)(n_cov + i) * N + v] = m_mean[(size_t)i * N + v];
            for (int i = 0; i < m_noise_params; i++)
                out.d[(size_t)(n_cov + P + i) * N + v] = noise_mean(i, v);
            out.d[(size_t)(n_cov + NP_all) * N + v] = 1.0;
        }
    }
    const bool s_mean = rundata.GetBool("save-mean"), s_std = rundata.GetBool("save-std"),
               s_z = rundata.GetBool("save-zstat"), s_var = rundata.GetBool("save-var");
    if (s_mean | s_std | s_z | s_var)
        for (int i = 0; i < P; i++)
        {
            /* model space: FwdModel::ToModel on mean and diagonal variance (fwdmodel.cc:326-337) */
            VoxelData *om = s_mean ? &rundata.NewVoxelData("mean_" + params[i].name, 1) : nullptr;
            VoxelData *oz = s_z ? &rundata.NewVoxelData("zstat_" + params[i].name, 1) : nullptr;
            VoxelData *os = s_std ? &rundata.NewVoxelData("std_" + params[i].name, 1) : nullptr;
            VoxelData *ov = s_var ? &rundata.NewVoxelData("var_" + params[i].name, 1) : nullptr;
            const char tr = params[i].transform;
            for (size_t v = 0; v < N; v++)
            {
                const double mean = transform_to_model(tr, m_mean[(size_t)i * N + v]);
                const double var = transform_to_model_var(tr, m_cov[(size_t)tri(i, i) * N + v]);
                const double sd = std::sqrt(var);
                if (om)
                    om->d[v] = mean;
                if (oz)
                    oz->d[v] = mean / sd;
                if (os)
                    os->d[v] = sd;
                if (ov)
                    ov->d[v] = var;
            }
        }
    const bool s_fit = rundata.GetBool("save-model-fit"), s_res = rundata.GetBool("save-residuals");
    if ((s_fit || s_res) && N > 0)
    {
        /* inference.cc:160-239: EvaluateFabber at the posterior means, as one batched device evaluation */
        fabber_cuda_vb_problem prob;
        memset(&prob, 0, sizeof(prob));
        prob.n_voxels = (int)N;
        prob.n_times = T;
        m_model->GetDeviceModel(prob.model);
        for (int i = 0; i < P; i++)
            prob.params[i].transform = params[i].transform;
        DeviceBuf d_mean(m_mean.size() * sizeof(double)), d_fit((size_t)T * N * sizeof(double));
        check(fabber_cuda_memcpy_h2d(d_mean.p, m_mean.data(), m_mean.size() * sizeof(double), nullptr), "model fit");
        check(fabber_cuda_model_fit(&prob, (const double *)d_mean.p, (double *)d_fit.p, nullptr), "model fit");
        std::vector<double> fit((size_t)T * N);
        check(fabber_cuda_memcpy_d2h(fit.data(), d_fit.p, fit.size() * sizeof(double), nullptr), "model fit");
        check(fabber_cuda_stream_sync(nullptr), "model fit");
        if (s_res)
        {
            const VoxelData &data = rundata.GetMainVoxelData();
            VoxelData &res = rundata.NewVoxelData("residuals", T);
            for (size_t i = 0; i < fit.size(); i++)
                res.d[i] = (double)data.f[i] - fit[i];
        }
        if (s_fit)
        {
            VoxelData &mf = rundata.NewVoxelData("modelfit", T);
            mf.d.swap(fit);
        }
    }
    if ((rundata.GetBool("save-noise-mean") | rundata.GetBool("save-noise-std")) && m_noise_params > 0)
    {
        VoxelData *nm = rundata.GetBool("save-noise-mean") ? &rundata.NewVoxelData("noise_means", m_noise_params) : nullptr;
        VoxelData *ns = rundata.GetBool("save-noise-std") ? &rundata.NewVoxelData("noise_stdevs", m_noise_params) : nullptr;
        for (int i = 0; i < m_noise_params; i++)
            for (size_t v = 0; v < N; v++)
            {
                if (nm)
                    nm->d[(size_t)i * N + v] = noise_mean(i, v);
                if (ns)
                    ns->d[(size_t)i * N + v] = std::sqrt(noise_cov(i, i, v));
            }
    }
    if (m_saveF && m_needF)
    {
        VoxelData &f = rundata.NewVoxelData("freeEnergy", 1);
        for (size_t v = 0; v < N; v++)
            f.d[v] = m_F[v];
    }
    if (N > 0 && m_saveFsHistory && m_fhist_len > 0)
    {
        /* one row per pass, plus the final value pushed after the loop (inference_vb.cc:553-554); voxels
         * that stopped early repeat their last value (:1038-1045) */
        int max_its = 0;
        for (size_t v = 0; v < N; v++)
            max_its = std::max(max_its, m_iterations[v]);
        const int rows = std::min(max_its + 1, m_fhist_len);
        VoxelData &h = rundata.NewVoxelData("freeEnergyHistory", rows);
        for (int r = 0; r < rows; r++)
            for (size_t v = 0; v < N; v++)
                h.d[(size_t)r * N + v] = r < m_iterations[v] ? m_Fhist[(size_t)r * N + v] : m_F[v];
    }
    rundata.Log() << "Vb::Done writing results." << std::endl;
}

} // namespace fabber_b200
