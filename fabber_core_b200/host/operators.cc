/*
 * operators.cc - host side of the reference's operator plug-in interface (see operators.h).
 * Behaviour follows convergence.cc:34-378 (detectors), noisemodel_white.cc:95-215 and noisemodel_ar.cc:305-403
 * (options, hard-coded initial distributions), priors.cc:108-219,490-528 (priors and their factory).
 */
#include <algorithm>
#include <cmath>

#include "operators.h"

#include <memory>

namespace fabber_b200
{
/* ---- MVNDist ----------------------------------------------------------------------------------------- */
void MVNDist::SetSize(int n)
{
    m_n = n;
    means.assign(n, 0.0);
    m_cov.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; i++)
        m_cov[(size_t)i * n + i] = 1.0;
    m_prec = m_cov;
    m_cov_valid = m_prec_valid = true;
}
void MVNDist::invert(const std::vector<double> &from, std::vector<double> &to) const
{
    /* Gauss-Jordan with partial pivoting; MVNDist's "+1e-10 I and retry" (dist_mvn.cc:215-224) */
    const int n = m_n;
    for (int attempt = 0; attempt < 2; attempt++)
    {
        std::vector<double> a(from), inv((size_t)n * n, 0.0);
        for (int i = 0; i < n; i++)
        {
            inv[(size_t)i * n + i] = 1.0;
            if (attempt)
                a[(size_t)i * n + i] += 1e-10;
        }
        bool singular = false;
        for (int c = 0; c < n && !singular; c++)
        {
            int p = c;
            for (int r = c + 1; r < n; r++)
                if (std::fabs(a[(size_t)r * n + c]) > std::fabs(a[(size_t)p * n + c]))
                    p = r;
            if (a[(size_t)p * n + c] == 0.0 || !std::isfinite(a[(size_t)p * n + c]))
            {
                singular = true;
                break;
            }
            for (int k = 0; k < n; k++)
            {
                std::swap(a[(size_t)c * n + k], a[(size_t)p * n + k]);
                std::swap(inv[(size_t)c * n + k], inv[(size_t)p * n + k]);
            }
            const double d = a[(size_t)c * n + c];
            for (int k = 0; k < n; k++)
            {
                a[(size_t)c * n + k] /= d;
                inv[(size_t)c * n + k] /= d;
            }
            for (int r = 0; r < n; r++)
                if (r != c)
                {
                    const double f = a[(size_t)r * n + c];
                    for (int k = 0; k < n; k++)
                    {
                        a[(size_t)r * n + k] -= f * a[(size_t)c * n + k];
                        inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k];
                    }
                }
        }
        if (!singular)
        {
            to = inv;
            return;
        }
    }
    throw FabberInternalError("MVNDist: matrix is singular");
}
double MVNDist::GetPrecisions(int i, int j) const
{
    if (!m_prec_valid)
    {
        invert(m_cov, m_prec);
        m_prec_valid = true;
    }
    return m_prec[(size_t)i * m_n + j];
}
void MVNDist::SetCovariance(int i, int j, double v)
{
    if (!m_cov_valid)
    {
        invert(m_prec, m_cov);
        m_cov_valid = true;
    }
    m_cov[(size_t)i * m_n + j] = m_cov[(size_t)j * m_n + i] = v;
    m_prec_valid = false;
}
void MVNDist::SetPrecisions(int i, int j, double v)
{
    GetPrecisions(0 < m_n ? 0 : 0, 0 < m_n ? 0 : 0); /* make the precision form current */
    m_prec[(size_t)i * m_n + j] = m_prec[(size_t)j * m_n + i] = v;
    m_cov_valid = false;
    /* covariance is refreshed lazily by the next reader */
    const_cast<MVNDist *>(this)->m_cov_valid = false;
}

/* special functions of the ARD free-energy term (tools.cc:87-98; MISCMATHS::digamma in single precision) */
double gammaln_host(double x)
{
    static const double cof[6] = { 76.18009172947146, -86.50532032941677, 24.01409824083091, -1.231739572450155,
        0.1208650973866179e-2, -0.5395239384953e-5 };
    double series = 1.000000000190015;
    for (int j = 0; j < 6; j++)
        series += cof[j] / (x + 1.0 + j);
    return std::log(2.5066282746310005 * series / x) + (x + 0.5) * std::log(x + 5.5) - x - 5.5;
}
double digamma_fsl_host(double xin)
{
    float y = (float)xin, acc = 0.0f;
    if (y <= 1e-5f)
        return (double)(-0.5772156649f - 1.0f / y);
    for (; y < 8.5f; y += 1.0f)
        acc -= 1.0f / y;
    float r = 1.0f / y;
    acc = (float)((double)acc + (double)(float)std::log((double)y) - 0.5 * (double)r);
    r *= r;
    acc -= r * (8.333333333e-2f - r * (8.333333333e-3f - r * 3.968253968e-3f));
    return (double)acc;
}

/* ---- convergence detectors ------------------------------------------------------------------------------ */
static const char *const DETECTOR_NAMES[] = { "maxits", "pointzeroone", "freduce", "trialmode", "lm" };
std::vector<std::string> ConvergenceDetector::GetKnown()
{
    return std::vector<std::string>(DETECTOR_NAMES, DETECTOR_NAMES + 5);
}
ConvergenceDetector *ConvergenceDetector::NewFromName(const std::string &name)
{
    if (name == "maxits")
        return new CountingConvergenceDetector();
    if (name == "pointzeroone")
        return new FchangeConvergenceDetector();
    if (name == "freduce")
        return new FreduceConvergenceDetector();
    if (name == "trialmode")
        return new TrialModeConvergenceDetector();
    if (name == "lm")
        return new LMConvergenceDetector();
    throw InvalidOptionValue("convergence", name, "Unrecognized convergence detector");
}
void ConvergenceDetector::Initialize(FabberRunData &) {}
void ConvergenceDetector::Describe(fabber_cuda_vb_problem &prob) const
{
    prob.conv_type = DeviceCode();
    prob.max_iterations = m_max_its - (DeviceCode() == FABBER_CONV_TRIALMODE ? 1 : 0); /* the kernel adds the 1 back */
    prob.fchange = m_fchange;
    prob.max_trials = m_max_trials;
}

void CountingConvergenceDetector::Initialize(FabberRunData &params)
{
    m_max_its = params.GetIntDefault("max-iterations", 10);
    if (m_max_its <= 0)
        throw InvalidOptionValue("max_iterations", stringify(m_max_its), "Must be positive"); /* sic, convergence.cc:39 */
    Reset();
}
void CountingConvergenceDetector::Reset(double)
{
    m_its = 0;
    m_reason = "";
}
bool CountingConvergenceDetector::Test(double)
{
    if (++m_its < m_max_its)
        return false;
    m_reason = "Max iterations reached";
    return true;
}

void FchangeConvergenceDetector::Initialize(FabberRunData &params)
{
    CountingConvergenceDetector::Initialize(params);
    m_fchange = params.GetDoubleDefault("min-fchange", 0.01);
    if (!(m_fchange > 0))
        throw InvalidOptionValue("min-fchange", stringify(m_fchange), "Must be positive");
    Reset();
}
void FchangeConvergenceDetector::Reset(double F)
{
    CountingConvergenceDetector::Reset();
    m_prev_f = F;
    m_save = m_revert = false;
}
bool FchangeConvergenceDetector::Test(double F)
{
    const double change = std::fabs(F - m_prev_f);
    m_prev_f = F;
    if (change < m_fchange)
    {
        m_reason = "Absolute difference less than minimum";
        return true;
    }
    return CountingConvergenceDetector::Test(F);
}
bool FreduceConvergenceDetector::Test(double F)
{
    if (F - m_prev_f < 0) /* F went down: stop and go back (to the pre-loop copies) */
    {
        m_reason = "F reduced";
        m_revert = true;
        return true;
    }
    return FchangeConvergenceDetector::Test(F);
}

void TrialModeConvergenceDetector::Initialize(FabberRunData &params)
{
    FchangeConvergenceDetector::Initialize(params);
    m_max_its += 1; /* convergence.cc:145: one more pass than asked for */
    m_max_trials = params.GetIntDefault("max-trials", 10);
    if (m_max_trials <= 0)
        throw InvalidOptionValue("max-trials", stringify(m_max_trials), "Must be positive");
    Reset();
}
void TrialModeConvergenceDetector::Reset(double)
{
    FchangeConvergenceDetector::Reset();
    m_trials = 0;
    m_trialmode = false;
    m_save = true;
}
bool TrialModeConvergenceDetector::Test(double F)
{
    const double diff = F - m_prev_f;
    const bool tiny = std::fabs(diff) < m_fchange;
    if (!m_trialmode)
    {
        if (diff < 0) /* first drop of F: start trialling from the saved state */
        {
            m_its = m_trials = 1;
            m_trialmode = m_revert = true;
            m_save = false;
            return false;
        }
        m_revert = false;
        if (tiny)
        {
            m_reason = "F increased by less than tolerance";
            m_save = false;
            return true;
        }
        m_save = true;
        m_prev_f = F;
        return ++m_its >= m_max_its;
    }
    ++m_trials;
    if (diff > 0)
    {
        m_revert = false;
        if (tiny)
        {
            m_reason = "F increased by less than tolerance during trial mode";
            m_save = false;
            return true;
        }
        m_trialmode = false; /* F recovered: carry on normally from here */
        m_trials = 0;
        m_save = true;
        m_prev_f = F;
        return false;
    }
    m_save = false;
    m_revert = m_trials >= m_max_trials;
    if (m_revert)
        m_reason = "Reached max trials";
    return m_revert;
}

void LMConvergenceDetector::Initialize(FabberRunData &params)
{
    m_max_its = params.GetIntDefault("max-iterations", 10);
    if (m_max_its <= 0)
        throw InvalidOptionValue("max-iterations", stringify(m_max_its), "Must be positive");
    m_fchange = params.GetDoubleDefault("max-fchange", 0.01);
    if (!(m_fchange > 0))
        throw InvalidOptionValue("max-fchange", stringify(m_fchange), "Must be positive");
    Reset();
}
void LMConvergenceDetector::Reset(double F)
{
    m_its = 0;
    m_prev_f = F;
    m_save = true; /* never changes afterwards: the snapshot is overwritten every pass */
    m_revert = m_LM = false;
    m_alpha = 0.0;
}
bool LMConvergenceDetector::Test(double F)
{
    const double diff = F - m_prev_f;
    if (!m_LM)
    {
        if (diff < 0) /* F dropped: retry with damping */
        {
            m_LM = m_revert = true;
            m_alpha = m_alphastart;
            return false;
        }
        m_revert = false;
        if (std::fabs(diff) < m_fchange)
        {
            m_reason = "F converged";
            return true;
        }
        if (m_its >= m_max_its)
        {
            m_reason = "Max iterations reached";
            return true;
        }
        m_prev_f = F;
        ++m_its;
        return false;
    }
    if (diff > 0) /* the damped step helped: relax the damping, accept the step */
    {
        if (m_alpha == m_alphastart)
            m_LM = false;
        else
            m_alpha /= 10;
        m_revert = false;
        m_prev_f = F;
        ++m_its;
        return false;
    }
    if (m_alpha >= m_alphamax)
    {
        m_reason = "Reached max m_alpha";
        m_revert = true;
        return true;
    }
    if (m_its >= m_max_its)
    {
        m_reason = "Max iterations reached";
        m_revert = false;
        return true;
    }
    m_alpha *= 10;
    m_revert = true;
    return false;
}

/* ---- noise models ------------------------------------------------------------------------------------------ */
MVNDist NoiseParams::OutputAsMVN() const
{
    const int na = alpha.GetSize(), n = na + (int)phis.size();
    MVNDist out(n);
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++)
            out.SetCovariance(i, j, 0.0);
    for (int i = 0; i < na; i++)
    {
        out.means[i] = alpha.means[i];
        for (int j = 0; j <= i; j++)
            out.SetCovariance(i, j, alpha.GetCovariance(i, j));
    }
    for (size_t i = 0; i < phis.size(); i++)
    {
        out.means[na + i] = phis[i].CalcMean();
        out.SetCovariance(na + (int)i, na + (int)i, phis[i].CalcVariance());
    }
    return out;
}
std::vector<std::string> NoiseModel::GetKnown()
{
    std::vector<std::string> k;
    k.push_back("ar");
    k.push_back("white");
    return k;
}
NoiseModel *NoiseModel::NewFromName(const std::string &name)
{
    if (name == "white")
        return new WhiteNoiseModel();
    if (name == "ar")
        return new Ar1cNoiseModel();
    throw InvalidOptionValue("noise", name, "Unrecognized noise type"); /* noisemodel.cc:29 */
}
void NoiseModel::Initialize(FabberRunData &args) { m_masked_tpoints = args.GetIntList("mt", 1); }
static void device_only(const char *what)
{
    throw FabberInternalError(std::string(what)
        + " runs inside the CUDA kernels (fabber_cuda_vb_*): this library has no CPU inference path");
}
void NoiseModel::UpdateNoise() const { device_only("NoiseModel::UpdateNoise"); }
void NoiseModel::UpdateTheta() const { device_only("NoiseModel::UpdateTheta"); }
double NoiseModel::CalcFreeEnergy() const
{
    device_only("NoiseModel::CalcFreeEnergy");
    return 0;
}

void WhiteNoiseModel::Initialize(FabberRunData &args)
{
    NoiseModel::Initialize(args);
    m_pattern = args.GetStringDefault("noise-pattern", "1"); /* noisemodel_white.cc:166-215 */
    m_digits.clear();
    for (size_t i = 0; i < m_pattern.size(); i++)
    {
        const char ch = m_pattern[i];
        if (ch >= '1' && ch <= '9')
            m_digits.push_back(ch - '0');
        else if (ch >= 'A' && ch <= 'Z')
            m_digits.push_back(ch - 'A' + 10);
        else if (ch >= 'a' && ch <= 'z')
            m_digits.push_back(ch - 'a' + 10);
        else
            throw InvalidOptionValue("noise-pattern", std::string(1, ch), "Invalid character"); /* noisemodel_white.cc:191 */
    }
    if (m_digits.empty())
        throw InvalidOptionValue("noise-pattern", m_pattern, "Pattern must not be empty");
    m_nphis = *std::max_element(m_digits.begin(), m_digits.end());
    if (m_nphis > FABBER_CUDA_MAX_PHIS)
        throw InvalidOptionValue("noise-pattern", m_pattern, "more noise precisions than the device kernels carry");
    m_phi_prior = args.GetDoubleDefault("prior-noise-stddev", -1);
    if (m_phi_prior < 0 && m_phi_prior != -1)
        throw InvalidOptionValue("prior-noise-stddev", stringify(m_phi_prior), "Must be > 0");
    m_locked = args.GetDoubleDefault("locked-noise-stdev", -1);
}
NoiseParams *WhiteNoiseModel::NewParams() const
{
    NoiseParams *p = new NoiseParams();
    p->phis.resize(m_nphis);
    return p;
}
void WhiteNoiseModel::HardcodedInitialDists(NoiseParams &prior, NoiseParams &posterior) const
{
    prior.phis.resize(m_nphis);
    posterior.phis.resize(m_nphis);
    for (int i = 0; i < m_nphis; i++)
    {
        if (m_phi_prior == -1) /* noisemodel_white.cc:142-149 */
        {
            prior.phis[i].b = 1e6;
            prior.phis[i].c = 1e-6;
            posterior.phis[i].b = 1e-8;
            posterior.phis[i].c = 50;
        }
        else /* :156-161 */
        {
            prior.phis[i].c = posterior.phis[i].c = 0.5;
            prior.phis[i].b = posterior.phis[i].b = 1 / (m_phi_prior * m_phi_prior * 0.5);
        }
    }
}
void WhiteNoiseModel::Describe(fabber_cuda_vb_problem &prob, int n_times, std::vector<unsigned char> &pattern) const
{
    prob.noise_type = FABBER_NOISE_WHITE;
    prob.n_phis = m_nphis;
    pattern.assign(n_times, 0);
    for (int t = 0; t < n_times; t++)
        pattern[t] = (unsigned char)(m_digits[t % m_digits.size()] - 1);
    NoiseParams prior, post;
    HardcodedInitialDists(prior, post);
    for (int i = 0; i < m_nphis; i++)
    {
        prob.noise_prior_b[i] = prior.phis[i].b;
        prob.noise_prior_c[i] = prior.phis[i].c;
        prob.noise_post_b[i] = post.phis[i].b;
        prob.noise_post_c[i] = post.phis[i].c;
    }
    prob.locked_noise_stdev = m_locked;
}

void Ar1cNoiseModel::Initialize(FabberRunData &args)
{
    NoiseModel::Initialize(args);
    /* noisemodel_ar.cc:305-349 */
    m_nphis = args.GetIntDefault("num-echoes", 1);
    m_type = args.GetStringDefault("ar1-cross-terms", "none");
    NumAlphas(); /* validates the type */
    if (m_nphis == 1 && m_type != "none")
        throw InvalidOptionValue("ar1-cross-terms", m_type, "You must use ar1-cross-terms=none with num-echoes=1");
    if (m_nphis != 1 && m_nphis != 2)
        throw InvalidOptionValue("num-echoes", stringify(m_nphis), "Must be 1 or 2");
    if (!m_masked_tpoints.empty())
        throw InvalidOptionValue("mt1", "", "Masked time points are not supported for the AR noise model");
}
int Ar1cNoiseModel::NumAlphas() const
{
    if (m_type == "none")
        return 2;
    if (m_type == "same")
        return 3;
    if (m_type == "dual")
        return 4;
    throw InvalidOptionValue("ar1-cross-terms", m_type, "Must be dual, same or none");
}
NoiseParams *Ar1cNoiseModel::NewParams() const
{
    NoiseParams *p = new NoiseParams();
    p->phis.resize(m_nphis);
    p->alpha.SetSize(NumAlphas());
    return p;
}
void Ar1cNoiseModel::HardcodedInitialDists(NoiseParams &prior, NoiseParams &posterior) const
{
    /* noisemodel_ar.cc:379-403 */
    const int na = NumAlphas();
    NoiseParams *both[2] = { &prior, &posterior };
    for (int w = 0; w < 2; w++)
    {
        both[w]->alpha.SetSize(na);
        for (int i = 0; i < na; i++)
            both[w]->alpha.SetCovariance(i, i, 1e4); /* precision 1e-4 */
        both[w]->phis.resize(m_nphis);
    }
    for (int i = 0; i < m_nphis; i++)
    {
        prior.phis[i].b = 1e6;
        prior.phis[i].c = 1e-6;
        posterior.phis[i].b = 1e-8;
        posterior.phis[i].c = 1e-6;
    }
}
void Ar1cNoiseModel::Describe(fabber_cuda_vb_problem &prob, int n_times, std::vector<unsigned char> &pattern) const
{
    prob.noise_type = FABBER_NOISE_AR1;
    prob.n_phis = m_nphis;
    prob.ar_cross_terms = NumAlphas() - 2; /* FABBER_AR_CROSS_NONE / SAME / DUAL */
    /* two echoes interleave (noisemodel_ar.cc:126-129): with an odd series the reference's alpha matrices and its
     * data vector disagree in size and NEWMAT throws on the first product */
    if (m_nphis == 2 && (n_times % 2 != 0 || n_times < 4))
        throw InvalidOptionValue("num-echoes", stringify(m_nphis),
            "the data must hold an even number (at least 4) of time points, the two echoes interleaved");
    pattern.assign(n_times, 0);
    NoiseParams prior, post;
    HardcodedInitialDists(prior, post);
    for (int i = 0; i < m_nphis; i++)
    {
        prob.noise_prior_b[i] = prior.phis[i].b;
        prob.noise_prior_c[i] = prior.phis[i].c;
        prob.noise_post_b[i] = post.phis[i].b;
        prob.noise_post_c[i] = post.phis[i].c;
    }
    prob.ar_alpha_prior_prec = prior.alpha.GetPrecisions(0, 0);
}

/* ---- inference techniques ---------------------------------------------------------------------------------------- */
std::vector<std::string> InferenceTechnique::GetKnown() { return Vb::GetKnownMethods(); }
InferenceTechnique *InferenceTechnique::NewFromName(const std::string &name)
{
    if (name == "vb" || name == "spatialvb")
        return new VariationalBayesInferenceTechnique();
    if (name == "nlls")
        return new NLLSInferenceTechnique();
    throw InvalidOptionValue("method", name, "Unrecognized inference method (vb, spatialvb, nlls)");
}
void InferenceTechnique::UsageFromName(const std::string &name, std::ostream &stream)
{
    std::unique_ptr<InferenceTechnique> t(NewFromName(name));
    stream << "Usage information for method: " << name << std::endl << std::endl;
    stream << t->GetDescription() << std::endl << std::endl << "Options: " << std::endl << std::endl;
    std::vector<OptionSpec> options;
    t->GetOptions(options);
    for (size_t i = 0; i < options.size(); i++)
        stream << "  --" << options[i].name << (options[i].optional ? " [optional]" : "") << ": " << options[i].description
               << std::endl;
}
std::string InferenceTechnique::GetVersion() const { return "fabber_core_b200"; }
void InferenceTechnique::Initialize(FwdModel *fwd_model, FabberRunData &args) { m_engine.Initialize(fwd_model, args); }
void InferenceTechnique::DoCalculations(FabberRunData &rundata) { m_engine.DoCalculations(rundata); }
void InferenceTechnique::SaveResults(FabberRunData &rundata) { m_engine.SaveResults(rundata); }
void NLLSInferenceTechnique::Initialize(FwdModel *fwd_model, FabberRunData &args)
{
    if (args.GetStringDefault("method", "nlls") != "nlls")
        throw InvalidOptionValue("method", args.GetString("method"), "NLLSInferenceTechnique runs method=nlls");
    args.Set("method", "nlls");
    InferenceTechnique::Initialize(fwd_model, args);
}

/* ---- priors --------------------------------------------------------------------------------------------------- */
DefaultPrior::DefaultPrior(const Parameter &p)
    : m_param_name(p.name)
    , m_idx(p.idx)
    , m_type_code(p.prior_type)
    , m_params(p.prior)
{
}
double DefaultPrior::ApplyToMVN(MVNDist *prior, const RunContext &)
{
    prior->means[m_idx] = m_params.mean();
    prior->SetPrecisions((int)m_idx, (int)m_idx, m_params.prec());
    return 0;
}
ImagePrior::ImagePrior(const Parameter &p, FabberRunData &rundata)
    : DefaultPrior(p)
{
    m_filename = p.options.find("image")->second;
    const VoxelData &img = rundata.GetVoxelData(m_filename);
    m_image.resize(img.cols);
    for (size_t v = 0; v < img.cols; v++)
        m_image[v] = img.at(0, v);
}
double ImagePrior::ApplyToMVN(MVNDist *prior, const RunContext &ctx)
{
    prior->means[m_idx] = m_image.at(ctx.v - 1);
    prior->SetPrecisions((int)m_idx, (int)m_idx, m_params.prec());
    return 0;
}
double ARDPrior::ApplyToMVN(MVNDist *prior, const RunContext &ctx)
{
    const MVNDist &post = ctx.fwd_post.at(ctx.v - 1);
    const double m = post.means[m_idx];
    const double second_moment = m * m + post.GetCovariance((int)m_idx, (int)m_idx); /* Chappell 2009 eq. D4 */
    if (ctx.it == 0)
    {
        prior->SetCovariance((int)m_idx, (int)m_idx, m_params.var());
        prior->means[m_idx] = m_params.mean();
    }
    else
        prior->SetCovariance((int)m_idx, (int)m_idx, second_moment);
    const double b = 2 / second_moment;
    return -1.5 * (std::log(b) + digamma_fsl_host(0.5)) - 0.5 - gammaln_host(0.5) - 0.5 * std::log(b);
}
SpatialPrior::SpatialPrior(const Parameter &p, FabberRunData &rundata)
    : DefaultPrior(p)
{
    /* priors.cc:183-219 */
    m_spatial_dims = rundata.GetIntDefault("spatial-dims", 3);
    if (m_spatial_dims < 0 || m_spatial_dims > 3)
        throw InvalidOptionValue("spatial-dims", stringify(m_spatial_dims), "Must be 0, 1, 2 or 3");
    m_spatial_speed = rundata.GetDoubleDefault("spatial-speed", -1); /* range unchecked, as in the reference */
    m_q1 = rundata.GetDoubleDefault("spatial-q1", 10.0);
    m_q2 = rundata.GetDoubleDefault("spatial-q2", 1.0);
    m_update_first_iter = rundata.GetBool("update-spatial-prior-on-first-iteration");
}
double SpatialPrior::ApplyToMVN(MVNDist *, const RunContext &)
{
    device_only("SpatialPrior::ApplyToMVN (MRF prior mean and precision, aK)");
    return 0;
}
void SpatialPrior::Describe(fabber_cuda_vb_problem &prob) const
{
    prob.spatial_dims = m_spatial_dims;
    prob.spatial_speed = m_spatial_speed;
    prob.spatial_q1 = m_q1;
    prob.spatial_q2 = m_q2;
    prob.update_first_iter = m_update_first_iter ? 1 : 0;
}
Prior *PriorFactory::CreatePrior(Parameter p)
{
    switch (p.prior_type) /* priors.cc:509-528 */
    {
    case 'N':
    case '-':
        p.prior_type = 'N';
        return new DefaultPrior(p);
    case 'I':
        return new ImagePrior(p, m_rundata);
    case 'A':
        return new ARDPrior(p, m_rundata);
    case 'M':
    case 'm':
    case 'P':
    case 'p':
        return new SpatialPrior(p, m_rundata);
    default:
        throw InvalidOptionValue("Prior type", std::string(1, p.prior_type), "Supported types: NMmPpAI");
    }
}
std::vector<Prior *> PriorFactory::CreatePriors(const std::vector<Parameter> &params)
{
    std::vector<Prior *> priors;
    try
    {
        for (size_t i = 0; i < params.size(); i++)
            priors.push_back(CreatePrior(params[i]));
    }
    catch (...)
    {
        for (size_t i = 0; i < priors.size(); i++)
            delete priors[i];
        throw;
    }
    return priors;
}

} // namespace fabber_b200
