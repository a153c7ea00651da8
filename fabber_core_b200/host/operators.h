/*
 * operators.h - host side of the reference's operator plug-in interface for the VB path:
 *   MVNDist / GammaDist / RunContext   dist_mvn.h, dist_gamma.h, run_context.h   (value types)
 *   ConvergenceDetector (+ 5 detectors) convergence.h:21-318, registry names of setup.cc:49-57
 *   NoiseModel (WhiteNoiseModel, Ar1cNoiseModel), NoiseParams   noisemodel.h:26-182, setup.cc:36-40
 *   Prior (Default / Image / ARD / Spatial), PriorFactory        priors.h:29-152
 * Same class names, virtual signatures, registry names and option keys as the reference, so code written
 * against fabber_core (a registration in setup.cc style, a test in test/test_convergence.cc or
 * test/test_priors.cc style) compiles against this header.
 *
 * What is DIFFERENT, by design: the arithmetic of the update loop lives in the CUDA kernels. A noise model, a
 * prior type and a convergence detector are therefore, for the device, a CODE (FABBER_NOISE_*, 'N'/'I'/'A'/
 * 'M'/.., FABBER_CONV_*) plus constants; every class here knows its code (DeviceCode / Describe) and
 * Vb::Prepare asks these classes - not string compares of its own - when it fills the plain-C problem
 * description. The cheap, data-independent operators are ALSO implemented on the host with the reference's
 * exact semantics (the detectors' state machines, Default / Image / ARD ApplyToMVN, the noise models' hard-coded
 * initial distributions): that is what the reference's own unit tests exercise. The per-voxel heavy operators
 * (UpdateNoise, UpdateTheta, CalcFreeEnergy, SpatialPrior::ApplyToMVN over a volume) have NO host implementation
 * - there is no CPU inference path in this library - and say so if called.
 */
#pragma once
#include <ostream>
#include <string>
#include <vector>

#include "fabber_host.h"

namespace fabber_b200
{
/* ---- value types ------------------------------------------------------------------------------------ */
/* dist_mvn.h: means + covariance, precision on demand. Dense row-major, 0-based (the reference is 1-based). */
class MVNDist
{
public:
    explicit MVNDist(int n = 0) { SetSize(n); }
    void SetSize(int n); /* means 0, covariance = precisions = identity (dist_mvn.cc:102-118) */
    int GetSize() const { return m_n; }
    std::vector<double> means;
    double GetCovariance(int i, int j) const { return m_cov[(size_t)i * m_n + j]; }
    double GetPrecisions(int i, int j) const;
    void SetCovariance(int i, int j, double v);
    void SetPrecisions(int i, int j, double v);

private:
    int m_n = 0;
    mutable std::vector<double> m_cov, m_prec;
    mutable bool m_cov_valid = true, m_prec_valid = true;
    void invert(const std::vector<double> &from, std::vector<double> &to) const;
};
struct GammaDist /* dist_gamma.h: b = scale, c = shape */
{
    double b = 1, c = 1;
    double CalcMean() const { return b * c; }
    double CalcVariance() const { return b * b * c; }
    void SetMeanVariance(double m, double v)
    {
        b = v / m;
        c = m / b;
    }
};
struct RunContext /* run_context.h:23-50 */
{
    int it = 0, v = 1, nvoxels = 0; /* v is 1-based, as in the reference */
    std::vector<int> ignore_voxels;
    std::vector<MVNDist> fwd_prior, fwd_post;
    std::vector<std::vector<int>> neighbours, neighbours2;
};

/* ---- convergence detectors (convergence.h) -------------------------------------------------------------- */
class ConvergenceDetector
{
public:
    static ConvergenceDetector *NewFromName(const std::string &name); /* maxits pointzeroone freduce trialmode lm */
    static std::vector<std::string> GetKnown();
    virtual ~ConvergenceDetector() {}
    virtual void Initialize(FabberRunData &params);
    virtual bool Test(double F) = 0;
    virtual void Reset(double F = -99e99) = 0;
    virtual bool UseF() const { return false; }
    virtual bool NeedSave() { return false; }
    virtual bool NeedRevert() { return false; }
    virtual float LMalpha() { return 0.0; }
    std::string GetReason() { return m_reason; }
    /* the kernels' code for this detector and the constants they need (include/fabber_cuda.h) */
    virtual int DeviceCode() const = 0;
    virtual void Describe(fabber_cuda_vb_problem &prob) const;

protected:
    std::string m_reason;
    int m_max_its = 10, m_max_trials = 10;
    double m_fchange = 0.01;
};
class CountingConvergenceDetector : public ConvergenceDetector
{
public:
    void Initialize(FabberRunData &params) override;
    bool Test(double F) override;
    void Reset(double F = -99e99) override;
    int DeviceCode() const override { return FABBER_CONV_MAXITS; }

protected:
    int m_its = 0;
};
class FchangeConvergenceDetector : public CountingConvergenceDetector
{
public:
    void Initialize(FabberRunData &params) override;
    bool Test(double F) override;
    void Reset(double F = -99e99) override;
    bool UseF() const override { return true; }
    bool NeedSave() override { return m_save; }
    bool NeedRevert() override { return m_revert; }
    int DeviceCode() const override { return FABBER_CONV_FCHANGE; }

protected:
    double m_prev_f = -99e99;
    bool m_save = false, m_revert = false;
};
class FreduceConvergenceDetector : public FchangeConvergenceDetector
{
public:
    bool Test(double F) override;
    int DeviceCode() const override { return FABBER_CONV_FREDUCE; }
};
class TrialModeConvergenceDetector : public FchangeConvergenceDetector
{
public:
    void Initialize(FabberRunData &params) override;
    bool Test(double F) override;
    void Reset(double F = -99e99) override;
    int DeviceCode() const override { return FABBER_CONV_TRIALMODE; }

protected:
    int m_trials = 0;
    bool m_trialmode = false;
};
class LMConvergenceDetector : public FchangeConvergenceDetector
{
public:
    void Initialize(FabberRunData &params) override;
    bool Test(double F) override;
    void Reset(double F = -99e99) override;
    float LMalpha() override { return (float)m_alpha; }
    int DeviceCode() const override { return FABBER_CONV_LM; }

protected:
    double m_alpha = 0, m_alphastart = 1e-6, m_alphamax = 1e6;
    bool m_LM = false;
};

/* ---- inference techniques (inference.h:22-166, registry names of setup.cc:28-33) --------------------------------
 * Same factory and virtuals as the reference. All three names run on the device behind one engine (class Vb of
 * fabber_host.h: block-wise upload, one voxel range or z-slab per GPU, device-side SaveResults); "vb" / "spatialvb"
 * are the reference's Vb, "nlls" its NLLSInferenceTechnique (csrc/vb_nlls.cuh). */
class InferenceTechnique
{
public:
    static std::vector<std::string> GetKnown();                      /* nlls, spatialvb, vb */
    static InferenceTechnique *NewFromName(const std::string &name); /* throws InvalidOptionValue("method", ..) */
    static void UsageFromName(const std::string &name, std::ostream &stream);
    virtual ~InferenceTechnique() {}
    virtual void GetOptions(std::vector<OptionSpec> &opts) const = 0;
    virtual std::string GetDescription() const = 0;
    virtual std::string GetVersion() const;
    virtual void Initialize(FwdModel *fwd_model, FabberRunData &args);
    virtual void DoCalculations(FabberRunData &rundata);
    virtual void SaveResults(FabberRunData &rundata);

protected:
    Vb m_engine;
};
class VariationalBayesInferenceTechnique : public InferenceTechnique /* "vb", "spatialvb": inference_vb.h */
{
public:
    void GetOptions(std::vector<OptionSpec> &opts) const override { Vb::GetOptions(opts, "vb"); }
    std::string GetDescription() const override { return Vb::GetDescription("vb"); }
};
class NLLSInferenceTechnique : public InferenceTechnique /* "nlls": inference_nlls.h */
{
public:
    void GetOptions(std::vector<OptionSpec> &opts) const override { Vb::GetOptions(opts, "nlls"); }
    std::string GetDescription() const override { return Vb::GetDescription("nlls"); }
    void Initialize(FwdModel *fwd_model, FabberRunData &args) override; /* insists on method=nlls in args */
};

/* ---- noise models (noisemodel.h) ------------------------------------------------------------------------- */
class NoiseParams /* noisemodel.h:26-58; white: one Gamma per phi, AR(1): alpha MVN + phis */
{
public:
    std::vector<GammaDist> phis;
    MVNDist alpha;
    MVNDist OutputAsMVN() const; /* noisemodel_white.cc:55-68, noisemodel_ar.cc:287-300 */
};
class NoiseModel
{
public:
    static NoiseModel *NewFromName(const std::string &name); /* "white", "ar" (setup.cc:36-40) */
    static std::vector<std::string> GetKnown();
    virtual ~NoiseModel() {}
    virtual void Initialize(FabberRunData &args); /* masked time points, inference.cc:96-103 */
    virtual NoiseParams *NewParams() const = 0;
    virtual void HardcodedInitialDists(NoiseParams &prior, NoiseParams &posterior) const = 0;
    virtual int NumParams() = 0;
    /* noise type, phi pattern, prior / initial Gammas, AR alpha precision into the plain-C problem. `pattern`
     * ([T], filled here) must outlive the launch. */
    virtual void Describe(fabber_cuda_vb_problem &prob, int n_times, std::vector<unsigned char> &pattern) const = 0;
    /* UpdateNoise / UpdateTheta / CalcFreeEnergy run inside the kernels (csrc/vb_voxelwise*.cuh); the host
     * classes have no CPU arithmetic behind them */
    void UpdateNoise() const;
    void UpdateTheta() const;
    double CalcFreeEnergy() const;
    const std::vector<int> &MaskedTimepoints() const { return m_masked_tpoints; }

protected:
    std::vector<int> m_masked_tpoints;
};
class WhiteNoiseModel : public NoiseModel
{
public:
    void Initialize(FabberRunData &args) override;
    NoiseParams *NewParams() const override;
    void HardcodedInitialDists(NoiseParams &prior, NoiseParams &posterior) const override;
    int NumParams() override { return (int)m_digits.empty() ? 0 : m_nphis; }
    void Describe(fabber_cuda_vb_problem &prob, int n_times, std::vector<unsigned char> &pattern) const override;

private:
    std::string m_pattern;
    std::vector<int> m_digits;
    int m_nphis = 1;
    double m_phi_prior = -1, m_locked = -1;
};
class Ar1cNoiseModel : public NoiseModel
{
public:
    void Initialize(FabberRunData &args) override;
    NoiseParams *NewParams() const override;
    void HardcodedInitialDists(NoiseParams &prior, NoiseParams &posterior) const override;
    int NumParams() override { return m_nphis; } /* quirk kept: nPhis, although the MVN has alphas too */
    int NumAlphas() const;
    void Describe(fabber_cuda_vb_problem &prob, int n_times, std::vector<unsigned char> &pattern) const override;

private:
    int m_nphis = 1;
    std::string m_type = "none";
};

/* ---- priors (priors.h) --------------------------------------------------------------------------------------- */
class Prior
{
public:
    virtual ~Prior() {}
    virtual double ApplyToMVN(MVNDist *prior, const RunContext &ctx) = 0;
    virtual char DeviceCode() const = 0; /* 'N' 'I' 'A' 'M' 'm' 'P' 'p' */
    static std::string ExpandPriorTypesString(std::string priors_str, unsigned num_params)
    {
        return fabber_b200::ExpandPriorTypesString(priors_str, num_params);
    }
};
class DefaultPrior : public Prior
{
public:
    explicit DefaultPrior(const Parameter &param);
    std::string m_param_name;
    unsigned m_idx;
    char m_type_code;
    DistParams m_params;
    double ApplyToMVN(MVNDist *prior, const RunContext &ctx) override;
    char DeviceCode() const override { return m_type_code; }
};
class ImagePrior : public DefaultPrior
{
public:
    ImagePrior(const Parameter &param, FabberRunData &rundata);
    double ApplyToMVN(MVNDist *prior, const RunContext &ctx) override;

protected:
    std::string m_filename;
    std::vector<double> m_image;
};
class ARDPrior : public DefaultPrior
{
public:
    ARDPrior(const Parameter &param, FabberRunData &)
        : DefaultPrior(param)
    {
    }
    double ApplyToMVN(MVNDist *prior, const RunContext &ctx) override;
};
class SpatialPrior : public DefaultPrior
{
public:
    SpatialPrior(const Parameter &param, FabberRunData &rundata); /* spatial-dims / speed / q1 / q2 validation */
    /* the MRF / Penny prior couples every voxel with its neighbours and its aK is a reduction over the volume:
     * device only (csrc/vb_spatial.cuh) - throws FabberInternalError */
    double ApplyToMVN(MVNDist *prior, const RunContext &ctx) override;
    void Describe(fabber_cuda_vb_problem &prob) const;

protected:
    double m_aK = 1e-8; /* priors.cc:185 */
    int m_spatial_dims = 3;
    double m_spatial_speed = -1, m_q1 = 10, m_q2 = 1;
    bool m_update_first_iter = false;
};
class PriorFactory
{
public:
    explicit PriorFactory(FabberRunData &rundata)
        : m_rundata(rundata)
    {
    }
    std::vector<Prior *> CreatePriors(const std::vector<Parameter> &params); /* caller owns them */

private:
    FabberRunData &m_rundata;
    Prior *CreatePrior(Parameter p);
};

/* MISCMATHS::digamma (single precision AS 103) and tools.cc:87-98 gammaln, host copies for ARDPrior */
double digamma_fsl_host(double x);
double gammaln_host(double x);

} // namespace fabber_b200
