/*
 * operators_test.cc - the reference's unit tests for its operator plug-in interface, restated against
 * fabber_core_b200/host/operators.h (same class names and calls):
 *   test/test_convergence.cc:35-263   exact true/false sequences of Test(F) for maxits, pointzeroone, freduce,
 *                                     trialmode (incl. the "one extra iteration" quirk), plus lm
 *   test/test_priors.cc:184-252       DefaultPrior / ImagePrior ApplyToMVN set the right element exactly
 *   plus ARDPrior (priors.cc:150-181), the noise models' hard-coded initial distributions
 *   (noisemodel_white.cc:127-164, noisemodel_ar.cc:379-403) and the registries' names (setup.cc:26-58).
 * Built and run by tests/test_host_operators.py; exit code 0 = all passed. No GPU needed.
 */
#include <cmath>
#include <cstdio>
#include <memory>

#include "../../fabber_core_b200/host/operators.h"

using namespace fabber_b200;

static int failures = 0;
#define CHECK(cond)                                                          \
    do                                                                       \
    {                                                                        \
        if (!(cond))                                                         \
        {                                                                    \
            printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);         \
            failures++;                                                      \
        }                                                                    \
    } while (0)

static void test_maxits()
{
    FabberRunData rd;
    rd.Set("max-iterations", "3");
    std::unique_ptr<ConvergenceDetector> c(ConvergenceDetector::NewFromName("maxits"));
    c->Initialize(rd);
    CHECK(!c->UseF());
    CHECK(!c->Test(1.0));
    CHECK(!c->Test(2.0));
    CHECK(c->Test(3.0)); /* third pass reaches the maximum */
    CHECK(!c->NeedSave() && !c->NeedRevert() && c->LMalpha() == 0.0f);
    c->Reset();
    CHECK(!c->Test(1.0));
    CHECK(c->DeviceCode() == FABBER_CONV_MAXITS);
}
static void test_fchange()
{
    FabberRunData rd;
    rd.Set("max-iterations", "10");
    rd.Set("min-fchange", "0.1");
    std::unique_ptr<ConvergenceDetector> c(ConvergenceDetector::NewFromName("pointzeroone"));
    c->Initialize(rd);
    CHECK(c->UseF());
    CHECK(!c->Test(-100.0));
    CHECK(!c->Test(-50.0));
    CHECK(!c->Test(-49.0));
    CHECK(c->Test(-48.95)); /* |dF| = 0.05 < 0.1 */
    c->Reset();
    for (int i = 0; i < 9; i++)
        CHECK(!c->Test(i * 10.0));
    CHECK(c->Test(1000.0)); /* max iterations */
}
static void test_freduce()
{
    FabberRunData rd;
    rd.Set("max-iterations", "10");
    rd.Set("min-fchange", "0.1");
    std::unique_ptr<ConvergenceDetector> c(ConvergenceDetector::NewFromName("freduce"));
    c->Initialize(rd);
    CHECK(!c->Test(-100.0));
    CHECK(!c->Test(-50.0));
    CHECK(c->Test(-60.0)); /* F reduced: stop and revert */
    CHECK(c->NeedRevert());
    CHECK(!c->NeedSave()); /* freduce never saves: the revert goes back to the pre-loop copies */
}
static void test_trialmode()
{
    FabberRunData rd;
    rd.Set("max-iterations", "3");
    rd.Set("min-fchange", "0.1");
    rd.Set("max-trials", "2");
    std::unique_ptr<ConvergenceDetector> c(ConvergenceDetector::NewFromName("trialmode"));
    c->Initialize(rd);
    /* one extra iteration (test_convergence.cc:211-222): max-iterations 3 stops on the FOURTH pass */
    CHECK(!c->Test(1.0));
    CHECK(c->NeedSave());
    CHECK(!c->Test(2.0));
    CHECK(!c->Test(3.0));
    CHECK(c->Test(4.0));
    c->Reset();
    CHECK(!c->Test(10.0));
    CHECK(!c->Test(5.0)); /* F dropped: trial mode, revert */
    CHECK(c->NeedRevert() && !c->NeedSave());
    CHECK(c->Test(4.0)); /* second trial fails: max-trials reached */
    CHECK(c->NeedRevert());
    c->Reset();
    CHECK(!c->Test(10.0));
    CHECK(!c->Test(5.0));
    CHECK(!c->Test(11.0)); /* recovered: back to normal, saving again */
    CHECK(c->NeedSave() && !c->NeedRevert());
}
static void test_lm()
{
    FabberRunData rd;
    rd.Set("max-iterations", "10");
    rd.Set("max-fchange", "0.01");
    std::unique_ptr<ConvergenceDetector> c(ConvergenceDetector::NewFromName("lm"));
    c->Initialize(rd);
    CHECK(c->NeedSave()); /* always true (convergence.cc:270) */
    CHECK(!c->Test(-100.0));
    CHECK(c->LMalpha() == 0.0f);
    CHECK(!c->Test(-120.0)); /* F dropped: LM mode, alpha = 1e-6, revert */
    CHECK(c->NeedRevert() && c->LMalpha() == (float)1e-6);
    CHECK(!c->Test(-130.0)); /* still worse: alpha * 10 */
    CHECK(c->LMalpha() == (float)1e-5 && c->NeedRevert());
    CHECK(!c->Test(-90.0)); /* better: alpha / 10, accept */
    CHECK(c->LMalpha() == (float)1e-6 && !c->NeedRevert());
    CHECK(!c->Test(-80.0)); /* better at alphastart: leave LM mode */
    CHECK(!c->Test(-70.0));
    CHECK(c->Test(-69.995)); /* converged */
    CHECK(c->NeedSave());
    fabber_cuda_vb_problem prob = fabber_cuda_vb_problem();
    c->Describe(prob);
    CHECK(prob.conv_type == FABBER_CONV_LM && prob.max_iterations == 10 && prob.fchange == 0.01);
    bool threw = false;
    try
    {
        std::unique_ptr<ConvergenceDetector> bad(ConvergenceDetector::NewFromName("nosuch"));
    }
    catch (InvalidOptionValue &)
    {
        threw = true;
    }
    CHECK(threw);
    CHECK(ConvergenceDetector::GetKnown().size() == 5);
}

static Parameter make_param(unsigned idx, const char *name, double mean, double var, char type)
{
    return Parameter(idx, name, DistParams(mean, var), DistParams(mean, var), type, 'I');
}
static void test_priors()
{
    FabberRunData rd;
    /* test_priors.cc:184-199: DefaultPrior sets mean and precision of ITS element only */
    std::vector<Parameter> params;
    params.push_back(make_param(0, "a", 1.5, 4.0, 'N'));
    params.push_back(make_param(1, "b", -2.0, 0.25, 'N'));
    params.push_back(make_param(2, "c", 7.0, 100.0, 'A'));
    PriorFactory factory(rd);
    std::vector<Prior *> priors = factory.CreatePriors(params);
    CHECK(priors.size() == 3 && priors[0]->DeviceCode() == 'N' && priors[2]->DeviceCode() == 'A');
    MVNDist prior(3);
    RunContext ctx;
    ctx.v = 1;
    ctx.nvoxels = 1;
    ctx.fwd_post.push_back(MVNDist(3));
    ctx.fwd_post[0].means[2] = 3.0;
    ctx.fwd_post[0].SetCovariance(2, 2, 0.5);
    CHECK(priors[0]->ApplyToMVN(&prior, ctx) == 0);
    CHECK(priors[1]->ApplyToMVN(&prior, ctx) == 0);
    CHECK(prior.means[0] == 1.5 && prior.means[1] == -2.0);
    CHECK(prior.GetPrecisions(0, 0) == 0.25 && prior.GetPrecisions(1, 1) == 4.0 && prior.GetPrecisions(1, 0) == 0.0);
    /* ARD, first iteration: model default; free-energy term from the posterior's second moment 9.5 */
    ctx.it = 0;
    const double F0 = priors[2]->ApplyToMVN(&prior, ctx);
    CHECK(prior.means[2] == 7.0 && std::fabs(prior.GetCovariance(2, 2) - 100.0) < 1e-12);
    const double b = 2 / 9.5;
    const double want = -1.5 * (std::log(b) + digamma_fsl_host(0.5)) - 0.5 - gammaln_host(0.5) - 0.5 * std::log(b);
    CHECK(F0 == want);
    CHECK(std::fabs(gammaln_host(0.5) - 0.5723649429) < 1e-9 && std::fabs(digamma_fsl_host(0.5) + 1.9635100) < 1e-5);
    ctx.it = 1; /* later iterations: prior variance = m^2 + Sigma */
    priors[2]->ApplyToMVN(&prior, ctx);
    CHECK(std::fabs(prior.GetCovariance(2, 2) - 9.5) < 1e-12 && prior.means[2] == 7.0);
    for (size_t i = 0; i < priors.size(); i++)
        delete priors[i];
    /* test_priors.cc:229-252: ImagePrior takes its mean from the voxel's image value */
    rd.SetExtent(3, 1, 1, nullptr);
    const float img[3] = { 0.5f, 1.5f, 2.5f };
    rd.SetVoxelDataArray("myimage", 1, img);
    Parameter pi = make_param(1, "b", 0.0, 0.25, 'I');
    pi.options["image"] = "myimage";
    ImagePrior ip(pi, rd);
    MVNDist prior2(2);
    ctx.v = 2;
    CHECK(ip.ApplyToMVN(&prior2, ctx) == 0);
    CHECK(prior2.means[1] == 1.5 && prior2.GetPrecisions(1, 1) == 4.0 && prior2.means[0] == 0.0);
    /* spatial priors are device-only: the class validates its options and refuses host arithmetic */
    rd.Set("spatial-dims", "2");
    SpatialPrior sp(make_param(0, "a", 0, 1, 'M'), rd);
    fabber_cuda_vb_problem prob = fabber_cuda_vb_problem();
    sp.Describe(prob);
    CHECK(prob.spatial_dims == 2 && prob.spatial_speed == -1 && prob.spatial_q1 == 10.0 && prob.spatial_q2 == 1.0);
    bool threw = false;
    try
    {
        sp.ApplyToMVN(&prior2, ctx);
    }
    catch (FabberInternalError &)
    {
        threw = true;
    }
    CHECK(threw);
    CHECK(Prior::ExpandPriorTypesString("M+", 4) == "MMMM" && Prior::ExpandPriorTypesString("NI+A", 5) == "NIIIA");
}

static void test_noise_models()
{
    FabberRunData rd;
    rd.Set("noise-pattern", "12");
    std::unique_ptr<NoiseModel> w(NoiseModel::NewFromName("white"));
    w->Initialize(rd);
    CHECK(w->NumParams() == 2);
    std::unique_ptr<NoiseParams> prior(w->NewParams()), post(w->NewParams());
    w->HardcodedInitialDists(*prior, *post);
    CHECK(prior->phis.size() == 2 && prior->phis[1].b == 1e6 && prior->phis[1].c == 1e-6);
    CHECK(post->phis[0].b == 1e-8 && post->phis[0].c == 50);
    fabber_cuda_vb_problem prob = fabber_cuda_vb_problem();
    std::vector<unsigned char> pattern;
    w->Describe(prob, 5, pattern);
    CHECK(prob.noise_type == FABBER_NOISE_WHITE && prob.n_phis == 2 && pattern.size() == 5);
    CHECK(pattern[0] == 0 && pattern[1] == 1 && pattern[4] == 0);
    MVNDist out = post->OutputAsMVN(); /* noisemodel_white.cc:55-68: means b c, variances b^2 c */
    CHECK(out.GetSize() == 2 && out.means[0] == 1e-8 * 50 && out.GetCovariance(1, 1) == 1e-8 * 1e-8 * 50);
    FabberRunData rd2;
    rd2.Set("prior-noise-stddev", "2");
    std::unique_ptr<NoiseModel> w2(NoiseModel::NewFromName("white"));
    w2->Initialize(rd2);
    std::unique_ptr<NoiseParams> p2(w2->NewParams()), q2(w2->NewParams());
    w2->HardcodedInitialDists(*p2, *q2);
    CHECK(p2->phis[0].c == 0.5 && p2->phis[0].b == 1 / (4 * 0.5) && q2->phis[0].b == p2->phis[0].b);

    FabberRunData rd3;
    std::unique_ptr<NoiseModel> ar(NoiseModel::NewFromName("ar"));
    ar->Initialize(rd3);
    CHECK(ar->NumParams() == 1); /* nPhis, although the MVN carries the alphas too (noisemodel_ar.cc:362-365) */
    std::unique_ptr<NoiseParams> pa(ar->NewParams()), qa(ar->NewParams());
    ar->HardcodedInitialDists(*pa, *qa);
    CHECK(pa->alpha.GetSize() == 2 && std::fabs(pa->alpha.GetPrecisions(0, 0) - 1e-4) < 1e-18 && qa->phis[0].c == 1e-6);
    CHECK(qa->OutputAsMVN().GetSize() == 3);
    ar->Describe(prob, 5, pattern);
    CHECK(prob.noise_type == FABBER_NOISE_AR1 && std::fabs(prob.ar_alpha_prior_prec - 1e-4) < 1e-18);
    FabberRunData rd4;
    rd4.Set("mt1", "3");
    bool threw = false;
    try
    {
        std::unique_ptr<NoiseModel> bad(NoiseModel::NewFromName("ar"));
        bad->Initialize(rd4); /* test_inference.cc:564-633: AR + masked time points must throw */
    }
    catch (InvalidOptionValue &)
    {
        threw = true;
    }
    CHECK(threw);
    threw = false;
    try
    {
        w->UpdateNoise(); /* no CPU inference path */
    }
    catch (FabberInternalError &)
    {
        threw = true;
    }
    CHECK(threw);
    CHECK(NoiseModel::GetKnown().size() == 2);
}

/* a model written against the DEPRECATED FwdModel API (fwdmodel.h:256-348): NameParams + HardcodedInitialDists +
 * Evaluate + ardindices; the current API's defaults are built from them (fwdmodel.cc:339-363, fwdmodel.h:152) */
class OldStyleModel : public FwdModel
{
public:
    void Initialize(FabberRunData &) override { ardindices.push_back(2); }
    void NameParams(std::vector<std::string> &names) const override
    {
        names.push_back("offset");
        names.push_back("slope");
    }
    void HardcodedInitialDists(MVNDist &prior, MVNDist &posterior) const override
    {
        prior.means[0] = 3.0;
        prior.SetCovariance(0, 0, 1e6);
        prior.SetCovariance(1, 1, 1e3);
        posterior.means[1] = 0.5;
        posterior.SetCovariance(1, 1, 10.0);
    }
    void Evaluate(const std::vector<double> &p, std::vector<double> &result) const override
    {
        for (size_t t = 0; t < result.size(); t++)
            result[t] = p[0] + p[1] * (double)t;
    }
    void GetDeviceModel(fabber_cuda_model &m) const override
    {
        m.id = FABBER_MODEL_POLY;
        m.n_params = 2;
        m.poly_degree = 1;
    }
};
static void test_deprecated_fwdmodel_api()
{
    FabberRunData rd;
    OldStyleModel model;
    model.Initialize(rd);
    std::vector<Parameter> params;
    model.GetParameters(rd, params);
    CHECK(params.size() == 2 && params[0].name == "offset" && params[1].name == "slope");
    CHECK(model.NumParams() == 2);
    CHECK(params[0].prior.mean() == 3.0 && params[0].prior.var() == 1e6 && params[0].prior_type == 'N');
    CHECK(params[1].prior.var() == 1e3 && params[1].post.mean() == 0.5 && params[1].post.var() == 10.0);
    CHECK(params[1].prior_type == 'A'); /* ardindices */
    std::vector<double> p(2), out;
    p[0] = 1.0;
    p[1] = 2.0;
    model.EvaluateModel(p, out, 4);
    CHECK(out.size() == 4 && out[0] == 1.0 && out[3] == 7.0);
}

/* inference.h:22-166, setup.cc:28-33, test/test_inference.cc:50-55 (CanCreate for every method) */
static void test_inference_techniques()
{
    std::vector<std::string> known = InferenceTechnique::GetKnown();
    CHECK(known.size() == 3 && known[0] == "nlls" && known[1] == "spatialvb" && known[2] == "vb");
    for (size_t i = 0; i < known.size(); i++)
    {
        std::unique_ptr<InferenceTechnique> t(InferenceTechnique::NewFromName(known[i]));
        CHECK(t.get() != nullptr && !t->GetDescription().empty());
        std::vector<OptionSpec> opts;
        t->GetOptions(opts);
        bool has_noise = false, has_lm = false;
        for (size_t k = 0; k < opts.size(); k++)
        {
            has_noise = has_noise || opts[k].name == "noise";
            has_lm = has_lm || opts[k].name == "vb-init"; /* the one NLLS option the reference lists (NUM_OPTIONS = 1) */
        }
        CHECK(has_noise == (known[i] != "nlls") && has_lm == (known[i] == "nlls"));
    }
    CHECK(dynamic_cast<NLLSInferenceTechnique *>(std::unique_ptr<InferenceTechnique>(InferenceTechnique::NewFromName("nlls")).get()));
    bool threw = false;
    try
    {
        InferenceTechnique::NewFromName("mcmc");
    }
    catch (InvalidOptionValue &)
    {
        threw = true;
    }
    CHECK(threw);
}

int main()
{
    test_inference_techniques();
    test_deprecated_fwdmodel_api();
    test_maxits();
    test_fchange();
    test_freduce();
    test_trialmode();
    test_lm();
    test_priors();
    test_noise_models();
    printf("%s (%d failure%s)\n", failures ? "FAILED" : "ok", failures, failures == 1 ? "" : "s");
    return failures ? 1 : 0;
}
