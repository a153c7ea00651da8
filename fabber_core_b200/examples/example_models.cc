/*
 * example_models.cc - host half of the example plug-in library: the FwdModel classes and the three symbols
 * the loader resolves (the reference's examples/exp_models.cc does the same for its CPU models).
 */
#include <cmath>
#include <cstring>

#include "../host/fabber_host.h"
#include "../../include/fabber_model_plugin.h"

extern "C" const void *fabber_example_sine_launchers();
extern "C" const void *fabber_example_exp_launchers();

using namespace fabber_b200;

namespace
{
class SineFwdModel : public FwdModel
{
public:
    std::string ModelVersion() const override { return "example plug-in 1.0"; }
    std::string GetDescription() const override { return "Example plug-in model: a*sin(b*(t-c))+d"; }
    void GetOptions(std::vector<OptionSpec> &opts) const override
    {
        OptionSpec dt = { "dt", OPT_FLOAT, "Time separation between samples", true, "1" };
        opts.push_back(dt);
    }
    void Initialize(FabberRunData &rundata) override { m_dt = rundata.GetDoubleDefault("dt", 1.0); }
    void GetParameterDefaults(std::vector<Parameter> &params) const override
    {
        params.clear();
        params.push_back(Parameter(0, "a", DistParams(1, 1e6), DistParams(1, 1e6)));
        params.push_back(Parameter(1, "b", DistParams(1, 1e6), DistParams(1, 1e6)));
        params.push_back(Parameter(2, "c", DistParams(0, 1e6), DistParams(0, 1e6)));
        params.push_back(Parameter(3, "d", DistParams(0, 1e6), DistParams(0, 1e6)));
    }
    void EvaluateModel(const std::vector<double> &p, std::vector<double> &result, int n_times,
        const std::string &) const override
    {
        result.resize(n_times);
        for (int i = 0; i < n_times; i++)
            result[i] = p[0] * std::sin(p[1] * (i * m_dt - p[2])) + p[3];
    }
    void GetDeviceModel(fabber_cuda_model &m) const override
    {
        m.id = FABBER_MODEL_PLUGIN;
        m.n_params = 4;
        m.plugin_launchers = fabber_example_sine_launchers();
        m.consts[0] = m_dt;
    }

private:
    double m_dt = 1.0;
};

/* examples/fwdmodel_exp.cc with num-exps fixed at 2 (one device struct per parameter count) */
class BiExpFwdModel : public FwdModel
{
public:
    std::string ModelVersion() const override { return "example plug-in 1.0"; }
    std::string GetDescription() const override { return "Example model of a sum of exponentials (plug-in build)"; }
    void GetOptions(std::vector<OptionSpec> &opts) const override
    {
        OptionSpec a = { "dt", OPT_FLOAT, "Time separation between samples", false, "" };
        OptionSpec b = { "num-exps", OPT_INT, "Number of independent decay rates (this build: 2)", true, "2" };
        opts.push_back(a);
        opts.push_back(b);
    }
    void Initialize(FabberRunData &rundata) override
    {
        m_dt = rundata.GetDouble("dt");
        if (rundata.GetIntDefault("num-exps", 2) != 2)
            throw InvalidOptionValue("num-exps", rundata.GetString("num-exps"), "the example plug-in is compiled for 2");
    }
    void GetParameterDefaults(std::vector<Parameter> &params) const override
    {
        params.clear();
        int p = 0;
        for (int i = 0; i < 2; i++)
        {
            params.push_back(Parameter(p++, "amp" + stringify(i + 1), DistParams(1, 1e5), DistParams(1, 1.5), 'N', 'L'));
            params.push_back(Parameter(p++, "r" + stringify(i + 1), DistParams(1, 1e5), DistParams(1, 1.5), 'N', 'L'));
        }
    }
    void EvaluateModel(const std::vector<double> &p, std::vector<double> &result, int n_times,
        const std::string &) const override
    {
        result.assign(n_times, 0.0);
        for (int k = 0; k < 2; k++)
            for (int i = 0; i < n_times; i++)
                result[i] += p[2 * k] * std::exp(-p[2 * k + 1] * (double(i) * m_dt));
    }
    void GetDeviceModel(fabber_cuda_model &m) const override
    {
        m.id = FABBER_MODEL_PLUGIN;
        m.n_params = 4;
        m.plugin_launchers = fabber_example_exp_launchers();
        m.consts[0] = m_dt;
    }

private:
    double m_dt = 1.0;
};

FwdModel *new_sine() { return new SineFwdModel(); }
FwdModel *new_biexp() { return new BiExpFwdModel(); }
} // namespace

extern "C" {
int fabber_b200_plugin_abi(void) { return FABBER_B200_PLUGIN_ABI; }
int get_num_models(void) { return 2; }
const char *get_model_name(int index)
{
    switch (index)
    {
    case 0:
        return "sine";
    case 1:
        return "exp";
    default:
        return nullptr;
    }
}
void *(*get_new_instance_func(const char *name))(void)
{
    typedef void *(*Fn)(void);
    if (strcmp(name, "sine") == 0)
        return (Fn)new_sine;
    if (strcmp(name, "exp") == 0)
        return (Fn)new_biexp;
    return nullptr;
}
}
