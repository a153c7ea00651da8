/*
 * fabber_main.cc - the command line tool: same options, same files in and out as the reference's `fabber`
 * (fabber_main.cc, fabber_core.cc:97-323), with the VB calculation on the GPU.
 *
 *   fabber_b200 --output=out --method=vb --model=poly --degree=2 --noise=white --data=data.nii.gz --mask=mask.nii.gz
 *   fabber_b200 -f options.txt          fabber_b200 --listmodels | --listmethods | --help [--model=..|--method=..]
 *
 * Differences, stated: --loadmodels of the reference's CPU model libraries is refused (models are compiled
 * __device__ hooks); the logfile is written when the run ends rather than line by line.
 */
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "fabber_host.h"

using namespace fabber_b200;

static std::ostream &operator<<(std::ostream &out, const OptionSpec &o)
{
    return out << "--" << o.name << " [" << option_type_name(o.type) << "," << (o.optional ? "NOT REQUIRED" : "REQUIRED")
               << "," << (o.def == "" ? std::string("NO DEFAULT") : "DEFAULT=" + o.def) << "]" << std::endl
               << "        " << o.description << std::endl;
}

static void Version() { std::cout << "Fabber " << fabber_b200_version() << std::endl; }

static void Usage()
{
    Version();
    std::cout << "Usage: fabber [--<option>|--<option>=<value> ...]" << std::endl
              << std::endl
              << "Use -f <file> to read options in option=value form" << std::endl
              << "Use -@ <file> to read options in command line form (DEPRECATED)." << std::endl
              << std::endl
              << "General options " << std::endl
              << std::endl;
    std::vector<OptionSpec> options;
    FabberRunData::GetOptions(options);
    for (size_t i = 0; i < options.size(); i++)
        std::cout << options[i] << std::endl;
}

static void ModelUsage(const std::string &name)
{
    std::unique_ptr<FwdModel> model(FwdModel::NewFromName(name));
    std::cout << name << ": " << model->ModelVersion() << std::endl << std::endl;
    std::cout << model->GetDescription() << std::endl << std::endl;
    std::cout << "Options: " << std::endl << std::endl;
    std::vector<OptionSpec> options;
    model->GetOptions(options);
    for (size_t i = 0; i < options.size(); i++)
        std::cout << options[i];
    std::vector<std::string> outputs;
    model->GetOutputs(outputs);
    if (!outputs.empty())
    {
        std::cout << std::endl << "Additional outputs: " << std::endl << std::endl;
        for (size_t i = 0; i < outputs.size(); i++)
            if (outputs[i] != "")
                std::cout << "  " << outputs[i] << std::endl;
    }
}

static void MethodUsage(const std::string &name)
{
    const std::vector<std::string> known = Vb::GetKnownMethods();
    if (std::find(known.begin(), known.end(), name) == known.end())
        throw InvalidOptionValue("method", name, "Unrecognized inference method");
    std::cout << "Usage information for method: " << name << std::endl << std::endl;
    std::cout << Vb::GetDescription(name) << std::endl << std::endl << "Options: " << std::endl << std::endl;
    std::vector<OptionSpec> options;
    Vb::GetOptions(options, name);
    for (size_t i = 0; i < options.size(); i++)
        std::cout << options[i] << std::endl;
}

/* PercentProgressCheck / SimpleProgressCheck (rundata.cc:102-137) */
static int g_last_percent = -1;
static void percent_progress(int voxel, int n_voxels)
{
    if (n_voxels == 0)
    {
        std::cout << "100%" << std::endl;
        return;
    }
    const int percent = (int)((100LL * voxel) / n_voxels);
    if (percent > g_last_percent)
    {
        std::cout << "\b\b\b";
        g_last_percent = percent;
        if (percent == 0)
            std::cout << " ";
        std::cout << percent << "%" << std::flush;
        if (percent == 100)
            std::cout << std::endl;
    }
}
static void simple_progress(int voxel, int n_voxels)
{
    if (n_voxels == 0)
    {
        std::cout << "100" << std::endl;
        return;
    }
    const int percent = (int)((100LL * voxel) / n_voxels);
    if (percent > g_last_percent)
    {
        g_last_percent = percent;
        std::cout << percent << std::endl << std::flush;
    }
}

static FwdModel *configured_model(FabberRunData &params)
{
    FwdModel *m = FwdModel::NewFromName(params.GetStringDefault("model", ""));
    m->Initialize(params);
    return m;
}

int execute(int argc, char **argv)
{
    FabberRunDataNewimage params(true);
    bool log_started = false, simple_output = false;
    std::string outdir;
    int ret = 1;
    try
    {
        setenv("FSLOUTPUTTYPE", "NIFTI_GZ", 0);
        params.Parse(argc, argv);
        if (!params.GetBool("no-compat-output")) /* rundata.cc:221-232 (applied after parsing here) */
        {
            const char *compat[] = { "save-mean", "save-std", "save-zstat", "save-noise-mean", "save-noise-std",
                "save-free-energy", "save-mvn" };
            for (size_t i = 0; i < sizeof(compat) / sizeof(compat[0]); i++)
                params.SetBool(compat[i]);
        }
        if (params.GetBool("help") || argc == 1)
        {
            const std::string model = params.GetStringDefault("model", ""), method = params.GetStringDefault("method", "");
            if (model != "")
                ModelUsage(model);
            else if (method != "")
                MethodUsage(method);
            else
                Usage();
            return 0;
        }
        if (params.GetBool("version"))
        {
            const std::string model = params.GetStringDefault("model", "");
            if (model != "")
                std::cout << std::unique_ptr<FwdModel>(FwdModel::NewFromName(model))->ModelVersion() << std::endl;
            else
                Version();
            return 0;
        }
        if (params.GetBool("listmodels"))
        {
            const std::vector<std::string> known = FwdModel::GetKnown();
            for (size_t i = 0; i < known.size(); i++)
                std::cout << known[i] << std::endl;
            return 0;
        }
        if (params.GetBool("listmethods"))
        {
            const std::vector<std::string> known = Vb::GetKnownMethods();
            for (size_t i = 0; i < known.size(); i++)
                std::cout << known[i] << std::endl;
            return 0;
        }
        if (params.GetBool("listparams") || params.GetBool("descparams"))
        {
            std::unique_ptr<FwdModel> model(configured_model(params));
            std::vector<Parameter> mp;
            model->GetParameters(params, mp);
            for (size_t i = 0; i < mp.size(); i++)
                std::cout << mp[i].name << std::endl; /* no descriptions / units are attached to these models */
            return 0;
        }
        if (params.GetBool("listoutputs"))
        {
            std::unique_ptr<FwdModel> model(configured_model(params));
            std::vector<std::string> outputs;
            model->GetOutputs(outputs);
            for (size_t i = 0; i < outputs.size(); i++)
                std::cout << outputs[i] << std::endl;
            return 0;
        }
        if (params.HaveKey("evaluate"))
        {
            std::unique_ptr<FwdModel> model(configured_model(params));
            std::vector<double> values, result;
            int rows = 0, cols = 0;
            read_matrix_file(params.GetString("evaluate-params"), values, rows, cols);
            std::vector<double> p(rows);
            for (int i = 0; i < rows; i++)
                p[i] = values[(size_t)i * cols]; /* first column */
            const int nt = params.GetInt("evaluate-nt", 0);
            model->EvaluateModel(p, result, nt, params.GetStringDefault("evaluate", ""));
            for (size_t i = 0; i < result.size(); i++)
                std::cout << result[i] << std::endl;
            return 0;
        }
        params.SetBool("dump-param-names");
        params.SetBool("link-to-latest");
        params.SetExtentFromData();
        simple_output = params.GetBool("simple-output");
        outdir = params.GetOutputDir();
        log_started = true;
        if (!simple_output)
        {
            std::cout << "----------------------" << std::endl;
            std::cout << "Welcome to FABBER " << fabber_b200_version() << std::endl;
            std::cout << "----------------------" << std::endl;
            std::cout << "Logfile started: " << outdir << "/logfile" << std::endl;
            params.Run(percent_progress);
        }
        else
            params.Run(simple_progress);
        params.ReissueWarnings();
        ret = 0;
    }
    catch (const std::exception &e)
    {
        params.ReissueWarnings();
        params.Log() << "Exception caught in fabber:\n  " << e.what() << std::endl;
        std::cerr << "Exception caught in fabber:\n  " << e.what() << std::endl;
    }
    catch (...)
    {
        params.ReissueWarnings();
        params.Log() << "Some other exception caught in fabber!" << std::endl;
        std::cerr << "Some other exception caught in fabber!" << std::endl;
    }
    if (log_started)
    {
        std::ofstream logfile((outdir + "/logfile").c_str());
        logfile << params.LogText();
        if (!simple_output)
            std::cout << std::endl << "Final logfile: " << outdir << "/logfile" << std::endl;
    }
    else
        std::cerr << params.LogText(); /* never got as far as an output directory */
    return ret;
}

int main(int argc, char **argv) { return execute(argc, argv); }
