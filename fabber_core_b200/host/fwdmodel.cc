/*
 * fwdmodel.cc - host side of the forward-model plugin API: parameters, transforms, the model registry
 * and the three models that have a compiled __device__ Evaluate hook.
 *
 * Reference: fwdmodel.cc:210-382 (GetParameters, EvaluateFabber), transforms.h:114-242,
 * priors.cc:35-91 (ExpandPriorTypesString), fwdmodel_linear.cc:53-96, fwdmodel_poly.cc:40-80,
 * examples/fwdmodel_exp.cc:43-91, tools.cc:27-40 (matrix files).
 * The host EvaluateModel implementations serve fabber_model_evaluate (fabber_capi.h:279); inference
 * itself never calls them - it runs the device hooks (csrc/vb_models.cuh).
 */
#include <algorithm>
#include <cmath>
#include <fstream>

#include <dlfcn.h>

#include <mutex>

#include "fabber_host.h"

#include "operators.h"

namespace fabber_b200
{
double transform_to_model(char code, double v)
{
    switch (code)
    {
    case 'L':
        return std::exp(v);
    case 'S':
        return v < 10 ? std::log(1 + std::exp(v)) : v;
    case 'F':
        return 1 / (1 + std::exp(v));
    case 'A':
        return std::fabs(v);
    default:
        return v;
    }
}
double transform_to_fabber(char code, double v)
{
    switch (code)
    {
    case 'L':
        return std::log(v);
    case 'S':
        return v < 10 ? std::log(std::exp(v) - 1) : v;
    case 'F':
        return std::log(1 / v - 1);
    default:
        return v;
    }
}
double transform_to_model_var(char code, double v)
{
    switch (code)
    {
    case 'L':
        return std::exp(v);
    case 'I':
    case 'F':
        return v;
    default: /* transforms.cc:17-20 */
        return std::pow(transform_to_model(code, std::sqrt(v)) - transform_to_model(code, 0), 2);
    }
}
double transform_to_fabber_var(char code, double v)
{
    switch (code)
    {
    case 'L':
        return std::log(v);
    case 'I':
    case 'F':
        return v;
    default: /* transforms.cc:22-25 */
        return std::pow(transform_to_fabber(code, transform_to_model(code, 0) + std::sqrt(v)), 2);
    }
}

/* priors.cc:35-91 */
std::string ExpandPriorTypesString(std::string priors_str, unsigned num_params)
{
    unsigned n_str_params = 0;
    char repeat_type = '-';
    bool plus_found = false;
    for (size_t i = 0; i < priors_str.size(); i++)
    {
        if (priors_str[i] != '+')
        {
            if (!plus_found)
                repeat_type = priors_str[i];
            n_str_params++;
        }
        else if (plus_found)
            throw InvalidOptionValue("param-spatial-priors", priors_str, "Only one + character allowed");
        else
            plus_found = true;
    }
    if (n_str_params > num_params)
        throw InvalidOptionValue("param-spatial-priors", priors_str, "Too many parameters");
    else if (n_str_params < num_params)
    {
        int deficit = num_params - n_str_params;
        size_t plus_pos = priors_str.find("+");
        if (plus_pos != std::string::npos)
            priors_str.insert(plus_pos, deficit - 1, '+');
        else
            priors_str.insert(priors_str.end(), deficit, '-');
    }
    else
        priors_str.erase(std::remove(priors_str.begin(), priors_str.end(), '+'), priors_str.end());
    std::replace(priors_str.begin(), priors_str.end(), '+', repeat_type);
    return priors_str;
}

/* ---- registry (setup.cc:44-47; "exp" is the reference's example model library) ---------------------- */
namespace
{
FwdModel *new_linear() { return new LinearFwdModel(); }
FwdModel *new_poly() { return new PolynomialFwdModel(); }
FwdModel *new_exp() { return new ExpFwdModel(); }
std::mutex g_registry_mu;
std::map<std::string, FwdModel::NewInstanceFptr> &registry()
{
    static std::map<std::string, FwdModel::NewInstanceFptr> r;
    if (r.empty())
    {
        r["linear"] = new_linear;
        r["poly"] = new_poly;
        r["exp"] = new_exp;
    }
    return r;
}
} // namespace

FwdModel *FwdModel::NewFromName(const std::string &name)
{
    NewInstanceFptr make = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_registry_mu);
        std::map<std::string, NewInstanceFptr>::const_iterator it = registry().find(name);
        if (it != registry().end())
            make = it->second;
    }
    if (!make)
    {
        std::string known;
        const std::vector<std::string> k = GetKnown();
        for (size_t i = 0; i < k.size(); i++)
            known += (i ? ", " : "") + k[i];
        throw InvalidOptionValue("model", name, "Unrecognized forward model (models with a device hook: " + known + ")");
    }
    return make();
}
std::vector<std::string> FwdModel::GetKnown()
{
    std::lock_guard<std::mutex> lock(g_registry_mu);
    std::vector<std::string> k;
    for (std::map<std::string, NewInstanceFptr>::const_iterator it = registry().begin(); it != registry().end(); ++it)
        k.push_back(it->first);
    return k;
}

/* fwdmodel.cc:63-129: a plug-in library exports get_num_models / get_model_name / get_new_instance_func,
 * exactly the reference's three symbols; what it returns are FwdModel objects of THIS library whose
 * GetDeviceModel() names the plug-in's own kernels (include/fabber_model_plugin.h). */
void FwdModel::LoadFromDynamicLibrary(const std::string &filename, std::ostream *log)
{
    typedef int (*GetNumModelsFptr)(void);
    typedef const char *(*GetModelNameFptr)(int);
    typedef NewInstanceFptr (*GetNewInstanceFptrFptr)(const char *);
    if (log)
        *log << "Loading dynamic models from " << filename << std::endl;
    void *lib = dlopen(filename.c_str(), RTLD_NOW | RTLD_GLOBAL);
    if (!lib)
        throw InvalidOptionValue("loadmodels", filename, std::string("Failed to open library ") + dlerror());
    GetNumModelsFptr get_num_models = (GetNumModelsFptr)dlsym(lib, "get_num_models");
    if (!get_num_models)
        throw InvalidOptionValue("loadmodels", filename, "Failed to resolve symbol 'get_num_models'");
    GetModelNameFptr get_model_name = (GetModelNameFptr)dlsym(lib, "get_model_name");
    if (!get_model_name)
        throw InvalidOptionValue("loadmodels", filename, "Failed to resolve symbol 'get_model_name'");
    GetNewInstanceFptrFptr get_new_instance = (GetNewInstanceFptrFptr)dlsym(lib, "get_new_instance_func");
    if (!get_new_instance)
        throw InvalidOptionValue("loadmodels", filename, "Failed to resolve symbol 'get_new_instance_func'");
    /* a plug-in built against another layout of the host classes or of fabber_cuda.h must not be used */
    typedef int (*AbiFptr)(void);
    AbiFptr abi = (AbiFptr)dlsym(lib, "fabber_b200_plugin_abi");
    if (!abi || abi() != FABBER_B200_PLUGIN_ABI)
        throw InvalidOptionValue("loadmodels", filename,
            "not a fabber_core_b200 model plug-in of this ABI version (the reference's CPU model libraries have no "
            "__device__ Evaluate hook: rebuild the model with include/fabber_model_plugin.h)");
    const int n = get_num_models();
    if (log)
        *log << "Loading " << n << " models" << std::endl;
    for (int i = 0; i < n; i++)
    {
        const char *name = get_model_name(i);
        if (!name)
            throw InvalidOptionValue("loadmodels", filename,
                "Dynamic library failed to return model name for index " + stringify(i));
        if (log)
            *log << "Loading model " << name << std::endl;
        NewInstanceFptr make = get_new_instance(name);
        if (!make)
            throw InvalidOptionValue("loadmodels", filename,
                std::string("Dynamic library failed to return new instance function for model") + name);
        std::lock_guard<std::mutex> lock(g_registry_mu);
        registry()[name] = make;
    }
}

/* fwdmodel.cc:210-282 */
void FwdModel::GetParameters(FabberRunData &rundata, std::vector<Parameter> &params)
{
    params.clear();
    GetParameterDefaults(params);
    m_params.clear();
    for (std::vector<Parameter>::iterator p = params.begin(); p < params.end(); ++p)
    {
        std::string types
            = ExpandPriorTypesString(rundata.GetStringDefault("param-spatial-priors", ""), params.size());
        if (types[p->idx] != '-')
            p->prior_type = types[p->idx];
        p->options["image"] = "image-prior" + stringify(p->idx + 1);
        for (int psp_idx = 1;; psp_idx++)
        {
            std::string name = rundata.GetStringDefault("PSP_byname" + stringify(psp_idx), "stop!");
            if (name == "stop!")
                break;
            if (name == p->name)
            {
                std::string s = stringify(psp_idx);
                std::string tcode = rundata.GetStringDefault("PSP_byname" + s + "_transform", "");
                if (tcode != "")
                {
                    if (tcode.size() != 1 || std::string("ILSFA").find(tcode[0]) == std::string::npos)
                        throw InvalidOptionValue("PSP_byname<n>_transform", tcode, "Supported transforms: I, L, S, F");
                    p->transform = tcode[0];
                }
                std::string ptype = rundata.GetStringDefault("PSP_byname" + s + "_type", std::string(1, p->prior_type));
                if (ptype.size() != 1)
                    throw InvalidOptionValue("PSP_byname<n>_type", ptype, "Must be a single character");
                if (ptype[0] != '-')
                    p->prior_type = ptype[0];
                double mean = rundata.GetDoubleDefault("PSP_byname" + s + "_mean", p->prior.mean());
                double prec = rundata.GetDoubleDefault("PSP_byname" + s + "_prec", p->prior.prec());
                p->prior = DistParams(mean, 1 / prec);
                p->options["image"] = "PSP_byname" + s + "_image";
            }
        }
        if (p->prior.prec() > 1e12)
        {
            rundata.Log() << "WARNING: Specified precision " << p->prior.prec()
                          << " is very high - this can trigger numerical instability. Using 1e12 instead" << std::endl;
            p->prior = DistParams(p->prior.mean(), 1e-12);
        }
        /* prior into Fabber space (fwdmodel.cc:277) */
        p->prior = DistParams(transform_to_fabber(p->transform, p->prior.mean()),
            transform_to_fabber_var(p->transform, p->prior.var()));
        m_params.push_back(*p);
    }
}

/* fwdmodel.cc:365-382 */
void FwdModel::EvaluateFabber(
    const std::vector<double> &theta, std::vector<double> &result, int n_times, const std::string &key) const
{
    std::vector<double> tp(theta.size());
    for (size_t i = 0; i < theta.size() && i < m_params.size(); i++)
        tp[i] = transform_to_model(m_params[i].transform, theta[i]);
    EvaluateModel(tp, result, n_times, key);
}

/* ---- matrix files (tools.cc:27-40 -> MISCMATHS::read_vest / read_ascii_matrix) ------------------------ */
void read_matrix_file(const std::string &filename, std::vector<double> &values, int &rows, int &cols)
{
    std::ifstream in(filename.c_str());
    if (!in)
        throw FabberRunDataError("Could not read matrix file: " + filename);
    std::vector<std::string> lines;
    std::string line;
    bool vest = false;
    size_t matrix_at = 0;
    while (std::getline(in, line))
    {
        if (line.compare(0, 7, "/Matrix") == 0)
        {
            vest = true;
            matrix_at = lines.size() + 1;
        }
        lines.push_back(line);
    }
    values.clear();
    rows = cols = 0;
    for (size_t i = vest ? matrix_at : 0; i < lines.size(); i++)
    {
        const std::string &l = lines[i];
        size_t first = l.find_first_not_of(" \t\r");
        if (first == std::string::npos)
            continue;
        if (!vest && (l[first] == '#' || l[first] == '/' || l[first] == '%'))
            continue;
        std::istringstream s(l);
        double x;
        int n = 0;
        while (s >> x)
        {
            values.push_back(x);
            n++;
        }
        if (n == 0)
            continue;
        if (cols == 0)
            cols = n;
        else if (n != cols)
            throw FabberRunDataError("Matrix file has rows of different lengths: " + filename);
        rows++;
    }
    if (rows == 0)
        throw FabberRunDataError("Matrix file is empty: " + filename);
}

/* ---- linear (fwdmodel_linear.cc:25-96) --------------------------------------------------------------- */
std::string LinearFwdModel::GetDescription() const
{
    return "Model in which output is a linear combination of input parameters";
}
void LinearFwdModel::GetOptions(std::vector<OptionSpec> &opts) const
{
    OptionSpec o = { "basis", OPT_MATRIX, "Design matrix", false, "" };
    opts.push_back(o);
}
void LinearFwdModel::Initialize(FabberRunData &args)
{
    std::string designFile = args.GetString("basis");
    args.Log() << "LinearFwdModel::Reading design file: " << designFile << std::endl;
    read_matrix_file(designFile, m_design, m_ntimes, m_nbasis);
    if (args.GetBool("add-ones-regressor"))
    {
        std::vector<double> d2((size_t)m_ntimes * (m_nbasis + 1));
        for (int t = 0; t < m_ntimes; t++)
        {
            for (int j = 0; j < m_nbasis; j++)
                d2[(size_t)t * (m_nbasis + 1) + j] = m_design[(size_t)t * m_nbasis + j];
            d2[(size_t)t * (m_nbasis + 1) + m_nbasis] = 1.0;
        }
        m_design.swap(d2);
        m_nbasis++;
    }
}
void LinearFwdModel::GetParameterDefaults(std::vector<Parameter> &params) const
{
    for (int i = 0; i < m_nbasis; i++)
        params.push_back(Parameter(i, "Parameter_" + stringify(i + 1), DistParams(0, 1e12), DistParams(0, 1e12)));
}
void LinearFwdModel::EvaluateModel(
    const std::vector<double> &p, std::vector<double> &result, int, const std::string &) const
{
    if ((int)p.size() != m_nbasis)
        throw InvalidOptionValue("num params", stringify(p.size()), "Incorrect number of parameters");
    result.assign(m_ntimes, 0.0);
    for (int t = 0; t < m_ntimes; t++)
    {
        double s = 0;
        for (int j = 0; j < m_nbasis; j++)
            s += m_design[(size_t)t * m_nbasis + j] * (p[j] - 0.0);
        result[t] = s + 0.0;
    }
}
void LinearFwdModel::GetDeviceModel(fabber_cuda_model &m) const
{
    m.id = FABBER_MODEL_LINEAR;
    m.n_params = m_nbasis;
    m.design = m_design.data();
}

/* ---- poly (fwdmodel_poly.cc) ------------------------------------------------------------------------- */
std::string PolynomialFwdModel::GetDescription() const
{
    return "Model which fits data to a simple polynomial function: c0 + c1x + c2x^2 ... etc";
}
void PolynomialFwdModel::GetOptions(std::vector<OptionSpec> &opts) const
{
    OptionSpec o = { "degree", OPT_INT, "Maximum power in the polynomial function", false, "" };
    opts.push_back(o);
}
void PolynomialFwdModel::Initialize(FabberRunData &args) { m_degree = args.GetInt("degree", 0); }
void PolynomialFwdModel::GetParameterDefaults(std::vector<Parameter> &params) const
{
    for (int i = 0; i < m_degree + 1; i++)
        params.push_back(Parameter(i, "c" + stringify(i), DistParams(0, 1e12), DistParams(0, 1e12)));
}
void PolynomialFwdModel::EvaluateModel(
    const std::vector<double> &p, std::vector<double> &result, int n_times, const std::string &) const
{
    if ((int)p.size() != m_degree + 1)
        throw InvalidOptionValue("num params", stringify(p.size()), "Incorrect number of parameters");
    result.assign(n_times, 0.0);
    for (int i = 1; i <= n_times; i++)
    {
        double res = 0;
        unsigned int pw = 1; /* the reference's `int` accumulator wraps the same way */
        for (int n = 0; n <= m_degree; n++)
        {
            res += p[n] * (double)(int)pw;
            pw *= (unsigned int)i;
        }
        result[i - 1] = res;
    }
}
void PolynomialFwdModel::GetDeviceModel(fabber_cuda_model &m) const
{
    m.id = FABBER_MODEL_POLY;
    m.n_params = m_degree + 1;
    m.poly_degree = m_degree;
}

/* ---- exp (examples/fwdmodel_exp.cc) ------------------------------------------------------------------ */
std::string ExpFwdModel::GetDescription() const { return "Example model of a sum of exponentials"; }
void ExpFwdModel::GetOptions(std::vector<OptionSpec> &opts) const
{
    OptionSpec a = { "dt", OPT_FLOAT, "Time separation between samples", false, "" };
    OptionSpec b = { "num-exps", OPT_INT, "Number of independent decay rates", true, "1" };
    opts.push_back(a);
    opts.push_back(b);
}
void ExpFwdModel::Initialize(FabberRunData &rundata)
{
    m_dt = rundata.GetDouble("dt");
    m_num = rundata.GetIntDefault("num-exps", 1);
    if (m_num < 1 || 2 * m_num > FABBER_CUDA_MAX_PARAMS)
        throw InvalidOptionValue("num-exps", stringify(m_num), "Must be between 1 and 3");
}
void ExpFwdModel::GetParameterDefaults(std::vector<Parameter> &params) const
{
    params.clear();
    int p = 0;
    for (int i = 0; i < m_num; i++)
    {
        params.push_back(Parameter(p++, "amp" + stringify(i + 1), DistParams(1, 1e5), DistParams(1, 1.5), 'N', 'L'));
        params.push_back(Parameter(p++, "r" + stringify(i + 1), DistParams(1, 1e5), DistParams(1, 1.5), 'N', 'L'));
    }
}
void ExpFwdModel::EvaluateModel(
    const std::vector<double> &p, std::vector<double> &result, int n_times, const std::string &) const
{
    if ((int)p.size() != 2 * m_num)
        throw InvalidOptionValue("num params", stringify(p.size()), "Incorrect number of parameters");
    result.assign(n_times, 0.0);
    for (int k = 0; k < m_num; k++)
    {
        const double amp = p[2 * k], r = p[2 * k + 1];
        for (int i = 0; i < n_times; i++)
        {
            const double t = double(i) * m_dt;
            result[i] += amp * std::exp(-r * t);
        }
    }
}
void ExpFwdModel::GetDeviceModel(fabber_cuda_model &m) const
{
    m.id = FABBER_MODEL_EXP;
    m.n_params = 2 * m_num;
    m.exp_num = m_num;
    m.exp_dt = m_dt;
}

/* fwdmodel.cc:339-363: a model written against the deprecated API names its parameters and fills two MVNs */
void FwdModel::GetParameterDefaults(std::vector<Parameter> &params) const
{
    params.clear();
    std::vector<std::string> names;
    NameParams(names);
    MVNDist priors((int)names.size()), posts((int)names.size());
    HardcodedInitialDists(priors, posts);
    for (size_t i = 0; i < names.size(); i++)
    {
        Parameter p((unsigned)i, names[i], DistParams(priors.means[i], priors.GetCovariance((int)i, (int)i)),
            DistParams(posts.means[i], posts.GetCovariance((int)i, (int)i)), 'N', 'I');
        if (std::find(ardindices.begin(), ardindices.end(), (int)i + 1) != ardindices.end())
            p.prior_type = 'A';
        params.push_back(p);
    }
}
/* fwdmodel.h:149-153 */
void FwdModel::EvaluateModel(const std::vector<double> &params, std::vector<double> &result, int n_times,
    const std::string &key) const
{
    if (key != "")
        throw FabberInternalError("This model does not provide the output '" + key + "'");
    result.assign(n_times, 0.0);
    Evaluate(params, result);
}

} // namespace fabber_b200
