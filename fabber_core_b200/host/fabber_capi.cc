/*
 * fabber_capi.cc - the reference's public C API (include/fabber_capi.h; fabber_capi.cc:45-623 upstream)
 * over the B200 host classes. Same conventions: 0 on success, <0 on error with the message copied into
 * the caller's err_buf; no exception crosses the ABI; every buffer belongs to the caller.
 */
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "../../include/fabber_capi.h"
#include "fabber_host.h"

using namespace fabber_b200;

static int fabber_err(int code, const char *msg, char *err_buf)
{
    if (!err_buf)
        return code;
    if (!msg)
        msg = "NULL message";
    strncpy(err_buf, msg, FABBER_ERR_MAXC - 1);
    err_buf[FABBER_ERR_MAXC - 1] = '\0';
    return code;
}

static int copy_out(const std::string &s, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (s.size() >= out_bufsize)
        return fabber_err(-1, "Buffer too small", err_buf);
    memcpy(out_buf, s.c_str(), s.size());
    out_buf[s.size()] = '\0';
    return 0;
}

extern "C" {

void *fabber_new(char *err_buf)
{
    try
    {
        /* one process per GPU: FABBER_CUDA_DEVICE picks the device of this process (default: the
         * calling thread's current device) */
        if (const char *dev = getenv("FABBER_CUDA_DEVICE"))
            if (fabber_cuda_set_device(atoi(dev)) != FABBER_CUDA_OK)
            {
                fabber_err(FABBER_ERR_FATAL, fabber_cuda_last_error(), err_buf);
                return NULL;
            }
        return new FabberRunDataArray();
    }
    catch (...)
    {
        fabber_err(FABBER_ERR_FATAL, "Failed to allocate memory for run data", err_buf);
        return NULL;
    }
}

int fabber_load_models(void *fab, const char *libpath, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!libpath)
        return fabber_err(FABBER_ERR_FATAL, "Library path is NULL", err_buf);
    try
    {
        FwdModel::LoadFromDynamicLibrary(libpath);
        return 0;
    }
    catch (std::exception &e)
    {
        return fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
}

int fabber_set_extent(void *fab, unsigned int nx, unsigned int ny, unsigned int nz, const int *mask, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!mask)
        return fabber_err(FABBER_ERR_FATAL, "Mask is NULL", err_buf);
    if ((nx <= 0) || (ny <= 0) || (nz <= 0))
        return fabber_err(FABBER_ERR_FATAL, "Dimensions must be >0", err_buf);
    try
    {
        ((FabberRunDataArray *)fab)->SetExtent(nx, ny, nz, mask);
        return 0;
    }
    catch (std::exception &e)
    {
        return fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
    catch (...)
    {
        return fabber_err(FABBER_ERR_FATAL, "Error setting extent", err_buf);
    }
}

void fabber_destroy(void *fab)
{
    if (fab)
        delete (FabberRunDataArray *)fab;
}

int fabber_set_opt(void *fab, const char *key, const char *value, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!key || !value)
        return fabber_err(FABBER_ERR_FATAL, "Option key or value is NULL", err_buf);
    try
    {
        ((FabberRunDataArray *)fab)->Set(key, value);
        return 0;
    }
    catch (std::exception &e)
    {
        return fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
}

int fabber_set_data(void *fab, const char *name, unsigned int data_size, const float *data, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!data)
        return fabber_err(FABBER_ERR_FATAL, "Data buffer is NULL", err_buf);
    if (!name)
        return fabber_err(FABBER_ERR_FATAL, "Data name is NULL", err_buf);
    if (data_size <= 0)
        return fabber_err(FABBER_ERR_FATAL, "Data size must be >0", err_buf);
    try
    {
        ((FabberRunDataArray *)fab)->SetVoxelDataArray(name, (int)data_size, data);
        return 0;
    }
    catch (std::exception &e)
    {
        return fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
    catch (...)
    {
        return fabber_err(FABBER_ERR_FATAL, "Error setting data", err_buf);
    }
}

int fabber_get_data_size(void *fab, const char *name, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!name)
        return fabber_err(FABBER_ERR_FATAL, "Data name is NULL", err_buf);
    try
    {
        return ((FabberRunDataArray *)fab)->GetVoxelDataSize(name);
    }
    catch (DataNotFound &e)
    {
        return fabber_err(-1, e.what(), err_buf);
    }
    catch (...)
    {
        return fabber_err(FABBER_ERR_FATAL, "Error getting data", err_buf);
    }
}

int fabber_get_data(void *fab, const char *name, float *data_buf, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!name)
        return fabber_err(FABBER_ERR_FATAL, "Data name is NULL", err_buf);
    if (!data_buf)
        return fabber_err(FABBER_ERR_FATAL, "Data name is NULL", err_buf);
    try
    {
        ((FabberRunDataArray *)fab)->GetVoxelDataArray(name, data_buf);
        return 0;
    }
    catch (DataNotFound &e)
    {
        return fabber_err(-1, e.what(), err_buf);
    }
    catch (...)
    {
        return fabber_err(FABBER_ERR_FATAL, "Error getting data", err_buf);
    }
}

int fabber_dorun(void *fab, unsigned int log_bufsize, char *log_buf, char *err_buf, void (*progress_cb)(int, int))
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!log_buf)
        return fabber_err(FABBER_ERR_FATAL, "Log buffer is NULL", err_buf);
    if (!err_buf)
        return fabber_err(FABBER_ERR_FATAL, "Error buffer is NULL", err_buf);
    int ret = 0;
    FabberRunDataArray *rundata = (FabberRunDataArray *)fab;
    rundata->ClearLog();
    try
    {
        rundata->Run(progress_cb);
    }
    catch (const FabberError &e)
    {
        rundata->Log() << e.what() << std::endl;
        ret = fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
    catch (const std::exception &e)
    {
        rundata->Log() << "STL exception caught in fabber:\n  " << e.what() << std::endl;
        ret = fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
    catch (...)
    {
        rundata->Log() << "Some other exception caught in fabber!" << std::endl;
        ret = fabber_err(FABBER_ERR_FATAL, "Unrecognized exception", err_buf);
    }
    if (log_bufsize > 0)
    {
        strncpy(log_buf, rundata->LogText().c_str(), log_bufsize - 1);
        log_buf[log_bufsize - 1] = '\0';
    }
    return ret;
}

int fabber_get_options(
    void *fab, const char *key, const char *value, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!out_buf)
        return fabber_err(FABBER_ERR_FATAL, "Output buffer is NULL", err_buf);
    if (key && !value)
        return fabber_err(FABBER_ERR_FATAL, "Key specified but no value", err_buf);
    try
    {
        std::vector<OptionSpec> options;
        std::string desc;
        if (!key || (strlen(key) == 0))
            FabberRunDataArray::GetOptions(options);
        else if (strcmp(key, "model") == 0)
        {
            std::unique_ptr<FwdModel> model(FwdModel::NewFromName(value));
            desc = model->GetDescription();
            model->GetOptions(options);
        }
        else if (strcmp(key, "method") == 0)
        {
            if (strcmp(value, "vb") != 0 && strcmp(value, "spatialvb") != 0)
                throw InvalidOptionValue("method", value, "Unrecognized inference method");
            desc = Vb::GetDescription();
            Vb::GetOptions(options);
        }
        desc.erase(std::remove(desc.begin(), desc.end(), '\n'), desc.end());
        std::ostringstream out;
        out << desc << std::endl;
        for (size_t i = 0; i < options.size(); i++)
            out << options[i].name << "\t" << options[i].description << "\t" << option_type_name(options[i].type) << "\t"
                << (options[i].optional ? 1 : 0) << "\t" << options[i].def << std::endl;
        return copy_out(out.str(), out_bufsize, out_buf, err_buf);
    }
    catch (std::exception &e)
    {
        return fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
    catch (...)
    {
        return fabber_err(FABBER_ERR_FATAL, "Error in get_options", err_buf);
    }
}

int fabber_get_models(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!out_buf)
        return fabber_err(FABBER_ERR_FATAL, "Output buffer is NULL", err_buf);
    std::ostringstream out;
    std::vector<std::string> known = FwdModel::GetKnown();
    for (size_t i = 0; i < known.size(); i++)
        out << known[i] << std::endl;
    return copy_out(out.str(), out_bufsize, out_buf, err_buf);
}

int fabber_get_methods(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!out_buf)
        return fabber_err(FABBER_ERR_FATAL, "Output buffer is NULL", err_buf);
    std::ostringstream out;
    std::vector<std::string> known = Vb::GetKnownMethods();
    for (size_t i = 0; i < known.size(); i++)
        out << known[i] << std::endl;
    return copy_out(out.str(), out_bufsize, out_buf, err_buf);
}

static int model_params_text(void *fab, bool descs, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!out_buf)
        return fabber_err(FABBER_ERR_FATAL, "Output buffer is NULL", err_buf);
    try
    {
        FabberRunDataArray *rundata = (FabberRunDataArray *)fab;
        std::unique_ptr<FwdModel> model(FwdModel::NewFromName(rundata->GetString("model")));
        model->Initialize(*rundata);
        std::vector<Parameter> params;
        model->GetParameters(*rundata, params);
        std::ostringstream out;
        for (size_t i = 0; i < params.size(); i++)
        {
            out << params[i].name;
            if (descs)
                out << " " << "No description available";
            out << std::endl;
        }
        return copy_out(out.str(), out_bufsize, out_buf, err_buf);
    }
    catch (std::exception &e)
    {
        return fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
    catch (...)
    {
        return fabber_err(FABBER_ERR_FATAL, "Error in get_model_params", err_buf);
    }
}
int fabber_get_model_params(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    return model_params_text(fab, false, out_bufsize, out_buf, err_buf);
}
int fabber_get_model_param_descs(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    return model_params_text(fab, true, out_bufsize, out_buf, err_buf);
}

int fabber_get_model_outputs(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!out_buf)
        return fabber_err(FABBER_ERR_FATAL, "Output buffer is NULL", err_buf);
    try
    {
        FabberRunDataArray *rundata = (FabberRunDataArray *)fab;
        std::unique_ptr<FwdModel> model(FwdModel::NewFromName(rundata->GetString("model")));
        model->Initialize(*rundata);
        std::vector<std::string> outputs;
        model->GetOutputs(outputs);
        std::ostringstream out;
        for (size_t i = 0; i < outputs.size(); i++)
            if (outputs[i] != "")
                out << outputs[i] << std::endl;
        return copy_out(out.str(), out_bufsize, out_buf, err_buf);
    }
    catch (std::exception &e)
    {
        return fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
    catch (...)
    {
        return fabber_err(FABBER_ERR_FATAL, "Error in get_model_outputs", err_buf);
    }
}

int fabber_model_evaluate_output(void *fab, unsigned int n_params, float *params, unsigned int n_ts, float *indata,
    const char *output_name, float *output, char *err_buf)
{
    if (!fab)
        return fabber_err(FABBER_ERR_FATAL, "Rundata is NULL", err_buf);
    if (!params)
        return fabber_err(FABBER_ERR_FATAL, "Params array is NULL", err_buf);
    if (!output)
        return fabber_err(FABBER_ERR_FATAL, "Output array is NULL", err_buf);
    try
    {
        /* fabber_capi.cc:520-623: model-space parameters, output padded with zeros / truncated to n_ts */
        FabberRunDataArray *rundata = (FabberRunDataArray *)fab;
        std::unique_ptr<FwdModel> model(FwdModel::NewFromName(rundata->GetString("model")));
        model->Initialize(*rundata);
        std::vector<Parameter> model_params;
        model->GetParameters(*rundata, model_params);
        if (n_params != model_params.size())
            return fabber_err(FABBER_ERR_FATAL, "Incorrect number of parameters specified", err_buf);
        std::vector<double> p(params, params + n_params), result;
        model->EvaluateModel(p, result, (int)n_ts, output_name ? output_name : "");
        (void)indata;
        for (unsigned int i = 0; i < n_ts; i++)
            output[i] = i < result.size() ? (float)result[i] : 0.0f;
        return 0;
    }
    catch (std::exception &e)
    {
        return fabber_err(FABBER_ERR_FATAL, e.what(), err_buf);
    }
    catch (...)
    {
        return fabber_err(FABBER_ERR_FATAL, "Error in model_evaluate", err_buf);
    }
}

int fabber_model_evaluate(void *fab, unsigned int n_params, float *params, unsigned int n_ts, float *indata,
    float *output, char *err_buf)
{
    return fabber_model_evaluate_output(fab, n_params, params, n_ts, indata, "", output, err_buf);
}

} /* extern "C" */
