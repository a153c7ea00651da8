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

namespace
{
int fabber_err(int code, const char *msg, char *err_buf)
{
    if (!err_buf)
        return code;
    if (!msg)
        msg = "NULL message";
    strncpy(err_buf, msg, FABBER_ERR_MAXC - 1);
    err_buf[FABBER_ERR_MAXC - 1] = '\0';
    return code;
}

/* argument checks: the first missing one names itself in err_buf (same wording as upstream, which callers match on) */
struct Need
{
    bool present;
    const char *complaint;
};
int first_missing(char *err_buf, std::initializer_list<Need> needs)
{
    for (const Need &n : needs)
        if (!n.present)
            return fabber_err(FABBER_ERR_FATAL, n.complaint, err_buf);
    return 0;
}

/* No exception crosses the ABI. DATA_KEY: an unknown data key is the caller's mistake, not a failure, and comes
 * back as -1 (fabber_capi.h:24-26); anything else is reported with the fixed text, as upstream does there. */
enum GuardMode
{
    REPORT_WHAT,
    DATA_KEY
};
template <class Body> int guarded(char *err_buf, GuardMode mode, const char *fallback, Body body)
{
    try
    {
        return body();
    }
    catch (DataNotFound &e)
    {
        return fabber_err(mode == DATA_KEY ? -1 : FABBER_ERR_FATAL, e.what(), err_buf);
    }
    catch (std::exception &e)
    {
        return fabber_err(FABBER_ERR_FATAL, mode == DATA_KEY ? fallback : e.what(), err_buf);
    }
    catch (...)
    {
        return fabber_err(FABBER_ERR_FATAL, fallback, err_buf);
    }
}

FabberRunDataArray &rundata_of(void *fab) { return *static_cast<FabberRunDataArray *>(fab); }

/* text results: one item per line into the caller's buffer, -1 if it does not fit (fabber_capi.cc:330-333) */
int text_out(const std::string &text, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (text.size() >= out_bufsize)
        return fabber_err(-1, "Buffer too small", err_buf);
    memcpy(out_buf, text.c_str(), text.size() + 1);
    return 0;
}
int lines_out(const std::vector<std::string> &items, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    std::string text;
    for (const std::string &item : items)
        if (!item.empty())
            text += item + "\n";
    return text_out(text, out_bufsize, out_buf, err_buf);
}

/* the model the options currently name, initialised from them - what every model query starts with */
std::unique_ptr<FwdModel> configured_model(FabberRunDataArray &rundata)
{
    std::unique_ptr<FwdModel> model(FwdModel::NewFromName(rundata.GetString("model")));
    model->Initialize(rundata);
    return model;
}
} // namespace

extern "C" {

void *fabber_new(char *err_buf)
{
    void *handle = NULL;
    guarded(err_buf, REPORT_WHAT, "Failed to allocate memory for run data", [&]() {
        /* one process per GPU: FABBER_CUDA_DEVICE picks the device of this process (default: the
         * calling thread's current device) */
        if (const char *dev = getenv("FABBER_CUDA_DEVICE"))
            if (fabber_cuda_set_device(atoi(dev)) != FABBER_CUDA_OK)
                return fabber_err(FABBER_ERR_FATAL, fabber_cuda_last_error(), err_buf);
        handle = new FabberRunDataArray();
        return 0;
    });
    return handle;
}

void fabber_destroy(void *fab) { delete static_cast<FabberRunDataArray *>(fab); }

int fabber_load_models(void *fab, const char *libpath, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { libpath != NULL, "Library path is NULL" } }))
        return rc;
    return guarded(err_buf, REPORT_WHAT, "Error loading models", [&]() {
        FwdModel::LoadFromDynamicLibrary(libpath);
        return 0;
    });
}

int fabber_set_extent(void *fab, unsigned int nx, unsigned int ny, unsigned int nz, const int *mask, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { mask != NULL, "Mask is NULL" },
                                            { nx > 0 && ny > 0 && nz > 0, "Dimensions must be >0" } }))
        return rc;
    return guarded(err_buf, REPORT_WHAT, "Error setting extent", [&]() {
        rundata_of(fab).SetExtent(nx, ny, nz, mask);
        return 0;
    });
}

int fabber_set_opt(void *fab, const char *key, const char *value, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { key && value, "Option key or value is NULL" } }))
        return rc;
    return guarded(err_buf, REPORT_WHAT, "Error setting option", [&]() {
        rundata_of(fab).Set(key, value);
        return 0;
    });
}

int fabber_set_data(void *fab, const char *name, unsigned int data_size, const float *data, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { data != NULL, "Data buffer is NULL" },
                                            { name != NULL, "Data name is NULL" }, { data_size > 0, "Data size must be >0" } }))
        return rc;
    return guarded(err_buf, REPORT_WHAT, "Error setting data", [&]() {
        rundata_of(fab).SetVoxelDataArray(name, (int)data_size, data);
        return 0;
    });
}

int fabber_get_data_size(void *fab, const char *name, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { name != NULL, "Data name is NULL" } }))
        return rc;
    return guarded(err_buf, DATA_KEY, "Error getting data", [&]() { return rundata_of(fab).GetVoxelDataSize(name); });
}

int fabber_get_data(void *fab, const char *name, float *data_buf, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { name != NULL, "Data name is NULL" },
                                            { data_buf != NULL, "Data name is NULL" } })) /* sic: upstream's wording, fabber_capi.cc:188-189 */
        return rc;
    return guarded(err_buf, DATA_KEY, "Error getting data", [&]() {
        rundata_of(fab).GetVoxelDataArray(name, data_buf);
        return 0;
    });
}

int fabber_dorun(void *fab, unsigned int log_bufsize, char *log_buf, char *err_buf, void (*progress_cb)(int, int))
{
    /* the one entry point where err_buf is mandatory (fabber_capi.cc:218-219) */
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { log_buf != NULL, "Log buffer is NULL" },
                                            { err_buf != NULL, "Error buffer is NULL" } }))
        return rc;
    int ret = 0;
    FabberRunDataArray *rundata = &rundata_of(fab);
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
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { out_buf != NULL, "Output buffer is NULL" },
                                            { !key || value, "Key specified but no value" } }))
        return rc;
    return guarded(err_buf, REPORT_WHAT, "Error in get_options", [&]() {
        /* first line: description (newlines removed); then name <TAB> description <TAB> type <TAB> optional <TAB> default */
        std::vector<OptionSpec> options;
        std::string desc;
        const std::string what = key ? key : "";
        if (what.empty())
            FabberRunDataArray::GetOptions(options);
        else if (what == "model")
        {
            std::unique_ptr<FwdModel> model(FwdModel::NewFromName(value));
            desc = model->GetDescription();
            model->GetOptions(options);
        }
        else if (what == "method")
        {
            const std::vector<std::string> methods = Vb::GetKnownMethods();
            if (std::find(methods.begin(), methods.end(), value) == methods.end())
                throw InvalidOptionValue("method", value, "Unrecognized inference method");
            desc = Vb::GetDescription(value);
            Vb::GetOptions(options, value);
        }
        desc.erase(std::remove(desc.begin(), desc.end(), '\n'), desc.end());
        std::ostringstream out;
        out << desc << std::endl;
        for (const OptionSpec &o : options)
            out << o.name << "\t" << o.description << "\t" << option_type_name(o.type) << "\t" << (o.optional ? 1 : 0)
                << "\t" << o.def << std::endl;
        return text_out(out.str(), out_bufsize, out_buf, err_buf);
    });
}

int fabber_get_models(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { out_buf != NULL, "Output buffer is NULL" } }))
        return rc;
    return guarded(err_buf, REPORT_WHAT, "Error in get_models",
        [&]() { return lines_out(FwdModel::GetKnown(), out_bufsize, out_buf, err_buf); });
}

int fabber_get_methods(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { out_buf != NULL, "Output buffer is NULL" } }))
        return rc;
    return guarded(err_buf, REPORT_WHAT, "Error in get_methods",
        [&]() { return lines_out(Vb::GetKnownMethods(), out_bufsize, out_buf, err_buf); });
}

static int model_params_text(void *fab, bool descs, unsigned int out_bufsize, char *out_buf, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { out_buf != NULL, "Output buffer is NULL" } }))
        return rc;
    return guarded(err_buf, REPORT_WHAT, "Error in get_model_params", [&]() {
        FabberRunDataArray &rundata = rundata_of(fab);
        std::vector<Parameter> params;
        configured_model(rundata)->GetParameters(rundata, params);
        std::vector<std::string> lines;
        for (const Parameter &p : params)
            lines.push_back(descs ? p.name + " No description available" : p.name);
        return lines_out(lines, out_bufsize, out_buf, err_buf);
    });
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
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { out_buf != NULL, "Output buffer is NULL" } }))
        return rc;
    return guarded(err_buf, REPORT_WHAT, "Error in get_model_outputs", [&]() {
        std::vector<std::string> outputs;
        configured_model(rundata_of(fab))->GetOutputs(outputs);
        return lines_out(outputs, out_bufsize, out_buf, err_buf);
    });
}

int fabber_model_evaluate_output(void *fab, unsigned int n_params, float *params, unsigned int n_ts, float *indata,
    const char *output_name, float *output, char *err_buf)
{
    if (int rc = first_missing(err_buf, { { fab != NULL, "Rundata is NULL" }, { params != NULL, "Params array is NULL" },
                                            { output != NULL, "Output array is NULL" } }))
        return rc;
    (void)indata; /* the built-in models do not look at the voxel's data */
    return guarded(err_buf, REPORT_WHAT, "Error in model_evaluate", [&]() {
        /* fabber_capi.cc:520-623: model-space parameters in; the series comes back padded with zeros or cut to n_ts */
        FabberRunDataArray &rundata = rundata_of(fab);
        std::unique_ptr<FwdModel> model = configured_model(rundata);
        std::vector<Parameter> declared;
        model->GetParameters(rundata, declared);
        if (n_params != declared.size())
            return fabber_err(FABBER_ERR_FATAL, "Incorrect number of parameters specified", err_buf);
        std::vector<double> p(params, params + n_params), series;
        model->EvaluateModel(p, series, (int)n_ts, output_name ? output_name : "");
        for (unsigned int i = 0; i < n_ts; i++)
            output[i] = i < series.size() ? (float)series[i] : 0.0f;
        return 0;
    });
}

int fabber_model_evaluate(void *fab, unsigned int n_params, float *params, unsigned int n_ts, float *indata,
    float *output, char *err_buf)
{
    return fabber_model_evaluate_output(fab, n_params, params, n_ts, indata, "", output, err_buf);
}

} /* extern "C" */
