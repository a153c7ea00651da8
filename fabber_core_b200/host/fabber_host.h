/*
 * fabber_host.h - C++ host side above the device C ABI (include/fabber_cuda.h).
 *
 * Mirrors the reference's plugin / operator interface for the VB path - same class names, option names,
 * argument meaning and error behaviour - so that code written against fabber_core reads the same here:
 *   FabberRunData / FabberRunDataArray   rundata.h:215-672, rundata_array.h  (options map, voxel data
 *                                        T x Nvox, extent + mask, Run())
 *   Parameter / DistParams / Transform   fwdmodel.h:24-57, transforms.h
 *   FwdModel                             fwdmodel.h:59-371 (Initialize, GetParameterDefaults,
 *                                        EvaluateModel, InitVoxelPosterior, GetOutputs, GetOptions ...)
 *                                        + one addition: GetDeviceModel(), the __device__ Evaluate hook id
 *   LinearFwdModel / PolynomialFwdModel / ExpFwdModel   fwdmodel_linear.*, fwdmodel_poly.*, examples/fwdmodel_exp.*
 *   Vb                                   inference_vb.h (Initialize / DoCalculations / SaveResults); its
 *                                        DoCalculations marshals SoA buffers and calls fabber_cuda_vb_*.
 * No NEWMAT: matrices are plain row-major arrays. No CPU inference path exists in here.
 */
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/fabber_cuda.h"

namespace fabber_b200
{
/* ---- exceptions (rundata.h:676-757) ----------------------------------------------------------- */
struct FabberError : std::runtime_error
{
    explicit FabberError(const std::string &m)
        : std::runtime_error(m)
    {
    }
};
struct FabberInternalError : FabberError
{
    explicit FabberInternalError(const std::string &m)
        : FabberError(m)
    {
    }
};
struct FabberRunDataError : FabberError
{
    explicit FabberRunDataError(const std::string &m)
        : FabberError(m)
    {
    }
};
struct InvalidOptionValue : FabberRunDataError
{
    InvalidOptionValue(const std::string &key, const std::string &value, const std::string &reason)
        : FabberRunDataError("Invalid value given for option: " + key + "=" + value + " (" + reason + ")") /* rundata.h:730 */
    {
    }
};
struct MandatoryOptionMissing : FabberRunDataError
{
    explicit MandatoryOptionMissing(const std::string &key)
        : FabberRunDataError("No value given for mandatory option: " + key)
    {
    }
};
struct DataNotFound : FabberRunDataError
{
    explicit DataNotFound(const std::string &key)
        : FabberRunDataError("Voxel data not found: " + key + " ()") /* rundata.h:756 */
    {
    }
};

/* ---- option description (rundata.h:45-99) ------------------------------------------------------- */
enum OptionType
{
    OPT_BOOL,
    OPT_STR,
    OPT_INT,
    OPT_FLOAT,
    OPT_FILE,
    OPT_IMAGE,
    OPT_TIMESERIES,
    OPT_MVN,
    OPT_MATRIX
};
struct OptionSpec
{
    std::string name;
    OptionType type;
    std::string description;
    bool optional;
    std::string def;
};
const char *option_type_name(OptionType t);

/* Host-side copies and output maps are memory-bound loops over 10^7..10^9 elements: split them over the
 * host cores (the reference does them single-threaded). fn(begin, end) must be independent per range. */
void parallel_for(size_t n, const std::function<void(size_t, size_t)> &fn, size_t min_chunk = 1 << 16);

template <class T> std::string stringify(const T &v)
{
    std::ostringstream s;
    s << v;
    return s.str();
}

/* ---- memory cache ---------------------------------------------------------------------------------
 * cudaMallocHost / cudaMalloc / cudaFree cost 0.1-1 s per GB and synchronise the device; a client that
 * calls set_data -> dorun -> get_data repeatedly would spend most of its time there (measured). Freed
 * blocks are parked by size and handed out again. */
void *cached_pinned_alloc(size_t bytes);
void cached_pinned_free(void *p, size_t bytes);
void *cached_device_alloc(size_t bytes); /* on the calling thread's current device */
void cached_device_free(void *p, size_t bytes);

/* ---- the GPUs of a run -------------------------------------------------------------------------------
 * One process drives every GPU of the box: voxelwise VB deals contiguous voxel ranges to the devices (voxels
 * are independent, inference_vb.cc:423-571), each with its own copy stream, kernels and result arrays.
 * FABBER_B200_DEVICES = "all" | comma list of ordinals picks them; unset: a process started by a one-process-
 * per-GPU launcher (LOCAL_RANK / LOCAL_WORLD_SIZE in the environment) stays on its own GPU, any other process
 * takes all visible ones. A device only takes part when it gets at least FABBER_B200_MIN_VOXELS_PER_DEVICE
 * (default 65536) voxels. */
const std::vector<int> &run_devices();
struct DeviceScope /* make `device` current for the calling thread, restore the previous one on exit */
{
    int prev;
    explicit DeviceScope(int device);
    ~DeviceScope();
    DeviceScope(const DeviceScope &) = delete;
    DeviceScope &operator=(const DeviceScope &) = delete;
};
void *copy_stream_of(int device); /* one non-blocking copy stream per device for the life of the process */

/* ---- host worker pool: threads are created once per process, not per call ---------------------------- */
size_t host_threads();
struct PoolJob;
PoolJob *pool_launch(size_t n_workers, const std::function<void()> &fn); /* fn runs once on each of n workers */
void pool_wait(PoolJob *job);                                            /* blocks, then frees the job */

/* ---- voxel data: rows x nvoxels, row t = volume t (rundata.h:628) --------------------------------
 * float32 throughout: every value enters (fabber_set_data) and leaves (fabber_get_data) the reference's
 * C API as float32, so nothing is lost; the buffers are pinned so device copies run at PCIe speed. */
struct VoxelData
{
    int rows;
    size_t cols;
    float *f;   /* pinned host memory [rows][cols] */
    /* optional device-resident copy (main data: uploaded while it is being set): the voxels [v0, v1) of part p
     * live on parts[p].device as a [rows][v1 - v0] array. One part = one GPU. */
    struct Part
    {
        int device;
        size_t v0, v1;
        float *dev;
        /* [own0, own1) is the sub-range this part is responsible for. Voxelwise VB: all of it. Spatial VB cuts
         * the volume into z-slabs: the part also holds the ghost planes just below and above its own planes
         * [z0, z1) (the same voxels are some other part's own), see fabber_cuda_vb_spatial_multi. */
        size_t own0, own1;
        int z0, z1; /* z0 == z1: not a z-slab */
    };
    std::vector<Part> parts;
    /* the upload goes block of voxels by block of voxels on the owning device's copy stream; `ready` is
     * recorded there after columns [v0, v1) have been queued, so a consumer can start on a block while later
     * ones travel */
    struct Block
    {
        size_t v0, v1;
        void *ready; /* fabber_cuda event */
        int part;
    };
    std::vector<Block> blocks;
    /* false while `f` has not been filled: the caller's buffer was page-locked and went to the devices by DMA
     * without a staging copy. ensure_host() downloads it on demand. */
    bool host_valid;
    void ensure_host();
    void wait_uploaded(int part); /* make that device's default stream wait for every block of the part */
    void release_device();        /* drain pending work, hand the device copies back to the cache */
    void upload_whole(int device); /* (re)load everything on to ONE device from the host copy */
    VoxelData();
    ~VoxelData();
    VoxelData(const VoxelData &) = delete;
    VoxelData &operator=(const VoxelData &) = delete;
    void alloc(int r, size_t c);
    size_t bytes() const { return (size_t)rows * cols * sizeof(float); }
    double at(int r, size_t c) const { return (double)f[(size_t)r * cols + c]; }
};

enum VoxelDataType /* rundata.h:37-43 */
{
    VDT_SCALAR,
    VDT_MVN
};

/* ---- run data (rundata.h:215-672 + rundata_array.cc) -------------------------------------------- */
class FabberRunData
{
public:
    FabberRunData();
    virtual ~FabberRunData();

    /* options (rundata.cc:427-600): booleans are "key present with empty value" */
    void Set(const std::string &key, const std::string &value);
    void SetBool(const std::string &key, bool value = true);
    void Unset(const std::string &key);
    bool HaveKey(const std::string &key) const;
    std::string GetString(const std::string &key);
    std::string GetStringDefault(const std::string &key, const std::string &def);
    bool GetBool(const std::string &key);
    int GetInt(const std::string &key, int min = INT32_MIN, int max = INT32_MAX);
    int GetIntDefault(const std::string &key, int def, int min = INT32_MIN, int max = INT32_MAX);
    double GetDouble(const std::string &key);
    double GetDoubleDefault(const std::string &key, double def);
    std::vector<std::string> GetStringList(const std::string &key); /* key1, key2, ... (rundata.cc:557-574) */
    std::vector<int> GetIntList(const std::string &key, int min = INT32_MIN, int max = INT32_MAX);

    /* command line and option files (rundata.cc:324-453): --key, --key=value, -f <file>, -@ <file>, --optfile */
    void Parse(int argc, char **argv);
    void ParseParamFile(const std::string &filename);
    void ParseOldStyleParamFile(const std::string &filename);
    void AddKeyEqualsValue(const std::string &exp, bool trim_comments = false);
    void LogParams();                  /* rundata.cc:239-246 */
    void CheckAllOptionsUsed();        /* "WARNING ONCE: Unused option specified: .." (rundata.cc:648-658) */
    void WarnOnce(const std::string &text); /* easylog.cc:105-127 */
    void ReissueWarnings();
    /* output directory: created on first use, '+' appended until it is new unless --overwrite (rundata.cc:660-737) */
    std::string GetOutputDir();

    /* extent, mask and coordinates (rundata_array.cc:23-66): voxel order x fastest, then y, then z */
    void SetExtent(int nx, int ny, int nz, const int *mask);
    const int *Extent() const { return m_extent; }
    size_t NumVoxels() const { return m_voxel_index.size(); }
    const std::vector<int> &VoxelIndex() const { return m_voxel_index; } /* grid offset of each masked voxel */
    const std::vector<int> &Coords() const { return m_coords; }         /* [3][N] */

    /* voxel data (rundata.cc:753-938, rundata_array.cc:68-133) */
    void SetVoxelDataArray(const std::string &key, int data_size, const float *data);
    void GetVoxelDataArray(const std::string &key, float *data);
    int GetVoxelDataSize(const std::string &key);
    const VoxelData &GetVoxelData(const std::string &key);
    const VoxelData &GetMainVoxelData(); /* "data", or data1..n combined by data-order (rundata.cc:753-905) */
    VoxelData &MutableMainVoxelData();
    /* called by SaveResults for every output it has stored under `key`; file-based front ends write it out
     * (rundata.cc:932-938, rundata_newimage.cc:140-183). The array front end keeps it in memory. */
    virtual void SaveVoxelData(const std::string &key, VoxelDataType type);
    VoxelData &NewVoxelData(const std::string &key, int rows); /* SaveVoxelData target (rundata.cc:932-938) */
    VoxelData &MutableVoxelData(const std::string &key);
    void ClearVoxelData(const std::string &key);

    /* Run: model + technique from their registries, Initialize -> DoCalculations -> SaveResults
     * (rundata.cc:248-311) */
    void Run(void (*progress_cb)(int, int) = nullptr);

    std::ostream &Log() { return m_log; }
    std::string LogText() const { return m_log.str(); }
    void ClearLog() { m_log.str(""); }
    void Progress(int v, int n)
    {
        if (m_progress)
            m_progress(v, n);
    }
    static void GetOptions(std::vector<OptionSpec> &opts);

protected:
    /* hook: `key` is not in memory - a file-based front end loads it (rundata_newimage.cc:89-138) and returns
     * true. The array front end has nothing to load from. */
    virtual bool LoadVoxelData(const std::string &key);

private:
    const VoxelData &GetMainVoxelDataMultiple();
    bool LooksSpatial() const;
    /* Speculative start. fabber_set_data("data") is normally the last call before fabber_dorun, and the options
     * are complete by then: the run is prepared right there and every block of voxels starts computing the
     * moment its upload is queued, while the host is still staging the blocks behind it. fabber_dorun adopts
     * that run if NOTHING was set or changed since (m_version), else it is thrown away and the run starts from
     * scratch - the result is the same either way. FABBER_B200_SPECULATE=0 turns it off. */
    struct SpeculativeRun;
    std::unique_ptr<SpeculativeRun> m_spec;
    unsigned long long m_version;
    bool m_in_run;
    void DiscardSpeculative();
    std::string m_outdir;
    std::set<std::string> m_used_params;
    std::map<std::string, int> m_warncount;
    std::map<std::string, std::string> m_params;
    std::map<std::string, std::unique_ptr<VoxelData>> m_voxel_data;
    int m_extent[3];
    bool m_have_extent;
    bool m_device_only_access;
    std::vector<int> m_mask, m_voxel_index, m_coords;
    std::ostringstream m_log;
    void (*m_progress)(int, int);
};
typedef FabberRunData FabberRunDataArray;

/* ---- file-based run data (rundata_newimage.h): NIfTI volumes in, NIfTI volumes out ------------------
 * Same name and behaviour as the reference's NEWIMAGE-backed class; the I/O underneath is nifti_io.cc. */
struct NiftiHeader;
class FabberRunDataNewimage : public FabberRunData
{
public:
    explicit FabberRunDataNewimage(bool compat_options = true);
    ~FabberRunDataNewimage();
    /* mask (binarised: > 1e-16) or, without one, the main data file gives the extent (rundata_newimage.cc:62-87) */
    void SetExtentFromData();
    void SaveVoxelData(const std::string &key, VoxelDataType type) override;

protected:
    bool LoadVoxelData(const std::string &filename) override;

private:
    std::unique_ptr<NiftiHeader> m_like; /* geometry every output copies (the mask's, else the first data file's) */
    bool m_have_mask;
};

/* ---- parameters and transforms (fwdmodel.h:24-57, transforms.h) ---------------------------------- */
struct DistParams
{
    double m_mean, m_var;
    DistParams(double mean = 0, double var = 1)
        : m_mean(mean)
        , m_var(var)
    {
    }
    double mean() const { return m_mean; }
    double var() const { return m_var; }
    double prec() const { return 1 / m_var; }
};
double transform_to_model(char code, double v);
double transform_to_fabber(char code, double v);
double transform_to_model_var(char code, double v);
double transform_to_fabber_var(char code, double v);

struct Parameter
{
    unsigned idx;
    std::string name;
    DistParams prior, post;
    char prior_type; /* 'N','I','A','M','m','P','p', '-' = model default */
    char transform;  /* 'I','L','S','F','A' */
    std::map<std::string, std::string> options;
    Parameter(unsigned i = 0, const std::string &n = "", DistParams pr = DistParams(), DistParams po = DistParams(),
        char ptype = 'N', char tr = 'I')
        : idx(i)
        , name(n)
        , prior(pr)
        , post(po)
        , prior_type(ptype)
        , transform(tr)
    {
    }
};
std::string ExpandPriorTypesString(std::string priors_str, unsigned num_params); /* priors.cc:35-106 */

/* ---- forward models ----------------------------------------------------------------------------- */
class MVNDist; /* host/operators.h */
class FwdModel
{
public:
    typedef FwdModel *(*NewInstanceFptr)(void);
    static FwdModel *NewFromName(const std::string &name); /* registry names: setup.cc:44-47 + "exp" */
    static std::vector<std::string> GetKnown();
    /* --loadmodels / fabber_load_models: register the models of a plug-in library (fwdmodel.cc:63-129) */
    static void LoadFromDynamicLibrary(const std::string &filename, std::ostream *log = nullptr);
    virtual ~FwdModel() {}
    virtual std::string ModelVersion() const { return "b200"; }
    virtual std::string GetDescription() const { return ""; }
    virtual void GetOptions(std::vector<OptionSpec> &) const {}
    virtual void Initialize(FabberRunData &rundata) = 0;
    /* default: built from the deprecated NameParams + HardcodedInitialDists + ardindices (fwdmodel.cc:339-363) */
    virtual void GetParameterDefaults(std::vector<Parameter> &params) const;
    /* one voxel's model prediction from model-space parameters (fwdmodel.h:149); default: forwards to the
     * deprecated Evaluate (fwdmodel.h:152) */
    virtual void EvaluateModel(const std::vector<double> &params, std::vector<double> &result, int n_times,
        const std::string &key = "") const;
    /* ---- deprecated API (fwdmodel.h:256-348; rundata.h:27 keeps it compiled in the reference): a model may
     * implement these instead of GetParameterDefaults / EvaluateModel ---- */
    virtual void Evaluate(const std::vector<double> &params, std::vector<double> &result) const
    {
        (void)params;
        (void)result;
    }
    virtual int NumParams() const { return (int)m_params.size(); }
    virtual void NameParams(std::vector<std::string> &names) const { (void)names; }
    virtual void HardcodedInitialDists(MVNDist &prior, MVNDist &posterior) const
    {
        (void)prior;
        (void)posterior;
    }
    std::vector<int> ardindices; /* 1-based indices of parameters under an ARD prior */
    virtual void GetOutputs(std::vector<std::string> &) const {}
    /* B200 addition: describe the compiled __device__ Evaluate hook for this model instance */
    virtual void GetDeviceModel(fabber_cuda_model &m) const = 0;

    /* fwdmodel.cc:210-282: defaults + param-spatial-priors + PSP_byname overrides, prior -> Fabber space */
    void GetParameters(FabberRunData &rundata, std::vector<Parameter> &params);
    /* fwdmodel.cc:365-382 */
    void EvaluateFabber(const std::vector<double> &theta, std::vector<double> &result, int n_times,
        const std::string &key = "") const;
    const std::vector<Parameter> &Params() const { return m_params; }

protected:
    std::vector<Parameter> m_params;
};

class LinearFwdModel : public FwdModel
{
public:
    std::string GetDescription() const override;
    void GetOptions(std::vector<OptionSpec> &opts) const override;
    void Initialize(FabberRunData &rundata) override;
    void GetParameterDefaults(std::vector<Parameter> &params) const override;
    void EvaluateModel(const std::vector<double> &p, std::vector<double> &result, int n_times,
        const std::string &key) const override;
    void GetDeviceModel(fabber_cuda_model &m) const override;

private:
    std::vector<double> m_design; /* [T][P] row-major */
    int m_ntimes = 0, m_nbasis = 0;
};
class PolynomialFwdModel : public FwdModel
{
public:
    std::string GetDescription() const override;
    void GetOptions(std::vector<OptionSpec> &opts) const override;
    void Initialize(FabberRunData &rundata) override;
    void GetParameterDefaults(std::vector<Parameter> &params) const override;
    void EvaluateModel(const std::vector<double> &p, std::vector<double> &result, int n_times,
        const std::string &key) const override;
    void GetDeviceModel(fabber_cuda_model &m) const override;

private:
    int m_degree = 0;
};
class ExpFwdModel : public FwdModel
{
public:
    std::string GetDescription() const override;
    void GetOptions(std::vector<OptionSpec> &opts) const override;
    void Initialize(FabberRunData &rundata) override;
    void GetParameterDefaults(std::vector<Parameter> &params) const override;
    void EvaluateModel(const std::vector<double> &p, std::vector<double> &result, int n_times,
        const std::string &key) const override;
    void GetDeviceModel(fabber_cuda_model &m) const override;

private:
    double m_dt = 1.0;
    int m_num = 1;
};

const char *fabber_b200_version();
/* bumped whenever FwdModel / Parameter / fabber_cuda_model / the launcher table change layout */
#define FABBER_B200_PLUGIN_ABI 2 /* 2: FwdModel gained the deprecated-API virtuals, ModelLaunchers gained sp_preload */

/* tools.cc:27-40: VEST or plain ASCII matrix file -> row-major values */
void read_matrix_file(const std::string &filename, std::vector<double> &values, int &rows, int &cols);

/* ---- inference technique ---------------------------------------------------------------------------- */
class Vb
{
public:
    static std::vector<std::string> GetKnownMethods(); /* nlls, spatialvb, vb (setup.cc:28-33) */
    static void GetOptions(std::vector<OptionSpec> &opts);
    static std::string GetDescription();
    static void GetOptions(std::vector<OptionSpec> &opts, const std::string &method); /* "nlls": inference_nlls.cc:31-45 */
    static std::string GetDescription(const std::string &method);
    void Initialize(FwdModel *model, FabberRunData &rundata); /* inference_vb.cc:100, inference.cc:62 */
    void DoCalculations(FabberRunData &rundata);              /* inference_vb.cc:360 - runs on the GPU */
    /* DoCalculations in three steps (it is Prepare + LaunchAll + Finish): the host library starts a run
     * speculatively while the series is still being uploaded, block by block */
    void Prepare(FabberRunData &rundata, VoxelData &data);
    void LaunchBlock(VoxelData &data, size_t block);
    void LaunchAll(VoxelData &data);
    void Finish(FabberRunData &rundata);
    bool LaunchesByBlock() const { return m_nvoxels > 0 && !m_spatial && !m_output_only; }
    const std::string &Description() const { return m_description; }
    void SaveResults(FabberRunData &rundata);                 /* inference_vb.cc:966, inference.cc:112 */

private:
    bool IsSpatial(FabberRunData &rundata, const std::vector<Parameter> &params) const; /* :334-358 */
    FwdModel *m_model = nullptr;
    int m_num_params = 0, m_noise_params = 0;
    bool m_ar = false, m_saveF = false, m_saveFsHistory = false, m_printF = false, m_needF = false;
    bool m_halt_bad_voxel = true;
    int m_nphis = 1, m_nalphas = 2;
    bool m_nlls = false, m_nlls_lm = false; /* --method=nlls through the same plumbing (inference_nlls.cc) */
    std::vector<double> m_nlls_start;
    /* results stay on the device, structure of arrays over voxels (the layout of include/fabber_cuda.h);
     * SaveResults turns them into float32 output maps there and downloads only what was asked for */
    size_t m_nvoxels = 0;
    int m_ntimes = 0, m_nn = 0, m_fhist_len = 0;
    fabber_cuda_vb_problem m_prob;
    std::vector<unsigned char> m_pattern, m_masked;
    struct DevArray
    {
        void *p = nullptr;
        size_t bytes = 0;
    };
    /* one context per GPU of the run: its contiguous voxel range [v0, v1), the device copy of that range's
     * series ([T][v1 - v0]) and its result arrays (all strides v1 - v0) */
    struct DevCtx
    {
        int device = 0;
        size_t v0 = 0, v1 = 0;
        size_t own0 = 0, own1 = 0; /* the columns this device publishes (== [v0, v1) unless it is a z-slab) */
        const float *data = nullptr;
        DevArray mean, cov, noise, F, hist, its, status;
    };
    std::vector<DevCtx> m_ctx;
    struct Scratch
    {
        int device;
        DevArray d;
    };
    std::vector<Scratch> m_scratch; /* per-voxel inputs on the devices (image priors, restart state, coordinates) */
    std::vector<fabber_cuda_vb_buffers> m_bufs;
    std::vector<fabber_cuda_vb_problem> m_probs;
    bool m_spatial = false, m_slabs = false, m_output_only = false;
    int m_launch_rc = 0;
    std::string m_launch_error, m_description;
    void FreeScratch();
    void ReleaseDevice();

public:
    ~Vb()
    {
        FreeScratch();
        ReleaseDevice();
    }
};

} // namespace fabber_b200
