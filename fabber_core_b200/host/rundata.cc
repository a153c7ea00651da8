/*
 * rundata.cc - FabberRunData / FabberRunDataArray: option map, extent + mask, voxel data T x Nvox.
 * Behaviour follows rundata.cc:427-600 (options), rundata_array.cc:23-133 (array I/O) and
 * rundata.cc:248-311 (Run). The main data keeps the caller's float32 samples in pinned host memory so
 * that the copy to the GPU runs at full PCIe speed; outputs are double and converted to float on read,
 * as in the reference.
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <ctime>

#include <cerrno>
#include <fstream>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <unordered_map>

#include <sched.h>
#include <sys/stat.h>
#include <immintrin.h>
#include <sys/types.h>
#include <unistd.h>

#include "fabber_host.h"
#include "operators.h"

namespace fabber_b200
{
const char *fabber_b200_version() { return "b200-r1 (VB path of fabber_core on sm_100a)"; }

size_t host_threads()
{
    /* the CPUs this process may run on (what `nproc` reports), not every CPU of the machine, divided by the
     * number of processes a one-process-per-GPU launcher started on this host (LOCAL_WORLD_SIZE: eight ranks
     * x 32 staging threads on 32 cores was measured to triple every rank's copy times);
     * FABBER_B200_HOST_THREADS overrides */
    if (const char *e = getenv("FABBER_B200_HOST_THREADS"))
        if (atol(e) > 0)
            return (size_t)std::min<long>(atol(e), 256);
    size_t threads = std::thread::hardware_concurrency();
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0)
        threads = threads ? std::min<size_t>(threads, (size_t)CPU_COUNT(&set)) : (size_t)CPU_COUNT(&set);
    if (threads == 0)
        threads = 4;
    if (const char *lw = getenv("LOCAL_WORLD_SIZE"))
        if (atol(lw) > 1 && !getenv("FABBER_B200_DEVICES"))
            threads = std::max<size_t>(2, threads / (size_t)atol(lw));
    if (threads > 64)
        threads = 64;
    return threads;
}

/* ---- worker pool ------------------------------------------------------------------------------------
 * Threads are created on first use and parked on a condition variable between jobs (a std::thread per call
 * per worker cost ~50 us each, 32 of them per set_data / get_data). */
struct PoolJob
{
    std::function<void()> fn;
    std::atomic<size_t> remaining;
    std::mutex mu;
    std::condition_variable done;
};
namespace
{
struct HostPool
{
    std::mutex mu;
    std::condition_variable wake;
    std::deque<PoolJob *> queue; /* one entry per worker slot still to be claimed */
    std::vector<std::thread> threads;
    bool stop = false;
    void grow(size_t n)
    {
        while (threads.size() < n)
            threads.emplace_back([this]() { loop(); });
    }
    void loop()
    {
        for (;;)
        {
            PoolJob *job = nullptr;
            {
                std::unique_lock<std::mutex> lock(mu);
                wake.wait(lock, [this]() { return stop || !queue.empty(); });
                if (stop && queue.empty())
                    return;
                job = queue.front();
                queue.pop_front();
            }
            job->fn();
            if (job->remaining.fetch_sub(1, std::memory_order_acq_rel) == 1)
            {
                std::lock_guard<std::mutex> lock(job->mu);
                job->done.notify_all();
            }
        }
    }
    ~HostPool()
    {
        {
            std::lock_guard<std::mutex> lock(mu);
            stop = true;
        }
        wake.notify_all();
        for (size_t i = 0; i < threads.size(); i++)
            threads[i].join();
    }
};
HostPool &host_pool()
{
    static HostPool *pool = new HostPool(); /* leaked on purpose: workers may outlive static destruction order */
    return *pool;
}
} // namespace

PoolJob *pool_launch(size_t n_workers, const std::function<void()> &fn)
{
    PoolJob *job = new PoolJob();
    job->fn = fn;
    job->remaining.store(n_workers, std::memory_order_relaxed);
    if (n_workers == 0)
        return job;
    HostPool &pool = host_pool();
    {
        std::lock_guard<std::mutex> lock(pool.mu);
        pool.grow(std::min<size_t>(std::max(n_workers, pool.threads.size()), 256));
        for (size_t i = 0; i < n_workers; i++)
            pool.queue.push_back(job);
    }
    pool.wake.notify_all();
    return job;
}
void pool_wait(PoolJob *job)
{
    if (!job)
        return;
    {
        std::unique_lock<std::mutex> lock(job->mu);
        job->done.wait(lock, [job]() { return job->remaining.load(std::memory_order_acquire) == 0; });
    }
    delete job;
}

void parallel_for(size_t n, const std::function<void(size_t, size_t)> &fn, size_t min_chunk)
{
    size_t threads = host_threads();
    if (n / min_chunk + 1 < threads)
        threads = n / min_chunk + 1;
    if (threads <= 1)
    {
        fn(0, n);
        return;
    }
    const size_t per = (n + threads - 1) / threads;
    std::atomic<size_t> next(0);
    pool_wait(pool_launch(threads, [&]() {
        const size_t t = next.fetch_add(1, std::memory_order_relaxed);
        const size_t b = t * per, e = std::min(n, b + per);
        if (b < e)
            fn(b, e);
    }));
}

/* Staging copy for SetVoxelDataArray: the destination is pinned memory that only the DMA engine reads next, so
 * the stores go around the cache (no read-for-ownership of the destination lines: 2 bytes of DRAM traffic per
 * byte copied instead of 3). glibc's memcpy only does this above a threshold far larger than the ~1 MB pieces
 * staged here. Falls back to memcpy without AVX2. */
__attribute__((target("avx2"))) static void stream_copy_avx2(float *dst, const float *src, size_t n)
{
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 31))
    {
        dst[i] = src[i];
        i++;
    }
    for (; i + 32 <= n; i += 32)
    {
        const __m256i a = _mm256_loadu_si256((const __m256i *)(src + i));
        const __m256i b = _mm256_loadu_si256((const __m256i *)(src + i + 8));
        const __m256i c = _mm256_loadu_si256((const __m256i *)(src + i + 16));
        const __m256i d = _mm256_loadu_si256((const __m256i *)(src + i + 24));
        _mm256_stream_si256((__m256i *)(dst + i), a);
        _mm256_stream_si256((__m256i *)(dst + i + 8), b);
        _mm256_stream_si256((__m256i *)(dst + i + 16), c);
        _mm256_stream_si256((__m256i *)(dst + i + 24), d);
    }
    for (; i < n; i++)
        dst[i] = src[i];
    _mm_sfence(); /* the copy must be globally visible before the host->device copy is queued */
}
void stage_copy(float *dst, const float *src, size_t n)
{
    static const bool avx2 = __builtin_cpu_supports("avx2") && !getenv("FABBER_B200_NO_STREAM_COPY");
    if (avx2 && n >= 4096)
        stream_copy_avx2(dst, src, n);
    else
        memcpy(dst, src, n * sizeof(float));
}

const char *option_type_name(OptionType t)
{
    switch (t)
    {
    case OPT_BOOL:
        return "BOOL";
    case OPT_STR:
        return "STR";
    case OPT_INT:
        return "INT";
    case OPT_FLOAT:
        return "FLOAT";
    case OPT_FILE:
        return "FILE";
    case OPT_IMAGE:
        return "IMAGE";
    case OPT_TIMESERIES:
        return "TIMESERIES";
    case OPT_MVN:
        return "MVN";
    case OPT_MATRIX:
        return "MATRIX";
    }
    return "UNKNOWN";
}

/* ---- memory cache (see fabber_host.h) ------------------------------------------------------------ */
namespace
{
struct BlockCache
{
    std::mutex mu;
    std::multimap<size_t, void *> free_blocks;
    size_t cached_bytes = 0;
};
BlockCache g_pinned;
/* device blocks are parked per device; the owner of every live block is remembered so that a block can be
 * handed back from any thread, whatever its current device */
std::mutex g_device_mu;
std::map<int, BlockCache> g_device;
std::unordered_map<void *, int> g_device_owner;
const size_t CACHE_LIMIT = (size_t)48 << 30; /* per kind; beyond it blocks really are freed */

void *cache_get(BlockCache &c, size_t bytes)
{
    std::lock_guard<std::mutex> lock(c.mu);
    std::multimap<size_t, void *>::iterator it = c.free_blocks.find(bytes);
    if (it == c.free_blocks.end())
        return nullptr;
    void *p = it->second;
    c.free_blocks.erase(it);
    c.cached_bytes -= bytes;
    return p;
}
bool cache_put(BlockCache &c, void *p, size_t bytes)
{
    std::lock_guard<std::mutex> lock(c.mu);
    if (c.cached_bytes + bytes > CACHE_LIMIT)
        return false;
    c.free_blocks.insert(std::make_pair(bytes, p));
    c.cached_bytes += bytes;
    return true;
}
} // namespace

/* Without a CUDA device there is nothing to pin for: plain memory, so that option handling, file I/O and
 * error reporting still work (and are testable) on a machine without a GPU. The inference itself has no
 * such fallback - Vb::DoCalculations fails with the device error. Fixed for the life of the process. */
static bool have_device()
{
    static const bool yes = fabber_cuda_device_count() > 0;
    return yes;
}

void *cached_pinned_alloc(size_t bytes)
{
    if (bytes == 0)
        bytes = 1;
    void *p = cache_get(g_pinned, bytes);
    if (!p)
        p = have_device() ? fabber_cuda_host_alloc(bytes) : malloc(bytes);
    if (!p)
        throw FabberInternalError(std::string("Could not allocate pinned host memory: ") + fabber_cuda_last_error());
    return p;
}
void cached_pinned_free(void *p, size_t bytes)
{
    if (p && !cache_put(g_pinned, p, bytes ? bytes : 1))
    {
        if (have_device())
            fabber_cuda_host_free(p);
        else
            free(p);
    }
}
void *cached_device_alloc(size_t bytes)
{
    if (bytes == 0)
        bytes = 1;
    const int dev = fabber_cuda_get_device();
    BlockCache *cache;
    {
        std::lock_guard<std::mutex> lock(g_device_mu);
        cache = &g_device[dev];
    }
    void *p = cache_get(*cache, bytes);
    if (!p)
        p = fabber_cuda_malloc(bytes);
    if (!p)
        throw FabberInternalError(std::string("GPU allocation failed: ") + fabber_cuda_last_error());
    std::lock_guard<std::mutex> lock(g_device_mu);
    g_device_owner[p] = dev;
    return p;
}
void cached_device_free(void *p, size_t bytes)
{
    if (!p)
        return;
    BlockCache *cache;
    int dev;
    {
        std::lock_guard<std::mutex> lock(g_device_mu);
        std::unordered_map<void *, int>::iterator it = g_device_owner.find(p);
        dev = it == g_device_owner.end() ? fabber_cuda_get_device() : it->second;
        if (it != g_device_owner.end())
            g_device_owner.erase(it);
        cache = &g_device[dev];
    }
    if (!cache_put(*cache, p, bytes ? bytes : 1))
    {
        DeviceScope scope(dev);
        fabber_cuda_free(p);
    }
}

/* ---- the GPUs of a run (see fabber_host.h) ------------------------------------------------------- */
const std::vector<int> &run_devices()
{
    static const std::vector<int> devices = []() {
        std::vector<int> d;
        const int n = fabber_cuda_device_count();
        if (n <= 0)
            return d;
        const char *env = getenv("FABBER_B200_DEVICES");
        if (env && *env && strcmp(env, "all") != 0)
        {
            std::istringstream in(env);
            std::string item;
            while (std::getline(in, item, ','))
            {
                const int v = atoi(item.c_str());
                if (!item.empty() && v >= 0 && v < n) /* a repeated ordinal gives that GPU two ranges (tests on one GPU) */
                    d.push_back(v);
            }
        }
        else if (!env && getenv("LOCAL_RANK") && getenv("LOCAL_WORLD_SIZE") && atoi(getenv("LOCAL_WORLD_SIZE")) > 1)
            d.push_back(std::max(0, fabber_cuda_get_device())); /* one process per GPU: the launcher's choice */
        else
            for (int i = 0; i < n; i++)
                d.push_back(i);
        if (d.empty())
            d.push_back(std::max(0, fabber_cuda_get_device()));
        return d;
    }();
    return devices;
}
DeviceScope::DeviceScope(int device)
    : prev(fabber_cuda_get_device())
{
    if (device >= 0 && device != prev)
        fabber_cuda_set_device(device);
    else
        prev = -1;
}
DeviceScope::~DeviceScope()
{
    if (prev >= 0)
        fabber_cuda_set_device(prev);
}
void *copy_stream_of(int device)
{
    static std::mutex mu;
    static std::map<int, void *> streams;
    std::lock_guard<std::mutex> lock(mu);
    std::map<int, void *>::iterator it = streams.find(device);
    if (it != streams.end())
        return it->second;
    DeviceScope scope(device);
    void *s = fabber_cuda_stream_create();
    streams[device] = s;
    return s;
}

VoxelData::VoxelData()
    : rows(0)
    , cols(0)
    , f(nullptr)
    , host_valid(true)
{
}

void VoxelData::wait_uploaded(int part)
{
    for (size_t i = 0; i < blocks.size(); i++)
        if (blocks[i].part == part)
            fabber_cuda_stream_wait_event(nullptr, blocks[i].ready);
}

void VoxelData::release_device()
{
    /* work queued on the devices may still read these blocks: drain it before they can be handed out again */
    for (size_t p = 0; p < parts.size(); p++)
    {
        DeviceScope scope(parts[p].device);
        fabber_cuda_stream_sync(copy_stream_of(parts[p].device));
        fabber_cuda_stream_sync(nullptr);
        cached_device_free(parts[p].dev, (size_t)rows * (parts[p].v1 - parts[p].v0) * sizeof(float));
    }
    for (size_t i = 0; i < blocks.size(); i++)
        fabber_cuda_event_destroy(blocks[i].ready);
    parts.clear();
    blocks.clear();
}

void VoxelData::ensure_host()
{
    if (host_valid)
        return;
    for (size_t p = 0; p < parts.size(); p++)
    {
        const Part &pt = parts[p];
        DeviceScope scope(pt.device);
        const size_t w = pt.v1 - pt.v0;
        fabber_cuda_stream_sync(copy_stream_of(pt.device));
        if (fabber_cuda_memcpy2d_d2h(f + pt.v0, cols * sizeof(float), pt.dev, w * sizeof(float), w * sizeof(float),
                (size_t)rows, nullptr)
                != FABBER_CUDA_OK
            || fabber_cuda_stream_sync(nullptr) != FABBER_CUDA_OK)
            throw FabberInternalError(std::string("copying data back from the GPU: ") + fabber_cuda_last_error());
    }
    host_valid = true;
}

void VoxelData::upload_whole(int device)
{
    if (parts.size() == 1 && parts[0].device == device && parts[0].v0 == 0 && parts[0].v1 == cols)
        return;
    ensure_host();
    release_device();
    DeviceScope scope(device);
    Part pt;
    pt.device = device;
    pt.v0 = pt.own0 = 0;
    pt.v1 = pt.own1 = cols;
    pt.z0 = pt.z1 = 0;
    pt.dev = (float *)cached_device_alloc(bytes());
    parts.push_back(pt);
    if (fabber_cuda_memcpy_h2d(pt.dev, f, bytes(), nullptr) != FABBER_CUDA_OK)
        throw FabberInternalError(std::string("copying data to the GPU: ") + fabber_cuda_last_error());
}

VoxelData::~VoxelData()
{
    release_device();
    cached_pinned_free(f, bytes());
}
void VoxelData::alloc(int r, size_t c)
{
    cached_pinned_free(f, bytes());
    f = nullptr;
    rows = r;
    cols = c;
    f = (float *)cached_pinned_alloc(bytes());
}

struct FabberRunData::SpeculativeRun
{
    std::unique_ptr<FwdModel> model;
    std::unique_ptr<Vb> vb;
    unsigned long long version = 0;
    const VoxelData *data = nullptr;
    bool by_block = false;
};

FabberRunData::FabberRunData()
    : m_version(0)
    , m_in_run(false)
    , m_have_extent(false)
    , m_device_only_access(false)
    , m_progress(nullptr)
{
    m_extent[0] = m_extent[1] = m_extent[2] = 0;
}
FabberRunData::~FabberRunData() { DiscardSpeculative(); }
void FabberRunData::DiscardSpeculative()
{
    m_spec.reset(); /* ~Vb drains its devices and hands the buffers back */
}

void FabberRunData::Set(const std::string &key, const std::string &value)
{
    m_version++;
    m_params[key] = value;
}
void FabberRunData::SetBool(const std::string &key, bool value)
{
    m_version++;
    if (value)
        m_params[key] = "";
    else
        m_params.erase(key);
}
void FabberRunData::Unset(const std::string &key)
{
    m_version++;
    m_params.erase(key);
}
bool FabberRunData::HaveKey(const std::string &key) const { return m_params.count(key) > 0; }

std::string FabberRunData::GetString(const std::string &key)
{
    if (m_params.count(key) == 0)
        throw MandatoryOptionMissing(key);
    if (m_params[key] == "")
        throw InvalidOptionValue(key, "<no value>", "Value must be given");
    m_used_params.insert(key);
    return m_params[key];
}
std::string FabberRunData::GetStringDefault(const std::string &key, const std::string &def)
{
    m_used_params.insert(key);
    if (m_params.count(key) == 0)
        return def;
    return m_params[key];
}
bool FabberRunData::GetBool(const std::string &key)
{
    m_used_params.insert(key);
    if (m_params.count(key) == 0)
        return false;
    if (m_params[key] == "")
        return true;
    throw InvalidOptionValue(key, m_params[key], "Value should not be given for boolean option");
}
static bool parse_int(const std::string &s, int &out)
{
    std::istringstream i(s);
    char c;
    if (!(i >> out) || (i >> c))
        return false;
    return true;
}
static bool parse_double(const std::string &s, double &out)
{
    std::istringstream i(s);
    char c;
    if (!(i >> out) || (i >> c))
        return false;
    return true;
}
int FabberRunData::GetInt(const std::string &key, int min, int max)
{
    std::string val = GetString(key);
    int i;
    if (!parse_int(val, i))
        throw InvalidOptionValue(key, val, "Failed to convert to required type"); /* rundata.h:768 */
    if (i < min)
        throw InvalidOptionValue(key, val, "Minimum " + stringify(min));
    if (i > max)
        throw InvalidOptionValue(key, val, "Maximum " + stringify(max));
    return i;
}
int FabberRunData::GetIntDefault(const std::string &key, int def, int min, int max)
{
    return m_params.count(key) == 0 ? def : GetInt(key, min, max);
}
double FabberRunData::GetDouble(const std::string &key)
{
    std::string val = GetString(key);
    double d;
    if (!parse_double(val, d))
        throw InvalidOptionValue(key, val, "Failed to convert to required type"); /* rundata.h:768 */
    return d;
}
double FabberRunData::GetDoubleDefault(const std::string &key, double def)
{
    return m_params.count(key) == 0 ? def : GetDouble(key);
}
std::vector<std::string> FabberRunData::GetStringList(const std::string &prefix)
{
    std::vector<std::string> ret;
    if (HaveKey(prefix))
        ret.push_back(GetString(prefix));
    else
        for (int n = 1; HaveKey(prefix + stringify(n)); n++)
            ret.push_back(GetString(prefix + stringify(n)));
    return ret;
}
std::vector<int> FabberRunData::GetIntList(const std::string &prefix, int min, int max)
{
    std::vector<int> ret;
    if (HaveKey(prefix))
        ret.push_back(GetInt(prefix, min, max));
    else
        for (int n = 1; HaveKey(prefix + stringify(n)); n++)
            ret.push_back(GetInt(prefix + stringify(n), min, max));
    return ret;
}

/* rundata_array.cc:23-66 */
void FabberRunData::SetExtent(int nx, int ny, int nz, const int *mask)
{
    if (nx <= 0 || ny <= 0 || nz <= 0)
        throw FabberRunDataError("Dimensions must be >0");
    m_version++;
    DiscardSpeculative();
    m_extent[0] = nx;
    m_extent[1] = ny;
    m_extent[2] = nz;
    m_have_extent = true;
    const size_t nv = (size_t)nx * ny * nz;
    if (mask)
        m_mask.assign(mask, mask + nv);
    else
        m_mask.assign(nv, 1);
    m_voxel_index.clear();
    std::vector<int> cx, cy, cz;
    size_t i = 0;
    for (int z = 0; z < nz; z++)
        for (int y = 0; y < ny; y++)
            for (int x = 0; x < nx; x++, i++)
                if (m_mask[i] != 0)
                {
                    m_voxel_index.push_back((int)i);
                    cx.push_back(x);
                    cy.push_back(y);
                    cz.push_back(z);
                }
    m_coords.clear();
    m_coords.insert(m_coords.end(), cx.begin(), cx.end());
    m_coords.insert(m_coords.end(), cy.begin(), cy.end());
    m_coords.insert(m_coords.end(), cz.begin(), cz.end());
}

/* do the options set so far describe a spatial run (method=spatialvb or a spatial prior type)? Spatial VB
 * couples the voxels, so its series is not dealt to several devices here. A wrong guess costs one re-upload
 * in Vb::DoCalculations, never a wrong result. */
bool FabberRunData::LooksSpatial() const
{
    std::map<std::string, std::string>::const_iterator it = m_params.find("method");
    if (it != m_params.end() && it->second == "spatialvb")
        return true;
    it = m_params.find("param-spatial-priors");
    if (it != m_params.end() && it->second.find_first_of("MmPp") != std::string::npos)
        return true;
    for (it = m_params.begin(); it != m_params.end(); ++it)
        if (it->first.compare(0, 10, "PSP_byname") == 0 && it->first.size() > 5
            && it->first.compare(it->first.size() - 5, 5, "_type") == 0 && it->second.find_first_of("MmPp") != std::string::npos)
            return true;
    return false;
}

/* rundata_array.cc:100-133: float[t][z][y][x] -> T x Nvox */
void FabberRunData::SetVoxelDataArray(const std::string &key, int data_size, const float *data)
{
    if (!m_have_extent)
        throw FabberRunDataError("Extent must be set before voxel data");
    const size_t n_grid = (size_t)m_extent[0] * m_extent[1] * m_extent[2];
    const size_t N = m_voxel_index.size();
    m_version++;
    DiscardSpeculative(); /* before its series can go away */
    m_voxel_data.erase(key); /* hand the old blocks back to the cache before asking for new ones */
    if (key.compare(0, 4, "data") == 0)
        m_voxel_data.erase("@maindata"); /* a stale combination of data1..n */
    std::unique_ptr<VoxelData> vd(new VoxelData());
    vd->alloc(data_size, N);
    float *dst_all = vd->f;
    const std::vector<int> &index = m_voxel_index;
    /* The main series goes straight on to the GPUs, BLOCK OF VOXELS BY BLOCK OF VOXELS: the voxel list is cut
     * into one contiguous range per device (voxels are independent in voxelwise VB), each range into blocks;
     * all host cores stage the columns [v0, v1) of every row of a block into pinned memory, the block's strided
     * host->device copy is queued on its device's copy stream, an event marks it - and the next block (for the
     * next device: the blocks are dealt round-robin so every PCIe link is busy) is staged while that one
     * travels. Voxelwise VB (Vb::DoCalculations) then starts each block's kernel on its event, so the
     * arithmetic of block k also overlaps the transfer of the blocks after it. */
    /* ("data" itself, or the file the data option names when a file-based front end loads it) */
    const bool is_main = key == "data" || (m_params.count("data") && m_params["data"] == key);
    const bool upload = is_main && N > 0 && fabber_cuda_device_count() > 0;
    const size_t T = (size_t)data_size;
    std::vector<int> devs(1, 0);
    if (upload)
    {
        devs = run_devices();
        const char *mv = getenv("FABBER_B200_MIN_VOXELS_PER_DEVICE");
        const size_t min_per = (mv && atol(mv) > 0) ? (size_t)atol(mv) : 65536;
        size_t g = std::min(devs.size(), std::max<size_t>(1, N / min_per));
        devs.resize(g);
    }
    const bool spatial_like = LooksSpatial();
    size_t G = devs.size();
    /* ~96 MB per block, at most 64 blocks per device, at least 64k voxels each (a block is also one launch) */
    const char *block_env = getenv("FABBER_B200_UPLOAD_BLOCK_MB"); /* tuning / test knob; 32-128 MB measure alike */
    const size_t block_mb = (block_env && atol(block_env) > 0) ? (size_t)atol(block_env) : 96;
    if (upload)
    {
        bool slabs = false;
        if (spatial_like && G > 1)
        {
            /* spatial VB couples neighbouring voxels: the volume is cut into z-slabs (the voxel list is z-major,
             * so a slab is a contiguous range), each device also gets the planes just below and above its own */
            const int nz = m_extent[2];
            G = std::min<size_t>(G, (size_t)nz);
            std::vector<size_t> first(nz + 2, N); /* first list position with z >= k */
            {
                size_t v = 0;
                for (int k = 0; k <= nz + 1; k++)
                {
                    while (v < N && m_coords[2 * N + v] < k)
                        v++;
                    first[k] = v;
                }
            }
            std::vector<VoxelData::Part> cut;
            for (size_t g = 0; g < G; g++)
            {
                const int base = nz / (int)G, extra = nz % (int)G;
                const int z0 = (int)g * base + std::min((int)g, extra), z1 = z0 + base + ((int)g < extra ? 1 : 0);
                VoxelData::Part pt;
                pt.device = devs[g];
                pt.z0 = z0;
                pt.z1 = z1;
                pt.own0 = first[z0];
                pt.own1 = first[z1];
                pt.v0 = first[std::max(z0 - 1, 0)];
                pt.v1 = first[std::min(z1 + 1, nz)];
                pt.dev = nullptr;
                cut.push_back(pt);
            }
            slabs = true;
            for (size_t g = 0; g < cut.size(); g++)
                if (cut[g].own1 <= cut[g].own0)
                    slabs = false; /* a slab without a voxel of the mask: one device runs it all */
            if (slabs)
                vd->parts = cut;
            else
                G = 1;
        }
        if (!slabs)
        {
            const size_t per_part = ((N + G - 1) / G + 127) / 128 * 128; /* whole CTAs of 128 voxels */
            for (size_t g = 0; g < G; g++)
            {
                VoxelData::Part pt;
                pt.device = devs[g];
                pt.v0 = std::min(N, g * per_part);
                pt.v1 = std::min(N, pt.v0 + per_part);
                pt.own0 = pt.v0;
                pt.own1 = pt.v1;
                pt.z0 = pt.z1 = 0;
                pt.dev = nullptr;
                if (pt.v1 > pt.v0)
                    vd->parts.push_back(pt);
            }
        }
    }
    std::vector<VoxelData::Block> blocks; /* round-robin over the devices */
    size_t max_block = 0;
    {
        std::vector<std::vector<VoxelData::Block>> per_part_blocks(upload ? vd->parts.size() : 1);
        size_t most = 0;
        for (size_t g = 0; g < per_part_blocks.size(); g++)
        {
            const size_t p0 = upload ? vd->parts[g].v0 : 0, p1 = upload ? vd->parts[g].v1 : N, pn = p1 - p0;
            size_t n_blocks = std::min<size_t>(64, std::max<size_t>(1, T * pn * sizeof(float) / (block_mb << 20)));
            n_blocks = std::max<size_t>(1, std::min(n_blocks, pn / 65536));
            if (!upload)
                n_blocks = 1;
            const size_t per_block = ((pn + n_blocks - 1) / n_blocks + 127) / 128 * 128;
            for (size_t bi = 0; bi < n_blocks; bi++)
            {
                VoxelData::Block blk;
                blk.v0 = p0 + bi * per_block;
                blk.v1 = std::min(p1, blk.v0 + per_block);
                blk.ready = nullptr;
                blk.part = (int)g;
                if (blk.v0 < blk.v1)
                {
                    per_part_blocks[g].push_back(blk);
                    max_block = std::max(max_block, blk.v1 - blk.v0);
                }
            }
            most = std::max(most, per_part_blocks[g].size());
        }
        for (size_t bi = 0; bi < most; bi++)
            for (size_t g = 0; g < per_part_blocks.size(); g++)
                if (bi < per_part_blocks[g].size())
                    blocks.push_back(per_part_blocks[g][bi]);
    }
    for (size_t g = 0; g < vd->parts.size(); g++)
    {
        DeviceScope scope(vd->parts[g].device);
        vd->parts[g].dev = (float *)cached_device_alloc(T * (vd->parts[g].v1 - vd->parts[g].v0) * sizeof(float));
    }
    /* speculative start (see fabber_host.h): prepare the run now; every block's kernel is queued right behind
     * its upload. Any failure here just means no speculation - fabber_dorun then reports it properly. */
    std::unique_ptr<SpeculativeRun> spec;
    {
        const char *se = getenv("FABBER_B200_SPECULATE");
        if (upload && !m_in_run && !(se && se[0] == '0') && m_params.count("model") && m_params.count("method")
            && m_params["method"] == "vb" && m_params.count("noise") && !spatial_like)
        {
            try
            {
                spec.reset(new SpeculativeRun());
                spec->model.reset(FwdModel::NewFromName(m_params["model"]));
                spec->model->Initialize(*this);
                std::vector<Parameter> params;
                spec->model->GetParameters(*this, params);
                spec->vb.reset(new Vb());
                spec->vb->Initialize(spec->model.get(), *this);
                spec->vb->Prepare(*this, *vd);
                spec->by_block = spec->vb->LaunchesByBlock();
                if (!spec->by_block)
                    spec.reset();
            }
            catch (...)
            {
                spec.reset();
            }
        }
    }
    auto queue_block = [&](VoxelData::Block &blk, const float *src, size_t src_pitch_elems) {
        const VoxelData::Part &pt = vd->parts[blk.part];
        DeviceScope scope(pt.device);
        void *cs = copy_stream_of(pt.device);
        const size_t w = blk.v1 - blk.v0, pw = pt.v1 - pt.v0;
        blk.ready = fabber_cuda_event_create();
        int rc = blk.ready ? fabber_cuda_memcpy2d_h2d(pt.dev + (blk.v0 - pt.v0), pw * sizeof(float), src + blk.v0,
                                 src_pitch_elems * sizeof(float), w * sizeof(float), T, cs)
                           : FABBER_CUDA_ERR_CUDA;
        if (rc == FABBER_CUDA_OK)
            rc = fabber_cuda_event_record(blk.ready, cs);
        vd->blocks.push_back(blk);
        if (rc == FABBER_CUDA_OK && spec)
            spec->vb->LaunchBlock(*vd, vd->blocks.size() - 1);
        return rc;
    };
    auto keep_spec = [&]() {
        if (spec)
        {
            spec->version = m_version;
            spec->data = vd.get();
            m_spec = std::move(spec);
        }
    };

    /* The caller's buffer is page-locked (cudaMallocHost / cudaHostRegister) and in voxel-list order (full
     * mask): the DMA engines read it in place, no staging copy - with several GPUs their PCIe links together
     * move data faster than the host cores can copy it. The call still returns only when every copy has
     * landed (the caller may reuse its buffer afterwards, as with the reference), unless the caller has
     * promised to leave it alone until fabber_dorun returns (FABBER_B200_ASYNC_SET_DATA=1). With ONE device
     * the staged path is kept: its copy (measured 76 GB/s) outruns one PCIe link and lets set_data return
     * while the tail of the upload is still in flight. */
    const char *direct_env = getenv("FABBER_B200_DIRECT_UPLOAD"); /* 0 = never, 1 = whenever possible */
    bool direct = upload && N == n_grid && fabber_cuda_host_is_pinned(data) == 1 && vd->parts.size() >= 2;
    if (direct_env && direct_env[0] == '0')
        direct = false;
    if (direct_env && direct_env[0] == '1')
        direct = upload && N == n_grid && fabber_cuda_host_is_pinned(data) == 1;
    if (direct)
    {
        for (size_t b = 0; b < blocks.size(); b++)
            if (queue_block(blocks[b], data, N) != FABBER_CUDA_OK)
                throw FabberInternalError(std::string("copying data to the GPU: ") + fabber_cuda_last_error());
        vd->host_valid = false;
        const char *async_env = getenv("FABBER_B200_ASYNC_SET_DATA");
        if (!(async_env && async_env[0] == '1'))
            for (size_t b = 0; b < vd->blocks.size(); b++)
                if (fabber_cuda_event_sync(vd->blocks[b].ready) != FABBER_CUDA_OK)
                    throw FabberInternalError(std::string("copying data to the GPU: ") + fabber_cuda_last_error());
        keep_spec();
        m_voxel_data[key] = std::move(vd);
        return;
    }

    /* One pool of workers for the whole volume, no barrier between blocks: work item = (block, row, piece of
     * the block's columns) - contiguous runs in both source and destination - handed out block-major from one
     * counter; a per-block count of unfinished items tells the calling thread when a block is staged, and it
     * queues that block's copy while the workers are already on the next one. */
    const size_t piece = (size_t)1 << 18, pieces = std::max<size_t>(1, (max_block + piece - 1) / piece), per_items = T * pieces;
    const size_t real_blocks = blocks.size(); /* empty mask: nothing to stage */
    std::vector<std::atomic<size_t>> unfinished(real_blocks);
    for (size_t b = 0; b < real_blocks; b++)
        unfinished[b].store(per_items, std::memory_order_relaxed);
    std::atomic<size_t> next(0);
    std::mutex staged_mu;
    std::condition_variable staged_cv;
    const size_t total_items = real_blocks * per_items;
    auto worker = [&]() {
        for (;;)
        {
            const size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= total_items)
                return;
            const size_t b = i / per_items, r = i - b * per_items, t = r / pieces;
            const size_t v0 = blocks[b].v0, v1 = blocks[b].v1;
            const size_t c0 = v0 + (r - t * pieces) * piece, c1 = std::min(v1, c0 + piece);
            if (c0 < c1)
            {
                if (N == n_grid)
                    stage_copy(dst_all + t * N + c0, data + t * N + c0, c1 - c0);
                else
                    for (size_t v = c0; v < c1; v++)
                        dst_all[t * N + v] = data[t * n_grid + index[v]];
            }
            if (unfinished[b].fetch_sub(1, std::memory_order_acq_rel) == 1)
            {
                std::lock_guard<std::mutex> lock(staged_mu);
                staged_cv.notify_all();
            }
        }
    };
    size_t threads = host_threads();
    if (total_items < threads)
        threads = total_items;
    if ((size_t)data_size * N < ((size_t)1 << 16))
        threads = 0; /* tiny: the calling thread does it */
    PoolJob *job = pool_launch(threads, worker);
    if (threads == 0)
        worker();
    struct Join
    {
        PoolJob *job;
        std::atomic<size_t> &next;
        size_t total;
        ~Join()
        {
            next.store(total, std::memory_order_relaxed); /* on an exception: stop the workers before unwinding */
            pool_wait(job);
        }
    } join = { job, next, total_items };
    for (size_t b = 0; b < real_blocks; b++)
    {
        {
            std::unique_lock<std::mutex> lock(staged_mu);
            staged_cv.wait(lock, [&]() { return unfinished[b].load(std::memory_order_acquire) == 0; });
        }
        if (upload && queue_block(blocks[b], dst_all, N) != FABBER_CUDA_OK)
            throw FabberInternalError(std::string("copying data to the GPU: ") + fabber_cuda_last_error());
    }
    keep_spec();
    m_voxel_data[key] = std::move(vd);
}

const VoxelData &FabberRunData::GetVoxelData(const std::string &key_in)
{
    /* indirection: an option may name the data key (rundata.cc:802-823) */
    std::string key = key_in;
    std::set<std::string> seen;
    while (m_voxel_data.count(key) == 0)
    {
        if (m_params.count(key) == 0 || m_params[key] == "" || seen.count(key))
        {
            /* end of the chain: the name of something a file-based front end can load */
            if (LoadVoxelData(key) && m_voxel_data.count(key))
                break;
            throw DataNotFound(key_in);
        }
        seen.insert(key);
        m_used_params.insert(key);
        key = m_params[key];
    }
    if (!m_device_only_access)
        m_voxel_data[key]->ensure_host(); /* every reader but Vb's device path wants the host copy */
    return *m_voxel_data[key];
}
bool FabberRunData::LoadVoxelData(const std::string &) { return false; }
void FabberRunData::SaveVoxelData(const std::string &, VoxelDataType) {}

const VoxelData &FabberRunData::GetMainVoxelData()
{
    if (m_voxel_data.count("@maindata"))
        return *m_voxel_data["@maindata"];
    try
    {
        return GetVoxelData("data");
    }
    catch (DataNotFound &e)
    {
        try
        {
            GetVoxelData("data1");
        }
        catch (DataNotFound &)
        {
            throw e;
        }
        return GetMainVoxelDataMultiple();
    }
}
VoxelData &FabberRunData::MutableMainVoxelData()
{
    /* Vb works on the device copy: do not pull a directly-uploaded series back to the host for it */
    struct Flag
    {
        bool &f;
        explicit Flag(bool &b)
            : f(b)
        {
            f = true;
        }
        ~Flag() { f = false; }
    } flag(m_device_only_access);
    return const_cast<VoxelData &>(GetMainVoxelData());
}

/* rundata.cc:821-905 */
const VoxelData &FabberRunData::GetMainVoxelDataMultiple()
{
    std::vector<const VoxelData *> sets;
    for (int n = 1;; n++)
    {
        try
        {
            sets.push_back(&GetVoxelData("data" + stringify(n)));
        }
        catch (DataNotFound &)
        {
            break;
        }
    }
    const std::string order = GetStringDefault("data-order", "interleave");
    const int n_sets = (int)sets.size();
    if (n_sets < 1)
        throw DataNotFound("data");
    if (order == "singlefile" && n_sets > 1)
        throw InvalidOptionValue("data-order", "singlefile", "More than one file specified");
    const size_t N = sets[0]->cols;
    int total = 0;
    for (int j = 0; j < n_sets; j++)
    {
        if (sets[j]->cols != N)
            throw FabberRunDataError("data" + stringify(j + 1) + " has a different number of voxels");
        total += sets[j]->rows;
    }
    std::unique_ptr<VoxelData> vd(new VoxelData());
    vd->alloc(total, N);
    if (order == "interleave")
    {
        m_log << "FabberRunData::Combining data into one big matrix by interleaving..." << std::endl;
        const int n_times = sets[0]->rows;
        for (int j = 0; j < n_sets; j++)
            if (sets[j]->rows != n_times)
                throw InvalidOptionValue("data-order", "interleave", "Data sets must all have the same number of time points");
        for (int i = 0; i < n_times; i++)
            for (int j = 0; j < n_sets; j++)
                memcpy(vd->f + (size_t)(n_sets * i + j) * N, sets[j]->f + (size_t)i * N, N * sizeof(float));
    }
    else if (order == "concatenate" || order == "singlefile")
    {
        if (order == "concatenate")
            m_log << "FabberRunData::Combining data into one big matrix by concatenating..." << std::endl;
        size_t row = 0;
        for (int j = 0; j < n_sets; j++)
        {
            memcpy(vd->f + row * N, sets[j]->f, sets[j]->bytes());
            row += sets[j]->rows;
        }
    }
    else
        throw InvalidOptionValue("data-order", order, "Value not recognized");
    m_log << "FabberRunData::Done loading data, size = " << total << " timepoints by " << N << " voxels" << std::endl;
    m_voxel_data["@maindata"] = std::move(vd);
    return *m_voxel_data["@maindata"];
}
int FabberRunData::GetVoxelDataSize(const std::string &key) { return GetVoxelData(key).rows; }
VoxelData &FabberRunData::NewVoxelData(const std::string &key, int rows)
{
    m_voxel_data.erase(key);
    std::unique_ptr<VoxelData> vd(new VoxelData());
    vd->alloc(rows, m_voxel_index.size());
    m_voxel_data[key] = std::move(vd);
    return *m_voxel_data[key];
}
VoxelData &FabberRunData::MutableVoxelData(const std::string &key)
{
    return const_cast<VoxelData &>(GetVoxelData(key));
}
void FabberRunData::ClearVoxelData(const std::string &key)
{
    if (m_spec && m_voxel_data.count(key) && m_voxel_data[key].get() == m_spec->data)
        DiscardSpeculative();
    m_voxel_data.erase(key);
}

/* rundata_array.cc:68-98: T x Nvox -> float[t][z][y][x], zeros outside the mask */
void FabberRunData::GetVoxelDataArray(const std::string &key, float *data)
{
    const VoxelData &vd = GetVoxelData(key);
    const size_t n_grid = (size_t)m_extent[0] * m_extent[1] * m_extent[2];
    const size_t N = m_voxel_index.size();
    const std::vector<int> &index = m_voxel_index;
    if (N == n_grid)
    {
        /* full mask: the stored layout IS the caller's layout */
        parallel_for((size_t)vd.rows * N, [&](size_t b, size_t e) { stage_copy(data + b, vd.f + b, e - b); },
            (size_t)1 << 17);
        return;
    }
    parallel_for((size_t)vd.rows * n_grid, [&](size_t b, size_t e) { memset(data + b, 0, (e - b) * sizeof(float)); },
        (size_t)1 << 20);
    parallel_for((size_t)vd.rows * N, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; i++)
        {
            const size_t t = i / N, v = i - t * N;
            data[t * n_grid + index[v]] = vd.f[i];
        }
    });
}

void FabberRunData::GetOptions(std::vector<OptionSpec> &opts)
{
    /* the general options, same names / types / order / defaults as rundata.cc:139-201 (what fabber_get_options
     * and --help list; the descriptions are this library's own wording) */
    static const OptionSpec O[] = {
        { "help", OPT_BOOL, "Print usage; with --method or --model, the usage of that method / model", true, "" },
        { "listmethods", OPT_BOOL, "List the known inference methods", true, "" },
        { "listmodels", OPT_BOOL, "List the known forward models", true, "" },
        { "listparams", OPT_BOOL, "List the model's parameters (needs the model's configuration options)", true, "" },
        { "descparams", OPT_BOOL, "Describe the model's parameters: name, description, units (needs the model's options)", true, "" },
        { "listoutputs", OPT_BOOL, "List the model's additional outputs (needs the model's configuration options)", true, "" },
        { "evaluate", OPT_STR, "Evaluate the model: name of the output wanted, or blank for the prediction; needs the model's options, --evaluate-params and --evaluate-nt", true, "" },
        { "evaluate-params", OPT_MATRIX, "Parameter values for the evaluation", true, "" },
        { "evaluate-nt", OPT_INT, "Number of time points for the evaluation", true, "" },
        { "simple-output", OPT_BOOL, "Print only progress lines (percentages) on standard output", true, "" },
        { "output", OPT_STR, "Directory for the output files (including the logfile)", false, "" },
        { "overwrite", OPT_BOOL, "Overwrite an existing output directory instead of appending '+' to its name", true, "" },
        { "link-to-latest", OPT_BOOL, "Try to create a link <output>_latest to the most recent output directory", true, "" },
        { "method", OPT_STR, "Use this inference method", false, "" },
        { "model", OPT_STR, "Use this forward model", false, "" },
        { "loadmodels", OPT_FILE, "Load forward models from this shared library (device model plug-ins, include/fabber_model_plugin.h)", true, "" },
        { "data", OPT_TIMESERIES, "The single input data file", false, "" },
        { "data<n>", OPT_TIMESERIES, "Several input data files, n = 1, 2, 3...", true, "" },
        { "data-order", OPT_STR, "How several data files are combined: concatenate (one after the other) or interleave (first sample of each, then the second, ...)", true, "interleave" },
        { "mask", OPT_IMAGE, "Mask: inference will only be performed where mask value > 0", true, "" },
        { "mt<n>", OPT_INT, "List of masked time points, indexed from 1: ignored in the parameter updates", true, "" },
        { "suppdata", OPT_TIMESERIES, "'Supplemental' timeseries data, required for some models", true, "" },
        { "dump-param-names", OPT_BOOL, "Write paramnames.txt with the names of the model's parameters", true, "" },
        { "save-model-fit", OPT_BOOL, "Output the model prediction as a 4d volume", true, "" },
        { "save-residuals", OPT_BOOL, "Output the difference between the data and the model prediction", true, "" },
        { "save-model-extras", OPT_BOOL, "Output the model's additional timeseries outputs", true, "" },
        { "save-mvn", OPT_BOOL, "Output the final MVN distributions", true, "" },
        { "save-mean", OPT_BOOL, "Output the parameter means", true, "" },
        { "save-std", OPT_BOOL, "Output the parameter standard deviations", true, "" },
        { "save-var", OPT_BOOL, "Output the parameter variances", true, "" },
        { "save-zstat", OPT_BOOL, "Output the parameter Zstats", true, "" },
        { "save-noise-mean", OPT_BOOL, "Output the noise means (the inferred distribution is the precision of a Gaussian noise source)", true, "" },
        { "save-noise-std", OPT_BOOL, "Output the noise standard deviations", true, "" },
        { "save-free-energy", OPT_BOOL, "Output the free energy, if calculated", true, "" },
        { "optfile", OPT_BOOL, "File with further options, one per line, written as on the command line", true, "" },
        { "debug", OPT_BOOL, "Very verbose logging (the reference's debug output is per voxel; nothing extra is logged here)", true, "" },
    };
    for (size_t i = 0; i < sizeof(O) / sizeof(O[0]); i++)
        opts.push_back(O[i]);
}

/* ---- command line and option files (rundata.cc:324-453) ------------------------------------------ */
static std::string trim(const std::string &s)
{
    const char *ws = " \t\r\n";
    const size_t b = s.find_first_not_of(ws);
    if (b == std::string::npos)
        return "";
    return s.substr(b, s.find_last_not_of(ws) - b + 1);
}

void FabberRunData::AddKeyEqualsValue(const std::string &exp, bool trim_comments)
{
    const size_t eq = exp.find("=");
    const std::string key = trim(exp.substr(0, eq));
    if (eq != std::string::npos)
    {
        size_t end = std::string::npos;
        if (trim_comments)
            end = exp.find("#");
        const std::string value = trim(exp.substr(eq + 1, end == std::string::npos ? end : end - (eq + 1)));
        if (m_params.count(key) > 0)
            throw InvalidOptionValue(key, value, "Already has a value: " + m_params[key]);
        if (key == "loadmodels") /* rundata.cc:440-443: acted on at once, not stored */
            FwdModel::LoadFromDynamicLibrary(value, &m_log);
        else
            m_params[key] = value;
    }
    else
        m_params[exp] = "";
    m_version++;
}

void FabberRunData::ParseParamFile(const std::string &filename)
{
    std::ifstream is(filename.c_str());
    if (!is.good())
        throw FabberRunDataError("Couldn't read input options file:" + filename);
    std::string input;
    while (std::getline(is, input))
    {
        input = trim(input);
        if (input.size() > 0 && input[0] != '#')
            AddKeyEqualsValue(input, true);
    }
}

void FabberRunData::ParseOldStyleParamFile(const std::string &filename)
{
    std::ifstream is(filename.c_str());
    if (!is.good())
        throw FabberRunDataError("Couldn't read input file: -@ " + filename);
    std::string param;
    char c;
    while (is.good())
    {
        if (!is.get(c))
            c = '\n'; /* end of file terminates the last word */
        if (!isspace((unsigned char)c))
            param += c;
        else if (param == "")
        {
        }
        else if (param.compare(0, 2, "--") == 0)
        {
            AddKeyEqualsValue(param.substr(2));
            param = "";
        }
        else if (param[0] == '#')
        {
            param = "";
            while (is.good() && c != '\n')
                is.get(c);
        }
        else if (param.compare(0, 2, "-@") == 0)
            throw FabberRunDataError("Can only use -@ on the command line");
        else
            throw FabberRunDataError("Invalid data '" + param + "' found in file '" + filename + "'");
    }
}

void FabberRunData::Parse(int argc, char **argv)
{
    m_params[""] = argv[0];
    for (int a = 1; a < argc; a++)
    {
        const std::string arg = argv[a];
        if (arg == "-f")
        {
            if (++a < argc)
                ParseParamFile(argv[a]);
            else
                throw InvalidOptionValue("-f", "", "No filename specified");
        }
        else if (arg.compare(0, 2, "--") == 0)
            AddKeyEqualsValue(arg.substr(2));
        else if (arg == "-@")
        {
            if (++a < argc)
                ParseOldStyleParamFile(argv[a]);
            else
                throw InvalidOptionValue("-@", "", "No filename specified");
        }
        else
            throw FabberRunDataError("Option '" + arg + "' doesn't begin with --");
    }
    if (HaveKey("optfile"))
        ParseOldStyleParamFile(GetString("optfile"));
}

void FabberRunData::LogParams()
{
    for (std::map<std::string, std::string>::const_iterator i = m_params.begin(); i != m_params.end(); ++i)
        m_log << "FabberRunData::Parameter " << i->first << "=" << i->second << std::endl;
}

void FabberRunData::WarnOnce(const std::string &text)
{
    if (++m_warncount[text] == 1)
        m_log << "WARNING ONCE: " << text << std::endl;
}

void FabberRunData::ReissueWarnings()
{
    if (m_warncount.empty())
        return;
    m_log << "\nSummary of warnings (" << m_warncount.size() << " distinct warnings)\n";
    for (std::map<std::string, int>::const_iterator it = m_warncount.begin(); it != m_warncount.end(); ++it)
        m_log << "Issued " << (it->second == 1 ? std::string("once: ") : stringify(it->second) + " times: ") << it->first
              << std::endl;
}

void FabberRunData::CheckAllOptionsUsed()
{
    for (std::map<std::string, std::string>::const_iterator i = m_params.begin(); i != m_params.end(); ++i)
        if (i->first != "" && m_used_params.count(i->first) == 0)
            WarnOnce("Unused option specified: " + i->first);
}

static bool is_dir(const std::string &path)
{
    struct stat s;
    return stat(path.c_str(), &s) == 0 && S_ISDIR(s.st_mode);
}

std::string FabberRunData::GetOutputDir()
{
    const bool link_to_latest = GetBool("link-to-latest");
    if (m_outdir != "")
        return m_outdir;
    const std::string basename = GetStringDefault("output", "");
    if (basename == "")
    {
        m_outdir = ".";
        return m_outdir;
    }
    const bool overwrite = GetBool("overwrite");
    m_outdir = basename;
    for (int count = 0;; count++)
    {
        if (count >= 50)
            throw FabberInternalError("Cannot create output directory (bad path, or too many + signs?): " + m_outdir);
        errno = 0;
        if (mkdir(m_outdir.c_str(), 0777) == 0)
            break;
        if (overwrite)
        {
            if (errno == EEXIST && is_dir(m_outdir))
                break;
            throw FabberInternalError("Unexpected problem creating output directory in overwrite mode: " + m_outdir);
        }
        m_outdir += "+";
    }
    if (link_to_latest)
    {
        /* "<output>_latest" -> the directory actually used; failure does not matter (rundata.cc:727-733) */
        const std::string link = basename + "_latest";
        unlink(link.c_str());
        if (symlink(m_outdir.c_str(), link.c_str()) != 0)
            m_log << "FabberRunData::link-to-latest failed" << std::endl;
    }
    return m_outdir;
}

/* rundata.cc:248-311 */
void FabberRunData::Run(void (*progress_cb)(int, int))
{
    struct InRun
    {
        bool &f;
        explicit InRun(bool &b)
            : f(b)
        {
            f = true;
        }
        ~InRun() { f = false; }
    } in_run(m_in_run);
    m_progress = progress_cb;
    time_t start;
    time(&start);
    m_log << "FabberRunData::Start time: " << ctime(&start);
    LogParams();
    std::unique_ptr<FwdModel> fwd_model(FwdModel::NewFromName(GetString("model")));
    fwd_model->Initialize(*this);
    std::vector<Parameter> params;
    fwd_model->GetParameters(*this, params);
    m_log << "FabberRunData::Forward Model version " << fwd_model->ModelVersion() << std::endl;
    if (GetBool("dump-param-names")) /* rundata.cc:277-286 */
    {
        std::ofstream param_file((GetStringDefault("output", ".") + "/paramnames.txt").c_str());
        for (size_t i = 0; i < params.size(); i++)
            param_file << params[i].name << std::endl;
    }

    const std::string method = GetString("method");
    if (method != "vb" && method != "spatialvb" && method != "nlls") /* setup.cc:28-33 */
        throw InvalidOptionValue("method", method, "Unrecognized inference method (vb, spatialvb, nlls)");
    /* a run that fabber_set_data started speculatively is adopted if nothing was set since */
    std::unique_ptr<SpeculativeRun> spec = std::move(m_spec);
    if (spec && (spec->version != m_version || m_voxel_data.count("data") == 0 || m_voxel_data["data"].get() != spec->data))
        spec.reset();
    if (spec)
    {
        m_log << "FabberRunData::Adopting the run started while the data was being set" << std::endl;
        m_log << spec->vb->Description() << std::endl;
        spec->vb->Finish(*this);
        spec->vb->SaveResults(*this);
    }
    else
    {
        std::unique_ptr<InferenceTechnique> infer(InferenceTechnique::NewFromName(method)); /* fabber_core.cc:258 */
        infer->Initialize(fwd_model.get(), *this);
        infer->DoCalculations(*this);
        infer->SaveResults(*this);
    }
    time_t end;
    time(&end);
    m_log << "FabberRunData::All done." << std::endl;
    CheckAllOptionsUsed();
    m_log << "FabberRunData::End time: " << ctime(&end);
    m_log << "FabberRunData::Duration: " << (long)difftime(end, start) << " seconds." << std::endl;
    m_progress = nullptr;
}

} // namespace fabber_b200
