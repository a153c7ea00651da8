/*
 * nifti_io.cc - see nifti_io.h. NIfTI-1 header layout: nifti1.h (public domain, NIH).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include <sys/stat.h>
#include <zlib.h>

#include "fabber_host.h"
#include "nifti_io.h"

namespace fabber_b200
{
namespace
{
/* byte offsets in the 348-byte header */
enum
{
    OFF_SIZEOF_HDR = 0,
    OFF_DIM = 40,
    OFF_INTENT_P1 = 56,
    OFF_INTENT_CODE = 68,
    OFF_DATATYPE = 70,
    OFF_BITPIX = 72,
    OFF_PIXDIM = 76,
    OFF_VOX_OFFSET = 108,
    OFF_SCL_SLOPE = 112,
    OFF_SCL_INTER = 116,
    OFF_CAL_MAX = 124,
    OFF_CAL_MIN = 128,
    OFF_QFORM_CODE = 252,
    OFF_SFORM_CODE = 254,
    OFF_QUATERN_B = 256,
    OFF_SROW_X = 280,
    OFF_SROW_Y = 296,
    OFF_SROW_Z = 312,
    OFF_MAGIC = 344
};

template <class T> T get(const unsigned char *raw, int off)
{
    T v;
    memcpy(&v, raw + off, sizeof(T));
    return v;
}
template <class T> void put(unsigned char *raw, int off, T v) { memcpy(raw + off, &v, sizeof(T)); }

void swap_bytes(unsigned char *p, size_t size, size_t count)
{
    for (size_t i = 0; i < count; i++, p += size)
        std::reverse(p, p + size);
}

/* every multi-byte field of the header, by size */
void swap_header(unsigned char *raw)
{
    swap_bytes(raw + 0, 4, 1);    /* sizeof_hdr */
    swap_bytes(raw + 32, 4, 1);   /* extents */
    swap_bytes(raw + 36, 2, 1);   /* session_error */
    swap_bytes(raw + 40, 2, 8);   /* dim */
    swap_bytes(raw + 56, 4, 3);   /* intent_p1..3 */
    swap_bytes(raw + 68, 2, 4);   /* intent_code, datatype, bitpix, slice_start */
    swap_bytes(raw + 76, 4, 8);   /* pixdim */
    swap_bytes(raw + 108, 4, 3);  /* vox_offset, scl_slope, scl_inter */
    swap_bytes(raw + 120, 2, 1);  /* slice_end */
    swap_bytes(raw + 124, 4, 4);  /* cal_max, cal_min, slice_duration, toffset */
    swap_bytes(raw + 140, 4, 2);  /* glmax, glmin */
    swap_bytes(raw + 252, 2, 2);  /* qform_code, sform_code */
    swap_bytes(raw + 256, 4, 18); /* quatern_b..d, qoffset_x..z, srow_x..z */
}

bool file_exists(const std::string &p)
{
    struct stat s;
    return stat(p.c_str(), &s) == 0 && S_ISREG(s.st_mode);
}

bool ends_with(const std::string &s, const std::string &e)
{
    return s.size() >= e.size() && s.compare(s.size() - e.size(), e.size(), e) == 0;
}

void read_all(gzFile f, void *dst, size_t bytes, const std::string &name)
{
    char *p = (char *)dst;
    while (bytes > 0)
    {
        const unsigned chunk = (unsigned)std::min<size_t>(bytes, (size_t)1 << 30);
        const int got = gzread(f, p, chunk);
        if (got <= 0)
            throw FabberRunDataError("Error loading file: " + name + " (truncated)");
        p += got;
        bytes -= (size_t)got;
    }
}

/* sign of the determinant of the voxel -> world matrix NEWIMAGE consults: sform if set, else qform, else
 * "no orientation information": treated as radiological */
bool is_neurological(const unsigned char *raw)
{
    const int sform = get<int16_t>(raw, OFF_SFORM_CODE), qform = get<int16_t>(raw, OFF_QFORM_CODE);
    double m[3][3];
    if (sform > 0)
    {
        for (int c = 0; c < 3; c++)
        {
            m[0][c] = get<float>(raw, OFF_SROW_X + 4 * c);
            m[1][c] = get<float>(raw, OFF_SROW_Y + 4 * c);
            m[2][c] = get<float>(raw, OFF_SROW_Z + 4 * c);
        }
    }
    else if (qform > 0)
    {
        const double b = get<float>(raw, OFF_QUATERN_B), c = get<float>(raw, OFF_QUATERN_B + 4),
                     d = get<float>(raw, OFF_QUATERN_B + 8);
        double a = 1.0 - (b * b + c * c + d * d);
        a = a > 0 ? sqrt(a) : 0.0;
        const double qfac = get<float>(raw, OFF_PIXDIM) < 0 ? -1.0 : 1.0;
        const double dx = get<float>(raw, OFF_PIXDIM + 4), dy = get<float>(raw, OFF_PIXDIM + 8),
                     dz = get<float>(raw, OFF_PIXDIM + 12) * qfac;
        m[0][0] = (a * a + b * b - c * c - d * d) * dx;
        m[0][1] = 2 * (b * c - a * d) * dy;
        m[0][2] = 2 * (b * d + a * c) * dz;
        m[1][0] = 2 * (b * c + a * d) * dx;
        m[1][1] = (a * a + c * c - b * b - d * d) * dy;
        m[1][2] = 2 * (c * d - a * b) * dz;
        m[2][0] = 2 * (b * d - a * c) * dx;
        m[2][1] = 2 * (c * d + a * b) * dy;
        m[2][2] = (a * a + d * d - c * c - b * b) * dz;
    }
    else
        return false;
    const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0])
        + m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    return det > 0;
}

void mirror_x(float *data, int nx, size_t rows)
{
    parallel_for(rows, [&](size_t b, size_t e) {
        for (size_t r = b; r < e; r++)
            std::reverse(data + r * nx, data + (r + 1) * nx);
    });
}

template <class T> void convert(const unsigned char *src, float *dst, size_t n, bool swap, double slope, double inter)
{
    parallel_for(n, [&](size_t b, size_t e) {
        for (size_t i = b; i < e; i++)
        {
            unsigned char tmp[sizeof(T)];
            memcpy(tmp, src + i * sizeof(T), sizeof(T));
            if (swap)
                std::reverse(tmp, tmp + sizeof(T));
            T v;
            memcpy(&v, tmp, sizeof(T));
            dst[i] = (float)((double)v * slope + inter);
        }
    });
}
} // namespace

NiftiHeader::NiftiHeader()
    : nx(0)
    , ny(0)
    , nz(0)
    , nt(0)
    , dx(1)
    , dy(1)
    , dz(1)
    , dt(1)
    , intent_code(0)
    , flip_x(false)
{
    memset(raw, 0, sizeof(raw));
    intent_p[0] = intent_p[1] = intent_p[2] = 0;
}

std::string nifti_find(const std::string &name)
{
    if (name == "")
        return "";
    static const char *ext[] = { "", ".nii.gz", ".nii", ".hdr", ".hdr.gz" };
    for (size_t i = 0; i < sizeof(ext) / sizeof(ext[0]); i++)
    {
        std::string p = name + ext[i];
        if (!file_exists(p))
            continue;
        if (i == 0 && (ends_with(p, ".img") || ends_with(p, ".img.gz")))
        {
            /* the data half of a pair: the header is next to it */
            std::string base = p.substr(0, p.rfind(".img"));
            if (file_exists(base + ".hdr"))
                return base + ".hdr";
            if (file_exists(base + ".hdr.gz"))
                return base + ".hdr.gz";
            continue;
        }
        return p;
    }
    return "";
}

void nifti_read(const std::string &name, NiftiHeader &hdr, std::vector<float> &data)
{
    const std::string path = nifti_find(name);
    if (path == "")
        throw DataNotFound(name + " (File is invalid or does not exist)");
    gzFile f = gzopen(path.c_str(), "rb");
    if (!f)
        throw DataNotFound(name + " (Error loading file)");
    try
    {
        gzbuffer(f, 1 << 20);
        read_all(f, hdr.raw, 348, path);
        bool swap = false;
        if (get<int32_t>(hdr.raw, OFF_SIZEOF_HDR) != 348)
        {
            swap_header(hdr.raw);
            swap = true;
            if (get<int32_t>(hdr.raw, OFF_SIZEOF_HDR) != 348)
                throw FabberRunDataError("Error loading file: " + path + " is not a NIfTI-1 / Analyze image");
        }
        const char *magic = (const char *)hdr.raw + OFF_MAGIC;
        const bool single = memcmp(magic, "n+1", 3) == 0;
        const int ndim = get<int16_t>(hdr.raw, OFF_DIM);
        if (ndim < 1 || ndim > 7)
            throw FabberRunDataError("Error loading file: " + path + " has a bad dim[0]");
        int dim[8];
        for (int i = 0; i < 8; i++)
            dim[i] = get<int16_t>(hdr.raw, OFF_DIM + 2 * i);
        hdr.nx = std::max(dim[1], 1);
        hdr.ny = ndim >= 2 ? std::max(dim[2], 1) : 1;
        hdr.nz = ndim >= 3 ? std::max(dim[3], 1) : 1;
        hdr.nt = ndim >= 4 ? std::max(dim[4], 1) : 1;
        for (int i = 5; i <= ndim; i++) /* higher dimensions fold into the 4th, like NEWIMAGE */
            hdr.nt *= std::max(dim[i], 1);
        hdr.dx = get<float>(hdr.raw, OFF_PIXDIM + 4);
        hdr.dy = get<float>(hdr.raw, OFF_PIXDIM + 8);
        hdr.dz = get<float>(hdr.raw, OFF_PIXDIM + 12);
        hdr.dt = get<float>(hdr.raw, OFF_PIXDIM + 16);
        hdr.intent_code = get<int16_t>(hdr.raw, OFF_INTENT_CODE);
        for (int i = 0; i < 3; i++)
            hdr.intent_p[i] = get<float>(hdr.raw, OFF_INTENT_P1 + 4 * i);
        hdr.flip_x = is_neurological(hdr.raw);

        const int datatype = get<int16_t>(hdr.raw, OFF_DATATYPE);
        size_t elem = 0;
        switch (datatype)
        {
        case 2:
        case 256:
            elem = 1;
            break;
        case 4:
        case 512:
            elem = 2;
            break;
        case 8:
        case 16:
        case 768:
            elem = 4;
            break;
        case 64:
        case 1024:
        case 1280:
            elem = 8;
            break;
        default:
            throw FabberRunDataError("Error loading file: " + path + " has unsupported datatype " + stringify(datatype));
        }
        const size_t n = (size_t)hdr.nx * hdr.ny * hdr.nz * hdr.nt;
        std::vector<unsigned char> bytes(n * elem);
        if (single)
        {
            const long off = (long)get<float>(hdr.raw, OFF_VOX_OFFSET);
            if (off < 348)
                throw FabberRunDataError("Error loading file: " + path + " has a bad vox_offset");
            std::vector<unsigned char> skip((size_t)off - 348);
            if (!skip.empty())
                read_all(f, skip.data(), skip.size(), path);
            read_all(f, bytes.data(), bytes.size(), path);
        }
        else
        {
            std::string base = path.substr(0, path.rfind(".hdr"));
            std::string img = file_exists(base + ".img") ? base + ".img" : base + ".img.gz";
            gzFile g = gzopen(img.c_str(), "rb");
            if (!g)
                throw DataNotFound(name + " (no .img next to the header)");
            try
            {
                const long off = (long)get<float>(hdr.raw, OFF_VOX_OFFSET);
                std::vector<unsigned char> skip(off > 0 ? (size_t)off : 0);
                if (!skip.empty())
                    read_all(g, skip.data(), skip.size(), img);
                read_all(g, bytes.data(), bytes.size(), img);
            }
            catch (...)
            {
                gzclose(g);
                throw;
            }
            gzclose(g);
        }
        double slope = get<float>(hdr.raw, OFF_SCL_SLOPE), inter = get<float>(hdr.raw, OFF_SCL_INTER);
        if (slope == 0 || !std::isfinite(slope) || !std::isfinite(inter))
        {
            slope = 1;
            inter = 0;
        }
        data.resize(n);
        switch (datatype)
        {
        case 2:
            convert<uint8_t>(bytes.data(), data.data(), n, false, slope, inter);
            break;
        case 256:
            convert<int8_t>(bytes.data(), data.data(), n, false, slope, inter);
            break;
        case 4:
            convert<int16_t>(bytes.data(), data.data(), n, swap, slope, inter);
            break;
        case 512:
            convert<uint16_t>(bytes.data(), data.data(), n, swap, slope, inter);
            break;
        case 8:
            convert<int32_t>(bytes.data(), data.data(), n, swap, slope, inter);
            break;
        case 768:
            convert<uint32_t>(bytes.data(), data.data(), n, swap, slope, inter);
            break;
        case 16:
            convert<float>(bytes.data(), data.data(), n, swap, slope, inter);
            break;
        case 64:
            convert<double>(bytes.data(), data.data(), n, swap, slope, inter);
            break;
        case 1024:
            convert<int64_t>(bytes.data(), data.data(), n, swap, slope, inter);
            break;
        case 1280:
            convert<uint64_t>(bytes.data(), data.data(), n, swap, slope, inter);
            break;
        }
        if (hdr.flip_x)
            mirror_x(data.data(), hdr.nx, n / hdr.nx);
    }
    catch (...)
    {
        gzclose(f);
        throw;
    }
    gzclose(f);
}

std::string nifti_write(const std::string &name, const NiftiHeader &like, int nt, int intent_code, const float *data)
{
    const char *type = getenv("FSLOUTPUTTYPE");
    const bool gz = !type || std::string(type) != "NIFTI";
    std::string path = name;
    if (!ends_with(path, ".nii") && !ends_with(path, ".nii.gz"))
        path += gz ? ".nii.gz" : ".nii";
    const bool gz_out = ends_with(path, ".gz");

    unsigned char raw[352];
    memset(raw, 0, sizeof(raw));
    memcpy(raw, like.raw, 348);
    put<int32_t>(raw, OFF_SIZEOF_HDR, 348);
    int16_t dim[8] = { (int16_t)(nt > 1 ? 4 : 3), (int16_t)like.nx, (int16_t)like.ny, (int16_t)like.nz, (int16_t)nt, 1, 1, 1 };
    for (int i = 0; i < 8; i++)
        put<int16_t>(raw, OFF_DIM + 2 * i, dim[i]);
    if (get<float>(raw, OFF_PIXDIM + 4) == 0) /* a blank template: unit voxels */
        for (int i = 1; i <= 4; i++)
            put<float>(raw, OFF_PIXDIM + 4 * i, 1.0f);
    put<int16_t>(raw, OFF_INTENT_CODE, (int16_t)intent_code);
    for (int i = 0; i < 3; i++)
        put<float>(raw, OFF_INTENT_P1 + 4 * i, 0.0f);
    put<int16_t>(raw, OFF_DATATYPE, 16);
    put<int16_t>(raw, OFF_BITPIX, 32);
    put<float>(raw, OFF_VOX_OFFSET, 352.0f);
    put<float>(raw, OFF_SCL_SLOPE, 1.0f);
    put<float>(raw, OFF_SCL_INTER, 0.0f);
    const size_t n = (size_t)like.nx * like.ny * like.nz * nt;
    float mn = 0, mx = 0;
    if (n > 0)
    {
        mn = mx = data[0];
        for (size_t i = 1; i < n; i++)
        {
            mn = std::min(mn, data[i]);
            mx = std::max(mx, data[i]);
        }
    }
    put<float>(raw, OFF_CAL_MAX, mx);
    put<float>(raw, OFF_CAL_MIN, mn);
    memcpy(raw + OFF_MAGIC, "n+1\0", 4);

    std::vector<float> mirrored;
    const float *out = data;
    if (like.flip_x && n > 0)
    {
        mirrored.assign(data, data + n);
        mirror_x(mirrored.data(), like.nx, n / like.nx);
        out = mirrored.data();
    }
    bool ok = true;
    if (gz_out)
    {
        gzFile f = gzopen(path.c_str(), "wb1"); /* speed over ratio: these are intermediate research outputs */
        if (!f)
            throw FabberRunDataError("Could not open " + path + " for writing");
        gzbuffer(f, 1 << 20);
        ok = gzwrite(f, raw, 352) == 352;
        const char *p = (const char *)out;
        size_t left = n * sizeof(float);
        while (ok && left > 0)
        {
            const unsigned chunk = (unsigned)std::min<size_t>(left, (size_t)1 << 30);
            ok = gzwrite(f, p, chunk) == (int)chunk;
            p += chunk;
            left -= chunk;
        }
        ok = (gzclose(f) == Z_OK) && ok;
    }
    else
    {
        FILE *f = fopen(path.c_str(), "wb");
        if (!f)
            throw FabberRunDataError("Could not open " + path + " for writing");
        ok = fwrite(raw, 1, 352, f) == 352 && fwrite(out, sizeof(float), n, f) == n;
        ok = (fclose(f) == 0) && ok;
    }
    if (!ok)
        throw FabberRunDataError("Error writing " + path);
    return path;
}
} // namespace fabber_b200
