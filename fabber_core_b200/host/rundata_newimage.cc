/*
 * rundata_newimage.cc - FabberRunDataNewimage: the file-based front end of the command line tool.
 * Follows rundata_newimage.cc:55-225 of the reference; NIfTI I/O is nifti_io.cc (zlib) instead of NEWIMAGE.
 */
#include <cstring>

#include "fabber_host.h"
#include "nifti_io.h"

namespace fabber_b200
{
FabberRunDataNewimage::FabberRunDataNewimage(bool)
    : m_have_mask(false)
{
}
FabberRunDataNewimage::~FabberRunDataNewimage() {}

static void dump_info(const NiftiHeader &h, std::ostream &out)
{
    out << "FabberRunDataNewimage::Dimensions: x=" << h.nx << ", y=" << h.ny << ", z=" << h.nz << ", vols=" << h.nt
        << std::endl;
    out << "FabberRunDataNewimage::Voxel size: x=" << h.dx << "mm, y=" << h.dy << "mm, z=" << h.dz << "mm, TR=" << h.dt
        << " sec\n";
    out << "FabberRunDataNewimage::Intents: " << h.intent_code << ", " << h.intent_p[0] << ", " << h.intent_p[1] << ", "
        << h.intent_p[2] << std::endl;
}

void FabberRunDataNewimage::SetExtentFromData()
{
    const std::string mask_fname = GetStringDefault("mask", "");
    m_have_mask = mask_fname != "";
    m_like.reset(new NiftiHeader());
    std::vector<float> vol;
    if (m_have_mask)
    {
        Log() << "FabberRunDataNewimage::Loading mask data from '" + mask_fname << "'" << std::endl;
        if (nifti_find(mask_fname) == "")
            throw DataNotFound(mask_fname + " (File is invalid or does not exist)");
        nifti_read(mask_fname, *m_like, vol);
        dump_info(*m_like, Log());
        /* volume::binarise(1e-16, max + 1, exclusive): 1 strictly inside the interval, else 0. Only the
         * first volume of a 4-D mask counts (read_volume). */
        const size_t n = (size_t)m_like->nx * m_like->ny * m_like->nz;
        float mx = vol.empty() ? 0.f : vol[0];
        for (size_t i = 1; i < n; i++)
            mx = std::max(mx, vol[i]);
        std::vector<int> mask(n);
        for (size_t i = 0; i < n; i++)
            mask[i] = (vol[i] > 1e-16f && vol[i] < mx + 1) ? 1 : 0;
        m_like->nt = 1;
        SetExtent(m_like->nx, m_like->ny, m_like->nz, mask.data());
    }
    else
    {
        Log() << "FabberRunDataNewimage::No mask, using data for extent" << std::endl;
        const std::string data_fname = GetStringDefault("data", GetStringDefault("data1", ""));
        if (nifti_find(data_fname) == "")
            throw DataNotFound(data_fname + " (File is invalid or does not exist)");
        nifti_read(data_fname, *m_like, vol);
        SetExtent(m_like->nx, m_like->ny, m_like->nz, nullptr);
    }
}

bool FabberRunDataNewimage::LoadVoxelData(const std::string &filename)
{
    if (nifti_find(filename) == "")
        return false;
    Log() << "FabberRunDataNewimage::Loading data from '" + filename << "'" << std::endl;
    NiftiHeader h;
    std::vector<float> vol;
    nifti_read(filename, h, vol);
    dump_info(h, Log());
    if (!m_like)
    {
        /* no SetExtentFromData yet: the first file defines the grid (rundata_newimage.cc:104-111) */
        m_like.reset(new NiftiHeader(h));
        SetExtent(h.nx, h.ny, h.nz, nullptr);
    }
    const int *ext = Extent();
    if (h.nx != ext[0] || h.ny != ext[1] || h.nz != ext[2])
        throw FabberRunDataError("Dimension mismatch between " + filename + " and the mask / main data");
    Log() << "FabberRunDataNewimage::Applying mask to data..." << std::endl;
    /* volume4D::matrix(mask): one row per volume, one column per in-mask voxel, x fastest */
    SetVoxelDataArray(filename, h.nt, vol.data());
    const VoxelData &vd = GetVoxelData(filename);
    double sum = 0;
    for (size_t i = 0; i < (size_t)vd.rows * vd.cols; i++)
        sum += vd.f[i];
    Log() << "FabberRunDataNewimage::GetVoxelData: " << filename << " mean value=" << sum / ((double)vd.rows * vd.cols)
          << std::endl;
    return true;
}

void FabberRunDataNewimage::SaveVoxelData(const std::string &key, VoxelDataType type)
{
    Log() << "FabberRunDataNewimage::Saving to nifti: " << key << std::endl;
    const VoxelData &vd = GetVoxelData(key);
    const int *ext = Extent();
    if (!m_like)
        m_like.reset(new NiftiHeader());
    m_like->nx = ext[0];
    m_like->ny = ext[1];
    m_like->nz = ext[2];
    std::vector<float> vol((size_t)vd.rows * ext[0] * ext[1] * ext[2]);
    GetVoxelDataArray(key, vol.data()); /* zeros outside the mask, like setmatrix(data, mask) */
    const std::string path = (key[0] == '/') ? key : GetOutputDir() + "/" + key;
    nifti_write(path, *m_like, vd.rows, type == VDT_MVN ? NIFTI_INTENT_SYMMATRIX_CODE : NIFTI_INTENT_NONE_CODE, vol.data());
    ClearVoxelData(key); /* on disk now: the command line tool never reads it back */
}
} // namespace fabber_b200
