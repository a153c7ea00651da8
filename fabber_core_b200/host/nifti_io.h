/*
 * nifti_io.h - the small part of NIfTI-1 I/O the command line front end needs, on zlib only.
 *
 * Stands where the reference uses FSL's NEWIMAGE (rundata_newimage.cc: read_volume / read_volume4D /
 * save_volume4D, volume4D::matrix(mask) / setmatrix). What is reproduced:
 *   - single-file .nii / .nii.gz and .hdr/.img(.gz) pairs, either byte order, every scalar datatype,
 *     scl_slope / scl_inter, returned as float32 in file order x fastest, then y, z, t;
 *   - NEWIMAGE's storage convention: volumes are held in RADIOLOGICAL voxel order, so a file whose
 *     sform (or, failing that, qform) has a positive determinant is mirrored in x on the way in and
 *     mirrored back on the way out. The voxel order is not cosmetic here: spatial VB sweeps the voxels
 *     sequentially (spatialvb.cc:428-437), so the result depends on it;
 *   - outputs are float32, geometry copied from the header of the mask (or of the first data file),
 *     intent code SYMMATRIX for MVNs, cal_min / cal_max set to the data range
 *     (rundata_newimage.cc:140-183), written as FSLOUTPUTTYPE says (default NIFTI_GZ).
 */
#pragma once
#include <string>
#include <vector>

namespace fabber_b200
{
const int NIFTI_INTENT_NONE_CODE = 0;
const int NIFTI_INTENT_SYMMATRIX_CODE = 1005;

struct NiftiHeader
{
    unsigned char raw[348]; /* native byte order */
    int nx, ny, nz, nt;
    float dx, dy, dz, dt;
    int intent_code;
    float intent_p[3];
    bool flip_x; /* file is in neurological order: mirrored in x between file and memory */
    NiftiHeader();
};

/* does `name`, `name`.nii.gz, `name`.nii, `name`.hdr(.gz) exist? Returns the header file's path or "" (fsl_imageexists) */
std::string nifti_find(const std::string &name);

/* read a 3-D or 4-D image; data comes back [t][z][y][x] in memory (radiological) order. Throws FabberRunDataError. */
void nifti_read(const std::string &name, NiftiHeader &hdr, std::vector<float> &data);

/* write float32 [nt][z][y][x] with the geometry of `like`; `name` without extension gets the FSLOUTPUTTYPE one */
std::string nifti_write(const std::string &name, const NiftiHeader &like, int nt, int intent_code, const float *data);
} // namespace fabber_b200
