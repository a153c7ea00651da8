/*
 * fabber_capi.h - the OUTER drop-in boundary: the reference's public C API, re-exported by
 * libfabbercore_b200.so with the same names, argument meaning and error behaviour, so that existing
 * bindings (the reference's py/fabber.py ctypes wrapper, Quantiphyse, pyfab) can load this library in
 * place of libfabbercore_shared.so for the VB path.
 *
 * Each entry point replaces the reference function of the same name, fabber_capi.h:40-279 /
 * fabber_capi.cc:45-623. Conventions kept (fabber_capi.cc:32-43): return 0 on success, <0 on error
 * (FABBER_ERR_FATAL, or -1 for "data not found" / "buffer too small"); the message is copied into the
 * caller-owned err_buf (>= FABBER_ERR_MAXC bytes, optional except in fabber_dorun); no exception
 * crosses the ABI; the caller owns every buffer; volumes are float arrays, x fastest, t slowest
 * (rundata_array.cc:100-133), masks are int arrays.
 *
 * What runs behind fabber_dorun here: --method=vb and --method=spatialvb (Vb::DoCalculations on the GPU
 * through include/fabber_cuda.h) for the models with a device Evaluate hook (linear, poly, exp), white
 * and AR(1) noise. Anything else (nlls, dynamically loaded CPU models) returns FABBER_ERR_FATAL with a
 * message: there is no CPU fallback.
 */
#ifndef FABBER_CAPI_H
#define FABBER_CAPI_H

#define FABBER_ERR_MAXC 255
#define FABBER_ERR_FATAL -255
#define FABBER_ERR_NEWMAT -254

#ifdef __cplusplus
extern "C" {
#endif

/* fabber_capi.h:40  - new run context (one live context at a time, as in the reference) */
void *fabber_new(char *err_buf);
/* fabber_capi.h:52  - dynamically loaded CPU models have no device hook: always an error here */
int fabber_load_models(void *fab, const char *libpath, char *err_buf);
/* fabber_capi.h:68  - extent + mask (mask != 0 inside), defines the voxel order x fastest */
int fabber_set_extent(void *fab, unsigned int nx, unsigned int ny, unsigned int nz, const int *mask, char *err_buf);
/* fabber_capi.h:76 */
void fabber_destroy(void *fab);
/* fabber_capi.h:90  - boolean options are set with an empty value */
int fabber_set_opt(void *fab, const char *key, const char *value, char *err_buf);
/* fabber_capi.h:108 - data_size volumes of nx*ny*nz floats, copied */
int fabber_set_data(void *fab, const char *name, unsigned int data_size, const float *data, char *err_buf);
/* fabber_capi.h:121 - number of volumes of a named output, -1 if not found */
int fabber_get_data_size(void *fab, const char *name, char *err_buf);
/* fabber_capi.h:137 - copies size * nx*ny*nz floats, zeros outside the mask */
int fabber_get_data(void *fab, const char *name, float *data_buf, char *err_buf);
/* fabber_capi.h:155 - run; progress_cb(voxel, nvoxels) is called on the calling thread */
int fabber_dorun(void *fab, unsigned int log_bufsize, char *log_buf, char *err_buf, void (*progress_cb)(int, int));
/* fabber_capi.h:178-279 - self description (newline / tab separated text) */
int fabber_get_options(
    void *fab, const char *key, const char *value, unsigned int out_bufsize, char *out_buf, char *err_buf);
int fabber_get_models(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf);
int fabber_get_methods(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf);
int fabber_get_model_params(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf);
int fabber_get_model_param_descs(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf);
int fabber_get_model_outputs(void *fab, unsigned int out_bufsize, char *out_buf, char *err_buf);
int fabber_model_evaluate(void *fab, unsigned int n_params, float *params, unsigned int n_ts, float *indata,
    float *output, char *err_buf);
int fabber_model_evaluate_output(void *fab, unsigned int n_params, float *params, unsigned int n_ts, float *indata,
    const char *output_name, float *output, char *err_buf);

#ifdef __cplusplus
}
#endif
#endif
